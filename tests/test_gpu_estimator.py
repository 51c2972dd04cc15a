"""-m gpu: the top-level drop-in surface NerfPoseEstimator.forward(batch) (nerf_pose_estimator.py:239-405), inference mode,
on a synthetic frame: every stage of the hot path runs through the CUDA library and the output dict has the reference's keys
and shapes.  (Weights are random, so the pose itself is meaningless here; PnP accuracy is covered by test_gpu_pnp.py.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(H=64, W=96, V=3, seed=5):
    from nerf_loc_b200 import synthetic as syn
    sc = syn.make_scene(H, W, V, seed=seed)
    g = torch.Generator().manual_seed(seed)
    dr = sc["depth_range"][0]
    b = {"image": torch.rand(1, 3, H, W, generator=g), "pose": sc["pose"][None], "K": sc["K"][None],
         "depth": sc["topk_depths"][:1].clone(), "near": dr[:1].clone(), "far": dr[1:].clone(),
         "topk_images": sc["topk_images"][None], "topk_depths": sc["topk_depths"][None], "topk_poses": sc["topk_poses"][None],
         "topk_Ks": sc["topk_Ks"][None], "points3d": torch.cat([torch.rand(1, 50, 3, generator=g) * 2, torch.rand(1, 50, 3, generator=g) * 255], -1)}
    b = {k: v.cuda() for k, v in b.items()}
    b["scene"], b["filename"] = ["synthetic"], ["frame0"]
    return b, sc


def test_forward_runs_the_whole_chain():
    from nerf_loc_b200.config import default_args
    from nerf_loc_b200.nerf_pose_estimator import NerfPoseEstimator
    torch.manual_seed(0)
    np.random.seed(0)
    args = default_args(16)
    args.matching.coarse_num_3d_keypoints = args.matching.fine_num_3d_keypoints = 128
    m = NerfPoseEstimator(args).eval().cuda()
    batch, sc = _batch()
    with pytest.raises(NotImplementedError):
        m.train()(batch)
    m.eval()
    batch["render_image"] = True
    out = m(batch)
    H, W = 64, 96
    n3, mc = out["score_matrix"].shape
    assert mc == (H // 8) * (W // 8) and 0 < n3 <= 128
    assert np.asarray(out["T"]).shape == (4, 4)
    assert out["pairs_gt"].shape[0] == 2
    assert tuple(out["rendered_image"].shape) == (H, W, 3) and tuple(out["rendered_depth"].shape) == (H, W, 1)  # render_image reshapes every key to [H, W, -1] (model.py:636)
    assert tuple(out["rendered_feat"].shape) == (H, W, 192) and tuple(out["rendered_feat_gt"].shape) == (H, W, 192)
    assert torch.isfinite(out["rendered_image"]).all() and torch.isfinite(out["score_matrix"]).all()
    assert len(out["pairs"]) == 2
    # per-frame caches were rebuilt for this frame (reference :289-290)
    assert m.model_3d.support_neural_points is not None


def test_estimate_recovers_pose_from_planted_matches():
    """estimate_pose on matches taken from the known geometry of the frame: the chain matcher output -> PnP -> c2w."""
    from nerf_loc_b200.config import default_args
    from nerf_loc_b200.nerf_pose_estimator import NerfPoseEstimator, camera_project
    m = NerfPoseEstimator(default_args(16)).eval().cuda()
    batch, sc = _batch()
    g = torch.Generator().manual_seed(1)
    pts = torch.rand(300, 3, generator=g) * torch.tensor([2.0, 1.5, 1.0]) + torch.tensor([-1.0, -0.75, 1.5])
    pose, K = sc["pose"], sc["K"]
    cam = (pose.inverse() @ torch.cat([pts, torch.ones(300, 1)], 1).t()).t()[:, :3]
    u, v, z = camera_project(cam, K)
    uv = torch.stack([u, v], 1) + 0.3 * torch.randn(300, 2, generator=g)
    T, inl = m.estimate_pose(uv.cuda(), pts.cuda(), K, 96, 64, ransac_thresh=8)
    assert np.abs(T - pose.numpy()).max() < 2e-2 and inl.sum() > 280
