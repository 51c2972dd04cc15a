"""-m gpu: device PnP-RANSAC (nlb_pnp_ransac through nerf_loc_b200.pnp) against the known pose and the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import pnp_oracle as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,seed", [(2048, 0), (512, 7), (64, 3)])
def test_pose_recovered_like_the_oracle(M, seed):
    from nerf_loc_b200 import pnp
    p2d, p3d, cam, Rgt, tgt, gt_inl = P.synthetic_correspondences(M=M, seed=seed)
    ret = pnp.absolute_pose_estimation(torch.from_numpy(p2d).cuda(), torch.from_numpy(p3d).cuda(), cam, 8.0, iters=1024, seed=seed)
    assert ret["success"]
    rot, pos = P.pose_error(ret["R"], ret["tvec"], Rgt, tgt)
    ref = P.absolute_pose_ransac(p2d, p3d, cam, 8.0, iters=128, seed=seed)
    rot_o, pos_o = P.pose_error(ref["R"], ref["t"], Rgt, tgt)
    print("gpu", rot, pos, "oracle", rot_o, pos_o, ret["num_inliers"], ref["num_inliers"])
    assert rot < 0.1 and pos < 5e-3
    # both converge to the same least-squares pose on (almost) the same inlier set
    inl = ret["inliers"].cpu().numpy()
    assert (inl != ref["inliers"]).sum() <= max(2, 0.01 * M)
    assert abs(rot - rot_o) < 0.02 and abs(pos - pos_o) < 1e-3


def test_estimate_pose_signature_and_failure():
    from nerf_loc_b200 import pnp
    p2d, p3d, cam, Rgt, tgt, _ = P.synthetic_correspondences(M=300, seed=11)
    K = torch.tensor([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1.0]])
    out = pnp.estimate_pose(torch.from_numpy(p2d).cuda(), torch.from_numpy(p3d).cuda(), K, 640, 480, ransac_thresh=8)
    c2w, inl = out
    assert c2w.shape == (4, 4) and inl.shape == (300,) and inl.dtype == bool
    w2c = np.linalg.inv(c2w)
    assert P.pose_error(w2c[:3, :3], w2c[:3, 3], Rgt, tgt)[0] < 0.2
    # too few matches -> None, like a failed pycolmap call (nerf_pose_estimator.py:576-577)
    assert pnp.estimate_pose(torch.zeros(3, 2).cuda(), torch.zeros(3, 3).cuda(), K, 640, 480) is None
    # pure noise: either no pose or a pose with (almost) no support
    g = torch.Generator().manual_seed(0)
    r = pnp.absolute_pose_estimation((torch.rand(200, 2, generator=g) * 400).cuda(), torch.randn(200, 3, generator=g).cuda(), cam, 2.0)
    assert (not r["success"]) or r["num_inliers"] < 20


def test_deterministic_for_a_seed():
    from nerf_loc_b200 import pnp
    p2d, p3d, cam, *_ = P.synthetic_correspondences(M=400, seed=2)
    a = pnp.absolute_pose_estimation(torch.from_numpy(p2d).cuda(), torch.from_numpy(p3d).cuda(), cam, 8.0, seed=5)
    b = pnp.absolute_pose_estimation(torch.from_numpy(p2d).cuda(), torch.from_numpy(p3d).cuda(), cam, 8.0, seed=5)
    assert np.array_equal(a["R"], b["R"]) and np.array_equal(a["tvec"], b["tvec"])
