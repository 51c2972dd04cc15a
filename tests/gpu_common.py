"""Builders shared by the -m gpu parity tests: the CUDA model / scene on the device and the oracle's view of the same
inputs on the CPU."""
import torch

from nerf_loc_b200 import params, synthetic as syn
from nerf_loc_b200.conditional_nerf import ConditionalNeRF
from nerf_loc_b200.config import default_args
from oracle import nerfloc_oracle as O

DATA_KEYS = ("K", "pose", "H", "W", "depth_range", "topk_images", "topk_depths", "topk_poses", "topk_Ks",
             "feat_fine_src", "feat_coarse_src", "stride_fine", "stride_coarse", "embedding_a")


def cuda_model(S, wseed):
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), wseed)
    m = ConditionalNeRF(default_args(S)).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert all("depth_fusion" in k for k in missing.missing_keys) and not missing.unexpected_keys
    return m.cuda(), sd


def cuda_data(sc):
    d = {}
    for k in DATA_KEYS:
        v = sc[k]
        d[k] = v.cuda() if torch.is_tensor(v) else v
    d["scene"], d["filename"] = "synthetic", "frame"
    return d


def oracle_scene(sc):
    return dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                depth_range=sc["depth_range"][0])


def setup_frame(model, sc):
    """Resets the per-frame caches like nerf_pose_estimator.py:289-290 and injects the synthetic visibility maps
    (DepthFusionNet is per-frame setup, outside the parity scope of these tests)."""
    data = cuda_data(sc)
    model.support_neural_points = None
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"].cuda()
    return data


def oracle_support(sd, sc):
    with torch.no_grad():
        return O.build_support_neural_points(sd, oracle_scene(sc), sc["feat_coarse_src"], sc["feat_fine_src"],
                                             sc["topk_depths"])
