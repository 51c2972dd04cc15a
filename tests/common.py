"""Shared builders for the parity tests (inputs are regenerated from seeds; see nerf_loc_b200/synthetic.py)."""
import os

import numpy as np
import torch

from nerf_loc_b200 import params, synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror oracle/make_golden.py RENDER_CASES
RENDER_CASES = {
    "render_s16": (16, 64, 96, 3, 24, 1234, 1234),
    "render_s64": (64, 64, 96, 4, 12, 4321, 77),
    "render_s192": (192, 64, 96, 3, 6, 555, 99),   # long rays (BASELINE configs[3]: 192 samples per ray)
    "render_v8_s128": (128, 64, 96, 8, 4, 2024, 8),   # the metric's configuration: 8 views, 128 samples per ray
    "render_v16_s32": (32, 64, 96, 16, 6, 777, 21),   # the upper end of the supported view count
}


# every case pins the oracle (CPU suite) AND the CUDA path (-m gpu tests iterate over RENDER_CASES)
ORACLE_CASES = dict(RENDER_CASES)


def relerr(a, b):
    """max-norm relative error (SURVEY.md section 8a 'Parity classes')."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def render_inputs(name):
    S, H, W, V, R, wseed, sseed = ORACLE_CASES[name]
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), wseed)
    sc = syn.make_scene(H, W, V, seed=sseed)
    px = syn.random_pixels(H, W, R)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    return S, sd, sc, scene, ro, rd


def golden(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}
