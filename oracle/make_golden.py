"""Generates tests/golden/*.npz by running the REAL reference (imported from /root/reference through
oracle/ref_harness.py) on seeded synthetic inputs.  Run in the build container only:

    python oracle/make_golden.py

Inputs are NOT stored: they are regenerated from the seeds by nerf_loc_b200/synthetic.py (same torch
version on the GPU box).  Only reference OUTPUTS are stored, so the fixtures stay small.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from nerf_loc_b200 import params, synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

RENDER_CASES = {
    # name: (S, H, W, V, n_rays, weight seed, scene seed)
    "render_s16": (16, 64, 96, 3, 24, 1234, 1234),
    "render_s64": (64, 64, 96, 4, 12, 4321, 77),
    "render_s192": (192, 64, 96, 3, 6, 555, 99),   # long rays (BASELINE configs[3]: 192 samples per ray)
    # the metric's own configuration (8 reference views, 128 samples per ray) on a small image
    "render_v8_s128": (128, 64, 96, 8, 4, 2024, 8),
    "render_v16_s32": (32, 64, 96, 16, 6, 777, 21),   # the upper end of the supported view count
}


def render_case(name):
    S, H, W, V, R_, wseed, sseed = RENDER_CASES[name]
    R = rh.load()
    model = R.ConditionalNeRF(rh.default_args(S)).eval()
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), wseed)
    model.load_state_dict(sd, strict=False)
    sc = syn.make_scene(H, W, V, seed=sseed)
    data = {k: sc[k] for k in ("K", "pose", "H", "W", "depth_range", "topk_images", "topk_depths", "topk_poses",
                               "topk_Ks", "feat_fine_src", "feat_coarse_src", "stride_fine", "stride_coarse",
                               "embedding_a")}
    data["scene"], data["filename"] = "synthetic", name
    # DepthFusionNet is per-frame setup: its output is an input of the hot path
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"]
    px = syn.random_pixels(H, W, R_)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    rays = {"rays_o": ro, "rays_d": rd, "depth_range": sc["depth_range"][0], "pixel_coordinates": px,
            "K": sc["K"], "pose": sc["pose"], "H": H, "W": W}
    with torch.no_grad():
        out = model.render_rays(data, rays)
        sup = model.support_neural_points
        z = model.sample_depths(S, *sc["depth_range"][0])
        xyz = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
        q = model.query(data, xyz, data["feat_fine_src"].permute(0, 3, 1, 2), sup["fine"], direction=None, K=8)
        knn = R.knn_points(xyz[None], sup["fine"]["xyz"][None], K=8)
        pts = sup["coarse"]["xyz"][::7][:40] + 0.01
        dc, _, _ = model.query_coarse(data, points=pts)
        df, _, _ = model.query_fine(data, pts)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        rgb=out["rgb"].numpy(), depth=out["depth"].numpy(), weights=out["weights"].numpy(),
        mask=out["mask"].numpy(), depth_uncertainty=out["depth_uncertainty"].numpy(), feat=out["feat"].numpy(),
        conf_fine=sup["fine"]["confidence"].numpy(), kp_coarse=sup["coarse"]["keypoint_score"].numpy(),
        n_fine=np.int64(sup["fine"]["xyz"].shape[0]),
        q_feature_agg=q["feature_agg"].numpy(), q_weights=q["weights"].numpy(),
        q_vis=q["multiview_visibility"].numpy(), knn_idx=knn.idx[0].numpy(), knn_d2=knn.dists[0].numpy(),
        desc_coarse=dc.numpy(), desc_fine=df.numpy())
    print(name, "written; mask true:", int(out["mask"].sum()), "/", R_)


def matcher_inputs(seed=5, N3=96, hc=6, wc=8):
    """Shared by make_golden and the tests (imported from here by tests only)."""
    from oracle import matcher_oracle as MO
    from oracle import nerfloc_oracle as O
    g = torch.Generator().manual_seed(seed)
    Mc = hc * wc
    d2 = torch.randn(Mc, 192, generator=g)
    d3 = torch.randn(N3, 192, generator=g)
    perm = torch.randperm(Mc, generator=g)[:32]
    d3[:32] = d2[perm] + 0.05 * torch.randn(32, 192, generator=g)
    kps3d = torch.rand(N3, 3, generator=g) * 2
    pe3 = O.positional_encoding(kps3d, 32, include_input=False)
    pe2 = MO.pos_embed_2d(torch.zeros(1, hc, wc))[0].reshape(Mc, -1)
    gy, gx = torch.meshgrid(torch.arange(hc), torch.arange(wc), indexing="ij")
    kps2d = torch.stack([gx, gy], -1).view(-1, 2).float()
    return dict(desc_3d=d3, pos_emd_3d=pe3, desc_2d_coarse=d2, pos_emd_2d=pe2, kps3d=kps3d, kps2d=kps2d,
                feat_fine=torch.randn(1, hc * 2, wc * 2, 192, generator=g),
                feat_coarse=torch.randn(1, hc, wc, 192, generator=g),
                desc_3d_fine=torch.randn(N3, 192, generator=g), stride_coarse=8, stride_fine=4)


def matcher_case():
    R = rh.load()
    m = R.Matcher(rh.default_args(), 192, 192, 192).eval()
    sd = syn.synthetic_state_dict(params.matcher_shapes(), 99)
    m.load_state_dict(sd)
    data = matcher_inputs()
    # reference-side check of the two positional encodings the inputs use
    pe_fn, _ = R.get_embedder(32, 0, include_input=False)
    assert torch.equal(pe_fn(data["kps3d"]), data["pos_emd_3d"])
    ref_pe2 = R.PositionEmbeddingSine(96, normalize=True, sine_type="lin_sine")(torch.zeros(1, 6, 8))[0].reshape(48, -1)
    assert torch.equal(ref_pe2, data["pos_emd_2d"])
    with torch.no_grad():
        out = m(dict(data))
    np.savez_compressed(os.path.join(GOLD, "matcher_small.npz"),
                        score_matrix=out["score_matrix"].numpy(), i_ids=out["i_ids"].numpy(),
                        j_ids=out["j_ids"].numpy(), expec_f=out["expec_f"].numpy(),
                        mkps2d_f=out["mkps2d_f"].numpy(), mkps2d_c=out["mkps2d_c"].numpy())
    print("matcher_small written; matches:", len(out["i_ids"]))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    for n in RENDER_CASES:
        if not only or n in only:
            render_case(n)
    if not only or "matcher_small" in only:
        matcher_case()
