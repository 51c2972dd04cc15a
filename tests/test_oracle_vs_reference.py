"""Pins the oracle against the reference ITSELF, imported from /root/reference (build container only)."""
import pytest
import torch

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not present (GPU box)")

from oracle import matcher_oracle as MO  # noqa: E402
from oracle import nerfloc_oracle as O  # noqa: E402
from oracle.make_golden import matcher_inputs  # noqa: E402
from nerf_loc_b200 import params, synthetic as syn  # noqa: E402
from tests.common import relerr  # noqa: E402


def test_parameter_inventory_matches_reference_state_dict():
    R = rh.load()
    for S in (16, 64):
        ref = {k: tuple(v.shape) for k, v in R.ConditionalNeRF(rh.default_args(S)).state_dict().items()
               if "depth_fusion" not in k}
        mine = {k: tuple(v) for k, v in params.conditional_nerf_shapes(S).items()}
        assert ref == mine
    ref = {k: tuple(v.shape) for k, v in R.Matcher(rh.default_args(), 192, 192, 192).state_dict().items()}
    assert ref == {k: tuple(v) for k, v in params.matcher_shapes().items()}


@pytest.mark.parametrize("S,H,W,V", [(8, 32, 64, 2), (32, 64, 64, 5)])
def test_render_rays_matches_reference(S, H, W, V):
    R = rh.load()
    model = R.ConditionalNeRF(rh.default_args(S)).eval()
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), 11)
    model.load_state_dict(sd, strict=False)
    sc = syn.make_scene(H, W, V, seed=3)
    data = {k: sc[k] for k in ("K", "pose", "H", "W", "depth_range", "topk_images", "topk_depths", "topk_poses",
                               "topk_Ks", "feat_fine_src", "feat_coarse_src", "stride_fine", "stride_coarse",
                               "embedding_a")}
    data["scene"], data["filename"] = "s", "f"
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"]
    px = syn.random_pixels(H, W, 16)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    rays = {"rays_o": ro, "rays_d": rd, "depth_range": sc["depth_range"][0], "pixel_coordinates": px,
            "K": sc["K"], "pose": sc["pose"], "H": H, "W": W}
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    with torch.no_grad():
        ref = model.render_rays(data, rays)
        sup = O.build_support_neural_points(sd, scene, sc["feat_coarse_src"], sc["feat_fine_src"], sc["topk_depths"])
        for lv in ("coarse", "fine"):
            for k, v in sup[lv].items():
                assert relerr(v, model.support_neural_points[lv][k]) < 2e-5, (lv, k)
        out = O.render_rays(sd, scene, sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S)
    for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat"):
        assert relerr(out[k], ref[k]) < 2e-5, k
    assert torch.equal(out["mask"], ref["mask"])
    # the ray generator itself
    rr = model.points_2d_to_rays(px, H, W, sc["K"], sc["pose"])
    assert torch.equal(rr["rays_d"], rd) and torch.equal(rr["rays_o"], ro)


def test_matcher_matches_reference():
    R = rh.load()
    m = R.Matcher(rh.default_args(), 192, 192, 192).eval()
    sd = syn.synthetic_state_dict(params.matcher_shapes(), 7)
    m.load_state_dict(sd)
    data = matcher_inputs(seed=21, N3=64, hc=5, wc=7)
    with torch.no_grad():
        ref = m(dict(data))
        out = MO.matcher_forward(sd, data)
    assert relerr(out["score_matrix"], ref["score_matrix"]) < 2e-5
    assert torch.equal(out["i_ids"], ref["i_ids"]) and torch.equal(out["j_ids"], ref["j_ids"])
    assert relerr(out["expec_f"], ref["expec_f"]) < 2e-5
    assert relerr(out["mkps2d_f"], ref["mkps2d_f"]) < 2e-5


def test_hierarchical_sampling_matches_reference():
    """N_importance > 0 (model.py:486-496): coarse weights from the NeuRay decoder, inverse-CDF sampling, sort.  The reference
    draws u with torch.rand; both sides run under the same CPU generator state, and the searchsorted indices (the 'sample
    indices' of SURVEY.md section 8a) must agree exactly."""
    R = rh.load()
    S, NI, H, W, V = 16, 16, 32, 64, 3
    model = R.ConditionalNeRF(rh.default_args(S, NI)).eval()
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S + NI), 5)
    model.load_state_dict(sd, strict=False)
    sc = syn.make_scene(H, W, V, seed=8)
    data = {k: sc[k] for k in ("K", "pose", "H", "W", "depth_range", "topk_images", "topk_depths", "topk_poses",
                               "topk_Ks", "feat_fine_src", "feat_coarse_src", "stride_fine", "stride_coarse",
                               "embedding_a")}
    data["scene"], data["filename"] = "s", "f"
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"]
    px = syn.random_pixels(H, W, 12)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    rays = {"rays_o": ro, "rays_d": rd, "depth_range": sc["depth_range"][0], "pixel_coordinates": px.float(),
            "K": sc["K"], "pose": sc["pose"], "H": H, "W": W}
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    with torch.no_grad():
        # the pieces
        zc = O.sample_depths(64, *sc["depth_range"][0]).expand(12, 64).contiguous()
        w_ref = model.multiview_aggregator.predict_weights_from_neuray(data, rays, zc)
        w_or = O.predict_weights_from_neuray(sd, "multiview_aggregator", scene, px.float(), sc["K"], sc["pose"], zc)
        assert relerr(w_or, w_ref) < 2e-5
        mid = 0.5 * (zc[:, :-1] + zc[:, 1:])
        torch.manual_seed(99)
        s_ref = R.sample_pdf(mid, w_ref[:, 1:-1], NI)
        torch.manual_seed(99)
        s_or, inds = O.sample_pdf(mid, w_ref[:, 1:-1], NI)
        assert torch.equal(s_or, s_ref)
        assert torch.equal(O.sample_pdf(mid, w_ref[:, 1:-1], NI, det=True)[0], R.sample_pdf(mid, w_ref[:, 1:-1], NI, det=True))
        # the whole render
        torch.manual_seed(7)
        ref = model.render_rays(data, rays)
        torch.manual_seed(7)
        sup = O.build_support_neural_points(sd, scene, sc["feat_coarse_src"], sc["feat_fine_src"], sc["topk_depths"])
        torch.manual_seed(7)
        z, depth_coarse, _ = O.hierarchical_depths(sd, scene, px.float(), sc["K"], sc["pose"], S, NI)
        out = O.render_rays(sd, scene, sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S, z_vals=z)
    assert relerr(depth_coarse, ref["depth_coarse"]) < 2e-5
    for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat"):
        assert relerr(out[k], ref[k]) < 5e-5, (k, relerr(out[k], ref[k]))
    assert torch.equal(out["mask"], ref["mask"])
