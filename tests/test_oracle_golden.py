"""The oracle restatement against the committed golden vectors (outputs of the REAL reference, written by
oracle/make_golden.py).  Runs everywhere (no /root/reference needed)."""
import pytest
import torch

from oracle import matcher_oracle as MO
from oracle import nerfloc_oracle as O
from oracle.make_golden import matcher_inputs
from nerf_loc_b200 import params, synthetic as syn
from tests.common import ORACLE_CASES, golden, relerr, render_inputs

TOL = 2e-5  # fp32 re-association noise between two CPU evaluations of the same algorithm


@pytest.mark.parametrize("name", list(ORACLE_CASES))
def test_render_oracle_vs_golden(name):
    S, sd, sc, scene, ro, rd = render_inputs(name)
    g = golden(name)
    with torch.no_grad():
        sup = O.build_support_neural_points(sd, scene, sc["feat_coarse_src"], sc["feat_fine_src"], sc["topk_depths"])
        assert sup["fine"]["xyz"].shape[0] == int(g["n_fine"])
        assert relerr(sup["fine"]["confidence"], g["conf_fine"]) < TOL
        assert relerr(sup["coarse"]["keypoint_score"], g["kp_coarse"]) < TOL
        fm = sc["feat_fine_src"].permute(0, 3, 1, 2)
        out = O.render_rays(sd, scene, sup["fine"], fm, ro, rd, sc["pose"], S, return_debug=True)
        for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat"):
            assert relerr(out[k], g[k]) < TOL, k
        assert torch.equal(out["mask"], g["mask"])
        assert torch.equal(out["knn_idx"], g["knn_idx"])  # bit-exact integer output
        assert torch.equal(out["knn_d2"], g["knn_d2"])
        z = O.sample_depths(S, *scene["depth_range"])
        xyz = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
        q = O.query(sd, scene, xyz, fm, sup["fine"], direction=None, K=8)
        assert relerr(q["feature_agg"], g["q_feature_agg"]) < TOL
        assert relerr(q["weights"], g["q_weights"]) < TOL
        assert relerr(q["multiview_visibility"], g["q_vis"]) < TOL
        pts = sup["coarse"]["xyz"][::7][:40] + 0.01
        dc = O.query_descriptor(sd, scene, sup["coarse"], sc["feat_coarse_src"].permute(0, 3, 1, 2), pts, "coarse")
        df = O.query_descriptor(sd, scene, sup["fine"], fm, pts, "fine")
        assert relerr(dc, g["desc_coarse"]) < TOL
        assert relerr(df, g["desc_fine"]) < TOL


def test_query_rows_identical_across_k():
    """Property the CUDA path exploits: q is broadcast over K in the neighbour attention
    (conditional_nerf/model.py:413-414), so `feature[n,k,:]` does not depend on k and softmax_K(corr) == 1/K."""
    S, sd, sc, scene, ro, rd = render_inputs("render_s16")
    with torch.no_grad():
        sup = O.build_support_neural_points(sd, scene, sc["feat_coarse_src"], sc["feat_fine_src"], sc["topk_depths"])
        z = O.sample_depths(S, *scene["depth_range"])
        xyz = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
        q = O.query(sd, scene, xyz, sc["feat_fine_src"].permute(0, 3, 1, 2), sup["fine"], K=8)
    f = q["feature"]
    assert float((f - f[:, :1]).abs().max()) < 1e-6


def test_matcher_oracle_vs_golden():
    sd = syn.synthetic_state_dict(params.matcher_shapes(), 99)
    g = golden("matcher_small")
    with torch.no_grad():
        out = MO.matcher_forward(sd, matcher_inputs())
    assert relerr(out["score_matrix"], g["score_matrix"]) < TOL
    assert torch.equal(out["i_ids"], g["i_ids"]) and torch.equal(out["j_ids"], g["j_ids"])
    for k in ("expec_f", "mkps2d_f", "mkps2d_c"):
        assert relerr(out[k], g[k]) < TOL, k
