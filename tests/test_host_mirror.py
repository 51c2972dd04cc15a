"""CPU-side checks of the host mirror: parameter names, the torch per-frame setup network, library exports."""
import ctypes
import os
import re

import pytest
import torch

from nerf_loc_b200 import _lib, params, synthetic as syn
from nerf_loc_b200.config import default_args
from nerf_loc_b200.conditional_nerf import ConditionalNeRF
from oracle import ref_harness as rh
from tests.common import relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nerfloc_b200.h")).read()
    declared = set(re.findall(r"\b(nlb_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("nlb_scene")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)  # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nlb_version() == 100
    assert lib.nlb_render_param_count() == len(params.conditional_nerf_shapes(64))


def test_mirror_state_dict_covers_the_inventory():
    for S in (16, 64):
        m = ConditionalNeRF(default_args(S))
        sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        inv = {k: tuple(v) for k, v in params.conditional_nerf_shapes(S).items()}
        assert {k: v for k, v in sd.items() if "depth_fusion" not in k} == inv
        assert list(k for k in sd if "depth_fusion" not in k) == list(inv)


def test_no_cpu_fallback():
    m = ConditionalNeRF(default_args(16))
    with pytest.raises(RuntimeError):
        m.packed_weights()  # CPU tensors are refused, nothing is computed on the host


@pytest.mark.skipif(not rh.available(), reason="/root/reference not present")
def test_mirror_matches_reference_names_and_depth_fusion():
    R = rh.load()
    torch.manual_seed(0)
    ref = R.ConditionalNeRF(rh.default_args(16)).eval()
    mine = ConditionalNeRF(default_args(16)).eval()
    ref_sd = ref.state_dict()
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    mine.load_state_dict(ref_sd)  # a reference checkpoint loads unchanged
    sc = syn.make_scene(64, 96, 3, seed=5)
    args = (sc["topk_images"], sc["feat_fine_src"].permute(0, 3, 1, 2), sc["topk_depths"], sc["topk_Ks"],
            sc["topk_poses"], sc["depth_range"][0])
    with torch.no_grad():
        a = ref.multiview_aggregator.depth_fusion(*args)
        b = mine.multiview_aggregator.depth_fusion(*args)
    assert a.shape == b.shape == (3, 32, 16, 24)
    assert relerr(b, a) < 2e-5
    # support-point construction (host plumbing)
    data = {k: sc[k] for k in ("topk_images", "topk_depths", "topk_poses", "topk_Ks", "feat_fine_src")}
    with torch.no_grad():
        r4 = ref.backproject_support_frame(sc["topk_images"], sc["feat_fine_src"], sc["topk_depths"], sc["topk_Ks"],
                                           sc["topk_poses"], stride=4)
        m4 = mine.backproject_support_frame(sc["topk_images"], sc["feat_fine_src"], sc["topk_depths"], sc["topk_Ks"],
                                            sc["topk_poses"], stride=4)
    for x, y in zip(r4, m4):
        assert torch.equal(x, y)


@pytest.mark.skipif(not rh.available(), reason="/root/reference not present")
def test_matcher_mirror_loads_reference_state_dict_and_transformer_matches():
    from nerf_loc_b200.matcher import Matcher, PositionEmbeddingSine
    from oracle import matcher_oracle as MO
    R = rh.load()
    ref = R.Matcher(rh.default_args(), 192, 192, 192).eval()
    mine = Matcher(default_args(), 192, 192, 192).eval()
    mine.load_state_dict(ref.state_dict())
    sd = {k: v for k, v in ref.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    a, pa = torch.randn(1, 20, 192, generator=g), torch.randn(1, 20, 192, generator=g)
    b, pb = torch.randn(1, 30, 192, generator=g), torch.randn(1, 30, 192, generator=g)
    with torch.no_grad():
        r0, r1 = ref.coarse_transformer(a, pa, b, pb)
        m0, m1 = mine.coarse_transformer(a, pa, b, pb)
        o0, o1 = MO.self_cross_transformer(sd, "coarse_transformer", a, pa, b, pb)
    assert relerr(m0, r0) < 2e-5 and relerr(m1, r1) < 2e-5 and relerr(o0, r0) < 2e-5 and relerr(o1, r1) < 2e-5
    pe = PositionEmbeddingSine(96, normalize=True)(torch.zeros(2, 7, 7))
    assert torch.equal(pe, R.PositionEmbeddingSine(96, normalize=True, sine_type="lin_sine")(torch.zeros(2, 7, 7)))


@pytest.mark.skipif(not rh.available(), reason="/root/reference not present")
def test_top_level_front_end_matches_reference_modules():
    """Backbone2D / appearance layers of nerf_loc_b200.nerf_pose_estimator against COTR/backbone2d.py and
    appearance_embedding.py: same parameter names, same outputs under the same weights."""
    import importlib
    import types
    rh.load()
    from nerf_loc_b200 import nerf_pose_estimator as NPE
    bb = importlib.import_module("nerf_loc.models.COTR.backbone2d")
    torch.manual_seed(0)
    ref = bb.Backbone(['conv1', 'layer1', 'layer2'], train_backbone=True, use_fpn=True, fpn_dim=192).eval()
    mine = NPE.Backbone2D(fpn_dim=192).eval()
    assert set(mine.state_dict()) == set(ref.state_dict())
    mine.load_state_dict(ref.state_dict())
    x = torch.rand(2, 3, 64, 96)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    for k in ('conv1', 'layer1', 'layer2'):
        assert a[k].shape == b[k].shape and torch.allclose(a[k], b[k], atol=1e-5), k
    assert mine.layer_to_channels['layer1'] == 192 and mine.layer_to_stride['layer2'] == 8
    ap = importlib.import_module("nerf_loc.models.appearance_embedding")
    args = types.SimpleNamespace(appearance_emb_dim=128)
    r_ad, m_ad = ap.AppearanceAdaptLayer(args, 192), NPE.AppearanceAdaptLayer(args, 192)
    m_ad.load_state_dict(r_ad.state_dict())
    feats = {'conv1': torch.randn(3, 64, 8, 12)}
    e_r, e_m = ap.AppearanceEmbedding(args)(None, feats), NPE.AppearanceEmbedding(args)(None, feats)
    assert torch.allclose(e_r, e_m, atol=1e-6)
    xf = torch.randn(3, 4, 5, 192)
    assert torch.allclose(r_ad(xf, e_r, e_r[:1]), m_ad(xf, e_m, e_m[:1]), atol=1e-6)


def test_knn_gather_masks_short_clouds_and_knn_points_validates_arguments():
    """knn_utils.py:171-222 (gather with `lengths`) and the argument checks of knn_points (no search: the search needs a GPU)."""
    from nerf_loc_b200.knn import knn_gather, knn_points
    x = torch.arange(2 * 5 * 2, dtype=torch.float32).reshape(2, 5, 2)
    idx = torch.tensor([[[0, 1, 4]], [[2, 0, 0]]])
    out = knn_gather(x, idx, torch.tensor([5, 1]))
    assert torch.equal(out[0, 0], x[0, [0, 1, 4]])
    assert torch.equal(out[1, 0, 0], x[1, 2]) and float(out[1, 0, 1:].abs().sum()) == 0.0
    assert torch.equal(knn_gather(x, idx), torch.stack([x[0, [0, 1, 4]][None], x[1, [2, 0, 0]][None]]))
    for bad in (dict(K=0), dict(K=17), dict(lengths1=torch.tensor([3])), dict(lengths2=torch.tensor([-1]))):
        with pytest.raises(ValueError):
            knn_points(torch.zeros(1, 2, 3), torch.zeros(1, 2, 3), **bad)
    with pytest.raises(ValueError):
        knn_points(torch.zeros(1, 2, 2), torch.zeros(1, 2, 2))
    with pytest.raises(ValueError):
        knn_points(torch.zeros(2, 2, 3), torch.zeros(1, 2, 3))
