"""-m gpu parity tests of the render path: CUDA (through the C ABI) vs the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star): bit-exact KNN indices / masks; <= 1e-4 max-norm relative on every float output.
"""
import pytest
import torch

from nerf_loc_b200 import synthetic as syn
from nerf_loc_b200.knn import KnnIndex, knn_points
from oracle import knn_oracle as KO
from oracle import nerfloc_oracle as O
from tests.common import RENDER_CASES, golden, relerr, render_inputs
from tests.gpu_common import cuda_model, oracle_scene, oracle_support, setup_frame

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _cloud(n, seed, lattice=False, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    p = (torch.rand(n, 3, generator=g) * 2 - 1) * scale
    return torch.round(p * 4) / 4 if lattice else p


@pytest.mark.parametrize("K", [1, 8])
@pytest.mark.parametrize("lattice", [False, True])
@pytest.mark.parametrize("M", [3, 100, 20000])
def test_knn_bit_exact(K, lattice, M):
    q, s = _cloud(3000, 1, lattice, 1.5), _cloud(M, 2, lattice)
    d_ref, i_ref = KO.knn_c(q, s, K)
    d, i = KnnIndex(s.cuda()).query(q.cuda(), K)
    assert torch.equal(i.cpu(), i_ref)
    assert torch.equal(d.cpu(), d_ref)


def test_knn_points_signature_and_surface_cloud():
    sc = syn.make_scene(64, 96, 3, seed=9)
    sd_dummy = None
    from nerf_loc_b200.conditional_nerf import ConditionalNeRF
    from nerf_loc_b200.config import default_args
    m = ConditionalNeRF(default_args(16))
    _, xyz, _, _ = m.backproject_support_frame(sc["topk_images"], sc["feat_fine_src"], sc["topk_depths"],
                                               sc["topk_Ks"], sc["topk_poses"], stride=4)
    px = syn.random_pixels(64, 96, 64)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    z = torch.linspace(0.3, 5.0, 32)
    q = (ro[:, None] + rd[:, None] * z[None, :, None]).reshape(-1, 3)
    out = knn_points(q[None].cuda(), xyz[None].cuda(), K=8, return_nn=True)
    d_ref, i_ref = KO.knn_c(q, xyz, 8)
    assert out.idx.dtype == torch.int64 and tuple(out.idx.shape) == (1, q.shape[0], 8)
    assert torch.equal(out.idx[0].cpu(), i_ref) and torch.equal(out.dists[0].cpu(), d_ref)
    assert torch.equal(out.knn[0].cpu(), xyz[i_ref])


@pytest.mark.parametrize("K", [1, 3, 5, 8, 11])
def test_knn_points_ragged_batches_and_any_k(K):
    """knn_utils.py:97-222 with lengths1 / lengths2 and neighbour counts the kernel is not instantiated for: indices and squared
    distances bit for bit against the C oracle per cloud, zeros beyond the lengths and where a cloud has fewer than K points."""
    from nerf_loc_b200.knn import knn_gather
    g = torch.Generator().manual_seed(100 + K)
    p1 = torch.rand(3, 200, 3, generator=g) * 4 - 2
    p2 = torch.rand(3, 500, 3, generator=g) * 4 - 2
    p2[1, 10:20] = p2[1, 0:10]                                  # exact ties
    l1 = torch.tensor([200, 57, 1]); l2 = torch.tensor([500, 123, 4])
    out = knn_points(p1.cuda(), p2.cuda(), lengths1=l1.cuda(), lengths2=l2.cuda(), K=K, return_nn=True)
    assert tuple(out.idx.shape) == (3, 200, K) and out.idx.dtype == torch.int64
    for b in range(3):
        n1, n2 = int(l1[b]), int(l2[b])
        kk = min(K, n2)
        d_ref, i_ref = KO.knn_c(p1[b, :n1].contiguous(), p2[b, :n2].contiguous(), kk)
        assert torch.equal(out.idx[b, :n1, :kk].cpu(), i_ref) and torch.equal(out.dists[b, :n1, :kk].cpu(), d_ref)
        assert int(out.idx[b, :n1, kk:].abs().sum()) == 0 and float(out.dists[b, :n1, kk:].abs().sum()) == 0.0
        assert int(out.idx[b, n1:].abs().sum()) == 0 and float(out.dists[b, n1:].abs().sum()) == 0.0
    want = knn_gather(p2.cuda(), out.idx, l2.cuda())
    assert torch.equal(out.knn, want)
    assert float(out.knn[2, :, 4:].abs().sum()) == 0.0 if K > 4 else True


@pytest.mark.parametrize("S,per_ray", [(128, False), (64, True), (24, False)])
def test_knn_ray_samples_bit_exact(S, per_ray):
    """The render path's own search (8 lanes per query, warm start along the ray; csrc/knn.cu) against the C oracle: indices and
    squared distances bit for bit, on a surface cloud (multi-view back-projection) with shared and per-ray depths."""
    import ctypes
    from nerf_loc_b200 import _lib
    from nerf_loc_b200.conditional_nerf import ConditionalNeRF
    from nerf_loc_b200.config import default_args
    sc = syn.make_scene(64, 96, 4, seed=11)
    m = ConditionalNeRF(default_args(16))
    _, xyz, _, _ = m.backproject_support_frame(sc["topk_images"], sc["feat_fine_src"], sc["topk_depths"],
                                               sc["topk_Ks"], sc["topk_poses"], stride=4)
    R = 301
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], syn.random_pixels(64, 96, R))
    g = torch.Generator().manual_seed(S)
    z = torch.linspace(0.2, 6.0, S)
    if per_ray:
        z = (z[None] + 0.03 * torch.rand(R, S, generator=g)).sort(dim=1).values.contiguous()
    q = (ro[:, None, :] + rd[:, None, :] * (z[..., None] if per_ray else z[None, :, None])).reshape(-1, 3)
    d_ref, i_ref = KO.knn_c(q, xyz, 8)
    L = _lib.load()
    index = KnnIndex(xyz.cuda())
    geo = torch.zeros(xyz.shape[0], 8)
    geo[:, :3] = xyz
    geo, rod, rdd, zd = geo.cuda(), ro.cuda().contiguous(), rd.cuda().contiguous(), z.cuda().contiguous()
    idx = torch.empty(R * S, 8, dtype=torch.int32, device="cuda")
    d2 = torch.empty(R * S, 8, device="cuda")
    _lib.check(L.nlb_debug_knn_rays(_lib.ptr(index.buf), _lib.ptr(rod), _lib.ptr(rdd), _lib.ptr(zd), S if per_ray else 0,
                                    _lib.ptr(geo), R, S, _lib.ptr(idx), _lib.ptr(d2), _lib.stream()))
    torch.cuda.synchronize()
    assert torch.equal(idx.long().cpu(), i_ref)
    assert torch.equal(d2.cpu(), d_ref)


@pytest.mark.parametrize("name", list(RENDER_CASES))
def test_aggregator_and_support_points(name):
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    model.build_support_neural_points(data)
    sup = oracle_support(sd, sc)
    g = golden(name)
    for lv in ("coarse", "fine"):
        for k, v in sup[lv].items():
            assert relerr(model.support_neural_points[lv][k].cpu(), v) < TOL, (lv, k)
    assert relerr(model.support_neural_points["fine"]["confidence"].cpu(), g["conf_fine"]) < TOL
    z = O.sample_depths(S, *scene["depth_range"])
    xyz = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
    fm = sc["feat_fine_src"].permute(0, 3, 1, 2)
    with torch.no_grad():
        out_o, rf_o, vis_o = O.aggregator_forward(sd, "multiview_aggregator", xyz, scene["Ks"], scene["c2ws"],
                                                  scene["images"], fm, scene["vis_maps"], scene["depth_range"])
    out, rf, vis = model.multiview_aggregator(xyz.cuda(), data["topk_Ks"], data["topk_poses"], data["topk_images"],
                                              fm.cuda(), data["topk_depths"], data["depth_range"][0])
    assert relerr(vis.cpu(), vis_o) < TOL
    assert relerr(rf.cpu(), rf_o) < TOL
    assert relerr(out.cpu(), out_o) < TOL


@pytest.mark.parametrize("name", list(RENDER_CASES))
def test_query_vs_oracle_and_golden(name):
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    g = golden(name)
    z = O.sample_depths(S, *scene["depth_range"])
    xyz = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3)
    model.build_support_neural_points(data)
    q = model.query(data, xyz.cuda(), support_featmaps=data["feat_fine_src"].permute(0, 3, 1, 2),
                    support_neural_points=model.support_neural_points["fine"], direction=None, K=8)
    assert torch.equal(q["knn_idx"].long().cpu(), g["knn_idx"])
    assert relerr(q["multiview_visibility"].cpu(), g["q_vis"]) < TOL
    assert relerr(q["feature_agg"].cpu(), g["q_feature_agg"]) < TOL
    assert relerr(q["weights"].cpu(), g["q_weights"]) < TOL
    assert tuple(q["feature"].shape) == (xyz.shape[0], 8, 128)
    pts = oracle_support(sd, sc)["coarse"]["xyz"][::7][:40] + 0.01
    dc, p3, ndc = model.query_coarse(data, points=pts.cuda())
    df, _, _ = model.query_fine(data, pts.cuda())
    assert relerr(dc.cpu(), g["desc_coarse"]) < TOL
    assert relerr(df.cpu(), g["desc_fine"]) < TOL


@pytest.mark.parametrize("name", list(RENDER_CASES))
def test_render_rays_vs_golden(name):
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    g = golden(name)
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]}
    out = model.render_rays(data, rays, _debug=True)
    sup = oracle_support(sd, sc)
    with torch.no_grad():
        ref = O.render_rays(sd, scene, sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S,
                            return_debug=True)
    report = {k: relerr(out[k].cpu(), ref[k]) for k in ("feature_agg", "sigma", "rgb", "depth", "weights",
                                                         "depth_uncertainty", "feat")}
    print(name, report)
    for k, v in report.items():
        assert v < TOL, (k, report)
    for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat"):
        assert relerr(out[k].cpu(), g[k]) < TOL, k
    assert torch.equal(out["mask"].cpu(), g["mask"])


@pytest.mark.parametrize("H,W,V,stride", [(64, 96, 3, 4), (64, 96, 3, 8), (480, 640, 8, 4)])
def test_backprojection_bit_identical_with_host_ops(H, W, V, stride):
    """model.py:203-265 on the device (nlb_backproject_points) against the reference's operators on the host: support
    features, xyz, xyz_ndc and directions bit for bit - the KNN input must not depend on where the frame was set up."""
    from nerf_loc_b200.conditional_nerf import ConditionalNeRF
    from nerf_loc_b200.config import default_args
    sc = syn.make_scene(H, W, V, seed=21)
    m = ConditionalNeRF(default_args(16))
    feats = sc["feat_fine_src"] if stride == 4 else sc["feat_coarse_src"]
    with torch.no_grad():
        host = O.backproject_support_frame(sc["topk_images"], feats, sc["topk_depths"], sc["topk_Ks"], sc["topk_poses"], stride)
        dev = m.backproject_support_frame(sc["topk_images"].cuda(), feats.cuda(), sc["topk_depths"].cuda(),
                                          sc["topk_Ks"].cuda(), sc["topk_poses"].cuda(), stride=stride)
    assert host[1].shape[0] > 0
    for name, a, b in zip(("feature", "xyz", "xyz_ndc", "direction"), host, dev):
        assert a.shape == b.shape, name
        assert torch.equal(a, b.cpu()), "%s: %d values differ" % (name, int((a != b.cpu()).sum()))


def test_render_v8_s128_multi_chunk_vs_oracle():
    """The metric's shape (8 views: the fused view-weight path of the aggregator; 128 samples: a full 128-row tensor-core tile
    per ray) over several chunks of the internal chunk loop, with a ragged last chunk, against the oracle."""
    name = "render_v8_s128"
    S, H, W, V, _, wseed, sseed = RENDER_CASES[name]
    sc = syn.make_scene(H, W, V, seed=sseed)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], syn.random_pixels(H, W, 21))
    model, sd = cuda_model(S, wseed)
    data = setup_frame(model, sc)
    model.chunk_rays = 8
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]}
    out = model.render_rays(data, rays, _debug=True)
    sup = oracle_support(sd, sc)
    with torch.no_grad():
        ref = O.render_rays(sd, oracle_scene(sc), sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S,
                            return_debug=True)
    report = {k: relerr(out[k].cpu(), ref[k]) for k in ("feature_agg", "sigma", "rgb", "depth", "weights",
                                                         "depth_uncertainty", "feat")}
    print(report)
    for k, v in report.items():
        assert v < TOL, (k, report)
    assert torch.equal(out["mask"].cpu(), ref["mask"])
    model.chunk_rays = 21
    one = model.render_rays(data, rays)
    for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat", "mask"):
        assert torch.equal(one[k], out[k]), k


def test_render_chunking_is_invisible():
    name = "render_s16"
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]}
    a = model.render_rays(data, rays)
    model.chunk_rays = 7
    b = model.render_rays(data, rays)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_render_image_shapes_and_white_background():
    name = "render_s16"
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    sc = dict(sc)
    data = setup_frame(model, sc)
    data["H"], data["W"] = 8, 12   # render a small crop of the frame
    img = model.render_image(data)
    assert tuple(img["rgb"].shape) == (8, 12, 3) and tuple(img["weights"].shape) == (8, 12, S)
    data["white_bkgd"] = True
    img2 = model.render_image(data)
    wsum = img["weights"].sum(-1, keepdim=True)
    assert relerr(img2["rgb"].cpu(), (img["rgb"] + (1 - wsum)).cpu()) < 1e-5


def test_render_256_samples_vs_oracle():
    """Long rays (128 < S <= 256) take the slab kernel of render_ray_long.cu; S = 256 is the top of BASELINE's sweep."""
    from nerf_loc_b200 import params, synthetic as syn
    S, H, W, V, R, wseed, sseed = 256, 64, 96, 3, 5, 31, 13
    sc = syn.make_scene(H, W, V, seed=sseed)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], syn.random_pixels(H, W, R))
    model, sd = cuda_model(S, wseed)
    data = setup_frame(model, sc)
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]}
    out = model.render_rays(data, rays, _debug=True)
    sup = oracle_support(sd, sc)
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    with torch.no_grad():
        ref = O.render_rays(sd, scene, sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S,
                            return_debug=True)
    report = {k: relerr(out[k].cpu(), ref[k]) for k in ("feature_agg", "sigma", "rgb", "depth", "weights",
                                                         "depth_uncertainty", "feat")}
    print(report)
    for k, v in report.items():
        assert v < TOL, (k, report)
    assert torch.equal(out["mask"].cpu(), ref["mask"])


def test_hierarchical_sampling_vs_oracle():
    """render.N_importance > 0 (model.py:486-496): nlb_hierarchical_depths + per-ray depths through nlb_render_rays.  The
    uniform draws are shared with the oracle; the searchsorted indices must agree exactly."""
    from nerf_loc_b200 import params, synthetic as syn
    from nerf_loc_b200.config import default_args
    from nerf_loc_b200.conditional_nerf import ConditionalNeRF
    S0, NI, H, W, V, R = 16, 16, 32, 64, 3, 12
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S0 + NI), 5)
    model = ConditionalNeRF(default_args(S0, NI)).eval()
    model.load_state_dict(sd, strict=False)
    model = model.cuda()
    sc = syn.make_scene(H, W, V, seed=8)
    data = setup_frame(model, sc)
    px = syn.random_pixels(H, W, R)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0], "pixel_coordinates": px.float().cuda(),
            "K": data["K"], "pose": data["pose"], "H": H, "W": W}
    u = torch.rand(R, NI, generator=torch.Generator().manual_seed(3))
    z, dc, inds = model.hierarchical_depths(data, rays, u)
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    with torch.no_grad():
        z_ref, dc_ref, inds_ref = O.hierarchical_depths(sd, scene, px.float(), sc["K"], sc["pose"], S0, NI, u=u)
        sup = oracle_support(sd, sc)
        ref = O.render_rays(sd, scene, sup["fine"], sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S0, z_vals=z_ref)
    assert torch.equal(inds.cpu(), inds_ref)
    assert relerr(z.cpu(), z_ref) < 1e-5 and relerr(dc.cpu(), dc_ref) < TOL
    out = model.render_rays(data, rays, _u=u)
    assert tuple(out["weights"].shape) == (R, S0 + NI) and relerr(out["depth_coarse"].cpu(), dc_ref) < TOL
    report = {k: relerr(out[k].cpu(), ref[k]) for k in ("rgb", "depth", "weights", "depth_uncertainty", "feat")}
    print(report)
    for k, v in report.items():
        assert v < TOL, (k, report)
    assert torch.equal(out["mask"].cpu(), ref["mask"])


def test_blend_prepare_is_the_linear_projection():
    """nlb_blend_prepare: the map-feature columns of rgb_blending_mlp.0 (model.py:90-96,532-535) applied per pixel; the
    render kernels rely on W f(x) = sum_t b_t (W f_t) for the bilinear fetch f(x)."""
    import ctypes
    from nerf_loc_b200 import _lib
    model, sd = cuda_model(16, 5)
    L = _lib.load()
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(1000, 192, generator=g).cuda()
    out = torch.empty(1000, 32, device="cuda")
    _lib.check(L.nlb_blend_prepare(_lib.ptr(model.packed_weights()), model.n_samples, _lib.ptr(feat), 1000, _lib.ptr(out),
                                   _lib.stream()))
    torch.cuda.synchronize()
    W = sd["rgb_blending_mlp.0.weight"].double()          # [32, 128 + 195 + 1 + 4]
    ref = feat.double().cpu() @ W[:, 128 + 3:128 + 195].t()
    assert relerr(out.cpu().double(), ref) < 1e-6


def test_render_edge_cases_empty_and_ragged():
    """R = 0 returns empty outputs; a ray count that is not a multiple of any tile size (neighbour tiles of 16 samples, chunks)
    gives the same rays as the full batch."""
    name = list(RENDER_CASES)[0]
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    full = model.render_rays(data, {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]})
    n = 13
    part = model.render_rays(data, {"rays_o": ro[:n].cuda(), "rays_d": rd[:n].cuda(), "depth_range": data["depth_range"][0]})
    for k in ("rgb", "depth", "weights", "feat", "depth_uncertainty"):
        assert torch.equal(part[k], full[k][:n]), k
    assert torch.equal(part["mask"], full["mask"][:n])
    empty = model.render_rays(data, {"rays_o": ro[:0].cuda(), "rays_d": rd[:0].cuda(), "depth_range": data["depth_range"][0]})
    assert empty["rgb"].shape[0] == 0 and empty["weights"].shape == (0, S)
