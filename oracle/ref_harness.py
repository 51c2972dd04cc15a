"""Import harness for the REAL reference (test infrastructure only).

Puts /root/reference on sys.path behind a handful of shims for third-party
names that are absent from this image, rebinds the
`Projector` / `fused_mean_variance` names to the in-repo implementations the
call sites were written against, and exposes the
reference classes.  It exists to (1) pin `oracle/nerfloc_oracle.py` against the
reference itself and (2) generate the golden vectors under tests/golden/.

/root/reference does not exist on the GPU box: nothing that runs there may
import this module.  `available()` tells callers whether the tree is present.
No reference source is copied; the modules are imported from where they lie.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("NERFLOC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nerf_loc", "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def default_args(n_samples=64, n_importance=0):
    """Defaults of nerf_loc/configs/__init__.py:4-92 as plain namespaces
    (yacs is not installed)."""
    ns = types.SimpleNamespace
    return ns(
        backbone2d_fpn_dim=192, model_3d_hidden_dim=128, matcher_hidden_dim=192,
        use_scene_coord_memorization=False, encode_appearance=True,
        appearance_emb_dim=128, multires=10, multires_views=4, i_embed=0,
        use_depth_supervision=False, fine_matching_loss_type='l2_with_std',
        render=ns(N_samples=n_samples, N_importance=n_importance, N_rand=1024,
                  chunk=2048, lindisp=False, white_bkgd=False,
                  use_render_uncertainty=True, render_feature=True),
        matching=ns(fine_num_3d_keypoints=1024, coarse_num_3d_keypoints=1024),
    )


_loaded = None


def load():
    """Returns a namespace with the reference classes (ConditionalNeRF, Matcher,
    S2DMatching, ...).  Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # --- pytorch3d.ops: exact KNN restated with torch (third-party behaviour:
    # K smallest squared-L2 by (dist, index), ascending) ---
    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1,
                   return_nn=False, return_sorted=True):
        from collections import namedtuple
        KNN = namedtuple("KNN", "dists idx knn")
        assert p1.shape[0] == 1 and p2.shape[0] == 1
        a, b = p1[0], p2[0]
        idx_all, d_all = [], []
        for s in range(0, a.shape[0], 4096):
            q = a[s:s + 4096]
            diff = q[:, None, :] - b[None, :, :]
            d2 = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
            d, i = torch.topk(d2, K, dim=1, largest=False, sorted=True)
            idx_all.append(i)
            d_all.append(d)
        idx = torch.cat(idx_all)[None]
        dists = torch.cat(d_all)[None]
        nn_pts = knn_gather(p2, idx) if return_nn else None
        return KNN(dists=dists, idx=idx, knn=nn_pts)

    def knn_gather(x, idx, lengths=None):
        N, M, U = x.shape
        _, L, K = idx.shape
        return x[0][idx[0].reshape(-1)].reshape(1, L, K, U)

    _stub("pytorch3d")
    _stub("pytorch3d.ops", knn_points=knn_points, knn_gather=knn_gather)
    _stub("inplace_abn", ABN=type("ABN", (), {}))

    # --- kornia 0.6.4 DSNT helpers (closed forms) ---
    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], dim=-1)[None]

    def spatial_expectation2d(inp, normalized_coordinates=True):
        b, c, h, w = inp.shape
        grid = create_meshgrid(h, w, normalized_coordinates, inp.device, inp.dtype)
        px = grid[..., 0].reshape(-1)
        py = grid[..., 1].reshape(-1)
        flat = inp.reshape(b, c, -1)
        ex = torch.sum(px * flat, -1, keepdim=True)
        ey = torch.sum(py * flat, -1, keepdim=True)
        return torch.cat([ex, ey], -1)

    _stub("kornia")
    _stub("kornia.geometry")
    _stub("kornia.geometry.subpix")
    dsnt = _stub("kornia.geometry.subpix.dsnt", spatial_expectation2d=spatial_expectation2d)
    sys.modules["kornia.geometry.subpix"].dsnt = dsnt
    _stub("kornia.utils")
    _stub("kornia.utils.grid", create_meshgrid=create_meshgrid)

    model = importlib.import_module("nerf_loc.models.conditional_nerf.model")
    agg = importlib.import_module("nerf_loc.models.conditional_nerf.multiview_aggregator")
    ibr = importlib.import_module("nerf_loc.models.ibrnet.ibrnet")
    # the call sites fit the in-repo copies, not third_party/IBRNet
    model.Projector = ibr.Projector
    agg.Projector = ibr.Projector
    agg.fused_mean_variance = ibr.fused_mean_variance

    # matcher.py imports coarse_matching (dead alternative) - importable as is
    matcher = importlib.import_module("nerf_loc.models.matcher")
    s2d = importlib.import_module("nerf_loc.models.matching.sparse_to_dense")
    fine = importlib.import_module("nerf_loc.models.matching.fine_matching")
    pe = importlib.import_module("nerf_loc.models.COTR.position_encoding")
    cn_utils = importlib.import_module("nerf_loc.models.conditional_nerf.utils")

    _loaded = types.SimpleNamespace(
        ConditionalNeRF=model.ConditionalNeRF, Matcher=matcher.Matcher,
        S2DMatching=s2d.S2DMatching, FineMatching=fine.FineMatching,
        FinePreprocess=fine.FinePreprocess, PositionEmbeddingSine=pe.PositionEmbeddingSine,
        get_embedder=cn_utils.get_embedder, get_rays=cn_utils.get_rays,
        sample_pdf=cn_utils.sample_pdf, knn_points=knn_points, knn_gather=knn_gather,
        model_module=model, agg_module=agg,
    )
    return _loaded
