"""`knn_points` / `knn_gather` with the reference's Python signatures (nerf_loc/models/ops/knn/knn_utils.py:97-222,
which is what `from pytorch3d.ops import knn_points, knn_gather` resolves to at conditional_nerf/model.py:20),
on top of the exact BVH search in csrc/knn.cu.  D = 3, one cloud per batch element."""
from collections import namedtuple

import torch

from . import _lib

_KNN = namedtuple("KNN", "dists idx knn")


class KnnIndex:
    """Per-frame search structure over a support cloud p2 [M,3] (built once, queried many times)."""

    def __init__(self, p2):
        L = _lib.load()
        self.p2 = _lib.f32(p2)
        if self.p2.dim() != 2 or self.p2.shape[1] != 3:
            raise ValueError("KnnIndex: expected points of shape [M,3]")
        self.M = self.p2.shape[0]
        nbytes = L.nlb_knn_index_bytes(self.M)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.p2.device)
        _lib.check(L.nlb_knn_build(_lib.ptr(self.p2), self.M, _lib.ptr(self.buf), nbytes, _lib.stream()))

    def query(self, p1, K):
        L = _lib.load()
        p1 = _lib.f32(p1)
        N = p1.shape[0]
        idx = torch.empty(N, K, dtype=torch.int64, device=p1.device)
        d2 = torch.empty(N, K, dtype=torch.float32, device=p1.device)
        _lib.check(L.nlb_knn_query(_lib.ptr(self.buf), _lib.ptr(p1), N, K, _lib.ptr(idx), _lib.ptr(d2), _lib.stream()))
        return d2, idx


_SEARCH_K = (1, 2, 4, 8, 16)   # neighbour counts the search kernel is instantiated for


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    """knn_utils.py:97-168.  `lengths1` / `lengths2` (ragged batches) and any K <= 16 are served by the same exact search:
    cloud b is searched over its first lengths2[b] points for its first lengths1[b] queries with the next instantiated
    neighbour count >= K, and the K nearest (ascending by (distance, index), which is what the reference returns for
    return_sorted=True and a valid answer for False) are kept.  As in the reference, outputs are zero where a cloud of p2
    has fewer than K points and for the rows of p1 beyond its length.  `version` selects among the reference's CUDA
    kernels and has no meaning here."""
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if p1.shape[2] != 3:
        raise ValueError("nerf_loc_b200.knn_points supports D = 3 only (the only dimension the hot path uses)")
    if K < 1 or K > _SEARCH_K[-1]:
        raise ValueError("nerf_loc_b200.knn_points: K must be in 1..16")
    Kq = next(k for k in _SEARCH_K if k >= K)
    N, P1 = p1.shape[0], p1.shape[1]
    l1 = [P1] * N if lengths1 is None else [int(v) for v in lengths1.tolist()]
    l2 = [p2.shape[1]] * N if lengths2 is None else [int(v) for v in lengths2.tolist()]
    if any(v < 0 or v > P1 for v in l1) or any(v < 0 or v > p2.shape[1] for v in l2):
        raise ValueError("nerf_loc_b200.knn_points: lengths out of range")
    dists = torch.zeros(N, P1, K, dtype=torch.float32, device=p1.device)
    idxs = torch.zeros(N, P1, K, dtype=torch.int64, device=p1.device)
    for b in range(N):
        if l1[b] == 0 or l2[b] == 0:
            continue
        d, i = KnnIndex(p2[b, :l2[b]]).query(p1[b, :l1[b]], Kq)   # (fewer than Kq support points: zero padded by the kernel)
        dists[b, :l1[b]] = d[:, :K]
        idxs[b, :l1[b]] = i[:, :K]
    nn = knn_gather(p2, idxs, lengths2) if return_nn else None
    return _KNN(dists=dists, idx=idxs, knn=nn)


def knn_gather(x, idx, lengths=None):
    """knn_utils.py:171-222: x_out[n, l, k] = x[n, idx[n, l, k]], zero where k >= lengths[n]."""
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    out = x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
    if lengths is not None and int(lengths.min()) < K:
        mask = lengths[:, None] <= torch.arange(K, device=x.device)[None]
        out = out.masked_fill(mask[:, None, :, None], 0.0)
    return out
