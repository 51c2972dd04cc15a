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


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if p1.shape[2] != 3:
        raise ValueError("nerf_loc_b200.knn_points supports D = 3 only (the only dimension the hot path uses)")
    if lengths1 is not None or lengths2 is not None:
        raise ValueError("nerf_loc_b200.knn_points: ragged batches are not on the hot path (lengths must be None)")
    dists, idxs = [], []
    for b in range(p1.shape[0]):
        d, i = KnnIndex(p2[b]).query(p1[b], K)
        dists.append(d)
        idxs.append(i)
    dists, idxs = torch.stack(dists), torch.stack(idxs)
    nn = knn_gather(p2, idxs) if return_nn else None
    return _KNN(dists=dists, idx=idxs, knn=nn)


def knn_gather(x, idx, lengths=None):
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    return x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
