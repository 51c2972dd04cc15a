"""KNN oracle pins: plain-C restatement vs the torch restatement vs the reference's own compiled knn_cpu.cpp."""
import pytest
import torch

from oracle import knn_oracle as KO
from oracle import nerfloc_oracle as O


def _cloud(n, seed, lattice=False):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n, 3, generator=g) * 2 - 1
    if lattice:  # many exact distance ties
        p = torch.round(p * 4) / 4
    return p


@pytest.mark.parametrize("K", [1, 8])
@pytest.mark.parametrize("lattice", [False, True])
def test_c_oracle_matches_torch_oracle(K, lattice):
    q, s = _cloud(300, 1, lattice), _cloud(1000, 2, lattice)
    d_c, i_c = KO.knn_c(q, s, K)
    d_t, i_t = O.knn_points(q, s, K)
    assert torch.equal(i_c, i_t)
    assert torch.equal(d_c, d_t)


def test_c_oracle_fewer_points_than_k():
    q, s = _cloud(5, 3), _cloud(3, 4)
    d, i = KO.knn_c(q, s, 8)
    assert (i[:, 3:] == 0).all() and (d[:, 3:] == 0).all()
    assert sorted(i[0, :3].tolist()) == [0, 1, 2]


@pytest.mark.skipif(not KO.ref_available(), reason="oracle/_ref/knn_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("K", [1, 8])
@pytest.mark.parametrize("lattice", [False, True])
def test_c_oracle_matches_reference_binary(K, lattice):
    q, s = _cloud(200, 5, lattice), _cloud(2000, 6, lattice)
    d_c, i_c = KO.knn_c(q, s, K)
    d_r, i_r = KO.knn_reference(q, s, K)
    assert torch.equal(d_c, d_r)
    assert torch.equal(i_c, i_r)
