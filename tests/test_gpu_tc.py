"""tcgen05 building blocks: bf16x3 GEMM (plain operands, shifted chunk-major A, A in tensor memory) against an fp64 reference."""
import pytest
import torch

from nerf_loc_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [16, 32, 64])
def test_tc_gemm_bf16x3(K):
    """bf16x3 (tcgen05.mma kind::f16, hi / lo bf16 split, three passes) against fp64: plain operands (mode 4), the chunk-major
    A tile read through a view shifted by one row (mode 5: the taps of the ray kernel's convolutions), A in tensor memory
    (mode 6).  Tolerance: the dropped terms are <= 3 * 2^-16 per product; a K-long dot product of unit normals lands near 3e-6
    of the output's max norm."""
    L = _lib.load()
    g = torch.Generator().manual_seed(7 + K)
    A = torch.randn(128, K, generator=g).cuda()
    W = torch.randn(128, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    ref_shift = torch.zeros_like(ref)
    ref_shift[:127] = ref[1:]
    errs = {}
    for mode in (4, 5, 6):
        C = torch.full((128, 128), 7.0, device="cuda")
        _lib.check(L.nlb_debug_tc_gemm(_lib.ptr(A), _lib.ptr(W), K, mode, _lib.ptr(C), _lib.stream()))
        torch.cuda.synchronize()
        r = ref_shift if mode == 5 else ref
        errs[mode] = float((C.double() - r).abs().max() / ref.abs().max())
    print("K", K, "bf16x3 err", errs[4], "shifted chunk-major A err", errs[5], "A-in-TMEM err", errs[6])
    for mode in (4, 5, 6):
        assert errs[mode] < 2e-5, errs
