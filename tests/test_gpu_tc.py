"""tcgen05 building blocks: single-pass tf32 and 3xTF32 GEMM against an fp64 reference."""
import pytest
import torch

from nerf_loc_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [8, 32, 64])
def test_tc_gemm_modes(K):
    L = _lib.load()
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).cuda()
    W = torch.randn(128, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    errs = {}
    for mode in (0, 1, 2):
        C = torch.zeros(128, 128, device="cuda")
        _lib.check(L.nlb_debug_tc_gemm(_lib.ptr(A), _lib.ptr(W), K, mode, _lib.ptr(C), _lib.stream()))
        torch.cuda.synchronize()
        errs[mode] = float((C.double() - ref).abs().max() / ref.abs().max())
    print("K", K, "tf32 err", errs[0], "3xTF32 err", errs[1], "3xTF32 A-in-TMEM err", errs[2])
    assert errs[0] < 5e-3
    assert errs[1] < 2e-6
    assert errs[2] < 2e-6   # A operand read from tensor memory (tcgen05.mma [d], [a], b-desc)


@pytest.mark.parametrize("K", [32, 64])
def test_mma_sync_3xtf32(K):
    """Warp-level path (rows16_mma, mma.sync.m16n8k8 tf32 with the hi / lo split) against fp64."""
    L = _lib.load()
    g = torch.Generator().manual_seed(100 + K)
    A = torch.randn(16, K, generator=g).cuda()
    W = torch.randn(128, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    Wt = W.t().contiguous()   # mode 3 takes the weight k-major, [K][128]
    C = torch.zeros(16, 128, device="cuda")
    _lib.check(L.nlb_debug_tc_gemm(_lib.ptr(A), _lib.ptr(Wt), K, 3, _lib.ptr(C), _lib.stream()))
    torch.cuda.synchronize()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    print("K", K, "mma.sync 3xTF32 err", err)
    assert err < 2e-6


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
