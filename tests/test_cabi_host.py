"""Host-only behaviour of the C ABI (include/nerfloc_b200.h): sizing helpers and argument validation.  None of these calls
reaches a CUDA API, so they run without a GPU; the error convention is the one INTEGRATION.md documents (non-zero return code,
message from nlb_last_error(), nothing written)."""
import ctypes

from nerf_loc_b200 import _lib


def test_scratch_size_and_launch_count():
    L = _lib.load()
    S, V = 128, 8
    small, big = L.nlb_render_scratch_bytes(18944, S, V), L.nlb_render_scratch_bytes(37888, S, V)
    # per sample: KNN idx + d2 [8], aggregator output + feature_agg [128], blend partials [V][32], rgb|vis [V][4],
    # view count, attention query + context [128] and the neighbour-weight sum, visibility | depth difference [V][2], out_fc input [416]
    per_sample = 2 * 8 * 4 + 2 * 128 * 4 + V * 32 * 4 + V * 16 + 1 + 2 * 128 * 4 + 4 + V * 8 + 416 * 4
    assert small >= 18944 * S * per_sample and small < 18944 * S * per_sample + (1 << 16)
    assert abs(big - 2 * small) < (1 << 16)
    assert L.nlb_render_scratch_bytes(0, S, V) == L.nlb_render_scratch_bytes(1, S, V)      # clamped, never zero
    assert L.nlb_render_scratch_bytes(1024, 192, V) > L.nlb_render_scratch_bytes(1024, 128, V) * 1.5   # S > 128 adds the slabs
    # seven kernels per chunk: KNN search, visibility, aggregate, q projection, neighbour, attention tail, ray
    assert L.nlb_render_launch_count(307200, 37888) == 7 * 9
    assert L.nlb_render_launch_count(307200, 0) == 7      # chunk_rays < 1: one chunk
    assert L.nlb_render_launch_count(0, 37888) == 0


def test_argument_validation_fails_loudly_without_touching_the_device():
    L = _lib.load()
    nul = ctypes.c_void_p(None)
    rc = L.nlb_render_rays(None, nul, 128, nul, nul, nul, 0, 16, 0, 16, nul, nul, nul, nul, nul, nul, nul, nul, nul, 0, nul)
    assert rc != 0 and b"scene is NULL" in L.nlb_last_error()
    rc = L.nlb_query_points(None, nul, 128, nul, nul, 16, 8, nul, nul, nul, nul, nul, nul, nul, nul, nul, 0, nul)
    assert rc != 0 and b"scene is NULL" in L.nlb_last_error()
    rc = L.nlb_knn_query(nul, nul, 16, 8, nul, nul, nul)
    assert rc != 0 and b"NULL pointer" in L.nlb_last_error()
    rc = L.nlb_debug_tc_gemm(nul, nul, 7, 4, nul, nul)
    assert rc != 0 and b"NULL pointer" in L.nlb_last_error()
