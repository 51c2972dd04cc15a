"""Lane / bank / fragment bookkeeping that the CUDA kernels rely on, restated in numpy (CPU, no GPU needed).

These are the non-obvious index identities behind
  * the visibility decoder of aggregate_kernel (csrc/render_point.cu, phase 3): XOR-swizzled weight staging, and the layer-1
    accumulator fragment reused as the layer-2 A operand with the B fragment read in the matching k order;
  * the attention scores of neighbor_kernel (csrc/neighbor_tc.cu, phase A): the 16-lane transposing butterfly.
A change of either kernel that breaks one of them shows up here before it shows up as a parity failure on the GPU.
"""
import numpy as np


def lanes():
    lane = np.arange(32)
    return lane >> 2, lane & 3   # g (row / column index inside the fragment), t


def banks(addr_floats):
    return np.asarray(addr_floats) % 32


def test_decoder_weight_swizzle_is_conflict_free():
    g, t = lanes()
    # layer 1: dec1 (k, n) staged at k * 128 + (n ^ 8 (k & 3)); B fragment of k-step ks, n-tile nt: rows 8 ks + t and + 4, column g
    for ks in range(4):
        for nt in range(16):
            for dk in (0, 4):
                k = 8 * ks + t + dk
                n = nt * 8 + g
                addr = k * 128 + (n ^ ((k & 3) << 3))
                assert len(set(banks(addr))) == 32
                # what the kernel computes: row (8 ks + t) [+ 4 rows], column n ^ (t << 3)
                assert np.array_equal(addr, (8 * ks + t) * 128 + dk * 128 + (n ^ (t << 3)))
    # layer 2: dec2 head block (k, n) staged at k * 32 + (n ^ 8 ((k >> 1) & 3)); B fragment rows 8 ks + 2t and + 1, column g
    for ks in range(4):
        for nt in range(4):
            for dk in (0, 1):
                k = 8 * ks + 2 * t + dk
                n = nt * 8 + g
                addr = k * 32 + (n ^ (((k >> 1) & 3) << 3))
                assert len(set(banks(addr))) == 32
                assert np.array_equal(addr, (8 * ks + 2 * t) * 32 + dk * 32 + (n ^ (t << 3)))
    # the staging copy moves 16-byte chunks: the swizzle must keep 4-float groups intact
    for k in range(32):
        for n4 in range(0, 128, 4):
            d = [k * 128 + ((n4 + j) ^ ((k & 3) << 3)) for j in range(4)]
            assert d == list(range(d[0], d[0] + 4)) and d[0] % 4 == 0


def mma_m16n8k8(a_frag, b_frag):
    """mma.sync.m16n8k8 (row.col) on per-lane fragments: a_frag [32][4] = (row g, k t), (g + 8, t), (g, t + 4), (g + 8, t + 4);
    b_frag [32][2] = (k t, n g), (k t + 4, n g); returns c_frag [32][4] = (row g, col 2t), (g, 2t + 1), (g + 8, 2t), (g + 8, 2t + 1)."""
    g, t = lanes()
    A = np.zeros((16, 8))
    B = np.zeros((8, 8))
    A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a_frag[:, 0], a_frag[:, 1], a_frag[:, 2], a_frag[:, 3]
    B[t, g], B[t + 4, g] = b_frag[:, 0], b_frag[:, 1]
    C = A @ B
    return np.stack([C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1]], axis=1)


def test_accumulator_fragment_as_next_a_operand():
    """Layer 2 of a decoder head: H2 = H1 W2 with H1 held as layer-1 ACCUMULATOR fragments.  The kernel feeds accumulator
    registers (c0, c2, c1, c3) of n-tile ks as the A fragment of k-step ks and reads W2 rows (8 ks + 2t, 8 ks + 2t + 1)."""
    rng = np.random.default_rng(0)
    g, t = lanes()
    H1 = rng.standard_normal((16, 32))      # 16-row tile, the head's 32 hidden columns
    W2 = rng.standard_normal((32, 32))      # [k][n]
    # layer-1 accumulator fragments: n-tile j holds columns 8 j + 2t, + 1 of rows g, g + 8
    c1 = [np.stack([H1[g, 8 * j + 2 * t], H1[g, 8 * j + 2 * t + 1], H1[g + 8, 8 * j + 2 * t], H1[g + 8, 8 * j + 2 * t + 1]], axis=1)
          for j in range(4)]
    out = np.zeros((16, 32))
    for nt in range(4):
        acc = np.zeros((32, 4))
        for ks in range(4):
            c = c1[ks]
            a_frag = np.stack([c[:, 0], c[:, 2], c[:, 1], c[:, 3]], axis=1)
            b_frag = np.stack([W2[8 * ks + 2 * t, nt * 8 + g], W2[8 * ks + 2 * t + 1, nt * 8 + g]], axis=1)
            acc += mma_m16n8k8(a_frag, b_frag)
        out[g, nt * 8 + 2 * t], out[g, nt * 8 + 2 * t + 1] = acc[:, 0], acc[:, 1]
        out[g + 8, nt * 8 + 2 * t], out[g + 8, nt * 8 + 2 * t + 1] = acc[:, 2], acc[:, 3]
    np.testing.assert_allclose(out, H1 @ W2, rtol=1e-12, atol=1e-12)


def test_head_output_reduction_over_fragment():
    """A head output is a 32-long dot product with a layer-2 row: 8 columns in-thread, then the 4 lanes t of a row group."""
    rng = np.random.default_rng(1)
    g, t = lanes()
    H2 = rng.standard_normal((16, 32))
    w3 = rng.standard_normal(32)
    p0 = np.zeros(32)
    p1 = np.zeros(32)
    for nt in range(4):
        p0 += H2[g, nt * 8 + 2 * t] * w3[nt * 8 + 2 * t] + H2[g, nt * 8 + 2 * t + 1] * w3[nt * 8 + 2 * t + 1]
        p1 += H2[g + 8, nt * 8 + 2 * t] * w3[nt * 8 + 2 * t] + H2[g + 8, nt * 8 + 2 * t + 1] * w3[nt * 8 + 2 * t + 1]
    lane = np.arange(32)
    for x in (1, 2):   # __shfl_xor_sync(.., 1) then (.., 2)
        p0 = p0 + p0[lane ^ x]
        p1 = p1 + p1[lane ^ x]
    ref = H2 @ w3
    # lane t = 0 finishes row g, lane t = 1 finishes row g + 8
    np.testing.assert_allclose(p0[t == 0], ref[g[t == 0]], rtol=1e-12)
    np.testing.assert_allclose(p1[t == 1], ref[g[t == 1] + 8], rtol=1e-12)


def test_transposing_butterfly_of_attention_scores():
    """16 lanes (column slices of one sample) each hold 32 partial sums v[h * 8 + k]; after the 16 + 8 + 4 + 2 exchange steps lane
    sl holds the two TOTALS of index 2 sl, 2 sl + 1, i.e. head sl >> 2, k = 2 (sl & 3) and + 1."""
    rng = np.random.default_rng(2)
    part = rng.standard_normal((16, 32))             # [lane sl][h * 8 + k]
    v = part.copy()
    sl = np.arange(16)
    w2 = 16
    while w2 >= 2:
        up = (sl & (w2 >> 1)) != 0
        send = np.where(up[:, None], v[:, :w2], v[:, w2:2 * w2])
        keep = np.where(up[:, None], v[:, w2:2 * w2], v[:, :w2])
        v = v.copy()
        v[:, :w2] = keep + send[sl ^ (w2 >> 1)]
        w2 >>= 1
    total = part.sum(axis=0)
    np.testing.assert_allclose(v[:, 0], total[2 * sl], rtol=1e-12)
    np.testing.assert_allclose(v[:, 1], total[2 * sl + 1], rtol=1e-12)
    # ... which is where the kernel stores them: sSc[p * 32 + head * 8 + k] with head * 8 + k == 2 sl (+ 1)
    assert np.array_equal((sl >> 2) * 8 + 2 * (sl & 3), 2 * sl)


def test_feature_store_lane_swap_is_conflict_free():
    """aggregate_kernel phase 5: lane l holds channels (2l, 2l + 1) of a 64-channel group stored at column 3 + ...; lanes 16-31
    store their odd channel first so that each of the two stores covers 32 distinct banks."""
    lane = np.arange(32)
    up = lane >> 4
    for first in (True, False):
        off = up if first else up ^ 1
        addr = 3 + lane * 2 + off
        assert len(set(banks(addr))) == 32
