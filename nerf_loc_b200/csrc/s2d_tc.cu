// S2DMatching scores (matching/sparse_to_dense.py:116-151) on tcgen05.
//
//   score[n, m] = sigmoid(w3 . relu(W2 relu(W1 (a_n * b_m) + b1) + b2) + b3),   n < N3 3D descriptors, m < Mc cells of the 2D map.
//
// W1 (a_n * b_m) = (W1 diag(a_n)) b_m: a CTA owns one 3D descriptor at a time, folds it into the first layer ONCE (B_n, a
// [128 x 192] bf16 hi | lo operand resident in shared memory next to W2) and sweeps the 2D map in tiles of 128 cells, whose
// descriptors were split into bf16 hi | lo K-slabs once per call and are streamed by bulk copies (vectorised sweeps over the
// feature grid; 98 KB per tile from L2).  Per tile: layer 1 = 36 tcgen05.mma (bf16x3, K = 192) into tensor memory, E1 = bias,
// ReLU, split back into tensor memory as the A operand of layer 2 (24 MMAs), E2 = bias, ReLU, dot with w3, sigmoid.  The
// [N3, Mc, 192] tensor of the reference (15 GB at 4096 x 4800) never exists; the FFMA2 kernel this replaces (match.cu) spent
// 42 ms on the 1.62 TFLOP.
#include "match_kernels.h"
#include "nlb_common.cuh"
#include "tc_bf16.cuh"
#include "tc_pipe.cuh"

namespace nlb {
namespace s2dtc {

constexpr int NS = 3;                                 // A-slab stages of 16 KB
constexpr uint32_t SLAB = 16384;                      // [128 x 32] hi 8 KB | lo 8 KB
constexpr uint32_t BN_OFF = 0;                        // B_n: 6 slabs
constexpr uint32_t W2_OFF = 6 * SLAB;                 // W2: 4 K-tiles
constexpr uint32_t STG_OFF = 10 * SLAB;
constexpr uint32_t SMALL_OFF = STG_OFF + NS * SLAB;   // a_n [192] | b1 [128] | b2 [128] | w3 [128] | red [2][128]
constexpr uint32_t SYNC_OFF = SMALL_OFF + (192 + 3 * 128 + 256) * 4;
constexpr uint32_t SMEM_BYTES = SYNC_OFF + 128;
static_assert(SMEM_BYTES <= 232448, "s2d_tc_kernel: shared memory budget");
constexpr uint32_t TM_D = 0, TM_AHI = 128, TM_ALO = 192;

struct Sync {
  uint64_t full[NS], empty[NS];
  uint64_t a_ready, d_ready;
  uint32_t tmem_slot;
};

// cells [M][192] fp32 -> per 128-cell tile and 32-column K-slab: hi plane | lo plane in the weight-tile layout (tc_bf16.cuh)
__global__ void split_cells_kernel(const float* __restrict__ desc1, const int64_t M, const int64_t ntiles, unsigned char* __restrict__ out) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= ntiles * 128 * 96) return;
  const int64_t cell = idx / 96;
  const int k = (int)(idx - cell * 96) * 2;
  const int64_t tile = cell >> 7;
  const int r = (int)(cell & 127), kt = k >> 5, kl = k & 31;
  float x0 = 0.f, x1 = 0.f;
  if (cell < M) { x0 = desc1[cell * 192 + k]; x1 = desc1[cell * 192 + k + 1]; }
  uint32_t hi, lo;
  tc::split_bf16x2(x0, x1, hi, lo);
  unsigned char* p = out + (size_t)(tile * 6 + kt) * SLAB + tc::wt_off(r, kl, 32);
  *reinterpret_cast<uint32_t*>(p) = hi;
  *reinterpret_cast<uint32_t*>(p + 8192) = lo;
}

__global__ void __launch_bounds__(NT + 64, 1)
s2d_tc_kernel(const PairMlp m, const float* __restrict__ w2_packed, const float* __restrict__ desc0, const unsigned char* __restrict__ cells,
              const int64_t N, const int64_t M, float* __restrict__ score) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* sAn = reinterpret_cast<float*>(sm + SMALL_OFF);
  float* sB1 = sAn + 192, *sB2 = sB1 + 128, *sW3 = sB2 + 128, *sRed = sW3 + 128;
  Sync& sy = *reinterpret_cast<Sync*>(sm + SYNC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (int)((M + 127) / 128);
  const int nmy = (int)((N - blockIdx.x + gridDim.x - 1) / gridDim.x);     // 3D descriptors of this CTA (grid <= N)
  if (warp == 8) {
    tc::tmem_alloc(&sy.tmem_slot, 256);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { tc::mbar_init(&sy.full[i], 1); tc::mbar_init(&sy.empty[i], 1); }
      tc::mbar_init(&sy.a_ready, NT);
      tc::mbar_init(&sy.d_ready, 1);
    }
  }
  if (tid < NT) {
    for (int i = tid; i < 4 * (int)SLAB / 16; i += NT) reinterpret_cast<uint4*>(sm + W2_OFF)[i] = __ldg(reinterpret_cast<const uint4*>(w2_packed) + i);
    if (tid < 128) { sB1[tid] = __ldg(m.b1 + tid); sB2[tid] = __ldg(m.b2 + tid); sW3[tid] = __ldg(m.w3 + tid); }
    tc::fence_async_smem();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sy.tmem_slot;

  if (warp == 9) {
    // ------------------------------------------------ producer: the K-slabs of every cell tile, once per 3D descriptor -------------
    uint32_t empty_par = 0;
    int i = 0;
    for (int q = 0; q < nmy; ++q)
      for (int t = 0; t < ntiles; ++t)
        for (int kt = 0; kt < 6; ++kt, ++i) {
          const int s = i % NS;
          if (i >= NS) {
            tc::mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
            empty_par ^= 1u << s;
          }
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&sy.full[s], SLAB);
            tc::bulk_copy(sm + STG_OFF + (size_t)s * SLAB, cells + (size_t)(t * 6 + kt) * SLAB, SLAB, &sy.full[s]);
          }
          __syncwarp();
        }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer -----------------------------------------------------------------
    uint32_t full_par = 0, a_par = 0;
    int i = 0;
    const uint32_t stage0 = tc::smem_u32(sm + STG_OFF), bn = tc::smem_u32(sm + BN_OFF), w2 = tc::smem_u32(sm + W2_OFF);
    const uint32_t idesc = tc::idesc_bf16(128, 128);
    const uint32_t sbo32 = tc::desc_hi(32u * 16u);
    for (int q = 0; q < nmy; ++q)
      for (int t = 0; t < ntiles; ++t) {
        // layer 1: A = cell slab (shared memory), B = folded first layer.  Waits for B_n (first tile) / for E2 of the previous
        // tile to have read the accumulator.
        tc::mbar_wait(&sy.a_ready, a_par); a_par ^= 1u;
        tc::fence_after_sync();
        for (int kt = 0; kt < 6; ++kt, ++i) {
          const int s = i % NS;
          tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u); full_par ^= 1u << s;
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t ab = stage0 + (uint32_t)s * SLAB, bb = bn + (uint32_t)kt * SLAB;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
              const uint32_t ap = ab + (pass == 0 ? 8192u : 0u), bp = bb + (pass == 1 ? 8192u : 0u);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_w(tmem + TM_D, tc::desc_lo(ap + (uint32_t)ks * 256u, 128u), sbo32, tc::desc_lo(bp + (uint32_t)ks * 256u, 128u), sbo32, idesc,
                               kt > 0 || pass > 0 || ks > 0);
            }
            tc::mma_commit(&sy.empty[s]);
            if (kt == 5) tc::mma_commit(&sy.d_ready);
          }
          __syncwarp();
        }
        // layer 2: A = hidden tile in tensor memory, B = W2 (resident)
        tc::mbar_wait(&sy.a_ready, a_par); a_par ^= 1u;
        tc::fence_after_sync();
        if (tc::elect_one()) {
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = tmem + (pass == 0 ? TM_ALO : TM_AHI) + (uint32_t)(kt * 16);
              const uint32_t bp = w2 + (uint32_t)kt * SLAB + (pass == 1 ? 8192u : 0u);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts_w(tmem + TM_D, a + (uint32_t)ks * 8u, tc::desc_lo(bp + (uint32_t)ks * 256u, 128u), sbo32, idesc, kt > 0 || pass > 0 || ks > 0);
            }
          tc::mma_commit(&sy.d_ready);
        }
        __syncwarp();
      }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t d_par = 0;
    auto wait_d = [&]() { tc::mbar_wait(&sy.d_ready, d_par); d_par ^= 1u; tc::fence_after_sync(); };
    const float b3 = __ldg(m.b3);
    for (int q = 0; q < nmy; ++q) {
      const int64_t n = (int64_t)blockIdx.x + (int64_t)q * gridDim.x;
      // ---- fold a_n into the first layer: B_n[j][k] = W1[j][k] a_n[k], split, weight-tile layout (the MMAs of the previous
      // descriptor have all completed: its last E1 waited for them)
      if (tid < 192) sAn[tid] = __ldg(desc0 + n * 192 + tid);
      cta_sync();
      for (int idx = tid; idx < 128 * 96; idx += NT) {
        const int j = idx & 127, k = (idx >> 7) * 2;
        const float x0 = __ldg(m.w1t + k * 128 + j) * sAn[k], x1 = __ldg(m.w1t + (k + 1) * 128 + j) * sAn[k + 1];
        uint32_t hi, lo;
        tc::split_bf16x2(x0, x1, hi, lo);
        unsigned char* p = sm + BN_OFF + (uint32_t)(k >> 5) * SLAB + tc::wt_off(j, k & 31, 32);
        *reinterpret_cast<uint32_t*>(p) = hi;
        *reinterpret_cast<uint32_t*>(p + 8192) = lo;
      }
      tc::fence_async_smem();
      for (int t = 0; t < ntiles; ++t) {
        tc::fence_before_sync();
        tc::mbar_arrive(&sy.a_ready);          // B_n is in place / the accumulator of the previous tile has been read
        // ---- E1: + bias, ReLU -> A operand of layer 2 in tensor memory
        wait_d();
        {
          const int c0 = half * 64;
#pragma unroll
          for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            tc::tmem_ld32(trow + TM_D + (uint32_t)(c0 + cc), v);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              tc::split_bf16x2(fmaxf(v[2 * j] + sB1[c0 + cc + 2 * j], 0.f), fmaxf(v[2 * j + 1] + sB1[c0 + cc + 2 * j + 1], 0.f), hi[j], lo[j]);
            tc::tmem_st16_u(trow + TM_AHI + (uint32_t)((c0 + cc) / 2), hi);
            tc::tmem_st16_u(trow + TM_ALO + (uint32_t)((c0 + cc) / 2), lo);
          }
          tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&sy.a_ready);
        // ---- E2: + bias, ReLU, dot with w3 (two column halves per row), sigmoid
        wait_d();
        {
          const int c0 = half * 64;
          float part = 0.f;
#pragma unroll
          for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            tc::tmem_ld32(trow + TM_D + (uint32_t)(c0 + cc), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) part = fmaf(fmaxf(v[j] + sB2[c0 + cc + j], 0.f), sW3[c0 + cc + j], part);
          }
          float* red = sRed + (t & 1) * 128;    // alternating buffers: one barrier per tile
          if (half == 1) red[row] = part;
          cta_sync();
          if (half == 0) {
            const int64_t cell = (int64_t)t * 128 + row;
            const float logit = part + red[row] + b3;
            if (cell < M) score[n * M + cell] = 1.f / (1.f + expf(-logit));
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 256);
  }
}

}  // namespace s2dtc

int launch_s2d_tc(const MatchW& w, const float* desc0, const float* desc1, int64_t N, int64_t M, float* score, cudaStream_t st) {
  const int64_t ntiles = (M + 127) / 128;
  // the split copy of the cell descriptors (98 KB per 128 cells) lives in a buffer kept per host thread and device and grown on
  // demand: nlb_s2d_scores has no scratch argument, and a stream-ordered allocation per call cost several milliseconds whenever
  // the pool had been trimmed at a synchronisation
  static thread_local unsigned char* cache[16] = {};
  static thread_local size_t cache_bytes[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return set_error("s2d: unsupported device ordinal");
  const size_t need = (size_t)ntiles * 6 * s2dtc::SLAB;
  cudaError_t e = cudaSuccess;
  if (cache_bytes[dev] < need) {
    if (cache[dev]) { cudaStreamSynchronize(st); cudaFree(cache[dev]); cache[dev] = nullptr; cache_bytes[dev] = 0; }
    e = cudaMalloc(reinterpret_cast<void**>(&cache[dev]), need);
    if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
    cache_bytes[dev] = need;
  }
  unsigned char* cells = cache[dev];
  const int64_t nthreads = ntiles * 128 * 96;
  s2dtc::split_cells_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(desc1, M, ntiles, cells);
  e = cudaFuncSetAttribute(s2dtc::s2d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2dtc::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(N < sms ? N : sms);
  s2dtc::s2d_tc_kernel<<<grid, NT + 64, s2dtc::SMEM_BYTES, st>>>(w.coarse, w.tb_w2c, desc0, cells, N, M, score);
  return check_launch("s2d_tc_kernel");
}

}  // namespace nlb
