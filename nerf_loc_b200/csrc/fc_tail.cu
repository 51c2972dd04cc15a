// fc_tail_kernel: out_fc (393 -> 64 -> 128, ELU; ibrnet/ibrnet.py:194-231 as used by multiview_aggregator.py:202-213) and the
// attention query projection q = W_q aggregated (ibrnet.py:69-90) as a chain of three tensor-core GEMMs over tiles of 128 SAMPLES.
//
// Inside aggregate_kernel the two out_fc layers were 8-row GEMMs per CTA on FFMA2 (7 k of its 27 k clk per tile, every CTA
// re-reading the 139 KB of weights from L2); a tile of samples only exists across 16 of its CTAs, so the statistics vector
// (416 floats per sample) takes one round trip through HBM instead: aggregate_kernel writes it, this kernel streams it.
//   GEMM 1  [128 x 416] x [416 x 64]   A: K-slabs of 32 columns converted to bf16 hi | lo by the compute warps (double buffered,
//                                      the loads of slab k+1 are in flight while slab k is converted), B: 8 KB tiles by bulk copy
//   E1      + bias, ELU -> A operand in tensor memory
//   GEMM 2  [128 x 64] x [64 x 128]    -> E2: + bias, ELU = `aggregated` (global, fp32) and the A operand of
//   GEMM 3  [128 x 128] x [128 x 128]  -> q (global, fp32)
// 256 TMEM columns and 100 KB of shared memory per CTA: two persistent CTAs per SM.  The kernel is bound by the 2.7 KB per sample
// it moves through HBM.
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_bf16.cuh"
#include "tc_pipe.cuh"

namespace nlb {
namespace fct {

constexpr int NS = 4;                                  // weight stages of 16 KB
constexpr uint32_t STG_BYTES = 16384;
constexpr int G_LD = 416;
constexpr int NKT1 = G_LD / 32;                        // 13 K-slabs of GEMM 1
constexpr uint32_t TM_D = 0, TM_AHI = 128, TM_ALO = 192;
constexpr uint32_t A_OFF = 0;                          // 2 x ([128 x 32] hi 8 KB | lo 8 KB)
constexpr uint32_t STG_OFF = 2 * 16384;
constexpr uint32_t SYNC_OFF = STG_OFF + NS * STG_BYTES;
constexpr uint32_t SMEM_BYTES = SYNC_OFF + 128;

struct Sync {
  uint64_t full[NS], empty[NS];
  // a_slab[b]: K-slab buffer b is written (its own barrier per buffer: the compute warps need a_free[b], i.e. the issuer's
  // progress, before they can arrive on it again, so it can never run two phases ahead of the issuer); a_ready: E1 / E2 operands
  uint64_t a_slab[2], a_ready, a_free[2], d_ready;
  uint32_t tmem_slot;
};
static_assert(sizeof(Sync) <= 128, "fct::Sync");

__global__ void __launch_bounds__(NT + 64, 2)
fc_tail_kernel(const RenderW w, const float* __restrict__ G, const int64_t N, float* __restrict__ agg_out, float* __restrict__ q_out,
               const bool agg_pm) {
  extern __shared__ __align__(1024) unsigned char sm[];
  Sync& sy = *reinterpret_cast<Sync*>(sm + SYNC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 8) {
    tc::tmem_alloc(&sy.tmem_slot, 256);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { tc::mbar_init(&sy.full[i], 1); tc::mbar_init(&sy.empty[i], 1); }
      tc::mbar_init(&sy.a_ready, NT);
      tc::mbar_init(&sy.a_slab[0], NT); tc::mbar_init(&sy.a_slab[1], NT);
      tc::mbar_init(&sy.a_free[0], 1); tc::mbar_init(&sy.a_free[1], 1);
      tc::mbar_init(&sy.d_ready, 1);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sy.tmem_slot;
  const int64_t ntiles = (N + 127) / 128;
  const int nmy = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp == 9) {
    // ------------------------------------------------ weight producer: fc1 (13 x 8 KB), fc2 (2 x 16 KB), W_q (4 x 16 KB) per tile --------
    uint32_t empty_par = 0;
    int i = 0;
    for (int t = 0; t < nmy; ++t) {
      for (int l = 0; l < 3; ++l) {
        const unsigned char* gB = reinterpret_cast<const unsigned char*>(l == 0 ? w.tb_fc1 : (l == 1 ? w.tb_fc2 : w.tb_wq));
        const int nkt = l == 0 ? NKT1 : (l == 1 ? 2 : 4);
        const uint32_t bytes = l == 0 ? 8192u : 16384u;
        for (int kt = 0; kt < nkt; ++kt, ++i) {
          const int s = i % NS;
          if (i >= NS) {
            tc::mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
            empty_par ^= 1u << s;
          }
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&sy.full[s], bytes);
            tc::bulk_copy(sm + STG_OFF + (size_t)s * STG_BYTES, gB + (size_t)kt * bytes, bytes, &sy.full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer -----------------------------------------------------------------
    uint32_t full_par = 0, a_par = 0, slab_par = 0;
    int i = 0;
    const uint32_t stage0 = tc::smem_u32(sm + STG_OFF), a0 = tc::smem_u32(sm + A_OFF);
    const uint32_t a_hi32 = tc::desc_hi(128u), b_hi32 = tc::desc_hi(32u * 16u);
    for (int t = 0; t < nmy; ++t) {
      // GEMM 1: one K-slab of A (shared memory, chunk-major) per weight tile
      for (int kt = 0; kt < NKT1; ++kt, ++i) {
        const int s = i % NS;
        tc::mbar_wait(&sy.a_slab[kt & 1], (slab_par >> (kt & 1)) & 1u); slab_par ^= 1u << (kt & 1);
        tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u); full_par ^= 1u << s;
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::idesc_bf16(128, 64);
          const uint32_t ab = a0 + (uint32_t)(kt & 1) * 16384u, bb = stage0 + (uint32_t)s * STG_BYTES;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
            const uint32_t ap = ab + (pass == 0 ? 8192u : 0u), bp = bb + (pass == 1 ? 4096u : 0u);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_bf16_w(tmem + TM_D, tc::desc_lo(ap + (uint32_t)ks * 2u * 2048u, 2048u), a_hi32, tc::desc_lo(bp + (uint32_t)ks * 256u, 128u), b_hi32, idesc,
                             kt > 0 || pass > 0 || ks > 0);
          }
          tc::mma_commit(&sy.empty[s]);
          tc::mma_commit(&sy.a_free[kt & 1]);
          if (kt == NKT1 - 1) tc::mma_commit(&sy.d_ready);
        }
        __syncwarp();
      }
      // GEMM 2 (K = 64) and GEMM 3 (K = 128): A from tensor memory
      for (int l = 1; l < 3; ++l) {
        const int nkt = l == 1 ? 2 : 4;
        tc::mbar_wait(&sy.a_ready, a_par); a_par ^= 1u;
        tc::fence_after_sync();
        for (int kt = 0; kt < nkt; ++kt, ++i) {
          const int s = i % NS;
          tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u); full_par ^= 1u << s;
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t idesc = tc::idesc_bf16(128, 128);
            const uint32_t bb = stage0 + (uint32_t)s * STG_BYTES;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = tmem + (pass == 0 ? TM_ALO : TM_AHI) + (uint32_t)(kt * 16);
              const uint32_t bp = bb + (pass == 1 ? 8192u : 0u);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts_w(tmem + TM_D, a + (uint32_t)ks * 8u, tc::desc_lo(bp + (uint32_t)ks * 256u, 128u), b_hi32, idesc, kt > 0 || pass > 0 || ks > 0);
            }
            tc::mma_commit(&sy.empty[s]);
            if (kt == nkt - 1) tc::mma_commit(&sy.d_ready);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t d_par = 0, free_par = 0;
    int slab = 0;   // K-slabs written so far (over all tiles)
    auto wait_d = [&]() { tc::mbar_wait(&sy.d_ready, d_par); d_par ^= 1u; tc::fence_after_sync(); };
    // this thread's part of a K-slab: rows tid / 4 and 64 + tid / 4, the 8-column chunk tid % 4 (four lanes share a row's
    // 128-byte run of the slab: coalesced, where one lane per row was 32 separate lines per load instruction)
    const int64_t n_base = ((int64_t)blockIdx.x) * 128;
    auto fetch = [&](int64_t nb, int kt, float4 (&dst)[4]) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t nn = nb + u * 64 + (tid >> 2);
        if (nn < N) {
          const float4* p = reinterpret_cast<const float4*>(G + nn * G_LD + kt * 32 + (tid & 3) * 8);
          dst[2 * u] = __ldcs(p); dst[2 * u + 1] = __ldcs(p + 1);
        } else {
          dst[2 * u] = make_float4(0.f, 0.f, 0.f, 0.f); dst[2 * u + 1] = dst[2 * u];
        }
      }
    };
    for (int it = 0; it < nmy; ++it) {
      const int64_t nb = n_base + (int64_t)it * gridDim.x * 128;
      const int64_t n = nb + row;
      float4 cur[4], nxt[4];
      fetch(nb, 0, cur);
      for (int kt = 0; kt < NKT1; ++kt, ++slab) {
        if (kt + 1 < NKT1) fetch(nb, kt + 1, nxt);
        const int buf = kt & 1;
        if (slab >= 2) {   // the MMAs that read this buffer two slabs ago have completed
          tc::mbar_wait(&sy.a_free[buf], (free_par >> buf) & 1u);
          free_par ^= 1u << buf;
        }
        unsigned char* hi = sm + A_OFF + buf * 16384;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float v[8] = {cur[2 * u].x, cur[2 * u].y, cur[2 * u].z, cur[2 * u].w, cur[2 * u + 1].x, cur[2 * u + 1].y, cur[2 * u + 1].z, cur[2 * u + 1].w};
          const uint32_t o = (uint32_t)(tid & 3) * 2048u + (uint32_t)(u * 64 + (tid >> 2)) * 16u;
          tc::split_store8(hi + o, hi + 8192 + o, v);
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        tc::mbar_arrive(&sy.a_slab[buf]);
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      }
      // ---- E1: + bias, ELU -> A operand (K = 64: 32 packed columns per plane; this thread: its 32 output columns)
      wait_d();
      {
        float v[32];
        tc::tmem_ld32(trow + TM_D + (uint32_t)(half * 32), v);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          tc::split_bf16x2(elu(v[2 * j] + __ldg(w.fc1_b + half * 32 + 2 * j)), elu(v[2 * j + 1] + __ldg(w.fc1_b + half * 32 + 2 * j + 1)), hi[j], lo[j]);
        tc::tmem_st16_u(trow + TM_AHI + (uint32_t)(half * 16), hi);
        tc::tmem_st16_u(trow + TM_ALO + (uint32_t)(half * 16), lo);
        tc::tmem_st_wait();
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&sy.a_ready);
      // ---- E2: + bias, ELU = aggregated -> global, and the A operand of the query projection (K = 128)
      wait_d();
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tc::tmem_ld32(trow + TM_D + (uint32_t)(half * 64 + cc), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = elu(v[j] + __ldg(w.fc2_b + half * 64 + cc + j));
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(agg_out + (agg_pm ? pm128_off(n, half * 64 + cc + j) : n * W_HID + half * 64 + cc + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tc::split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        tc::tmem_st16_u(trow + TM_AHI + (uint32_t)((half * 64 + cc) / 2), hi);
        tc::tmem_st16_u(trow + TM_ALO + (uint32_t)((half * 64 + cc) / 2), lo);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(&sy.a_ready);
      // ---- E3: q -> global
      wait_d();
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tc::tmem_ld32(trow + TM_D + (uint32_t)(half * 64 + cc), v);
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(q_out + pm128_off(n, half * 64 + cc + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc::fence_before_sync();
      cta_sync();   // every thread has read the accumulator: the next tile's first MMA may overwrite it
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 256);
  }
}

}  // namespace fct

int launch_fc_tail(const RenderW& w, const float* g, int64_t N, float* agg, float* q, bool agg_pm, cudaStream_t st) {
  if (N <= 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(fct::fc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fct::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (N + 127) / 128;
  const unsigned grid = (unsigned)(tiles < 2 * sms ? tiles : 2 * sms);
  fct::fc_tail_kernel<<<grid, NT + 64, fct::SMEM_BYTES, st>>>(w, g, N, agg, q, agg_pm);
  return check_launch("fc_tail_kernel");
}

}  // namespace nlb
