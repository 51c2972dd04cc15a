// Self-test of the tcgen05 building blocks (tc_bf16.cuh): C[128 x 128] = A[128 x K] * W[128 x K]^T, one shot, no staging.
// Exposed as nlb_debug_tc_gemm for tests/test_gpu_tc.py.
#include "nlb_internal.h"
#include "tc_common.cuh"
#include "tc_bf16.cuh"

namespace nlb {

// bf16x3 building blocks (tc_bf16.cuh): C[128 x 128] = A[128 x K] * W[128 x K]^T, K a multiple of 16, K <= 64.
//   mode 4: A and B in the weight-tile layout (core matrices adjacent in K contiguous);
//   mode 5: A in the chunk-major layout, read through a view SHIFTED by one row: C[m] = A[m + 1] * W^T (row 127 reads the zero
//           row behind the tile) - the mechanism the ray kernel uses for the taps of its convolutions;
//   mode 6: A (hi and lo) in tensor memory, two bf16 per 32-bit column.
__global__ void __launch_bounds__(256, 1)
tc_test_bf16_kernel(const float* __restrict__ A, const float* __restrict__ W, const int K, const int mode, float* __restrict__ C) {
  extern __shared__ __align__(128) unsigned char tsm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int RA = 130;                            // chunk-major A: rows 0..127 data, 128..129 zero
  const uint32_t wtile = 128u * (uint32_t)K * 2u;    // one [128 x K] bf16 plane in the weight-tile layout
  const uint32_t atile = mode == 5 ? (uint32_t)(K / 8) * RA * 16u : wtile;
  unsigned char* aHi = tsm;
  unsigned char* aLo = aHi + atile;
  unsigned char* bHi = aLo + atile;
  unsigned char* bLo = bHi + wtile;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 32) tc::mbar_init(&mbar, 1);
  if (mode == 5) for (int i = tid; i < (int)(2 * atile / 4); i += 256) reinterpret_cast<uint32_t*>(aHi)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < 128 * (K / 2); i += 256) {
    const int r = i / (K / 2), k = (i - r * (K / 2)) * 2;
    uint32_t hi, lo;
    tc::split_bf16x2(A[r * K + k], A[r * K + k + 1], hi, lo);
    const uint32_t ao = mode == 5 ? tc::cm_off(r, k, RA) : tc::wt_off(r, k, K);
    *reinterpret_cast<uint32_t*>(aHi + ao) = hi;
    *reinterpret_cast<uint32_t*>(aLo + ao) = lo;
    tc::split_bf16x2(W[r * K + k], W[r * K + k + 1], hi, lo);
    *reinterpret_cast<uint32_t*>(bHi + tc::wt_off(r, k, K)) = hi;
    *reinterpret_cast<uint32_t*>(bLo + tc::wt_off(r, k, K)) = lo;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_bf16(128, 128);
  const uint32_t b_hi32 = tc::desc_hi((uint32_t)K * 16u);
  if (mode == 6) {
    if (warp < 4) {
      const int row = warp * 32 + lane;
      const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::split_bf16x2(A[row * K + k0 + 2 * j], A[row * K + k0 + 2 * j + 1], hi[j], lo[j]);
        tc::tmem_st8_u(base + 128u + (uint32_t)(k0 / 2), hi);
        tc::tmem_st8_u(base + 192u + (uint32_t)(k0 / 2), lo);
      }
      tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
      bool acc = false;
      for (int pass = 0; pass < 3; ++pass) {                       // lo*hi, hi*lo, hi*hi
        const uint32_t a = tmem + (pass == 0 ? 192u : 128u);
        const unsigned char* b = pass == 1 ? bLo : bHi;
        for (int s = 0; s < K / 16; ++s) {
          tc::mma_bf16_ts_w(tmem, a + (uint32_t)(s * 8), tc::desc_lo(tc::smem_u32(b) + s * 256, 128), b_hi32, idesc, acc);
          acc = true;
        }
      }
      tc::mma_commit(&mbar);
    }
  } else if (tid == 0) {
    bool acc = false;
    const uint32_t a_lbo = mode == 5 ? RA * 16u : 128u, a_sbo = mode == 5 ? 128u : (uint32_t)K * 16u;
    const uint32_t a_step = mode == 5 ? 2u * RA * 16u : 256u;     // bytes per K = 16 step
    const uint32_t a_shift = mode == 5 ? 16u : 0u;                 // one row down
    for (int pass = 0; pass < 3; ++pass) {
      const unsigned char* a = pass == 0 ? aLo : aHi;
      const unsigned char* b = pass == 1 ? bLo : bHi;
      for (int s = 0; s < K / 16; ++s) {
        tc::mma_bf16_w(tmem, tc::desc_lo(tc::smem_u32(a) + a_shift + s * a_step, a_lbo), tc::desc_hi(a_sbo),
                       tc::desc_lo(tc::smem_u32(b) + s * 256, 128), b_hi32, idesc, acc);
        acc = true;
      }
    }
    tc::mma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::fence_after_sync();
  {
    const int row = (warp & 3) * 32 + lane;
    const int c0 = (warp >> 2) * 64;
    for (int cc = 0; cc < 64; cc += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c0 + cc), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) C[row * 128 + c0 + cc + j] = v[j];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

int launch_tc_test(const float* A, const float* W, int K, int mode, float* C, cudaStream_t st) {
  if (mode < 4 || mode > 6) return set_error("tc_test: mode must be 4 (plain operands), 5 (shifted chunk-major A) or 6 (A in tensor memory)");
  if (K % 16 != 0 || K < 16 || K > 64) return set_error("tc_test: K must be 16, 32, 48 or 64");
  const size_t smem_b = (size_t)2 * (K / 8) * 130 * 16 + (size_t)4 * 128 * K * 2;
  cudaError_t eb = cudaFuncSetAttribute(tc_test_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
  if (eb != cudaSuccess) return set_error(cudaGetErrorString(eb));
  tc_test_bf16_kernel<<<1, 256, smem_b, st>>>(A, W, K, mode, C);
  return check_launch("tc_test_bf16_kernel");
}

}  // namespace nlb
