// bf16x3 operands for tcgen05.mma.kind::f16 (sm_100a), the second MMA mode of this library.
//
// An fp32 operand x is split into hi = bf16(x) and lo = bf16(x - hi) (round to nearest even both times); a product is
// evaluated as  Alo*Bhi + Ahi*Blo + Ahi*Bhi  with fp32 accumulation in tensor memory.  The dropped terms (lo*lo and the two
// third-order residuals) are <= 3 * 2^-16 of |a||b| per product, typically 2e-6 of a dot product's magnitude; measured against
// fp64 in tests/test_gpu_tc.py.  Compared with 3xTF32 (tc_common.cuh) every pass moves twice the K per instruction (K = 16
// per tcgen05.mma for 16-bit operands at the same 128 x N x 32-byte issue cost) and every operand takes half the bytes in
// shared memory and in L2 - the ray kernel is bound by exactly those two.
//
// Shared-memory operand layouts (K-major, no swizzle: 8-row x 16-byte core matrices = 8 rows x 8 bf16):
//   * weights (B):  tile [N x KT]: element (n, k) at (n/8)*SBO + (k/8)*128 + (n%8)*16 + (k%8)*2, SBO = KT*16, LBO = 128
//     (core matrices adjacent in K are contiguous) - what pack.cu writes and the producer warp streams;
//   * activations (A), "chunk-major":  element (r, k) at (k/8)*RA*16 + r*16 + (k%8)*2, i.e. LBO = RA*16 (one 8-column chunk of
//     ALL rows is contiguous), SBO = 128.  Consecutive rows of a chunk are 16 bytes apart, so a view of the tile shifted by j
//     rows is the same descriptor with its start address moved by 16*j bytes: a k=3 Conv1d along the ray accumulates its three
//     taps into ONE accumulator from three shifted views of the same tile (zero rows before and after the data), where the
//     3xTF32 kernel needed three accumulators and a row-shift exchange in the epilogue.
#pragma once
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace nlb {
namespace tc {

// instruction descriptor: bf16 x bf16 -> f32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 16 step; descriptors as (low, high) words; issued by ONE thread
__device__ __forceinline__ void mma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, bool accumulate) {
  if (accumulate) {
    asm volatile(
        "{\n"
        ".reg .b64 da, db;\n"
        ".reg .pred p;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .b64 da, db;\n"
        ".reg .pred p;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.eq.u32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  }
}

// Same with the A operand in tensor memory: lane = row, one 32-bit column per TWO consecutive k (k even in the low half)
__device__ __forceinline__ void mma_bf16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              bool accumulate) {
  if (accumulate) {
    asm volatile(
        "{\n"
        ".reg .b64 db;\n"
        ".reg .pred p;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .b64 db;\n"
        ".reg .pred p;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.eq.u32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  }
}

// descriptor words of a no-swizzle K-major operand
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

// ---- bf16 hi / lo split ------------------------------------------------------------------------------------------------
// two fp32 -> packed (hi0, hi1) and (lo0, lo1), element 0 in the low 16 bits
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// eight consecutive k of one row -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void split_store8(unsigned char* hi_ptr, unsigned char* lo_ptr, const float (&v)[8]) {
  uint4 h, l;
  split_bf16x2(v[0], v[1], h.x, l.x);
  split_bf16x2(v[2], v[3], h.y, l.y);
  split_bf16x2(v[4], v[5], h.z, l.z);
  split_bf16x2(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(hi_ptr) = h;
  *reinterpret_cast<uint4*>(lo_ptr) = l;
}

// chunk-major activation tile: byte offset of (physical row r, column k) inside a plane whose chunks hold RA rows
__device__ __forceinline__ uint32_t cm_off(int r, int k, int RA) { return (uint32_t)(k >> 3) * (uint32_t)RA * 16u + (uint32_t)r * 16u + (uint32_t)(k & 7) * 2u; }

// weight tile [N x KT] (see above): byte offset of (n, k)
__host__ __device__ __forceinline__ uint32_t wt_off(int n, int k, int KT) {
  return (uint32_t)(n >> 3) * (uint32_t)KT * 16u + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
}

// registers -> TMEM, 32-bit columns given as raw words
__device__ __forceinline__ void tmem_st16_u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8_u(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

}  // namespace tc
}  // namespace nlb
