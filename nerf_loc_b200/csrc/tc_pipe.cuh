// Warp-specialised tcgen05 pipeline for chains of 128-row GEMMs inside one CTA.
//
// A CTA has NT = 256 compute threads (warps 0-7) plus one controller warp (warp 8).  One lane of the controller
//   * streams the pre-split weight tiles (B operand, hi|lo, K-tile of 16) from L2 into a ring of NSTAGE shared-memory
//     stages with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier),
//   * issues the 3xTF32 tcgen05.mma sequence for every tile (lo*hi, hi*lo, hi*hi; fp32 accumulate in TMEM),
//   * commits stage reuse (`empty`) and layer completion (`d_ready`) through tcgen05.commit.
// The compute warps write the A operand of a layer (canonical K-major hi / lo tiles), arrive on `a_ready`, wait on
// `d_ready`, and run the epilogue straight out of TMEM (tcgen05.ld: warp w reads lanes 32*(w&3).., columns by w>>2).
// Copies are issued NSTAGE-1 tiles ahead and a refill always waits on the tile BEFORE the one just issued, so the tensor
// pipe never drains inside a layer.
#pragma once
#include "tc_common.cuh"

namespace nlb {
namespace tc {

constexpr int KTB = 16;        // K extent of one B stage
constexpr int NSTAGE = 4;
constexpr uint32_t STAGE_BYTES = 2u * 128u * KTB * 4u;   // hi + lo for N = 128: 16 KB
constexpr uint32_t B_SBO = KTB * 32u;                   // bytes between 8-row groups of a B tile

struct Layer {        // D[128 x N] (+)= A[128 x K] * B[N x K]^T
  const unsigned char* gB;   // packed weights: per K-tile [hi: N x 16 canonical][lo: N x 16 canonical]
  uint32_t a_hi, a_lo;       // shared-memory addresses of the A operand (canonical, 8-row groups a_sbo bytes apart)
  uint32_t a_sbo;
  int nkt;                   // K / 16
  int N;                     // multiple of 16, <= 128
  uint32_t tmem_col;         // accumulator column offset inside the CTA's TMEM allocation
  uint32_t flags;            // WAIT_A | SIGNAL_D | ACCUM
};
constexpr uint32_t WAIT_A = 1u;    // wait for the compute warps' a_ready before the first MMA of this GEMM
constexpr uint32_t SIGNAL_D = 2u;  // commit d_ready after the last MMA of this GEMM
constexpr uint32_t SIGNAL_AUX = 8u; // commit d_aux instead of d_ready
constexpr uint32_t ACCUM = 4u;     // accumulate onto what is already in the TMEM columns (K split over several GEMMs)

template <int NS>
struct SyncT {
  uint64_t full[NS];
  uint64_t empty[NS];
  uint64_t a_ready;
  uint64_t d_ready;
  uint64_t d_aux;     // completion of a GEMM that runs right behind another signalled one (an mbarrier must not run two
                      // phases ahead of its waiters)
  uint32_t tmem_slot;
};
using Sync = SyncT<NSTAGE>;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// executed by ALL threads of the CTA (NT compute + 32 controller) at kernel start
template <int NS>
__device__ __forceinline__ uint32_t setup(SyncT<NS>& sy, int warp, int lane, uint32_t tmem_cols) {
  if (warp == 8) {
    tmem_alloc(&sy.tmem_slot, tmem_cols);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { mbar_init(&sy.full[i], 1); mbar_init(&sy.empty[i], 1); }
      mbar_init(&sy.a_ready, 256);
      mbar_init(&sy.d_ready, 1);
      mbar_init(&sy.d_aux, 1);
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return sy.tmem_slot;
}

// executed by ALL threads at kernel end (after the last TMEM read)
template <int NS>
__device__ __forceinline__ void teardown(SyncT<NS>& sy, int warp, uint32_t tmem, uint32_t tmem_cols) {
  fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem, tmem_cols);
  }
}

// compute threads: the A operand of the next layer is in shared memory
template <int NS>
__device__ __forceinline__ void a_ready(SyncT<NS>& sy) {
  fence_async_smem();
  fence_before_sync();
  mbar_arrive(&sy.a_ready);
}
// compute threads: wait for the accumulator of the current layer
template <int NS>
__device__ __forceinline__ void wait_d(SyncT<NS>& sy, uint32_t& parity) {
  mbar_wait(&sy.d_ready, parity);
  parity ^= 1u;
  fence_after_sync();
}

// controller lane: runs `nlayers` GEMMs back to back
template <int NS>
__device__ __forceinline__ void controller(SyncT<NS>& sy, unsigned char* stages, uint32_t tmem, const Layer* L, int nlayers) {
  int total = 0;
  for (int i = 0; i < nlayers; ++i) total += L[i].nkt;
  int c_layer = 0, c_kt = 0, copied = 0;
  int m_layer = 0, m_kt = 0, issued = 0;
  uint32_t full_par = 0, empty_par = 0, a_par = 0;
  auto copy_next = [&]() {
    const int s = copied % NS;
    if (copied >= NS) {
      mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
      empty_par ^= 1u << s;
    }
    const uint32_t bytes = 2u * (uint32_t)L[c_layer].N * KTB * 4u;
    mbar_expect_tx(&sy.full[s], bytes);
    bulk_copy(stages + (size_t)s * STAGE_BYTES, L[c_layer].gB + (size_t)c_kt * bytes, bytes, &sy.full[s]);
    ++copied;
    if (++c_kt == L[c_layer].nkt) { c_kt = 0; ++c_layer; }
  };
  while (copied < total && copied < NS - 1) copy_next();
  while (issued < total) {
    const Layer& l = L[m_layer];
    if (m_kt == 0 && (l.flags & WAIT_A)) {
      mbar_wait(&sy.a_ready, a_par);
      a_par ^= 1u;
    }
    const int s = issued % NS;
    mbar_wait(&sy.full[s], (full_par >> s) & 1u);
    full_par ^= 1u << s;
    fence_after_sync();
    const uint32_t idesc = idesc_tf32(128, l.N);
    const uint32_t b_hi = smem_u32(stages + (size_t)s * STAGE_BYTES);
    const uint32_t b_lo = b_hi + (uint32_t)l.N * KTB * 4u;
    const uint32_t d = tmem + l.tmem_col;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a = pass == 0 ? l.a_lo : l.a_hi;   // lo*hi, hi*lo, hi*hi
      const uint32_t b = pass == 1 ? b_lo : b_hi;
#pragma unroll
      for (int ks = 0; ks < KTB / 8; ++ks) {
        const uint64_t ad = smem_desc(a + (uint32_t)(m_kt * (KTB / 4) + ks * 2) * 128u, 128u, l.a_sbo);
        const uint64_t bd = smem_desc(b + (uint32_t)ks * 256u, 128u, B_SBO);
        mma_tf32(d, ad, bd, idesc, ((m_kt | pass | ks) != 0 || (l.flags & ACCUM)) ? 1u : 0u);
      }
    }
    mma_commit(&sy.empty[s]);
    if (++m_kt == l.nkt) {
      if (l.flags & SIGNAL_D) mma_commit(&sy.d_ready);
      if (l.flags & SIGNAL_AUX) mma_commit(&sy.d_aux);
      m_kt = 0;
      ++m_layer;
    }
    ++issued;
    if (copied < total) copy_next();
  }
}

// canonical A-operand addressing helper: byte offset of (row r, col k), 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint32_t a_off(int r, int k, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 2) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

// store 4 consecutive k (k % 4 == 0) of row r as hi / lo
__device__ __forceinline__ void store_split4(unsigned char* hi_base, unsigned char* lo_base, int r, int k, uint32_t sbo,
                                             float x0, float x1, float x2, float x3) {
  float4 h, l;
  split_tf32(x0, h.x, l.x); split_tf32(x1, h.y, l.y); split_tf32(x2, h.z, l.z); split_tf32(x3, h.w, l.w);
  const uint32_t o = a_off(r, k, sbo);
  *reinterpret_cast<float4*>(hi_base + o) = h;
  *reinterpret_cast<float4*>(lo_base + o) = l;
}

}  // namespace tc
}  // namespace nlb
