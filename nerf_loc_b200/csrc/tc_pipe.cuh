// Warp-specialised tcgen05 pipeline for chains of 128-row GEMMs inside one CTA.
//
// A CTA has NT = 256 compute threads (warps 0-7) plus two single-lane service warps:
//   * warp 9, the producer, streams the pre-split weight tiles (B operand, hi|lo, 16 KB each) from L2 into a ring of
//     shared-memory stages with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier);
//   * warp 8, the MMA issuer, issues the 3xTF32 tcgen05.mma sequence of every tile (lo*hi, hi*lo, hi*hi; fp32 accumulate
//     in TMEM) and commits stage reuse (`empty`) and layer completion (`d_ready`) through tcgen05.commit.
// The compute warps write the A operand of a layer (canonical K-major hi / lo tiles), arrive on `a_ready`, wait on
// `d_ready`, and run the epilogue straight out of TMEM (tcgen05.ld: warp w reads lanes 32*(w&3).., columns by w>>2).
// The issuer's per-MMA work is two adds: descriptors are kept as (low, high) words and only the low word moves.
#pragma once
#include "tc_common.cuh"

namespace nlb {
namespace tc {

constexpr int KTB = 16;        // K extent of one B stage
constexpr int NSTAGE = 4;
constexpr uint32_t STAGE_BYTES = 2u * 128u * KTB * 4u;   // hi + lo for N = 128: 16 KB

struct Layer {        // D[128 x N] (+)= A[128 x K] * B[N x K]^T
  const unsigned char* gB;   // packed weights: per K-tile [hi: N x 16 canonical][lo: N x 16 canonical]
  uint32_t a_hi, a_lo;       // shared-memory addresses of the A operand (canonical, 8-row groups a_sbo bytes apart)
  uint32_t a_sbo;
  int nkt;                   // K / ktile
  int N;                     // multiple of 16, <= 128
  int ktile;                 // K extent of one weight tile: 2048 / N (16, 32 or 64), so that a tile is always 16 KB
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

// D[tmem] (+)= A * B^T for one K = 8 step; descriptors passed as (low, high) words so the issuer only adds to the low word
__device__ __forceinline__ void mma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, bool accumulate) {
  if (accumulate) {
    asm volatile(
        "{\n"
        ".reg .b64 da, db;\n"
        ".reg .pred p;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .b64 da, db;\n"
        ".reg .pred p;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.eq.u32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  }
}

// Same with the A operand in tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void mma_tf32_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              bool accumulate) {
  if (accumulate) {
    asm volatile(
        "{\n"
        ".reg .b64 db;\n"
        ".reg .pred p;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .b64 db;\n"
        ".reg .pred p;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.eq.u32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  }
}

// One elected lane of a converged warp (PTX elect.sync).  The service warps run their loops warp-uniformly and predicate only
// the tcgen05 / bulk-copy instructions with this, so descriptor words live in uniform registers - inside an `if (lane == 0)`
// region the compiler has to wrap every UTCHMMA in an R2UR waterfall loop instead.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}

// producer warp (warp 9, all lanes): streams every weight tile of the GEMM list through the stage ring, as far ahead as the ring allows
template <int NS>
__device__ __forceinline__ void producer(SyncT<NS>& sy, unsigned char* stages, const Layer* L, int nlayers) {
  uint32_t empty_par = 0;
  int i = 0;
  for (int l = 0; l < nlayers; ++l) {
    const uint32_t bytes = 2u * (uint32_t)L[l].N * (uint32_t)L[l].ktile * 4u;
    for (int kt = 0; kt < L[l].nkt; ++kt, ++i) {
      const int s = i % NS;
      if (i >= NS) {
        mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
        empty_par ^= 1u << s;
      }
      if (elect_one()) {
        mbar_expect_tx(&sy.full[s], bytes);
        bulk_copy(stages + (size_t)s * STAGE_BYTES, L[l].gB + (size_t)kt * bytes, bytes, &sy.full[s]);
      }
      __syncwarp();
    }
  }
}

// MMA warp (warp 8, all lanes; one elected lane issues): issues the 3xTF32 sequence of every tile and the completion commits
template <int NS>
__device__ __forceinline__ void mma_issuer(SyncT<NS>& sy, unsigned char* stages, uint32_t tmem, const Layer* L, int nlayers) {
  uint32_t full_par = 0, a_par = 0;
  int i = 0;
  const uint32_t stage0 = smem_u32(stages);
  for (int l = 0; l < nlayers; ++l) {
    const Layer lay = L[l];
    if (lay.flags & WAIT_A) {
      mbar_wait(&sy.a_ready, a_par);
      a_par ^= 1u;
    }
    const uint32_t idesc = idesc_tf32(128, lay.N);
    const uint32_t d = tmem + lay.tmem_col;
    const uint32_t a_hi32 = ((lay.a_sbo >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t b_hi32 = (((uint32_t)lay.ktile * 32u >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t lbo = (128u >> 4) << 16;
    const uint32_t a_lo_hi = ((lay.a_hi & 0x3FFFFu) >> 4) | lbo;   // low word of the descriptor of the A-hi tile
    const uint32_t a_lo_lo = ((lay.a_lo & 0x3FFFFu) >> 4) | lbo;
    const uint32_t half_tile = (uint32_t)lay.N * (uint32_t)lay.ktile * 4u;
    const int ksteps = lay.ktile / 8;
    bool acc = (lay.flags & ACCUM) != 0;
    for (int kt = 0; kt < lay.nkt; ++kt, ++i) {
      const int s = i % NS;
      mbar_wait(&sy.full[s], (full_par >> s) & 1u);
      full_par ^= 1u << s;
      fence_after_sync();
      const uint32_t b_base = stage0 + (uint32_t)s * STAGE_BYTES;
      const uint32_t b_lo_hi = ((b_base & 0x3FFFFu) >> 4) | lbo;
      const uint32_t b_lo_lo = (((b_base + half_tile) & 0x3FFFFu) >> 4) | lbo;
      const uint32_t a_step = (uint32_t)(kt * (lay.ktile / 4)) * 8u;   // (k/4)*128 bytes >> 4
      if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
          const uint32_t al = (pass == 0 ? a_lo_lo : a_lo_hi) + a_step;
          const uint32_t bl = pass == 1 ? b_lo_lo : b_lo_hi;
          for (int ks = 0; ks < ksteps; ++ks)
            mma_tf32_w(d, al + (uint32_t)ks * 16u, a_hi32, bl + (uint32_t)ks * 16u, b_hi32, idesc, acc || pass > 0 || ks > 0);
        }
        mma_commit(&sy.empty[s]);
      }
      __syncwarp();
      acc = true;
    }
    if (elect_one()) {
      if (lay.flags & SIGNAL_D) mma_commit(&sy.d_ready);
      if (lay.flags & SIGNAL_AUX) mma_commit(&sy.d_aux);
    }
    __syncwarp();
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
