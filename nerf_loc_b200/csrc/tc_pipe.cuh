// Pieces shared by the warp-specialised tcgen05 kernels (ray2, neighbor2, visibility, fc_tail, s2d_tc): a CTA has NT = 256
// compute threads (warps 0-7, which write A operands and run the epilogues out of TMEM: warp w reads lanes 32 (w & 3) ..,
// columns by w >> 2) plus service warps - an MMA issuer (one elected lane issues the bf16x3 tcgen05.mma sequence and commits
// stage reuse / accumulator completion through tcgen05.commit) and a producer that streams pre-split weight tiles from L2
// into a ring of shared-memory stages with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier).
#pragma once
#include "tc_common.cuh"

namespace nlb {
namespace tc {

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

}  // namespace tc
}  // namespace nlb
