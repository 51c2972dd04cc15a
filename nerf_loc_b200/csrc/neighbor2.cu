// The K = 8 support-point MLP + attention of ConditionalNeRF.query (conditional_nerf/model.py:371-427), second generation:
// every dense product of the stage runs on tcgen05 with bf16x3 operands (tc_bf16.cuh).
//
//   qproj_kernel      q = W_q agg                                    one 128-sample tile per MMA          (ibrnet.py:69-90)
//   neighbor2_kernel  base_mlp (3 layers) on the 128 (sample, neighbour) rows of 16 samples, then the key and value projections
//                     of base_mlp_attn ON THE SAME ROWS (W_k pf, W_v pf), scores q_h . k_h, softmax over the 8 neighbours and the
//                     per-head context o = sum_k a_hk v_k in the epilogues; inverse-distance x confidence neighbour weights
//   attn_tail_kernel  feature = LayerNorm(W_fc o + agg), feature_agg = feature * sum_k w_k      one 128-sample tile per MMA
//
// neighbor_tc.cu folded the key / value projections onto the query side to save FLOPs (M = 16 per tile instead of 128) and paid
// for it with four 16-row GEMMs per tile on the SIMT / mma.sync path, each re-reading a 64 KB weight matrix from L2: 34 k of its
// 57 k clk per tile.  Unfolded, the projections are two more 128-row tcgen05 layers (1.5 k clk each) and the per-sample
// projections that remain (q, fc) are batched over 128 samples in the two small kernels around this one.
//
// neighbor2_kernel works on a super-tile of 32 samples = two 128-row sub-tiles with their own accumulator and A operand in
// tensor memory.  The weight tiles of a layer (64 KB) stay in the 8-stage ring while BOTH sub-tiles use them: the issuer runs
// the layer for sub-tile 0 as soon as its operand is ready and for sub-tile 1 when its operand follows, so the tensor cores
// work on one sub-tile while the compute warps run the epilogue of the other.
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_bf16.cuh"
#include "tc_pipe.cuh"

namespace nlb {
namespace nb2 {

__device__ long long g_prof_nb2[32];
#define NB2_STAMP(i) do { if (stamp) g_prof_nb2[i] = clock64(); } while (0)

__device__ __forceinline__ void fast_sincos2(float x, float& s, float& c) {
  const float k = rintf(x * 0.63661977236758134f);
  float r = fmaf(-k, 1.5707963705062866f, x);
  r = fmaf(-k, -4.3711388286737929e-8f, r);
  const int q = (int)k;
  const float sr = __sinf(r), cr = __cosf(r);
  const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
  s = (q & 2) ? -a : a;
  c = ((q + 1) & 2) ? -b : b;
}

constexpr int NS = 10;                       // weight stages: a layer's tiles stay until both sub-tiles used them, 6 more run ahead
constexpr uint32_t STG_BYTES = 16384;        // one [128 x 32] hi | lo tile
constexpr int TP = 16;                       // samples per sub-tile
constexpr int NTC = 512;                     // compute threads: one group of 8 warps per sub-tile
constexpr int LDQ = 132;
// tensor-memory map of sub-tile u (columns): accumulator | A hi | A lo (two bf16 per column)
constexpr uint32_t TM_SUB = 256, TM_D = 0, TM_AHI = 128, TM_ALO = 192;
// shared-memory map (bytes)
constexpr uint32_t STG_OFF = 0;
constexpr uint32_t Q_OFF = STG_OFF + NS * STG_BYTES;                  // q [32][132] fp32
constexpr uint32_t D_OFF = Q_OFF + 2 * TP * LDQ * 4;                  // [2][256]: squared distances | confidences per sub-tile
constexpr uint32_t W_OFF = D_OFF + 2 * 256 * 4;                       // ray_diff_fc weights (544 floats)
constexpr uint32_t IDX_OFF = W_OFF + 544 * 4;                         // int idx[2][128]
constexpr int ROW_LD = 20;                                            // floats per staged row record (16 used; 20: conflict-free float4 reads)
constexpr uint32_t ROW_OFF = IDX_OFF + 2 * 128 * 4;                   // [2][128][ROW_LD]: g0 | g1 | x y z dx dy dz d2 id of the NEXT super-tile
constexpr uint32_t SYNC_OFF = ROW_OFF + 2 * 128 * ROW_LD * 4;
constexpr uint32_t SMEM_BYTES = SYNC_OFF + 256;
static_assert(SMEM_BYTES <= 232448, "neighbor2_kernel: shared memory budget");

struct Sync {
  uint64_t full[NS], empty[NS];
  uint64_t a_ready[2], d_ready[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(Sync) <= 256, "nb2::Sync");

constexpr int N_LAYERS = 5;                  // base_mlp 0, 2, 4; key projection; value projection
__device__ __forceinline__ int layer_tiles(int l) { return l == 0 ? 3 : 4; }

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory"); }

__global__ void __launch_bounds__(NTC + 128, 1)
neighbor2_kernel(const SceneDev sc, const RenderW w, const PointSrc ps, const int64_t N, const int K,
                 const int* __restrict__ knn_idx, const float* __restrict__ knn_d2, const float* __restrict__ q_in,
                 float* __restrict__ o_out, float* __restrict__ wsum_out, float* __restrict__ weights_out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* stg = smraw + STG_OFF;
  float* sQ = reinterpret_cast<float*>(smraw + Q_OFF);
  float* sD = reinterpret_cast<float*>(smraw + D_OFF);
  float* sW = reinterpret_cast<float*>(smraw + W_OFF);
  int* sIdx = reinterpret_cast<int*>(smraw + IDX_OFF);
  float* sRow = reinterpret_cast<float*>(smraw + ROW_OFF);
  Sync& sy = *reinterpret_cast<Sync*>(smraw + SYNC_OFF);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == NTC / 32) {
    tc::tmem_alloc(&sy.tmem_slot, 512);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { tc::mbar_init(&sy.full[i], 1); tc::mbar_init(&sy.empty[i], 1); }
      for (int u = 0; u < 2; ++u) { tc::mbar_init(&sy.a_ready[u], NTC / 2); tc::mbar_init(&sy.d_ready[u], 1); }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sy.tmem_slot;
  const int64_t nst = (N + 2 * TP - 1) / (2 * TP);                              // super-tiles
  const int nmy = (int)((nst - blockIdx.x + gridDim.x - 1) / gridDim.x);        // of this CTA (grid <= nst)

  if (warp >= NTC / 32) {
   // service warpgroup (its last two warps are only there so that the register hand-over is warpgroup-aligned)
   tc::reg_dec<56>();
   if (warp == NTC / 32 + 1) {
    // ------------------------------------------------ weight producer: 19 tiles of 16 KB per super-tile ------------------------
    uint32_t empty_par = 0;
    int i = 0;
    for (int t = 0; t < nmy; ++t) {
      for (int l = 0; l < N_LAYERS; ++l) {
        const float* wl = l == 0 ? w.tb_w1b : (l == 1 ? w.tb_w2 : (l == 2 ? w.tb_w3 : (l == 3 ? w.tb_wk : w.tb_wv)));
        const unsigned char* gB = reinterpret_cast<const unsigned char*>(wl);
        const int nkt = layer_tiles(l);
        for (int kt = 0; kt < nkt; ++kt, ++i) {
          const int s = i % NS;
          if (i >= NS) {
            tc::mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
            empty_par ^= 1u << s;
          }
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&sy.full[s], STG_BYTES);
            tc::bulk_copy(stg + (size_t)s * STG_BYTES, gB + (size_t)kt * STG_BYTES, STG_BYTES, &sy.full[s]);
          }
          __syncwarp();
        }
      }
    }
   } else if (warp == NTC / 32) {
    // ------------------------------------------------ MMA issuer: bf16x3, A from tensor memory --------------------------------
    // layer l of sub-tile 0 on the layer's tiles (kept in the ring), then the same tiles for sub-tile 1 (released one by one)
    uint32_t full_par = 0, a_par = 0;
    int i = 0;
    const uint32_t stage0 = tc::smem_u32(stg);
    const uint32_t idesc = tc::idesc_bf16(128, 128);
    const uint32_t b_hi32 = tc::desc_hi(32u * 16u);
    for (int t = 0; t < nmy; ++t) {
      for (int l = 0; l < N_LAYERS; ++l) {
        const int nkt = layer_tiles(l);
        for (int u = 0; u < 2; ++u) {
          tc::mbar_wait(&sy.a_ready[u], (a_par >> u) & 1u);
          a_par ^= 1u << u;
          tc::fence_after_sync();
          const uint32_t d = tmem + u * TM_SUB + TM_D;
          for (int kt = 0; kt < nkt; ++kt) {
            const int s = (i + kt) % NS;
            if (u == 0) {
              tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u);
              full_par ^= 1u << s;
              tc::fence_after_sync();
            }
            const uint32_t b_base = stage0 + (uint32_t)s * STG_BYTES;
            const uint32_t bd_hi = tc::desc_lo(b_base, 128u), bd_lo = tc::desc_lo(b_base + STG_BYTES / 2, 128u);
            if (tc::elect_one()) {
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
                const uint32_t a = tmem + u * TM_SUB + (pass == 0 ? TM_ALO : TM_AHI) + (uint32_t)(kt * 16);
                const uint32_t bl = pass == 1 ? bd_lo : bd_hi;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  tc::mma_bf16_ts_w(d, a + (uint32_t)ks * 8u, bl + (uint32_t)ks * 16u, b_hi32, idesc, kt > 0 || pass > 0 || ks > 0);
              }
              if (u == 1) tc::mma_commit(&sy.empty[s]);
            }
            __syncwarp();
          }
          if (tc::elect_one()) tc::mma_commit(&sy.d_ready[u]);
          __syncwarp();
        }
        i += nkt;
      }
    }
   }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    // Two groups of 8 warps, group u owns sub-tile u from its geometry to its context: the epilogues of the two sub-tiles run
    // side by side (four warps per scheduler instead of two) while the issuer alternates between their accumulators.
    tc::reg_inc<104>();
    const float range = sc.far_ - sc.near_;
    const int u = warp >> 3, gtid = tid & 255;
    const int row = (warp & 3) * 32 + lane;            // TMEM lane == (sample, neighbour) row of the sub-tile
    const int half = (warp >> 2) & 1;                   // column half owned in the epilogues
    const int p = row >> 3, k = row & 7;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)u * TM_SUB;
    uint32_t d_par = 0;
    for (int i = tid; i < 64 + 16 + 432 + 27; i += NTC)
      sW[i] = i < 64 ? __ldg(w.rd1 + i) : (i < 80 ? __ldg(w.rd1_b + i - 64) : (i < 512 ? __ldg(w.rd2 + i - 80) : __ldg(w.rd2_b + i - 512)));
    asm volatile("bar.sync 1, 512;" ::: "memory");
    auto wait_d = [&]() {
      tc::mbar_wait(&sy.d_ready[u], d_par);
      d_par ^= 1u;
      tc::fence_after_sync();
    };
    auto a_ready = [&]() { tc::fence_before_sync(); tc::mbar_arrive(&sy.a_ready[u]); };

    // The (sample, neighbour) records of the NEXT super-tile are staged in shared memory by the half-0 threads in two steps, so
    // that neither the index -> geometry dependency nor the records themselves cost registers or stalls in the phases between:
    //   fetch_ids(n0)   neighbour id, sample position and direction -> a few registers (loads in flight)
    //   stage_rows()    id / position / direction -> sRow, geometry (two 16-byte pieces) and squared distance by cp.async
    struct RowPre { int id; float x, y, z, dx, dy, dz; };
    auto fetch_ids = [&](const int64_t n0t) {
      RowPre q;
      q.id = -1; q.x = q.y = q.z = q.dx = q.dy = q.dz = 0.f;
      const int64_t n = n0t + p;
      if (n < N && k < K) {
        q.id = knn_idx[n * K + k];
        if (ps.xyz) {
          q.x = ps.xyz[n * 3]; q.y = ps.xyz[n * 3 + 1]; q.z = ps.xyz[n * 3 + 2];
        } else {
          const int64_t r = n / ps.S;
          const float t = ps.z[r * ps.zs + (n - r * ps.S)];
          q.x = __fadd_rn(ps.rays_o[r * 3 + 0], __fmul_rn(ps.rays_d[r * 3 + 0], t));
          q.y = __fadd_rn(ps.rays_o[r * 3 + 1], __fmul_rn(ps.rays_d[r * 3 + 1], t));
          q.z = __fadd_rn(ps.rays_o[r * 3 + 2], __fmul_rn(ps.rays_d[r * 3 + 2], t));
        }
        if (ps.dirs) {
          q.dx = ps.dirs[n * 3]; q.dy = ps.dirs[n * 3 + 1]; q.dz = ps.dirs[n * 3 + 2];
        } else if (ps.rays_d && !ps.xyz) {
          const int64_t r = n / ps.S;
          q.dx = ps.rays_d[r * 3]; q.dy = ps.rays_d[r * 3 + 1]; q.dz = ps.rays_d[r * 3 + 2];
        } else {  // direction=None: the nearest neighbour's own direction (model.py:391-392)
          const int id0 = knn_idx[n * K];
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id0 * 8));
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id0 * 8 + 4));
          q.dx = h0.w; q.dy = h1.x; q.dz = h1.y;
        }
      }
      return q;
    };
    auto stage_rows = [&](const RowPre& q, const int64_t n0t) {
      float* rr = sRow + (u * 128 + row) * ROW_LD;
      *reinterpret_cast<float4*>(rr + 8) = make_float4(q.x, q.y, q.z, q.dx);
      *reinterpret_cast<float4*>(rr + 12) = make_float4(q.dy, q.dz, 1.f, __int_as_float(q.id));
      if (q.id >= 0) {
        cp_async16(rr, sc.sup_geo + (size_t)q.id * 8);
        cp_async16(rr + 4, sc.sup_geo + (size_t)q.id * 8 + 4);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(rr + 14);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(knn_d2 + (n0t + p) * K + k));
      } else {
        *reinterpret_cast<float4*>(rr) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(rr + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };

    // ---- P0: per (sample, neighbour) geometry -> A operand of layer 1 (PE 63 | ray_diff_fc 27 | 0 x 6, permuted) in tensor memory.
    // Two threads per row, each with half of the work: positional-encoding octaves 0-4 / 5-9 and ray_diff_fc outputs 0-13 / 14-26
    // (K order: pack.cu::tcb_src_index); 48 values = 24 packed columns per plane and thread, written 16 values at a time.
    auto phase0 = [&]() {
      const float* rrow = sRow + (u * 128 + row) * ROW_LD;
      const float4 g0 = *reinterpret_cast<const float4*>(rrow), g1 = *reinterpret_cast<const float4*>(rrow + 4);
      const float4 r2 = *reinterpret_cast<const float4*>(rrow + 8), r3 = *reinterpret_cast<const float4*>(rrow + 12);
      const int rid = __float_as_int(r3.w);
      const bool live = rid >= 0;
      float vals[48];
      const float off[3] = {live ? __fdiv_rn(__fsub_rn(r2.x, g0.x), range) : 0.f, live ? __fdiv_rn(__fsub_rn(r2.y, g0.y), range) : 0.f,
                            live ? __fdiv_rn(__fsub_rn(r2.z, g0.z), range) : 0.f};
      if (half == 0) {
        sD[u * 256 + row] = live ? r3.z : 1.f;
        sD[u * 256 + 128 + row] = live ? g1.z : 0.f;   // confidence
        sIdx[u * 128 + row] = rid;
        vals[0] = off[0]; vals[1] = off[1]; vals[2] = off[2]; vals[3] = 0.f;
      } else {
#pragma unroll
        for (int c = 43; c < 48; ++c) vals[c] = 0.f;
      }
      // octaves 5 * half + {0, 2, 4} by range-reduced sincos, {1, 3} from their predecessors by the double-angle identities
      // (absolute error doubles: 4e-7 -> 8e-7, two orders below what the 1e-4 parity bar needs)
      const float f = half ? 32.f : 1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sn[5], co[5];
#pragma unroll
        for (int ii = 0; ii < 5; ii += 2) {
          sn[ii] = 0.f; co[ii] = 0.f;
          if (live) fast_sincos2(off[c] * (f * (float)(1 << ii)), sn[ii], co[ii]);
        }
#pragma unroll
        for (int ii = 1; ii < 5; ii += 2) {
          sn[ii] = 2.f * sn[ii - 1] * co[ii - 1];
          co[ii] = live ? fmaf(-2.f * sn[ii - 1], sn[ii - 1], 1.f) : 0.f;
        }
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          if (half == 0) { vals[4 + ii * 6 + c] = sn[ii]; vals[4 + ii * 6 + 3 + c] = co[ii]; }
          else { vals[ii * 6 + c] = sn[ii]; vals[ii * 6 + 3 + c] = co[ii]; }
        }
      }
      {
        // ray difference (model.py:396-399) and ray_diff_fc (4 -> 16 -> 27, LeakyReLU)
        const float nx = g0.w, ny = g1.x, nz = g1.y;
        const float rx = r2.w - nx, ry = r3.x - ny, rz = r3.y - nz;
        const float rn = sqrtf(rx * rx + ry * ry + rz * rz) + 1e-8f;
        const float rd[4] = {rx / rn, ry / rn, rz / rn, r2.w * nx + r3.x * ny + r3.y * nz};
        float h1[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const float4 w4 = *reinterpret_cast<const float4*>(sW + o * 4);
          h1[o] = leaky(fmaf(w4.w, rd[3], fmaf(w4.z, rd[2], fmaf(w4.y, rd[1], fmaf(w4.x, rd[0], sW[64 + o])))));
        }
#pragma unroll
        for (int oo = 0; oo < 14; ++oo) {
          const int o = half ? 14 + oo : oo;   // half 1 has 13 outputs (14..26)
          if (o < 27) {
            float a = sW[512 + o];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(sW + 80 + o * 16 + c);
              a = fmaf(w4.x, h1[c], a); a = fmaf(w4.y, h1[c + 1], a); a = fmaf(w4.z, h1[c + 2], a); a = fmaf(w4.w, h1[c + 3], a);
            }
            const float v = live ? leaky(a) : 0.f;
            if (half == 0) vals[34 + oo] = v; else vals[30 + oo] = v;
          }
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < 3; ++c8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::split_bf16x2(vals[c8 * 16 + 2 * j], vals[c8 * 16 + 2 * j + 1], hi[j], lo[j]);
        tc::tmem_st8_u(trow + TM_AHI + (uint32_t)(half * 24 + c8 * 8), hi);
        tc::tmem_st8_u(trow + TM_ALO + (uint32_t)(half * 24 + c8 * 8), lo);
      }
      tc::tmem_st_wait();
    };

    // ---- epilogue of a base_mlp layer: accumulator (+ per-frame support part or bias) -> LeakyReLU -> A operand of the next layer
    // MODE 0: layer 1 (adds the gathered sup_pre row), 1: layer 2 (bias b2), 2: layer 3 (bias b3).  16 columns at a time; the
    // addend of the first 32 columns is requested before the accumulator is waited for (layer 1: an L2 round trip underneath the
    // MMA), each quarter's registers are refilled with the addend 32 columns further as soon as they are consumed.
    auto mlp_epilogue = [&](const int mode) {
      const int c0 = half * 64;
      const int id = sIdx[u * 128 + row];
      const float* add = mode == 0 ? (id >= 0 ? sc.sup_pre + (size_t)id * W_HID : nullptr) : (mode == 1 ? w.b2 : w.b3);
      float4 a4[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a4[j] = add ? __ldg(reinterpret_cast<const float4*>(add + c0 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      wait_d();
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        float v[16];
        tc::tmem_ld16(trow + TM_D + (uint32_t)(c0 + cc), v);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = a4[cc / 4 + j];
          tc::split_bf16x2(leaky(v[4 * j] + b4.x), leaky(v[4 * j + 1] + b4.y), hi[2 * j], lo[2 * j]);
          tc::split_bf16x2(leaky(v[4 * j + 2] + b4.z), leaky(v[4 * j + 3] + b4.w), hi[2 * j + 1], lo[2 * j + 1]);
        }
        tc::tmem_st8_u(trow + TM_AHI + (uint32_t)((c0 + cc) / 2), hi);
        tc::tmem_st8_u(trow + TM_ALO + (uint32_t)((c0 + cc) / 2), lo);
      }
      tc::tmem_st_wait();
    };

    float prob[2];   // per head of this thread's column half
    // ---- scores q_h . k_h / sqrt(d_k), softmax over the K neighbours (8 consecutive lanes) --------------------------------------
    auto scores = [&]() {
      const float* qrow = sQ + (u * TP + p) * LDQ + half * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 32; cc += 16) {
          float v[16];
          tc::tmem_ld16(trow + TM_D + (uint32_t)(half * 64 + h * 32 + cc), v);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(qrow + h * 32 + cc + j);
            fma2_v(a0, a1, q4.x, q4.y, v[j], v[j + 1]);
            fma2_v(a0, a1, q4.z, q4.w, v[j + 2], v[j + 3]);
          }
        }
        const float a = k < K ? (a0 + a1) * 0.17677669529663687f : -FLT_MAX;   // 1 / sqrt(d_k = 32)
        float m = a;
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        const float e = k < K ? expf(a - m) : 0.f;
        float ssum = e;
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
        prob[h] = e / ssum;
      }
    };

    // ---- context o = sum_k a_hk v_k, one head (32 columns) at a time: reduce-scatter over the 8 lanes of a sample, lane k ends
    // with columns 4 k .. 4 k + 3 of the head; neighbour weights ---------------------------------------------------------------------
    auto context = [&](const int64_t n0t) {
      const int64_t n = n0t + p;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[32];
        tc::tmem_ld32(trow + TM_D + (uint32_t)(half * 64 + h * 32), v);
        const float a = prob[h];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= a;
        // of its 2 w values a lane keeps [0, w) if bit (w / 4) of k is clear, [w, 2 w) otherwise, and adds what the partner lane
        // (which keeps the other half) sends for the same positions
#pragma unroll
        for (int w2 = 16; w2 >= 4; w2 >>= 1) {
          const bool up = (k & (w2 >> 2)) != 0;
#pragma unroll
          for (int j = 0; j < w2; ++j) {
            const float send = up ? v[j] : v[j + w2];
            const float keep = up ? v[j + w2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w2 >> 2);
          }
        }
        if (n < N) __stcs(reinterpret_cast<float4*>(o_out + n * W_HID + half * 64 + h * 32 + k * 4), make_float4(v[0], v[1], v[2], v[3]));
      }
      if (half == 0) {
        // weights = (1/clamp(dist)) * softmax_K(corr) * conf, normalised (model.py:415-426); the rows of `feature` are identical
        // across K, so softmax_K(corr) is exactly 1/K and feature_agg = feature * sum_k w_k
        float wv = 0.f;
        if (k < K) {
          wv = 1.f / fmaxf(sqrtf(sD[u * 256 + row]), 1e-8f);
          wv *= 1.f / (float)K;
          wv *= sD[u * 256 + 128 + row];
        }
        float ssum = wv;
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
        const float wk = wv / fmaxf(ssum, 1e-8f);
        float wt = wk;
        wt += __shfl_xor_sync(0xffffffffu, wt, 1);
        wt += __shfl_xor_sync(0xffffffffu, wt, 2);
        wt += __shfl_xor_sync(0xffffffffu, wt, 4);
        if (n < N) {
          if (k == 0) wsum_out[n] = wt;
          if (weights_out && k < K) weights_out[n * K + k] = wk;
        }
      }
    };

    if (half == 0) stage_rows(fetch_ids((int64_t)blockIdx.x * 2 * TP + u * TP), (int64_t)blockIdx.x * 2 * TP + u * TP);
    cp_async_commit();
    cp_async_wait<0>();
    group_sync(u);
    for (int it = 0; it < nmy; ++it) {
      const bool stamp = it == 1 && blockIdx.x == gridDim.x / 2 && tid == 0;
      const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 2 * TP + u * TP;   // first sample of this group's sub-tile
      NB2_STAMP(0);
      // q rows of the group's 16 samples (written by fc_tail / qproj): asynchronous copy, consumed by the score phase
      for (int i = gtid; i < TP * 32; i += 256) {
        const int pp = i >> 5, c4 = i & 31;
        if (n0 + pp < N) cp_async16(sQ + (u * TP + pp) * LDQ + c4 * 4, q_in + pm128_off(n0 + pp, c4 * 4));
        else *reinterpret_cast<float4*>(sQ + (u * TP + pp) * LDQ + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_commit();
      phase0();
      a_ready();
      NB2_STAMP(1);
      group_sync(u);   // sIdx / sD of this sub-tile visible; sRow consumed
      const bool more = it + 1 < nmy;
      const int64_t n0n = ((int64_t)blockIdx.x + (int64_t)(it + 1) * gridDim.x) * 2 * TP + u * TP;
      RowPre pre;
      if (more && half == 0) pre = fetch_ids(n0n);
      NB2_STAMP(2);
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        mlp_epilogue(l);   // (waits for the accumulator itself)
        a_ready();
        if (l == 0) {
          if (more && half == 0) stage_rows(pre, n0n);
          cp_async_commit();
        }
        NB2_STAMP(3 + l);
      }
      cp_async_wait<0>();
      group_sync(u);   // q rows and the next records landed
      wait_d();
      NB2_STAMP(6);
      scores();
      a_ready();   // the key accumulator may be overwritten by the value projection
      NB2_STAMP(7);
      wait_d();
      NB2_STAMP(8);
      context(n0);
      NB2_STAMP(9);
      group_sync(u);   // sIdx / sD / sQ are rewritten by the next super-tile
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == NTC / 32) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

// ---- row GEMM over 128-sample tiles: out = A W^T with a [128 x 128] weight resident in shared memory ---------------------------
//   MODE 0 (qproj):  q = W_q agg
//   MODE 1 (tail):   feature = LayerNorm_eps1e-6(W_fc o + agg); feature_agg = feature * wsum   (ibrnet.py:104-119, model.py:426)
constexpr int RG_RA = 128;
constexpr uint32_t RG_W_OFF = 0;                                   // 4 K-tiles of [hi | lo] [128 x 32]: 64 KB
constexpr uint32_t RG_A_OFF = 65536;                               // A hi | lo, chunk-major, 128 rows: 64 KB
constexpr uint32_t RG_RED_OFF = RG_A_OFF + 65536;                  // [2][128] row partial sums
constexpr uint32_t RG_SYNC_OFF = RG_RED_OFF + 1024;
constexpr uint32_t RG_SMEM = RG_SYNC_OFF + 64;

template <int MODE>
__global__ void __launch_bounds__(NT, 1)
row_gemm128_kernel(const float* __restrict__ A, const float* __restrict__ wpacked, const int64_t N, const float* __restrict__ resid,
                   const float* __restrict__ ln_g, const float* __restrict__ ln_b, const float* __restrict__ wsum,
                   float* __restrict__ out, float* __restrict__ out_feature, unsigned char* __restrict__ out_split, const int S_split,
                   const bool resid_pm) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sWt = smraw + RG_W_OFF;
  unsigned char* aHi = smraw + RG_A_OFF;
  unsigned char* aLo = aHi + 32768;
  float* sRed = reinterpret_cast<float*>(smraw + RG_RED_OFF);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smraw + RG_SYNC_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smraw + RG_SYNC_OFF + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(tmem_slot, 128);
  if (tid == 32) tc::mbar_init(mbar, 1);
  for (int i = tid; i < 65536 / 16; i += NT) reinterpret_cast<uint4*>(sWt)[i] = __ldg(reinterpret_cast<const uint4*>(wpacked) + i);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int64_t ntiles = (N + 127) / 128;
  uint32_t par = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t n0 = t * 128;
    // A rows -> chunk-major bf16 hi | lo.  Four lanes share a row (one 8-column chunk = 32 bytes each: a 128-byte run), eight
    // rows per warp and instruction; 8 items in flight per thread.  (One lane per row, 512 bytes apart, was 32 separate lines
    // per load instruction.)
    {
      float4 a[8][2];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (u >> 2) * 64 + (tid >> 2), c = (u & 3) * 4 + (tid & 3);
        a[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); a[u][1] = a[u][0];
        if (n0 + r < N) {
          const float4* pp = reinterpret_cast<const float4*>(A + (n0 + r) * W_HID + c * 8);
          a[u][0] = __ldcs(pp); a[u][1] = __ldcs(pp + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (u >> 2) * 64 + (tid >> 2), c = (u & 3) * 4 + (tid & 3);
        const float v[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
        const uint32_t o = (uint32_t)c * RG_RA * 16u + (uint32_t)r * 16u;
        tc::split_store8(aHi + o, aLo + o, v);
      }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) {
      if (tc::elect_one()) {
        const uint32_t idesc = tc::idesc_bf16(128, 128);
        const uint32_t a_hi32 = tc::desc_hi(128u), b_hi32 = tc::desc_hi(32u * 16u);
        const uint32_t lbo_a = RG_RA * 16u;
        bool acc = false;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t ab = tc::smem_u32(pass == 0 ? aLo : aHi);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t bb = tc::smem_u32(sWt) + (uint32_t)(ks >> 1) * 16384u + (pass == 1 ? 8192u : 0u) + (uint32_t)(ks & 1) * 256u;
            tc::mma_bf16_w(tmem, tc::desc_lo(ab + (uint32_t)ks * 2u * lbo_a, lbo_a), a_hi32, tc::desc_lo(bb, 128u), b_hi32, idesc, acc);
            acc = true;
          }
        }
        tc::mma_commit(mbar);
      }
      __syncwarp();
    }
    tc::mbar_wait(mbar, par);
    par ^= 1u;
    tc::fence_after_sync();
    const int64_t n = n0 + row;
    const int c0 = half * 64;
    if (MODE == 0) {
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tc::tmem_ld32(trow + (uint32_t)(c0 + cc), v);
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + pm128_off(n, c0 + cc + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    } else {
      float v[64];
      float sum = 0.f;
      // residual row: 4-float pieces `rstep` floats apart (row-major: 4; piece-major, pm128_off: 128)
      const float* rbase = resid + (resid_pm ? (n >> 5) * 4096 + (n & 31) * 4 : n * W_HID);
      const int64_t rstep = resid_pm ? 128 : 4;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float t32[32];
        tc::tmem_ld32(trow + (uint32_t)(c0 + cc), t32);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 r4 = n < N ? __ldcs(reinterpret_cast<const float4*>(rbase + (int64_t)((c0 + cc + j) >> 2) * rstep)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[cc + j] = t32[j] + r4.x; v[cc + j + 1] = t32[j + 1] + r4.y; v[cc + j + 2] = t32[j + 2] + r4.z; v[cc + j + 3] = t32[j + 3] + r4.w;
          sum += (v[cc + j] + v[cc + j + 1]) + (v[cc + j + 2] + v[cc + j + 3]);
        }
      }
      sRed[half * 128 + row] = sum;
      __syncthreads();
      const float mean = (sRed[row] + sRed[128 + row]) * (1.f / 128.f);
      __syncthreads();
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) { const float d = v[j] - mean; q += d * d; }
      sRed[half * 128 + row] = q;
      __syncthreads();
      const float rstd = 1.f / sqrtf((sRed[row] + sRed[128 + row]) * (1.f / 128.f) + 1e-6f);
      const float ws = n < N ? wsum[n] : 0.f;
      if (n < N) {
        // pre-split copy for the pair ray kernel: sample s of ray r -> plane | chunk | row s, 16 bytes per (plane, chunk)
        unsigned char* sp = nullptr;
        if (out_split) {
          const int64_t r = n / S_split;
          sp = out_split + (size_t)r * S_split * 512 + (size_t)(n - r * S_split) * 16;
        }
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          float o8[8];
#pragma unroll
          for (int h4 = 0; h4 < 8; h4 += 4) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(ln_g + c0 + j + h4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + j + h4));
            float4 f;
            f.x = (v[j + h4] - mean) * rstd * g4.x + b4.x; f.y = (v[j + h4 + 1] - mean) * rstd * g4.y + b4.y;
            f.z = (v[j + h4 + 2] - mean) * rstd * g4.z + b4.z; f.w = (v[j + h4 + 3] - mean) * rstd * g4.w + b4.w;
            if (out_feature) *reinterpret_cast<float4*>(out_feature + n * W_HID + c0 + j + h4) = f;
            o8[h4] = f.x * ws; o8[h4 + 1] = f.y * ws; o8[h4 + 2] = f.z * ws; o8[h4 + 3] = f.w * ws;
            if (out) __stcs(reinterpret_cast<float4*>(out + n * W_HID + c0 + j + h4), make_float4(o8[h4], o8[h4 + 1], o8[h4 + 2], o8[h4 + 3]));
          }
          if (sp) {
            const size_t co = (size_t)((c0 + j) >> 3) * S_split * 16;
            tc::split_store8(sp + co, sp + (size_t)S_split * 256 + co, o8);
          }
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();   // accumulator, A tile and sRed are reused by the next tile
    tc::fence_after_sync();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 128);
  }
}

}  // namespace nb2

int read_prof_nb2(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, nb2::g_prof_nb2, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : set_error("read_prof_nb2 failed");
}

size_t neighbor2_scratch_floats(int64_t N) { return pm128_floats(N) + (size_t)N * W_HID + (size_t)((N + 63) / 64 * 64); }

// scratch: q (piece-major, whole groups of 32 rows) | o [N][128] | wsum [N]  (neighbor2_scratch_floats(N) floats)
int launch_neighbor2(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, int K, const int* idx,
                     const float* d2, const float* agg, float* fagg, unsigned char* fagg_split, int S_split, float* feature,
                     float* weights, float* scratch, bool q_ready, cudaStream_t st, bool agg_pm) {
  if (N <= 0) return 0;
  if (K < 1 || K > 8) return set_error("neighbor: K must be in 1..8");
  if (!scratch) return set_error("neighbor: scratch is NULL");
  float* q = scratch;
  float* o = q + pm128_floats(N);
  float* wsum = o + (size_t)N * W_HID;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(nb2::row_gemm128_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb2::RG_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(nb2::row_gemm128_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb2::RG_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(nb2::neighbor2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb2::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const int64_t t128 = (N + 127) / 128;
  const unsigned g128 = (unsigned)(t128 < sms ? t128 : sms);
  if (fagg_split && (S_split < 1 || N % S_split != 0)) return set_error("neighbor: the pre-split output needs whole rays");
  if (!q_ready) {
    nb2::row_gemm128_kernel<0><<<g128, NT, nb2::RG_SMEM, st>>>(agg, w.tb_wq, N, nullptr, nullptr, nullptr, nullptr, q, nullptr, nullptr, 1, false);
    if (check_launch("qproj_kernel")) return 1;
    prof_mark("qproj");
  }
  const int64_t nst = (N + 2 * nb2::TP - 1) / (2 * nb2::TP);
  const unsigned grid = (unsigned)(nst < sms ? nst : sms);   // persistent: one CTA per SM
  nb2::neighbor2_kernel<<<grid, nb2::NTC + 128, nb2::SMEM_BYTES, st>>>(sc, w, ps, N, K, idx, d2, q, o, wsum, weights);
  if (check_launch("neighbor2_kernel")) return 1;
  prof_mark("neighbor2");
  nb2::row_gemm128_kernel<1><<<g128, NT, nb2::RG_SMEM, st>>>(o, w.tb_wfc, N, agg, w.ln_g, w.ln_b, wsum, fagg, feature, fagg_split,
                                                             S_split < 1 ? 1 : S_split, agg_pm);
  if (check_launch("attn_tail_kernel")) return 1;
  prof_mark("attn_tail");
  return 0;
}

}  // namespace nlb
