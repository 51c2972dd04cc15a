// neighbor_kernel: the K=8 support-point MLP + attention of ConditionalNeRF.query (conditional_nerf/model.py:371-427).
//
// A tile is 16 samples = 128 (sample, neighbour) rows; one persistent CTA per SM walks the tiles.  The three 128-row layers of
// base_mlp run on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split operands, accumulator AND A operand
// in tensor memory): warps 0-7 produce the A operand / run the epilogues, warp 9 streams the pre-split weight tiles with bulk
// copies, warp 8 issues the MMAs.  The small per-sample pieces (query / folded key+value projections, softmax over the 8
// neighbours, output projection, LayerNorm, neighbour weights) stay on fp32 FFMA2 GEMMs and are software-pipelined against
// the tensor-core layers of the NEXT tile.
//
// Exact rewrites used here: layer 1 = gathered per-frame `sup_pre` + W1[:,195:285] [PE | ray_diff_fc];
// attention with a broadcast query (K/V projections folded, one distinct `feature` row per sample).
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_pipe.cuh"

namespace nlb {

// phase timestamps of block 0 (debug aid, read with nlb_debug_read_prof)
__device__ long long g_prof[32];
#define NLB_STAMP(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_prof[i] = clock64(); } while (0)

// sin / cos with a two-term Cody-Waite reduction to [-pi/4, pi/4] and the hardware approximations there: absolute error
// ~4e-7 (the library sincosf costs ~45 instructions per call; the positional encoding needs 60 per row).
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
  const float k = rintf(x * 0.63661977236758134f);
  float r = fmaf(-k, 1.5707963705062866f, x);
  r = fmaf(-k, -4.3711388286737929e-8f, r);
  const int q = (int)k;
  const float sr = __sinf(r), cr = __cosf(r);
  const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
  s = (q & 2) ? -a : a;
  c = ((q + 1) & 2) ? -b : b;
}

constexpr int NB_LDH = 132;
constexpr int NB_TP = 16;                      // samples per tile
// tensor-memory map (columns): accumulator | A operand hi | A operand lo
constexpr uint32_t NB_TM_D = 0, NB_TM_AHI = 128, NB_TM_ALO = 256;
// shared-memory map (bytes)
constexpr uint32_t NB_STG = 0;                                   // 4 x 16 KB weight stages
constexpr uint32_t NB_PF = NB_STG + 4 * 16384;                   // pf [128][132] fp32 (layer-3 output of the previous tile)
constexpr uint32_t NB_QT = NB_PF + 128 * NB_LDH * 4;             // q~ / ctx [64][132]
constexpr uint32_t NB_RED = NB_QT + 64 * NB_LDH * 4;             // small-M GEMM K-split scratch (32 KB)
constexpr uint32_t NB_SMALL = NB_RED + 32768;                    // sAgg, sQ, sO [16][132] | sSc [512] | sD [2][256] | sW [544]
constexpr uint32_t NB_SMALL_BYTES = (3 * NB_TP * NB_LDH + 512 + 512 + 544) * 4;
constexpr uint32_t NB_IDX = NB_SMALL + NB_SMALL_BYTES;           // int idx[128]
constexpr uint32_t NB_SYNC = NB_IDX + 512;                       // tc::Sync
constexpr uint32_t NB_SMEM_BYTES = NB_SYNC + 256;
static_assert(NB_SMEM_BYTES <= 232448, "neighbor_kernel: shared memory budget");

// Persistent, software-pipelined: a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The A operand of the three
// tensor-core layers lives in TENSOR MEMORY (tcgen05.mma [d], [a], b-desc; hi at columns 128.., lo at 256..): the epilogue of
// layer l reads the accumulator with tcgen05.ld and writes the split activations back with tcgen05.st, so no operand ever
// goes through shared memory and the 128 KB that used to hold it carry the attention state of the PREVIOUS tile instead.
// While the tensor cores run layer l of tile t the compute warps do the per-sample attention pieces of tile t-1 (and the
// query projections of tile t), so the MMA / weight-streaming latency is hidden behind SIMT work:
//   slot t:  P0(t) | A: scores, softmax, ctx (t-1) | E1(t) | B: value + output projections (t-1) | E2(t)
//            | C: LayerNorm, weights, output (t-1); q, q~ (t) | E3(t): pf(t) -> shared memory
__global__ void __launch_bounds__(NT + 64, 1)
neighbor_kernel(const SceneDev sc, const RenderW w, const PointSrc ps, const int64_t N, const int K,
                const int* __restrict__ knn_idx, const float* __restrict__ knn_d2, const float* __restrict__ agg_in,
                float* __restrict__ fagg_out, float* __restrict__ feature_out, float* __restrict__ weights_out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* stg = smraw + NB_STG;
  float* sA = reinterpret_cast<float*>(smraw + NB_PF);             // pf [128][132]
  float* sQT = reinterpret_cast<float*>(smraw + NB_QT);            // [64][132]
  float* sB = reinterpret_cast<float*>(smraw + NB_RED);            // K-split scratch of rows16_gemm
  float* sAgg = reinterpret_cast<float*>(smraw + NB_SMALL);
  float* sQ = sAgg + NB_TP * NB_LDH;
  float* sO = sQ + NB_TP * NB_LDH;
  float* sSc = sO + NB_TP * NB_LDH;
  float* sD = sSc + 512;                                           // [2][256]: squared distances | confidences, per tile parity
  float* sW = sD + 512;                                            // ray_diff_fc weights: rd1 [16][4] | b1 [16] | rd2 [27][16] | b2 [27]
  int* sIdx = reinterpret_cast<int*>(smraw + NB_IDX);
  tc::Sync& sy = *reinterpret_cast<tc::Sync*>(smraw + NB_SYNC);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tmem = tc::setup(sy, warp, lane, 512);
  const int64_t ntiles = (N + NB_TP - 1) / NB_TP;
  const int nmy = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (grid <= ntiles)

  if (warp == 9) {
    // ------------------------------------------------ weight producer: 22 tiles of 16 KB per sample tile ---------------------
    uint32_t empty_par = 0;
    int i = 0;
    for (int t = 0; t < nmy; ++t) {
      for (int l = 0; l < 3; ++l) {
        const unsigned char* gB = reinterpret_cast<const unsigned char*>(l == 0 ? w.tc_w1b : (l == 1 ? w.tc_w2 : w.tc_w3));
        const int nkt = l == 0 ? 6 : 8;
        for (int kt = 0; kt < nkt; ++kt, ++i) {
          const int s = i % tc::NSTAGE;
          if (i >= tc::NSTAGE) {
            tc::mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
            empty_par ^= 1u << s;
          }
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&sy.full[s], tc::STAGE_BYTES);
            tc::bulk_copy(stg + (size_t)s * tc::STAGE_BYTES, gB + (size_t)kt * tc::STAGE_BYTES, tc::STAGE_BYTES, &sy.full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer: 3xTF32, A from tensor memory -------------------------------
    uint32_t full_par = 0, a_par = 0;
    int i = 0;
    const uint32_t stage0 = tc::smem_u32(stg);
    const uint32_t idesc = tc::idesc_tf32(128, 128);
    const uint32_t b_hi32 = (((uint32_t)tc::KTB * 32u >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t lbo = (128u >> 4) << 16;
    for (int t = 0; t < nmy; ++t) {
      for (int l = 0; l < 3; ++l) {
        const int nkt = l == 0 ? 6 : 8;
        tc::mbar_wait(&sy.a_ready, a_par);
        a_par ^= 1u;
        tc::fence_after_sync();
        for (int kt = 0; kt < nkt; ++kt, ++i) {
          const int s = i % tc::NSTAGE;
          tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u);
          full_par ^= 1u << s;
          tc::fence_after_sync();
          const uint32_t b_base = stage0 + (uint32_t)s * tc::STAGE_BYTES;
          const uint32_t b_lo_hi = ((b_base & 0x3FFFFu) >> 4) | lbo;
          const uint32_t b_lo_lo = (((b_base + tc::STAGE_BYTES / 2) & 0x3FFFFu) >> 4) | lbo;
          if (tc::elect_one()) {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
              const uint32_t a = tmem + (pass == 0 ? NB_TM_ALO : NB_TM_AHI) + (uint32_t)(kt * tc::KTB);
              const uint32_t bl = pass == 1 ? b_lo_lo : b_lo_hi;
#pragma unroll
              for (int ks = 0; ks < tc::KTB / 8; ++ks)
                tc::mma_tf32_ts_w(tmem + NB_TM_D, a + (uint32_t)ks * 8u, bl + (uint32_t)ks * 16u, b_hi32, idesc, kt > 0 || pass > 0 || ks > 0);
            }
            tc::mma_commit(&sy.empty[s]);
          }
          __syncwarp();
        }
        if (tc::elect_one()) tc::mma_commit(&sy.d_ready);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    const float range = sc.far_ - sc.near_;
    const int row = (warp & 3) * 32 + lane;            // TMEM lane == (sample, neighbour) row
    const int half = warp >> 2;                         // column half owned in the TMEM epilogues / phase 0
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t d_par = 0;
    for (int i = tid; i < 64 + 16 + 432 + 27; i += NT)
      sW[i] = i < 64 ? __ldg(w.rd1 + i) : (i < 80 ? __ldg(w.rd1_b + i - 64) : (i < 512 ? __ldg(w.rd2 + i - 80) : __ldg(w.rd2_b + i - 512)));
    cta_sync();
    // LayerNorm affine parameters of this lane's four channels (loop invariant: loaded once per CTA)
    const float4 ln_g4 = __ldg(reinterpret_cast<const float4*>(w.ln_g + lane * 4));
    const float4 ln_b4 = __ldg(reinterpret_cast<const float4*>(w.ln_b + lane * 4));
    int64_t n0p = 0;   // previous tile
    int npp = 0;
    // this thread's (sample, neighbour) record of a tile: neighbour id + geometry, sample position and viewing direction.
    // Requested one slot ahead so that the dependent index -> geometry loads are off the critical path of phase 0.
    struct RowRec { int id; bool live; float4 g0, g1; float x, y, z, dx, dy, dz; };
    auto load_row = [&](const int64_t n0t, const int npt) {
      RowRec q;
      q.id = -1; q.g0 = make_float4(0.f, 0.f, 0.f, 0.f); q.g1 = q.g0;
      q.x = q.y = q.z = q.dx = q.dy = q.dz = 0.f;
      const int p = row >> 3, k = row & 7;
      q.live = p < npt && k < K;
      if (q.live) {
        const int64_t n = n0t + p;
        q.id = knn_idx[n * K + k];
        q.g0 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)q.id * 8));
        q.g1 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)q.id * 8 + 4));
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
    RowRec nxt = load_row((int64_t)blockIdx.x * NB_TP, (int)min((int64_t)NB_TP, N - (int64_t)blockIdx.x * NB_TP));
    for (int it = 0; it <= nmy; ++it) {
      const bool cur = it < nmy, prev = it > 0;
      const bool stamp = it == 1 && blockIdx.x == gridDim.x / 2 && tid == 0;
      const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * NB_TP;
      const int np = cur ? (int)min((int64_t)NB_TP, N - n0) : 0;
      float* sDc = sD + (it & 1) * 256;
      const float* sDp = sD + ((it & 1) ^ 1) * 256;
      if (stamp) g_prof[0] = clock64();

      if (cur) {
        // ---- phase 0: per (sample, neighbour) geometry -> A operand of layer 1 (PE 63 | ray_diff_fc 27 | 0 x 6, permuted) ----
        // two threads per row, each with half of the work: positional-encoding octaves 0-4 / 5-9 and ray_diff_fc outputs
        // 0-13 / 14-26 (K order: pack.cu::tcb_src_index); the 48 columns go straight to tensor memory
        const int p = row >> 3, k = row & 7;
        const bool live = nxt.live;
        const int id = nxt.id;
        const float4 g0 = nxt.g0, g1 = nxt.g1;
        const float x = nxt.x, y = nxt.y, z = nxt.z, dx = nxt.dx, dy = nxt.dy, dz = nxt.dz;
        float vals[48];
        const float off[3] = {live ? __fdiv_rn(__fsub_rn(x, g0.x), range) : 0.f, live ? __fdiv_rn(__fsub_rn(y, g0.y), range) : 0.f,
                              live ? __fdiv_rn(__fsub_rn(z, g0.z), range) : 0.f};
        if (half == 0) {
          if (live) {
            sDc[row] = knn_d2[(n0 + p) * K + k];
            sDc[128 + row] = g1.z;  // confidence
          } else {
            sDc[row] = 1.f; sDc[128 + row] = 0.f;
          }
          sIdx[row] = id;
          vals[0] = off[0]; vals[1] = off[1]; vals[2] = off[2]; vals[3] = 0.f;
        } else {
#pragma unroll
          for (int c = 43; c < 48; ++c) vals[c] = 0.f;
        }
        // positional encoding (utils.py:5-53): octaves 5*half .. 5*half+4
        float f = half ? 32.f : 1.f;
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn = 0.f, co = 0.f;
            if (live) fast_sincos(off[c] * f, sn, co);
            if (half == 0) { vals[4 + ii * 6 + c] = sn; vals[4 + ii * 6 + 3 + c] = co; }
            else { vals[ii * 6 + c] = sn; vals[ii * 6 + 3 + c] = co; }
          }
          f *= 2.f;
        }
        {
          // ray difference (model.py:396-399) and ray_diff_fc (4 -> 16 -> 27, LeakyReLU): hidden layer on both threads,
          // output rows split between them
          const float nx = g0.w, ny = g1.x, nz = g1.y;
          const float rx = dx - nx, ry = dy - ny, rz = dz - nz;
          const float rn = sqrtf(rx * rx + ry * ry + rz * rz) + 1e-8f;
          const float rd[4] = {rx / rn, ry / rn, rz / rn, dx * nx + dy * ny + dz * nz};
          float h1[16];
#pragma unroll
          for (int o = 0; o < 16; ++o) {
            float a = sW[64 + o];
#pragma unroll
            for (int c = 0; c < 4; ++c) a = fmaf(sW[o * 4 + c], rd[c], a);
            h1[o] = leaky(a);
          }
#pragma unroll
          for (int oo = 0; oo < 14; ++oo) {
            const int o = half ? 14 + oo : oo;   // half 1 has 13 outputs (14..26)
            if (o < 27) {
              float a = sW[512 + o];
#pragma unroll
              for (int c = 0; c < 16; ++c) a = fmaf(sW[80 + o * 16 + c], h1[c], a);
              const float v = live ? leaky(a) : 0.f;
              if (half == 0) vals[34 + oo] = v; else vals[30 + oo] = v;
            }
          }
        }
#pragma unroll
        for (int c16 = 0; c16 < 3; ++c16) {
          float hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) tc::split_tf32(vals[c16 * 16 + j], hi[j], lo[j]);
          tc::tmem_st16(trow + NB_TM_AHI + (uint32_t)(half * 48 + c16 * 16), hi);
          tc::tmem_st16(trow + NB_TM_ALO + (uint32_t)(half * 48 + c16 * 16), lo);
        }
        tc::tmem_st_wait();
        tc::a_ready(sy);
      }
      cta_sync();  // sIdx / sD of this tile visible
      if (stamp) g_prof[1] = clock64();

      // per-frame support part of layer 1 for this thread's row (gathered by neighbour id, 64 columns): the first 32 columns
      // are requested here, a whole attention chunk ahead of their use in E1
      float4 spre[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) spre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cur && sIdx[row] >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) spre[j] = __ldg(reinterpret_cast<const float4*>(sc.sup_pre + (size_t)sIdx[row] * W_HID + half * 64 + j * 4));
      }

      if (prev) {
        // ---- A (tile t-1): attention scores + softmax over the K neighbours ------------------------------------------------
        // thread = (sample p, column slice sl of 8): partial dot products of the 4 q~ rows with the 8 neighbour rows over its 8
        // columns (every q~ / pf element is read from shared memory once), then a transposing butterfly over the 16 lanes of
        // the sample (16 + 8 + 4 + 2 shuffles) that leaves lane sl with the two scores (head sl >> 2, k = 2 (sl & 3), + 1)
        {
          const int p = tid >> 4, sl = tid & 15;
          float4 q[4][2];
#pragma unroll
          for (int hd = 0; hd < 4; ++hd) {
            q[hd][0] = *reinterpret_cast<const float4*>(sQT + (p * 4 + hd) * NB_LDH + sl * 8);
            q[hd][1] = *reinterpret_cast<const float4*>(sQT + (p * 4 + hd) * NB_LDH + sl * 8 + 4);
          }
          float v[32];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 k0 = *reinterpret_cast<const float4*>(sA + (p * 8 + k) * NB_LDH + sl * 8);
            const float4 k1 = *reinterpret_cast<const float4*>(sA + (p * 8 + k) * NB_LDH + sl * 8 + 4);
#pragma unroll
            for (int hd = 0; hd < 4; ++hd) {
              float a = 0.f, a1 = 0.f;
              fma2_v(a, a1, q[hd][0].x, q[hd][0].y, k0.x, k0.y);
              fma2_v(a, a1, q[hd][0].z, q[hd][0].w, k0.z, k0.w);
              fma2_v(a, a1, q[hd][1].x, q[hd][1].y, k1.x, k1.y);
              fma2_v(a, a1, q[hd][1].z, q[hd][1].w, k1.z, k1.w);
              v[hd * 8 + k] = a + a1;
            }
          }
#pragma unroll
          for (int w2 = 16; w2 >= 2; w2 >>= 1) {
            // of its 2 * w2 values a lane keeps [0, w2) if bit (w2 / 2) of its slice index is clear, [w2, 2 * w2) otherwise, and
            // adds what the partner lane (which keeps the other half) sends for the same positions
            const bool up = (sl & (w2 >> 1)) != 0;
#pragma unroll
            for (int j = 0; j < w2; ++j) {
              const float send = up ? v[j] : v[j + w2];
              const float keep = up ? v[j + w2] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w2 >> 1);
            }
          }
          const int k0i = 2 * (sl & 3);
          float a0 = v[0] * 0.17677669529663687f, a1 = v[1] * 0.17677669529663687f;  // 1/sqrt(d_k = 32)
          if (k0i >= K) a0 = -FLT_MAX;
          if (k0i + 1 >= K) a1 = -FLT_MAX;
          float m = fmaxf(a0, a1);
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          const float e0 = k0i < K ? expf(a0 - m) : 0.f, e1 = k0i + 1 < K ? expf(a1 - m) : 0.f;
          float ssum = e0 + e1;
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
          *reinterpret_cast<float2*>(sSc + 2 * tid) = make_float2(e0 / ssum, e1 / ssum);   // index p * 32 + head * 8 + k
        }
        cta_sync();
        // ---- per-head context = sum_k a_k * point_feature_k (overwrites q~) -----------------------------------------------------
        {
          // thread = (sample-head row ph, quarter q): columns q*4 + 16*j .. +3, so the 4 quarter-lanes of a row read one
          // contiguous 64-byte run per j and the rows of a warp stay on distinct banks
          const int ph = tid >> 2, q4 = (tid & 3) * 4;
          const int p = ph >> 2;
          float4 acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k = 0; k < 8; ++k) {
            const float a = sSc[ph * 8 + k];
            const float* kv = sA + (p * 8 + k) * NB_LDH + q4;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v4 = *reinterpret_cast<const float4*>(kv + 16 * j);
              fma2_s(acc[j].x, acc[j].y, a, v4.x, v4.y);
              fma2_s(acc[j].z, acc[j].w, a, v4.z, v4.w);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(sQT + ph * NB_LDH + q4 + 16 * j) = acc[j];
        }
      }
      if (stamp) g_prof[2] = clock64();

      if (cur) {
        // ---- E1 (tile t): layer-1 accumulator + per-frame support part -> LeakyReLU -> A operand of layer 2 ------------------
        tc::wait_d(sy, d_par);
        if (stamp) g_prof[3] = clock64();
        const int id = sIdx[row];
        const int c0 = half * 64;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32], lo[32];
          tc::tmem_ld32(trow + NB_TM_D + (uint32_t)(c0 + cc), v);
          float4 snx[8];
          if (cc == 0) {   // second half of the row: requested before the first half is consumed
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              snx[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (id >= 0) snx[j] = __ldg(reinterpret_cast<const float4*>(sc.sup_pre + (size_t)id * W_HID + c0 + 32 + j * 4));
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = spre[j >> 2];
            v[j] += s4.x; v[j + 1] += s4.y; v[j + 2] += s4.z; v[j + 3] += s4.w;
          }
          if (cc == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) spre[j] = snx[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) tc::split_tf32(leaky(v[j]), v[j], lo[j]);
          tc::tmem_st32(trow + NB_TM_AHI + (uint32_t)(c0 + cc), v);
          tc::tmem_st32(trow + NB_TM_ALO + (uint32_t)(c0 + cc), lo);
        }
        tc::tmem_st_wait();
        tc::a_ready(sy);
      }
      if (stamp) g_prof[4] = clock64();

      if (prev) {
        // ---- B (tile t-1): o_h = Wv_h ctx_h ; fc + residual ------------------------------------------------------------------
        cta_sync();  // ctx complete
        // (these two stay on the FFMA2 small-M GEMM: the warp-level mma.sync path shares the tensor pipe with the tcgen05
        // layer that runs underneath this chunk and measured 14.1 k vs 11.4 k clk here; it wins for the q / q~ pair below)
        // a thread's K slice of either weight matrix is one batch of 32 float2 registers: the slice of the output projection is
        // requested as soon as the FMA loop of the value projection has consumed its own, so that its L2 round trip runs
        // underneath the K-split reduction and the epilogue
        float2 wreg[32];
        rows16_load<128, 128>(w.wv, 128, wreg);
        rows16_compute<128, 16, 128>(wreg, [&](int r, int c) { return sQT + (r * 4 + (c >> 5)) * NB_LDH; }, sB,
                                     [&] { rows16_load<128, 128>(w.wfc, 128, wreg); },
                                     [&](int r, int c, float v) { sO[r * NB_LDH + c] = v; });
        cta_sync();
        rows16_compute<128, 16, 128>(wreg, [&](int r, int) { return sO + r * NB_LDH; }, sB, [] {},
                                     [&](int r, int c, float v) { sQ[r * NB_LDH + c] = v + sAgg[r * NB_LDH + c]; });
      }
      cta_sync();  // sAgg of tile t-1 is dead
      if (cur) {
        // aggregated features of this tile's 16 samples: asynchronous copy, consumed by the q GEMM in chunk C
        for (int i = tid; i < NB_TP * 32; i += NT) {
          const int p = i >> 5, c4 = i & 31;
          if (p < np) cp_async16(sAgg + p * NB_LDH + c4 * 4, agg_in + (n0 + p) * W_HID + c4 * 4);
          else *reinterpret_cast<float4*>(sAgg + p * NB_LDH + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
      }
      if (stamp) g_prof[5] = clock64();

      if (cur) {
        // ---- E2 (tile t) --------------------------------------------------------------------------------------------------------
        tc::wait_d(sy, d_par);
        if (stamp) g_prof[6] = clock64();
        const int c0 = half * 64;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32], lo[32];
          tc::tmem_ld32(trow + NB_TM_D + (uint32_t)(c0 + cc), v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.b2 + c0 + cc + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) tc::split_tf32(leaky(v[j]), v[j], lo[j]);
          tc::tmem_st32(trow + NB_TM_AHI + (uint32_t)(c0 + cc), v);
          tc::tmem_st32(trow + NB_TM_ALO + (uint32_t)(c0 + cc), lo);
        }
        tc::tmem_st_wait();
        tc::a_ready(sy);
      }
      if (stamp) g_prof[7] = clock64();

      if (prev) {
        // ---- C1 (tile t-1): LayerNorm(eps 1e-6), neighbour weights, weighted sum -------------------------------------------------
        if (tid < NB_TP * 8) {
          // weights = (1/clamp(dist)) * softmax_K(corr) * conf, normalised (model.py:415-426); corr rows are identical
          // across K, so softmax_K(corr) is exactly 1/K.  One thread per (sample, neighbour), sum over the 8 lanes of a sample.
          const int k = tid & 7;
          float v = 0.f;
          if (k < K) {
            v = 1.f / fmaxf(sqrtf(sDp[tid]), 1e-8f);
            v *= 1.f / (float)K;
            v *= sDp[128 + tid];
          }
          float ssum = v;
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
          const float wk = v / fmaxf(ssum, 1e-8f);
          sSc[tid] = wk;
          // the K rows of `feature` are identical, so feature_agg = feature * sum_k w_k: the sum is taken once per sample here
          float wt = wk;
          wt += __shfl_xor_sync(0xffffffffu, wt, 1);
          wt += __shfl_xor_sync(0xffffffffu, wt, 2);
          wt += __shfl_xor_sync(0xffffffffu, wt, 4);
          if (k == 0) sSc[128 + (tid >> 3)] = wt;
          if (weights_out && (tid >> 3) < npp && k < K) weights_out[(n0p + (tid >> 3)) * K + k] = wk;
        }
        if (stamp) g_prof[12] = clock64();
        for (int p = warp; p < NB_TP; p += NT / 32) {
          const float4 y = *reinterpret_cast<const float4*>(sQ + p * NB_LDH + lane * 4);
          const float mean = warp_sum(y.x + y.y + y.z + y.w) * (1.f / 128.f);
          const float d0 = y.x - mean, d1 = y.y - mean, d2 = y.z - mean, d3 = y.w - mean;
          const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / 128.f);
          const float rstd = 1.f / sqrtf(var + 1e-6f);
          const float4 g = ln_g4, b = ln_b4;
          float4 f;
          f.x = d0 * rstd * g.x + b.x; f.y = d1 * rstd * g.y + b.y; f.z = d2 * rstd * g.z + b.z; f.w = d3 * rstd * g.w + b.w;
          *reinterpret_cast<float4*>(sO + p * NB_LDH + lane * 4) = f;
        }
        if (stamp) g_prof[13] = clock64();
        cta_sync();
        if (stamp) g_prof[14] = clock64();
        for (int i = tid; i < npp * W_HID; i += NT) {
          const int p = i >> 7, c = i & 127;
          const float f = sO[p * NB_LDH + c];
          __stcs(fagg_out + (n0p + p) * W_HID + c, f * sSc[128 + p]);   // read once by the ray kernel, much later
          if (feature_out) feature_out[(n0p + p) * W_HID + c] = f;
        }
      }
      if (stamp) g_prof[8] = clock64();

      if (cur) {
        // ---- C2 (tile t): q = Wq agg ; q~_h = Wk_h^T q_h ------------------------------------------------------------------------
        cp_async_wait<0>();
        cta_sync();  // every thread's part of the agg tile has landed
        rows16_mma<128, 128, 1>([&](int r, int, int) { return sAgg + r * NB_LDH; }, w.wq, 128,
                                [&](int r, int c, float v, int) { sQ[r * NB_LDH + c] = v; });
        cta_sync();
        // q~[p][h][c] = sum_j q[p][32h+j] Wk[32h+j][c]: four K = 32 products sharing the 128 output columns
        rows16_mma<128, 32, 4>([&](int r, int, int h) { return sQ + r * NB_LDH + 32 * h; }, w.wk, 128,
                               [&](int r, int c, float v, int h) { sQT[(r * 4 + h) * NB_LDH + c] = v; });
      }
      if (stamp) g_prof[9] = clock64();

      if (it + 1 < nmy) {
        const int64_t n0n = ((int64_t)blockIdx.x + (int64_t)(it + 1) * gridDim.x) * NB_TP;
        nxt = load_row(n0n, (int)min((int64_t)NB_TP, N - n0n));
      }
      if (cur) {
        // ---- E3 (tile t): point features -> shared memory as plain fp32 rows (read by chunk A of the next slot) ---------------
        tc::wait_d(sy, d_par);
        if (stamp) g_prof[10] = clock64();
        const int c0 = half * 64;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32];
          tc::tmem_ld32(trow + NB_TM_D + (uint32_t)(c0 + cc), v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.b3 + c0 + cc + j));
            *reinterpret_cast<float4*>(sA + row * NB_LDH + c0 + cc + j) =
                make_float4(leaky(v[j] + b4.x), leaky(v[j + 1] + b4.y), leaky(v[j + 2] + b4.z), leaky(v[j + 3] + b4.w));
          }
        }
        tc::fence_before_sync();
      }
      cta_sync();  // pf and q~ of tile t complete; the accumulator columns may be overwritten by the next tile
      if (stamp) g_prof[11] = clock64();
      n0p = n0; npp = np;
    }
  }
  tc::teardown(sy, warp, tmem, 512);
}

int read_prof(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_prof, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : set_error("read_prof failed");
}

int launch_neighbor(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, int K, const int* idx,
                    const float* d2, const float* agg, float* fagg, float* feature, float* weights, cudaStream_t st) {
  if (N <= 0) return 0;
  if (K < 1 || K > 8) return set_error("neighbor: K must be in 1..8");
  cudaError_t e = cudaFuncSetAttribute(neighbor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NB_SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const int64_t ntiles = (N + NB_TP - 1) / NB_TP;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);   // persistent: one CTA per SM
  neighbor_kernel<<<grid, NT + 64, NB_SMEM_BYTES, st>>>(sc, w, ps, N, K, idx, d2, agg, fagg, feature, weights);
  return check_launch("neighbor_kernel");
}

}  // namespace nlb
