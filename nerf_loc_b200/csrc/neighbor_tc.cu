// neighbor_kernel: the K=8 support-point MLP + attention of ConditionalNeRF.query (conditional_nerf/model.py:371-427).
//
// A CTA owns 16 samples = 128 (sample, neighbour) rows.  The three 128-row layers of base_mlp run on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, 3xTF32 split operands, accumulator in TMEM) through the warp-specialised pipeline
// of tc_pipe.cuh: warps 0-7 produce the A operand / run the epilogues, warp 8 streams the pre-split weight tiles with bulk
// copies and issues the MMAs.  The small per-sample pieces (query / folded key+value projections, softmax over the 8
// neighbours, output projection, LayerNorm, neighbour weights) stay on the fp32 tile GEMM.
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
constexpr int NB_TP = 16;                      // samples per CTA
constexpr uint32_t NB_SBO1 = 96u * 32u;        // A operand of layer 1: K = 96
constexpr uint32_t NB_SBO2 = 128u * 32u;       // A operand of layers 2, 3: K = 128
// shared-memory map (bytes)
constexpr uint32_t NB_STG = 0;                                   // 4 x 16 KB weight stages | later: fp32 staging + q~/ctx
constexpr uint32_t NB_STG_BYTES = 67584;                         // >= 65536 and >= 32768 + 64*132*4
constexpr uint32_t NB_ACT = NB_STG + NB_STG_BYTES;               // A hi (64 KB) | A lo (64 KB) | later: pf [128][132] fp32
constexpr uint32_t NB_ACT_BYTES = 131072;
constexpr uint32_t NB_SMALL = NB_ACT + NB_ACT_BYTES;             // sAgg, sQ, sO [16][132], sSc [512], sD [256]
constexpr uint32_t NB_SMALL_BYTES = (3 * NB_TP * NB_LDH + 512 + 256) * 4;
constexpr uint32_t NB_IDX = NB_SMALL + NB_SMALL_BYTES;           // int idx[128]
constexpr uint32_t NB_SYNC = NB_IDX + 512;                       // tc::Sync + tc::Layer[3]
constexpr uint32_t NB_SMEM_BYTES = NB_SYNC + 256 + 3 * 64;

__global__ void __launch_bounds__(NT + 64, 1)
neighbor_kernel(const SceneDev sc, const RenderW w, const PointSrc ps, const int64_t N, const int K,
                const int* __restrict__ knn_idx, const float* __restrict__ knn_d2, const float* __restrict__ agg_in,
                float* __restrict__ fagg_out, float* __restrict__ feature_out, float* __restrict__ weights_out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* stg = smraw + NB_STG;
  unsigned char* actHi = smraw + NB_ACT;
  unsigned char* actLo = actHi + 65536;
  float* sA = reinterpret_cast<float*>(smraw + NB_ACT);           // [128][132] after layer 3
  float* sB = reinterpret_cast<float*>(smraw + NB_STG);           // fp32 staging ring (after the tensor-core phase)
  float* sQT = sB + STAGE_FLOATS;                                  // [64][132]
  float* sAgg = reinterpret_cast<float*>(smraw + NB_SMALL);
  float* sQ = sAgg + NB_TP * NB_LDH;
  float* sO = sQ + NB_TP * NB_LDH;
  float* sSc = sO + NB_TP * NB_LDH;
  float* sD = sSc + 512;
  int* sIdx = reinterpret_cast<int*>(smraw + NB_IDX);
  tc::Sync& sy = *reinterpret_cast<tc::Sync*>(smraw + NB_SYNC);
  tc::Layer* layer_buf = reinterpret_cast<tc::Layer*>(smraw + NB_SYNC + 128);  // one private copy per service warp

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tmem = tc::setup(sy, warp, lane, 128);

  if (warp >= 8) {
    // ------------------------------------------------ service warps: 8 = MMA issuer, 9 = weight producer -----------------
    {
      // warp-uniform: every lane builds the same list and runs the same loop, one elected lane issues
      tc::Layer* layers = layer_buf + (warp == 9 ? 3 : 0);
      const uint32_t hi = tc::smem_u32(actHi), lo = tc::smem_u32(actLo);
      layers[0] = tc::Layer{reinterpret_cast<const unsigned char*>(w.tc_w1b), hi, lo, NB_SBO1, 6, 128, 16, 0, tc::WAIT_A | tc::SIGNAL_D};
      layers[1] = tc::Layer{reinterpret_cast<const unsigned char*>(w.tc_w2), hi, lo, NB_SBO2, 8, 128, 16, 0, tc::WAIT_A | tc::SIGNAL_D};
      layers[2] = tc::Layer{reinterpret_cast<const unsigned char*>(w.tc_w3), hi, lo, NB_SBO2, 8, 128, 16, 0, tc::WAIT_A | tc::SIGNAL_D};
      __syncwarp();
      if (warp == 8) tc::mma_issuer(sy, stg, tmem, layers, 3);
      else tc::producer(sy, stg, layers, 3);
    }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    const int64_t n0 = (int64_t)blockIdx.x * NB_TP;
    const int np = (int)min((int64_t)NB_TP, N - n0);
    const float range = sc.far_ - sc.near_;
    uint32_t d_par = 0;
    NLB_STAMP(0);
    // aggregated features of the 16 samples: asynchronous copy, consumed after the tensor-core phase
    for (int i = tid; i < NB_TP * 32; i += NT) {
      const int p = i >> 5, c4 = i & 31;
      if (p < np) cp_async16(sAgg + p * NB_LDH + c4 * 4, agg_in + (n0 + p) * W_HID + c4 * 4);
      else *reinterpret_cast<float4*>(sAgg + p * NB_LDH + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();

    // ---- phase 0: per (sample, neighbour) geometry -> A operand of layer 1 (PE 63 | ray_diff_fc 27 | 0 x 6, permuted) --
    // two threads per row, each with half of the work: positional-encoding octaves 0-4 / 5-9 and ray_diff_fc outputs
    // 0-13 / 14-26
    float* sW = sQ;  // ray_diff_fc weights: rd1 [16][4] | b1 [16] | rd2 [27][16] | b2 [27]  (sQ is free until the q GEMM)
    for (int i = tid; i < 64 + 16 + 432 + 27; i += NT)
      sW[i] = i < 64 ? __ldg(w.rd1 + i) : (i < 80 ? __ldg(w.rd1_b + i - 64) : (i < 512 ? __ldg(w.rd2 + i - 80) : __ldg(w.rd2_b + i - 512)));
    {
      const int rowp = tid & 127, half = tid >> 7;
      const int p = rowp >> 3, k = rowp & 7;
      const bool live = p < np && k < K;
      int id = -1;
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
      if (live) {
        const int64_t n = n0 + p;
        id = knn_idx[n * K + k];
        g0 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id * 8));
        g1 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id * 8 + 4));
        if (ps.xyz) {
          x = ps.xyz[n * 3]; y = ps.xyz[n * 3 + 1]; z = ps.xyz[n * 3 + 2];
        } else {
          const int64_t r = n / ps.S;
          const float t = ps.z[r * ps.zs + (n - r * ps.S)];
          x = __fadd_rn(ps.rays_o[r * 3 + 0], __fmul_rn(ps.rays_d[r * 3 + 0], t));
          y = __fadd_rn(ps.rays_o[r * 3 + 1], __fmul_rn(ps.rays_d[r * 3 + 1], t));
          z = __fadd_rn(ps.rays_o[r * 3 + 2], __fmul_rn(ps.rays_d[r * 3 + 2], t));
        }
        if (ps.dirs) {
          dx = ps.dirs[n * 3]; dy = ps.dirs[n * 3 + 1]; dz = ps.dirs[n * 3 + 2];
        } else if (ps.rays_d && !ps.xyz) {
          const int64_t r = n / ps.S;
          dx = ps.rays_d[r * 3]; dy = ps.rays_d[r * 3 + 1]; dz = ps.rays_d[r * 3 + 2];
        } else {  // direction=None: the nearest neighbour's own direction (model.py:391-392)
          const int id0 = knn_idx[n * K];
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id0 * 8));
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(sc.sup_geo + (size_t)id0 * 8 + 4));
          dx = h0.w; dy = h1.x; dz = h1.y;
        }
      }
      cta_sync();  // sW loaded
      // this thread's 48 consecutive columns of the layer-1 operand (K order: see pack.cu::tcb_src_index), stored as 12
      // conflict-free float4s (8 consecutive rows of a core-matrix column are one 128-byte run)
      float vals[48];
      const float off[3] = {live ? __fdiv_rn(__fsub_rn(x, g0.x), range) : 0.f, live ? __fdiv_rn(__fsub_rn(y, g0.y), range) : 0.f,
                            live ? __fdiv_rn(__fsub_rn(z, g0.z), range) : 0.f};
      if (half == 0) {
        if (live) {
          sD[rowp] = knn_d2[(n0 + p) * K + k];
          sD[128 + rowp] = g1.z;  // confidence
        } else {
          sD[rowp] = 1.f; sD[128 + rowp] = 0.f;
        }
        sIdx[rowp] = id;
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
      for (int c4 = 0; c4 < 12; ++c4)
        tc::store_split4(actHi, actLo, rowp, half * 48 + c4 * 4, NB_SBO1, vals[c4 * 4], vals[c4 * 4 + 1], vals[c4 * 4 + 2],
                         vals[c4 * 4 + 3]);
    }
    cp_async_wait<0>();  // the agg tile requested at kernel start
    tc::a_ready(sy);
    NLB_STAMP(1);

    // ---- base_mlp on the tensor cores: three 128 x 128 layers, epilogues out of TMEM -------------------------------------
    const int row = (warp & 3) * 32 + lane;            // TMEM lane == (sample, neighbour) row
    const int c0 = (warp >> 2) * 64;                    // this thread's 64 output columns
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    cta_sync();  // sIdx visible
    for (int layer = 0; layer < 3; ++layer) {
      tc::wait_d(sy, d_par);
      NLB_STAMP(2 + 2 * layer);
      const int id = sIdx[row];
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float v[32];
        tc::tmem_ld32(trow + (uint32_t)(c0 + cc), v);
        if (layer == 0) {
          // + per-frame precomputed support part of layer 1 (includes the bias)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (id >= 0) s4 = __ldg(reinterpret_cast<const float4*>(sc.sup_pre + (size_t)id * W_HID + c0 + cc + j));
            v[j] += s4.x; v[j + 1] += s4.y; v[j + 2] += s4.z; v[j + 3] += s4.w;
          }
        } else {
          const float* bias = layer == 1 ? w.b2 : w.b3;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + cc + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = leaky(v[j]);
        if (layer < 2) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            tc::store_split4(actHi, actLo, row, c0 + cc + j, NB_SBO2, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          // every thread must have finished reading TMEM / the MMAs are done: pf goes out as plain fp32 rows
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(sA + row * NB_LDH + c0 + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      if (layer < 2) tc::a_ready(sy);
      NLB_STAMP(3 + 2 * layer);
    }
    tc::fence_before_sync();
    cta_sync();

    NLB_STAMP(8);
    // ---- q = Wq agg ; q~_h = Wk_h^T q_h ------------------------------------------------------------------------------------
    rows16_gemm<128>([&](int r, int) { return sAgg + r * NB_LDH; }, w.wq, 128, 128, sB,
                     [&](int r, int c, float v) { sQ[r * NB_LDH + c] = v; });
    cta_sync();
    {
      // q~[p][h][c] = sum_j q[p][32h+j] Wk[32h+j][c]: thread = (column c, head pair); all 64 weight loads of a thread are
      // independent, no K split and no reduction
      const int c = tid & 127, hp = tid >> 7;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int hd = hp * 2 + hh;
        float acc[16][2];
#pragma unroll
        for (int r = 0; r < 16; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
        const float* wp = w.wk + (size_t)(32 * hd) * 128 + c;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          float b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = __ldg(wp + (size_t)(k + j) * 128);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float* ap = sQ + r * NB_LDH + 32 * hd + k;
            const float4 a0 = *reinterpret_cast<const float4*>(ap);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + 4);
            fma2_v(acc[r][0], acc[r][1], a0.x, a0.y, b[0], b[1]);
            fma2_v(acc[r][0], acc[r][1], a0.z, a0.w, b[2], b[3]);
            fma2_v(acc[r][0], acc[r][1], a1.x, a1.y, b[4], b[5]);
            fma2_v(acc[r][0], acc[r][1], a1.z, a1.w, b[6], b[7]);
          }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) sQT[(r * 4 + hd) * NB_LDH + c] = acc[r][0] + acc[r][1];
      }
    }
    cta_sync();
    NLB_STAMP(9);
    // ---- attention scores + softmax over the K neighbours -----------------------------------------------------------------
    for (int it = 0; it < 2; ++it) {
      const int i = tid + it * NT;  // (p, h, k), k fastest
      const int p = i >> 5, hd = (i >> 3) & 3, k = i & 7;
      const float* qv = sQT + (p * 4 + hd) * NB_LDH;
      const float* kv = sA + (p * 8 + k) * NB_LDH;
      float a = 0.f, a1 = 0.f;
#pragma unroll 8
      for (int c = 0; c < 128; c += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qv + c);
        const float4 k4 = *reinterpret_cast<const float4*>(kv + c);
        fma2_v(a, a1, q4.x, q4.y, k4.x, k4.y);
        fma2_v(a, a1, q4.z, q4.w, k4.z, k4.w);
      }
      a = (a + a1) * 0.17677669529663687f;  // 1/sqrt(d_k = 32)
      if (k >= K) a = -FLT_MAX;
      float m = a;
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      const float e = k < K ? expf(a - m) : 0.f;
      float s = e;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      sSc[i] = e / s;
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
    NLB_STAMP(10);
    // ---- o_h = Wv_h ctx_h ; fc + residual -------------------------------------------------------------------------------------
    cta_sync();
    rows16_gemm<128>([&](int r, int c) { return sQT + (r * 4 + (c >> 5)) * NB_LDH; }, w.wv, 128, 128, sB,
                     [&](int r, int c, float v) { sO[r * NB_LDH + c] = v; });
    cta_sync();
    rows16_gemm<128>([&](int r, int) { return sO + r * NB_LDH; }, w.wfc, 128, 128, sB,
                     [&](int r, int c, float v) { sQ[r * NB_LDH + c] = v + sAgg[r * NB_LDH + c]; });
    cta_sync();
    NLB_STAMP(11);
    // ---- LayerNorm(eps 1e-6), neighbour weights, weighted sum -------------------------------------------------------------------
    if (tid < NB_TP) {
      // weights = (1/clamp(dist)) * softmax_K(corr) * conf, normalised (model.py:415-426); corr rows are identical
      // across K, so softmax_K(corr) is exactly 1/K.
      float wk[8], s = 0.f;
      const float corr = 1.f / (float)K;
      for (int k = 0; k < 8; ++k) {
        float v = 0.f;
        if (k < K) {
          v = 1.f / fmaxf(sqrtf(sD[tid * 8 + k]), 1e-8f);
          v *= corr;
          v *= sD[128 + tid * 8 + k];
        }
        wk[k] = v; s += v;
      }
      s = fmaxf(s, 1e-8f);
      for (int k = 0; k < 8; ++k) {
        wk[k] = wk[k] / s;
        sSc[tid * 8 + k] = wk[k];
        if (weights_out && tid < np && k < K) weights_out[(n0 + tid) * K + k] = wk[k];
      }
    }
    for (int p = warp; p < NB_TP; p += NT / 32) {
      const float4 y = *reinterpret_cast<const float4*>(sQ + p * NB_LDH + lane * 4);
      const float mean = warp_sum(y.x + y.y + y.z + y.w) * (1.f / 128.f);
      const float d0 = y.x - mean, d1 = y.y - mean, d2 = y.z - mean, d3 = y.w - mean;
      const float var = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) * (1.f / 128.f);
      const float rstd = 1.f / sqrtf(var + 1e-6f);
      const float4 g = __ldg(reinterpret_cast<const float4*>(w.ln_g + lane * 4));
      const float4 b = __ldg(reinterpret_cast<const float4*>(w.ln_b + lane * 4));
      float4 f;
      f.x = d0 * rstd * g.x + b.x; f.y = d1 * rstd * g.y + b.y; f.z = d2 * rstd * g.z + b.z; f.w = d3 * rstd * g.w + b.w;
      *reinterpret_cast<float4*>(sO + p * NB_LDH + lane * 4) = f;
    }
    cta_sync();
    for (int i = tid; i < np * W_HID; i += NT) {
      const int p = i >> 7, c = i & 127;
      const float f = sO[p * NB_LDH + c];
      float a = 0.f;
      for (int k = 0; k < K; ++k) a += f * sSc[p * 8 + k];
      fagg_out[(n0 + p) * W_HID + c] = a;
      if (feature_out) feature_out[(n0 + p) * W_HID + c] = f;
    }
  }
  NLB_STAMP(12);
  tc::teardown(sy, warp, tmem, 128);
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
  const unsigned grid = (unsigned)((N + NB_TP - 1) / NB_TP);
  neighbor_kernel<<<grid, NT + 64, NB_SMEM_BYTES, st>>>(sc, w, ps, N, K, idx, d2, agg, fagg, feature, weights);
  return check_launch("neighbor_kernel");
}

}  // namespace nlb
