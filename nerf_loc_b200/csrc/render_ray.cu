// Per-ray stage of ConditionalNeRF.render_rays (conditional_nerf/model.py:521-598), one CTA per ray (sm_100a):
//   colour blend over views (model.py:528-538) -> RayUnet along the ray (conditional_nerf/ray_unet.py:5-69) ->
//   softplus density (model.py:525) -> alpha compositing, depth, depth variance, validity mask (model.py:541-575) ->
//   rendered 192-d feature (model.py:594-598).
//
// The S <= 128 samples of a ray are the 128 rows of a tcgen05 tile, so every convolution of the RayUnet runs on the tensor
// cores (3xTF32, accumulators in TMEM) through the warp-specialised pipeline of tc_pipe.cuh:
//   * a k=3 Conv1d is three GEMMs, one per tap, on the UNSHIFTED input, into three TMEM column ranges; the epilogue adds
//     them with the row shift on the OUTPUT side (out[s] = Y0[s-1] + Y1[s] + Y2[s+1]) - a row shift cannot be expressed in
//     a UMMA shared-memory descriptor, a lane shift in the epilogue is a warp shuffle;
//   * a stride-2 ConvTranspose1d is three GEMMs as well: even[j] = Y1[j], odd[j] = Y2[j] + Y0[j+1];
//   * activations live in shared memory as canonical K-major hi / lo tiles; skip connections are column ranges of the same
//     tile (c1|x1, c2|x0) or a second K chunk accumulated into the same TMEM columns (x | x2 for conv_out);
//   * the joint LayerNorm([C, S_level]) of every block is a CTA-wide two-pass reduction over an fp32 copy of the block
//     output (encoder: scratch; decoder: in place in the destination tile), conv_out keeps its rows in registers.
// Layers with fewer than 128 valid rows still issue M = 128 MMAs; the extra rows read whatever lies behind the tile (always
// inside this CTA's shared memory) and their outputs are never used.
//
// Exact rewrite: feat = sum_s w_s (W2 h_s + b2) = W2 (sum_s w_s h_s) + b2 sum_s w_s, so the second feat_mlp layer
// runs once per ray instead of once per sample.
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_pipe.cuh"

namespace nlb {

__device__ long long g_prof_ray[32];
#define RAY_STAMP(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_prof_ray[i] = clock64(); } while (0)

constexpr int RY_NS = 3;  // weight stages
// ---- shared-memory map (bytes) ---------------------------------------------------------------------------------------
constexpr uint32_t RY_X_HI = 0, RY_X_LO = 65536;            // x     [128 x 128]  SBO 4096
constexpr uint32_t RY_B2_HI = 0, RY_B2_LO = 32768;          // c2|x0 [ 32 x 256]  SBO 8192
constexpr uint32_t RY_B1_HI = 65536, RY_B1_LO = 98304;      // c1|x1 [ 64 x 128]  SBO 4096
constexpr uint32_t RY_X2_HI = 131072, RY_X2_LO = 147456;    // x2    [128 x  32]  SBO 1024
constexpr uint32_t RY_B3_HI = 131072, RY_B3_LO = 139264;    // c3    [ 16 x 128]  SBO 4096
constexpr uint32_t RY_RAW = 131072;                         // fp32 scratch of conv1 / conv2 outputs (32 KB)
constexpr uint32_t RY_RAW3 = 147456;                        // fp32 scratch of conv3 output (16 KB)
constexpr uint32_t RY_BLEND = 131072;                       // blend scratch: sBl [S][36] | logits [S*V] | small weights
constexpr uint32_t RY_STG = 163840;                         // RY_NS x 16 KB weight stages
constexpr uint32_t RY_MISC = RY_STG + RY_NS * 16384;        // small arrays, barriers, layer list
constexpr uint32_t RY_MISC_FLOATS = 512 + 512 + 64 + 256 + 256 + 256;
constexpr uint32_t RY_SYNC = RY_MISC + RY_MISC_FLOATS * 4;
constexpr uint32_t RY_SMEM_BYTES = RY_SYNC + 128 + 2 * 28 * 40;  // two layer lists, sizeof(tc::Layer) == 40
constexpr uint32_t SBO128 = 4096, SBO256 = 8192, SBO32 = 1024;

struct RayCtx {
  unsigned char* sm;
  float *xchU, *xchD, *red;
  int tid, lane, warp, wq, half, row;
  uint32_t trow;  // TMEM address of this thread's lane quarter
};

// ---- TMEM column chunk loads ---------------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void tld(uint32_t addr, float (&v)[CH]);
template <>
__device__ __forceinline__ void tld<32>(uint32_t addr, float (&v)[32]) { tc::tmem_ld32(addr, v); }
template <>
__device__ __forceinline__ void tld<16>(uint32_t addr, float (&v)[16]) { tc::tmem_ld16(addr, v); }

// One chunk of CH output columns of a 3-tap block for this thread's row: y1 + (row above of y0) + (row below of y2).
// `below_only`: transposed conv odd rows (y0 term absent).  Rows at or beyond L contribute nothing to their neighbours.
template <int CH>
__device__ __forceinline__ void shift_combine(const RayCtx& c, float (&y0)[CH], float (&y1)[CH], float (&y2)[CH], const int L,
                                              const bool use_up, float (&out)[CH]) {
  if (use_up && c.lane == 31) {
#pragma unroll
    for (int j = 0; j < CH; ++j) c.xchU[c.warp * 32 + j] = y0[j];
  }
  if (c.lane == 0) {
#pragma unroll
    for (int j = 0; j < CH; ++j) c.xchD[c.warp * 32 + j] = y2[j];
  }
  cta_sync();
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    float up = __shfl_up_sync(0xffffffffu, y0[j], 1);
    float dn = __shfl_down_sync(0xffffffffu, y2[j], 1);
    if (c.lane == 0) up = c.wq > 0 ? c.xchU[(c.warp - 1) * 32 + j] : 0.f;
    if (c.lane == 31) dn = c.wq < 3 ? c.xchD[(c.warp + 1) * 32 + j] : 0.f;
    if (c.row + 1 >= L) dn = 0.f;
    if (!use_up || c.row == 0) up = 0.f;
    out[j] = up + y1[j] + dn;
  }
  cta_sync();  // exchange buffers are reused by the next chunk
}

// Encoder block epilogue: conv output (3 taps, bias) -> fp32 scratch raw[L][COUT]; LayerNorm([COUT, L]) + ELU + MaxPool(2)
// -> destination tile rows [0, L/2), columns [0, COUT) as hi / lo.
template <int COUT>
__device__ __forceinline__ void enc_epilogue(const RayCtx& c, const uint32_t tmem, const int L, const UnetLayer& U, float* raw,
                                             unsigned char* dHi, unsigned char* dLo, const uint32_t dsbo) {
  constexpr int NC = COUT / 2, CH = NC < 32 ? NC : 32;
  // raw[row][COUT] without padding (it has to fit the 32 KB scratch); the float4 group index is XOR-swizzled with the row so
  // that the 32 rows a warp writes at once do not all fall on the same banks
  auto raw4 = [&](int r, int g4) { return reinterpret_cast<float4*>(raw + r * COUT + ((g4 ^ (r & 7)) << 2)); };
  const int c0 = c.half * NC;
  float sum = 0.f;
#pragma unroll 1
  for (int cc = 0; cc < NC; cc += CH) {
    float y0[CH], y1[CH], y2[CH], o[CH];
    tld<CH>(c.trow + tmem + 0 * COUT + c0 + cc, y0);
    tld<CH>(c.trow + tmem + 1 * COUT + c0 + cc, y1);
    tld<CH>(c.trow + tmem + 2 * COUT + c0 + cc, y2);
    shift_combine<CH>(c, y0, y1, y2, L, true, o);
    if (c.row < L) {
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.b + c0 + cc + j));
        const float4 v = make_float4(o[j] + b4.x, o[j + 1] + b4.y, o[j + 2] + b4.z, o[j + 3] + b4.w);
        sum += (v.x + v.y) + (v.z + v.w);
        *raw4(c.row, (c0 + cc + j) >> 2) = v;
      }
    }
  }
  const float n = (float)(L * COUT);
  const float mean = block_sum(sum, c.red) / n;  // (block_sum's barriers also publish `raw`)
  float q = 0.f;
  for (int i = c.tid; i < L * (COUT / 4); i += NT) {
    const int r = i / (COUT / 4), c4 = i % (COUT / 4);
    const float4 v = *raw4(r, c4);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = 1.f / sqrtf(block_sum(q, c.red) / n + 1e-5f);
  for (int i = c.tid; i < (L / 2) * (COUT / 4); i += NT) {
    const int r2 = i / (COUT / 4), c4 = i % (COUT / 4);
    const float4 va = *raw4(2 * r2, c4), vb = *raw4(2 * r2 + 1, c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(U.g + (2 * r2) * COUT + c4 * 4));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(U.g + (2 * r2 + 1) * COUT + c4 * 4));
    const float4 ba = __ldg(reinterpret_cast<const float4*>(U.be + (2 * r2) * COUT + c4 * 4));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(U.be + (2 * r2 + 1) * COUT + c4 * 4));
    tc::store_split4(dHi, dLo, r2, c4 * 4, dsbo,
                     fmaxf(elu((va.x - mean) * rstd * ga.x + ba.x), elu((vb.x - mean) * rstd * gb.x + bb.x)),
                     fmaxf(elu((va.y - mean) * rstd * ga.y + ba.y), elu((vb.y - mean) * rstd * gb.y + bb.y)),
                     fmaxf(elu((va.z - mean) * rstd * ga.z + ba.z), elu((vb.z - mean) * rstd * gb.z + bb.z)),
                     fmaxf(elu((va.w - mean) * rstd * ga.w + ba.w), elu((vb.w - mean) * rstd * gb.w + bb.w)));
  }
}

// Decoder block epilogue: transposed conv (operands in the order tap 1, 2, 0) -> rows 2j (even) and 2j+1 (odd) of the
// destination tile, columns [col_off, col_off + COUT); the fp32 values are parked in the hi tile, then LayerNorm + ELU and
// the hi / lo split happen in place.
template <int COUT>
__device__ __forceinline__ void dec_epilogue(const RayCtx& c, const uint32_t tmem, const int L, const UnetLayer& U,
                                             unsigned char* dHi, unsigned char* dLo, const uint32_t dsbo, const int col_off) {
  constexpr int NC = COUT / 2, CH = NC < 32 ? NC : 32;
  const int c0 = c.half * NC;
  float sum = 0.f;
#pragma unroll 1
  for (int cc = 0; cc < NC; cc += CH) {
    float ye[CH], y1[CH], y2[CH], o[CH];
    tld<CH>(c.trow + tmem + 0 * COUT + c0 + cc, ye);   // tap 1: even rows
    tld<CH>(c.trow + tmem + 1 * COUT + c0 + cc, y1);   // tap 2: odd rows, same input row
    tld<CH>(c.trow + tmem + 2 * COUT + c0 + cc, y2);   // tap 0: odd rows, next input row
    shift_combine<CH>(c, y1, y1, y2, L, false, o);
    if (c.row < L) {
#pragma unroll
      for (int j = 0; j < CH; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.b + c0 + cc + j));
        const float4 ve = make_float4(ye[j] + b4.x, ye[j + 1] + b4.y, ye[j + 2] + b4.z, ye[j + 3] + b4.w);
        const float4 vo = make_float4(o[j] + b4.x, o[j + 1] + b4.y, o[j + 2] + b4.z, o[j + 3] + b4.w);
        sum += ((ve.x + ve.y) + (ve.z + ve.w)) + ((vo.x + vo.y) + (vo.z + vo.w));
        *reinterpret_cast<float4*>(dHi + tc::a_off(2 * c.row, col_off + c0 + cc + j, dsbo)) = ve;
        *reinterpret_cast<float4*>(dHi + tc::a_off(2 * c.row + 1, col_off + c0 + cc + j, dsbo)) = vo;
      }
    }
  }
  const int R2 = 2 * L;
  const float n = (float)(R2 * COUT);
  const float mean = block_sum(sum, c.red) / n;
  float q = 0.f;
  for (int i = c.tid; i < R2 * (COUT / 4); i += NT) {
    const int r = i / (COUT / 4), c4 = i % (COUT / 4);
    const float4 v = *reinterpret_cast<const float4*>(dHi + tc::a_off(r, col_off + c4 * 4, dsbo));
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = 1.f / sqrtf(block_sum(q, c.red) / n + 1e-5f);
  for (int i = c.tid; i < R2 * (COUT / 4); i += NT) {
    const int r = i / (COUT / 4), c4 = i % (COUT / 4);
    const float4 v = *reinterpret_cast<const float4*>(dHi + tc::a_off(r, col_off + c4 * 4, dsbo));
    const float4 g = __ldg(reinterpret_cast<const float4*>(U.g + r * COUT + c4 * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(U.be + r * COUT + c4 * 4));
    tc::store_split4(dHi, dLo, r, col_off + c4 * 4, dsbo, elu((v.x - mean) * rstd * g.x + be.x), elu((v.y - mean) * rstd * g.y + be.y),
                     elu((v.z - mean) * rstd * g.z + be.z), elu((v.w - mean) * rstd * g.w + be.w));
  }
}

__device__ __forceinline__ void load_x(unsigned char* sm, const float* __restrict__ fagg, int64_t s0, int S, int tid) {
  // eight independent 16-byte loads in flight per thread before the first split / store
  for (int i0 = tid; i0 < S * 32; i0 += 8 * NT) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * NT;
      v[u] = i < S * 32 ? __ldcs(reinterpret_cast<const float4*>(fagg + (s0 + (i >> 5)) * W_HID + (i & 31) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * NT;
      if (i < S * 32) tc::store_split4(sm + RY_X_HI, sm + RY_X_LO, i >> 5, (i & 31) * 4, SBO128, v[u].x, v[u].y, v[u].z, v[u].w);
    }
  }
}

// 128 registers (not the 204 a 320-thread CTA could have): a third of the register file stays free for the KNN search of the
// next chunk, which nlb_render_rays runs underneath this kernel on a side stream
__global__ void __maxnreg__(128)
ray_kernel(const SceneDev sc, const RenderW w, const float* __restrict__ z_vals, const int64_t zs, const int S, const int white_bkgd,
           const float* __restrict__ fagg, const float* __restrict__ partial, const float* __restrict__ rgbvis,
           const unsigned char* __restrict__ nvalid, float* __restrict__ rgb_out, float* __restrict__ depth_out,
           float* __restrict__ weights_out, unsigned char* __restrict__ mask_out, float* __restrict__ unc_out,
           float* __restrict__ feat_out, float* __restrict__ sigma_dbg, const FeatPeers peers) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* misc = reinterpret_cast<float*>(sm + RY_MISC);
  float* sRGB = misc;               // [S][4]
  float* sV = sRGB + 512;           // sigma/alpha [S], T [S], weights [S], z [S]
  float* red = sV + 512;            // [64]
  float* xchU = red + 64;           // [8][32]
  float* xchD = xchU + 256;         // [8][32]
  float* sSigP = xchD + 256;        // [2][128] sigma partial dots
  float* sSig = sV, *sT = sV + 128, *sWt = sV + 256, *sZ = sV + 384;
  tc::SyncT<RY_NS>& sy = *reinterpret_cast<tc::SyncT<RY_NS>*>(sm + RY_SYNC);
  tc::Layer* layer_buf = reinterpret_cast<tc::Layer*>(sm + RY_SYNC + 128);  // one private copy per service warp

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tmem = tc::setup(sy, warp, lane, 512);
  const int64_t ray = blockIdx.x;
  const int64_t s0 = ray * S;
  const int V = sc.V;

  if (warp >= 8) {
    // ------------------------------------------------ service warps: 8 = MMA issuer, 9 = weight producer -----------------
    {
      // warp-uniform: every lane builds the same list and runs the same loop, one elected lane issues
      tc::Layer* layers = layer_buf + (warp == 9 ? 28 : 0);
      const uint32_t b = tc::smem_u32(sm);
      auto B = [](const float* p) { return reinterpret_cast<const unsigned char*>(p); };
      int n = 0;
      // GEMM list: {weights, A hi, A lo, A SBO, K tiles, N, K per tile (2048 / N), TMEM column, flags}
      // blend (feature_agg half of layer 1) and conv1 read x
      layers[n++] = tc::Layer{B(w.tc_bl1a), b + RY_X_HI, b + RY_X_LO, SBO128, 2, 32, 64, 448, tc::WAIT_A | tc::SIGNAL_D};
      for (int t = 0; t < 3; ++t)
        layers[n++] = tc::Layer{B(w.tcu[0][t]), b + RY_X_HI, b + RY_X_LO, SBO128, 4, 64, 32, (uint32_t)(64 * t), t == 2 ? tc::SIGNAL_D : 0u};
      for (int t = 0; t < 3; ++t)  // conv2 on c1 (K = 64)
        layers[n++] = tc::Layer{B(w.tcu[1][t]), b + RY_B1_HI, b + RY_B1_LO, SBO128, 4, 128, 16, (uint32_t)(128 * t),
                                (t == 0 ? tc::WAIT_A : 0u) | (t == 2 ? tc::SIGNAL_D : 0u)};
      for (int t = 0; t < 3; ++t)  // conv3 on c2 (K = 128 of the 256-wide tile)
        layers[n++] = tc::Layer{B(w.tcu[2][t]), b + RY_B2_HI, b + RY_B2_LO, SBO256, 8, 128, 16, (uint32_t)(128 * t),
                                (t == 0 ? tc::WAIT_A : 0u) | (t == 2 ? tc::SIGNAL_D : 0u)};
      for (int t = 0; t < 3; ++t)  // trans_conv3 on c3
        layers[n++] = tc::Layer{B(w.tcu[3][t]), b + RY_B3_HI, b + RY_B3_LO, SBO128, 8, 128, 16, (uint32_t)(128 * t),
                                (t == 0 ? tc::WAIT_A : 0u) | (t == 2 ? tc::SIGNAL_D : 0u)};
      for (int t = 0; t < 3; ++t)  // trans_conv2 on c2|x0 (K = 256)
        layers[n++] = tc::Layer{B(w.tcu[4][t]), b + RY_B2_HI, b + RY_B2_LO, SBO256, 8, 64, 32, (uint32_t)(64 * t),
                                (t == 0 ? tc::WAIT_A : 0u) | (t == 2 ? tc::SIGNAL_D : 0u)};
      for (int t = 0; t < 3; ++t)  // trans_conv1 on c1|x1 (K = 128)
        layers[n++] = tc::Layer{B(w.tcu[5][t]), b + RY_B1_HI, b + RY_B1_LO, SBO128, 2, 32, 64, (uint32_t)(32 * t),
                                (t == 0 ? tc::WAIT_A : 0u) | (t == 2 ? tc::SIGNAL_D : 0u)};
      for (int t = 0; t < 3; ++t) {  // conv_out on x (K = 128) and x2 (K = 32), same accumulator
        layers[n++] = tc::Layer{B(w.tcu[6][t]), b + RY_X_HI, b + RY_X_LO, SBO128, 8, 128, 16, (uint32_t)(128 * t), t == 0 ? tc::WAIT_A : 0u};
        layers[n++] = tc::Layer{B(w.tcu_x2[t]), b + RY_X2_HI, b + RY_X2_LO, SBO32, 2, 128, 16, (uint32_t)(128 * t),
                                tc::ACCUM | (t == 2 ? tc::SIGNAL_D : 0u)};
      }
      layers[n++] = tc::Layer{B(w.tc_ft1), b + RY_X_HI, b + RY_X_LO, SBO128, 8, 128, 16, 384, tc::SIGNAL_AUX};  // feat_mlp layer 1
      __syncwarp();
      if (warp == 8) tc::mma_issuer<RY_NS>(sy, sm + RY_STG, tmem, layers, n);
      else tc::producer<RY_NS>(sy, sm + RY_STG, layers, n);
    }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    RayCtx c;
    c.sm = sm; c.xchU = xchU; c.xchD = xchD; c.red = red;
    c.tid = tid; c.lane = lane; c.warp = warp; c.wq = warp & 3; c.half = warp >> 2; c.row = c.wq * 32 + lane;
    c.trow = (uint32_t)(c.wq * 32) << 16;
    uint32_t dpar = 0;

    RAY_STAMP(0);
    load_x(sm, fagg, s0, S, tid);
    if (tid < S) sZ[tid] = z_vals[ray * zs + tid];
    tc::a_ready(sy);                                                       // a#0: x

    // ---- colour blend (model.py:528-538) while conv1 runs on the tensor cores -----------------------------------------
    float* sBl = reinterpret_cast<float*>(sm + RY_BLEND);   // [S][36]
    float* sLogit = sBl + 128 * 36;                          // [S*V]
    float* sW2 = sLogit + 2048;                              // [16][32] | b2[16] | w3[16] | b3
    for (int i = tid; i < 512; i += NT) sW2[i] = __ldg(w.bl2 + i);
    if (tid < 16) { sW2[512 + tid] = __ldg(w.bl2_b + tid); sW2[528 + tid] = __ldg(w.bl3 + tid); }
    if (tid == 0) sW2[544] = __ldg(w.bl3_b);
    RAY_STAMP(1);
    tc::wait_d(sy, dpar);                                                  // d#0: blend GEMM
    {
      float v[16];
      tc::tmem_ld16(c.trow + tmem + 448 + c.half * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) sBl[c.row * 36 + c.half * 16 + j] = v[j];
    }
    cta_sync();
    // two (sample, view) items per thread at a time: both rows of `partial` (streamed from HBM, written by aggregate_kernel a
    // chunk ago) are requested together - half as many exposed round trips - and the layer-2 weights are read once per pair
    for (int i0 = tid; i0 < S * V; i0 += 2 * NT) {
      const int i1 = i0 + NT;
      const bool two = i1 < S * V;
      const int ib[2] = {i0, two ? i1 : i0};
      float4 pa[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4* pp = reinterpret_cast<const float4*>(partial + (s0 * V + ib[u]) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) pa[u][q] = __ldcs(pp + q);   // streamed once
      }
      float h1[2][32];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int s = ib[u] / V;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a = pa[u][q];
          const float4 b = *reinterpret_cast<const float4*>(sBl + s * 36 + q * 4);
          h1[u][q * 4 + 0] = leaky(a.x + b.x); h1[u][q * 4 + 1] = leaky(a.y + b.y);
          h1[u][q * 4 + 2] = leaky(a.z + b.z); h1[u][q * 4 + 3] = leaky(a.w + b.w);
        }
      }
      float logit0 = sW2[544], logit1 = logit0;
#pragma unroll 2
      for (int o = 0; o < 16; ++o) {
        // even-k / odd-k partial sums: one FFMA2 per two k, weights read as float4 (broadcast)
        float a = sW2[512 + o], a1 = 0.f, c = a, c1 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sW2 + o * 32 + k);
          fma2_v(a, a1, w4.x, w4.y, h1[0][k], h1[0][k + 1]);
          fma2_v(c, c1, w4.x, w4.y, h1[1][k], h1[1][k + 1]);
          fma2_v(a, a1, w4.z, w4.w, h1[0][k + 2], h1[0][k + 3]);
          fma2_v(c, c1, w4.z, w4.w, h1[1][k + 2], h1[1][k + 3]);
        }
        const float w3 = sW2[528 + o];
        logit0 = fmaf(w3, leaky(a + a1), logit0);
        logit1 = fmaf(w3, leaky(c + c1), logit1);
      }
      sLogit[i0] = __ldg(rgbvis + (s0 * V + i0) * 4 + 3) == 0.f ? -1e9f : logit0;
      if (two) sLogit[i1] = __ldg(rgbvis + (s0 * V + i1) * 4 + 3) == 0.f ? -1e9f : logit1;
    }
    cta_sync();
    if (tid < S) {
      float m = -FLT_MAX;
      for (int v = 0; v < V; ++v) m = fmaxf(m, sLogit[tid * V + v]);
      float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
      for (int v = 0; v < V; ++v) {
        const float e = expf(sLogit[tid * V + v] - m);
        const float4 cv = __ldg(reinterpret_cast<const float4*>(rgbvis + ((s0 + tid) * V + v) * 4));
        den += e; r += cv.x * e; g += cv.y * e; b += cv.z * e;
      }
      sRGB[tid * 4] = r / den; sRGB[tid * 4 + 1] = g / den; sRGB[tid * 4 + 2] = b / den;
    }
    cta_sync();  // the blend scratch region becomes the encoder scratch

    RAY_STAMP(2);
    // ---- RayUnet -------------------------------------------------------------------------------------------------------
    tc::wait_d(sy, dpar);                                                  // d#1: conv1
    RAY_STAMP(3);
    enc_epilogue<64>(c, tmem, S, w.u[0], reinterpret_cast<float*>(sm + RY_RAW), sm + RY_B1_HI, sm + RY_B1_LO, SBO128);
    tc::a_ready(sy);                                                       // a#1: c1
    RAY_STAMP(4);
    tc::wait_d(sy, dpar);                                                  // d#2: conv2
    RAY_STAMP(5);
    enc_epilogue<128>(c, tmem, S / 2, w.u[1], reinterpret_cast<float*>(sm + RY_RAW), sm + RY_B2_HI, sm + RY_B2_LO, SBO256);
    tc::a_ready(sy);                                                       // a#2: c2
    RAY_STAMP(6);
    tc::wait_d(sy, dpar);                                                  // d#3: conv3
    RAY_STAMP(7);
    enc_epilogue<128>(c, tmem, S / 4, w.u[2], reinterpret_cast<float*>(sm + RY_RAW3), sm + RY_B3_HI, sm + RY_B3_LO, SBO128);
    tc::a_ready(sy);                                                       // a#3: c3
    RAY_STAMP(8);
    tc::wait_d(sy, dpar);                                                  // d#4: trans_conv3 -> x0 = columns 128.. of c2|x0
    RAY_STAMP(9);
    dec_epilogue<128>(c, tmem, S / 8, w.u[3], sm + RY_B2_HI, sm + RY_B2_LO, SBO256, 128);
    tc::a_ready(sy);                                                       // a#4
    RAY_STAMP(10);
    tc::wait_d(sy, dpar);                                                  // d#5: trans_conv2 -> x1 = columns 64.. of c1|x1
    RAY_STAMP(11);
    dec_epilogue<64>(c, tmem, S / 4, w.u[4], sm + RY_B1_HI, sm + RY_B1_LO, SBO128, 64);
    tc::a_ready(sy);                                                       // a#5
    RAY_STAMP(12);
    tc::wait_d(sy, dpar);                                                  // d#6: trans_conv1 -> x2
    RAY_STAMP(13);
    dec_epilogue<32>(c, tmem, S / 2, w.u[5], sm + RY_X2_HI, sm + RY_X2_LO, SBO32, 0);
    load_x(sm, fagg, s0, S, tid);                                          // x again (its tile was recycled)
    tc::a_ready(sy);                                                       // a#6: x | x2

    // ---- conv_out + LayerNorm + ELU in registers, sigma = softplus(w . y + b) -----------------------------------------------
    RAY_STAMP(14);
    tc::wait_d(sy, dpar);                                                  // d#7: conv_out
    RAY_STAMP(15);
    {
      float v[64];
      float sum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float y0[32], y1[32], y2[32], o[32];
        tc::tmem_ld32(c.trow + tmem + 0 + c.half * 64 + cc, y0);
        tc::tmem_ld32(c.trow + tmem + 128 + c.half * 64 + cc, y1);
        tc::tmem_ld32(c.trow + tmem + 256 + c.half * 64 + cc, y2);
        shift_combine<32>(c, y0, y1, y2, S, true, o);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.u[6].b + c.half * 64 + cc + j));
          v[cc + j] = o[j] + b4.x; v[cc + j + 1] = o[j + 1] + b4.y; v[cc + j + 2] = o[j + 2] + b4.z; v[cc + j + 3] = o[j + 3] + b4.w;
          if (c.row < S) sum += (v[cc + j] + v[cc + j + 1]) + (v[cc + j + 2] + v[cc + j + 3]);
        }
      }
      const float n = (float)(S * 128);
      const float mean = block_sum(sum, red) / n;
      float q = 0.f;
      if (c.row < S)
#pragma unroll
        for (int j = 0; j < 64; ++j) { const float d = v[j] - mean; q += d * d; }
      const float rstd = 1.f / sqrtf(block_sum(q, red) / n + 1e-5f);
      float part = 0.f;
      if (c.row < S) {
        const float4* g4 = reinterpret_cast<const float4*>(w.u[6].g + c.row * 128 + c.half * 64);
        const float4* b4 = reinterpret_cast<const float4*>(w.u[6].be + c.row * 128 + c.half * 64);
        const float4* s4 = reinterpret_cast<const float4*>(w.sig_w + c.half * 64);
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 g = __ldg(g4 + (j >> 2)), be = __ldg(b4 + (j >> 2)), sw = __ldg(s4 + (j >> 2));
          part = fmaf(elu((v[j] - mean) * rstd * g.x + be.x), sw.x, part);
          part = fmaf(elu((v[j + 1] - mean) * rstd * g.y + be.y), sw.y, part);
          part = fmaf(elu((v[j + 2] - mean) * rstd * g.z + be.z), sw.z, part);
          part = fmaf(elu((v[j + 3] - mean) * rstd * g.w + be.w), sw.w, part);
        }
      }
      sSigP[c.half * 128 + c.row] = part;
    }
    cta_sync();

    RAY_STAMP(16);
    // ---- compositing (model.py:541-575) -------------------------------------------------------------------------------------
    if (tid < S) {
      const float sg = softplus(sSigP[tid] + sSigP[128 + tid] + __ldg(w.sig_b));
      if (sigma_dbg) sigma_dbg[s0 + tid] = sg;
      const float delta = tid + 1 < S ? sZ[tid + 1] - sZ[tid] : 1e2f;
      sSig[tid] = 1.f - expf(-delta * sg);  // alpha
    }
    cta_sync();
    if (tid == 0) {
      float T = 1.f;
      for (int s = 0; s < S; ++s) { sT[s] = T; T *= (1.f - sSig[s]); }
    }
    cta_sync();
    float wv = 0.f, zz = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, nv = 0.f;
    if (tid < S) {
      wv = sSig[tid] * sT[tid];
      sWt[tid] = wv;
      weights_out[s0 + tid] = wv;
      zz = sZ[tid];
      cr = sRGB[tid * 4]; cg = sRGB[tid * 4 + 1]; cb = sRGB[tid * 4 + 2];
      nv = nvalid[s0 + tid] > 1 ? 1.f : 0.f;
    }
    const float wsum = block_sum(wv, red);
    const float depth = block_sum(wv * zz, red);
    const float unc = block_sum(wv * (zz - depth) * (zz - depth), red);
    float r = block_sum(wv * cr, red), g = block_sum(wv * cg, red), b = block_sum(wv * cb, red);
    const float cnt = block_sum(nv, red);
    if (tid == 0) {
      if (white_bkgd) { r += 1.f - wsum; g += 1.f - wsum; b += 1.f - wsum; }
      rgb_out[ray * 3] = r; rgb_out[ray * 3 + 1] = g; rgb_out[ray * 3 + 2] = b;
      depth_out[ray] = depth;
      unc_out[ray] = unc;
      mask_out[ray] = cnt > 8.f ? 1 : 0;
    }

    RAY_STAMP(17);
    // ---- rendered feature (model.py:594-598) ---------------------------------------------------------------------------------
    tc::mbar_wait(&sy.d_aux, 0);                                           // feat_mlp layer 1 done (x tile is free now)
    tc::fence_after_sync();
    if (feat_out || peers.n > 0) {
      float* sPart = reinterpret_cast<float*>(sm);  // [128][132] fp32 over the dead x tile
      const float wrow = c.row < S ? sWt[c.row] : 0.f;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        float y[32];
        tc::tmem_ld32(c.trow + tmem + 384 + c.half * 64 + cc, y);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.ft1_b + c.half * 64 + cc + j));
          float4 o;
          o.x = c.row < S ? leaky(y[j] + b4.x) * wrow : 0.f;
          o.y = c.row < S ? leaky(y[j + 1] + b4.y) * wrow : 0.f;
          o.z = c.row < S ? leaky(y[j + 2] + b4.z) * wrow : 0.f;
          o.w = c.row < S ? leaky(y[j + 3] + b4.w) * wrow : 0.f;
          *reinterpret_cast<float4*>(sPart + c.row * 132 + c.half * 64 + cc + j) = o;
        }
      }
      cta_sync();
      float* sHs = sPart + 128 * 132;
      if (tid < 128) {
        float a = 0.f;
        for (int rr = 0; rr < 128; ++rr) a += sPart[rr * 132 + tid];
        sHs[tid] = a;
      }
      cta_sync();
      if (tid < C_FEAT) {
        float a = __ldg(w.ft2_b + tid) * wsum;
        for (int k = 0; k < 128; ++k) a = fmaf(__ldg(w.ft2 + k * C_FEAT + tid), sHs[k], a);
        if (feat_out) feat_out[ray * C_FEAT + tid] = a;
        // fused all-gather: the same 768-byte row goes to every rank's gathered matrix (peer stores over NVLink)
        for (int p = 0; p < peers.n; ++p) peers.p[p][(peers.row0 + ray) * C_FEAT + tid] = a;
      }
    }
  }
  RAY_STAMP(18);
  tc::teardown(sy, warp, tmem, 512);
}

int read_prof_ray(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_prof_ray, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : set_error("read_prof_ray failed");
}

int launch_ray(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t zs, int64_t R, int S, int white_bkgd,
               const float* fagg, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
               float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
               const FeatPeers& peers, cudaStream_t st) {
  if (R <= 0) return 0;
  if (S % 8 != 0 || S < 8 || S > 128) return set_error("ray stage: samples per ray must be a multiple of 8 in [8, 128]");
  if (w.S != S) return set_error("ray stage: weights were packed for a different number of samples per ray");
  if (sc.V > 16) return set_error("ray stage: at most 16 reference views");
  cudaError_t e = cudaFuncSetAttribute(ray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RY_SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  ray_kernel<<<(unsigned)R, NT + 64, RY_SMEM_BYTES, st>>>(sc, w, z_vals, zs, S, white_bkgd, fagg, partial, rgbvis, nvalid, rgb,
                                                        depth, weights, mask, depth_unc, feat, sigma_dbg, peers);
  return check_launch("ray_kernel");
}

}  // namespace nlb
