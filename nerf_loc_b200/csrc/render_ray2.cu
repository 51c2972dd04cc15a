// Per-ray stage of ConditionalNeRF.render_rays (conditional_nerf/model.py:521-598), second generation: one CTA per PAIR of
// rays, bf16x3 operands (tc_bf16.cuh) on tcgen05, S <= 128.
//
//   colour blend over views (model.py:528-538) -> RayUnet along the ray (conditional_nerf/ray_unet.py:5-69) -> softplus density
//   (model.py:525) -> alpha compositing, depth, depth variance, validity mask (model.py:541-575) -> rendered 192-d feature
//   (model.py:594-598).
//
// What changed against render_ray.cu (3xTF32, one ray per CTA, 273 k clk per ray of which 118 k were waits on a 2.27 MB weight
// stream and 109 k three-accumulator epilogues):
//   * activations are chunk-major bf16 hi | lo tiles (one 8-column chunk of all rows contiguous, rows 16 bytes apart), so the
//     three taps of a Conv1d are three descriptors on the SAME tile whose start address differs by one sample's worth of rows
//     and which accumulate into ONE TMEM accumulator: a third of the TMEM traffic, no row-shift exchange in the epilogue;
//   * a stride-2 ConvTranspose1d is two accumulators: even[j] = W1 x[j], odd[j] = W2 x[j] + W0 x[j+1];
//   * the two rays of a CTA are interleaved along M below the first level (row = sample * 2 + ray): a tile of level 2 is
//     64 samples x 2 rays = 128 rows, so the half-resolution layers fill the 128-row MMA that a single ray left half empty, a
//     "one sample" shift is two rows, and the only padding rows are before the first and after the last sample;
//   * at full resolution each ray has its own tile and every weight tile is used for two MMAs (one per ray) while it sits in
//     shared memory: 1.16 MB of bf16 hi | lo weights are streamed per PAIR instead of 2.27 MB of tf32 hi | lo per ray;
//   * LayerNorm([C, S_level]) statistics are taken from TMEM in three cheap passes (sum, squared deviations, normalise) instead
//     of going through an fp32 scratch copy in shared memory.
// A ray's results do not depend on the ray it is paired with: MMA rows are independent and every reduction has the same shape
// for either slot of the pair.
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_bf16.cuh"
#include "tc_pipe.cuh"

namespace nlb {
namespace r2 {

__device__ long long g_prof_ray2[32];
#define R2_STAMP(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_prof_ray2[i] = clock64(); } while (0)

constexpr int NS = 3;                      // weight stages
constexpr uint32_t STG_BYTES = 16384;
// rows per chunk of the activation tiles: data rows + one sample of zero rows before and after
constexpr int RA_X = 130;                  // level 1, one ray per tile: 1 + 128 + 1
constexpr int RA_L2 = 132;                 // level 2, two rays: 2 + 128 + 2
constexpr int RA_L3 = 68;                  // level 3: 2 + 64 + 2
constexpr int RA_L4 = 36;                  // level 4: 2 + 32 + 2
constexpr uint32_t plane_bytes(int channels, int RA) { return (uint32_t)(channels / 8) * (uint32_t)RA * 16u; }
// ---- shared-memory map (bytes) ---------------------------------------------------------------------------------------
// region P: the two x tiles (first and last phase); the level 3 / 4 tiles and x1 in between
constexpr uint32_t X_PLANE = plane_bytes(128, RA_X);                 // 33,280
constexpr uint32_t X_TILE = 2 * X_PLANE;                             // hi | lo
constexpr uint32_t P_BYTES = 2 * X_TILE;                             // 133,120
constexpr uint32_t C2_HI = 0, C2_LO = C2_HI + plane_bytes(128, RA_L3);
constexpr uint32_t X0_HI = C2_LO + plane_bytes(128, RA_L3), X0_LO = X0_HI + plane_bytes(128, RA_L3);
constexpr uint32_t C3_HI = X0_LO + plane_bytes(128, RA_L3), C3_LO = C3_HI + plane_bytes(128, RA_L4);
constexpr uint32_t X1_HI = C3_LO + plane_bytes(128, RA_L4), X1_LO = X1_HI + plane_bytes(64, RA_L2);
static_assert(X1_LO + plane_bytes(64, RA_L2) + 4096 <= P_BYTES, "region P");
// region Q: c1 (level 2), later the two x2 tiles; the colour-blend scratch before either exists
constexpr uint32_t Q_OFF = P_BYTES;
constexpr uint32_t C1_HI = Q_OFF, C1_LO = C1_HI + plane_bytes(64, RA_L2);
constexpr uint32_t Q_BYTES = 2 * plane_bytes(64, RA_L2);             // 33,792
constexpr uint32_t X2_PLANE = plane_bytes(32, RA_X);                 // 8,320
constexpr uint32_t X2_TILE = 2 * X2_PLANE;
static_assert(2 * X2_TILE <= Q_BYTES, "region Q");
constexpr uint32_t STG_OFF = Q_OFF + Q_BYTES;
constexpr uint32_t MISC_OFF = STG_OFF + NS * STG_BYTES;
constexpr uint32_t MISC_FLOATS = 2 * 512 + 2 * 512 + 64 + 2 * 256;   // sRGB [2][128][4] | sV [2][4][128] | red | sSigP [2][2][128]
constexpr uint32_t SYNC_OFF = MISC_OFF + MISC_FLOATS * 4;
constexpr int MAX_SEGS = 36;
constexpr int N_DBAR = 9;

struct Seg {                 // D[128 x N] (+)= A_view[128 x K] * B[N x K]^T, for one or two rays sharing the weight tiles
  const unsigned char* gB;   // packed weights: per K-tile [hi: N x ktile][lo: N x ktile] (tc_bf16.cuh weight-tile layout)
  uint32_t a_hi[2], a_lo[2]; // shared-memory addresses of the A planes, row shift already applied, per ray
  uint32_t a_lbo;            // bytes between 8-column chunks (RA * 16)
  uint32_t tmem_col[2];
  uint16_t nkt, N, ktile, nray;
  uint32_t flags;            // F_WAIT_A | F_ACCUM | (d barrier index + 1) << 8
};
constexpr uint32_t F_WAIT_A = 1u, F_ACCUM = 2u;

struct Sync2 {
  uint64_t full[NS], empty[NS];
  uint64_t a_ready;
  uint64_t d_bar[N_DBAR];
  uint64_t x_full;           // the x tiles of both rays have landed (bulk copies of the pre-split feature_agg)
  uint64_t bl2;              // a batch of colour-blend layer-2 MMAs (issued by a compute warp) has completed
  uint32_t tmem_slot;
  int nseg;
};
constexpr uint32_t SEG_OFF = SYNC_OFF + 256;
constexpr uint32_t SMEM_BYTES = SEG_OFF + MAX_SEGS * sizeof(Seg);
static_assert(sizeof(Sync2) <= 256, "Sync2");
static_assert(SMEM_BYTES <= 232448, "ray2_kernel: shared memory budget");

enum { D_BLEND = 0, D_CONV1, D_CONV2, D_CONV3, D_TCONV3, D_TCONV2, D_TCONV1, D_CONVOUT, D_FEAT };

struct Ctx {
  unsigned char* sm;
  float* red;
  int tid, lane, warp, wq, half, m;
  uint32_t trow;   // TMEM address of this warp's lane quarter (column 0)
  int S;
};

__device__ __forceinline__ void tld16(uint32_t addr, float (&v)[16]) { tc::tmem_ld16(addr, v); }
// 16 consecutive bias values (64-byte aligned: packed arrays start on 256-byte boundaries) as four 16-byte loads
__device__ __forceinline__ void ldb16(const float* __restrict__ p, float (&b)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
}

// block-wide sums of two independent quantities (slot 0 / slot 1 of the ray pair); every thread gets both
__device__ __forceinline__ float2 block_sum2(float a, float b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  cta_sync();
  if ((threadIdx.x & 31) == 0) { red[(threadIdx.x >> 5) * 2] = a; red[(threadIdx.x >> 5) * 2 + 1] = b; }
  cta_sync();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) { ta += red[2 * i]; tb += red[2 * i + 1]; }
  return make_float2(ta, tb);
}

// zero rows [r0, r0 + nr) of every chunk of both planes of a tile
__device__ __forceinline__ void zero_rows(unsigned char* hi, unsigned char* lo, int RA, int chunks, int r0, int nr, int tid) {
  for (int i = tid; i < chunks * nr * 2; i += NT) {
    const int pl = i & 1, r = (i >> 1) % nr, c = (i >> 1) / nr;
    *reinterpret_cast<uint4*>((pl ? lo : hi) + (size_t)c * RA * 16 + (size_t)(r0 + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// The x tile of a ray arrives by bulk copies (feature_agg is written pre-split by attn_tail_kernel, one [S x 8] bf16 run per
// plane and chunk): the compute warps only clear the zero rows before and after the data - or the whole tile of a missing ray.
__device__ __forceinline__ void prep_x(unsigned char* tile, int S, bool live, int tid) {
  unsigned char* hi = tile;
  unsigned char* lo = tile + X_PLANE;
  if (live) {
    zero_rows(hi, lo, RA_X, 16, 0, 1, tid);
    zero_rows(hi, lo, RA_X, 16, S + 1, 1, tid);
  } else {
    zero_rows(hi, lo, RA_X, 16, 0, S + 2, tid);
  }
}

// ---- encoder block at level 1 (conv1): per-ray accumulators [128 x 64] at TMEM columns 64 * ray -> LayerNorm([64, S]) + ELU +
// MaxPool(2) -> rows (j * 2 + ray) of the level-2 tile c1.  LayerNorm affine rows are loaded once for both rays.
__device__ __forceinline__ void epi_conv1(const Ctx& c, const uint32_t tmem, const UnetLayer& U) {
  const int S = c.S;
  const bool valid = c.m < S;
  const int c0 = c.half * 32;
  const float n = (float)(S * 64);
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int ray = 0; ray < 2; ++ray)
#pragma unroll
    for (int cc = 0; cc < 32; cc += 16) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + 64 * ray + c0 + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) sum[ray] += v[j] + bv[j];
    }
  const float2 tot = block_sum2(valid ? sum[0] : 0.f, valid ? sum[1] : 0.f, c.red);
  const float mean[2] = {tot.x / n, tot.y / n};
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int ray = 0; ray < 2; ++ray)
#pragma unroll
    for (int cc = 0; cc < 32; cc += 16) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + 64 * ray + c0 + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) { const float d = (v[j] + bv[j]) - mean[ray]; q[ray] += d * d; }
    }
  const float2 qt = block_sum2(valid ? q[0] : 0.f, valid ? q[1] : 0.f, c.red);
  const float rstd[2] = {1.f / sqrtf(qt.x / n + 1e-5f), 1.f / sqrtf(qt.y / n + 1e-5f)};
  unsigned char* dhi = c.sm + C1_HI;
  unsigned char* dlo = c.sm + C1_LO;
  const int srow = valid ? c.m : 0;
#pragma unroll
  for (int cc = 0; cc < 32; cc += 16) {
    float g[16], be[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(U.g2 + ln_off(srow, c0 + cc + j, 64)));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.be2 + ln_off(srow, c0 + cc + j, 64)));
      g[j] = g4.x; g[j + 1] = g4.y; g[j + 2] = g4.z; g[j + 3] = g4.w;
      be[j] = b4.x; be[j + 1] = b4.y; be[j + 2] = b4.z; be[j + 3] = b4.w;
    }
#pragma unroll
    for (int ray = 0; ray < 2; ++ray) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + 64 * ray + c0 + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float y = elu(((v[j] + bv[j]) - mean[ray]) * rstd[ray] * g[j] + be[j]);
        v[j] = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, 1));     // MaxPool over the sample pair (2j, 2j + 1)
      }
      // both lanes of a pair hold the pooled row: the even lane stores ray 0's, the odd lane ray 1's
      if (valid && (c.lane & 1) == ray) {
        const int drow = (c.m >> 1) * 2 + ray + 2;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const float w8[8] = {v[h8 * 8], v[h8 * 8 + 1], v[h8 * 8 + 2], v[h8 * 8 + 3], v[h8 * 8 + 4], v[h8 * 8 + 5], v[h8 * 8 + 6], v[h8 * 8 + 7]};
          const uint32_t o = tc::cm_off(drow, c0 + cc + h8 * 8, RA_L2);
          tc::split_store8(dhi + o, dlo + o, w8);
        }
      }
    }
  }
  zero_rows(dhi, dlo, RA_L2, 8, 0, 2, c.tid);
  zero_rows(dhi, dlo, RA_L2, 8, 2 + S, 2, c.tid);
}

// ---- encoder block on a pair tile (conv2, conv3): accumulator [rows = sample * 2 + ray][128] at TMEM column 0, L samples per
// ray -> LayerNorm([128, L]) per ray + ELU + MaxPool(2) -> rows ((s / 2) * 2 + ray) of the next level's tile
__device__ __forceinline__ void epi_enc_pair(const Ctx& c, const uint32_t tmem, const int L, const UnetLayer& U, unsigned char* dhi,
                                             unsigned char* dlo, const int RAd) {
  const bool valid = c.m < 2 * L;
  const int b = c.lane & 1, s = c.m >> 1;
  const int c0 = c.half * 64;
  const float n = (float)(L * 128);
  float sum = 0.f;
#pragma unroll
  for (int cc = 0; cc < 64; cc += 16) {
    float v[16];
    float bv[16];
    ldb16(U.b + c0 + cc, bv);
    tld16(c.trow + tmem + c0 + cc, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += v[j] + bv[j];
  }
  if (!valid) sum = 0.f;
  const float2 tot = block_sum2(b == 0 ? sum : 0.f, b == 1 ? sum : 0.f, c.red);
  const float mean = (b == 0 ? tot.x : tot.y) / n;
  float q = 0.f;
#pragma unroll
  for (int cc = 0; cc < 64; cc += 16) {
    float v[16];
    float bv[16];
    ldb16(U.b + c0 + cc, bv);
    tld16(c.trow + tmem + c0 + cc, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float d = (v[j] + bv[j]) - mean; q += d * d; }
  }
  if (!valid) q = 0.f;
  const float2 qt = block_sum2(b == 0 ? q : 0.f, b == 1 ? q : 0.f, c.red);
  const float rstd = 1.f / sqrtf((b == 0 ? qt.x : qt.y) / n + 1e-5f);
  const int srow = valid ? s : 0;
  const int sp = (c.lane >> 1) & 1;                 // parity of the sample: which half of a 16-column chunk this lane stores
  const int drow = (s >> 1) * 2 + b + 2;
#pragma unroll
  for (int cc = 0; cc < 64; cc += 16) {
    float v[16];
    float bv[16];
    ldb16(U.b + c0 + cc, bv);
    tld16(c.trow + tmem + c0 + cc, v);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(U.g2 + ln_off(srow, c0 + cc + j, 128)));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.be2 + ln_off(srow, c0 + cc + j, 128)));
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float y = elu(((v[j + u] + bv[j + u]) - mean) * rstd * gg[u] + bb[u]);
        v[j + u] = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, 2));   // partner row: same ray, sample s ^ 1
      }
    }
    if (valid) {
      const float w8[8] = {v[sp * 8], v[sp * 8 + 1], v[sp * 8 + 2], v[sp * 8 + 3], v[sp * 8 + 4], v[sp * 8 + 5], v[sp * 8 + 6], v[sp * 8 + 7]};
      const uint32_t o = tc::cm_off(drow, c0 + cc + sp * 8, RAd);
      tc::split_store8(dhi + o, dlo + o, w8);
    }
  }
  zero_rows(dhi, dlo, RAd, 16, 0, 2, c.tid);
  zero_rows(dhi, dlo, RAd, 16, 2 + L, 2, c.tid);   // L / 2 samples x 2 rays = L data rows
}

// ---- decoder block (stride-2 transposed conv) on a pair tile with Lin samples per ray: even accumulator at TMEM column 0, odd
// accumulator at column N; output sample 2j (+1) of ray b <- even (odd) row j * 2 + b.  LayerNorm([N, 2 Lin]) per ray + ELU.
// PER_RAY_DST: the destination is a pair of level-1 tiles (one per ray, row = sample + 1) instead of a pair tile.
template <int N, bool PER_RAY_DST>
__device__ __forceinline__ void epi_dec(const Ctx& c, const uint32_t tmem, const int Lin, const UnetLayer& U, unsigned char* dhi,
                                        unsigned char* dlo, const int RAd, const uint32_t ray_stride) {
  constexpr int NC = N / 2, CH = NC < 16 ? NC : 16;
  static_assert(CH == 16, "decoder epilogue reads 16-column chunks");
  const bool valid = c.m < 2 * Lin;
  const int b = c.lane & 1, j = c.m >> 1;
  const int c0 = c.half * NC;
  const float n = (float)(2 * Lin * N);
  float sum = 0.f;
#pragma unroll
  for (int par = 0; par < 2; ++par)
#pragma unroll
    for (int cc = 0; cc < NC; cc += 16) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + par * N + c0 + cc, v);
#pragma unroll
      for (int u = 0; u < 16; ++u) sum += v[u] + bv[u];
    }
  if (!valid) sum = 0.f;
  const float2 tot = block_sum2(b == 0 ? sum : 0.f, b == 1 ? sum : 0.f, c.red);
  const float mean = (b == 0 ? tot.x : tot.y) / n;
  float q = 0.f;
#pragma unroll
  for (int par = 0; par < 2; ++par)
#pragma unroll
    for (int cc = 0; cc < NC; cc += 16) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + par * N + c0 + cc, v);
#pragma unroll
      for (int u = 0; u < 16; ++u) { const float d = (v[u] + bv[u]) - mean; q += d * d; }
    }
  if (!valid) q = 0.f;
  const float2 qt = block_sum2(b == 0 ? q : 0.f, b == 1 ? q : 0.f, c.red);
  const float rstd = 1.f / sqrtf((b == 0 ? qt.x : qt.y) / n + 1e-5f);
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    const int so = valid ? 2 * j + par : 0;                      // output sample
    unsigned char* th = dhi + (PER_RAY_DST ? (size_t)b * ray_stride : 0);
    unsigned char* tl = dlo + (PER_RAY_DST ? (size_t)b * ray_stride : 0);
    const int drow = PER_RAY_DST ? so + 1 : so * 2 + b + 2;
#pragma unroll
    for (int cc = 0; cc < NC; cc += 16) {
      float v[16];
      float bv[16];
      ldb16(U.b + c0 + cc, bv);
      tld16(c.trow + tmem + par * N + c0 + cc, v);
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(U.g2 + ln_off(so, c0 + cc + u, N)));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.be2 + ln_off(so, c0 + cc + u, N)));
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) v[u + t] = elu(((v[u + t] + bv[u + t]) - mean) * rstd * gg[t] + bb[t]);
      }
      if (valid) {
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const float w8[8] = {v[h8 * 8], v[h8 * 8 + 1], v[h8 * 8 + 2], v[h8 * 8 + 3], v[h8 * 8 + 4], v[h8 * 8 + 5], v[h8 * 8 + 6], v[h8 * 8 + 7]};
          const uint32_t o = tc::cm_off(drow, c0 + cc + h8 * 8, RAd);
          tc::split_store8(th + o, tl + o, w8);
        }
      }
    }
  }
  if (PER_RAY_DST) {
#pragma unroll
    for (int ray = 0; ray < 2; ++ray) {
      zero_rows(dhi + ray * ray_stride, dlo + ray * ray_stride, RAd, N / 8, 0, 1, c.tid);
      zero_rows(dhi + ray * ray_stride, dlo + ray * ray_stride, RAd, N / 8, 1 + 2 * Lin, 1, c.tid);
    }
  } else {
    zero_rows(dhi, dlo, RAd, N / 8, 0, 2, c.tid);
    zero_rows(dhi, dlo, RAd, N / 8, 2 + 4 * Lin, 2, c.tid);      // 2 Lin samples x 2 rays
  }
}

__global__ void __launch_bounds__(NT + 128, 1)
ray2_kernel(const SceneDev sc, const RenderW w, const float* __restrict__ z_vals, const int64_t zs, const int S, const int64_t R,
            const int white_bkgd, const unsigned char* __restrict__ xsplit, const float* __restrict__ partial,
            const float* __restrict__ rgbvis, const unsigned char* __restrict__ nvalid, float* __restrict__ rgb_out,
            float* __restrict__ depth_out, float* __restrict__ weights_out, unsigned char* __restrict__ mask_out,
            float* __restrict__ unc_out, float* __restrict__ feat_out, float* __restrict__ sigma_dbg, const FeatPeers peers) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* misc = reinterpret_cast<float*>(sm + MISC_OFF);
  float* sRGB = misc;                 // [2][128][4]
  float* sV = sRGB + 1024;            // [2][4][128]: alpha, T, weight, z
  float* red = sV + 1024;             // [64]
  float* sSigP = red + 64;            // [2][2][128] sigma partial dots
  Sync2& sy = *reinterpret_cast<Sync2*>(sm + SYNC_OFF);
  Seg* segs = reinterpret_cast<Seg*>(sm + SEG_OFF);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ray0 = (int64_t)blockIdx.x * 2;
  const bool live1 = ray0 + 1 < R;
  const int V = sc.V;

  // ---- setup: TMEM, barriers, the GEMM list (one thread builds it while the others allocate / initialise) ----------------
  if (warp == 8) {
    tc::tmem_alloc(&sy.tmem_slot, 512);
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { tc::mbar_init(&sy.full[i], 1); tc::mbar_init(&sy.empty[i], 1); }
      tc::mbar_init(&sy.a_ready, NT);
      tc::mbar_init(&sy.x_full, 1);
      tc::mbar_init(&sy.bl2, 1);
      for (int i = 0; i < N_DBAR; ++i) tc::mbar_init(&sy.d_bar[i], 1);
    }
  } else if (warp == 9 && lane == 0) {
    const uint32_t b = tc::smem_u32(sm);
    auto B = [](const float* p) { return reinterpret_cast<const unsigned char*>(p); };
    int n = 0;
    // one tap of a layer on one source tile: `tiles` K-tiles of the packed [N x Cin] block starting at K-tile `kt0`
    auto add = [&](const float* wts, int N, int ktile, int kt0, int tiles, uint32_t hi0, uint32_t lo0, uint32_t ray_stride, int nray,
                   int RA, int shift_rows, uint32_t col0, uint32_t col_stride, uint32_t flags) {
      Seg s;
      s.gB = B(wts) + (size_t)kt0 * (4u * N * ktile);
      for (int r = 0; r < 2; ++r) {
        s.a_hi[r] = b + hi0 + r * ray_stride + (uint32_t)shift_rows * 16u;
        s.a_lo[r] = b + lo0 + r * ray_stride + (uint32_t)shift_rows * 16u;
        s.tmem_col[r] = col0 + r * col_stride;
      }
      s.a_lbo = (uint32_t)RA * 16u;
      s.nkt = (uint16_t)tiles; s.N = (uint16_t)N; s.ktile = (uint16_t)ktile; s.nray = (uint16_t)nray;
      s.flags = flags;
      segs[n++] = s;
    };
    auto sig = [](int d) { return (uint32_t)(d + 1) << 8; };
    // row offset of the view for tap t of a Conv1d (x[s-1], x[s], x[s+1]) on a tile with `bt` rows per sample: (t) * bt
    // level 1 (per-ray tiles, both rays per weight tile): blend layer-1 (feature_agg half), conv1
    add(w.tb_bl1a, 32, 128, 0, 1, 0, X_PLANE, X_TILE, 2, RA_X, 1, 448, 32, F_WAIT_A | sig(D_BLEND));
    for (int t = 0; t < 3; ++t)
      add(w.tb_u[0][t], 64, 64, 0, 2, 0, X_PLANE, X_TILE, 2, RA_X, t, 0, 64, (t ? F_ACCUM : 0u) | (t == 2 ? sig(D_CONV1) : 0u));
    // conv2 on c1
    for (int t = 0; t < 3; ++t)
      add(w.tb_u[1][t], 128, 32, 0, 2, C1_HI, C1_LO, 0, 1, RA_L2, 2 * t, 0, 0,
          (t ? F_ACCUM : F_WAIT_A) | (t == 2 ? sig(D_CONV2) : 0u));
    // conv3 on c2
    for (int t = 0; t < 3; ++t)
      add(w.tb_u[2][t], 128, 32, 0, 4, C2_HI, C2_LO, 0, 1, RA_L3, 2 * t, 0, 0,
          (t ? F_ACCUM : F_WAIT_A) | (t == 2 ? sig(D_CONV3) : 0u));
    // trans_conv3 on c3: even = W1 x[j]; odd = W2 x[j] + W0 x[j+1]
    add(w.tb_u[3][1], 128, 32, 0, 4, C3_HI, C3_LO, 0, 1, RA_L4, 2, 0, 0, F_WAIT_A);
    add(w.tb_u[3][2], 128, 32, 0, 4, C3_HI, C3_LO, 0, 1, RA_L4, 2, 128, 0, 0u);
    add(w.tb_u[3][0], 128, 32, 0, 4, C3_HI, C3_LO, 0, 1, RA_L4, 4, 128, 0, F_ACCUM | sig(D_TCONV3));
    // trans_conv2 on c2 | x0 (K-tiles 0-1: c2, 2-3: x0)
    add(w.tb_u[4][1], 64, 64, 0, 2, C2_HI, C2_LO, 0, 1, RA_L3, 2, 0, 0, F_WAIT_A);
    add(w.tb_u[4][1], 64, 64, 2, 2, X0_HI, X0_LO, 0, 1, RA_L3, 2, 0, 0, F_ACCUM);
    add(w.tb_u[4][2], 64, 64, 0, 2, C2_HI, C2_LO, 0, 1, RA_L3, 2, 64, 0, 0u);
    add(w.tb_u[4][2], 64, 64, 2, 2, X0_HI, X0_LO, 0, 1, RA_L3, 2, 64, 0, F_ACCUM);
    add(w.tb_u[4][0], 64, 64, 0, 2, C2_HI, C2_LO, 0, 1, RA_L3, 4, 64, 0, F_ACCUM);
    add(w.tb_u[4][0], 64, 64, 2, 2, X0_HI, X0_LO, 0, 1, RA_L3, 4, 64, 0, F_ACCUM | sig(D_TCONV2));
    // trans_conv1 on c1 | x1 (K-tile 0: c1, 1: x1)
    add(w.tb_u[5][1], 32, 64, 0, 1, C1_HI, C1_LO, 0, 1, RA_L2, 2, 0, 0, F_WAIT_A);
    add(w.tb_u[5][1], 32, 64, 1, 1, X1_HI, X1_LO, 0, 1, RA_L2, 2, 0, 0, F_ACCUM);
    add(w.tb_u[5][2], 32, 64, 0, 1, C1_HI, C1_LO, 0, 1, RA_L2, 2, 32, 0, 0u);
    add(w.tb_u[5][2], 32, 64, 1, 1, X1_HI, X1_LO, 0, 1, RA_L2, 2, 32, 0, F_ACCUM);
    add(w.tb_u[5][0], 32, 64, 0, 1, C1_HI, C1_LO, 0, 1, RA_L2, 4, 32, 0, F_ACCUM);
    add(w.tb_u[5][0], 32, 64, 1, 1, X1_HI, X1_LO, 0, 1, RA_L2, 4, 32, 0, F_ACCUM | sig(D_TCONV1));
    // conv_out on x | x2 (K-tiles 0-3: x, 4: x2), both rays per weight tile; then feat_mlp layer 1 on x
    for (int t = 0; t < 3; ++t) {
      add(w.tb_u[6][t], 128, 32, 0, 4, 0, X_PLANE, X_TILE, 2, RA_X, t, 0, 128, t ? F_ACCUM : F_WAIT_A);
      add(w.tb_u[6][t], 128, 32, 4, 1, Q_OFF, Q_OFF + X2_PLANE, X2_TILE, 2, RA_X, t, 0, 128, F_ACCUM | (t == 2 ? sig(D_CONVOUT) : 0u));
    }
    add(w.tb_ft1, 128, 32, 0, 4, 0, X_PLANE, X_TILE, 2, RA_X, 1, 256, 128, sig(D_FEAT));
    sy.nseg = n;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sy.tmem_slot;

  if (warp >= 8) {
   // service warpgroup: 8 = MMA issuer, 9 = weight producer, 10 = x loader (11 only keeps the register hand-over aligned)
   tc::reg_dec<56>();
   if (warp == 10) {
    // ------------------------------------------------ x loader: pre-split feature_agg -> the two x tiles, twice -----------------
    const int nlive = live1 ? 2 : 1;
    for (int round = 0; round < 2; ++round) {
      if (round == 1) { tc::mbar_wait(&sy.d_bar[D_TCONV1], 0); }     // region P is free again (trans_conv1 has consumed x1)
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&sy.x_full, (uint32_t)(nlive * S * 512));
        for (int ray = 0; ray < nlive; ++ray) {
          const unsigned char* src = xsplit + (size_t)(ray0 + ray) * S * 512;
          for (int pl = 0; pl < 2; ++pl)
            for (int cch = 0; cch < 16; ++cch)
              tc::bulk_copy(sm + ray * X_TILE + pl * X_PLANE + (uint32_t)cch * RA_X * 16u + 16u, src + (size_t)pl * S * 256 + (size_t)cch * S * 16,
                            (uint32_t)(S * 16), &sy.x_full);
        }
      }
      __syncwarp();
    }
   } else if (warp == 9) {
    // ------------------------------------------------ weight producer ---------------------------------------------------------
    uint32_t empty_par = 0;
    int i = 0;
    const int nseg = sy.nseg;
    for (int l = 0; l < nseg; ++l) {
      const Seg sg = segs[l];
      const uint32_t bytes = 4u * sg.N * sg.ktile;
      for (int kt = 0; kt < sg.nkt; ++kt, ++i) {
        const int s = i % NS;
        if (i >= NS) {
          tc::mbar_wait(&sy.empty[s], (empty_par >> s) & 1u);
          empty_par ^= 1u << s;
        }
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&sy.full[s], bytes);
          tc::bulk_copy(sm + STG_OFF + (size_t)s * STG_BYTES, sg.gB + (size_t)kt * bytes, bytes, &sy.full[s]);
        }
        __syncwarp();
      }
    }
   } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer -----------------------------------------------------------------
    uint32_t full_par = 0, a_par = 0;
    int i = 0;
    const uint32_t stage0 = tc::smem_u32(sm + STG_OFF);
    const int nseg = sy.nseg;
    const uint32_t a_hi32 = tc::desc_hi(128u);
    for (int l = 0; l < nseg; ++l) {
      const Seg sg = segs[l];
      if (sg.flags & F_WAIT_A) {
        tc::mbar_wait(&sy.a_ready, a_par);
        a_par ^= 1u;
        tc::fence_after_sync();
      }
      const uint32_t idesc = tc::idesc_bf16(128, sg.N);
      const uint32_t b_hi32 = tc::desc_hi((uint32_t)sg.ktile * 16u);
      const uint32_t half_tile = 2u * sg.N * sg.ktile;
      const uint32_t lbo_a = (sg.a_lbo >> 4) << 16;
      const uint32_t a_kstep = (2u * sg.a_lbo) >> 4;                 // 16 k = two chunks, in descriptor address units
      const int ksteps = sg.ktile / 16;
      bool acc = (sg.flags & F_ACCUM) != 0;
      for (int kt = 0; kt < sg.nkt; ++kt, ++i) {
        const int s = i % NS;
        tc::mbar_wait(&sy.full[s], (full_par >> s) & 1u);
        full_par ^= 1u << s;
        tc::fence_after_sync();
        const uint32_t b_base = stage0 + (uint32_t)s * STG_BYTES;
        const uint32_t bd_hi = tc::desc_lo(b_base, 128u), bd_lo = tc::desc_lo(b_base + half_tile, 128u);
        const uint32_t a_koff = (uint32_t)(kt * ksteps) * a_kstep;
        if (tc::elect_one()) {
          for (int r = 0; r < sg.nray; ++r) {
            const uint32_t ad_hi = (((sg.a_hi[r] & 0x3FFFFu) >> 4) | lbo_a) + a_koff;
            const uint32_t ad_lo = (((sg.a_lo[r] & 0x3FFFFu) >> 4) | lbo_a) + a_koff;
            const uint32_t d = tmem + sg.tmem_col[r];
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
              const uint32_t al = pass == 0 ? ad_lo : ad_hi;
              const uint32_t bl = pass == 1 ? bd_lo : bd_hi;
              for (int ks = 0; ks < ksteps; ++ks)
                tc::mma_bf16_w(d, al + (uint32_t)ks * a_kstep, a_hi32, bl + (uint32_t)ks * 16u, b_hi32, idesc, acc || pass > 0 || ks > 0);
            }
          }
          tc::mma_commit(&sy.empty[s]);
        }
        __syncwarp();
        acc = true;
      }
      const uint32_t d_idx = sg.flags >> 8;
      if (d_idx) {
        if (tc::elect_one()) tc::mma_commit(&sy.d_bar[d_idx - 1]);
        __syncwarp();
      }
    }
   }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    tc::reg_inc<224>();
    Ctx c;
    c.sm = sm; c.red = red; c.tid = tid; c.lane = lane; c.warp = warp; c.wq = warp & 3; c.half = warp >> 2;
    c.m = c.wq * 32 + lane; c.trow = (uint32_t)(c.wq * 32) << 16; c.S = S;
    auto wait_d = [&](int d) { tc::mbar_wait(&sy.d_bar[d], 0); tc::fence_after_sync(); };
    auto a_ready = [&]() { tc::fence_async_smem(); tc::fence_before_sync(); tc::mbar_arrive(&sy.a_ready); };
    const int64_t sbase[2] = {ray0 * S, (ray0 + 1) * S};

    R2_STAMP(0);
    prep_x(sm, S, true, tid);
    prep_x(sm + X_TILE, S, live1, tid);
    {
      const int ray = tid >> 7, s = tid & 127;
      if (s < S) sV[ray * 512 + 384 + s] = (ray == 0 || live1) ? z_vals[(ray0 + ray) * zs + s] : 0.f;
    }
    // colour blend: the (sample, view) items of a ray are rows of 128-row tiles, processed in batches of four tiles; this
    // thread's two items of the first batch are requested before anything else is waited for
    const int bl_items = S * V, bl_tiles = (bl_items + 127) / 128, bl_nb = (bl_tiles + 3) / 4;
    const int bl_steps = (live1 ? 2 : 1) * bl_nb;
    float4 pa[2][8];
    float visf[2];
    auto bl_fetch = [&](int step) {
      const int ray = step / bl_nb, t0 = (step - ray * bl_nb) * 4;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + 2 * u + (warp >> 2), i = t * 128 + (warp & 3) * 32 + lane;
        const bool on = t < bl_tiles && i < bl_items;
        visf[u] = on ? __ldg(rgbvis + (sbase[ray] * V + i) * 4 + 3) : 0.f;
        const int64_t r = sbase[ray] * V + (on ? i : 0);
#pragma unroll
        for (int q = 0; q < 8; ++q)   // streamed once; 32 lanes x one piece = 512 contiguous bytes (partial_off)
          pa[u][q] = on ? __ldcs(reinterpret_cast<const float4*>(partial + partial_off(r, q))) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    bl_fetch(0);
    tc::mbar_wait(&sy.x_full, 0);
    a_ready();                                                             // a#0: x of both rays
    R2_STAMP(1);

    // ---- colour blend (model.py:528-538), one ray after the other in the scratch of region Q -----------------------------
    float* sBl = reinterpret_cast<float*>(sm + Q_OFF);      // [S][36]
    float* sLogit = sBl + 128 * 36;                          // [S * V]
    float* sW2 = sLogit + 2048;                              // b2[16] | w3[16] | b3
    unsigned char* sW2b = reinterpret_cast<unsigned char*>(sW2 + 64);   // layer 2 [16 x 32] as a bf16 hi | lo weight tile (1 KB | 1 KB)
    if (tid < 16) { sW2[tid] = __ldg(w.bl2_b + tid); sW2[16 + tid] = __ldg(w.bl3 + tid); }
    if (tid == 0) sW2[32] = __ldg(w.bl3_b);
    {
      const int n = tid >> 4, k = (tid & 15) * 2;           // 256 threads: one pair of the 16 x 32 weights each
      uint32_t hi, lo;
      tc::split_bf16x2(__ldg(w.bl2 + n * 32 + k), __ldg(w.bl2 + n * 32 + k + 1), hi, lo);
      *reinterpret_cast<uint32_t*>(sW2b + tc::wt_off(n, k, 32)) = hi;
      *reinterpret_cast<uint32_t*>(sW2b + 1024 + tc::wt_off(n, k, 32)) = lo;
    }
    tc::fence_async_smem();
    wait_d(D_BLEND);
    R2_STAMP(2);
    // Layer 2 (32 -> 16) of the blend MLP runs on the tensor cores as well: the S * V (sample, view) items of a ray are rows of
    // 128-row tiles; a thread builds layer-1 activations of its items (per-view half streamed from HBM + per-sample half from
    // the blend GEMM), writes them as the A operand into free tensor-memory columns (128 ..), one compute warp issues the MMAs
    // of a batch of four tiles, and the 16 outputs per item come back for the 16 -> 1 head.  (On FFMA2 this layer alone took
    // 50 k of the 244 k clk per pair of rays: 2 x 1024 items x 512 FMAs on eight warps.)
    constexpr uint32_t BL_TM = 128, BL_TILE = 48;            // per tile: A hi 16 | A lo 16 | D 16 columns
    uint32_t bl_par = 0;
    for (int step = 0; step < bl_steps; ++step) {
      const int ray = step / bl_nb, t0 = (step - ray * bl_nb) * 4;
      const int64_t s0 = sbase[ray];
      if (t0 == 0) {
        float v[16];
        tc::tmem_ld16(c.trow + tmem + 448 + 32 * ray + c.half * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) sBl[c.m * 36 + c.half * 16 + j] = v[j];
        cta_sync();
      }
      if (step == 0) R2_STAMP(20);
      float vis_cur[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tl = 2 * u + c.half, t = t0 + tl, i = t * 128 + c.m;
        const bool on = t < bl_tiles && i < bl_items;
        vis_cur[u] = visf[u];
        const int s = (on ? i : 0) / V;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(sBl + s * 36 + q * 4);
          tc::split_bf16x2(leaky(pa[u][q].x + bq.x), leaky(pa[u][q].y + bq.y), hi[2 * q], lo[2 * q]);
          tc::split_bf16x2(leaky(pa[u][q].z + bq.z), leaky(pa[u][q].w + bq.w), hi[2 * q + 1], lo[2 * q + 1]);
        }
        if (t < bl_tiles) {   // warp-uniform
          tc::tmem_st16_u(c.trow + tmem + BL_TM + tl * BL_TILE, hi);
          tc::tmem_st16_u(c.trow + tmem + BL_TM + tl * BL_TILE + 16, lo);
        }
      }
      if (step == 0) R2_STAMP(25);
      // the per-view halves of the NEXT batch (HBM, written by aggregate_kernel a chunk ago: ~7 k clk when waited for in place)
      // and, at a ray's last batch, its colour rows are requested now and land underneath the MMAs and the 16 -> 1 head
      if (step + 1 < bl_steps) bl_fetch(step + 1);
      const bool last_of_ray = t0 + 4 >= bl_tiles;
      float4 cv[16];
      if (last_of_ray && tid < S) {
#pragma unroll
        for (int v = 0; v < 16; ++v)
          cv[v] = v < V ? __ldg(reinterpret_cast<const float4*>(rgbvis + ((s0 + tid) * V + v) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (step == 0) R2_STAMP(26);
      tc::tmem_st_wait();
      if (step == 0) R2_STAMP(27);
      tc::fence_before_sync();
      cta_sync();
      if (step == 0) R2_STAMP(21);
      if (warp == 0) {
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::idesc_bf16(128, 16);
          const uint32_t wb = tc::smem_u32(sW2b), b_hi32 = tc::desc_hi(32u * 16u);
          for (int tl = 0; tl < 4 && t0 + tl < bl_tiles; ++tl) {
            const uint32_t base = tmem + BL_TM + tl * BL_TILE;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
              const uint32_t a = base + (pass == 0 ? 16u : 0u);
              const uint32_t bp = wb + (pass == 1 ? 1024u : 0u);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts_w(base + 32u, a + (uint32_t)ks * 8u, tc::desc_lo(bp + (uint32_t)ks * 256u, 128u), b_hi32, idesc, pass > 0 || ks > 0);
            }
          }
          tc::mma_commit(&sy.bl2);
        }
        __syncwarp();
      }
      tc::mbar_wait(&sy.bl2, bl_par);
      bl_par ^= 1u;
      tc::fence_after_sync();
      if (step == 0) R2_STAMP(22);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tl = 2 * u + c.half, t = t0 + tl, i = t * 128 + c.m;
        if (t < bl_tiles) {   // warp-uniform
          float d[16];
          tc::tmem_ld16(c.trow + tmem + BL_TM + tl * BL_TILE + 32, d);
          float logit = sW2[32];
#pragma unroll
          for (int o = 0; o < 16; ++o) logit = fmaf(sW2[16 + o], leaky(d[o] + sW2[o]), logit);
          if (i < bl_items) sLogit[i] = vis_cur[u] == 0.f ? -1e9f : logit;
        }
      }
      tc::fence_before_sync();   // the accumulator columns are rewritten by the next batch
      if (last_of_ray) {
        cta_sync();
        if (ray == 0) R2_STAMP(23);
        if (tid < S) {
          // softmax over the views and the blended colour
          float mx = -FLT_MAX;
          for (int v = 0; v < V; ++v) mx = fmaxf(mx, sLogit[tid * V + v]);
          float den = 0.f, r = 0.f, g = 0.f, bl = 0.f;
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            if (v < V) {
              const float e = expf(sLogit[tid * V + v] - mx);
              den += e; r += cv[v].x * e; g += cv[v].y * e; bl += cv[v].z * e;
            }
          }
          sRGB[ray * 512 + tid * 4] = r / den; sRGB[ray * 512 + tid * 4 + 1] = g / den; sRGB[ray * 512 + tid * 4 + 2] = bl / den;
        }
        if (ray == 0) R2_STAMP(24);
        cta_sync();   // sLogit / sBl are rewritten for the next ray
      }
    }
    R2_STAMP(3);

    // ---- RayUnet -------------------------------------------------------------------------------------------------------
    wait_d(D_CONV1);
    R2_STAMP(4);
    epi_conv1(c, tmem, w.u[0]);
    a_ready();                                                             // a#1: c1
    R2_STAMP(5);
    wait_d(D_CONV2);
    R2_STAMP(6);
    epi_enc_pair(c, tmem, S / 2, w.u[1], sm + C2_HI, sm + C2_LO, RA_L3);
    a_ready();                                                             // a#2: c2
    R2_STAMP(7);
    wait_d(D_CONV3);
    R2_STAMP(8);
    epi_enc_pair(c, tmem, S / 4, w.u[2], sm + C3_HI, sm + C3_LO, RA_L4);
    a_ready();                                                             // a#3: c3
    R2_STAMP(9);
    wait_d(D_TCONV3);
    R2_STAMP(10);
    epi_dec<128, false>(c, tmem, S / 8, w.u[3], sm + X0_HI, sm + X0_LO, RA_L3, 0);
    a_ready();                                                             // a#4: x0
    R2_STAMP(11);
    wait_d(D_TCONV2);
    R2_STAMP(12);
    epi_dec<64, false>(c, tmem, S / 4, w.u[4], sm + X1_HI, sm + X1_LO, RA_L2, 0);
    a_ready();                                                             // a#5: x1
    R2_STAMP(13);
    wait_d(D_TCONV1);
    R2_STAMP(14);
    epi_dec<32, true>(c, tmem, S / 2, w.u[5], sm + Q_OFF, sm + Q_OFF + X2_PLANE, RA_X, X2_TILE);
    prep_x(sm, S, true, tid);                                              // x again (region P was recycled)
    prep_x(sm + X_TILE, S, live1, tid);
    tc::mbar_wait(&sy.x_full, 1);
    a_ready();                                                             // a#6: x | x2
    R2_STAMP(15);

    // ---- conv_out + LayerNorm([128, S]) + ELU; sigma = softplus(w . y + b) (model.py:525) ---------------------------------------
    wait_d(D_CONVOUT);
    R2_STAMP(16);
    {
      const bool valid = c.m < S;
      const int c0 = c.half * 64;
      const float n = (float)(S * 128);
      const UnetLayer& U = w.u[6];
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int ray = 0; ray < 2; ++ray)
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
          float v[16];
          float bv[16];
          ldb16(U.b + c0 + cc, bv);
          tld16(c.trow + tmem + 128 * ray + c0 + cc, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[ray] += v[j] + bv[j];
        }
      const float2 tot = block_sum2(valid ? sum[0] : 0.f, valid ? sum[1] : 0.f, red);
      const float mean[2] = {tot.x / n, tot.y / n};
      float q[2] = {0.f, 0.f};
#pragma unroll
      for (int ray = 0; ray < 2; ++ray)
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
          float v[16];
          float bv[16];
          ldb16(U.b + c0 + cc, bv);
          tld16(c.trow + tmem + 128 * ray + c0 + cc, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) { const float d = (v[j] + bv[j]) - mean[ray]; q[ray] += d * d; }
        }
      const float2 qt = block_sum2(valid ? q[0] : 0.f, valid ? q[1] : 0.f, red);
      const float rstd[2] = {1.f / sqrtf(qt.x / n + 1e-5f), 1.f / sqrtf(qt.y / n + 1e-5f)};
      float part[2] = {0.f, 0.f};
      const int srow = valid ? c.m : 0;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        float g[16], be[16], sw[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(U.g2 + ln_off(srow, c0 + cc + j, 128)));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(U.be2 + ln_off(srow, c0 + cc + j, 128)));
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(w.sig_w + c0 + cc + j));
          g[j] = g4.x; g[j + 1] = g4.y; g[j + 2] = g4.z; g[j + 3] = g4.w;
          be[j] = b4.x; be[j + 1] = b4.y; be[j + 2] = b4.z; be[j + 3] = b4.w;
          sw[j] = s4.x; sw[j + 1] = s4.y; sw[j + 2] = s4.z; sw[j + 3] = s4.w;
        }
#pragma unroll
        for (int ray = 0; ray < 2; ++ray) {
          float v[16];
          float bv[16];
          ldb16(U.b + c0 + cc, bv);
          tld16(c.trow + tmem + 128 * ray + c0 + cc, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            part[ray] = fmaf(elu(((v[j] + bv[j]) - mean[ray]) * rstd[ray] * g[j] + be[j]), sw[j], part[ray]);
        }
      }
      sSigP[0 * 256 + c.half * 128 + c.m] = part[0];
      sSigP[1 * 256 + c.half * 128 + c.m] = part[1];
    }
    cta_sync();
    R2_STAMP(17);

    // ---- compositing (model.py:541-575): threads 0-127 ray 0, threads 128-255 ray 1 -------------------------------------------
    const int cr_ray = tid >> 7, cs = tid & 127;
    const bool cr_live = cr_ray == 0 || live1;
    float* sSig = sV + cr_ray * 512;
    float* sT = sSig + 128;
    float* sWt = sSig + 256;
    float* sZ = sSig + 384;
    if (cs < S) {
      const float sg = softplus(sSigP[cr_ray * 256 + cs] + sSigP[cr_ray * 256 + 128 + cs] + __ldg(w.sig_b));
      if (sigma_dbg && cr_live) sigma_dbg[sbase[cr_ray] + cs] = sg;
      const float delta = cs + 1 < S ? sZ[cs + 1] - sZ[cs] : 1e2f;
      sSig[cs] = 1.f - expf(-delta * sg);  // alpha
    }
    cta_sync();
    if (cs == 0) {
      float T = 1.f;
      for (int s = 0; s < S; ++s) { sT[s] = T; T *= (1.f - sSig[s]); }
    }
    cta_sync();
    float wv = 0.f, zz = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, nv = 0.f;
    if (cs < S && cr_live) {
      wv = sSig[cs] * sT[cs];
      sWt[cs] = wv;
      weights_out[sbase[cr_ray] + cs] = wv;
      zz = sZ[cs];
      cr = sRGB[cr_ray * 512 + cs * 4]; cg = sRGB[cr_ray * 512 + cs * 4 + 1]; cb = sRGB[cr_ray * 512 + cs * 4 + 2];
      nv = nvalid[sbase[cr_ray] + cs] > 1 ? 1.f : 0.f;
    } else if (cs < S) {
      sWt[cs] = 0.f;
    }
    auto sum_mine = [&](float v) {
      const float2 t = block_sum2(cr_ray == 0 ? v : 0.f, cr_ray == 1 ? v : 0.f, red);
      return cr_ray == 0 ? t.x : t.y;
    };
    const float wsum = sum_mine(wv);
    const float depth = sum_mine(wv * zz);
    const float unc = sum_mine(wv * (zz - depth) * (zz - depth));
    float r = sum_mine(wv * cr), g = sum_mine(wv * cg), bch = sum_mine(wv * cb);
    const float cnt = sum_mine(nv);
    if (cs == 0 && cr_live) {
      const int64_t ray = ray0 + cr_ray;
      if (white_bkgd) { r += 1.f - wsum; g += 1.f - wsum; bch += 1.f - wsum; }
      rgb_out[ray * 3] = r; rgb_out[ray * 3 + 1] = g; rgb_out[ray * 3 + 2] = bch;
      depth_out[ray] = depth;
      unc_out[ray] = unc;
      mask_out[ray] = cnt > 8.f ? 1 : 0;
    }
    // per-ray sum of the weights, kept for the feature epilogue
    if (cs == 0) red[32 + cr_ray] = wsum;
    R2_STAMP(18);

    // ---- rendered feature (model.py:594-598): feat = W2 (sum_s w_s leaky(W1 x_s + b1)) + b2 sum_s w_s ----------------------------
    wait_d(D_FEAT);                                                        // both x tiles are dead now
    if (feat_out || peers.n > 0) {
      float* sPart = reinterpret_cast<float*>(sm);  // [128][132] fp32 over the dead x tiles
      float* sHs = sPart + 128 * 132;
      for (int ray = 0; ray < 2; ++ray) {
        if (ray == 1 && !live1) break;   // uniform
        cta_sync();                      // sPart / sHs of the previous ray consumed; sWt visible
        const float wrow = c.m < S ? sV[ray * 512 + 256 + c.m] : 0.f;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
          float y[16];
          tld16(c.trow + tmem + 256 + 128 * ray + c.half * 64 + cc, y);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.ft1_b + c.half * 64 + cc + j));
            float4 o;
            o.x = c.m < S ? leaky(y[j] + b4.x) * wrow : 0.f;
            o.y = c.m < S ? leaky(y[j + 1] + b4.y) * wrow : 0.f;
            o.z = c.m < S ? leaky(y[j + 2] + b4.z) * wrow : 0.f;
            o.w = c.m < S ? leaky(y[j + 3] + b4.w) * wrow : 0.f;
            *reinterpret_cast<float4*>(sPart + c.m * 132 + c.half * 64 + cc + j) = o;
          }
        }
        cta_sync();
        if (tid < 128) {
          float a = 0.f;
          for (int rr = 0; rr < 128; ++rr) a += sPart[rr * 132 + tid];
          sHs[tid] = a;
        }
        cta_sync();
        if (tid < C_FEAT) {
          const int64_t rayi = ray0 + ray;
          float a = __ldg(w.ft2_b + tid) * red[32 + ray];
          for (int k = 0; k < 128; ++k) a = fmaf(__ldg(w.ft2 + k * C_FEAT + tid), sHs[k], a);
          if (feat_out) feat_out[rayi * C_FEAT + tid] = a;
          // fused all-gather: the same 768-byte row goes to every rank's gathered matrix (peer stores over NVLink)
          for (int p = 0; p < peers.n; ++p) peers.p[p][(peers.row0 + rayi) * C_FEAT + tid] = a;
        }
      }
    }
    R2_STAMP(19);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace r2

int read_prof_ray2(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, r2::g_prof_ray2, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : set_error("read_prof_ray2 failed");
}

int launch_ray2(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t zs, int64_t R, int S, int white_bkgd,
                const unsigned char* xsplit, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
                float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
                const FeatPeers& peers, cudaStream_t st) {
  if (R <= 0) return 0;
  if (S % 8 != 0 || S < 8 || S > 128) return set_error("ray stage: samples per ray must be a multiple of 8 in [8, 128]");
  if (w.S != S) return set_error("ray stage: weights were packed for a different number of samples per ray");
  if (sc.V > 16) return set_error("ray stage: at most 16 reference views");
  cudaError_t e = cudaFuncSetAttribute(r2::ray2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r2::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  r2::ray2_kernel<<<(unsigned)((R + 1) / 2), NT + 128, r2::SMEM_BYTES, st>>>(sc, w, z_vals, zs, S, R, white_bkgd, xsplit, partial, rgbvis,
                                                                           nvalid, rgb, depth, weights, mask, depth_unc, feat,
                                                                           sigma_dbg, peers);
  return check_launch("ray2_kernel");
}

}  // namespace nlb
