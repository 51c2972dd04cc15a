// Per-sample stage of the conditional-NeRF render / query path (sm_100a).
//
//   aggregate_kernel  - MultiviewFeatureAggregator.forward (conditional_nerf/multiview_aggregator.py:156-222):
//                       two projections per (sample, view) (IBRNet convention ibrnet/ibrnet.py:169-231 and NeuRay
//                       convention conditional_nerf/depth_fusion.py:78-147), bilinear fetches from the reference
//                       views, the mixture-of-logistics visibility decoder (visibility_decoder.py:62-148),
//                       visibility-weighted mean/variance over views and the 393->64->128 MLP.  It also emits the
//                       per-view half of the colour-blend MLP's first layer (model.py:532-535), which is linear in
//                       its concatenated input, so the 195-channel per-view features never leave the SM.
//   neighbor_kernel   - the K=8 support-point MLP + attention of ConditionalNeRF.query (model.py:371-427).
//
// One CTA owns a tile of samples; activations stay in shared memory between layers (row-major, padded ld), weights
// are streamed from L2 by nlb::tile_gemm.  Two exact algebraic rewrites are used (DESIGN.md "rewrites"):
//   (1) base_mlp layer 1 is split into the support-feature part, precomputed once per frame per support point
//       (sup_pre = W1[:, :195] f + b1), and the per-pair part (positional encoding + ray difference);
//   (2) in base_mlp_attn the query is the same vector for all K positions (model.py:413-414), so the K and V
//       projections are folded onto the query / context side and `feature` has one distinct row per sample.
#include <float.h>
#include <stdlib.h>
#include <type_traits>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"

namespace nlb {


// phase timestamps of a mid-grid CTA (debug aid, read with nlb_debug_read_prof: slots 16..31)
__device__ long long g_prof[32];
#define AGG_STAMP(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_prof[16 + i] = clock64(); } while (0)

// ------------------------------------------------------------------------------------------------------------------
// shared geometry helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_point(const PointSrc& ps, int64_t n, float& x, float& y, float& z) {
  if (ps.xyz) {
    x = ps.xyz[n * 3]; y = ps.xyz[n * 3 + 1]; z = ps.xyz[n * 3 + 2];
  } else {
    // xyz = rays_o + rays_d * z, one rounding per op (model.py:498)
    const int64_t r = n / ps.S;
    const float t = ps.z[r * ps.zs + (n - r * ps.S)];
    x = __fadd_rn(ps.rays_o[r * 3 + 0], __fmul_rn(ps.rays_d[r * 3 + 0], t));
    y = __fadd_rn(ps.rays_o[r * 3 + 1], __fmul_rn(ps.rays_d[r * 3 + 1], t));
    z = __fadd_rn(ps.rays_o[r * 3 + 2], __fmul_rn(ps.rays_d[r * 3 + 2], t));
  }
}

struct Taps {  // bilinear footprint: 4 taps, weight 0 marks a skipped (out of range) tap
  int x0, y0;
  float w[4];  // nw, ne, sw, se
};

// torch grid_sample, bilinear.  `zeros`: padding_mode='zeros' (taps outside contribute nothing);
// otherwise 'border' (coordinate clipped first).
__device__ __forceinline__ Taps make_taps(float ix, float iy, int w, int h, bool zeros) {
  if (!zeros) {
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  }
  const float fx = floorf(ix), fy = floorf(iy);
  Taps t;
  // keep the int conversion well defined for far-away projections (weights are zero there anyway)
  t.x0 = (int)fminf(fmaxf(fx, -2.f), (float)w);
  t.y0 = (int)fminf(fmaxf(fy, -2.f), (float)h);
  const float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy;  // ix_se - ix, iy_se - iy
  const float wx = ix - fx, wy = iy - fy;
  t.w[0] = ex * ey; t.w[1] = wx * ey; t.w[2] = ex * wy; t.w[3] = wx * wy;
  const bool inx0 = fx >= 0.f && fx <= (float)(w - 1), inx1 = fx + 1.f >= 0.f && fx + 1.f <= (float)(w - 1);
  const bool iny0 = fy >= 0.f && fy <= (float)(h - 1), iny1 = fy + 1.f >= 0.f && fy + 1.f <= (float)(h - 1);
  if (!(inx0 && iny0)) t.w[0] = 0.f;
  if (!(inx1 && iny0)) t.w[1] = 0.f;
  if (!(inx0 && iny1)) t.w[2] = 0.f;
  if (!(inx1 && iny1)) t.w[3] = 0.f;
  // NaN coordinates: every comparison above is false -> all weights zero
  return t;
}

// row info slots
// per map a bilinear footprint is stored ready to use: 4 pixel indices (y * width + x, clamped into the map; as int bits) and
// 4 weights (0 for taps that do not contribute) - computed once per row in phase 1 instead of once per lane in every gather
enum { RI_TF = 0, RI_TI = 8, RI_TV = 16, RI_DEPTH = 24, RI_VALID, RI_MASK, RI_DD, RI_VIS /* 28: float4 VIS W RD0 RD1 */, RI_W,
       RI_RD0, RI_RD1, RI_RD2 /* 32: RD2 RD3 */, RI_RD3, RI_V, RI_P, RI_N = 36 };   // RI_V / RI_P: view and sample index of the row (int bits)

__device__ __forceinline__ void store_taps(float* slot, const Taps& t, int w, int h) {
  const int x0 = min(max(t.x0, 0), w - 1), x1 = min(max(t.x0 + 1, 0), w - 1);
  const int y0 = min(max(t.y0, 0), h - 1), y1 = min(max(t.y0 + 1, 0), h - 1);
  *reinterpret_cast<int4*>(slot) = make_int4(y0 * w + x0, y0 * w + x1, y1 * w + x0, y1 * w + x1);
  *reinterpret_cast<float4*>(slot + 4) = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
}

constexpr int LDF = 196;   // rgb_feat rows: 195 (+1)
constexpr int LDH = 132;
constexpr int LDX = 36;
constexpr int LDG = 420;   // 393 -> 416 (+4)
#ifndef AGG_ROWS
#define AGG_ROWS 64
#endif
#ifndef AGG_PERSIST
#define AGG_PERSIST 0   // 1: persistent CTAs with next-tile prefetch for the render variant (measured slower: 160 vs 152 ms, the
#endif                  // two CTAs of an SM fall into lockstep and their latency-bound phases stop covering each other)

// ROWS = (sample, view) rows per CTA.  128 rows fill the SM with one CTA; 64 rows fit two CTAs per SM (about 110 KB each), which
// doubles the resident warps for the latency-bound gather phases.
// FUSED (V <= 8): the gather of phase 5 feeds the mean / variance of phase 6 through registers, so the [ROWS][LDF] tile of
// interpolated features does not exist and the arena only holds the decoder input; the 50 KB it saves per CTA go to the L1.
// SLIM (the persistent render variant: decoder and out_fc live in their own kernels): no weight staging ring, no row tile
template <int ROWS, bool FUSED, bool SLIM = false>
constexpr int agg_smem_floats() {
  return (SLIM ? 0 : STAGE_FLOATS + ROWS * (FUSED ? LDX : LDF)) + (ROWS / 8) * (SLIM ? 32 : LDG) + (ROWS / 8) * 68 + ROWS * RI_N + (ROWS / 8) * 4 +
         2 * (ROWS / 8) * 8 + 2 * ROWS * 2;   // + the next tile's point data and visibility rows (persistent GOUT variant)
}

// visibility-weighted mean / variance over the views of one sample (ibrnet.py:8-12): one warp, lanes over channels
template <int VM>
__device__ __forceinline__ void mean_var_rows(const float* __restrict__ f0, const float* __restrict__ ri0, const int V,
                                              const int lane, float* __restrict__ g) {
  float wv[VM];
#pragma unroll
  for (int v = 0; v < VM; ++v) wv[v] = v < V ? ri0[v * RI_N + RI_W] : 0.f;
  // all loads of a lane's (up to) seven channels are requested before the first is consumed (V <= 8; two at a time above)
  constexpr int NC = (C_RGBF + 31) / 32;
  constexpr int JB = VM <= 8 ? NC : 2;
#pragma unroll
  for (int j0 = 0; j0 < NC; j0 += JB) {
    float f[JB][VM];
#pragma unroll
    for (int j = 0; j < JB; ++j) {
      const int c = min(lane + 32 * (j0 + j), C_RGBF - 1);
#pragma unroll
      for (int v = 0; v < VM; ++v) f[j][v] = v < V ? f0[v * LDF + c] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < JB; ++j) {
      const int c = lane + 32 * (j0 + j);
      float m = 0.f;
#pragma unroll
      for (int v = 0; v < VM; ++v) m += f[j][v] * wv[v];
      float var = 0.f;
#pragma unroll
      for (int v = 0; v < VM; ++v) { const float d = f[j][v] - m; var += wv[v] * (d * d); }
      if (j0 + j < NC && c < C_RGBF) {
        g[c] = m;
        g[C_RGBF + c] = var;
      }
    }
  }
}

// EXT: visibility and depth difference per (sample, view) come from visibility_kernel (`visdd_in`, [N][V] float2) - the NeuRay
// projection, the visibility-feature gather and the decoder are compiled out.
// GOUT (with FUSED and EXT): the per-sample statistics vector (mean | variance | 3 extras, the input of out_fc) goes to global
// memory (`g_out`, [N][416]: map-channel means 0..191, variances 192..383, rgb mean / variance 384..389, extras 390..392, zeros)
// and out_fc + the attention query projection run as 128-sample tensor-core GEMMs in fc_tail_kernel.
template <int ROWS, bool FUSED, bool EXT, bool GOUT = false>
__global__ void __launch_bounds__(NT, (FUSED && EXT && GOUT) ? 2 : 128 / ROWS)
aggregate_kernel(const SceneDev sc, const RenderW w, const PointSrc ps, const int64_t N, const int with_blend,
                 const float2* __restrict__ visdd_in, float* __restrict__ g_out,
                 float* __restrict__ agg_out, float* __restrict__ partial_out, float* __restrict__ rgbvis_out,
                 unsigned char* __restrict__ nvalid_out, float* __restrict__ mvf_out, float* __restrict__ mvv_out) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;
  constexpr bool SLIM = FUSED && EXT && GOUT;
  float* arena = sB + (SLIM ? 0 : STAGE_FLOATS);
  constexpr int TP_MAX = ROWS / 8;
  constexpr int PARTS = NT / ROWS;  // threads per row in the per-row scalar phases
  float* sG = arena + (SLIM ? 0 : ROWS * (FUSED ? LDX : LDF));
  // SLIM: of a sample's statistics vector only the three extras 390..392 (+ zero padding) pass through shared memory
  constexpr int LDGS = SLIM ? 32 : LDG, G0 = SLIM ? 390 : 0;
  float* sO1 = sG + TP_MAX * LDGS;
  float* sRI = sO1 + TP_MAX * 68;
  float* sPt = sRI + ROWS * RI_N;
  float* sPre = sPt + TP_MAX * 4;        // [2][TP_MAX][8]: o | d | z (or xyz) of the samples of this / the next tile
  float* sVd = sPre + 2 * TP_MAX * 8;    // [2][ROWS] float2: visibility | depth difference rows of this / the next tile
  float* sX = arena;                // [ROWS][LDX]
  float* sF = arena;                // [ROWS][LDF] (after the decoder is done)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = sc.V;
  const int TP = min(TP_MAX, ROWS / V);
  // Which samples a tile holds.  Explicit point lists: TP consecutive points.  Ray samples (n = ray * S + s): the SAME sample
  // index of TP consecutive rays - neighbouring rays at one depth project onto the same few pixels of a reference view (a
  // quarter of a feature-map pixel apart for adjacent query pixels), consecutive samples of one ray do not (more than a pixel
  // apart along the epipolar line), and the bilinear gathers of a tile are what this kernel waits for.
  const bool by_ray = ps.xyz == nullptr;
  const int64_t n_items = by_ray ? N / ps.S : N;                     // rays or points
  const int64_t n_tiles = by_ray ? ((n_items + TP - 1) / TP) * ps.S : (n_items + TP - 1) / TP;
  const float near_ = sc.near_, far_ = sc.far_;
  // GOUT (the render path): persistent CTAs.  The point data and the visibility rows of a CTA's NEXT tile are copied into shared
  // memory (cp.async) while the current tile is gathered, so a tile starts with its projections instead of with a round trip to
  // global memory, and the launch cost of a CTA is paid once per 16 k tiles.  The other variants run one tile per CTA.
  auto prefetch = [&](const int64_t t, const int buf) {
    const int64_t tg = by_ray ? t / ps.S : t;
    const int ts = by_ray ? (int)(t - tg * ps.S) : 0;
    const int npn = (int)min((int64_t)TP, n_items - tg * TP);
    if (tid < npn * 8) {
      const int pp = tid >> 3, f = tid & 7;
      const int64_t item = tg * TP + pp;
      const float* src = nullptr;
      if (by_ray) src = f < 3 ? ps.rays_o + item * 3 + f : (f < 6 ? ps.rays_d + item * 3 + (f - 3) : (f == 6 ? ps.z + item * ps.zs + ts : nullptr));
      else src = f < 3 ? ps.xyz + item * 3 + f : nullptr;
      if (src) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sPre + (buf * TP_MAX + pp) * 8 + f);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src));
      }
    }
    if (tid >= 64 && tid < 64 + npn * V) {
      const int r = tid - 64, pp = r / V;
      const int64_t n = by_ray ? (tg * TP + pp) * ps.S + ts : tg * TP + pp;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sVd + (buf * ROWS + r) * 2);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(visdd_in + n * V + (r - pp * V)));
    }
    cp_async_commit();
  };
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
  const int64_t tile_g = by_ray ? tile / ps.S : tile;
  const int tile_s = by_ray ? (int)(tile - tile_g * ps.S) : 0;
  const int np = (int)min((int64_t)TP, n_items - tile_g * TP);
  auto nidx = [&](int p) -> int64_t { return by_ray ? (tile_g * TP + p) * ps.S + tile_s : tile_g * TP + p; };
  const int rows = np * V;
  const int pbuf = it & 1;

  AGG_STAMP(0);
  float2 vd_pre = make_float2(0.f, 0.f);
  if (GOUT && AGG_PERSIST) {
    if (it == 0) prefetch(tile, 0);
    cp_async_wait<0>();
    cta_sync();   // this tile's point data and visibility rows landed; the previous tile is done with sRI / sG
    if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x, pbuf ^ 1);
  } else if (EXT && tid < rows) {
    // visibility | depth difference of this thread's (sample, view) row: requested now, consumed after the projections
    vd_pre = __ldcs(visdd_in + nidx(tid / V) * V + (tid - (tid / V) * V));   // written once by visibility_kernel
  }
  // decoder weights (32 KB) go into the staging ring, which is idle until phase 7: dec1 as [32][128] (the four heads side by
  // side), dec2 as 4 x [32][32].  Both are read as mma.sync B fragments (thread (g, t) reads rows t / t + 4 resp. 2t / 2t + 1,
  // column g), so 8-column groups are XOR-swizzled with the row to keep those reads bank-conflict free without padding:
  // dec1 (k, n) -> k * 128 + (n ^ 8 (k & 3)), dec2 (k, n) -> k * 32 + (n ^ 8 ((k >> 1) & 3)).  Requested now, consumed in phase 3.
  if (!EXT) {
    for (int i = tid; i < 1024; i += NT) {
      const int k = i >> 5, n = (i & 31) * 4;
      cp_async16(sB + k * 128 + (n ^ ((k & 3) << 3)), w.dec1 + i * 4);
    }
    for (int i = tid; i < 1024; i += NT) {
      const int k = (i >> 3) & 31, n = (i & 7) * 4;
      cp_async16(sB + 4096 + (i >> 3) * 32 + (n ^ (((k >> 1) & 3) << 3)), w.dec2 + i * 4);
    }
    cp_async_commit();
  }

  // ---- phase 1: projections; PARTS threads per (sample, view) row share the three independent pieces -------------
  {
    const int r = tid % ROWS, part = tid / ROWS;
    float* ri = sRI + r * RI_N;
    if (r < rows) {
      const int p = r / V, v = r - p * V;
      float x, y, z;
      if (GOUT && AGG_PERSIST) {
        const float* pr = sPre + (pbuf * TP_MAX + p) * 8;
        if (by_ray) {   // xyz = rays_o + rays_d * z, one rounding per op (model.py:498)
          x = __fadd_rn(pr[0], __fmul_rn(pr[3], pr[6]));
          y = __fadd_rn(pr[1], __fmul_rn(pr[4], pr[6]));
          z = __fadd_rn(pr[2], __fmul_rn(pr[5], pr[6]));
        } else {
          x = pr[0]; y = pr[1]; z = pr[2];
        }
      } else {
        load_point(ps, nidx(p), x, y, z);
      }
      const float* cam = sc.cams + v * 32;
      // (EXT: the NeuRay projection, job 1, lives in visibility_kernel; the two remaining jobs go to different threads of the row)
      for (int jj = part; jj < (EXT ? 2 : 3); jj += PARTS) {
        const int job = EXT ? jj * 2 : jj;
        if (job == 0) {
          // the (sample, view) split of the row index is taken once here: a runtime division per row in each of the later
          // per-row loops was 11 % of the kernel's instructions
          ri[RI_V] = __int_as_float(v); ri[RI_P] = __int_as_float(p);
          if (v == 0) { sPt[p * 4] = x; sPt[p * 4 + 1] = y; sPt[p * 4 + 2] = z; }
          // IBRNet convention
          const float ph0 = fmaf(cam[2], z, fmaf(cam[1], y, cam[0] * x)) + cam[3];
          const float ph1 = fmaf(cam[6], z, fmaf(cam[5], y, cam[4] * x)) + cam[7];
          const float ph2 = fmaf(cam[10], z, fmaf(cam[9], y, cam[8] * x)) + cam[11];
          const float zc = fmaxf(ph2, 1e-8f);
          float px = ph0 / zc, py = ph1 / zc;
          px = fminf(fmaxf(px, -1e6f), 1e6f);
          py = fminf(fmaxf(py, -1e6f), 1e6f);
          const bool inb = px <= (float)(sc.W - 1) && px >= 0.f && py <= (float)(sc.H - 1) && py >= 0.f;
          ri[RI_MASK] = (inb && ph2 > 0.f) ? 1.f : 0.f;
          const float gx = 2.f * px / (float)(sc.W - 1) - 1.f, gy = 2.f * py / (float)(sc.H - 1) - 1.f;
          store_taps(ri + RI_TI, make_taps(((gx + 1.f) / 2.f) * (float)(sc.W - 1), ((gy + 1.f) / 2.f) * (float)(sc.H - 1), sc.W, sc.H, true), sc.W, sc.H);
          store_taps(ri + RI_TF, make_taps(((gx + 1.f) / 2.f) * (float)(sc.w - 1), ((gy + 1.f) / 2.f) * (float)(sc.h - 1), sc.w, sc.h, true), sc.w, sc.h);
        } else if (job == 1) {
          // NeuRay convention
          const float* kr = cam + 12;
          const float c0 = fmaf(kr[2], z, fmaf(kr[1], y, kr[0] * x)) + kr[3];
          const float c1 = fmaf(kr[6], z, fmaf(kr[5], y, kr[4] * x)) + kr[7];
          float dep = fmaf(kr[10], z, fmaf(kr[9], y, kr[8] * x)) + kr[11];
          const bool bad = fabsf(dep) < 1e-4f;
          if (bad) dep = 1e-3f;
          const float qx = c0 / dep, qy = c1 / dep;
          const bool outside = qx < -0.5f || qx >= (float)sc.W - 0.5f || qy < -0.5f || qy >= (float)sc.H - 0.5f;
          ri[RI_VALID] = (!bad && !outside) ? 1.f : 0.f;
          ri[RI_DEPTH] = dep;
          const float xn = qx / (float)(sc.W - 1) * 2.f - 1.f, yn = qy / (float)(sc.H - 1) * 2.f - 1.f;
          float vx, vy;
          if (sc.vh == sc.H && sc.vw == sc.W) {  // align_corners=True only when the map has the image size
            vx = ((xn + 1.f) / 2.f) * (float)(sc.vw - 1);
            vy = ((yn + 1.f) / 2.f) * (float)(sc.vh - 1);
          } else {
            vx = ((xn + 1.f) * (float)sc.vw - 1.f) / 2.f;
            vy = ((yn + 1.f) * (float)sc.vh - 1.f) / 2.f;
          }
          store_taps(ri + RI_TV, make_taps(vx, vy, sc.vw, sc.vh, false), sc.vw, sc.vh);
        } else {
          // colour-blend ray difference (ibrnet.py:144-167) between the query camera and view v
          const float* cc = cam + 24;
          float ax = sc.qc[0] - x, ay = sc.qc[1] - y, az = sc.qc[2] - z;
          const float an = sqrtf(ax * ax + ay * ay + az * az) + 1e-6f;
          ax /= an; ay /= an; az /= an;
          float bx = cc[0] - x, by = cc[1] - y, bz = cc[2] - z;
          const float bn = sqrtf(bx * bx + by * by + bz * bz) + 1e-6f;
          bx /= bn; by /= bn; bz /= bn;
          const float dx = ax - bx, dy = ay - by, dz = az - bz;
          const float dn = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);
          ri[RI_RD0] = dx / dn; ri[RI_RD1] = dy / dn; ri[RI_RD2] = dz / dn;
          ri[RI_RD3] = ax * bx + ay * by + az * bz;
        }
      }
    } else {
      for (int i = part; i < RI_N; i += PARTS) ri[i] = 0.f;
    }
  }
  cta_sync();

  AGG_STAMP(1);
  if (!EXT) {
  // ---- phase 2: 32-channel visibility features (border padding), one warp per row ------------------------------
  // All four taps of all ROWS / 8 rows of a warp are requested before any is consumed (addresses clamped into the map, taps
  // outside carry weight 0): one L2 round trip for the whole phase.
  constexpr int VU = ROWS / (NT / 32);
  for (int rb = warp; rb < ROWS; rb += VU * (NT / 32)) {
    float q[VU][4], wgt[VU][4], valid[VU];
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int r = rb + u * (NT / 32);
      valid[u] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) { q[u][t] = 0.f; wgt[u][t] = 0.f; }
      if (r < rows) {
        const float* ri = sRI + r * RI_N;
        const int v = __float_as_int(ri[RI_V]);
        const int4 ti = *reinterpret_cast<const int4*>(ri + RI_TV);
        const float4 tw = *reinterpret_cast<const float4*>(ri + RI_TV + 4);
        const float* base = sc.vis + ((size_t)v * sc.vh * sc.vw) * C_VIS + lane;
        q[u][0] = __ldg(base + (size_t)ti.x * C_VIS);
        q[u][1] = __ldg(base + (size_t)ti.y * C_VIS);
        q[u][2] = __ldg(base + (size_t)ti.z * C_VIS);
        q[u][3] = __ldg(base + (size_t)ti.w * C_VIS);
        wgt[u][0] = tw.x; wgt[u][1] = tw.y; wgt[u][2] = tw.z; wgt[u][3] = tw.w;
        valid[u] = ri[RI_VALID];
      }
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int r = rb + u * (NT / 32);
      if (r < ROWS) {
        float a = q[u][0] * wgt[u][0];
        a += q[u][1] * wgt[u][1];
        a += q[u][2] * wgt[u][2];
        a += q[u][3] * wgt[u][3];
        sX[r * LDX + lane] = a * valid[u];
      }
    }
  }
  // (tile_gemm starts with a __syncthreads)

  AGG_STAMP(2);
  // ---- phase 3: visibility decoder ------------------------------------------------------------------------------
  cp_async_wait<0>();
  cta_sync();  // decoder weights landed, sX complete
  AGG_STAMP(9);
  // Both decoder layers and the head outputs run on the warp-level tensor-core path (mma.sync m16n8k8 tf32 with the 3xTF32
  // hi / lo split, fp32 accumulation) and never leave registers: warp (mt, hp) owns the 16-row tile mt and the head pair hp
  // (heads 2hp, 2hp + 1 = 64 of the 128 hidden columns).  Layer 2 is block diagonal (head h maps its own 32 columns to
  // themselves), so the layer-1 accumulator fragment IS the layer-2 A operand: an accumulator holds columns 2t, 2t + 1 of an
  // 8-column tile where the A fragment wants k = t, t + 4, and since a product does not care in which order k is summed the
  // B fragment is simply read in the matching order (rows 2t, 2t + 1).  The six head outputs (visibility_decoder.py:99-148)
  // are 32-long dot products with the layer-2 rows: in-thread over the 8 columns a thread holds, then over the four t lanes.
  {
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp & 3, hp = warp >> 2;
    const float* xa = sX + (mt * 16 + g) * LDX;   // rows g and g + 8 of the tile
    const float* xb = xa + 8 * LDX;
    float c1[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { c1[nt][0] = 0.f; c1[nt][1] = 0.f; c1[nt][2] = 0.f; c1[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      float ah[4], al[4];
      split_hi_lo(xa[ks * 8 + t], ah[0], al[0]);
      split_hi_lo(xb[ks * 8 + t], ah[1], al[1]);
      split_hi_lo(xa[ks * 8 + t + 4], ah[2], al[2]);
      split_hi_lo(xb[ks * 8 + t + 4], ah[3], al[3]);
      const float* b0p = sB + (ks * 8 + t) * 128;   // rows k = 8 ks + t and k + 4: both have k & 3 == t
      // MMAs into one accumulator are dependent: the three 3xTF32 passes are issued pass-major over the eight n-tiles
      float bh[8][2], bl[8][2];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = (hp * 64 + nt * 8 + g) ^ (t << 3);
        split_hi_lo(b0p[n], bh[nt][0], bl[nt][0]);
        split_hi_lo(b0p[4 * 128 + n], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(c1[nt], al, bh[nt]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(c1[nt], ah, bl[nt]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(c1[nt], ah, bh[nt]);
    }
#pragma unroll
    for (int hl = 0; hl < 2; ++hl) {
      const int hh = hp * 2 + hl;
      float c2[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { c2[nt][0] = 0.f; c2[nt][1] = 0.f; c2[nt][2] = 0.f; c2[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 b1 = __ldg(reinterpret_cast<const float2*>(w.dec1_b + hh * 32 + ks * 8 + 2 * t));
        const float* c = c1[hl * 4 + ks];
        float ah[4], al[4];
        split_hi_lo(elu(c[0] + b1.x), ah[0], al[0]);   // (row g,     k = 2t)
        split_hi_lo(elu(c[2] + b1.x), ah[1], al[1]);   // (row g + 8, k = 2t)
        split_hi_lo(elu(c[1] + b1.y), ah[2], al[2]);   // (row g,     k = 2t + 1)
        split_hi_lo(elu(c[3] + b1.y), ah[3], al[3]);   // (row g + 8, k = 2t + 1)
        const float* b0p = sB + 4096 + hh * 1024 + (ks * 8 + 2 * t) * 32;   // rows k = 8 ks + 2t, k + 1: (k >> 1) & 3 == t
        float bh[4][2], bl[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int n = (nt * 8 + g) ^ (t << 3);
          split_hi_lo(b0p[n], bh[nt][0], bl[nt][0]);
          split_hi_lo(b0p[32 + n], bh[nt][1], bl[nt][1]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_16x8x8(c2[nt], al, bh[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_16x8x8(c2[nt], ah, bl[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_16x8x8(c2[nt], ah, bh[nt]);
      }
      // layer-2 bias + ELU in place: c2[nt] = rows g | g + 8, columns 8 nt + 2t, + 1 of head hh
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 b2 = __ldg(reinterpret_cast<const float2*>(w.dec2_b + hh * 32 + nt * 8 + 2 * t));
        c2[nt][0] = elu(c2[nt][0] + b2.x); c2[nt][1] = elu(c2[nt][1] + b2.y);
        c2[nt][2] = elu(c2[nt][2] + b2.x); c2[nt][3] = elu(c2[nt][3] + b2.y);
      }
      // head outputs: mean head -> j = 0, 1 (softplus), scale head -> j = 2, 3 (softplus + 0.05), mixture weight -> j = 4
      // and visibility scale -> j = 5 (sigmoid)
      const int nj = hp == 0 ? 2 : 1, jb = hp == 0 ? hl * 2 : 4 + hl;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj < nj) {   // warp-uniform
          const int j = jb + jj;
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float2 w3 = __ldg(reinterpret_cast<const float2*>(w.dec3 + j * 32 + nt * 8 + 2 * t));
            p0 = fmaf(c2[nt][0], w3.x, p0); p0 = fmaf(c2[nt][1], w3.y, p0);
            p1 = fmaf(c2[nt][2], w3.x, p1); p1 = fmaf(c2[nt][3], w3.y, p1);
          }
          p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
          p0 += __shfl_xor_sync(0xffffffffu, p0, 2); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
          if (t < 2) {   // lane t = 0 finishes row g, lane t = 1 row g + 8
            const float o = (t == 0 ? p0 : p1) + __ldg(w.dec3_b + j);
            sO1[(mt * 16 + g + 8 * t) * 6 + j] = j < 2 ? softplus_fast(o) : (j < 4 ? softplus_fast(o) + 0.05f : sigmoid_fast(o));
          }
        }
      }
    }
  }
  }
  AGG_STAMP(10);
  AGG_STAMP(11);
  cta_sync();
  const bool fused_w = V == 8;   // rows of a sample are 8 consecutive lanes: the view weights follow in the same threads
  if (tid < ROWS && (fused_w || tid < rows)) {
    float* ri = sRI + tid * RI_N;
    const bool live = tid < rows;
    float vis = 0.f, dd = 0.f;
    if (EXT) {
      if (live) {
        if (GOUT && AGG_PERSIST) vd_pre = *reinterpret_cast<const float2*>(sVd + (pbuf * ROWS + tid) * 2);
        vis = vd_pre.x; dd = vd_pre.y;
      }
    } else {
      const float m0 = sO1[tid * 6], m1 = sO1[tid * 6 + 1], v0 = sO1[tid * 6 + 2], v1 = sO1[tid * 6 + 3];
      const float aw = sO1[tid * 6 + 4], vs = sO1[tid * 6 + 5];
      const float dep = ri[RI_DEPTH];
      const float near_inv = -1.f / near_, far_inv = -1.f / far_;
      // (divisions of this serial per-row tail use the hardware reciprocal: 2 ulp, operands far from the denormal range)
      float refd = __fdividef(-1.f, m0 * (far_inv - near_inv) + near_inv);
      refd = fminf(fmaxf(refd, near_), far_);
      dd = live ? __fdividef(fabsf(dep - refd), far_ - near_) : 0.f;
      const float dn = __fdividef(__fdividef(-1.f, fmaxf(dep, 1e-5f)) - near_inv, far_inv - near_inv);
      const float cdf0 = (0.5f + 0.5f * tanh_fast((dn - m0) * v0)) * vs;
      const float cdf1 = (0.5f + 0.5f * tanh_fast((dn - m1) * v1)) * vs;
      vis = live ? ((1.f - cdf0) * aw + (1.f - cdf1) * (1.f - aw)) * ri[RI_VALID] : 0.f;
    }
    if (live) {
      ri[RI_DD] = dd;
      ri[RI_VIS] = vis;
      if (!EXT && mvv_out) mvv_out[nidx(tid / V) * V + (tid - (tid / V) * V)] = vis;
    }
    if (fused_w) {
      // ---- phase 4 fused (V == 8): per-sample view weights over the 8 lanes of a sample (same summation tree as warp_sum) --
      auto sum8 = [](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
      };
      const int p = tid >> 3, v = tid & 7;
      const float den = sum8(vis) + 1e-8f;
      const float wv = __fdividef(vis, den);
      const float ddm = sum8(dd * wv);
      const float wsum = sum8(wv);
      const unsigned bal = __ballot_sync(0xffffffffu, live && ri[RI_MASK] != 0.f);
      const unsigned nval = __popc((bal >> (8 * (lane >> 3))) & 0xffu);
      const float d = dd - ddm;
      const float ddv = sum8(wv * (d * d));
      if (live) {
        ri[RI_W] = wv;
        float* g = sG + p * LDGS - G0;
        if (v == 0) {
          g[390] = ddm; g[391] = ddv; g[392] = wsum / (float)V;
          if (nvalid_out) nvalid_out[nidx(p)] = (unsigned char)nval;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (v * 3 + j < 23) g[393 + v * 3 + j] = 0.f;
      }
    }
  }
  if (!fused_w) {
    cta_sync();
    // ---- phase 4: per-sample view weights ------------------------------------------------------------------------
    for (int p = warp; p < np; p += NT / 32) {
      // one warp per sample, lane = view
      const bool on = lane < V;
      float* ri = sRI + (p * V + (on ? lane : 0)) * RI_N;
      const float vis = on ? ri[RI_VIS] : 0.f, dd = on ? ri[RI_DD] : 0.f;
      const float den = warp_sum(vis) + 1e-8f;
      const float wv = vis / den;
      if (on) ri[RI_W] = wv;
      const float ddm = warp_sum(dd * wv);
      const float wsum = warp_sum(wv);
      const unsigned nval = __popc(__ballot_sync(0xffffffffu, on && ri[RI_MASK] != 0.f));
      const float d = dd - ddm;
      const float ddv = warp_sum(wv * (d * d));
      float* g = sG + p * LDGS - G0;
      if (lane == 0) {
        g[390] = ddm; g[391] = ddv; g[392] = wsum / (float)V;
        if (nvalid_out) nvalid_out[nidx(p)] = (unsigned char)nval;
      }
      if (lane < 23) g[393 + lane] = 0.f;
    }
  }
  AGG_STAMP(3);
  cta_sync();  // sX (arena) is dead from here on; view weights visible

  AGG_STAMP(4);
  if (FUSED) {
    // ---- phases 5 + 6 fused (V <= 8): one warp per SAMPLE.  The warp walks the sample's views two at a time (24 independent
    // 8-byte loads in flight per lane), keeps the interpolated 192 map channels of all views in registers (lane l: channels
    // 3 + 64 j + 2 l, + 1) and takes the visibility-weighted mean / variance (ibrnet.py:8-12) from there.  The warps of a CTA
    // work on CONSECUTIVE samples of a ray at the same time: their footprints in a reference view are the same few pixels, so
    // most of the 29 KB a sample gathers are L1 hits on lines a neighbouring warp has just requested.
    float wb[8], bias = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) wb[i] = 0.f;
    if (with_blend) {
#pragma unroll
      for (int i = 0; i < 3; ++i) wb[i] = __ldg(w.bl1v + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < 5; ++i) wb[3 + i] = __ldg(w.bl1v + (195 + i) * 32 + lane);
      bias = __ldg(w.bl1_b + lane);
    }
    const float* fb0 = sc.feat + lane * 2;
    const float* bb0 = sc.featb + lane;
    const unsigned hw = (unsigned)(sc.h * sc.w);
    // FULL (V == 8, the rendering configuration): no per-view guards, no zero-filled buffers
    auto gather_samples = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      // (the pipeline does not drain between the samples of a warp: views 0 and 1 of the next sample are requested while views 6
      // and 7 of this one are consumed, its colour taps one sample ahead)
      // Software pipeline over the views, two register buffers: while view v is interpolated the loads of view v + 1 are in flight
      // and those of view v + 2 are issued into the buffer v just left.  The statistics are accumulated in one pass about the first
      // view's value K (d = x - K, exactly 0 for view 0): mean = K sw + sum w d, var = sum w (d - c)^2 = sum w d^2 - 2 c sum w d +
      // c^2 sw with c = mean - K - so the views' channels need not stay in registers for a second pass.
      f32x2 q[2][4][3];
      float wt[2][4], bq[2][4];
      auto issue = [&](const int b, const int ps_, const int v) {
        const float* ri0 = sRI + ps_ * V * RI_N;
        if (!FULL) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            wt[b][t] = 0.f; bq[b][t] = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) q[b][t][j] = 0ull;
          }
        }
        if (FULL || v < V) {
          const float* ri = ri0 + v * RI_N;
          const int4 ti = *reinterpret_cast<const int4*>(ri + RI_TF);
          const float4 tw = *reinterpret_cast<const float4*>(ri + RI_TF + 4);
          const unsigned vb = (unsigned)v * hw;
          const unsigned pg[4] = {vb + (unsigned)ti.x, vb + (unsigned)ti.y, vb + (unsigned)ti.z, vb + (unsigned)ti.w};
          const float twv[4] = {tw.x, tw.y, tw.z, tw.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            wt[b][t] = twv[t];
            const float* tp = fb0 + (size_t)pg[t] * C_FEAT;
#pragma unroll
            for (int j = 0; j < 3; ++j) q[b][t][j] = __ldg(reinterpret_cast<const f32x2*>(tp + j * 64));
          }
          if (with_blend) {
#pragma unroll
            for (int t = 0; t < 4; ++t) bq[b][t] = __ldg(bb0 + (size_t)pg[t] * 32);
          }
        }
      };
      auto colour = [&](const int ps_) {
        // colour: lane (v, t) = (lane / 4, lane % 4) fetches tap t of view v, the taps are summed over the 4 lanes of a view
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FULL || (lane >> 2) < V) {
          const float* ri = sRI + (ps_ * V + (lane >> 2)) * RI_N;
          const float wi = ri[RI_TI + 4 + (lane & 3)];
          if (wi != 0.f) {
            const int pix = __float_as_int(ri[RI_TI + (lane & 3)]);
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc.images + ((size_t)(lane >> 2) * sc.H * sc.W + pix) * 4));
            c = make_float4(c4.x * wi, c4.y * wi, c4.z * wi, 0.f);
          }
        }
        return c;
      };
      float4 crgb_next = make_float4(0.f, 0.f, 0.f, 0.f);
      if (warp < np) {
        crgb_next = colour(warp);
        issue(0, warp, 0);
        issue(1, warp, 1);
      }
      for (int p = warp; p < np; p += NT / 32) {
      const bool has_next = p + NT / 32 < np;
      const float* ri0 = sRI + p * V * RI_N;
      // output addresses once per sample; per view they differ by compile-time offsets (8 views: rows 8 n .. 8 n + 7 of partial_off)
      const int64_t n = nidx(p);
      float* rv_ptr = rgbvis_out ? rgbvis_out + n * V * 4 : nullptr;
      float* pp_ptr = partial_out + (FULL ? (n >> 2) * 1024 + (lane >> 2) * 128 + (n & 3) * 32 + (lane & 3) : 0);
      f32x2 nkk[3], sx[3], sxx[3];   // -K | sum w d | sum w d^2, channel pairs
      float rk = 0.f, rsx = 0.f, rsxx = 0.f, sw = 0.f;   // the same for the lane's colour channel
      float4 crgb = crgb_next;
      if (has_next) crgb_next = colour(p + NT / 32);
      crgb.x += __shfl_xor_sync(0xffffffffu, crgb.x, 1); crgb.y += __shfl_xor_sync(0xffffffffu, crgb.y, 1); crgb.z += __shfl_xor_sync(0xffffffffu, crgb.z, 1);
      crgb.x += __shfl_xor_sync(0xffffffffu, crgb.x, 2); crgb.y += __shfl_xor_sync(0xffffffffu, crgb.y, 2); crgb.z += __shfl_xor_sync(0xffffffffu, crgb.z, 2);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int b = v & 1;
        // d = sum_t q_t w_t - K: the interpolation starts from -K (0 for view 0, whose value becomes K)
        f32x2 d[3];
        {
          const f32x2 w2[4] = {pk2(wt[b][0], wt[b][0]), pk2(wt[b][1], wt[b][1]), pk2(wt[b][2], wt[b][2]), pk2(wt[b][3], wt[b][3])};
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            f32x2 acc = v == 0 ? 0ull : nkk[j];
#pragma unroll
            for (int t = 0; t < 4; ++t) acc = fma2p(q[b][t][j], w2[t], acc);
            d[j] = acc;
          }
        }
        float bacc = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) bacc = fmaf(bq[b][t], wt[b][t], bacc);
        float rgbl = 0.f, wv = 0.f;
        if (FULL || v < V) {   // warp-uniform
          const float cx = __shfl_sync(0xffffffffu, crgb.x, 4 * v), cy = __shfl_sync(0xffffffffu, crgb.y, 4 * v),
                      cz = __shfl_sync(0xffffffffu, crgb.z, 4 * v);
          const float4 r0 = *reinterpret_cast<const float4*>(ri0 + v * RI_N + RI_VIS);   // vis | w | rd0 | rd1
          const float2 r1 = *reinterpret_cast<const float2*>(ri0 + v * RI_N + RI_RD2);   // rd2 | rd3
          rgbl = lane == 0 ? cx : (lane == 1 ? cy : cz);
          wv = r0.y;
          if (lane == 0 && rv_ptr) __stcs(reinterpret_cast<float4*>(rv_ptr + v * 4), make_float4(cx, cy, cz, r0.x));
          if (with_blend) {
            float a = bacc;
            a = fmaf(cx, wb[0], a); a = fmaf(cy, wb[1], a); a = fmaf(cz, wb[2], a);
            a = fmaf(r0.x, wb[3], a);
            a = fmaf(r0.z, wb[4], a); a = fmaf(r0.w, wb[5], a); a = fmaf(r1.x, wb[6], a); a = fmaf(r1.y, wb[7], a);
            if (FULL) __stcs(pp_ptr + v * 4, a + bias);
            else __stcs(partial_out + partial_off(n * V + v, lane >> 2) + (lane & 3), a + bias);
          }
          if (mvf_out) {
            float* mo = mvf_out + (n * V + v) * C_RGBF;
            if (lane < 3) mo[lane] = rgbl;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              float2 f = up2(d[j]);
              if (v != 0) { const float2 k2 = up2(nkk[j]); f.x -= k2.x; f.y -= k2.y; }
              mo[3 + j * 64 + lane * 2] = f.x; mo[3 + j * 64 + lane * 2 + 1] = f.y;
            }
          }
        }
        if (v + 2 < 8) issue(b, p, v + 2);   // the buffer is free: request the view after next
        else if (has_next) issue(b, p + NT / 32, v + 2 - 8);
        sw += wv;
        if (v == 0) {
          rk = rgbl;
        } else {
          const float dr = rgbl - rk, wd = wv * dr;
          rsx += wd;
          rsxx = fmaf(wd, dr, rsxx);
        }
        if (v == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float2 k2 = up2(d[j]);
            nkk[j] = pk2(-k2.x, -k2.y);
            sx[j] = 0ull;
            sxx[j] = 0ull;
          }
        } else {
          const f32x2 wv2 = pk2(wv, wv);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const f32x2 wd = mul2p(wv2, d[j]);
            sx[j] = add2p(sx[j], wd);
            sxx[j] = fma2p(wd, d[j], sxx[j]);
          }
        }
      }
      // visibility-weighted mean / variance over the views (ibrnet.py:8-12)
      float* g = sG + p * LDGS - G0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float2 nk = up2(nkk[j]), s1 = up2(sx[j]), s2 = up2(sxx[j]);
        const float mx = fmaf(-nk.x, sw, s1.x), my = fmaf(-nk.y, sw, s1.y);
        const float cx = mx + nk.x, cy = my + nk.y;
        const float vx = fmaxf(fmaf(cx * cx, sw, fmaf(-2.f * cx, s1.x, s2.x)), 0.f);
        const float vy = fmaxf(fmaf(cy * cy, sw, fmaf(-2.f * cy, s1.y, s2.y)), 0.f);
        if (GOUT) {
          float* go = g_out + n * 416 + j * 64 + lane * 2;
          __stcs(reinterpret_cast<float2*>(go), make_float2(mx, my));
          __stcs(reinterpret_cast<float2*>(go + 192), make_float2(vx, vy));
        } else {
          const int c = 3 + j * 64 + lane * 2;
          g[c] = mx; g[c + 1] = my;
          g[C_RGBF + c] = vx; g[C_RGBF + c + 1] = vy;
        }
      }
      if (lane < 3) {
        const float m = fmaf(rk, sw, rsx), cm = m - rk;
        const float var = fmaxf(fmaf(cm * cm, sw, fmaf(-2.f * cm, rsx, rsxx)), 0.f);
        if (GOUT) {
          g_out[n * 416 + 384 + lane] = m;
          g_out[n * 416 + 387 + lane] = var;
        } else {
          g[lane] = m;
          g[C_RGBF + lane] = var;
        }
      }
      if (GOUT && lane < 26) g_out[n * 416 + 390 + lane] = lane < 3 ? g[390 + lane] : 0.f;   // extras (phase 4), zero padding up to 416
      }
    };
    if (V == 8) gather_samples(std::true_type{});
    else gather_samples(std::false_type{});
  } else {
  // ---- phase 5: rgb + 192-channel feature gather (zeros padding, align_corners=True), one warp per row ----------
  // A warp handles two rows per iteration and requests everything they need before it consumes any of it: 2 x 4 taps x 3
  // chunks of the feature map (24 independent 8-byte loads per lane), the rgb taps (lanes 0..3, one tap each) and, for
  // rendering, 4 taps of the pre-projected colour-blend channels; addresses are clamped into the maps and taps outside carry
  // weight 0.
  // Colour blend, per-view half of the first layer (model.py:532-535): that layer is linear, and so is the bilinear fetch:
  // W f(x) = sum_t b_t (W f_t), so the 192 map channels are pre-projected once per frame (sc.featb, [V][h][w][32]) and a row
  // gathers 32 more channels instead of running a [224 x 32] GEMM; the remaining inputs (rgb, visibility, ray difference)
  // are 8 FMAs per output.
  {
    float wb[8], bias = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) wb[i] = 0.f;
    if (with_blend) {
#pragma unroll
      for (int i = 0; i < 3; ++i) wb[i] = __ldg(w.bl1v + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < 5; ++i) wb[3 + i] = __ldg(w.bl1v + (195 + i) * 32 + lane);
      bias = __ldg(w.bl1_b + lane);
    }
    for (int rb = warp; rb < ROWS; rb += 2 * (NT / 32)) {
      float2 q[2][4][3];
      float wt[2][4], bq[2][4];
      float4 cq[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * (NT / 32);
        cq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          wt[u][t] = 0.f; bq[u][t] = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) q[u][t][j] = make_float2(0.f, 0.f);
        }
        if (r < rows) {
          const float* ri = sRI + r * RI_N;
          const int v = __float_as_int(ri[RI_V]);
          const int4 ti = *reinterpret_cast<const int4*>(ri + RI_TF);
          const float4 tw = *reinterpret_cast<const float4*>(ri + RI_TF + 4);
          const float* fb = sc.feat + ((size_t)v * sc.h * sc.w) * C_FEAT + lane * 2;
          const float* tp[4] = {fb + (size_t)ti.x * C_FEAT, fb + (size_t)ti.y * C_FEAT, fb + (size_t)ti.z * C_FEAT, fb + (size_t)ti.w * C_FEAT};
          const float twv[4] = {tw.x, tw.y, tw.z, tw.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            wt[u][t] = twv[t];
#pragma unroll
            for (int j = 0; j < 3; ++j) q[u][t][j] = __ldg(reinterpret_cast<const float2*>(tp[t] + j * 64));
          }
          if (lane < 4) {
            const float wi = ri[RI_TI + 4 + lane];
            if (wi != 0.f) {
              const int pix = __float_as_int(ri[RI_TI + lane]);
              const float4 c4 = __ldg(reinterpret_cast<const float4*>(sc.images + ((size_t)v * sc.H * sc.W + pix) * 4));
              cq[u] = make_float4(c4.x * wi, c4.y * wi, c4.z * wi, 0.f);
            }
          }
          if (with_blend) {
            const float* bb = sc.featb + ((size_t)v * sc.h * sc.w) * 32 + lane;
            bq[u][0] = __ldg(bb + (size_t)ti.x * 32);
            bq[u][1] = __ldg(bb + (size_t)ti.y * 32);
            bq[u][2] = __ldg(bb + (size_t)ti.z * 32);
            bq[u][3] = __ldg(bb + (size_t)ti.w * 32);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * (NT / 32);
        float* frow = sF + r * LDF;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            acc.x = fmaf(q[u][t][j].x, wt[u][t], acc.x);
            acc.y = fmaf(q[u][t][j].y, wt[u][t], acc.y);
          }
          // lanes 0-15 store their even channel first, lanes 16-31 their odd one: each store covers 32 distinct banks
          const int up = lane >> 4;
          frow[3 + j * 64 + lane * 2 + up] = up ? acc.y : acc.x;
          frow[3 + j * 64 + lane * 2 + (up ^ 1)] = up ? acc.x : acc.y;
        }
        if (r < rows) {   // warp-uniform
          float4 c = cq[u];
          c.x += __shfl_xor_sync(0xffffffffu, c.x, 1); c.y += __shfl_xor_sync(0xffffffffu, c.y, 1); c.z += __shfl_xor_sync(0xffffffffu, c.z, 1);
          c.x += __shfl_xor_sync(0xffffffffu, c.x, 2); c.y += __shfl_xor_sync(0xffffffffu, c.y, 2); c.z += __shfl_xor_sync(0xffffffffu, c.z, 2);
          c.x = __shfl_sync(0xffffffffu, c.x, 0); c.y = __shfl_sync(0xffffffffu, c.y, 0); c.z = __shfl_sync(0xffffffffu, c.z, 0);
          const float* ri = sRI + r * RI_N;
          const int p = __float_as_int(ri[RI_P]), v = __float_as_int(ri[RI_V]);
          if (lane == 0) {
            frow[0] = c.x; frow[1] = c.y; frow[2] = c.z;
            if (rgbvis_out) __stcs(reinterpret_cast<float4*>(rgbvis_out + (nidx(p) * V + v) * 4), make_float4(c.x, c.y, c.z, ri[RI_VIS]));
          }
          if (with_blend) {
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) a = fmaf(bq[u][t], wt[u][t], a);
            a = fmaf(c.x, wb[0], a); a = fmaf(c.y, wb[1], a); a = fmaf(c.z, wb[2], a);
            a = fmaf(ri[RI_VIS], wb[3], a);
            a = fmaf(ri[RI_RD0], wb[4], a); a = fmaf(ri[RI_RD1], wb[5], a); a = fmaf(ri[RI_RD2], wb[6], a); a = fmaf(ri[RI_RD3], wb[7], a);
            // per-chunk intermediates are written once and read once by the next kernel: streaming stores keep them from
            // evicting the reference feature maps (118 MB, gathered by every sample) out of the 126 MB L2
            __stcs(partial_out + partial_off(nidx(p) * V + v, lane >> 2) + (lane & 3), a + bias);
          }
        }
      }
    }
  }
  cta_sync();
  if (mvf_out) {
    for (int i = tid; i < rows * C_RGBF; i += NT) {
      const int r = i / C_RGBF, c = i - r * C_RGBF;
      mvf_out[(nidx(r / V) * V + (r - (r / V) * V)) * C_RGBF + c] = sF[r * LDF + c];
    }
  }

  AGG_STAMP(5);
  AGG_STAMP(6);
  // ---- phase 6: visibility-weighted mean / variance over views (ibrnet.py:8-12) ---------------------------------
  for (int p = warp; p < np; p += NT / 32) {
    if (V <= 8) mean_var_rows<8>(sF + (p * V) * LDF, sRI + (p * V) * RI_N, V, lane, sG + p * LDG);
    else mean_var_rows<16>(sF + (p * V) * LDF, sRI + (p * V) * RI_N, V, lane, sG + p * LDG);
  }
  }
  for (int i = tid; i < (TP_MAX - np) * LDGS; i += NT) sG[np * LDGS + i] = 0.f;

  AGG_STAMP(7);
  if (GOUT) { if (AGG_PERSIST) continue; else return; }   // (persistent: the next tile starts with a CTA-wide barrier)
  // ---- phase 7: out_fc 393 -> 64 -> 128 (ELU) --------------------------------------------------------------------
  cta_sync();  // sG complete
  rows16_gemm<64, TP_MAX, 416, 8>([&](int r, int) { return sG + r * LDG; }, w.fc1, 64, sB,
                  [&](int r, int c, float v) { sO1[r * 68 + c] = elu(v + __ldg(w.fc1_b + c)); });
  cta_sync();
  rows16_gemm<128, TP_MAX, 64, 8>([&](int r, int) { return sO1 + r * 68; }, w.fc2, 128, sB, [&](int r, int c, float v) {
    if (r < np) __stcs(agg_out + nidx(r) * W_HID + c, elu(v + __ldg(w.fc2_b + c)));
  });

  AGG_STAMP(8);
  }
}

// Per-frame pre-projection of the reference feature maps through the map-feature rows of the colour-blend first layer
// (rgb_blending_mlp.0, model.py:90-96,532-535): out[p][n] = sum_c feat[p][c] * Wt[3 + c][n], Wt = RenderW::bl1v [224][32].
// One warp per pixel, lane = output channel.
__global__ void __launch_bounds__(256) blend_project_kernel(const float* __restrict__ feat, const int64_t P,
                                                            const float* __restrict__ Wt, float* __restrict__ out) {
  __shared__ float sW[C_FEAT * 32];
  for (int i = threadIdx.x; i < C_FEAT * 32; i += blockDim.x) sW[i] = Wt[3 * 32 + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < P; p += (int64_t)gridDim.x * 8) {
    float f[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) f[j] = __ldg(feat + p * C_FEAT + j * 32 + lane);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int l = 0; l < 32; ++l) a = fmaf(__shfl_sync(0xffffffffu, f[j], l), sW[(j * 32 + l) * 32 + lane], a);
    out[p * 32 + lane] = a;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// generic row-tile Linear: out[n][0..Nout) = act(A[n][0..K) Wt + b), Wt [Kp][Nout] (Kp = K rounded up to 32)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
linear_kernel(const float* __restrict__ A, const int64_t N, const int K, const int lda, const float* __restrict__ Wt,
              const float* __restrict__ bias, const int Nout, const int act, float* __restrict__ out, const int ldo) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;
  float* sA = sB + STAGE_FLOATS;
  const int Kp = (K + 31) / 32 * 32;
  const int ld = Kp + 4;
  const int64_t n0 = (int64_t)blockIdx.x * 128;
  const int nr = (int)min((int64_t)128, N - n0);
  for (int i = threadIdx.x; i < 128 * Kp; i += NT) {
    const int r = i / Kp, k = i - r * Kp;
    sA[r * ld + k] = (r < nr && k < K) ? A[(n0 + r) * lda + k] : 0.f;
  }
  for (int c0 = 0; c0 < Nout; c0 += 64) {
    tile_gemm<4, 8, 64, false>(plainA(sA, ld), 128, Wt + c0, Nout, Kp, sB, [&](int r, int c, float v) {
      if (r < nr) {
        float o = v + (bias ? __ldg(bias + c0 + c) : 0.f);
        if (act == 1) o = leaky(o);
        else if (act == 2) o = 1.f / (1.f + expf(-o));
        out[(n0 + r) * ldo + c0 + c] = o;
      }
    });
  }
}

// support-point geometry pack: [M][8] = xyz, direction(3), confidence, 0
__global__ void sup_geo_kernel(const float* __restrict__ xyz, const float* __restrict__ dir, const float* __restrict__ conf,
                               int64_t M, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float4 a = make_float4(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], dir[i * 4]);
  float4 b = make_float4(dir[i * 4 + 1], dir[i * 4 + 2], conf[i], 0.f);
  *reinterpret_cast<float4*>(out + i * 8) = a;
  *reinterpret_cast<float4*>(out + i * 8 + 4) = b;
}

// ------------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------------
template <class Kern>
static int set_smem(Kern k, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : set_error(cudaGetErrorString(e));
}

int launch_aggregate(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, int with_blend, float* agg,
                     float* partial, float* rgbvis, unsigned char* nvalid, float* mvf, float* mvv, float* visdd_scratch,
                     float* g_scratch, float* q_out, cudaStream_t st, bool agg_pm) {
  if (N <= 0) return 0;
  if (sc.V < 1 || sc.V > 16) return set_error("aggregate: number of reference views must be in 1..16");
  if (with_blend && !sc.featb) return set_error("aggregate: scene.featmaps_blend is NULL (call nlb_blend_prepare once per frame)");
  constexpr int ROWS = AGG_ROWS;
  const int TP = ROWS / 8 < ROWS / sc.V ? ROWS / 8 : ROWS / sc.V;
  if (!ps.xyz && (ps.S < 1 || N % ps.S != 0)) return set_error("aggregate: ray samples must come as whole rays");
  const int64_t tiles = ps.xyz ? (N + TP - 1) / TP : ((N / ps.S + TP - 1) / TP) * ps.S;
  if (tiles > 0x7fffffffLL) return set_error("aggregate: too many tiles for one launch");
  const unsigned grid = (unsigned)tiles;
  static const bool unfused = getenv("NLB_AGG_UNFUSED") != nullptr;   // A/B switches
  static const bool one_kernel = getenv("NLB_AGG_V1") != nullptr;      // visibility decoder inside aggregate_kernel (mma.sync)
  const bool ext = !one_kernel && visdd_scratch != nullptr;
  if (ext) {
    if (launch_visibility(sc, w, ps, N, visdd_scratch, mvv, st)) return 1;
    prof_mark("visibility");
  }
  const float2* vd = reinterpret_cast<const float2*>(visdd_scratch);
  static const bool fc_inside = getenv("NLB_AGG_FC_V1") != nullptr;    // out_fc inside aggregate_kernel (FFMA2 small-M GEMMs)
  if (sc.V <= 8 && !unfused) {
    const size_t smem = agg_smem_floats<ROWS, true>() * sizeof(float);
    if (ext && g_scratch && q_out && !fc_inside) {
      // render variant: 32 samples (the same sample index of 32 adjacent rays) per tile - the serial per-tile phases (projection,
      // view weights, barriers) cost little more for 256 rows than for 64 - in 40 KB of shared memory, two CTAs per SM
      constexpr int ROWS_G = 256;
      const int TPG = ROWS_G / 8 < ROWS_G / sc.V ? ROWS_G / 8 : ROWS_G / sc.V;
      const int64_t tiles_g = ps.xyz ? (N + TPG - 1) / TPG : ((N / ps.S + TPG - 1) / TPG) * ps.S;
      const size_t smem_p = agg_smem_floats<ROWS_G, true, true>() * sizeof(float);
      if (set_smem(aggregate_kernel<ROWS_G, true, true, true>, smem_p)) return 1;
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const unsigned pgrid = AGG_PERSIST ? (unsigned)(tiles_g < 2 * (int64_t)sms ? tiles_g : 2 * (int64_t)sms) : (unsigned)tiles_g;
      aggregate_kernel<ROWS_G, true, true, true><<<pgrid, NT, smem_p, st>>>(sc, w, ps, N, with_blend, vd, g_scratch, agg, partial, rgbvis, nvalid, mvf, mvv);
      if (check_launch("aggregate_kernel")) return 1;
      prof_mark("aggregate");
      if (launch_fc_tail(w, g_scratch, N, agg, q_out, agg_pm, st)) return 1;
      prof_mark("fc_tail");
      return 2;   // aggregated AND the attention query are done
    } else if (ext) {
      if (set_smem(aggregate_kernel<ROWS, true, true>, smem)) return 1;
      aggregate_kernel<ROWS, true, true><<<grid, NT, smem, st>>>(sc, w, ps, N, with_blend, vd, nullptr, agg, partial, rgbvis, nvalid, mvf, mvv);
    } else {
      if (set_smem(aggregate_kernel<ROWS, true, false>, smem)) return 1;
      aggregate_kernel<ROWS, true, false><<<grid, NT, smem, st>>>(sc, w, ps, N, with_blend, nullptr, nullptr, agg, partial, rgbvis, nvalid, mvf, mvv);
    }
  } else {
    const size_t smem = agg_smem_floats<ROWS, false>() * sizeof(float);
    if (ext) {
      if (set_smem(aggregate_kernel<ROWS, false, true>, smem)) return 1;
      aggregate_kernel<ROWS, false, true><<<grid, NT, smem, st>>>(sc, w, ps, N, with_blend, vd, nullptr, agg, partial, rgbvis, nvalid, mvf, mvv);
    } else {
      if (set_smem(aggregate_kernel<ROWS, false, false>, smem)) return 1;
      aggregate_kernel<ROWS, false, false><<<grid, NT, smem, st>>>(sc, w, ps, N, with_blend, nullptr, nullptr, agg, partial, rgbvis, nvalid, mvf, mvv);
    }
  }
  if (check_launch("aggregate_kernel")) return 1;
  prof_mark("aggregate");
  return 0;
}

int read_prof(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_prof, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : set_error("read_prof failed");
}

int launch_blend_project(const float* feat, int64_t P, const float* bl1v, float* out, cudaStream_t st) {
  if (P <= 0) return 0;
  const int64_t blocks = (P + 7) / 8;
  blend_project_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(feat, P, bl1v, out);
  return check_launch("blend_project_kernel");
}

int launch_linear(const float* A, int64_t N, int K, int lda, const float* Wt, const float* bias, int Nout, int act,
                  float* out, int ldo, cudaStream_t st) {
  if (N <= 0) return 0;
  if (Nout % 64 != 0) return set_error("linear: Nout must be a multiple of 64");
  const int Kp = (K + 31) / 32 * 32;
  if (Kp > 352) return set_error("linear: K too large for one shared-memory tile");
  const size_t smem = (STAGE_FLOATS + 128 * (Kp + 4)) * sizeof(float);
  if (set_smem(linear_kernel, smem)) return 1;
  linear_kernel<<<(unsigned)((N + 127) / 128), NT, smem, st>>>(A, N, K, lda, Wt, bias, Nout, act, out, ldo);
  return check_launch("linear_kernel");
}

int launch_sup_geo(const float* xyz, const float* dir, const float* conf, int64_t M, float* out, cudaStream_t st) {
  if (M <= 0) return 0;
  sup_geo_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(xyz, dir, conf, M, out);
  return check_launch("sup_geo_kernel");
}

}  // namespace nlb
