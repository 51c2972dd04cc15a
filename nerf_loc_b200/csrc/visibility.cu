// visibility_kernel: the per-(sample, view) half of MultiviewFeatureAggregator.forward (conditional_nerf/
// multiview_aggregator.py:156-222) - NeuRay projection (conditional_nerf/depth_fusion.py:78-147), bilinear fetch of the 32
// DepthFusionNet channels, the mixture-of-logistics visibility decoder (visibility_decoder.py:62-148) and the visibility /
// depth-difference outputs - split off aggregate_kernel so that the decoder runs on tcgen05.
//
// A tile is 128 (sample, view) rows = 128 / V samples x V views; persistent CTAs, two per SM (256 TMEM columns and 66 KB of
// shared memory each), so that one CTA's gather latency and MMA waits are covered by the other's epilogues:
//   gather     one warp per row, lane = channel (a row's four taps are four coalesced 128-byte reads), eight rows of a warp in
//              flight at a time, interpolated into the bf16 hi | lo layer-1 operand (chunk-major shared memory tile);
//   layer 1    [128 x 32] x [32 x 128] (the four heads side by side), bf16x3, accumulator in tensor memory;
//   E1         + bias, ELU, split -> layer-2 operand in tensor memory (two bf16 per column);
//   layer 2    block diagonal: four [128 x 32] x [32 x 32] products into the accumulator columns of their head;
//   E2         + bias, ELU, the six head outputs (32-long dot products), softplus / sigmoid, visibility and depth difference.
// aggregate_kernel spent 9-13 k of its 39 k clk per 64 rows on this decoder (1,536 mma.sync + operand splits per tile, issue
// bound); here it is 48 tcgen05.mma per 128 rows and two TMEM epilogues.
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"
#include "tc_bf16.cuh"
#include "tc_pipe.cuh"

namespace nlb {
namespace vis {

__device__ long long g_prof_vis[16];
#define VIS_STAMP(i) do { if (stamp) g_prof_vis[i] = clock64(); } while (0)

constexpr int RA = 128;
// tensor-memory map (columns)
constexpr uint32_t TM_D = 0, TM_AHI = 128, TM_ALO = 192;
// shared-memory map (bytes)
constexpr uint32_t W1_OFF = 0;                       // dec1 [128 x 32]: hi 8 KB | lo 8 KB
constexpr uint32_t W2_OFF = 16384;                   // dec2: 4 heads x ([32 x 32] hi 2 KB | lo 2 KB)
constexpr uint32_t A1_OFF = 32768;                   // layer-1 operand [128 x 32], chunk-major: hi 8 KB | lo 8 KB
constexpr uint32_t TAP_OFF = A1_OFF + 16384;         // [2][128][12]: 4 pixel indices, 4 weights, valid, depth, -, -
constexpr int TAP_LD = 12;
constexpr uint32_t SMALL_OFF = TAP_OFF + 2 * 128 * TAP_LD * 4;   // dec1_b [128] | dec2_b [128] | dec3 [6][32] | dec3_b [8]
constexpr uint32_t O1_OFF = SMALL_OFF + (128 + 128 + 192 + 8) * 4;   // head outputs [128][6]
constexpr uint32_t SYNC_OFF = O1_OFF + 128 * 6 * 4;
constexpr uint32_t SMEM_BYTES = SYNC_OFF + 64;

struct Sync {
  uint64_t a_ready, d_ready;
  uint32_t tmem_slot;
};

__device__ __forceinline__ int tile_rows_samples(int V) { return 128 / V; }

__global__ void __launch_bounds__(NT + 32, 2)
visibility_kernel(const SceneDev sc, const RenderW w, const PointSrc ps, const int64_t N, float2* __restrict__ visdd_out,
                  float* __restrict__ mvv_out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* sTap = reinterpret_cast<float*>(sm + TAP_OFF);
  float* sSmall = reinterpret_cast<float*>(sm + SMALL_OFF);
  float* sB1 = sSmall, *sB2 = sSmall + 128, *sW3 = sSmall + 256, *sB3 = sSmall + 448;
  float* sO1 = reinterpret_cast<float*>(sm + O1_OFF);
  Sync& sy = *reinterpret_cast<Sync*>(sm + SYNC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = sc.V;
  const int TP = 128 / V;                       // samples per tile
  const int rows_full = TP * V;
  // tiles: explicit point lists take TP consecutive points; ray samples the same sample index of TP consecutive rays (see
  // aggregate_kernel: neighbouring rays at one depth hit the same pixels of a reference view)
  const bool by_ray = ps.xyz == nullptr;
  const int64_t n_items = by_ray ? N / ps.S : N;
  const int64_t groups = (n_items + TP - 1) / TP;
  const int64_t ntiles = by_ray ? groups * ps.S : groups;
  if (warp == 8) {
    tc::tmem_alloc(&sy.tmem_slot, 256);
    if (lane == 0) { tc::mbar_init(&sy.a_ready, NT); tc::mbar_init(&sy.d_ready, 1); }
  }
  if (tid < NT) {
    // decoder weights: the packed bf16 hi | lo tiles are copied as they are
    const uint4* g1 = reinterpret_cast<const uint4*>(w.tb_dec1);
    const uint4* g2 = reinterpret_cast<const uint4*>(w.tb_dec2);
    for (int i = tid; i < 1024; i += NT) reinterpret_cast<uint4*>(sm + W1_OFF)[i] = __ldg(g1 + i);
    for (int i = tid; i < 1024; i += NT) reinterpret_cast<uint4*>(sm + W2_OFF)[i] = __ldg(g2 + i);
    for (int i = tid; i < 128; i += NT) { sB1[i] = __ldg(w.dec1_b + i); sB2[i] = __ldg(w.dec2_b + i); }
    for (int i = tid; i < 192; i += NT) sW3[i] = __ldg(w.dec3 + i);
    if (tid < 6) sB3[tid] = __ldg(w.dec3_b + tid);
    tc::fence_async_smem();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sy.tmem_slot;
  const int nmy = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp >= 8) {
    {
      // ------------------------------------------------ MMA issuer ---------------------------------------------------------------
      uint32_t a_par = 0;
      const uint32_t a_hi32 = tc::desc_hi(128u);
      const uint32_t w1 = tc::smem_u32(sm + W1_OFF), w2 = tc::smem_u32(sm + W2_OFF), a1 = tc::smem_u32(sm + A1_OFF);
      for (int t = 0; t < nmy; ++t) {
        // layer 1: A from shared memory (chunk-major, LBO = RA * 16), B = dec1 tile (KT = 32)
        tc::mbar_wait(&sy.a_ready, a_par); a_par ^= 1u;
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::idesc_bf16(128, 128);
          const uint32_t b_hi32 = tc::desc_hi(32u * 16u);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {                  // lo*hi, hi*lo, hi*hi
            const uint32_t ab = a1 + (pass == 0 ? 8192u : 0u);
            const uint32_t bb = w1 + (pass == 1 ? 8192u : 0u);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_bf16_w(tmem + TM_D, tc::desc_lo(ab + (uint32_t)ks * 2u * RA * 16u, RA * 16u), a_hi32, tc::desc_lo(bb + (uint32_t)ks * 256u, 128u), b_hi32,
                             idesc, pass > 0 || ks > 0);
          }
          tc::mma_commit(&sy.d_ready);
        }
        __syncwarp();
        // layer 2: A from tensor memory (head h: packed columns 16 h ..), B = dec2 head h, D columns 32 h ..
        tc::mbar_wait(&sy.a_ready, a_par); a_par ^= 1u;
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::idesc_bf16(128, 32);
          const uint32_t b_hi32 = tc::desc_hi(32u * 16u);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = tmem + (pass == 0 ? TM_ALO : TM_AHI) + (uint32_t)(h * 16);
              const uint32_t bb = w2 + (uint32_t)h * 4096u + (pass == 1 ? 2048u : 0u);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_bf16_ts_w(tmem + TM_D + (uint32_t)(h * 32), a + (uint32_t)ks * 8u, tc::desc_lo(bb + (uint32_t)ks * 256u, 128u), b_hi32, idesc,
                                  pass > 0 || ks > 0);
            }
          }
          tc::mma_commit(&sy.d_ready);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ compute warps ------------------------------------------------------
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float near_ = sc.near_, far_ = sc.far_;
    uint32_t d_par = 0;
    auto wait_d = [&]() { tc::mbar_wait(&sy.d_ready, d_par); d_par ^= 1u; tc::fence_after_sync(); };
    auto tile_of = [&](int it, int64_t& g, int& s) {
      const int64_t b = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      g = by_ray ? b / ps.S : b;
      s = by_ray ? (int)(b - g * ps.S) : 0;
    };
    auto nidx = [&](int64_t g, int s, int p) -> int64_t { return by_ray ? (g * TP + p) * ps.S + s : g * TP + p; };

    // NeuRay projection of this thread's row (threads 0..127) of tile `it` -> tap record
    auto project = [&](int it) {
      int64_t g; int s;
      tile_of(it, g, s);
      const int np = (int)min((int64_t)TP, n_items - g * TP);
      float* tp = sTap + (it & 1) * 128 * TAP_LD;
      if (tid < 128) {
        float* ri = tp + tid * TAP_LD;
        const int p = tid / V, v = tid - p * V;
        if (tid < rows_full && p < np) {
          const int64_t n = nidx(g, s, p);
          float x, y, z;
          if (ps.xyz) {
            x = ps.xyz[n * 3]; y = ps.xyz[n * 3 + 1]; z = ps.xyz[n * 3 + 2];
          } else {
            const int64_t r = g * TP + p;
            const float t = ps.z[r * ps.zs + s];
            x = __fadd_rn(ps.rays_o[r * 3 + 0], __fmul_rn(ps.rays_d[r * 3 + 0], t));
            y = __fadd_rn(ps.rays_o[r * 3 + 1], __fmul_rn(ps.rays_d[r * 3 + 1], t));
            z = __fadd_rn(ps.rays_o[r * 3 + 2], __fmul_rn(ps.rays_d[r * 3 + 2], t));
          }
          const float* kr = sc.cams + v * 32 + 12;
          const float c0 = fmaf(kr[2], z, fmaf(kr[1], y, kr[0] * x)) + kr[3];
          const float c1 = fmaf(kr[6], z, fmaf(kr[5], y, kr[4] * x)) + kr[7];
          float dep = fmaf(kr[10], z, fmaf(kr[9], y, kr[8] * x)) + kr[11];
          const bool bad = fabsf(dep) < 1e-4f;
          if (bad) dep = 1e-3f;
          const float qx = c0 / dep, qy = c1 / dep;
          const bool outside = qx < -0.5f || qx >= (float)sc.W - 0.5f || qy < -0.5f || qy >= (float)sc.H - 0.5f;
          const float xn = qx / (float)(sc.W - 1) * 2.f - 1.f, yn = qy / (float)(sc.H - 1) * 2.f - 1.f;
          float vx, vy;
          if (sc.vh == sc.H && sc.vw == sc.W) {  // align_corners=True only when the map has the image size
            vx = ((xn + 1.f) / 2.f) * (float)(sc.vw - 1);
            vy = ((yn + 1.f) / 2.f) * (float)(sc.vh - 1);
          } else {
            vx = ((xn + 1.f) * (float)sc.vw - 1.f) / 2.f;
            vy = ((yn + 1.f) * (float)sc.vh - 1.f) / 2.f;
          }
          // torch grid_sample, bilinear, padding_mode='border': coordinate clipped first
          const float ix = fminf((float)(sc.vw - 1), fmaxf(vx, 0.f)), iy = fminf((float)(sc.vh - 1), fmaxf(vy, 0.f));
          const float fx = floorf(ix), fy = floorf(iy);
          const int x0i = (int)fminf(fmaxf(fx, -2.f), (float)sc.vw), y0i = (int)fminf(fmaxf(fy, -2.f), (float)sc.vh);
          const float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy, wx = ix - fx, wy = iy - fy;
          float tw[4] = {ex * ey, wx * ey, ex * wy, wx * wy};
          const bool inx0 = fx >= 0.f && fx <= (float)(sc.vw - 1), inx1 = fx + 1.f >= 0.f && fx + 1.f <= (float)(sc.vw - 1);
          const bool iny0 = fy >= 0.f && fy <= (float)(sc.vh - 1), iny1 = fy + 1.f >= 0.f && fy + 1.f <= (float)(sc.vh - 1);
          if (!(inx0 && iny0)) tw[0] = 0.f;
          if (!(inx1 && iny0)) tw[1] = 0.f;
          if (!(inx0 && iny1)) tw[2] = 0.f;
          if (!(inx1 && iny1)) tw[3] = 0.f;
          const int xa = min(max(x0i, 0), sc.vw - 1), xb = min(max(x0i + 1, 0), sc.vw - 1);
          const int ya = min(max(y0i, 0), sc.vh - 1), yb = min(max(y0i + 1, 0), sc.vh - 1);
          const int vbase = v * sc.vh * sc.vw;
          *reinterpret_cast<int4*>(ri) = make_int4(vbase + ya * sc.vw + xa, vbase + ya * sc.vw + xb, vbase + yb * sc.vw + xa, vbase + yb * sc.vw + xb);
          *reinterpret_cast<float4*>(ri + 4) = make_float4(tw[0], tw[1], tw[2], tw[3]);
          ri[8] = (!bad && !outside) ? 1.f : 0.f;
          ri[9] = dep;
          ri[10] = 1.f;   // live
        } else {
          *reinterpret_cast<int4*>(ri) = make_int4(0, 0, 0, 0);
          *reinterpret_cast<float4*>(ri + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          ri[8] = 0.f; ri[9] = 1.f; ri[10] = 0.f;
        }
      }
      cta_sync();
    };

    for (int it = 0; it < nmy; ++it) {
      const bool stamp = it == 1 && blockIdx.x == gridDim.x / 2 && tid == 0;
      VIS_STAMP(0);
      int64_t g; int s;
      tile_of(it, g, s);
      const int np = (int)min((int64_t)TP, n_items - g * TP);
      const float* tp = sTap + (it & 1) * 128 * TAP_LD;
      project(it);
      VIS_STAMP(1);
      // ---- X: gather the taps of this warp's 16 rows (lane = channel), eight rows in flight, and interpolate -> layer-1 operand
      // (bf16 hi | lo, chunk-major).  Pairs of lanes are packed with a shuffle and the even lane stores 4 bytes per plane.
#pragma unroll
      for (int ub = 0; ub < 16; ub += 8) {
        float q[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int4 ti = *reinterpret_cast<const int4*>(tp + (warp * 16 + ub + u) * TAP_LD);
          q[u][0] = __ldg(sc.vis + (size_t)ti.x * C_VIS + lane);
          q[u][1] = __ldg(sc.vis + (size_t)ti.y * C_VIS + lane);
          q[u][2] = __ldg(sc.vis + (size_t)ti.z * C_VIS + lane);
          q[u][3] = __ldg(sc.vis + (size_t)ti.w * C_VIS + lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = warp * 16 + ub + u;
          const float4 tw = *reinterpret_cast<const float4*>(tp + r * TAP_LD + 4);
          float a = q[u][0] * tw.x;
          a += q[u][1] * tw.y;
          a += q[u][2] * tw.z;
          a += q[u][3] * tw.w;
          a *= tp[r * TAP_LD + 8];   // valid
          const float b = __shfl_down_sync(0xffffffffu, a, 1);
          if ((lane & 1) == 0) {
            uint32_t hi, lo;
            tc::split_bf16x2(a, b, hi, lo);
            const uint32_t o = tc::cm_off(r, lane, RA);
            *reinterpret_cast<uint32_t*>(sm + A1_OFF + o) = hi;
            *reinterpret_cast<uint32_t*>(sm + A1_OFF + 8192 + o) = lo;
          }
        }
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      tc::mbar_arrive(&sy.a_ready);
      VIS_STAMP(2);
      // ---- E1: + bias, ELU -> layer-2 operand in tensor memory
      wait_d();
      {
        const int c0 = half * 64;
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32];
          tc::tmem_ld32(trow + TM_D + (uint32_t)(c0 + cc), v);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            tc::split_bf16x2(elu(v[2 * j] + sB1[c0 + cc + 2 * j]), elu(v[2 * j + 1] + sB1[c0 + cc + 2 * j + 1]), hi[j], lo[j]);
          tc::tmem_st16_u(trow + TM_AHI + (uint32_t)((c0 + cc) / 2), hi);
          tc::tmem_st16_u(trow + TM_ALO + (uint32_t)((c0 + cc) / 2), lo);
        }
        tc::tmem_st_wait();
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&sy.a_ready);
      VIS_STAMP(3);
      // ---- E2: + bias, ELU, head outputs.  Half 0 holds the mean and scale heads (outputs 0-3), half 1 the mixture-weight and
      // visibility-scale heads (outputs 4, 5)
      wait_d();
      {
        const int c0 = half * 64;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tc::tmem_ld32(trow + TM_D + (uint32_t)(c0 + hh * 32), v);
          // outputs of this head: half 0 -> head hh has outputs 2 hh, 2 hh + 1; half 1 -> head 2 + hh has output 4 + hh
          const float* w3a = sW3 + (half == 0 ? (2 * hh) * 32 : (4 + hh) * 32);
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 wa = *reinterpret_cast<const float4*>(w3a + j);
            const float4 wb = *reinterpret_cast<const float4*>(w3a + 32 + j);   // (only used by half 0)
            const float e0 = elu(v[j] + sB2[c0 + hh * 32 + j]), e1 = elu(v[j + 1] + sB2[c0 + hh * 32 + j + 1]);
            const float e2 = elu(v[j + 2] + sB2[c0 + hh * 32 + j + 2]), e3 = elu(v[j + 3] + sB2[c0 + hh * 32 + j + 3]);
            a0 = fmaf(e0, wa.x, a0); a0 = fmaf(e1, wa.y, a0); a0 = fmaf(e2, wa.z, a0); a0 = fmaf(e3, wa.w, a0);
            a1 = fmaf(e0, wb.x, a1); a1 = fmaf(e1, wb.y, a1); a1 = fmaf(e2, wb.z, a1); a1 = fmaf(e3, wb.w, a1);
          }
          o[2 * hh] = a0; o[2 * hh + 1] = a1;
        }
        if (half == 0) {
          sO1[row * 6 + 0] = softplus_fast(o[0] + sB3[0]);
          sO1[row * 6 + 1] = softplus_fast(o[1] + sB3[1]);
          sO1[row * 6 + 2] = softplus_fast(o[2] + sB3[2]) + 0.05f;
          sO1[row * 6 + 3] = softplus_fast(o[3] + sB3[3]) + 0.05f;
        } else {
          sO1[row * 6 + 4] = sigmoid_fast(o[0] + sB3[4]);
          sO1[row * 6 + 5] = sigmoid_fast(o[2] + sB3[5]);
        }
      }
      tc::fence_before_sync();
      cta_sync();
      VIS_STAMP(4);
      // ---- per-row tail (visibility_decoder.py:99-148, multiview_aggregator.py:95-154)
      if (tid < 128) {
        const float* ri = tp + tid * TAP_LD;
        const int p = tid / V, v = tid - p * V;
        if (tid < rows_full && p < np) {
          const float m0 = sO1[tid * 6], m1 = sO1[tid * 6 + 1], v0 = sO1[tid * 6 + 2], v1 = sO1[tid * 6 + 3];
          const float aw = sO1[tid * 6 + 4], vs = sO1[tid * 6 + 5];
          const float dep = ri[9];
          const float near_inv = -1.f / near_, far_inv = -1.f / far_;
          float refd = __fdividef(-1.f, m0 * (far_inv - near_inv) + near_inv);
          refd = fminf(fmaxf(refd, near_), far_);
          const float dd = __fdividef(fabsf(dep - refd), far_ - near_);
          const float dn = __fdividef(__fdividef(-1.f, fmaxf(dep, 1e-5f)) - near_inv, far_inv - near_inv);
          const float cdf0 = (0.5f + 0.5f * tanh_fast((dn - m0) * v0)) * vs;
          const float cdf1 = (0.5f + 0.5f * tanh_fast((dn - m1) * v1)) * vs;
          const float visv = ((1.f - cdf0) * aw + (1.f - cdf1) * (1.f - aw)) * ri[8];
          const int64_t n = nidx(g, s, p);
          visdd_out[n * V + v] = make_float2(visv, dd);
          if (mvv_out) mvv_out[n * V + v] = visv;
        }
      }
      cta_sync();   // sO1 and this tile's tap records are free
      VIS_STAMP(5);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 256);
  }
}

}  // namespace vis

int read_prof_vis(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, vis::g_prof_vis, sizeof(long long) * (n < 16 ? n : 16)) == cudaSuccess ? 0 : set_error("read_prof_vis failed");
}

int launch_visibility(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, float* visdd, float* mvv, cudaStream_t st) {
  if (N <= 0) return 0;
  if (sc.V < 1 || sc.V > 16) return set_error("visibility: number of reference views must be in 1..16");
  if (!ps.xyz && (ps.S < 1 || N % ps.S != 0)) return set_error("visibility: ray samples must come as whole rays");
  cudaError_t e = cudaFuncSetAttribute(vis::visibility_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vis::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const int TP = 128 / sc.V;
  const int64_t items = ps.xyz ? N : N / ps.S;
  const int64_t tiles = ps.xyz ? (items + TP - 1) / TP : ((items + TP - 1) / TP) * ps.S;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(tiles < 2 * sms ? tiles : 2 * sms);   // persistent, two CTAs per SM
  vis::visibility_kernel<<<grid, NT + 32, vis::SMEM_BYTES, st>>>(sc, w, ps, N, reinterpret_cast<float2*>(visdd), mvv);
  return check_launch("visibility_kernel");
}

}  // namespace nlb
