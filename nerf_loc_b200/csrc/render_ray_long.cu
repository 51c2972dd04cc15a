// Per-ray stage of ConditionalNeRF.render_rays for rays with MORE than 128 samples (128 < S <= 256; BASELINE configs 4 and 5:
// S = 192, 256).  The tensor-core ray_kernel of render_ray.cu maps one ray onto one 128-row tcgen05 tile and keeps every
// activation in shared memory; a 256-sample ray does not fit either.  This kernel keeps the same layer order and the same
// algebra (conditional_nerf/model.py:521-598, ray_unet.py:5-69) but
//   * parks the [S][C] activations of a ray in a per-CTA global scratch slab (L2 resident, ~330 KB per CTA at S = 256),
//   * runs every convolution as a 3-tap fp32 tile GEMM (nlb::tile_gemm, A fragments through L1) that may take several row
//     passes, writes the pre-LayerNorm values to a raw slab and applies the joint LayerNorm([C, S_level]) as a two-pass
//     reduction over that slab,
//   * loops persistently over rays (grid = a multiple of the SM count), so the scratch does not grow with the ray count.
// Reads of the scratch use plain generic loads (never ld.global.nc): the data is produced inside this kernel.
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"

namespace nlb {

constexpr int RL_LDX = 164;   // [x 128 | x2 32] + 4
constexpr int RL_LDC1 = 132;  // [c1 64 | x1 64] + 4
constexpr int RL_LDC2 = 260;  // [c2 128 | x0 128] + 4
constexpr int RL_LDC3 = 132;  // [c3 128] + 4
constexpr int RL_MAX_S = 256;

size_t ray_long_slab_floats(int S) {
  const size_t n = (size_t)(S + 2) * RL_LDX + (size_t)(S / 2 + 2) * RL_LDC1 + (size_t)(S / 4 + 2) * RL_LDC2 +
                   (size_t)(S / 8 + 2) * RL_LDC3 + (size_t)S * 128;
  return (n + 63) / 64 * 64;
}

int ray_long_grid(int64_t R) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t g = (int64_t)sms * 2;
  if (g > RL_MAX_GRID) g = RL_MAX_GRID;
  return (int)(R < g ? R : g);
}

// mean / rstd of a raw [rows][cols] slab (contiguous), two-pass like torch.nn.LayerNorm
__device__ __forceinline__ void slab_stats(const float* raw, int n, float* red, float& mean, float& rstd) {
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < n; i += NT * 4) {
    const float4 v = *reinterpret_cast<const float4*>(raw + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  mean = block_sum(s, red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x * 4; i < n; i += NT * 4) {
    const float4 v = *reinterpret_cast<const float4*>(raw + i);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  rstd = 1.f / sqrtf(block_sum(q, red) / (float)n + 1e-5f);
}

// Encoder block: Conv1d k3 (+bias) -> raw, LayerNorm + ELU + MaxPool1d(2) -> dst rows [0, rows/2), columns [0, COLS)
template <int COLS>
__device__ __forceinline__ void enc_long(const ASrc A, int rows, const UnetLayer& L, int cin, float* sB, float* red, float* raw,
                                         float* dst, int ldd) {
  tile_gemm<4, 8, COLS, false>(A, rows, L.w, COLS, 3 * cin, sB,
                               [&](int r, int c, float v) { raw[r * COLS + c] = v + __ldg(L.b + c); });
  cta_sync();
  float mean, rstd;
  slab_stats(raw, rows * COLS, red, mean, rstd);
  for (int i = threadIdx.x; i < (rows / 2) * COLS; i += NT) {
    const int r2 = i / COLS, c = i - r2 * COLS;
    const int i0 = (2 * r2) * COLS + c, i1 = i0 + COLS;
    const float y0 = elu((raw[i0] - mean) * rstd * __ldg(L.g + i0) + __ldg(L.be + i0));
    const float y1 = elu((raw[i1] - mean) * rstd * __ldg(L.g + i1) + __ldg(L.be + i1));
    dst[r2 * ldd + c] = fmaxf(y0, y1);
  }
  cta_sync();
}

// Decoder block: stride-2 ConvTranspose1d k3 -> raw rows [0, 2*rows), LayerNorm + ELU -> dst columns [0, COLS)
template <int COLS>
__device__ __forceinline__ void dec_long(const float* in, int ldi, int rows, const UnetLayer& L, int cin, float* sB, float* red,
                                         float* raw, float* dst, int ldd) {
  tile_gemm<2, 8, COLS, false>(ASrc{in, ldi, cin, 0, 0, 0}, rows, L.w, COLS, cin, sB,
                               [&](int r, int c, float v) { raw[(2 * r) * COLS + c] = v + __ldg(L.b + c); });
  tile_gemm<2, 8, COLS, false>(ASrc{in, ldi, cin, 0, 1, 0}, rows, L.w + (size_t)cin * COLS, COLS, 2 * cin, sB,
                               [&](int r, int c, float v) { raw[(2 * r + 1) * COLS + c] = v + __ldg(L.b + c); });
  cta_sync();
  float mean, rstd;
  slab_stats(raw, 2 * rows * COLS, red, mean, rstd);
  for (int i = threadIdx.x; i < 2 * rows * COLS; i += NT) {
    const int r = i / COLS, c = i - r * COLS;
    dst[r * ldd + c] = elu((raw[i] - mean) * rstd * __ldg(L.g + i) + __ldg(L.be + i));
  }
  cta_sync();
}

__global__ void __launch_bounds__(NT, 2)
ray_long_kernel(const SceneDev sc, const RenderW w, const float* __restrict__ z_vals, const int64_t zs, const int S, const int white_bkgd,
                const int64_t R, const float* __restrict__ fagg, const float* __restrict__ partial,
                const float* __restrict__ rgbvis, const unsigned char* __restrict__ nvalid, float* __restrict__ rgb_out,
                float* __restrict__ depth_out, float* __restrict__ weights_out, unsigned char* __restrict__ mask_out,
                float* __restrict__ unc_out, float* __restrict__ feat_out, float* __restrict__ sigma_dbg, float* slabs,
                const FeatPeers peers,
                const size_t slab_floats) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;                        // weight staging ring
  float* sBl = sB + STAGE_FLOATS;          // [S][36] feature_agg half of the blend layer
  float* sLogit = sBl + RL_MAX_S * 36;     // [S*V]
  float* sRGB = sLogit + RL_MAX_S * 16;    // [S][4]
  float* sV = sRGB + RL_MAX_S * 4;         // sigma/alpha, T, weights, z: 4 x [S]
  float* red = sV + RL_MAX_S * 4;          // [64]
  float* sSig = sV, *sT = sV + RL_MAX_S, *sWt = sV + 2 * RL_MAX_S, *sZ = sV + 3 * RL_MAX_S;

  float* slab = slabs + (size_t)blockIdx.x * slab_floats;
  float* bX = slab;
  float* bC1 = bX + (S + 2) * RL_LDX;
  float* bC2 = bC1 + (S / 2 + 2) * RL_LDC1;
  float* bC3 = bC2 + (S / 4 + 2) * RL_LDC2;
  float* raw = bC3 + (S / 8 + 2) * RL_LDC3;   // [<= S][128]
  float* X = bX + RL_LDX;                      // logical row 0 (one zero halo row above and below every level)
  float* C1 = bC1 + RL_LDC1;
  float* C2 = bC2 + RL_LDC2;
  float* C3 = bC3 + RL_LDC3;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = sc.V;

  for (int64_t ray = blockIdx.x; ray < R; ray += gridDim.x) {
    const int64_t s0 = ray * S;
    cta_sync();  // previous ray fully consumed
    // ---- load feature_agg rows, clear the halo rows ---------------------------------------------------------------------
    for (int i = tid; i < RL_LDX; i += NT) { bX[i] = 0.f; bX[(S + 1) * RL_LDX + i] = 0.f; }
    for (int i = tid; i < RL_LDC1; i += NT) { bC1[i] = 0.f; bC1[(S / 2 + 1) * RL_LDC1 + i] = 0.f; }
    for (int i = tid; i < RL_LDC2; i += NT) { bC2[i] = 0.f; bC2[(S / 4 + 1) * RL_LDC2 + i] = 0.f; }
    for (int i = tid; i < RL_LDC3; i += NT) { bC3[i] = 0.f; bC3[(S / 8 + 1) * RL_LDC3 + i] = 0.f; }
    for (int i = tid; i < S * 32; i += NT) {
      const int s = i >> 5, c4 = i & 31;
      *reinterpret_cast<float4*>(X + s * RL_LDX + c4 * 4) = __ldg(reinterpret_cast<const float4*>(fagg + (s0 + s) * W_HID + c4 * 4));
    }
    for (int i = tid; i < S; i += NT) sZ[i] = z_vals[ray * zs + i];

    // ---- colour blend (model.py:528-538) ----------------------------------------------------------------------------------
    tile_gemm<4, 4, 32, false>(plainA(X, RL_LDX), S, w.bl1a, 32, 128, sB, [&](int r, int c, float v) { sBl[r * 36 + c] = v; });
    cta_sync();
    float* sW2 = sB;  // [16][32] | b2[16] | w3[16] | b3
    for (int i = tid; i < 512; i += NT) sW2[i] = __ldg(w.bl2 + i);
    if (tid < 16) { sW2[512 + tid] = __ldg(w.bl2_b + tid); sW2[528 + tid] = __ldg(w.bl3 + tid); }
    if (tid == 0) sW2[544] = __ldg(w.bl3_b);
    cta_sync();
    for (int i = tid; i < S * V; i += NT) {
      const int s = i / V;
      float h1[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(partial + partial_off(s0 * V + i, q)));
        const float4 b = *reinterpret_cast<const float4*>(sBl + s * 36 + q * 4);
        h1[q * 4 + 0] = leaky(a.x + b.x); h1[q * 4 + 1] = leaky(a.y + b.y);
        h1[q * 4 + 2] = leaky(a.z + b.z); h1[q * 4 + 3] = leaky(a.w + b.w);
      }
      float logit = sW2[544];
#pragma unroll 4
      for (int o = 0; o < 16; ++o) {
        float a = sW2[512 + o];
#pragma unroll
        for (int c = 0; c < 32; ++c) a = fmaf(sW2[o * 32 + c], h1[c], a);
        logit = fmaf(sW2[528 + o], leaky(a), logit);
      }
      const float vis = __ldg(rgbvis + (s0 * V + i) * 4 + 3);
      sLogit[i] = vis == 0.f ? -1e9f : logit;
    }
    cta_sync();
    for (int s = tid; s < S; s += NT) {
      float m = -FLT_MAX;
      for (int v = 0; v < V; ++v) m = fmaxf(m, sLogit[s * V + v]);
      float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
      for (int v = 0; v < V; ++v) {
        const float e = expf(sLogit[s * V + v] - m);
        const float4 c = __ldg(reinterpret_cast<const float4*>(rgbvis + ((s0 + s) * V + v) * 4));
        den += e; r += c.x * e; g += c.y * e; b += c.z * e;
      }
      sRGB[s * 4] = r / den; sRGB[s * 4 + 1] = g / den; sRGB[s * 4 + 2] = b / den;
    }

    // ---- RayUnet ---------------------------------------------------------------------------------------------------------------
    enc_long<64>(ASrc{X, RL_LDX, 128, -1, 0, 1}, S, w.u[0], 128, sB, red, raw, C1, RL_LDC1);          // conv1 -> c1 [S/2][64]
    enc_long<128>(ASrc{C1, RL_LDC1, 64, -1, 0, 1}, S / 2, w.u[1], 64, sB, red, raw, C2, RL_LDC2);      // conv2 -> c2 [S/4][128]
    enc_long<128>(ASrc{C2, RL_LDC2, 128, -1, 0, 1}, S / 4, w.u[2], 128, sB, red, raw, C3, RL_LDC3);    // conv3 -> c3 [S/8][128]
    dec_long<128>(C3, RL_LDC3, S / 8, w.u[3], 128, sB, red, raw, C2 + 128, RL_LDC2);                   // trans_conv3 -> x0
    dec_long<64>(C2, RL_LDC2, S / 4, w.u[4], 256, sB, red, raw, C1 + 64, RL_LDC1);                     // trans_conv2(c2|x0) -> x1
    dec_long<32>(C1, RL_LDC1, S / 2, w.u[5], 128, sB, red, raw, X + 128, RL_LDX);                      // trans_conv1(c1|x1) -> x2
    {
      // conv_out(x|x2) + LayerNorm + ELU, then sigma = softplus(w . y + b), one warp per sample
      tile_gemm<4, 8, 128, false>(ASrc{X, RL_LDX, 160, -1, 0, 1}, S, w.u[6].w, 128, 480, sB,
                                  [&](int r, int c, float v) { raw[r * 128 + c] = v + __ldg(w.u[6].b + c); });
      cta_sync();
      float mean, rstd;
      slab_stats(raw, S * 128, red, mean, rstd);
      const float4 sw = __ldg(reinterpret_cast<const float4*>(w.sig_w + lane * 4));
      for (int r = warp; r < S; r += NT / 32) {
        const int i = r * 128 + lane * 4;
        const float4 v = *reinterpret_cast<const float4*>(raw + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(w.u[6].g + i));
        const float4 be = __ldg(reinterpret_cast<const float4*>(w.u[6].be + i));
        float part = elu((v.x - mean) * rstd * g.x + be.x) * sw.x;
        part = fmaf(elu((v.y - mean) * rstd * g.y + be.y), sw.y, part);
        part = fmaf(elu((v.z - mean) * rstd * g.z + be.z), sw.z, part);
        part = fmaf(elu((v.w - mean) * rstd * g.w + be.w), sw.w, part);
        part = warp_sum(part);
        if (lane == 0) sSig[r] = softplus(part + __ldg(w.sig_b));
      }
    }
    cta_sync();

    // ---- compositing (model.py:541-575) ---------------------------------------------------------------------------------------
    for (int s = tid; s < S; s += NT) {
      if (sigma_dbg) sigma_dbg[s0 + s] = sSig[s];
      const float delta = s + 1 < S ? sZ[s + 1] - sZ[s] : 1e2f;
      sSig[s] = 1.f - expf(-delta * sSig[s]);  // alpha
    }
    cta_sync();
    if (tid == 0) {
      float T = 1.f;
      for (int s = 0; s < S; ++s) { sT[s] = T; T *= (1.f - sSig[s]); }
    }
    cta_sync();
    float wv = 0.f, zz = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, nv = 0.f;
    if (tid < S) {  // S <= NT
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

    // ---- rendered feature (model.py:594-598): feat = W2 (sum_s w_s h_s) + b2 sum_s w_s -----------------------------------------
    if (feat_out || peers.n > 0) {
      tile_gemm<4, 8, 128, false>(plainA(X, RL_LDX), S, w.ft1, 128, 128, sB,
                                  [&](int rr, int c, float v) { raw[rr * 128 + c] = leaky(v + __ldg(w.ft1_b + c)) * sWt[rr]; });
      cta_sync();
      float* sHs = sB;
      if (tid < 128) {
        float a = 0.f;
        for (int rr = 0; rr < S; ++rr) a += raw[rr * 128 + tid];
        sHs[tid] = a;
      }
      cta_sync();
      if (tid < C_FEAT) {
        float a = __ldg(w.ft2_b + tid) * wsum;
        for (int k = 0; k < 128; ++k) a = fmaf(__ldg(w.ft2 + k * C_FEAT + tid), sHs[k], a);
        if (feat_out) feat_out[ray * C_FEAT + tid] = a;
        for (int p = 0; p < peers.n; ++p) peers.p[p][(peers.row0 + ray) * C_FEAT + tid] = a;   // fused all-gather (peer stores)
      }
    }
  }
}

int launch_ray_long(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t zs, int64_t R, int S, int white_bkgd,
                    const float* fagg, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
                    float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
                    float* slabs, const FeatPeers& peers, cudaStream_t st) {
  if (R <= 0) return 0;
  if (S % 8 != 0 || S <= 128 || S > RL_MAX_S) return set_error("long-ray stage: samples per ray must be a multiple of 8 in (128, 256]");
  if (w.S != S) return set_error("ray stage: weights were packed for a different number of samples per ray");
  if (sc.V > 16) return set_error("ray stage: at most 16 reference views");
  if (!slabs) return set_error("long-ray stage: no scratch slabs");
  const size_t smem = (size_t)(STAGE_FLOATS + RL_MAX_S * (36 + 16 + 4 + 4) + 64) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ray_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  ray_long_kernel<<<(unsigned)ray_long_grid(R), NT, smem, st>>>(sc, w, z_vals, zs, S, white_bkgd, R, fagg, partial, rgbvis, nvalid,
                                                                rgb, depth, weights, mask, depth_unc, feat, sigma_dbg, slabs,
                                                                peers, ray_long_slab_floats(S));
  return check_launch("ray_long_kernel");
}

}  // namespace nlb
