// Per-ray stage of ConditionalNeRF.render_rays (conditional_nerf/model.py:521-598), one CTA per ray (sm_100a):
//   colour blend over views (model.py:528-538) -> RayUnet along the ray (conditional_nerf/ray_unet.py:5-69) ->
//   softplus density (model.py:525) -> alpha compositing, depth, depth variance, validity mask (model.py:541-575) ->
//   rendered 192-d feature (model.py:594-598).
//
// The [S][C] activations of a ray live in shared memory, row-major, with one zero row above and below so that a
// k=3 Conv1d is a 3-tap tile GEMM (K = 3*Cin) and a stride-2 ConvTranspose1d is an "even" GEMM (tap 1) plus an "odd"
// GEMM (tap 2 on row j, tap 0 on row j+1).  The joint LayerNorm([C, S_level]) of every block is a CTA-wide
// two-pass reduction; for the pooled encoder blocks and conv_out it is applied to the register accumulators, so the
// pre-pool tensors are never written anywhere.  Skip connections are column ranges of the same buffers.
//
// Exact rewrite: feat = sum_s w_s (W2 h_s + b2) = W2 (sum_s w_s h_s) + b2 sum_s w_s, so the second feat_mlp layer
// runs once per ray instead of once per sample.
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"

namespace nlb {

constexpr int LDXR = 164;  // [x 128 | x2 32] + 4
constexpr int LDC1 = 132;  // [c1 64 | x1 64] + 4
constexpr int LDC2 = 260;  // [c2 128 | x0 128] + 4
constexpr int LDC3 = 132;  // [c3 128] + 4

static size_t ray_smem_floats(int S) {
  return (size_t)STAGE_FLOATS + (size_t)(S + 2) * LDXR + (size_t)(S / 2 + 2) * LDC1 + (size_t)(S / 4 + 2) * LDC2 +
         (size_t)(S / 8 + 2) * LDC3 + (size_t)S * 4 + (size_t)S * 4 + 64;
}

// LayerNorm over a [rows x cols] slab held in shared memory (two-pass), then ELU, in place.
__device__ __forceinline__ void ln_elu_smem(float* base, int ld, int rows, int cols, const float* __restrict__ g,
                                            const float* __restrict__ be, float* red) {
  const int n = rows * cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) s += base[(i / cols) * ld + (i % cols)];
  const float mean = block_sum(s, red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) {
    const float d = base[(i / cols) * ld + (i % cols)] - mean;
    q += d * d;
  }
  const float rstd = 1.f / sqrtf(block_sum(q, red) / (float)n + 1e-5f);
  for (int i = threadIdx.x; i < n; i += NT) {
    const int r = i / cols, c = i % cols;
    float* p = base + r * ld + c;
    *p = elu((*p - mean) * rstd * __ldg(g + i) + __ldg(be + i));
  }
  cta_sync();
}

// LayerNorm statistics of a register fragment (bias already added).  Returns mean / rstd to every thread.
template <int TM, int TN, int COLS>
__device__ __forceinline__ void frag_stats(const Frag<TM, TN, COLS>& f, int rows, float* red, float& mean, float& rstd) {
  const float n = (float)(rows * COLS);
  float s = 0.f;
  if (f.active)
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) s += f.acc[i][j];
  mean = block_sum(s, red) / n;
  float q = 0.f;
  if (f.active)
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const float d = f.acc[i][j] - mean;
        q += d * d;
      }
  rstd = 1.f / sqrtf(block_sum(q, red) / n + 1e-5f);
}

// Encoder block: conv (3 taps) + LayerNorm + ELU + MaxPool1d(2); result written to dst rows [0, rows/2).
template <int TM, int COLS>
__device__ __forceinline__ void enc_block(const ASrc A, int rows, const UnetLayer& L, int cin, float* sB, float* red,
                                          float* dst, int ldd) {
  Frag<TM, 8, COLS> f;
  tile_gemm_frag<TM, 8, COLS>(A, rows, L.w, COLS, 3 * cin, sB, f);
  if (f.active)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = __ldg(L.b + f.col(j));
#pragma unroll
      for (int i = 0; i < TM; ++i) f.acc[i][j] += b;
    }
  float mean, rstd;
  frag_stats(f, rows, red, mean, rstd);
  if (f.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = f.col(j);
#pragma unroll
      for (int i = 0; i < TM; i += 2) {
        const int r = f.r0 + i;
        const float y0 = elu((f.acc[i][j] - mean) * rstd * __ldg(L.g + r * COLS + c) + __ldg(L.be + r * COLS + c));
        const float y1 = elu((f.acc[i + 1][j] - mean) * rstd * __ldg(L.g + (r + 1) * COLS + c) + __ldg(L.be + (r + 1) * COLS + c));
        dst[(r >> 1) * ldd + c] = fmaxf(y0, y1);
      }
    }
  }
  cta_sync();
}

// Decoder block: stride-2 transposed conv written to dst rows [0, 2*rows), then LayerNorm + ELU in place.
template <int COLS>
__device__ __forceinline__ void dec_block(const float* in, int ldi, int rows, const UnetLayer& L, int cin, float* sB,
                                          float* red, float* dst, int ldd) {
  tile_gemm<1, 8, COLS, false>(ASrc{in, ldi, cin, 0, 0, 0}, rows, L.w, COLS, cin, sB,
                               [&](int r, int c, float v) { dst[(2 * r) * ldd + c] = v + __ldg(L.b + c); });
  tile_gemm<1, 8, COLS, false>(ASrc{in, ldi, cin, 0, 1, 0}, rows, L.w + (size_t)cin * COLS, COLS, 2 * cin, sB,
                               [&](int r, int c, float v) { dst[(2 * r + 1) * ldd + c] = v + __ldg(L.b + c); });
  cta_sync();
  ln_elu_smem(dst, ldd, 2 * rows, COLS, L.g, L.be, red);
}

__global__ void __launch_bounds__(NT, 1)
ray_kernel(const SceneDev sc, const RenderW w, const float* __restrict__ z_vals, const int S, const int white_bkgd,
           const float* __restrict__ fagg, const float* __restrict__ partial, const float* __restrict__ rgbvis,
           const unsigned char* __restrict__ nvalid, float* __restrict__ rgb_out, float* __restrict__ depth_out,
           float* __restrict__ weights_out, unsigned char* __restrict__ mask_out, float* __restrict__ unc_out,
           float* __restrict__ feat_out, float* __restrict__ sigma_dbg) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;
  float* bX = sB + STAGE_FLOATS;
  float* bC1 = bX + (S + 2) * LDXR;
  float* bC2 = bC1 + (S / 2 + 2) * LDC1;
  float* bC3 = bC2 + (S / 4 + 2) * LDC2;
  float* sRGB = bC3 + (S / 8 + 2) * LDC3;   // [S][4]
  float* sV = sRGB + S * 4;                 // sigma/alpha [S], T [S], weights [S], z [S]
  float* red = sV + S * 4;                  // [64] scratch
  float* sSig = sV, *sT = sV + S, *sWt = sV + 2 * S, *sZ = sV + 3 * S;
  float* X = bX + LDXR;      // logical row 0
  float* C1 = bC1 + LDC1;
  float* C2 = bC2 + LDC2;
  float* C3 = bC3 + LDC3;

  const int tid = threadIdx.x;
  const int64_t ray = blockIdx.x;
  const int64_t s0 = ray * S;
  const int V = sc.V;

  // ---- load feature_agg rows, clear the halo rows -------------------------------------------------------------------
  for (int i = tid; i < LDXR; i += NT) { bX[i] = 0.f; bX[(S + 1) * LDXR + i] = 0.f; }
  for (int i = tid; i < LDC1; i += NT) { bC1[i] = 0.f; bC1[(S / 2 + 1) * LDC1 + i] = 0.f; }
  for (int i = tid; i < LDC2; i += NT) { bC2[i] = 0.f; bC2[(S / 4 + 1) * LDC2 + i] = 0.f; }
  for (int i = tid; i < LDC3; i += NT) { bC3[i] = 0.f; bC3[(S / 8 + 1) * LDC3 + i] = 0.f; }
  for (int i = tid; i < S * 32; i += NT) {
    const int s = i >> 5, c4 = i & 31;
    *reinterpret_cast<float4*>(X + s * LDXR + c4 * 4) = __ldg(reinterpret_cast<const float4*>(fagg + (s0 + s) * W_HID + c4 * 4));
  }
  if (tid < S) sZ[tid] = z_vals[tid];

  // ---- colour blend (model.py:528-538) ------------------------------------------------------------------------------
  float* sBl = bC1;      // [S][36]: feature_agg half of layer 1
  float* sLogit = bC2;   // [S][V]
  tile_gemm<4, 4, 32, false>(plainA(X, LDXR), S, w.bl1a, 32, 128, sB, [&](int r, int c, float v) { sBl[r * 36 + c] = v; });
  cta_sync();
  float* sW2 = sB;  // [16][32] | b2[16] | w3[16] | b3
  for (int i = tid; i < 512; i += NT) sW2[i] = __ldg(w.bl2 + i);
  if (tid < 16) { sW2[512 + tid] = __ldg(w.bl2_b + tid); sW2[528 + tid] = __ldg(w.bl3 + tid); }
  if (tid == 0) sW2[544] = __ldg(w.bl3_b);
  cta_sync();
  for (int i = tid; i < S * V; i += NT) {
    const int s = i / V;
    const float4* pp = reinterpret_cast<const float4*>(partial + (s0 * V + i) * 32);
    float h1[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = __ldg(pp + q);
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
  if (tid < S) {
    float m = -FLT_MAX;
    for (int v = 0; v < V; ++v) m = fmaxf(m, sLogit[tid * V + v]);
    float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
    for (int v = 0; v < V; ++v) {
      const float e = expf(sLogit[tid * V + v] - m);
      const float4 c = __ldg(reinterpret_cast<const float4*>(rgbvis + ((s0 + tid) * V + v) * 4));
      den += e; r += c.x * e; g += c.y * e; b += c.z * e;
    }
    sRGB[tid * 4] = r / den; sRGB[tid * 4 + 1] = g / den; sRGB[tid * 4 + 2] = b / den;
  }
  cta_sync();
  // the blend scratch aliased the halo rows of bC1 / bC2: clear them again
  for (int i = tid; i < LDC1; i += NT) bC1[i] = 0.f;
  for (int i = tid; i < LDC2; i += NT) bC2[i] = 0.f;
  cta_sync();

  // ---- RayUnet ---------------------------------------------------------------------------------------------------------
  enc_block<4, 64>(ASrc{X, LDXR, 128, -1, 0, 1}, S, w.u[0], 128, sB, red, C1, LDC1);            // conv1 -> c1 [S/2][64]
  enc_block<4, 128>(ASrc{C1, LDC1, 64, -1, 0, 1}, S / 2, w.u[1], 64, sB, red, C2, LDC2);         // conv2 -> c2 [S/4][128]
  enc_block<2, 128>(ASrc{C2, LDC2, 128, -1, 0, 1}, S / 4, w.u[2], 128, sB, red, C3, LDC3);       // conv3 -> c3 [S/8][128]
  dec_block<128>(C3, LDC3, S / 8, w.u[3], 128, sB, red, C2 + 128, LDC2);                         // trans_conv3 -> x0
  dec_block<64>(C2, LDC2, S / 4, w.u[4], 256, sB, red, C1 + 64, LDC1);                           // trans_conv2(c2|x0) -> x1
  dec_block<32>(C1, LDC1, S / 2, w.u[5], 128, sB, red, X + 128, LDXR);                           // trans_conv1(c1|x1) -> x2
  {
    // conv_out(x|x2) + LayerNorm + ELU, then sigma = softplus(w . y + b) without materialising y
    Frag<8, 8, 128> f;
    tile_gemm_frag<8, 8, 128>(ASrc{X, LDXR, 160, -1, 0, 1}, S, w.u[6].w, 128, 480, sB, f);
    if (f.active)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = __ldg(w.u[6].b + f.col(j));
#pragma unroll
        for (int i = 0; i < 8; ++i) f.acc[i][j] += b;
      }
    float mean, rstd;
    frag_stats(f, S, red, mean, rstd);
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[i] = 0.f;
    if (f.active)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = f.col(j);
        const float sw = __ldg(w.sig_w + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = f.r0 + i;
          const float y = elu((f.acc[i][j] - mean) * rstd * __ldg(w.u[6].g + r * 128 + c) + __ldg(w.u[6].be + r * 128 + c));
          part[i] = fmaf(y, sw, part[i]);
        }
      }
    // the 16 threads that share a row group are 16 consecutive lanes
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = part[i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (f.active && (tid & 15) == 0) sSig[f.r0 + i] = softplus(v + __ldg(w.sig_b));
    }
  }
  cta_sync();

  // ---- compositing (model.py:541-575) -------------------------------------------------------------------------------------
  if (tid < S) {
    if (sigma_dbg) sigma_dbg[s0 + tid] = sSig[tid];
    const float delta = tid + 1 < S ? sZ[tid + 1] - sZ[tid] : 1e2f;
    sSig[tid] = 1.f - expf(-delta * sSig[tid]);  // alpha
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

  // ---- rendered feature (model.py:594-598) ---------------------------------------------------------------------------------
  if (feat_out) {
    Frag<8, 8, 128> f;
    tile_gemm_frag<8, 8, 128>(plainA(X, LDXR), S, w.ft1, 128, 128, sB, f);
    cta_sync();  // staging ring is free: reuse it for the per-row-group partial sums
    float* sPart = sB;  // [16][128]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = f.col(j);
      float a = 0.f;
      if (f.active) {
        const float bb = __ldg(w.ft1_b + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(leaky(f.acc[i][j] + bb), sWt[f.r0 + i], a);
      }
      sPart[(tid / 16) * 128 + c] = a;
    }
    cta_sync();
    float* sHs = sB + 16 * 128;
    if (tid < 128) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) a += sPart[t * 128 + tid];
      sHs[tid] = a;
    }
    cta_sync();
    if (tid < C_FEAT) {
      float a = __ldg(w.ft2_b + tid) * wsum;
      for (int k = 0; k < 128; ++k) a = fmaf(__ldg(w.ft2 + k * C_FEAT + tid), sHs[k], a);
      feat_out[ray * C_FEAT + tid] = a;
    }
  }
}

int launch_ray(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t R, int S, int white_bkgd,
               const float* fagg, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
               float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
               cudaStream_t st) {
  if (R <= 0) return 0;
  if (S % 8 != 0 || S < 8 || S > 128) return set_error("ray stage: samples per ray must be a multiple of 8 in [8, 128]");
  if (w.S != S) return set_error("ray stage: weights were packed for a different number of samples per ray");
  const size_t smem = ray_smem_floats(S) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  ray_kernel<<<(unsigned)R, NT, smem, st>>>(sc, w, z_vals, S, white_bkgd, fagg, partial, rgbvis, nvalid, rgb, depth,
                                          weights, mask, depth_unc, feat, sigma_dbg);
  return check_launch("ray_kernel");
}

}  // namespace nlb
