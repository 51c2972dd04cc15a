// Hierarchical (importance) sampling of the render depths, N_importance > 0 (dead in the shipped configs):
//   model.py:486-496        64 coarse depths -> weights from the reference views -> inverse-CDF samples -> sort
//   multiview_aggregator.py:95-154   predict_weights_from_neuray (NeuRay projection, visibility features, mixture-of-
//                           logistics decoder, compute_prob(is_ref=True), visibility-weighted alpha over views, cumprod)
//   utils.py:73-112         sample_pdf: searchsorted(right=True), gather, linear interpolation
// One CTA per ray.  Thread = one (coarse depth, view) row for the decoder; the decoder weights (dec1 | dec2 | dec3, 34 KB)
// are resident in shared memory.  The uniform draws `u` come from the host (torch.rand in the reference).
#include <float.h>
#include "nlb_common.cuh"
#include "nlb_internal.h"
#include "render_kernels.h"

namespace nlb {

constexpr int HS_D = 64;       // coarse depths (model.py:489)
constexpr int HS_MAXV = 16;

__global__ void __launch_bounds__(NT, 2)
hier_sample_kernel(const SceneDev sc, const RenderW w, const float3 center, const float* __restrict__ dirs,
                   const float* __restrict__ z_coarse, const float* __restrict__ z_reg, const int S, const float* __restrict__ u,
                   const int NI, float* __restrict__ z_out, float* __restrict__ depth_coarse, int64_t* __restrict__ inds_out) {
  extern __shared__ __align__(16) float smem[];
  float* sW1 = smem;               // [32][128]
  float* sW2 = sW1 + 4096;         // 4 x [32][32]
  float* sW3 = sW2 + 4096;         // [6][32]
  float* sB1 = sW3 + 192;          // [128]
  float* sB2 = sB1 + 128;          // [128]
  float* sB3 = sB2 + 128;          // [8]
  float* sAl = sB3 + 8;            // alpha value [V][64]
  float* sVi = sAl + HS_MAXV * HS_D;   // visibility
  float* sMk = sVi + HS_MAXV * HS_D;   // mask
  float* sDn = sMk + HS_MAXV * HS_D;   // normalised inverse coarse depths [64] | dists [64]
  float* sWt = sDn + 2 * HS_D;     // weights [64]
  float* sCdf = sWt + HS_D;        // [64]
  float* sZ = sCdf + HS_D;         // merged depths before the sort [<= 256]
  const int tid = threadIdx.x;
  const int64_t ray = blockIdx.x;
  const int V = sc.V;
  for (int i = tid; i < 4096; i += NT) { sW1[i] = __ldg(w.dec1 + i); sW2[i] = __ldg(w.dec2 + i); }
  if (tid < 192) sW3[tid] = __ldg(w.dec3 + tid);
  if (tid < 128) { sB1[tid] = __ldg(w.dec1_b + tid); sB2[tid] = __ldg(w.dec2_b + tid); }
  if (tid < 6) sB3[tid] = __ldg(w.dec3_b + tid);
  const float near_ = sc.near_, far_ = sc.far_;
  const float ni = -1.f / near_, fi = -1.f / far_;
  if (tid < HS_D) sDn[tid] = (-1.f / z_coarse[tid] - ni) / (fi - ni);
  __syncthreads();
  if (tid < HS_D) sDn[HS_D + tid] = tid + 1 < HS_D ? sDn[tid + 1] - sDn[tid] : 1e6f;   // depth2dists (depth_fusion.py:47-50)
  __syncthreads();
  const float dx = dirs[ray * 3], dy = dirs[ray * 3 + 1], dz = dirs[ray * 3 + 2];

  for (int row = tid; row < HS_D * V; row += NT) {
    const int v = row / HS_D, d = row - v * HS_D;
    const float t = z_coarse[d];
    const float x = __fadd_rn(center.x, __fmul_rn(dx, t)), y = __fadd_rn(center.y, __fmul_rn(dy, t)),
                z = __fadd_rn(center.z, __fmul_rn(dz, t));
    // NeuRay-convention projection (depth_fusion.py:78-126), same as aggregate_kernel phase 1
    const float* kr = sc.cams + v * 32 + 12;
    const float c0 = fmaf(kr[2], z, fmaf(kr[1], y, kr[0] * x)) + kr[3];
    const float c1 = fmaf(kr[6], z, fmaf(kr[5], y, kr[4] * x)) + kr[7];
    float dep = fmaf(kr[10], z, fmaf(kr[9], y, kr[8] * x)) + kr[11];
    const bool bad = fabsf(dep) < 1e-4f;
    if (bad) dep = 1e-3f;
    const float qx = c0 / dep, qy = c1 / dep;
    const bool outside = qx < -0.5f || qx >= (float)sc.W - 0.5f || qy < -0.5f || qy >= (float)sc.H - 0.5f;
    const float valid = (!bad && !outside) ? 1.f : 0.f;
    const float xn = qx / (float)(sc.W - 1) * 2.f - 1.f, yn = qy / (float)(sc.H - 1) * 2.f - 1.f;
    float ix, iy;
    if (sc.vh == sc.H && sc.vw == sc.W) {
      ix = ((xn + 1.f) / 2.f) * (float)(sc.vw - 1); iy = ((yn + 1.f) / 2.f) * (float)(sc.vh - 1);
    } else {
      ix = ((xn + 1.f) * (float)sc.vw - 1.f) / 2.f; iy = ((yn + 1.f) * (float)sc.vh - 1.f) / 2.f;
    }
    // bilinear, border padding
    ix = fminf((float)(sc.vw - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(sc.vh - 1), fmaxf(iy, 0.f));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const int x1 = min(x0 + 1, sc.vw - 1), y1 = min(y0 + 1, sc.vh - 1);
    const float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy, wx = ix - fx, wy = iy - fy;
    const float w00 = ex * ey, w01 = (x0 + 1 <= sc.vw - 1) ? wx * ey : 0.f, w10 = (y0 + 1 <= sc.vh - 1) ? ex * wy : 0.f,
                w11 = (x0 + 1 <= sc.vw - 1 && y0 + 1 <= sc.vh - 1) ? wx * wy : 0.f;
    const float* base = sc.vis + ((size_t)v * sc.vh * sc.vw) * C_VIS;
    const float4* p00 = reinterpret_cast<const float4*>(base + ((size_t)y0 * sc.vw + x0) * C_VIS);
    const float4* p01 = reinterpret_cast<const float4*>(base + ((size_t)y0 * sc.vw + x1) * C_VIS);
    const float4* p10 = reinterpret_cast<const float4*>(base + ((size_t)y1 * sc.vw + x0) * C_VIS);
    const float4* p11 = reinterpret_cast<const float4*>(base + ((size_t)y1 * sc.vw + x1) * C_VIS);
    float xin[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = __ldg(p00 + q), b = __ldg(p01 + q), c = __ldg(p10 + q), e = __ldg(p11 + q);
      xin[q * 4 + 0] = (((a.x * w00 + b.x * w01) + c.x * w10) + e.x * w11) * valid;
      xin[q * 4 + 1] = (((a.y * w00 + b.y * w01) + c.y * w10) + e.y * w11) * valid;
      xin[q * 4 + 2] = (((a.z * w00 + b.z * w01) + c.z * w10) + e.z * w11) * valid;
      xin[q * 4 + 3] = (((a.w * w00 + b.w * w01) + c.w * w10) + e.w * w11) * valid;
    }
    // mixture-of-logistics decoder, head by head (visibility_decoder.py:62-107)
    float o[6];
#pragma unroll 1
    for (int hd = 0; hd < 4; ++hd) {
      float h1[32], h2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) h1[j] = sB1[hd * 32 + j];
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const float xv = xin[k];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sW1 + k * 128 + hd * 32 + j);
          h1[j] = fmaf(xv, w4.x, h1[j]); h1[j + 1] = fmaf(xv, w4.y, h1[j + 1]);
          h1[j + 2] = fmaf(xv, w4.z, h1[j + 2]); h1[j + 3] = fmaf(xv, w4.w, h1[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) { h1[j] = elu(h1[j]); h2[j] = sB2[hd * 32 + j]; }
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const float xv = h1[k];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sW2 + hd * 1024 + k * 32 + j);
          h2[j] = fmaf(xv, w4.x, h2[j]); h2[j + 1] = fmaf(xv, w4.y, h2[j + 1]);
          h2[j + 2] = fmaf(xv, w4.z, h2[j + 2]); h2[j + 3] = fmaf(xv, w4.w, h2[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) h2[j] = elu(h2[j]);
      const int j0 = hd < 2 ? hd * 2 : hd + 2, nj = hd < 2 ? 2 : 1;
      for (int jj = 0; jj < nj; ++jj) {
        float a = sB3[j0 + jj];
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(sW3[(j0 + jj) * 32 + k], h2[k], a);
        o[j0 + jj] = a;
      }
    }
    const float m0 = softplus(o[0]), m1 = softplus(o[1]);
    const float v0 = softplus(o[2]) + 0.05f, v1 = softplus(o[3]) + 0.05f;
    const float aw = sigmoidf(o[4]), vs = sigmoidf(o[5]);
    // compute_prob(is_ref=True) (visibility_decoder.py:6-51,150-181)
    const float dn = (-1.f / fmaxf(dep, 1e-5f) - ni) / (fi - ni);
    const float half_lo = sDn[HS_D + (d > 0 ? d - 1 : 0)] * 0.5f, half_hi = sDn[HS_D + d] * 0.5f;
    const float nearp = dn - half_lo, farp = dn + half_hi;
    const float cdf00 = (0.5f + 0.5f * tanhf((nearp - m0) * v0)) * vs, cdf01 = (0.5f + 0.5f * tanhf((nearp - m1) * v1)) * vs;
    const float cdf10 = (0.5f + 0.5f * tanhf((farp - m0) * v0)) * vs, cdf11 = (0.5f + 0.5f * tanhf((farp - m1) * v1)) * vs;
    const float visib = (1.f - cdf00) * aw + (1.f - cdf01) * (1.f - aw);
    const float hit = (cdf10 - cdf00) * aw + (cdf11 - cdf01) * (1.f - aw);
    sAl[row] = logf(hit / (visib - hit + 1e-5f) + 1e-5f);
    sVi[row] = visib;
    sMk[row] = valid;
  }
  __syncthreads();
  // visibility-weighted alpha over the views (multiview_aggregator.py:138-148)
  if (tid < HS_D) {
    float num = 0.f, den = 0.f;
    int nvalid = 0;
    for (int v = 0; v < V; ++v) {
      const float m = sMk[v * HS_D + tid];
      const float a = sAl[v * HS_D + tid] * m + (1.f - m) * -15.f;
      const float vi = sVi[v * HS_D + tid] * m;
      num += a * vi; den += vi;
      nvalid += m != 0.f;
    }
    float a = num / fmaxf(den, 1e-8f);
    const float inv = nvalid == 0 ? 1.f : 0.f;
    a = a * (1.f - inv) + inv * -15.f;
    sWt[tid] = sigmoidf(a);   // alpha
  }
  __syncthreads();
  if (tid == 0) {
    float T = 1.f, dc = 0.f;
    for (int d = 0; d < HS_D; ++d) {
      const float a = sWt[d];
      const float wgt = a * T;
      T *= (1.f - a);
      sWt[d] = wgt;
      dc += wgt * z_coarse[d];
    }
    depth_coarse[ray] = dc;
    // sample_pdf on weights[1:-1] (62 bins), bins = interval mid points (63)
    float tot = 0.f;
    for (int d = 1; d < HS_D - 1; ++d) tot += sWt[d] + 1e-5f;
    float c = 0.f;
    sCdf[0] = 0.f;
    for (int d = 1; d < HS_D - 1; ++d) { c += (sWt[d] + 1e-5f) / tot; sCdf[d] = c; }   // cdf[0..62]
  }
  __syncthreads();
  constexpr int NB = HS_D - 2;  // 62 pdf bins, cdf has NB + 1 entries, bins (mid points) NB + 1
  for (int j = tid; j < NI; j += NT) {
    const float uu = u[ray * NI + j];
    int lo = 0, hi = NB + 1;   // first index with cdf[idx] > uu  (searchsorted right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sCdf[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, NB);
    const float c0 = sCdf[below], c1 = sCdf[above];
    const float b0 = 0.5f * (z_coarse[below] + z_coarse[below + 1]), b1 = 0.5f * (z_coarse[above] + z_coarse[above + 1]);
    float denom = c1 - c0;
    if (denom < 1e-5f) denom = 1.f;
    sZ[S + j] = b0 + (uu - c0) / denom * (b1 - b0);
    if (inds_out) inds_out[ray * NI + j] = lo;
  }
  for (int j = tid; j < S; j += NT) sZ[j] = z_reg[j];
  __syncthreads();
  // rank sort of the S + NI depths (torch.sort, model.py:495; equal values keep their input order)
  const int n = S + NI;
  for (int i = tid; i < n; i += NT) {
    const float zi = sZ[i];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const float zk = sZ[k];
      rank += (zk < zi) || (zk == zi && k < i);
    }
    z_out[ray * n + rank] = zi;
  }
}

int launch_hier_sample(const SceneDev& sc, const RenderW& w, const float* center_host, const float* dirs, int64_t R,
                       const float* z_coarse, const float* z_reg, int S, const float* u, int NI, float* z_out,
                       float* depth_coarse, int64_t* inds, cudaStream_t st) {
  if (R <= 0) return 0;
  if (sc.V < 1 || sc.V > HS_MAXV) return set_error("hierarchical sampling: number of reference views must be in 1..16");
  if (NI < 1 || S < 1 || S + NI > 256) return set_error("hierarchical sampling: N_samples + N_importance must be <= 256");
  const size_t smem = (size_t)(4096 * 2 + 192 + 128 * 2 + 8 + 3 * HS_MAXV * HS_D + 2 * HS_D + 2 * HS_D + 256) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(hier_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const float3 c = make_float3(center_host[0], center_host[1], center_host[2]);
  hier_sample_kernel<<<(unsigned)R, NT, smem, st>>>(sc, w, c, dirs, z_coarse, z_reg, S, u, NI, z_out, depth_coarse, inds);
  return check_launch("hier_sample_kernel");
}

}  // namespace nlb
