// 3D<->2D coarse-to-fine matcher kernels (sm_100a).
//
//   s2d_kernel          - S2DMatching scores (matching/sparse_to_dense.py:125-127).  The reference materialises
//                         x[n,m,:] = a_n * b_m ([N3,Mc,192], 15 GB at 4096 x 4800) and runs a 192->128->128->1 MLP on
//                         it.  Here a CTA keeps a tile of 128 cell descriptors b_m in shared memory and, for each 3D
//                         descriptor a_n of its slice, folds a_n into the first layer while the weight tile is staged
//                         ((W1 diag(a_n)) b_m), runs both layers from shared memory and writes only score[n, m-tile].
//   colmax/rowmatch/compact - the mutual-nearest rule of sparse_to_dense.py:136-142.
//   fine_windows_kernel - gathers the 7x7 fine-map windows of the matched cells only (instead of F.unfold over the
//                         whole map, fine_matching.py:53-57) and applies fine_preprocess.proj.
//   fine_match_kernel   - FineMatching.forward: pair MLP, softmax heat-map, DSNT expectation and std
//                         (fine_matching.py:122-149).
#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include "match_kernels.h"
#include "nlb_common.cuh"

namespace nlb {

// ---- packed matcher weights ---------------------------------------------------------------------------------------
static MatchW match_layout(const float* base, int C, size_t* total) {
  size_t off = 0;
  auto take = [&](size_t n) { const float* p = base + off; off += (n + 63) / 64 * 64; return p; };
  MatchW w{};
  w.C = C;
  for (int i = 0; i < 2; ++i) {
    PairMlp& m = i == 0 ? w.coarse : w.fine;
    m.w1t = take(192 * 128); m.b1 = take(128);
    m.w2t = take(128 * 128); m.b2 = take(128);
    m.w3 = take(128); m.b3 = take(1);
  }
  const int Cp = (C + 31) / 32 * 32;
  w.projt = take((size_t)Cp * 192); w.proj_b = take(192);
  w.tb_w2c = take(128 * 128);
  if (total) *total = off;
  return w;
}
size_t match_weights_floats(int C) { size_t t; match_layout(nullptr, C, &t); return t; }
MatchW match_weights_view(const float* packed, int C) { return match_layout(packed, C, nullptr); }

int match_weights_pack(const float* const* p, int n_params, int C, float* packed, size_t packed_floats, cudaStream_t st) {
  if (n_params != 14) return set_error("match_weights_pack: expected 14 parameter tensors");
  if (C < 1 || C > 320) return set_error("match_weights_pack: fine feature channels must be in 1..320");
  if (packed_floats < match_weights_floats(C)) return set_error("match_weights_pack: packed buffer too small");
  for (int i = 0; i < 14; ++i) if (!p[i]) return set_error("match_weights_pack: null parameter pointer");
  cudaMemsetAsync(packed, 0, match_weights_floats(C) * sizeof(float), st);
  const MatchW w = match_weights_view(packed, C);
  auto T = [&](const float* dst, const float* src, int Kp, int N, int src_ld, int Kv) {
    const int n = Kp * N;
    pack_t_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(dst), src, Kp, N, N, src_ld, 0, Kv);
  };
  auto Cpy = [&](const float* dst, const float* src, int n) {
    pack_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(dst), src, n);
  };
  for (int i = 0; i < 2; ++i) {
    const PairMlp& m = i == 0 ? w.coarse : w.fine;
    const float* const* q = p + 6 * i;
    T(m.w1t, q[0], 192, 128, 192, 192); Cpy(m.b1, q[1], 128);
    T(m.w2t, q[2], 128, 128, 128, 128); Cpy(m.b2, q[3], 128);
    Cpy(m.w3, q[4], 128); Cpy(m.b3, q[5], 1);
  }
  const int Cp = (C + 31) / 32 * 32;
  T(w.projt, p[12], Cp, 192, C, C); Cpy(w.proj_b, p[13], 192);
  pack_tcb16_kernel<<<(128 * 128 + 255) / 256, 256, 0, st>>>(reinterpret_cast<uint16_t*>(const_cast<float*>(w.tb_w2c)), p[2], 128, 128, 128, 0, 1,
                                                           32, 128, 0, 0, 128);
  return check_launch("match_weights_pack");
}

// ---- S2D scores -----------------------------------------------------------------------------------------------------
// A CTA owns S2D_ROWS = 64 cells of the 2D map and walks a slice of the 3D points; two CTAs share an SM (109 KB of shared
// memory and <= 128 registers each), so the barrier / staging stalls of one overlap the FFMA2 stream of the other (with one
// 128-cell CTA per SM the kernel sat at 50 % of the fp32 pipe).
constexpr int LDD = 196;  // 192 + 4
constexpr int LDH2 = 132;
constexpr int S2D_ROWS = 64;
constexpr int S2D_NS = 3;  // weight staging slots (8 KB each)
constexpr int S2D_SMEM_FLOATS = S2D_NS * KT * 128 + S2D_ROWS * LDD + S2D_ROWS * LDH2 + 192 + 64;

// second layer + 128->1 head on a hidden tile held in shared memory; returns the logit of row r0+i in lane tc==0
template <int TM, int NS, class Out>
__device__ __forceinline__ void pair_tail(const PairMlp& m, const float* sH, const int rows, float* sB, Out out) {
  Frag<TM, 8, 128> f;
  tile_gemm_frag<TM, 8, 128, false, NS>(plainA(sH, LDH2), rows, m.w2t, 128, 128, sB, f);
  float part[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) part[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = f.col(j);
    const float b = __ldg(m.b2 + c), w3 = __ldg(m.w3 + c);
#pragma unroll
    for (int i = 0; i < TM; ++i) part[i] = fmaf(fmaxf(f.acc[i][j] + b, 0.f), w3, part[i]);
  }
  const float b3 = __ldg(m.b3);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    float v = part[i];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if ((threadIdx.x & 15) == 0) out(f.r0 + i, v + b3);
  }
}

__global__ void __launch_bounds__(NT, 2)
s2d_kernel(const PairMlp m, const float* __restrict__ desc0, const float* __restrict__ desc1, const int64_t N,
           const int64_t M, const int n_per_cta, float* __restrict__ score) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TM = S2D_ROWS / 16;
  float* sB = smem;
  float* sD = sB + S2D_NS * KT * 128;   // [S2D_ROWS][LDD] cell descriptors
  float* sH = sD + S2D_ROWS * LDD;      // [S2D_ROWS][LDH2]
  float* sAn = sH + S2D_ROWS * LDH2;    // [192]
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * S2D_ROWS;
  const int mr = (int)min((int64_t)S2D_ROWS, M - m0);
  for (int i = tid; i < S2D_ROWS * 48; i += NT) {
    const int r = i / 48, c4 = i % 48;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < mr) v = __ldg(reinterpret_cast<const float4*>(desc1 + (m0 + r) * 192 + c4 * 4));
    *reinterpret_cast<float4*>(sD + r * LDD + c4 * 4) = v;
  }
  const int64_t nb = (int64_t)blockIdx.y * n_per_cta;
  const int64_t ne = min(N, nb + n_per_cta);
  for (int64_t n = nb; n < ne; ++n) {
    __syncthreads();  // previous iteration is done with sAn / sH
    if (tid < 192) sAn[tid] = __ldg(desc0 + n * 192 + tid);
    {
      Frag<TM, 8, 128> f;
      tile_gemm_frag<TM, 8, 128, true, S2D_NS>(plainA(sD, LDD), S2D_ROWS, m.w1t, 128, 192, sB, f, sAn);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = f.col(j);
        const float b = __ldg(m.b1 + c);
#pragma unroll
        for (int i = 0; i < TM; ++i) sH[(f.r0 + i) * LDH2 + c] = fmaxf(f.acc[i][j] + b, 0.f);
      }
    }
    pair_tail<TM, S2D_NS>(m, sH, S2D_ROWS, sB, [&](int r, float logit) {
      if (r < mr) score[n * M + m0 + r] = 1.f / (1.f + expf(-logit));
    });
  }
}

int launch_s2d(const MatchW& w, const float* desc0, const float* desc1, int64_t N, int64_t M, float* score,
               cudaStream_t st) {
  static const bool v1 = getenv("NLB_S2D_V1") != nullptr;   // A/B switch: the fp32 FFMA2 kernel below
  if (!v1) return launch_s2d_tc(w, desc0, desc1, N, M, score, st);
  const size_t smem = S2D_SMEM_FLOATS * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const unsigned gx = (unsigned)((M + S2D_ROWS - 1) / S2D_ROWS);
  // aim for ~4 waves of 296 resident CTAs
  int64_t gy = (1184 + gx - 1) / gx;
  if (gy > N) gy = N;
  if (gy < 1) gy = 1;
  const int n_per = (int)((N + gy - 1) / gy);
  gy = (N + n_per - 1) / n_per;
  s2d_kernel<<<dim3(gx, (unsigned)gy), NT, smem, st>>>(w.coarse, desc0, desc1, N, M, n_per, score);
  return check_launch("s2d_kernel");
}

// ---- mutual nearest -------------------------------------------------------------------------------------------------
__global__ void colmax_kernel(const float* __restrict__ score, int64_t N, int64_t M, float* __restrict__ colmax) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  float v = -FLT_MAX;
  for (int64_t n = 0; n < N; ++n) v = fmaxf(v, score[n * M + m]);
  colmax[m] = v;
}

// one CTA per row: row max, then the first column that passes all three tests (or -1)
__global__ void rowmatch_kernel(const float* __restrict__ score, int64_t N, int64_t M, float thr,
                                const float* __restrict__ colmax, long long* __restrict__ rowj) {
  __shared__ float sred[32];
  __shared__ long long sidx[32];
  const int64_t n = blockIdx.x;
  const float* row = score + n * M;
  float v = -FLT_MAX;
  for (int64_t m = threadIdx.x; m < M; m += blockDim.x) v = fmaxf(v, row[m]);
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float rmax = -FLT_MAX;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) rmax = fmaxf(rmax, sred[i]);
  long long best = LLONG_MAX;
  for (int64_t m = threadIdx.x; m < M; m += blockDim.x) {
    const float s = row[m];
    if (s > thr && s == rmax && s == colmax[m]) { best = m; break; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) sidx[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long b = LLONG_MAX;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) b = sidx[i] < b ? sidx[i] : b;
    rowj[n] = b == LLONG_MAX ? -1 : b;
  }
}

// single CTA, ordered compaction
__global__ void compact_kernel(const long long* __restrict__ rowj, int64_t N, long long* __restrict__ i_ids,
                               long long* __restrict__ j_ids, int* __restrict__ count) {
  __shared__ int swarp[32];
  __shared__ int sbase;
  if (threadIdx.x == 0) sbase = 0;
  __syncthreads();
  for (int64_t b = 0; b < N; b += blockDim.x) {
    const int64_t n = b + threadIdx.x;
    const long long j = n < N ? rowj[n] : -1;
    const int flag = j >= 0;
    const unsigned ball = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) swarp[wid] = __popc(ball);
    __syncthreads();
    int pre = 0;
    for (int i = 0; i < wid; ++i) pre += swarp[i];
    int total = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += swarp[i];
    const int pos = sbase + pre + __popc(ball & ((1u << lane) - 1));
    if (flag) { i_ids[pos] = n; j_ids[pos] = j; }
    __syncthreads();
    if (threadIdx.x == 0) sbase += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = sbase;
}

int launch_mutual(const float* score, int64_t N, int64_t M, float thr, int64_t* i_ids, int64_t* j_ids, int* count,
                  void* scratch, cudaStream_t st) {
  long long* rowj = (long long*)scratch;
  float* colmax = (float*)((char*)scratch + ((size_t)N * 8 + 255) / 256 * 256);
  colmax_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(score, N, M, colmax);
  rowmatch_kernel<<<(unsigned)N, 256, 0, st>>>(score, N, M, thr, colmax, rowj);
  compact_kernel<<<1, 1024, 0, st>>>(rowj, N, (long long*)i_ids, (long long*)j_ids, count);
  return check_launch("mutual_matches");
}

// ---- fine windows ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
fine_windows_kernel(const MatchW w, const float* __restrict__ feat, const int h, const int wd, const int C, const int stride,
                    const int coarse_w, const long long* __restrict__ j_ids, const int64_t Mm, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;
  float* sA = sB + STAGE_FLOATS;
  const int Cp = (C + 31) / 32 * 32;
  const int ld = Cp + 4;
  const int64_t rows_total = Mm * 49;
  const int64_t r0 = (int64_t)blockIdx.x * 128;
  const int nr = (int)min((int64_t)128, rows_total - r0);
  for (int i = threadIdx.x; i < 128 * Cp; i += NT) {
    const int r = i / Cp, c = i - r * Cp;
    float v = 0.f;
    if (r < nr && c < C) {
      const int64_t g = r0 + r;
      const int64_t t = g / 49;
      const int ww = (int)(g - t * 49);
      const long long j = j_ids[t];
      const int y = (int)(j / coarse_w) * stride - 3 + ww / 7;
      const int x = (int)(j % coarse_w) * stride - 3 + ww % 7;
      if (y >= 0 && y < h && x >= 0 && x < wd) v = __ldg(feat + ((size_t)y * wd + x) * C + c);
    }
    sA[r * ld + c] = v;
  }
  for (int c0 = 0; c0 < 192; c0 += 64) {
    tile_gemm<4, 8, 64, false>(plainA(sA, ld), 128, w.projt + c0, 192, Cp, sB, [&](int r, int c, float v) {
      if (r < nr) out[(r0 + r) * 192 + c0 + c] = v + __ldg(w.proj_b + c0 + c);
    });
  }
}

int launch_fine_windows(const MatchW& w, const float* feat_fine, int h, int wd, int C, int stride, int coarse_w,
                        const int64_t* j_ids, int64_t Mm, float* out, cudaStream_t st) {
  if (C != w.C) return set_error("fine_windows: channel count differs from the packed weights");
  const int Cp = (C + 31) / 32 * 32;
  const size_t smem = (STAGE_FLOATS + 128 * (Cp + 4)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fine_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const int64_t rows = Mm * 49;
  fine_windows_kernel<<<(unsigned)((rows + 127) / 128), NT, smem, st>>>(w, feat_fine, h, wd, C, stride, coarse_w,
                                                                     (const long long*)j_ids, Mm, out);
  return check_launch("fine_windows_kernel");
}

// ---- fine matching: two matches (2 x 49 rows) per CTA -----------------------------------------------------------------
constexpr int FM_SMEM_FLOATS = STAGE_FLOATS + 128 * LDD + 128 * LDH2 + 128 + 64;

__global__ void __launch_bounds__(NT, 1)
fine_match_kernel(const PairMlp m, const float* __restrict__ f0, const float* __restrict__ f1, const int64_t Mm,
                  const float* __restrict__ mk_c, float* __restrict__ expec, float* __restrict__ mk_f) {
  extern __shared__ __align__(16) float smem[];
  float* sB = smem;
  float* sX = sB + STAGE_FLOATS;  // [128][LDD]  f0 * f1 rows
  float* sH = sX + 128 * LDD;
  float* sL = sH + 128 * LDH2;    // [128] logits
  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * 2;
  for (int i = tid; i < 128 * 48; i += NT) {
    const int r = i / 48, c4 = i % 48;
    const int64_t t = t0 + r / 49;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < 98 && t < Mm) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(f0 + t * 192 + c4 * 4));
      const float4 b = __ldg(reinterpret_cast<const float4*>(f1 + (t * 49 + r % 49) * 192 + c4 * 4));
      v = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    }
    *reinterpret_cast<float4*>(sX + r * LDD + c4 * 4) = v;
  }
  tile_gemm<8, 8, 128, false>(plainA(sX, LDD), 128, m.w1t, 128, 192, sB,
                              [&](int r, int c, float v) { sH[r * LDH2 + c] = fmaxf(v + __ldg(m.b1 + c), 0.f); });
  pair_tail<8, NSTG>(m, sH, 128, sB, [&](int r, float logit) { sL[r] = logit; });
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < 2 && t0 + warp < Mm) {
    const int64_t t = t0 + warp;
    const float temp = 1.f / sqrtf(192.f);  // softmax_temp = 1 / C**.5
    const float l0 = lane < 49 ? sL[warp * 49 + lane] * temp : -FLT_MAX;
    const float l1 = lane + 32 < 49 ? sL[warp * 49 + lane + 32] * temp : -FLT_MAX;
    float mx = fmaxf(l0, l1);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = expf(l0 - mx), e1 = lane + 32 < 49 ? expf(l1 - mx) : 0.f;
    const float den = warp_sum(e0 + e1);
    const float h0 = e0 / den, h1 = e1 / den;
    // normalised 7x7 grid: (i / 6 - 0.5) * 2
    const int i0 = lane, i1 = lane + 32;
    const float gx0 = ((float)(i0 % 7) / 6.f - 0.5f) * 2.f, gy0 = ((float)(i0 / 7) / 6.f - 0.5f) * 2.f;
    const float gx1 = ((float)(i1 % 7) / 6.f - 0.5f) * 2.f, gy1 = ((float)(i1 / 7) / 6.f - 0.5f) * 2.f;
    const float ex = warp_sum(h0 * gx0 + h1 * gx1), ey = warp_sum(h0 * gy0 + h1 * gy1);
    const float exx = warp_sum(gx0 * gx0 * h0 + gx1 * gx1 * h1), eyy = warp_sum(gy0 * gy0 * h0 + gy1 * gy1 * h1);
    if (lane == 0) {
      const float vx = exx - ex * ex, vy = eyy - ey * ey;
      const float sd = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));
      expec[t * 3] = ex; expec[t * 3 + 1] = ey; expec[t * 3 + 2] = sd;
      mk_f[t * 2] = mk_c[t * 2] + ex * 3.f;
      mk_f[t * 2 + 1] = mk_c[t * 2 + 1] + ey * 3.f;
    }
  }
}

int launch_fine_match(const MatchW& w, const float* f0, const float* f1, int64_t Mm, const float* mkps2d_c,
                      float* expec_f, float* mkps2d_f, cudaStream_t st) {
  const size_t smem = FM_SMEM_FLOATS * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fine_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  fine_match_kernel<<<(unsigned)((Mm + 1) / 2), NT, smem, st>>>(w.fine, f0, f1, Mm, mkps2d_c, expec_f, mkps2d_f);
  return check_launch("fine_match_kernel");
}

// ---- out[n] = sigmoid(x[n, :K] . w + b) -------------------------------------------------------------------------------
__global__ void rowdot_sigmoid_kernel(const float* __restrict__ x, int64_t N, int K, const float* __restrict__ wv,
                                      const float* __restrict__ b, float* __restrict__ out) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = __ldg(b);
  for (int k = 0; k < K; ++k) a = fmaf(x[n * K + k], __ldg(wv + k), a);
  out[n] = 1.f / (1.f + expf(-a));
}

int launch_rowdot_sigmoid(const float* x, int64_t N, int K, const float* wv, const float* b, float* out, cudaStream_t st) {
  if (N <= 0) return 0;
  rowdot_sigmoid_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(x, N, K, wv, b, out);
  return check_launch("rowdot_sigmoid_kernel");
}

}  // namespace nlb
