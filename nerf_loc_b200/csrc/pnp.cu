// Absolute pose from 2D-3D matches on the device: the stage right after matching
// (nerf_loc/models/nerf_pose_estimator.py:557-583, `pycolmap.absolute_pose_estimation(p2d, p3d, PINHOLE, thresh)`).
//
// COLMAP is a third-party dependency that is not part of the reference tree (parity unpinned); what is
// built here is its published algorithm shape, all in fp64:
//   pnp_hypotheses_kernel  one thread per RANSAC sample: three correspondences -> Grunert's P3P (quartic in v = s3/s1 built
//                          by polynomial arithmetic, roots by Durand-Kerner + Newton polish) -> up to 4 poses
//   pnp_score_kernel       one CTA per candidate pose: MSAC score sum_i min(e_i^2, thr^2) over all correspondences
//   pnp_refine_kernel      one CTA: arg-min over the scores, then local optimisation: inliers -> Levenberg-Marquardt on the
//                          reprojection error (6x6 normal equations by block reduction) -> inliers, repeated lo_rounds times
// The restatement in oracle/pnp_oracle.py follows the same steps; both are validated against the known synthetic pose.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nlb_internal.h"

namespace nlb {

struct PnpPose {
  double R[9];
  double t[3];
  double valid;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Cplx { double re, im; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return Cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
  const double d = b.re * b.re + b.im * b.im;
  return Cplx{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// real roots of c[0] x^4 + c[1] x^3 + c[2] x^2 + c[3] x + c[4]; returns how many were written
__device__ int quartic_real_roots(const double* c, double* out) {
  if (fabs(c[0]) < 1e-14) return 0;
  const double a3 = c[1] / c[0], a2 = c[2] / c[0], a1 = c[3] / c[0], a0 = c[4] / c[0];
  const double rad = 1.0 + fmax(fmax(fabs(a3), fabs(a2)), fmax(fabs(a1), fabs(a0)));
  if (!(rad < 1e12)) return 0;
  Cplx r[4];
  Cplx seed{0.4, 0.9}, p{1.0, 0.0};
  const double r0 = fmin(rad, 1.0 + pow(fabs(a0), 0.25));
  for (int k = 0; k < 4; ++k) { r[k] = Cplx{p.re * r0, p.im * r0}; p = cmul(p, seed); }
  for (int it = 0; it < 120; ++it) {
    double delta = 0.0;
    for (int k = 0; k < 4; ++k) {
      const Cplx x = r[k];
      // Horner
      Cplx v{1.0, 0.0};
      v = cmul(v, x); v.re += a3;
      v = cmul(v, x); v.re += a2;
      v = cmul(v, x); v.re += a1;
      v = cmul(v, x); v.re += a0;
      Cplx den{1.0, 0.0};
      for (int j = 0; j < 4; ++j)
        if (j != k) den = cmul(den, csub(x, r[j]));
      if (den.re * den.re + den.im * den.im < 1e-300) continue;
      const Cplx d = cdiv(v, den);
      r[k] = csub(x, d);
      delta = fmax(delta, fabs(d.re) + fabs(d.im));
    }
    if (delta < 1e-14 * rad) break;
  }
  int n = 0;
  for (int k = 0; k < 4; ++k) {
    if (fabs(r[k].im) > 1e-7 * fmax(1.0, fabs(r[k].re))) continue;
    double x = r[k].re;
    for (int it = 0; it < 2; ++it) {  // Newton polish on the real polynomial
      const double f = (((x + a3) * x + a2) * x + a1) * x + a0;
      const double df = ((4.0 * x + 3.0 * a3) * x + 2.0 * a2) * x + a1;
      if (fabs(df) > 1e-300) x -= f / df;
    }
    out[n++] = x;
  }
  return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// orthonormal frame of three points, columns e1 | e2 | e3 (row-major 3x3); false if degenerate
__device__ bool triad(const double A[3][3], double* F) {
  double e1[3], d2[3], e3[3], e2[3];
  for (int i = 0; i < 3; ++i) { e1[i] = A[1][i] - A[0][i]; d2[i] = A[2][i] - A[0][i]; }
  const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  cross3(e1, d2, e3);
  const double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  if (n1 < 1e-12 || n3 < 1e-12) return false;
  for (int i = 0; i < 3; ++i) { e1[i] /= n1; e3[i] /= n3; }
  cross3(e3, e1, e2);
  for (int i = 0; i < 3; ++i) { F[i * 3 + 0] = e1[i]; F[i * 3 + 1] = e2[i]; F[i * 3 + 2] = e3[i]; }
  return true;
}

__global__ void __launch_bounds__(128)
pnp_hypotheses_kernel(const float* __restrict__ p2d, const float* __restrict__ p3d, const int64_t M, const double fx,
                      const double fy, const double cx, const double cy, const int iters, const uint64_t seed,
                      PnpPose* __restrict__ poses) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= iters) return;
  for (int s = 0; s < 4; ++s) poses[(size_t)h * 4 + s].valid = 0.0;
  uint64_t st = seed * 0xD1342543DE82EF95ull + (uint64_t)h * 0x9E3779B97F4A7C15ull + 1;
  int64_t id[3];
  id[0] = (int64_t)(splitmix64(st) % (uint64_t)M);
  do { id[1] = (int64_t)(splitmix64(st) % (uint64_t)M); } while (id[1] == id[0]);
  do { id[2] = (int64_t)(splitmix64(st) % (uint64_t)M); } while (id[2] == id[0] || id[2] == id[1]);
  double j[3][3], P[3][3];
  for (int k = 0; k < 3; ++k) {
    const double bx = ((double)p2d[id[k] * 2] - cx) / fx, by = ((double)p2d[id[k] * 2 + 1] - cy) / fy;
    const double n = sqrt(bx * bx + by * by + 1.0);
    j[k][0] = bx / n; j[k][1] = by / n; j[k][2] = 1.0 / n;
    for (int i = 0; i < 3; ++i) P[k][i] = (double)p3d[id[k] * 3 + i];
  }
  auto d2 = [&](int a, int b) {
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (P[a][i] - P[b][i]) * (P[a][i] - P[b][i]);
    return s;
  };
  auto dot = [&](int a, int b) { return j[a][0] * j[b][0] + j[a][1] * j[b][1] + j[a][2] * j[b][2]; };
  const double a2 = d2(1, 2), b2 = d2(0, 2), c2 = d2(0, 1);
  if (fmin(a2, fmin(b2, c2)) < 1e-18) return;
  const double ca = dot(1, 2), cb = dot(0, 2), cg = dot(0, 1);
  const double Kq = (a2 - c2) / b2;
  // u = N(v) / D(v), highest power first
  const double N[3] = {Kq - 1.0, -2.0 * Kq * cb, 1.0 + Kq};
  const double D[2] = {-2.0 * ca, 2.0 * cg};
  const double Q[3] = {-c2 / b2, 2.0 * c2 / b2 * cb, 1.0 - c2 / b2};
  double DD[3] = {D[0] * D[0], 2.0 * D[0] * D[1], D[1] * D[1]};
  double q[5] = {0, 0, 0, 0, 0};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) q[a + b] += DD[a] * Q[b] + N[a] * N[b];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 2; ++b) q[a + b + 1] -= 2.0 * cg * N[a] * D[b];
  double roots[4];
  const int nr = quartic_real_roots(q, roots);
  double Fw[9];
  if (!triad(P, Fw)) return;
  int ns = 0;
  for (int k = 0; k < nr; ++k) {
    const double v = roots[k];
    if (!(v > 0.0)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (fabs(den) < 1e-12) continue;
    const double u = ((Kq - 1.0) * v * v - 2.0 * Kq * cb * v + 1.0 + Kq) / den;
    const double wq = 1.0 + v * v - 2.0 * v * cb;
    if (!(u > 0.0) || !(wq > 0.0)) continue;
    const double s1 = sqrt(b2 / wq);
    const double sc[3] = {s1, u * s1, v * s1};
    double X[3][3];
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < 3; ++i) X[a][i] = sc[a] * j[a][i];
    double Fc[9];
    if (!triad(X, Fc)) continue;
    PnpPose& o = poses[(size_t)h * 4 + ns];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int m = 0; m < 3; ++m) s += Fc[r * 3 + m] * Fw[c * 3 + m];   // Fc * Fw^T
        o.R[r * 3 + c] = s;
      }
    for (int r = 0; r < 3; ++r) o.t[r] = X[0][r] - (o.R[r * 3] * P[0][0] + o.R[r * 3 + 1] * P[0][1] + o.R[r * 3 + 2] * P[0][2]);
    o.valid = 1.0;
    ++ns;
  }
}

__device__ __forceinline__ double reproj_err2(const double* R, const double* t, const float* p2d, const float* p3d, int64_t i,
                                              double fx, double fy, double cx, double cy, double* Xc) {
  const double X = p3d[i * 3], Y = p3d[i * 3 + 1], Z = p3d[i * 3 + 2];
  const double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
  const double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
  const double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
  if (Xc) { Xc[0] = x; Xc[1] = y; Xc[2] = z; }
  if (!(z > 1e-9)) return 1e300;
  const double du = fx * x / z + cx - (double)p2d[i * 2], dv = fy * y / z + cy - (double)p2d[i * 2 + 1];
  return du * du + dv * dv;
}

template <int NTH>
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NTH / 32; ++i) s += sh[i];
  return s;
}

__global__ void __launch_bounds__(128)
pnp_score_kernel(const float* __restrict__ p2d, const float* __restrict__ p3d, const int64_t M, const double fx, const double fy,
                 const double cx, const double cy, const double thr2, const PnpPose* __restrict__ poses,
                 double* __restrict__ score) {
  __shared__ double sh[4];
  const PnpPose& ps = poses[blockIdx.x];
  if (ps.valid == 0.0) {
    if (threadIdx.x == 0) score[blockIdx.x] = 1e300;
    return;
  }
  double R[9], t[3];
  for (int i = 0; i < 9; ++i) R[i] = ps.R[i];
  for (int i = 0; i < 3; ++i) t[i] = ps.t[i];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < M; i += 128) s += fmin(reproj_err2(R, t, p2d, p3d, i, fx, fy, cx, cy, nullptr), thr2);
  s = block_sum_d<128>(s, sh);
  if (threadIdx.x == 0) score[blockIdx.x] = s;
}

// in-place solve of the symmetric 6x6 system A x = b by Gaussian elimination with partial pivoting; false if singular
__device__ bool solve6(double A[6][6], double* b) {
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (fabs(A[p][c]) < 1e-300) return false;
    if (p != c) {
      for (int k = 0; k < 6; ++k) { const double tmp = A[c][k]; A[c][k] = A[p][k]; A[p][k] = tmp; }
      const double tmp = b[c]; b[c] = b[p]; b[p] = tmp;
    }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < 6; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int c = 5; c >= 0; --c) {
    double s = b[c];
    for (int k = c + 1; k < 6; ++k) s -= A[c][k] * b[k];
    b[c] = s / A[c][c];
  }
  return true;
}

__device__ void exp_so3(const double* w, double* E) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double a = 1.0, b = 0.5;
  if (th > 1e-12) { a = sin(th) / th; b = (1.0 - cos(th)) / (th * th); }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double kk = 0.0;
      for (int m = 0; m < 3; ++m) kk += K[r * 3 + m] * K[m * 3 + c];
      E[r * 3 + c] = (r == c ? 1.0 : 0.0) + a * K[r * 3 + c] + b * kk;
    }
}

constexpr int PNP_NT = 256;

__global__ void __launch_bounds__(PNP_NT)
pnp_refine_kernel(const float* __restrict__ p2d, const float* __restrict__ p3d, const int64_t M, const double fx, const double fy,
                  const double cx, const double cy, const double thr2, const PnpPose* __restrict__ poses,
                  const double* __restrict__ score, const int n_poses, const int lo_rounds, double* __restrict__ pose_out,
                  unsigned char* __restrict__ inliers, int* __restrict__ result) {
  __shared__ double sh[PNP_NT / 32];
  __shared__ double sBest[PNP_NT];
  __shared__ int sBestI[PNP_NT];
  __shared__ double sR[9], sT[3], cR[9], cT[3], sHg[27];
  __shared__ int sFlag;
  const int tid = threadIdx.x;
  // ---- arg-min over the candidate scores (ties -> lowest index) ----------------------------------------------------------
  double bs = 1e300;
  int bi = -1;
  for (int i = tid; i < n_poses; i += PNP_NT)
    if (score[i] < bs) { bs = score[i]; bi = i; }
  sBest[tid] = bs; sBestI[tid] = bi;
  __syncthreads();
  for (int o = PNP_NT / 2; o > 0; o >>= 1) {
    if (tid < o) {
      const double s2 = sBest[tid + o];
      const int i2 = sBestI[tid + o];
      if (i2 >= 0 && (s2 < sBest[tid] || (s2 == sBest[tid] && (sBestI[tid] < 0 || i2 < sBestI[tid])))) { sBest[tid] = s2; sBestI[tid] = i2; }
    }
    __syncthreads();
  }
  if (sBestI[0] < 0 || !(sBest[0] < 1e299)) {
    if (tid == 0) { result[0] = 0; result[1] = 0; }
    for (int64_t i = tid; i < M; i += PNP_NT) inliers[i] = 0;
    return;
  }
  if (tid < 9) sR[tid] = poses[sBestI[0]].R[tid];
  if (tid < 3) sT[tid] = poses[sBestI[0]].t[tid];
  __syncthreads();

  auto mark_inliers = [&]() -> int {
    double R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = sR[i];
    for (int i = 0; i < 3; ++i) t[i] = sT[i];
    double cnt = 0.0;
    for (int64_t i = tid; i < M; i += PNP_NT) {
      const bool in = reproj_err2(R, t, p2d, p3d, i, fx, fy, cx, cy, nullptr) < thr2;
      inliers[i] = in ? 1 : 0;
      cnt += in ? 1.0 : 0.0;
    }
    return (int)block_sum_d<PNP_NT>(cnt, sh);
  };
  auto cost_of = [&](const double* Rp, const double* tp) -> double {
    double R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = Rp[i];
    for (int i = 0; i < 3; ++i) t[i] = tp[i];
    double c = 0.0;
    for (int64_t i = tid; i < M; i += PNP_NT)
      if (inliers[i]) c += fmin(reproj_err2(R, t, p2d, p3d, i, fx, fy, cx, cy, nullptr), 1e12);
    return block_sum_d<PNP_NT>(c, sh);
  };

  int n_in = mark_inliers();
  for (int round = 0; round < lo_rounds && n_in >= 4; ++round) {
    double lam = 1e-3;
    double c0 = cost_of(sR, sT);
    for (int it = 0; it < 10; ++it) {
      // normal equations over the inliers: 21 upper-triangular entries of J^T J and 6 of J^T r
      double acc[27];
      for (int k = 0; k < 27; ++k) acc[k] = 0.0;
      {
        double R[9], t[3];
        for (int i = 0; i < 9; ++i) R[i] = sR[i];
        for (int i = 0; i < 3; ++i) t[i] = sT[i];
        for (int64_t i = tid; i < M; i += PNP_NT) {
          if (!inliers[i]) continue;
          double X[3];
          reproj_err2(R, t, p2d, p3d, i, fx, fy, cx, cy, X);
          const double x = X[0], y = X[1], z = fmax(X[2], 1e-9);
          const double ru = fx * x / z + cx - (double)p2d[i * 2], rv = fy * y / z + cy - (double)p2d[i * 2 + 1];
          const double du[3] = {fx / z, 0.0, -fx * x / (z * z)};
          const double dv[3] = {0.0, fy / z, -fy * y / (z * z)};
          double Ju[6], Jv[6];
          Ju[0] = du[2] * y - du[1] * z; Ju[1] = du[0] * z - du[2] * x; Ju[2] = du[1] * x - du[0] * y;
          Jv[0] = dv[2] * y - dv[1] * z; Jv[1] = dv[0] * z - dv[2] * x; Jv[2] = dv[1] * x - dv[0] * y;
          for (int k = 0; k < 3; ++k) { Ju[3 + k] = du[k]; Jv[3 + k] = dv[k]; }
          int q = 0;
          for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) acc[q++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
          for (int a = 0; a < 6; ++a) acc[21 + a] += Ju[a] * ru + Jv[a] * rv;
        }
      }
      for (int k = 0; k < 27; ++k) {
        const double s = block_sum_d<PNP_NT>(acc[k], sh);
        if (tid == 0) sHg[k] = s;
      }
      __syncthreads();
      bool accepted = false;
      for (int tr = 0; tr < 8 && !accepted; ++tr) {
        if (tid == 0) {
          double A[6][6], b[6];
          int q = 0;
          for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) { A[a][c] = sHg[q]; A[c][a] = sHg[q]; ++q; }
          for (int a = 0; a < 6; ++a) { A[a][a] += lam * A[a][a]; b[a] = -sHg[21 + a]; }
          sFlag = solve6(A, b) ? 1 : 0;
          if (sFlag) {
            double E[9];
            exp_so3(b, E);
            for (int r = 0; r < 3; ++r) {
              for (int c = 0; c < 3; ++c) cR[r * 3 + c] = E[r * 3] * sR[c] + E[r * 3 + 1] * sR[3 + c] + E[r * 3 + 2] * sR[6 + c];
              cT[r] = E[r * 3] * sT[0] + E[r * 3 + 1] * sT[1] + E[r * 3 + 2] * sT[2] + b[3 + r];
            }
          }
        }
        __syncthreads();
        if (sFlag) {
          const double c1 = cost_of(cR, cT);
          if (c1 < c0) {
            __syncthreads();
            if (tid < 9) sR[tid] = cR[tid];
            if (tid < 3) sT[tid] = cT[tid];
            c0 = c1;
            lam = fmax(lam * 0.1, 1e-9);
            accepted = true;
          } else {
            lam *= 10.0;
          }
        } else {
          lam *= 10.0;
        }
        __syncthreads();
      }
      if (!accepted) break;
    }
    n_in = mark_inliers();
  }
  if (tid < 9) pose_out[tid] = sR[tid];
  if (tid < 3) pose_out[9 + tid] = sT[tid];
  if (tid == 0) { result[0] = n_in >= 4 ? 1 : 0; result[1] = n_in; }
}

size_t pnp_scratch_bytes(int iters) {
  const size_t n = (size_t)(iters < 1 ? 1 : iters) * 4;
  return (n * sizeof(PnpPose) + 255) / 256 * 256 + n * sizeof(double) + 256;
}

int launch_pnp(const float* p2d, const float* p3d, int64_t M, const float* cam, float thresh, int iters, uint64_t seed,
               int lo_rounds, double* pose_out, unsigned char* inliers, int* result, void* scratch, cudaStream_t st) {
  if (M < 4) return set_error("pnp: at least 4 correspondences are required");
  if (iters < 1) return set_error("pnp: iters must be positive");
  char* p = reinterpret_cast<char*>(scratch);
  PnpPose* poses = reinterpret_cast<PnpPose*>(p);
  double* score = reinterpret_cast<double*>(p + (((size_t)iters * 4 * sizeof(PnpPose) + 255) / 256) * 256);
  const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], thr2 = (double)thresh * thresh;
  pnp_hypotheses_kernel<<<(iters + 127) / 128, 128, 0, st>>>(p2d, p3d, M, fx, fy, cx, cy, iters, seed, poses);
  if (check_launch("pnp_hypotheses_kernel")) return 1;
  pnp_score_kernel<<<iters * 4, 128, 0, st>>>(p2d, p3d, M, fx, fy, cx, cy, thr2, poses, score);
  if (check_launch("pnp_score_kernel")) return 1;
  pnp_refine_kernel<<<1, PNP_NT, 0, st>>>(p2d, p3d, M, fx, fy, cx, cy, thr2, poses, score, iters * 4, lo_rounds, pose_out,
                                          inliers, result);
  return check_launch("pnp_refine_kernel");
}

}  // namespace nlb
