// Back-projection of the depth pixels of one reference view (conditional_nerf/model.py:203-265) with the ROUNDING of
// the reference's CPU operators, so that the support points - the KNN's input, where one ulp flips near-tied neighbours -
// are bit-identical with the reference run on the host:
//   * torch.matmul on a reduction of 3 / 4 (MKL sgemm): ascending fused multiply-add chain, first term a plain product;
//   * elementwise mul / add / sub / div: one IEEE rounding each;
//   * torch.sum over the last dimension of 3: sequential adds of separately rounded products;
//   * torch.norm over 3 elements: sqrt of an ascending fma chain of the squares.
// The 3x3 / 4x4 matrices (inverse intrinsics, pose products) are computed by the caller with the reference's own host ops.
#include "nlb_internal.h"

namespace nlb {

struct BackprojMats {
  float kinv[9];   // torch.inverse(K)                      model.py:236
  float rot[9];    // c2w[:3,:3]
  float trans[3];  // c2w[:3,3]
  float toref[12]; // rows 0..2 of (inverse(c2w_ref) @ c2w)  model.py:240-241
  float fx, fy, cx, cy;  // K after the stride division      model.py:214-216 (get_rays, utils.py:56-70)
};

__global__ void __launch_bounds__(256) backproject_kernel(BackprojMats m, const long long* __restrict__ uu,
                                                          const long long* __restrict__ vv, const float* __restrict__ zz,
                                                          long long M, float* __restrict__ world, float* __restrict__ ref,
                                                          float* __restrict__ dir) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float u = (float)uu[i], v = (float)vv[i], z = zz[i];
  float cam[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float p = __fmaf_rn(m.kinv[j * 3 + 2], 1.0f, __fmaf_rn(m.kinv[j * 3 + 1], v, __fmul_rn(m.kinv[j * 3], u)));
    cam[j] = __fmul_rn(p, z);
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float r = __fmaf_rn(m.rot[j * 3 + 2], cam[2], __fmaf_rn(m.rot[j * 3 + 1], cam[1], __fmul_rn(m.rot[j * 3], cam[0])));
    world[i * 3 + j] = __fadd_rn(r, m.trans[j]);
    const float* t = m.toref + j * 4;
    ref[i * 3 + j] = __fmaf_rn(t[3], 1.0f, __fmaf_rn(t[2], cam[2], __fmaf_rn(t[1], cam[1], __fmul_rn(t[0], cam[0]))));
  }
  const float d0 = __fdiv_rn(__fsub_rn(u, m.cx), m.fx), d1 = __fdiv_rn(__fsub_rn(v, m.cy), m.fy), d2 = 1.0f;
  float rd[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    rd[j] = __fadd_rn(__fadd_rn(__fmul_rn(d0, m.rot[j * 3]), __fmul_rn(d1, m.rot[j * 3 + 1])), __fmul_rn(d2, m.rot[j * 3 + 2]));
  const float nrm = __fsqrt_rn(__fmaf_rn(rd[2], rd[2], __fmaf_rn(rd[1], rd[1], __fmul_rn(rd[0], rd[0]))));
  float4 o;
  o.x = __fdiv_rn(rd[0], nrm);
  o.y = __fdiv_rn(rd[1], nrm);
  o.z = __fdiv_rn(rd[2], nrm);
  o.w = z;
  reinterpret_cast<float4*>(dir)[i] = o;
}

int launch_backproject(const float* mats_host, const long long* uu, const long long* vv, const float* zz, long long M,
                       float* world, float* ref, float* dir, cudaStream_t st) {
  if (M <= 0) return 0;
  BackprojMats m;
  memcpy(&m, mats_host, sizeof(m));
  backproject_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(m, uu, vv, zz, M, world, ref, dir);
  return check_launch("backproject_kernel");
}

}  // namespace nlb
