// Launchers of the matcher kernels (internal).
#pragma once
#include "nlb_internal.h"

namespace nlb {

struct PairMlp {          // 192 -> 128 -> 128 -> 1, ReLU
  const float *w1t, *b1;  // [192][128], [128]
  const float *w2t, *b2;  // [128][128], [128]
  const float *w3, *b3;   // [128], [1]
};
struct MatchW {
  PairMlp coarse, fine;
  const float *projt, *proj_b;  // [Cp][192] (Cp = C rounded up to 32), [192]
  const float* tb_w2c;          // coarse pair MLP, layer 2 [128 x 128] as bf16 hi | lo K-tiles of 32 (s2d_tc.cu)
  int C;
};

size_t match_weights_floats(int C);
int match_weights_pack(const float* const* params, int n_params, int C, float* packed, size_t packed_floats,
                       cudaStream_t st);
MatchW match_weights_view(const float* packed, int C);

int launch_s2d(const MatchW& w, const float* desc0, const float* desc1, int64_t N, int64_t M, float* score,
               cudaStream_t st);
int launch_s2d_tc(const MatchW& w, const float* desc0, const float* desc1, int64_t N, int64_t M, float* score,
                  cudaStream_t st);
int launch_mutual(const float* score, int64_t N, int64_t M, float thr, int64_t* i_ids, int64_t* j_ids, int* count,
                  void* scratch, cudaStream_t st);
int launch_fine_windows(const MatchW& w, const float* feat_fine, int h, int wd, int C, int stride, int coarse_w,
                        const int64_t* j_ids, int64_t Mm, float* out, cudaStream_t st);
int launch_fine_match(const MatchW& w, const float* f0, const float* f1, int64_t Mm, const float* mkps2d_c,
                      float* expec_f, float* mkps2d_f, cudaStream_t st);
int launch_rowdot_sigmoid(const float* x, int64_t N, int K, const float* wv, const float* b, float* out, cudaStream_t st);

// pack.cu helpers reused by match.cu
__global__ void pack_t_kernel(float* dst, const float* __restrict__ src, int Kp, int N, int dst_ld, int src_ld,
                              int src_off, int Kv);
__global__ void pack_copy_kernel(float* dst, const float* __restrict__ src, int n);
__global__ void pack_tcb16_kernel(uint16_t* dst, const float* __restrict__ src, int N, int K, int src_ld, int src_off, int src_ks,
                                  int ktile, int Kv, int perm, int n_off, int N_total);

// absolute pose (pnp.cu)
size_t pnp_scratch_bytes(int iters);
int launch_pnp(const float* p2d, const float* p3d, int64_t M, const float* cam, float thresh, int iters, uint64_t seed,
               int lo_rounds, double* pose_out, unsigned char* inliers, int* result, void* scratch, cudaStream_t st);

}  // namespace nlb
