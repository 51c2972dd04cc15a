// Launchers of the render-path kernels (internal).
#pragma once
#include "nlb_internal.h"

namespace nlb {

// Where the sample points come from: an explicit list (query API) or rays x depth samples (render API).
struct PointSrc {
  const float* xyz;     // [N][3] or null
  const float* dirs;    // [N][3] viewing direction per point, or null
  const float* rays_o;  // [R][3]
  const float* rays_d;  // [R][3]
  const float* z;       // [S] shared by all rays (zs == 0) or [R][S] per ray (zs == S, hierarchical sampling)
  int S;
  int64_t zs;
};

// Peer copies of the rendered-feature matrix (ray sharding across the GPUs of one NVSwitch box): the ray kernels store
// feat row (row0 + ray) into every listed buffer - the local one and the peers' mapped over NVLink - so the all-gather
// of the rendered features that precedes matching is part of the render epilogue instead of a separate collective.
struct FeatPeers {
  float* p[8];
  int n;
  int64_t row0;
};

// `visdd_scratch` ([N][V] float2, or null): visibility / depth difference per (sample, view) computed by visibility_kernel
// (visibility.cu, decoder on tcgen05) ahead of aggregate_kernel; null keeps the decoder inside aggregate_kernel
// `g_scratch` ([N][416]) and `q_out` ([N][128]), or null: out_fc and the attention query projection as 128-sample tensor-core
// GEMMs (fc_tail_kernel) instead of inside aggregate_kernel.  Returns 0 (done), 2 (done, and `q_out` holds W_q aggregated) or 1.
int launch_aggregate(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, int with_blend, float* agg,
                     float* partial, float* rgbvis, unsigned char* nvalid, float* mvf, float* mvv, float* visdd_scratch,
                     float* g_scratch, float* q_out, cudaStream_t st, bool agg_pm = false);
// `q` is written in the piece-major layout of pm128_off, `agg` too if `agg_pm` (both hold pm128_floats(N) floats then)
int launch_fc_tail(const RenderW& w, const float* g, int64_t N, float* agg, float* q, bool agg_pm, cudaStream_t st);
int launch_visibility(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, float* visdd, float* mvv, cudaStream_t st);
// second generation (neighbor2.cu): q projection, 32-sample super-tiles with the attention projections on tcgen05, fc + LayerNorm
// tail; `scratch` holds neighbor2_scratch_floats(N) floats
size_t neighbor2_scratch_floats(int64_t N);
// outputs: `fagg` fp32 [N][128] and / or `fagg_split` (the pair ray kernel's operand layout, samples grouped into rays of
// `S_split`); either may be null
// `q_ready`: scratch already holds q = W_q aggregated (fc_tail_kernel wrote it, piece-major: pm128_off)
int launch_neighbor2(const SceneDev& sc, const RenderW& w, const PointSrc& ps, int64_t N, int K, const int* idx,
                     const float* d2, const float* agg, float* fagg, unsigned char* fagg_split, int S_split, float* feature,
                     float* weights, float* scratch, bool q_ready, cudaStream_t st, bool agg_pm = false);   // agg_pm: `agg` is in the pm128_off layout
int launch_blend_project(const float* feat, int64_t P, const float* bl1v, float* out, cudaStream_t st);
int launch_linear(const float* A, int64_t N, int K, int lda, const float* Wt, const float* bias, int Nout, int act,
                  float* out, int ldo, cudaStream_t st);
int launch_sup_geo(const float* xyz, const float* dir, const float* conf, int64_t M, float* out, cudaStream_t st);

// pair kernel (render_ray2.cu): two rays per CTA, bf16x3 tcgen05; same contract as launch_ray
// `xsplit`: feature_agg pre-split into bf16 hi | lo planes per ray, [R][2][16 chunks][S][8] (written by launch_neighbor2)
int launch_ray2(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t zs, int64_t R, int S, int white_bkgd,
                const unsigned char* xsplit, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
                float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
                const FeatPeers& peers, cudaStream_t st);

// hierarchical sampling (hier_sample.cu)
int launch_hier_sample(const SceneDev& sc, const RenderW& w, const float* center_host, const float* dirs, int64_t R,
                       const float* z_coarse, const float* z_reg, int S, const float* u, int NI, float* z_out,
                       float* depth_coarse, int64_t* inds, cudaStream_t st);
// rays with 128 < S <= 256 samples (render_ray_long.cu): persistent CTAs, activations in per-CTA global slabs
constexpr int RL_MAX_GRID = 296;
size_t ray_long_slab_floats(int S);
int launch_ray_long(const SceneDev& sc, const RenderW& w, const float* z_vals, int64_t zs, int64_t R, int S, int white_bkgd,
                    const float* fagg, const float* partial, const float* rgbvis, const unsigned char* nvalid, float* rgb,
                    float* depth, float* weights, unsigned char* mask, float* depth_unc, float* feat, float* sigma_dbg,
                    float* slabs, const FeatPeers& peers, cudaStream_t st);

}  // namespace nlb
