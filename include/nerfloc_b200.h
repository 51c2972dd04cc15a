/* nerfloc_b200 - C ABI of the B200-native NeRF-Loc render-and-match hot path.
 *
 * Plain C, device pointers and sizes only (no torch types).  All tensors are fp32 unless stated, contiguous,
 * resident on the current CUDA device; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Every function returns 0 on success and a non-zero code on failure; nlb_last_error() then holds the reason.
 * Nothing is allocated behind the caller's back: scratch and outputs are caller-provided (torch owns memory).
 *
 * Each entry point names the reference interface it replaces (paths below /root/reference/nerf_loc/models/).
 * The reference has no C FFI for this path except the pybind KNN module (ops/knn/src/knn_api.cpp:10-14); the rest of
 * the boundary is Python call signatures on nn.Modules, mirrored by nerf_loc_b200/*.py on top of these functions.
 */
#ifndef NERFLOC_B200_H
#define NERFLOC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* nlb_last_error(void);
int nlb_version(void);

/* ---- exact K nearest neighbours --------------------------------------------------------------------------------
 * Replaces knn_points_idx (ops/knn/src/knn_api.cpp:12, ops/knn/src/knn.h:58-75; python wrapper
 * ops/knn/knn_utils.py:97-171) for D = 3, one cloud per call.  Squared L2, K smallest by (distance, index),
 * ascending, int64 indices, zero padding when the support set has fewer than K points - bit-identical to
 * knn_cpu.cpp:13-64.  The index over the support points is built once per frame. */
size_t nlb_knn_index_bytes(int64_t M);
int nlb_knn_build(const float* p2 /*[M,3]*/, int64_t M, void* index, size_t index_bytes, void* stream);
int nlb_knn_query(const void* index, const float* p1 /*[N,3]*/, int64_t N, int K /*1,2,4,8,16*/,
                  int64_t* idx /*[N,K]*/, float* dist2 /*[N,K]*/, void* stream);

/* ---- ConditionalNeRF weights -------------------------------------------------------------------------------------
 * `params` is a HOST array of device pointers to the tensors of the reference state_dict
 * (conditional_nerf/model.py:29-135) in the order of nerf_loc_b200/params.py::conditional_nerf_shapes(S)
 * (nlb_render_param_count() entries).  S = render.N_samples + render.N_importance the RayUnet was built for
 * (conditional_nerf/ray_unet.py:11-52); S = 0 packs everything but the RayUnet (query-only use). */
int nlb_render_param_count(void);
size_t nlb_render_weights_floats(int S);
int nlb_render_pack_weights(const float* const* params, int n_params, int S, float* packed, size_t packed_floats,
                            void* stream);

/* ---- per-frame scene --------------------------------------------------------------------------------------------
 * The tensors ConditionalNeRF reads from `data` plus its two per-frame caches (support_neural_points,
 * multiview_aggregator.vis_featmaps; reset protocol nerf_pose_estimator.py:289-290). */
typedef struct nlb_scene {
  int32_t V, H, W, h, w;     /* reference views, image size, feature-map size */
  int32_t vh, vw;            /* visibility-map size (always the fine level: H/4 x W/4, also for coarse queries) */
  float near_plane, far_plane;
  const float* images;       /* [V,H,W,4]  rgb + one pad channel (topk_images, channels last) */
  const float* featmaps;     /* [V,h,w,192] (feat_fine_src / feat_coarse_src, already channels last) */
  const float* vis_maps;     /* [V,vh,vw,32] DepthFusionNet output, channels last */
  const float* cams;         /* [V,32]: rows 0..2 of K_hom*inv(c2w) (12) | K*Rt (12) | camera centre (3) | pad (5) */
  int64_t M;                 /* support neural points of this level */
  const float* sup_pre;      /* [M,128] from nlb_support_prepare */
  const float* sup_geo;      /* [M,8]   from nlb_support_prepare */
  const void* knn_index;     /* from nlb_knn_build over the support xyz */
  float query_center[3];     /* camera centre of the query pose (render only) */
  const float* featmaps_blend; /* [V,h,w,32] from nlb_blend_prepare (render only; may be NULL for query / aggregate) */
} nlb_scene;

/* Per-frame pre-projection of the reference feature maps through the map-feature columns of the colour-blend MLP's first
 * layer (rgb_blending_mlp.0, conditional_nerf/model.py:90-96,532-535).  That layer is linear and so is the bilinear fetch of
 * the 192 map channels (ibrnet/ibrnet.py:194-231), so W f(x) = sum_t b_t (W f_t): a rendered sample gathers 32 projected
 * channels per view instead of multiplying its 195 fetched channels by a [195 x 32] matrix.
 * featmaps [n_pixels,192] (n_pixels = V*h*w) -> featmaps_blend [n_pixels,32]. */
int nlb_blend_prepare(const float* packed_weights, int S, const float* featmaps, int64_t n_pixels, float* featmaps_blend,
                      void* stream);

/* Per-frame precompute over the support points (conditional_nerf/model.py:372-375,404-405):
 * sup_pre = base_mlp.0.weight[:, :195] * feature + bias, sup_geo = (xyz, direction[:3], confidence). */
int nlb_support_prepare(const float* packed_weights, int S, const float* xyz /*[M,3]*/, const float* feature /*[M,195]*/,
                        const float* confidence /*[M,1]*/, const float* direction /*[M,4]*/, int64_t M,
                        float* sup_pre /*[M,128]*/, float* sup_geo /*[M,8]*/, void* stream);

/* ---- ConditionalNeRF.query (conditional_nerf/model.py:344-436) --------------------------------------------------
 * xyz [N,3]; direction [N,3] or NULL (then the nearest neighbour's direction is used, model.py:391-392); K <= 8.
 * Outputs (any may be NULL except feature_agg): feature_agg [N,128]; feature [N,128] (identical for every k, the
 * caller expands to [N,K,128]); weights [N,K]; mv_feature [N,V,195]; mv_visibility [N,V]; aggregated [N,128]
 * (MultiviewFeatureAggregator.forward output, multiview_aggregator.py:156-222); knn_idx int32 [N,K]; knn_d2 [N,K].
 * scratch: nlb_query_scratch_bytes(N, K). */
size_t nlb_query_scratch_bytes(int64_t N, int K);
int nlb_query_points(const nlb_scene* scene, const float* packed_weights, int S, const float* xyz,
                     const float* direction, int64_t N, int K, float* feature_agg, float* feature, float* weights,
                     float* mv_feature, float* mv_visibility, float* aggregated, int32_t* knn_idx, float* knn_d2,
                     void* scratch, size_t scratch_bytes, void* stream);

/* ---- MultiviewFeatureAggregator.forward alone (conditional_nerf/multiview_aggregator.py:156-222) ---------------------
 * Used by the per-frame confidence pass over the support points (model.py:137-142,172-175); the scene needs no
 * support points yet.  aggregated [N,128]; mv_feature [N,V,195] / mv_visibility [N,V] may be NULL. */
int nlb_aggregate_points(const nlb_scene* scene, const float* packed_weights, int S, const float* xyz, int64_t N,
                         float* aggregated, float* mv_feature, float* mv_visibility, void* stream);

/* Linear heads applied to query results: proj_layer_3d_{coarse,fine} (model.py:132-133,303-306,333-336),
 * level 0 = coarse, 1 = fine; x [N,323] = cat(feature_agg, support feature) -> out [N,192]. */
int nlb_descriptor_head(const float* packed_weights, int S, int level, const float* x, int64_t N, float* out,
                        void* stream);
/* confidence_mlp (model.py:52-57,137-142): aggregated [N,128] -> conf [N,1]; scratch N*64 floats. */
int nlb_confidence_head(const float* packed_weights, int S, const float* aggregated, int64_t N, float* conf,
                        float* scratch, void* stream);

/* Back-projection of one reference view's depth pixels (model.py:203-265, the matmul / get_rays arithmetic at :236-247 and
 * conditional_nerf/utils.py:56-70) with the rounding of the reference's HOST operators (ascending fma chains of MKL sgemm on a
 * reduction of 3 / 4, single roundings elsewhere), so that xyz / xyz_ndc / direction are bit-identical with the reference run
 * on the CPU.  mats_host: 37 floats in HOST memory computed with the reference's own host ops - inverse(K)[9], c2w[:3,:3][9],
 * c2w[:3,3][3], rows 0..2 of inverse(c2w_ref) @ c2w [12], fx, fy, cx, cy of the stride-scaled K.  uu, vv [M] int64 pixel
 * coordinates (torch.nonzero order), zz [M] depths.  Outputs: world [M,3], ref [M,3], dir [M,4] = (unit ray direction, depth). */
int nlb_backproject_points(const float* mats_host, const int64_t* uu, const int64_t* vv, const float* zz, int64_t M, float* world,
                           float* ref, float* dir, void* stream);

/* ---- ConditionalNeRF.render_rays (conditional_nerf/model.py:472-600) ---------------------------------------------------
 * rays_o, rays_d [R,3] (unit directions, as conditional_nerf/utils.py:56-70 produces them); z_vals [S] with z_stride = 0
 * (sample_depths, model.py:451-458, shared by all rays) or [R,S] with z_stride = S (per-ray depths from
 * nlb_hierarchical_depths when N_importance > 0).  S in [8, 256], multiple of 8.  Outputs: rgb [R,3], depth [R], weights [R,S], mask [R] (uint8 0/1), depth_uncertainty [R],
 * feat [R,192] (NULL to skip render_feature).  Optional debug outputs (NULL to skip): feature_agg [R*S,128],
 * sigma [R*S].  Rays are processed in chunks of `chunk_rays`; scratch: nlb_render_scratch_bytes(chunk_rays, S, V). */
size_t nlb_render_scratch_bytes(int64_t chunk_rays, int S, int V);
int nlb_render_rays(const nlb_scene* scene, const float* packed_weights, int S, const float* rays_o,
                    const float* rays_d, const float* z_vals, int64_t z_stride, int64_t R, int white_bkgd,
                    int64_t chunk_rays,
                    float* rgb, float* depth, float* weights, uint8_t* mask, float* depth_uncertainty, float* feat,
                    float* dbg_feature_agg, float* dbg_sigma, void* scratch, size_t scratch_bytes, void* stream);
/* Ray-sharded rendering across the GPUs of one NVSwitch box: same as nlb_render_rays for this rank's
 * R rays, and the rendered 192-d feature of ray i is additionally stored as row (feat_row0 + i) of every buffer in
 * feat_peers[0..n_peers) (device pointers, n_peers <= 8): the rank's own gathered matrix and the peers', mapped into this
 * process (CUDA IPC / symmetric memory; the stores cross NVLink).  The all-gather of the rendered 3D features that the
 * matcher needs is thereby part of the ray kernel's epilogue; the caller only synchronises the ranks afterwards.  `feat`
 * (the local [R,192] output) may be NULL. */
int nlb_render_rays_gather(const nlb_scene* scene, const float* packed_weights, int S, const float* rays_o,
                           const float* rays_d, const float* z_vals, int64_t z_stride, int64_t R, int white_bkgd,
                           int64_t chunk_rays, float* rgb, float* depth, float* weights, uint8_t* mask,
                           float* depth_uncertainty, float* feat, void* scratch, size_t scratch_bytes,
                           float* const* feat_peers /*HOST array*/, int n_peers, int64_t feat_row0, void* stream);
/* ---- hierarchical sampling, render.N_importance > 0 (model.py:486-496; multiview_aggregator.py:95-154; utils.py:73-112) ---
 * center [3] (HOST) and dirs [R,3] (device): the query camera centre and the UN-normalised NeuRay ray directions
 * (depth_fusion.py:9-30) of the pixels; z_coarse [64] and z_regular [n_samples] from sample_depths; u [R,n_importance] the
 * uniform draws (torch.rand in the reference, utils.py:97).  Outputs: z_out [R, n_samples + n_importance] sorted per ray
 * (feed to nlb_render_rays with z_stride = S_total), depth_coarse [R], inds [R,n_importance] (the searchsorted indices,
 * NULL to skip).  The packed weights are the ones of S_total = n_samples + n_importance (model.py:81). */
int nlb_hierarchical_depths(const nlb_scene* scene, const float* packed_weights, int S_total, const float* center,
                            const float* dirs, int64_t R, const float* z_coarse, int n_samples, const float* z_regular,
                            const float* u, int n_importance, float* z_out, float* depth_coarse, int64_t* inds, void* stream);

/* number of kernels nlb_render_rays launches for R rays (bench.py's gpu_launches claim) */
int64_t nlb_render_launch_count(int64_t R, int64_t chunk_rays);

/* Optional per-kernel device timing of nlb_render_rays (CUDA events on the launch stream after every kernel, accumulated per
 * kernel name).  Enabling resets the table; while enabled every chunk ends with one synchronisation, the launch order is
 * unchanged.  nlb_profile_report writes "name:milliseconds:launches;" records into `buf`. */
void nlb_profile_enable(int on);
int nlb_profile_report(char* buf, size_t n);

/* ---- Matcher weights --------------------------------------------------------------------------------------------------
 * `params` is a HOST array of 14 device pointers, reference layouts (nerf_loc/models/matcher.py:22,40-61):
 *   coarse_matcher.mlps.{0,2,4}.{weight,bias}  [128,192],[128],[128,128],[128],[1,128],[1]      (entries 0..5)
 *   fine_matcher.mlps.{0,2,4}.{weight,bias}    same shapes                                         (entries 6..11)
 *   fine_preprocess.proj.{weight,bias}         [192,C],[192]                                       (entries 12,13)
 * C = channels of the fine feature map (<= 320). */
size_t nlb_match_weights_floats(int C);
int nlb_match_pack_weights(const float* const* params, int n_params, int C, float* packed, size_t packed_floats,
                           void* stream);

/* ---- S2DMatching (matching/sparse_to_dense.py:116-151) ------------------------------------------------------------
 * score[n,m] = sigmoid(MLP(desc0[n] * desc1[m])), MLP = 192->128->128->1 with ReLU (coarse_matcher weights).
 * nlb_mutual_matches applies the rule of sparse_to_dense.py:136-142 (score > thr AND row max AND column max, exact
 * float equality; a row's j is its first qualifying column) and writes the matches in ascending i order; count
 * receives the number of matches. */
int nlb_s2d_scores(const float* packed_match_weights, int C, const float* desc0 /*[N,192]*/,
                   const float* desc1 /*[M,192]*/, int64_t N, int64_t M, float* score /*[N,M]*/, void* stream);
size_t nlb_mutual_scratch_bytes(int64_t N, int64_t M);
int nlb_mutual_matches(const float* score, int64_t N, int64_t M, float thr, int64_t* i_ids /*[N]*/,
                       int64_t* j_ids /*[N]*/, int32_t* count /*[1], device*/, void* scratch, size_t scratch_bytes,
                       void* stream);

/* ---- fine stage (matching/fine_matching.py:35-153) ------------------------------------------------------------------
 * nlb_fine_windows: the 7x7 windows F.unfold(kernel 7, stride, padding 3) would produce for the matched coarse cells
 * only (fine_matching.py:53-57), projected by fine_preprocess.proj (fine_matching.py:74):
 * feat_fine [h,w,C] channels last, j_ids [Mm] (row-major cell index on a grid `coarse_w` wide) -> out [Mm,49,192].
 * nlb_fine_match: FineMatching.forward (fine_matching.py:109-153): f0 [Mm,192], f1 [Mm,49,192] (fine_matcher weights),
 * mkps2d_c [Mm,2] -> expec_f [Mm,3] (x, y, std), mkps2d_f [Mm,2]. */
int nlb_fine_windows(const float* packed_match_weights, int C, const float* feat_fine, int h, int w, int stride,
                     int coarse_w, const int64_t* j_ids, int64_t Mm, float* out, void* stream);
int nlb_fine_match(const float* packed_match_weights, int C, const float* f0, const float* f1, int64_t Mm,
                   const float* mkps2d_c, float* expec_f, float* mkps2d_f, void* stream);

/* ---- absolute pose from the 2D-3D matches (models/nerf_pose_estimator.py:557-583) ---------------------------------------
 * Replaces `pycolmap.absolute_pose_estimation(p2d, p3d, {PINHOLE, [fx, fy, cx, cy]}, ransac_thresh)` (line 574; COLMAP is a
 * third-party dependency outside the reference tree - parity unpinned, validated against known poses): P3P-RANSAC with an
 * MSAC score and Levenberg-Marquardt local optimisation on the inliers, fp64, all on the device.
 * p2d [M,2] pixels, p3d [M,3] world points (device); camera = {fx, fy, cx, cy} (HOST); pose_w2c [12] doubles (device):
 * R row-major then t, x_cam = R x_world + t; inliers [M] 0/1; result [2] = {success, number of inliers}. */
size_t nlb_pnp_scratch_bytes(int iters);
int nlb_pnp_ransac(const float* p2d, const float* p3d, int64_t M, const float* camera, float thresh_px, int iters,
                   uint64_t seed, int lo_rounds, double* pose_w2c, uint8_t* inliers, int32_t* result, void* scratch,
                   size_t scratch_bytes, void* stream);

/* ---- self-test of the tcgen05 building blocks: C[128,128] = A[128,K] * W[128,K]^T, K multiple of 8 <= 64;
 * mode 0 = single-pass tf32, 1 = 3xTF32 (fp32-equivalent), 2 = 3xTF32 with the A operand in tensor memory;
 * mode 3 = the warp-level path (mma.sync.m16n8k8 tf32, 3xTF32): C[16,128] = A[16,K] * W[K,128], W k-major, K = 32 or 64 ---- */
/* clock64() phase stamps of block 0 of the last neighbor_kernel launch (debug aid) */
int nlb_debug_read_prof(long long* out /*[n]*/, int n /*<= 64: 0..31 neighbour/aggregate, 32..63 ray kernel*/);
/* Test hook: the ray-sample K = 8 search the render path runs per chunk (sample n = r * S + s at o_r + d_r * z_s; results as
 * int32 indices / squared distances [R * S, 8]); same definition as nlb_knn_query (ops/knn/src/knn_cpu.cpp:13-64). */
int nlb_debug_knn_rays(const void* index, const float* rays_o, const float* rays_d, const float* z_vals, int64_t z_stride,
                       const float* sup_geo /*[M,8]*/, int64_t R, int S, int32_t* idx, float* dist2, void* stream);
int nlb_debug_tc_gemm(const float* A, const float* W, int K, int mode, float* C, void* stream);

#ifdef __cplusplus
}
#endif
#endif
