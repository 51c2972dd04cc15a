// extern "C" surface of libnerfloc_b200.so (declared in include/nerfloc_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/nerfloc_b200.h"
#include "match_kernels.h"
#include "nlb_internal.h"
#include "render_kernels.h"

namespace nlb {

static thread_local char g_err[512] = "";

int set_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "unknown error");
  return 1;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  char buf[480];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  return set_error(buf);
}

static SceneDev to_dev(const nlb_scene* s) {
  SceneDev d{};
  d.V = s->V; d.H = s->H; d.W = s->W; d.h = s->h; d.w = s->w; d.vh = s->vh; d.vw = s->vw;
  d.images = s->images; d.feat = s->featmaps; d.vis = s->vis_maps; d.cams = s->cams;
  d.near_ = s->near_plane; d.far_ = s->far_plane;
  d.M = s->M; d.sup_pre = s->sup_pre; d.sup_geo = s->sup_geo; d.knn = s->knn_index;
  d.qc[0] = s->query_center[0]; d.qc[1] = s->query_center[1]; d.qc[2] = s->query_center[2];
  d.featb = s->featmaps_blend;
  return d;
}

static int check_scene(const nlb_scene* s, bool need_support = true) {
  if (!s) return set_error("scene is NULL");
  if (!s->images || !s->featmaps || !s->vis_maps || !s->cams) return set_error("scene: NULL map / camera pointer");
  if (need_support && (!s->sup_pre || !s->sup_geo || !s->knn_index || s->M < 1))
    return set_error("scene: support points not prepared");
  if (s->V < 1 || s->V > 16) return set_error("scene: number of reference views must be in 1..16");
  if (s->H < 2 || s->W < 2 || s->h < 2 || s->w < 2 || s->vh < 2 || s->vw < 2) return set_error("scene: maps must be at least 2x2");
  return 0;
}

// Optional per-kernel device timing of nlb_render_rays (bench.py's roofline figures): a CUDA event on the launch stream after
// every kernel of a chunk, named by the launcher that recorded it (prof_mark).  One synchronisation per chunk while enabled,
// otherwise the schedule is the one that is timed.  The accumulators are guarded by a mutex: callers on several host threads
// add into the same table.
}  // namespace nlb
#include <atomic>
#include <mutex>
namespace nlb {
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
constexpr int PROF_MAX = 12;
static struct { const char* name; double ms; int64_t n; } g_prof_tab[PROF_MAX];
static int g_prof_rows = 0;
struct Prof {
  cudaStream_t st;
  cudaEvent_t ev[PROF_MAX + 1];
  const char* names[PROF_MAX + 1];
  int n = 0;
  bool made = false;
  explicit Prof(cudaStream_t s) : st(s) {}
  void mark(const char* name) {   // `name`: the kernel that ran since the previous mark (null for the first mark of a chunk)
    if (!g_prof_on.load(std::memory_order_relaxed) || n > PROF_MAX) return;
    if (!made) { for (auto& e : ev) cudaEventCreate(&e); made = true; }
    names[n] = name;
    cudaEventRecord(ev[n++], st);
  }
  void flush() {
    if (!g_prof_on.load(std::memory_order_relaxed) || n < 2) { n = 0; return; }
    cudaEventSynchronize(ev[n - 1]);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      int r = 0;
      while (r < g_prof_rows && strcmp(g_prof_tab[r].name, names[i]) != 0) ++r;
      if (r == g_prof_rows) { if (r == PROF_MAX) continue; g_prof_tab[r] = {names[i], 0.0, 0}; ++g_prof_rows; }
      g_prof_tab[r].ms += ms; g_prof_tab[r].n += 1;
    }
    n = 0;
  }
  ~Prof() { if (made) for (auto& e : ev) cudaEventDestroy(e); }
};
static thread_local Prof* g_cur_prof = nullptr;
void prof_mark(const char* name) { if (g_cur_prof) g_cur_prof->mark(name); }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }



static FeatPeers peers_at(FeatPeers p, int64_t row0) {
  p.row0 = row0;
  return p;
}

struct Carver {
  char* p;
  size_t left;
  bool ok = true;
  template <class T>
  T* take(size_t n) {
    const size_t b = align256(n * sizeof(T));
    if (b > left) { ok = false; return nullptr; }
    T* r = reinterpret_cast<T*>(p);
    p += b; left -= b;
    return r;
  }
};

}  // namespace nlb

namespace nlb { int launch_backproject(const float* mats_host, const long long* uu, const long long* vv, const float* zz, long long M,
                                      float* world, float* ref, float* dir, cudaStream_t st); }
namespace nlb { int launch_tc_test(const float* A, const float* W, int K, int mode, float* C, cudaStream_t st); }
namespace nlb { int read_prof(long long* out, int n); int read_prof_ray2(long long* out, int n); int read_prof_nb2(long long* out, int n); }
using namespace nlb;

extern "C" {

const char* nlb_last_error(void) { return g_err; }
int nlb_version(void) { return 100; }

size_t nlb_knn_index_bytes(int64_t M) { return knn_index_bytes(M); }

int nlb_knn_build(const float* p2, int64_t M, void* index, size_t index_bytes, void* stream) {
  if (!p2 || !index) return set_error("nlb_knn_build: NULL pointer");
  return knn_build(p2, M, index, index_bytes, (cudaStream_t)stream);
}

int nlb_knn_query(const void* index, const float* p1, int64_t N, int K, int64_t* idx, float* dist2, void* stream) {
  if (!index || (N > 0 && (!p1 || !idx || !dist2))) return set_error("nlb_knn_query: NULL pointer");
  return knn_query(index, p1, N, K, idx, nullptr, dist2, (cudaStream_t)stream);
}

int nlb_render_param_count(void) { return 100; }
size_t nlb_render_weights_floats(int S) { return render_weights_floats(S); }

int nlb_render_pack_weights(const float* const* params, int n_params, int S, float* packed, size_t packed_floats,
                            void* stream) {
  if (!params || !packed) return set_error("nlb_render_pack_weights: NULL pointer");
  return render_weights_pack(params, n_params, S, packed, packed_floats, (cudaStream_t)stream);
}

int nlb_support_prepare(const float* packed_weights, int S, const float* xyz, const float* feature,
                        const float* confidence, const float* direction, int64_t M, float* sup_pre, float* sup_geo,
                        void* stream) {
  if (!packed_weights || !xyz || !feature || !confidence || !direction || !sup_pre || !sup_geo)
    return set_error("nlb_support_prepare: NULL pointer");
  const RenderW w = render_weights_view(packed_weights, S);
  if (launch_linear(feature, M, C_RGBF, C_RGBF, w.w1a, w.b1, W_HID, 0, sup_pre, W_HID, (cudaStream_t)stream)) return 1;
  return launch_sup_geo(xyz, direction, confidence, M, sup_geo, (cudaStream_t)stream);
}

int nlb_blend_prepare(const float* packed_weights, int S, const float* featmaps, int64_t n_pixels, float* featmaps_blend,
                      void* stream) {
  if (!packed_weights || !featmaps || !featmaps_blend) return set_error("nlb_blend_prepare: NULL pointer");
  const RenderW w = render_weights_view(packed_weights, S);
  return launch_blend_project(featmaps, n_pixels, w.bl1v, featmaps_blend, (cudaStream_t)stream);
}

size_t nlb_query_scratch_bytes(int64_t N, int K) {
  if (N < 1) N = 1;
  return align256((size_t)N * K * 4) + align256((size_t)N * K * 4) + align256((size_t)N * W_HID * 4) +
         align256(neighbor2_scratch_floats(N) * 4) + align256((size_t)N * 16 * 8) +   // visibility | depth difference: <= 16 views
         align256((size_t)N * 416 * 4) + 1024;                                            // out_fc input of fc_tail_kernel
}

int nlb_query_points(const nlb_scene* scene, const float* packed_weights, int S, const float* xyz,
                     const float* direction, int64_t N, int K, float* feature_agg, float* feature, float* weights,
                     float* mv_feature, float* mv_visibility, float* aggregated, int32_t* knn_idx, float* knn_d2,
                     void* scratch, size_t scratch_bytes, void* stream) {
  if (check_scene(scene)) return 1;
  if (N <= 0) return 0;
  if (!packed_weights || !xyz || !feature_agg) return set_error("nlb_query_points: NULL pointer");
  if (K != 1 && K != 2 && K != 4 && K != 8) return set_error("nlb_query_points: K must be 1, 2, 4 or 8");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{(char*)scratch, scratch_bytes};
  int* idx = knn_idx ? knn_idx : c.take<int>((size_t)N * K);
  float* d2 = knn_d2 ? knn_d2 : c.take<float>((size_t)N * K);
  float* agg = aggregated ? aggregated : c.take<float>((size_t)N * W_HID);
  float* nb2 = c.take<float>(neighbor2_scratch_floats(N));
  float* visdd = c.take<float>((size_t)N * scene->V * 2);
  float* gvec = c.take<float>((size_t)N * 416);
  if (!c.ok) return set_error("nlb_query_points: scratch too small (see nlb_query_scratch_bytes)");
  const SceneDev sc = to_dev(scene);
  const RenderW w = render_weights_view(packed_weights, S);
  PointSrc ps{xyz, direction, nullptr, nullptr, nullptr, 1, 0};
  if (knn_query(sc.knn, xyz, N, K, nullptr, idx, d2, st)) return 1;
  const int ar = launch_aggregate(sc, w, ps, N, 0, agg, nullptr, nullptr, nullptr, mv_feature, mv_visibility, visdd,
                                  gvec, nb2, st);
  if (ar == 1) return 1;
  return launch_neighbor2(sc, w, ps, N, K, idx, d2, agg, feature_agg, nullptr, 0, feature, weights, nb2, ar == 2, st);
}

int nlb_aggregate_points(const nlb_scene* scene, const float* packed_weights, int S, const float* xyz, int64_t N,
                         float* aggregated, float* mv_feature, float* mv_visibility, void* stream) {
  if (check_scene(scene, false)) return 1;
  if (N <= 0) return 0;
  if (!packed_weights || !xyz || !aggregated) return set_error("nlb_aggregate_points: NULL pointer");
  const SceneDev sc = to_dev(scene);
  const RenderW w = render_weights_view(packed_weights, S);
  PointSrc ps{xyz, nullptr, nullptr, nullptr, nullptr, 1, 0};
  // (no scratch in this entry point's signature: the decoder stays inside aggregate_kernel)
  return launch_aggregate(sc, w, ps, N, 0, aggregated, nullptr, nullptr, nullptr, mv_feature, mv_visibility, nullptr, nullptr,
                          nullptr, (cudaStream_t)stream);
}

int nlb_descriptor_head(const float* packed_weights, int S, int level, const float* x, int64_t N, float* out,
                        void* stream) {
  if (!packed_weights || !x || !out) return set_error("nlb_descriptor_head: NULL pointer");
  const RenderW w = render_weights_view(packed_weights, S);
  return launch_linear(x, N, 323, 323, level == 0 ? w.pj_c : w.pj_f, level == 0 ? w.pj_c_b : w.pj_f_b, 192, 0, out, 192,
                       (cudaStream_t)stream);
}

int nlb_confidence_head(const float* packed_weights, int S, const float* aggregated, int64_t N, float* conf,
                        float* scratch, void* stream) {
  if (!packed_weights || !aggregated || !conf || !scratch) return set_error("nlb_confidence_head: NULL pointer");
  const RenderW w = render_weights_view(packed_weights, S);
  if (launch_linear(aggregated, N, W_HID, W_HID, w.cf1, w.cf1_b, 64, 1, scratch, 64, (cudaStream_t)stream)) return 1;
  return launch_rowdot_sigmoid(scratch, N, 64, w.cf2, w.cf2_b, conf, (cudaStream_t)stream);
}

int nlb_backproject_points(const float* mats_host, const int64_t* uu, const int64_t* vv, const float* zz, int64_t M, float* world,
                           float* ref, float* dir, void* stream) {
  if (!mats_host) return set_error("nlb_backproject_points: NULL matrices");
  if (M > 0 && (!uu || !vv || !zz || !world || !ref || !dir)) return set_error("nlb_backproject_points: NULL pointer");
  if (M < 0) return set_error("nlb_backproject_points: negative point count");
  return launch_backproject(mats_host, (const long long*)uu, (const long long*)vv, zz, M, world, ref, dir, (cudaStream_t)stream);
}

size_t nlb_render_scratch_bytes(int64_t chunk_rays, int S, int V) {
  const size_t n = (size_t)(chunk_rays < 1 ? 1 : chunk_rays) * S;
  const size_t slabs = S > 128 ? align256((size_t)RL_MAX_GRID * ray_long_slab_floats(S) * 4) : 0;
  return align256(n * KNN_K * 4) * 2 + align256(pm128_floats((int64_t)n) * 4) + align256(n * W_HID * 4) + align256((n * V + 31) / 32 * 32 * 32 * 4) + align256(n * V * 16) +
         align256(n) + align256(neighbor2_scratch_floats((int64_t)n) * 4) + align256(n * V * 8) + align256(n * 416 * 4) + slabs + 2048;
}

int64_t nlb_render_launch_count(int64_t R, int64_t chunk_rays) {
  if (R <= 0) return 0;
  if (chunk_rays < 1) chunk_rays = R;
  // KNN search, visibility, aggregate, fc_tail, neighbor2, attention tail, ray
  return 7 * ((R + chunk_rays - 1) / chunk_rays);
}

static int render_rays_impl(const nlb_scene* scene, const float* packed_weights, int S, const float* rays_o,
                            const float* rays_d, const float* z_vals, int64_t z_stride, int64_t R, int white_bkgd, int64_t chunk_rays,
                            float* rgb, float* depth, float* weights, uint8_t* mask, float* depth_uncertainty, float* feat,
                            float* dbg_feature_agg, float* dbg_sigma, void* scratch, size_t scratch_bytes, void* stream,
                            float* const* feat_peers, int n_peers, int64_t feat_row0) {
  if (check_scene(scene)) return 1;
  if (R <= 0) return 0;
  if (!packed_weights || !rays_o || !rays_d || !z_vals || !rgb || !depth || !weights || !mask || !depth_uncertainty)
    return set_error("nlb_render_rays: NULL pointer");
  if (!scene->featmaps_blend) return set_error("nlb_render_rays: scene.featmaps_blend is NULL (call nlb_blend_prepare once per frame)");
  if (S % 8 != 0 || S < 8 || S > 256) return set_error("nlb_render_rays: S must be a multiple of 8 in [8, 256]");
  if (z_stride != 0 && z_stride != S) return set_error("nlb_render_rays: z_stride must be 0 (shared depths) or S (per-ray depths)");
  if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !feat_peers)) return set_error("nlb_render_rays_gather: 0..8 peer buffers");
  FeatPeers peers{};
  peers.n = n_peers;
  for (int p = 0; p < n_peers; ++p) {
    if (!feat_peers[p]) return set_error("nlb_render_rays_gather: NULL peer buffer");
    peers.p[p] = feat_peers[p];
  }
  if (chunk_rays < 1) chunk_rays = R;
  if (chunk_rays > R) chunk_rays = R;
  cudaStream_t st = (cudaStream_t)stream;
  const int V = scene->V;
  const size_t n = (size_t)chunk_rays * S;
  Carver c{(char*)scratch, scratch_bytes};
  int* idx = c.take<int>(n * KNN_K);
  float* d2 = c.take<float>(n * KNN_K);
  float* agg = c.take<float>(pm128_floats((int64_t)n));   // piece-major (pm128_off), whole groups of 32 rows
  float* fagg = c.take<float>(n * W_HID);
  float* partial = c.take<float>((n * V + 31) / 32 * 32 * 32);   // whole groups of 32 rows (partial_off)
  float* rgbvis = c.take<float>(n * V * 4);
  unsigned char* nvalid = c.take<unsigned char>(n);
  float* nb2 = c.take<float>(neighbor2_scratch_floats((int64_t)n));
  float* visdd = c.take<float>(n * V * 2);
  float* gvec = c.take<float>(n * 416);
  float* slabs = S > 128 ? c.take<float>((size_t)RL_MAX_GRID * ray_long_slab_floats(S)) : nullptr;
  if (!c.ok) return set_error("nlb_render_rays: scratch too small (see nlb_render_scratch_bytes)");
  const SceneDev sc = to_dev(scene);
  const RenderW w = render_weights_view(packed_weights, S);
  const int64_t nchunks = (R + chunk_rays - 1) / chunk_rays;
  // (The KNN search of chunk i+1 used to run on a side stream underneath the ray kernel of chunk i; the pair ray kernel fills
  // the SM's registers and shared memory, so the overlap only cost launch gaps - measured 417 k vs 429 k rays/s - and is gone.)
  auto knn_chunk = [&](int64_t i, cudaStream_t s) {
    const int64_t r0 = i * chunk_rays;
    const int64_t rc = (R - r0) < chunk_rays ? (R - r0) : chunk_rays;
    return knn_query_rays(sc.knn, rays_o + r0 * 3, rays_d + r0 * 3, z_vals + r0 * z_stride, z_stride, sc.sup_geo, rc, S,
                          idx, d2, s);
  };
  int rc_err = 0;
  Prof prof(st);
  g_cur_prof = &prof;
  for (int64_t i = 0; i < nchunks && !rc_err; ++i) {
    const int64_t r0 = i * chunk_rays;
    const int64_t rc = (R - r0) < chunk_rays ? (R - r0) : chunk_rays;
    const int64_t nc = rc * S;
    const float* ro = rays_o + r0 * 3;
    const float* rd = rays_d + r0 * 3;
    const float* zc = z_vals + r0 * z_stride;
    PointSrc ps{nullptr, nullptr, ro, rd, zc, S, z_stride};
    float* fa = dbg_feature_agg ? dbg_feature_agg + r0 * S * W_HID : fagg;
    // the pair ray kernel takes feature_agg pre-split into its bf16 operand layout (written into the same scratch by the
    // attention tail); an fp32 copy is only produced for the debug output
    const bool split_x = S <= 128;
    unsigned char* fsplit = split_x ? reinterpret_cast<unsigned char*>(fagg) : nullptr;
    float* fa32 = split_x ? (dbg_feature_agg ? fa : nullptr) : fa;
    prof.mark(nullptr);
    if (knn_chunk(i, st)) { rc_err = 1; break; }
    prof.mark("knn_query_rays");
    const int ar = launch_aggregate(sc, w, ps, nc, 1, agg, partial, rgbvis, nvalid, nullptr, nullptr, visdd, gvec, nb2, st, true);
    if (ar == 1) { rc_err = 1; break; }
    // (agg is piece-major only if fc_tail_kernel wrote it: ar == 2)
    if (launch_neighbor2(sc, w, ps, nc, KNN_K, idx, d2, agg, fa32, fsplit, S, nullptr, nullptr, nb2, ar == 2, st, ar == 2)) { rc_err = 1; break; }
    if (split_x) {
      if (launch_ray2(sc, w, zc, z_stride, rc, S, white_bkgd, fsplit, partial, rgbvis, nvalid, rgb + r0 * 3, depth + r0,
                      weights + r0 * S, mask + r0, depth_uncertainty + r0, feat ? feat + r0 * C_FEAT : nullptr,
                      dbg_sigma ? dbg_sigma + r0 * S : nullptr, peers_at(peers, feat_row0 + r0), st)) { rc_err = 1; break; }
    } else if (launch_ray_long(sc, w, zc, z_stride, rc, S, white_bkgd, fa, partial, rgbvis, nvalid, rgb + r0 * 3, depth + r0,
                               weights + r0 * S, mask + r0, depth_uncertainty + r0, feat ? feat + r0 * C_FEAT : nullptr,
                               dbg_sigma ? dbg_sigma + r0 * S : nullptr, slabs, peers_at(peers, feat_row0 + r0), st)) {
      rc_err = 1; break;
    }
    prof.mark(S <= 128 ? "ray" : "ray_long");
    prof.flush();
  }
  g_cur_prof = nullptr;
  return rc_err;
}

int nlb_render_rays(const nlb_scene* scene, const float* packed_weights, int S, const float* rays_o,
                    const float* rays_d, const float* z_vals, int64_t z_stride, int64_t R, int white_bkgd, int64_t chunk_rays,
                    float* rgb, float* depth, float* weights, uint8_t* mask, float* depth_uncertainty, float* feat,
                    float* dbg_feature_agg, float* dbg_sigma, void* scratch, size_t scratch_bytes, void* stream) {
  return render_rays_impl(scene, packed_weights, S, rays_o, rays_d, z_vals, z_stride, R, white_bkgd, chunk_rays, rgb, depth, weights,
                          mask, depth_uncertainty, feat, dbg_feature_agg, dbg_sigma, scratch, scratch_bytes, stream, nullptr, 0, 0);
}

int nlb_render_rays_gather(const nlb_scene* scene, const float* packed_weights, int S, const float* rays_o,
                           const float* rays_d, const float* z_vals, int64_t z_stride, int64_t R, int white_bkgd,
                           int64_t chunk_rays, float* rgb, float* depth, float* weights, uint8_t* mask, float* depth_uncertainty,
                           float* feat, void* scratch, size_t scratch_bytes, float* const* feat_peers, int n_peers,
                           int64_t feat_row0, void* stream) {
  return render_rays_impl(scene, packed_weights, S, rays_o, rays_d, z_vals, z_stride, R, white_bkgd, chunk_rays, rgb, depth, weights,
                          mask, depth_uncertainty, feat, nullptr, nullptr, scratch, scratch_bytes, stream, feat_peers, n_peers,
                          feat_row0);
}

int nlb_hierarchical_depths(const nlb_scene* scene, const float* packed_weights, int S_total, const float* center,
                            const float* dirs, int64_t R, const float* z_coarse, int n_samples, const float* z_regular,
                            const float* u, int n_importance, float* z_out, float* depth_coarse, int64_t* inds, void* stream) {
  if (check_scene(scene)) return 1;
  if (!packed_weights || !center || !dirs || !z_coarse || !z_regular || !u || !z_out || !depth_coarse)
    return set_error("nlb_hierarchical_depths: NULL pointer");
  if (n_samples + n_importance != S_total) return set_error("nlb_hierarchical_depths: S_total must be n_samples + n_importance");
  const SceneDev sc = to_dev(scene);
  const RenderW w = render_weights_view(packed_weights, S_total);
  return launch_hier_sample(sc, w, center, dirs, R, z_coarse, z_regular, n_samples, u, n_importance, z_out, depth_coarse, inds,
                            (cudaStream_t)stream);
}

void nlb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on.store(on != 0);
  g_prof_rows = 0;
}

int nlb_profile_report(char* buf, size_t n) {
  if (!buf || n == 0) return set_error("nlb_profile_report: NULL buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  size_t off = 0;
  buf[0] = 0;
  for (int r = 0; r < g_prof_rows; ++r) {
    const int k = snprintf(buf + off, n - off, "%s:%.6f:%lld;", g_prof_tab[r].name, g_prof_tab[r].ms, (long long)g_prof_tab[r].n);
    if (k < 0 || (size_t)k >= n - off) return set_error("nlb_profile_report: buffer too small");
    off += (size_t)k;
  }
  return 0;
}

size_t nlb_pnp_scratch_bytes(int iters) { return pnp_scratch_bytes(iters); }

int nlb_pnp_ransac(const float* p2d, const float* p3d, int64_t M, const float* camera, float thresh_px, int iters,
                   uint64_t seed, int lo_rounds, double* pose_w2c, uint8_t* inliers, int32_t* result, void* scratch,
                   size_t scratch_bytes, void* stream) {
  if (!p2d || !p3d || !camera || !pose_w2c || !inliers || !result || !scratch) return set_error("nlb_pnp_ransac: NULL pointer");
  if (scratch_bytes < pnp_scratch_bytes(iters)) return set_error("nlb_pnp_ransac: scratch too small (see nlb_pnp_scratch_bytes)");
  if (!(thresh_px > 0.f)) return set_error("nlb_pnp_ransac: threshold must be positive");
  return launch_pnp(p2d, p3d, M, camera, thresh_px, iters, seed, lo_rounds < 0 ? 0 : lo_rounds, pose_w2c, inliers, result,
                    scratch, (cudaStream_t)stream);
}

size_t nlb_match_weights_floats(int C) { return match_weights_floats(C); }

int nlb_match_pack_weights(const float* const* params, int n_params, int C, float* packed, size_t packed_floats,
                           void* stream) {
  if (!params || !packed) return set_error("nlb_match_pack_weights: NULL pointer");
  return match_weights_pack(params, n_params, C, packed, packed_floats, (cudaStream_t)stream);
}

int nlb_s2d_scores(const float* packed, int C, const float* desc0, const float* desc1, int64_t N, int64_t M,
                   float* score, void* stream) {
  if (N <= 0 || M <= 0) return set_error("nlb_s2d_scores: both descriptor sets must be non-empty (sparse_to_dense.py:123)");
  if (!packed || !desc0 || !desc1 || !score) return set_error("nlb_s2d_scores: NULL pointer");
  return launch_s2d(match_weights_view(packed, C), desc0, desc1, N, M, score, (cudaStream_t)stream);
}

size_t nlb_mutual_scratch_bytes(int64_t N, int64_t M) {
  return align256((size_t)(N < 1 ? 1 : N) * 8) + align256((size_t)(M < 1 ? 1 : M) * 4) + 1024;
}

int nlb_mutual_matches(const float* score, int64_t N, int64_t M, float thr, int64_t* i_ids, int64_t* j_ids,
                       int32_t* count, void* scratch, size_t scratch_bytes, void* stream) {
  if (!score || !i_ids || !j_ids || !count || !scratch) return set_error("nlb_mutual_matches: NULL pointer");
  if (N <= 0 || M <= 0) return set_error("nlb_mutual_matches: empty score matrix");
  if (scratch_bytes < nlb_mutual_scratch_bytes(N, M)) return set_error("nlb_mutual_matches: scratch too small");
  return launch_mutual(score, N, M, thr, i_ids, j_ids, count, scratch, (cudaStream_t)stream);
}

int nlb_fine_windows(const float* packed, int C, const float* feat_fine, int h, int w, int stride, int coarse_w,
                     const int64_t* j_ids, int64_t Mm, float* out, void* stream) {
  if (Mm <= 0) return 0;
  if (!packed || !feat_fine || !j_ids || !out) return set_error("nlb_fine_windows: NULL pointer");
  return launch_fine_windows(match_weights_view(packed, C), feat_fine, h, w, C, stride, coarse_w, j_ids, Mm, out,
                             (cudaStream_t)stream);
}

int nlb_fine_match(const float* packed, int C, const float* f0, const float* f1, int64_t Mm, const float* mkps2d_c,
                   float* expec_f, float* mkps2d_f, void* stream) {
  if (Mm <= 0) return 0;
  if (!packed || !f0 || !f1 || !mkps2d_c || !expec_f || !mkps2d_f) return set_error("nlb_fine_match: NULL pointer");
  return launch_fine_match(match_weights_view(packed, C), f0, f1, Mm, mkps2d_c, expec_f, mkps2d_f, (cudaStream_t)stream);
}

int nlb_debug_knn_rays(const void* index, const float* rays_o, const float* rays_d, const float* z_vals, int64_t z_stride,
                       const float* sup_geo, int64_t R, int S, int32_t* idx, float* dist2, void* stream) {
  if (!index || !rays_o || !rays_d || !z_vals || !sup_geo || !idx || !dist2) return set_error("nlb_debug_knn_rays: NULL pointer");
  return knn_query_rays(index, rays_o, rays_d, z_vals, z_stride, sup_geo, R, S, idx, dist2, (cudaStream_t)stream);
}

int nlb_debug_tc_gemm(const float* A, const float* W, int K, int mode, float* C, void* stream) {
  if (!A || !W || !C) return set_error("nlb_debug_tc_gemm: NULL pointer");
  return launch_tc_test(A, W, K, mode, C, (cudaStream_t)stream);
}

int nlb_debug_read_prof(long long* out, int n) {
  if (n <= 32) return read_prof(out, n);
  if (read_prof_ray2(out + 32, n - 32) || read_prof(out, 32)) return 1;
  return read_prof_nb2(out, 16);   // slots 0-15: neighbour stamps, 16-31: aggregate stamps, 32-63: ray stamps
}

}  // extern "C"
