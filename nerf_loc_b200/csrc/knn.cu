// Exact K-nearest-neighbour search over the support neural points (sm_100a).
//
// Replaces the reference's brute-force KNN (nerf_loc/models/ops/knn/src/knn.cu:131-241, the vendored pytorch3d
// kernels selected for D=3, K=8 / K=1 at conditional_nerf/model.py:289,318,377) with an exact search over a
// bounding-volume hierarchy that is rebuilt once per frame:
//   * support points are sorted along a 30-bit Morton curve (cub radix sort, per-frame setup);
//   * leaves hold LEAF consecutive sorted points, inner levels group FAN consecutive children; every node
//     stores an ORIENTED bounding box (centre, principal axes of its points, half extents: four float4).  The
//     support points are samples of surfaces: an axis-aligned box around a tilted patch is as thick as the
//     patch is wide, and a far query's search sphere (tangent to the surface) then cuts through dozens of boxes
//     that hold no neighbour; along the patch's own normal the box is as thin as the surface is rough;
//   * the tree is walked nearest child first, pruning a node only when its box distance is STRICTLY larger than
//     the current K-th distance.
// Results are identical to the reference definition: the K smallest squared distances by (distance, index),
// ascending, with the squared distance accumulated as ((dx*dx + dy*dy) + dz*dz) with one rounding per
// operation (no FMA), exactly like knn_cpu.cpp:43-47.  The box distance is a lower bound of every contained
// point's COMPUTED distance by construction: extents are inflated and the per-axis gaps deflated by more than
// the rounding error of the projections (node_d2), so pruning is exact, ties included.
#include <cub/cub.cuh>
#include <float.h>
#include <stdlib.h>
#include "nlb_internal.h"

namespace nlb {

#ifndef NLB_KNN_LEAF
#define NLB_KNN_LEAF 8
#endif
#ifndef NLB_KNN_FAN
#define NLB_KNN_FAN 4      // measured on the 153,600-point frame (ray search, ms): fan-out 8: 134, 4: 116, 3: 116, 2: 142
#endif
constexpr int LEAF = NLB_KNN_LEAF;
constexpr int FAN = NLB_KNN_FAN;
constexpr int NODE_F4 = 4;     // float4 per node: (centre, h0) (axis 0, h1) (axis 1, h2) (axis 2, -)

// ---- index layout inside the caller-provided buffer -------------------------------------------------------
// header (KnnHeader) | sorted points float4[M] (xyz, original index bits) | level 0 boxes | level 1 boxes ...
struct KnnHeader {
  int64_t M;
  int n_levels;
  int level_count[12];
  int64_t level_off[12];  // float4 offsets (NODE_F4 float4 per node) from the start of the box area
  int64_t pts_off;        // byte offsets from buffer start
  int64_t box_off;
  int64_t keys_off, vals_off, keys2_off, vals2_off, bbox_off, cub_off;
  int64_t cub_bytes;
  int64_t total_bytes;
};

static void knn_layout(int64_t M, KnnHeader& h) {
  h.M = M;
  int64_t n = (M + LEAF - 1) / LEAF;
  int L = 0;
  int64_t boxes = 0;
  while (true) {
    h.level_count[L] = (int)n;
    h.level_off[L] = boxes * NODE_F4;
    boxes += n;
    ++L;
    if (n <= FAN || L >= 12) break;
    n = (n + FAN - 1) / FAN;
  }
  h.n_levels = L;
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t o = align(sizeof(KnnHeader));
  h.pts_off = o; o = align(o + M * 16);
  h.box_off = o; o = align(o + boxes * 16 * NODE_F4);
  h.keys_off = o; o = align(o + M * 4);
  h.vals_off = o; o = align(o + M * 4);
  h.keys2_off = o; o = align(o + M * 4);
  h.vals2_off = o; o = align(o + M * 4);
  h.bbox_off = o; o = align(o + 64);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)M, 0, 30);
  h.cub_off = o; h.cub_bytes = (int64_t)cub_bytes; o = align(o + cub_bytes);
  h.total_bytes = o;
}

size_t knn_index_bytes(int64_t M) {
  KnnHeader h;
  knn_layout(M < 1 ? 1 : M, h);
  return (size_t)h.total_bytes;
}

// ---- build kernels ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void knn_bbox_init(unsigned* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

__global__ void knn_bbox_kernel(const float* __restrict__ xyz, int64_t M, unsigned* bb) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = xyz[i * 3 + c];
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bb[c], f2ord(lo[c]));
      atomicMax(&bb[3 + c], f2ord(hi[c]));
    }
  }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void knn_morton_kernel(const float* __restrict__ xyz, int64_t M, const unsigned* __restrict__ bb,
                                  uint32_t* keys, uint32_t* vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  unsigned q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float lo = ord2f(bb[c]), hi = ord2f(bb[3 + c]);
    float e = fmaxf(hi - lo, 1e-20f);
    float t = (xyz[i * 3 + c] - lo) / e * 1023.f;
    t = fminf(fmaxf(t, 0.f), 1023.f);
    q[c] = (unsigned)t;
  }
  keys[i] = (spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]);
  vals[i] = (uint32_t)i;
}

__global__ void knn_gather_kernel(const float* __restrict__ xyz, int64_t M, const uint32_t* __restrict__ order,
                                  float4* pts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint32_t j = order[i];
  pts[i] = make_float4(xyz[(int64_t)j * 3], xyz[(int64_t)j * 3 + 1], xyz[(int64_t)j * 3 + 2], __uint_as_float(j));
}

// One warp per node: oriented bounding box of the node's points (a contiguous run of `span` sorted points).
//   pass 1: mean; pass 2: covariance about the mean, its eigenvectors (cyclic Jacobi, re-orthonormalised) are the box axes;
//   pass 3: extent of the projections along each axis -> box centre (mid-range) and half extents;
//   pass 4: the half extents are re-measured about the STORED centre with the arithmetic the queries use and inflated by more
//           than its rounding error, so |a_i . (p - c)| <= h_i holds for every point in exact arithmetic.
// aabb != 0 keeps the coordinate axes (an axis-aligned box in the same record; A/B switch NLB_KNN_AABB).
__global__ void __launch_bounds__(256) knn_node_obb_kernel(const float4* __restrict__ pts, int64_t M, int64_t span, int n_nodes,
                                                           float4* __restrict__ out, int aabb) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= n_nodes) return;
  const int64_t p0 = (int64_t)b * span, p1 = min(M, p0 + span);
  const float n = (float)(p1 - p0);
  auto wsum = [](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  auto wmax = [](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  };
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int64_t i = p0 + lane; i < p1; i += 32) { const float4 p = pts[i]; sx += p.x; sy += p.y; sz += p.z; }
  const float mx = wsum(sx) / n, my = wsum(sy) / n, mz = wsum(sz) / n;
  float ax[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};   // rows: box axes
  if (!aabb) {
    float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
    for (int64_t i = p0 + lane; i < p1; i += 32) {
      const float4 p = pts[i];
      const float dx = p.x - mx, dy = p.y - my, dz = p.z - mz;
      c00 += dx * dx; c01 += dx * dy; c02 += dx * dz; c11 += dy * dy; c12 += dy * dz; c22 += dz * dz;
    }
    float A[3][3];
    A[0][0] = wsum(c00); A[0][1] = A[1][0] = wsum(c01); A[0][2] = A[2][0] = wsum(c02);
    A[1][1] = wsum(c11); A[1][2] = A[2][1] = wsum(c12); A[2][2] = wsum(c22);
    float V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};   // columns: eigenvectors
    // every lane runs the same deterministic iteration on the same (shuffled) numbers
    for (int sweep = 0; sweep < 6; ++sweep) {
#pragma unroll
      for (int pq = 0; pq < 3; ++pq) {
        const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
        const float apq = A[p][q];
        if (fabsf(apq) <= 1e-30f) continue;
        const float theta = (A[q][q] - A[p][p]) / (2.f * apq);
        const float t = (theta >= 0.f ? 1.f : -1.f) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
        const float cs = 1.f / sqrtf(t * t + 1.f), sn = t * cs;
#pragma unroll
        for (int k = 0; k < 3; ++k) {   // A <- A J
          const float akp = A[k][p], akq = A[k][q];
          A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {   // A <- J^T A
          const float apk = A[p][k], aqk = A[q][k];
          A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {   // V <- V J
          const float vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq;
        }
      }
    }
    // Gram-Schmidt on the columns of V; a degenerate column falls back to a coordinate axis / cross product
    float e0[3] = {V[0][0], V[1][0], V[2][0]}, e1[3] = {V[0][1], V[1][1], V[2][1]};
    float l0 = sqrtf(e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2]);
    if (!(l0 > 1e-20f)) { e0[0] = 1.f; e0[1] = 0.f; e0[2] = 0.f; l0 = 1.f; }
    e0[0] /= l0; e0[1] /= l0; e0[2] /= l0;
    float dp = e1[0] * e0[0] + e1[1] * e0[1] + e1[2] * e0[2];
    e1[0] -= dp * e0[0]; e1[1] -= dp * e0[1]; e1[2] -= dp * e0[2];
    float l1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    if (!(l1 > 1e-6f)) {   // pick the coordinate axis least aligned with e0
      const int m = fabsf(e0[0]) <= fabsf(e0[1]) ? (fabsf(e0[0]) <= fabsf(e0[2]) ? 0 : 2) : (fabsf(e0[1]) <= fabsf(e0[2]) ? 1 : 2);
      e1[0] = m == 0; e1[1] = m == 1; e1[2] = m == 2;
      dp = e1[0] * e0[0] + e1[1] * e0[1] + e1[2] * e0[2];
      e1[0] -= dp * e0[0]; e1[1] -= dp * e0[1]; e1[2] -= dp * e0[2];
      l1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    }
    e1[0] /= l1; e1[1] /= l1; e1[2] /= l1;
    ax[0][0] = e0[0]; ax[0][1] = e0[1]; ax[0][2] = e0[2];
    ax[1][0] = e1[0]; ax[1][1] = e1[1]; ax[1][2] = e1[2];
    ax[2][0] = e0[1] * e1[2] - e0[2] * e1[1]; ax[2][1] = e0[2] * e1[0] - e0[0] * e1[2]; ax[2][2] = e0[0] * e1[1] - e0[1] * e1[0];
  }
  // extent of the projections about the mean -> centre at the mid-range
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = p0 + lane; i < p1; i += 32) {
    const float4 p = pts[i];
    const float dx = p.x - mx, dy = p.y - my, dz = p.z - mz;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float t = ax[a][0] * dx + ax[a][1] * dy + ax[a][2] * dz;
      lo[a] = fminf(lo[a], t); hi[a] = fmaxf(hi[a], t);
    }
  }
  float cx = mx, cy = my, cz = mz;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mid = 0.5f * (-wmax(-lo[a]) + wmax(hi[a]));
    cx += mid * ax[a][0]; cy += mid * ax[a][1]; cz += mid * ax[a][2];
  }
  float h[3] = {0.f, 0.f, 0.f}, span1 = 0.f;
  for (int64_t i = p0 + lane; i < p1; i += 32) {
    const float4 p = pts[i];
    const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
    span1 = fmaxf(span1, fabsf(dx) + fabsf(dy) + fabsf(dz));
#pragma unroll
    for (int a = 0; a < 3; ++a) h[a] = fmaxf(h[a], fabsf(fmaf(ax[a][2], dz, fmaf(ax[a][1], dy, ax[a][0] * dx))));
  }
  span1 = wmax(span1);
#pragma unroll
  for (int a = 0; a < 3; ++a) h[a] = wmax(h[a]) * 1.00001f + 2e-6f * span1;
  if (lane == 0) {
    out[(size_t)b * NODE_F4 + 0] = make_float4(cx, cy, cz, h[0]);
    out[(size_t)b * NODE_F4 + 1] = make_float4(ax[0][0], ax[0][1], ax[0][2], h[1]);
    out[(size_t)b * NODE_F4 + 2] = make_float4(ax[1][0], ax[1][1], ax[1][2], h[2]);
    out[(size_t)b * NODE_F4 + 3] = make_float4(ax[2][0], ax[2][1], ax[2][2], 0.f);
  }
}

int knn_build(const float* xyz, int64_t M, void* buf, size_t bytes, cudaStream_t st) {
  if (M < 1) return set_error("knn_build: empty support set");
  if (M >= (1ll << 31)) return set_error("knn_build: too many points");
  KnnHeader h;
  knn_layout(M, h);
  if (bytes < (size_t)h.total_bytes) return set_error("knn_build: index buffer too small");
  char* base = (char*)buf;
  unsigned* bb = (unsigned*)(base + h.bbox_off);
  uint32_t* keys = (uint32_t*)(base + h.keys_off);
  uint32_t* vals = (uint32_t*)(base + h.vals_off);
  uint32_t* keys2 = (uint32_t*)(base + h.keys2_off);
  uint32_t* vals2 = (uint32_t*)(base + h.vals2_off);
  float4* pts = (float4*)(base + h.pts_off);
  float4* boxes = (float4*)(base + h.box_off);
  const int T = 256;
  const int G = (int)((M + T - 1) / T);
  knn_bbox_init<<<1, 32, 0, st>>>(bb);
  knn_bbox_kernel<<<G < 592 ? G : 592, T, 0, st>>>(xyz, M, bb);
  knn_morton_kernel<<<G, T, 0, st>>>(xyz, M, bb, keys, vals);
  size_t cub_bytes = (size_t)h.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(base + h.cub_off, cub_bytes, keys, keys2, vals, vals2, (int)M, 0, 30, st);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  knn_gather_kernel<<<G, T, 0, st>>>(xyz, M, vals2, pts);
  static const int aabb = getenv("NLB_KNN_AABB") ? 1 : 0;   // A/B switch: axis-aligned instead of oriented boxes
  int64_t span = LEAF;
  for (int l = 0; l < h.n_levels; ++l, span *= FAN)
    knn_node_obb_kernel<<<(h.level_count[l] + 7) / 8, 256, 0, st>>>(pts, M, span, h.level_count[l], boxes + h.level_off[l], aabb);
  e = cudaMemcpyAsync(buf, &h, sizeof(h), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  // the header is read back from pageable host memory: make sure the copy has consumed it
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  return 0;
}

// ---- query -------------------------------------------------------------------------------------------------
struct KnnTree {
  const float4* pts;
  const float4* boxes;
  int64_t M;
  int n_levels;
  int level_count[12];
  int64_t level_off[12];
};

__device__ __forceinline__ float d2_exact(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Lower bound of the computed squared distance from q to every point of a node.  d = q - c is a correctly rounded difference
// of exact inputs; the three projections carry a rounding error below 3e-7 |d|_1, the stored half extents exceed the exact ones
// (knn_node_obb_kernel), so every per-axis gap below is an underestimate; the final factor covers the axes' deviation from
// orthonormality and the rounding of the point distance itself.
__device__ __forceinline__ float node_d2(float qx, float qy, float qz, const float4* __restrict__ nd) {
  const float4 f0 = nd[0], f1 = nd[1], f2 = nd[2], f3 = nd[3];
  const float dx = qx - f0.x, dy = qy - f0.y, dz = qz - f0.z;
  const float eps = 2e-6f * (fabsf(dx) + fabsf(dy) + fabsf(dz));
  const float l0 = fabsf(fmaf(f1.z, dz, fmaf(f1.y, dy, f1.x * dx)));
  const float l1 = fabsf(fmaf(f2.z, dz, fmaf(f2.y, dy, f2.x * dx)));
  const float l2 = fabsf(fmaf(f3.z, dz, fmaf(f3.y, dy, f3.x * dx)));
  const float g0 = fmaxf(l0 - f0.w - eps, 0.f), g1 = fmaxf(l1 - f1.w - eps, 0.f), g2 = fmaxf(l2 - f2.w - eps, 0.f);
  return (g0 * g0 + g1 * g1 + g2 * g2) * 0.99999f;
}

template <int K>
struct TopK {
  float d[K];
  int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) { d[i] = FLT_MAX; id[i] = 0x7fffffff; }
  }
  __device__ __forceinline__ float worst() const { return d[K - 1]; }
  // keep ascending by (dist, index)
  __device__ __forceinline__ void push(float dist, int idx) {
    if (!(dist < d[K - 1] || (dist == d[K - 1] && idx < id[K - 1]))) return;
    d[K - 1] = dist; id[K - 1] = idx;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      const bool sw = d[i] < d[i - 1] || (d[i] == d[i - 1] && id[i] < id[i - 1]);
      if (sw) {
        float td = d[i]; d[i] = d[i - 1]; d[i - 1] = td;
        int ti = id[i]; id[i] = id[i - 1]; id[i - 1] = ti;
      }
    }
  }
};

// `bound`: a squared distance that at least K support points are known to lie within (FLT_MAX if unknown).  Nothing
// strictly farther than it can be part of the answer, so it prunes exactly like the running K-th distance does.
template <int K>
__device__ __forceinline__ void knn_search(const KnnTree& t, float qx, float qy, float qz, TopK<K>& best,
                                           const float bound = FLT_MAX) {
  constexpr int STACK = 12 * FAN;
  unsigned long long stk[STACK];   // (distance bits << 32) | level << 28 | node: one 8-byte local-memory access per push / pop
  int sp = 0;
  const int top = t.n_levels - 1;
  // push the top level (<= FAN nodes unless the level cap was hit), farthest first
  for (int n = t.level_count[top] - 1; n >= 0; --n) {
    if (sp < STACK) {
      const float d0 = node_d2(qx, qy, qz, t.boxes + t.level_off[top] + NODE_F4 * n);
      stk[sp++] = ((unsigned long long)__float_as_uint(d0) << 32) | (((unsigned)top << 28) | (unsigned)n);
    }
  }
  // The nearest child of a node is visited next without a round trip through the stack (which lives in local memory): `cur`
  // holds it; the visiting order is the one of a stack that had it on top.
  unsigned cur = 0;
  float cur_d = 0.f;
  bool have_cur = false;
  while (have_cur || sp > 0) {
    float nd;
    unsigned code;
    if (have_cur) {
      nd = cur_d; code = cur; have_cur = false;
    } else {
      const unsigned long long e = stk[--sp];
      nd = __uint_as_float((unsigned)(e >> 32)); code = (unsigned)e;
    }
    if (nd > fminf(best.worst(), bound)) continue;  // strict: equal distance may still hide a smaller index
    const int lvl = code >> 28;
    const int node = code & 0x0fffffffu;
    if (lvl == 0) {
      const int64_t p0 = (int64_t)node * LEAF;
#pragma unroll
      for (int k = 0; k < LEAF; ++k) {
        if (p0 + k < t.M) {
          const float4 p = t.pts[p0 + k];
          const float d = d2_exact(qx, qy, qz, p.x, p.y, p.z);
          if (d <= bound) best.push(d, __float_as_int(p.w));
        }
      }
    } else {
      const int cl = lvl - 1;
      const int c0 = node * FAN;
      const int nc = min(FAN, t.level_count[cl] - c0);
      float cd[FAN];
      int nearest = -1;
      float nearest_d = FLT_MAX;
#pragma unroll
      for (int k = 0; k < FAN; ++k) {
        cd[k] = FLT_MAX;
        if (k < nc) {
          cd[k] = node_d2(qx, qy, qz, t.boxes + t.level_off[cl] + NODE_F4 * (c0 + k));
          if (cd[k] < nearest_d) { nearest_d = cd[k]; nearest = k; }
        }
      }
      const float w = fminf(best.worst(), bound);
#pragma unroll
      for (int k = 0; k < FAN; ++k) {
        if (k < nc && k != nearest && cd[k] <= w && sp < STACK) {
          stk[sp++] = ((unsigned long long)__float_as_uint(cd[k]) << 32) | (((unsigned)cl << 28) | (unsigned)(c0 + k));
        }
      }
      if (nearest >= 0 && nearest_d <= w) {
        cur = ((unsigned)cl << 28) | (unsigned)(c0 + nearest); cur_d = nearest_d; have_cur = true;
      }
    }
  }
}

__device__ __forceinline__ KnnTree load_tree(const void* index) {
  const KnnHeader* h = (const KnnHeader*)index;
  KnnTree t;
  t.M = h->M;
  t.n_levels = h->n_levels;
  t.pts = (const float4*)((const char*)index + h->pts_off);
  t.boxes = (const float4*)((const char*)index + h->box_off);
#pragma unroll
  for (int i = 0; i < 12; ++i) { t.level_count[i] = h->level_count[i]; t.level_off[i] = h->level_off[i]; }
  return t;
}

// Generic query: p1 [N,3] -> idx (int64 [N,K]) and/or idx32 (int32 [N,K]), dist2 [N,K].
template <int K>
__global__ void __launch_bounds__(128) knn_query_kernel(const void* __restrict__ index, const float* __restrict__ p1, int64_t N,
                                                        int64_t* idx64, int* idx32, float* dist2) {
  __shared__ KnnTree tree;
  if (threadIdx.x == 0) tree = load_tree(index);
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float qx = p1[i * 3], qy = p1[i * 3 + 1], qz = p1[i * 3 + 2];
  TopK<K> best;
  best.init();
  knn_search<K>(tree, qx, qy, qz, best);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool ok = best.id[k] != 0x7fffffff;  // fewer than K support points: pad with zeros like the reference
    if (idx64) idx64[i * K + k] = ok ? best.id[k] : 0;
    if (idx32) idx32[i * K + k] = ok ? best.id[k] : 0;
    if (dist2) dist2[i * K + k] = ok ? best.d[k] : 0.f;
  }
}

// Ray-sample query used by the render path: sample n = r*S + s sits at o_r + d_r * z_s, computed with the
// reference's operation order (one rounding per multiply/add, conditional_nerf/model.py:498).
// One thread walks SEG consecutive samples of a ray.  The K neighbours of sample s are real support points, so the largest
// of their distances to sample s+1 bounds the K-th distance there: every search after the first of a segment starts with a
// tight, exact pruning radius instead of +inf.  Adjacent lanes hold adjacent rays at the same depth (coherent traversal).
template <int K>
__global__ void __launch_bounds__(128) knn_query_rays_kernel(const void* __restrict__ index, const float* __restrict__ rays_o,
                                                             const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                                                             const float* __restrict__ sup_geo, int64_t R, int S, int SEG,
                                                             int64_t zs, int* idx32, float* dist2) {
  __shared__ KnnTree tree;
  if (threadIdx.x == 0) tree = load_tree(index);
  __syncthreads();
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nseg = (S + SEG - 1) / SEG;
  if (g >= R * nseg) return;
  const int64_t r = g % R;
  const int s0 = (int)(g / R) * SEG;
  const float ox = rays_o[r * 3 + 0], oy = rays_o[r * 3 + 1], oz = rays_o[r * 3 + 2];
  const float dx = rays_d[r * 3 + 0], dy = rays_d[r * 3 + 1], dz = rays_d[r * 3 + 2];
  TopK<K> best;
  bool have_prev = false;
  for (int s = s0; s < min(S, s0 + SEG); ++s) {
    const float z = z_vals[r * zs + s];
    const float qx = __fadd_rn(ox, __fmul_rn(dx, z));
    const float qy = __fadd_rn(oy, __fmul_rn(dy, z));
    const float qz = __fadd_rn(oz, __fmul_rn(dz, z));
    float bound = FLT_MAX;
    if (have_prev) {
      bound = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(sup_geo + (size_t)best.id[k] * 8));
        bound = fmaxf(bound, d2_exact(qx, qy, qz, p.x, p.y, p.z));
      }
    }
    best.init();
    knn_search<K>(tree, qx, qy, qz, best, bound);
    const int64_t i = r * S + s;
    have_prev = best.id[K - 1] != 0x7fffffff;
    int oi[K];
    float od[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool ok = best.id[k] != 0x7fffffff;
      oi[k] = ok ? best.id[k] : 0;
      od[k] = ok ? best.d[k] : 0.f;
    }
    if (K % 4 == 0) {   // (rows of K * 4 bytes: 16-byte aligned)
#pragma unroll
      for (int k = 0; k < K; k += 4) {
        __stcs(reinterpret_cast<int4*>(idx32 + i * K + k), make_int4(oi[k], oi[k + 1], oi[k + 2], oi[k + 3]));
        __stcs(reinterpret_cast<float4*>(dist2 + i * K + k), make_float4(od[k], od[k + 1], od[k + 2], od[k + 3]));
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) { __stcs(idx32 + i * K + k, oi[k]); __stcs(dist2 + i * K + k, od[k]); }
    }
  }
}


int knn_query(const void* index, const float* p1, int64_t N, int K, int64_t* idx64, int* idx32, float* dist2,
              cudaStream_t st) {
  if (N <= 0) return 0;
  const int T = 128;
  const unsigned G = (unsigned)((N + T - 1) / T);
  switch (K) {
    case 1: knn_query_kernel<1><<<G, T, 0, st>>>(index, p1, N, idx64, idx32, dist2); break;
    case 2: knn_query_kernel<2><<<G, T, 0, st>>>(index, p1, N, idx64, idx32, dist2); break;
    case 4: knn_query_kernel<4><<<G, T, 0, st>>>(index, p1, N, idx64, idx32, dist2); break;
    case 8: knn_query_kernel<8><<<G, T, 0, st>>>(index, p1, N, idx64, idx32, dist2); break;
    case 16: knn_query_kernel<16><<<G, T, 0, st>>>(index, p1, N, idx64, idx32, dist2); break;
    default: return set_error("knn_query: K must be one of 1,2,4,8,16");
  }
  return check_launch("knn_query");
}

int knn_query_rays(const void* index, const float* rays_o, const float* rays_d, const float* z_vals, int64_t zs,
                   const float* sup_geo, int64_t R, int S, int* idx32, float* dist2, cudaStream_t st) {
  if (R * S <= 0) return 0;
  const int T = 128;
  static const int seg_env = getenv("NLB_KNN_SEG") ? atoi(getenv("NLB_KNN_SEG")) : 0;   // A/B switch
  const int SEG = seg_env > 0 ? seg_env : (S >= 64 ? 16 : 8);
  const int64_t threads = R * ((S + SEG - 1) / SEG);
  knn_query_rays_kernel<8><<<(unsigned)((threads + T - 1) / T), T, 0, st>>>(index, rays_o, rays_d, z_vals, sup_geo, R, S, SEG,
                                                                           zs, idx32, dist2);
  return check_launch("knn_query_rays");
}

}  // namespace nlb
