// Shared device helpers for the NeRF-Loc B200 kernels (sm_100a).
//
// Everything on the render path is a chain of small dense layers whose activations never leave shared
// memory.  The workhorse is `tile_gemm`: C[rows x COLS] = A[rows x K] * Wt[K x COLS] where
//   * A lives in shared memory, row-major with a padded leading dimension (ld % 32 == 4 keeps the float4
//     fragment loads conflict free), optionally addressed through conv "taps" (row shift per K-block) so a
//     Conv1d / ConvTranspose1d along the ray is the same code path as a Linear layer;
//   * Wt is the pre-transposed, zero-padded weight in global memory (L2 resident, packed once per model by
//     pack.cu) and is streamed through a double-buffered cp.async staging ring;
//   * accumulation is fp32 FFMA with a TM x TN register micro-tile per thread (MMA mode "fp32-FFMA" in
//     the parity bar of 1e-4 rules out single-pass bf16/tf32 operands).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nlb {

constexpr int NT = 256;  // threads per CTA for all tile kernels
constexpr int KT = 16;   // K-tile of the weight staging ring
constexpr int NSTG = 4;  // ring depth: loads run NSTG-1 tiles ahead of the FFMA loop
constexpr int STAGE_FLOATS = NSTG * KT * 128;  // NSTG slots of [KT][<=128]

// Barrier over the NT compute threads of a CTA (named barrier 1).  Kernels that add a tcgen05 controller warp on top
// of the NT compute threads keep that warp out of these barriers; for plain NT-thread kernels it is a __syncthreads.
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Packed fp32 FMA (FFMA2, sm_100+): two independent IEEE fused multiply-adds in one issue slot, bit-identical to two fmaf.
// The FFMA GEMMs here are issue-bound (ncu: issue active 55 %, FMA pipe 31 % in aggregate_kernel), so halving the FMA
// instruction count is what matters.  A scalar operand is broadcast by the instruction itself (SASS `Rn.F32`).
__device__ __forceinline__ void fma2_s(float& c0, float& c1, const float a, const float b0, const float b1) {  // c += a * (b0, b1)
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %2};\nmov.b64 rb, {%3, %4};\nmov.b64 rc, {%0, %1};\n"
      "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}" : "+f"(c0), "+f"(c1) : "f"(a), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2_v(float& c0, float& c1, const float a0, const float a1, const float b0, const float b1) {  // c += (a0, a1) * (b0, b1)
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%0, %1};\n"
      "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}" : "+f"(c0), "+f"(c1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// packed fp32 pairs (one issue slot for two FMAs): operands are 64-bit registers holding two floats
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up2(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2 fma2p(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2p(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2p(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }


// LeakyReLU(0.01) as max(x, 0.01 x): the same value for every finite x (signed zeros included) in two instructions instead of three
__device__ __forceinline__ float leaky(float x) { return fmaxf(x, 0.01f * x); }
// ELU(alpha=1).  exp(x) - 1 with the hardware exponential: absolute error ~1e-7 for x < 0 (the library expm1f costs ~40
// instructions per element and was 15 % of aggregate_kernel's samples); far inside the 1e-4 parity bar.
// `ex2.approx.ftz` directly: __expf without -ftz compiles to 9 SASS instructions (denormal-range scaling), this form to 5, and
// the two agree bit for bit after the "- 1" (a flushed result is below 2^-126 next to -1).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : ex2_ftz(x * 1.4426950408889634f) - 1.f; }
__device__ __forceinline__ float softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }
// Hardware-exponential variants for the visibility decoder's output activations (64 rows x 8 transcendentals per tile sit on
// a short serial path): absolute error ~1e-7 on outputs of order 1, two orders below what the 1e-4 parity bar needs.
__device__ __forceinline__ float softplus_fast(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum over NT threads; `red` is >= 8 floats of shared scratch.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  cta_sync();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cta_sync();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += red[i];
  return t;
}

// A-operand addressing.  For K-tile kt (KT consecutive k) the fragment of logical row r starts at
//   base + (r + off[tap]) * ld + (kt*KT - tap*cin),  tap = (kt*KT) / cin.
// A plain Linear layer has cin >= K and off[0] = 0.  cin must be a multiple of KT.
struct ASrc {
  const float* base;
  int ld;
  int cin;
  int off0, off1, off2;
};
__device__ __forceinline__ ASrc plainA(const float* base, int ld) { return ASrc{base, ld, 1 << 30, 0, 0, 0}; }

// Core of the tile GEMM: acc[TM][TN] (+)= A[r0.., :] * Wt for the calling thread's micro-tile.
//   Wt: global, row k at Wt + k*ldb (column offset already applied by the caller), K % KT == 0.
//   sB: shared staging, 2*KT*COLS floats.
//   kscale (SCALE only): shared/global array of K factors; staged row k of Wt is multiplied by kscale[k] by the
//   thread that copied it, i.e. the GEMM computes A * (diag(kscale) Wt) - the S2D "fold a_n into layer 1" form.
template <int TM, int TN, int COLS, bool SCALE, int NS = NSTG>
__device__ __forceinline__ void gemm_core(const ASrc& A, const int r0, const bool active, const float* __restrict__ Wt,
                                          const int ldb, const int K, float* sB, const float* kscale,
                                          float (&acc)[TM][TN]) {
  static_assert(TN % 4 == 0 && COLS % TN == 0, "bad tile");
  constexpr int TC = COLS / TN;  // column groups
  constexpr int NG = TN / 4;     // float4 column groups per thread
  constexpr int GSTRIDE = COLS / NG;
  static_assert(NT % TC == 0, "bad tile");
  constexpr int CHUNKS = KT * COLS / 4;  // 16-byte chunks per stage
  constexpr int STG = KT * 128;          // floats per stage slot
  const int tid = threadIdx.x;
  const int tc = tid % TC;
  const int nkt = K / KT;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto stage_load = [&](int kt) {
    if (kt < nkt) {
      float* dst = sB + (kt % NS) * STG;
      const float* src = Wt + (size_t)kt * KT * ldb;
      for (int c = tid; c < CHUNKS; c += NT) {
        const int k = c / (COLS / 4), n4 = c % (COLS / 4);
        cp_async16(dst + k * COLS + n4 * 4, src + (size_t)k * ldb + n4 * 4);
      }
    }
    cp_async_commit();  // always commit: keeps the group count uniform
  };

  cta_sync();  // previous users of sB (and producers of A) are done
#pragma unroll
  for (int i = 0; i < NS - 1; ++i) stage_load(i);

  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<NS - 2>();  // tile kt has landed (this thread's part)
    if (SCALE) {
      float* cur = sB + (kt % NS) * STG;
      for (int c = tid; c < CHUNKS; c += NT) {
        const int k = c / (COLS / 4), n4 = c % (COLS / 4);
        const float sc = kscale[kt * KT + k];
        float4* p = reinterpret_cast<float4*>(cur + k * COLS + n4 * 4);
        float4 v = *p;
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        *p = v;
      }
    }
    cta_sync();                  // everybody's part of tile kt is visible; tile kt-1's slot is free
    stage_load(kt + NS - 1);   // refill the slot tile kt-1 used
    if (active) {
      const int k0 = kt * KT;
      const int tap = k0 / A.cin;
      const int roff = tap == 0 ? A.off0 : (tap == 1 ? A.off1 : A.off2);
      const float* arow = A.base + (r0 + roff) * A.ld + (k0 - tap * A.cin);
      const float* bt = sB + (kt % NS) * STG + tc * 4;
#pragma unroll
      for (int kk = 0; kk < KT; kk += 4) {
        float4 a[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * A.ld + kk);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          float b[TN];
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const float4 bv = *reinterpret_cast<const float4*>(bt + (kk + k4) * COLS + g * GSTRIDE);
            b[g * 4 + 0] = bv.x; b[g * 4 + 1] = bv.y; b[g * 4 + 2] = bv.z; b[g * 4 + 3] = bv.w;
          }
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            const float av = k4 == 0 ? a[i].x : (k4 == 1 ? a[i].y : (k4 == 2 ? a[i].z : a[i].w));
#pragma unroll
            for (int j = 0; j < TN; j += 2) fma2_s(acc[i][j], acc[i][j + 1], av, b[j], b[j + 1]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
}

// C = A * Wt (+ epilogue).  rows: runtime; threads whose rows fall outside are idle but still help staging.
//   epi(row, col, value) is called once per output element by its owning thread.
// If SYNC_BEFORE_EPI, a __syncthreads() separates the last A read from the first epilogue write (in-place use).
template <int TM, int TN, int COLS, bool SYNC_BEFORE_EPI, class Epi>
__device__ __forceinline__ void tile_gemm(const ASrc A, const int rows, const float* __restrict__ Wt, const int ldb,
                                          const int K, float* sB, Epi epi) {
  constexpr int TC = COLS / TN;
  constexpr int TR = NT / TC;  // row groups per pass
  constexpr int NG = TN / 4;
  constexpr int GSTRIDE = COLS / NG;
  const int tc = threadIdx.x % TC, tr = threadIdx.x / TC;
  for (int rb = 0; rb < rows; rb += TR * TM) {
    const int r0 = rb + tr * TM;
    const bool active = r0 < rows;
    float acc[TM][TN];
    gemm_core<TM, TN, COLS, false>(A, r0, active, Wt, ldb, K, sB, nullptr, acc);
    if (SYNC_BEFORE_EPI) cta_sync();
    if (active) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int j = 0; j < 4; ++j) epi(r0 + i, g * GSTRIDE + tc * 4 + j, acc[i][g * 4 + j]);
    }
  }
}

// Variant that leaves the accumulators with the caller (single pass: rows <= TR*TM).  Used where a LayerNorm
// over the whole output slab, or a reduction over the columns, has to happen before anything is written.
template <int TM, int TN, int COLS>
struct Frag {
  float acc[TM][TN];
  int r0;
  bool active;
  static constexpr int TC = COLS / TN;
  static constexpr int NG = TN / 4;
  static constexpr int GSTRIDE = COLS / NG;
  __device__ __forceinline__ int col(int j) const { return (j / 4) * GSTRIDE + (threadIdx.x % TC) * 4 + (j % 4); }
};

template <int TM, int TN, int COLS, bool SCALE = false, int NS = NSTG>
__device__ __forceinline__ void tile_gemm_frag(const ASrc A, const int rows, const float* __restrict__ Wt, const int ldb,
                                               const int K, float* sB, Frag<TM, TN, COLS>& f,
                                               const float* kscale = nullptr) {
  constexpr int TC = COLS / TN;
  f.r0 = (threadIdx.x / TC) * TM;
  f.active = f.r0 < rows;
  gemm_core<TM, TN, COLS, SCALE, NS>(A, f.r0, f.active, Wt, ldb, K, sB, kscale, f.acc);
}

// Small-M GEMM: out[r][c] = sum_k A_r[k] * Wt[k*ldb + c] for r < R (<= 16), c < COLS.  A tile GEMM would give every thread one
// row and re-read the staged weight tile R times; here a thread owns two adjacent output COLUMNS for all R rows (R FFMA2
// accumulator pairs), reads its weight pair straight from L2 (a warp reads 256 contiguous bytes per k, no staging, no
// barrier per K-tile) and the 256 threads split K in 2*NT/COLS slices that are reduced through `red`
// (>= (2*NT/COLS)*R*COLS floats of shared scratch).  Many K slices keep the number of dependent L2 round trips per thread
// small - the loop is latency bound.  arow(r, c) -> shared-memory pointer to row r as seen by column c (c even).
struct NoMid { __device__ __forceinline__ void operator()() const {} };
// `mid()` runs right after the FMA loop, when the weight registers are dead (see rows16_load / rows16_compute below).
template <int COLS, int R, int K, int KB, class ARow, class Epi, class Mid = NoMid>
__device__ __forceinline__ void rows16_gemm(ARow arow, const float* __restrict__ Wt, const int ldb, float* red, Epi epi, Mid mid = Mid()) {
  constexpr int KSPLIT = 2 * NT / COLS;
  constexpr int KPER = K / KSPLIT;                 // k per thread, multiple of 4
  constexpr int NBATCH = (KPER + KB - 1) / KB;     // weight batches: batch n+1 is requested before the FMAs of batch n
  static_assert(K % KSPLIT == 0 && KPER % 4 == 0 && KB % 4 == 0, "rows16_gemm: bad K split");
  const int tid = threadIdx.x;
  const int c = (tid % (COLS / 2)) * 2, ks = tid / (COLS / 2);
  const int k0 = ks * KPER;
  float acc[R][2];
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
  const float* wp = Wt + (size_t)k0 * ldb + c;
  float2 b[KB], bn[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) b[j] = (j < KPER) ? __ldg(reinterpret_cast<const float2*>(wp + (size_t)j * ldb)) : make_float2(0.f, 0.f);
#pragma unroll
  for (int nb = 0; nb < NBATCH; ++nb) {
    const int k = nb * KB;
    if (nb + 1 < NBATCH) {
#pragma unroll
      for (int j = 0; j < KB; ++j)
        bn[j] = (k + KB + j < KPER) ? __ldg(reinterpret_cast<const float2*>(wp + (size_t)(k + KB + j) * ldb)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float* ap = arow(r, c) + k0 + k;
#pragma unroll
      for (int g = 0; g < KB / 4; ++g) {
        if (k + 4 * g < KPER) {
          const float4 a = *reinterpret_cast<const float4*>(ap + 4 * g);
          fma2_s(acc[r][0], acc[r][1], a.x, b[4 * g].x, b[4 * g].y);
          fma2_s(acc[r][0], acc[r][1], a.y, b[4 * g + 1].x, b[4 * g + 1].y);
          fma2_s(acc[r][0], acc[r][1], a.z, b[4 * g + 2].x, b[4 * g + 2].y);
          fma2_s(acc[r][0], acc[r][1], a.w, b[4 * g + 3].x, b[4 * g + 3].y);
        }
      }
    }
    if (nb + 1 < NBATCH) {
#pragma unroll
      for (int j = 0; j < KB; ++j) b[j] = bn[j];
    }
  }
  mid();
  cta_sync();  // `red` may alias a buffer an earlier phase still reads
#pragma unroll
  for (int r = 0; r < R; ++r) *reinterpret_cast<float2*>(red + (ks * R + r) * COLS + c) = make_float2(acc[r][0], acc[r][1]);
  cta_sync();
  for (int i = tid; i < R * COLS; i += NT) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < KSPLIT; ++q) v += red[q * R * COLS + i];
    epi(i / COLS, i % COLS, v);
  }
}

// ---- warp-level tensor-core path (mma.sync.m16n8k8 tf32, 3xTF32 split): the in-kernel visibility decoder of aggregate_kernel
// (V > 8 and the scratch-less nlb_aggregate_points entry point).  3xTF32: hi = value with the low 13 mantissa bits cleared,
// lo = value - hi; lo*hi + hi*lo + hi*hi.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
__device__ __forceinline__ void split_hi_lo(const float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
}  // namespace nlb
