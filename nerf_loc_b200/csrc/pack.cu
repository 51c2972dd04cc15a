// Weight packing for the render path: reference-layout tensors (the `state_dict` of ConditionalNeRF,
// nerf_loc/models/conditional_nerf/model.py:29-135) -> one flat device buffer of transposed, zero-padded
// matrices in the layout the tile GEMM streams (see nlb_internal.h::RenderW).  Runs once per model load.
#include "nlb_internal.h"

namespace nlb {

// index of every tensor in the order of nerf_loc_b200/params.py::conditional_nerf_shapes
enum P {
  RD0_W = 0, RD0_B, RD2_W, RD2_B,
  DEC = 4,  // 4 heads x (0.w,0.b,2.w,2.b,4.w,4.b): mean, var, aw, vis
  FC0_W = 28, FC0_B, FC2_W, FC2_B,
  CF0_W = 32, CF0_B, CF2_W, CF2_B,
  KP_W = 36, KP_B,
  BM0_W = 38, BM0_B, BM2_W, BM2_B, BM4_W, BM4_B,
  AT_Q = 44, AT_K, AT_V, AT_FC, AT_LNG, AT_LNB,
  AW0_W = 50, AW0_B, AW2_W, AW2_B,
  UNET = 54,  // 7 layers x (w, b, g, be)
  SIG_W = 82, SIG_B,
  FT0_W = 84, FT0_B, FT2_W, FT2_B,
  BL0_W = 88, BL0_B, BL2_W, BL2_B, BL4_W, BL4_B,
  BETA_W = 94, BETA_B,
  PJC_W = 96, PJC_B, PJF_W, PJF_B,
  N_PARAMS = 100
};

struct Alloc {
  const float* base;
  size_t off = 0;
  const float* take(size_t n) {
    const float* p = base + off;
    off += (n + 63) / 64 * 64;
    return p;
  }
};

static const int UN_CIN[7] = {128, 64, 128, 128, 256, 128, 160};
static const int UN_COUT[7] = {64, 128, 128, 128, 64, 32, 128};
static const int UN_SDIV[7] = {1, 2, 4, 4, 2, 1, 1};
static const bool UN_TR[7] = {false, false, false, true, true, true, false};

static RenderW layout(const float* base, int S, size_t* total) {
  Alloc a{base};
  RenderW w{};
  w.S = S;
  w.dec1 = a.take(32 * 128); w.dec1_b = a.take(128);
  w.dec2 = a.take(4 * 32 * 32); w.dec2_b = a.take(128);
  w.dec3 = a.take(6 * 32); w.dec3_b = a.take(6);
  w.fc1 = a.take(416 * 64); w.fc1_b = a.take(64);
  w.fc2 = a.take(64 * 128); w.fc2_b = a.take(128);
  w.bl1v = a.take(224 * 32); w.bl1_b = a.take(32);
  w.bl1a = a.take(128 * 32);
  w.bl2 = a.take(16 * 32); w.bl2_b = a.take(16);
  w.bl3 = a.take(16); w.bl3_b = a.take(1);
  w.w1a = a.take(224 * 128); w.b1 = a.take(128);
  w.rd1 = a.take(64); w.rd1_b = a.take(16);
  w.rd2 = a.take(27 * 16); w.rd2_b = a.take(27);
  w.b2 = a.take(128);
  w.b3 = a.take(128);
  w.ln_g = a.take(128); w.ln_b = a.take(128);
  for (int l = 0; l < 7; ++l) {
    const int sl = S > 0 ? S / UN_SDIV[l] : 0;
    w.u[l].w = a.take((size_t)3 * UN_CIN[l] * UN_COUT[l]);
    w.u[l].b = a.take(UN_COUT[l]);
    w.u[l].g = a.take((size_t)sl * UN_COUT[l]);
    w.u[l].be = a.take((size_t)sl * UN_COUT[l]);
    w.u[l].g2 = a.take((size_t)((sl + 31) / 32 * 32) * UN_COUT[l]);
    w.u[l].be2 = a.take((size_t)((sl + 31) / 32 * 32) * UN_COUT[l]);
  }
  // bf16 hi | lo: 2 planes x 2 bytes = one float per weight
  for (int l = 0; l < 7; ++l)
    for (int t = 0; t < 3; ++t) w.tb_u[l][t] = a.take((size_t)UN_COUT[l] * UN_CIN[l]);
  w.tb_bl1a = a.take(32 * 128);
  w.tb_ft1 = a.take(128 * 128);
  w.tb_w1b = a.take(128 * 96);
  w.tb_w2 = a.take(128 * 128); w.tb_w3 = a.take(128 * 128);
  w.tb_dec1 = a.take(128 * 32); w.tb_dec2 = a.take(4 * 32 * 32);
  w.tb_fc1 = a.take(64 * 416); w.tb_fc2 = a.take(128 * 64);
  w.tb_wq = a.take(128 * 128); w.tb_wk = a.take(128 * 128); w.tb_wv = a.take(128 * 128); w.tb_wfc = a.take(128 * 128);
  w.sig_w = a.take(128); w.sig_b = a.take(1);
  w.ft1 = a.take(128 * 128); w.ft1_b = a.take(128);
  w.ft2 = a.take(128 * 192); w.ft2_b = a.take(192);
  w.cf1 = a.take(128 * 64); w.cf1_b = a.take(64); w.cf2 = a.take(64); w.cf2_b = a.take(1);
  w.pj_c = a.take(352 * 192); w.pj_c_b = a.take(192);
  w.pj_f = a.take(352 * 192); w.pj_f_b = a.take(192);
  if (total) *total = a.off;
  return w;
}

size_t render_weights_floats(int S) {
  size_t t = 0;
  layout(nullptr, S, &t);
  return t;
}

RenderW render_weights_view(const float* packed, int S) { return layout(packed, S, nullptr); }

// dst[k*dst_ld + n] = (k < Kv) ? src[n*src_ld + src_off + k] : 0   for k < Kp, n < N
__global__ void pack_t_kernel(float* dst, const float* __restrict__ src, int Kp, int N, int dst_ld, int src_ld,
                              int src_off, int Kv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kp * N) return;
  const int k = i / N, n = i % N;
  dst[(size_t)k * dst_ld + n] = k < Kv ? src[(size_t)n * src_ld + src_off + k] : 0.f;
}
// LayerNorm affine [C][S] (reference layout) -> ln_off layout over [S][C]
__global__ void pack_ln_kernel(float* dst, const float* __restrict__ src, int S, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * C) return;
  const int s = i / C, c = i % C;
  dst[ln_off(s, c, C)] = src[(size_t)c * S + s];
}
__global__ void pack_copy_kernel(float* dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// conv taps: dst[(j*Cin+ci)*Cout + co] = transposed ? src[(ci*Cout+co)*3 + tap_j] : src[(co*Cin+ci)*3 + tap_j]
__global__ void pack_conv_kernel(float* dst, const float* __restrict__ src, int Cin, int Cout, int ntaps, int t0, int t1,
                                 int t2, int transposed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntaps * Cin * Cout) return;
  const int co = i % Cout, ci = (i / Cout) % Cin, j = i / (Cout * Cin);
  const int t = j == 0 ? t0 : (j == 1 ? t1 : t2);
  dst[i] = transposed ? src[((size_t)ci * Cout + co) * 3 + t] : src[((size_t)co * Cin + ci) * 3 + t];
}

// K order of the neighbour MLP's per-pair layer-1 operand as neighbor2_kernel writes it (perm == 1; two threads per row, each
// storing 48 consecutive columns): [offset xyz, 0 | PE octaves 0-4 | ray_diff_fc 0-13] [PE octaves 5-9 | ray_diff_fc 14-26 | 0 x 5];
// the source order is [offset xyz | PE octaves 0-9 | ray_diff_fc 0-26].
__device__ __forceinline__ int tcb_src_index(int k, int perm) {
  if (perm == 0) return k;
  if (perm == 2) {
    // out_fc input as aggregate_kernel (GOUT) writes it: map-channel means, map-channel variances, rgb mean, rgb variance,
    // extras; the source order is [rgb + map means (195) | rgb + map variances (195) | extras (3)]
    if (k < 192) return 3 + k;
    if (k < 384) return 195 + 3 + (k - 192);
    if (k < 387) return k - 384;
    if (k < 390) return 195 + (k - 387);
    if (k < 393) return k;
    return -1;
  }
  if (k < 3) return k;
  if (k == 3) return -1;
  if (k < 34) return k - 1;
  if (k < 48) return k + 29;
  if (k < 78) return k - 15;
  if (k < 91) return k - 1;
  return -1;
}
// bf16x3 B operand (tc_bf16.cuh): W [N][K] -> per K-tile of `ktile` columns: hi tile then lo tile, each in the weight-tile
// layout (8-row x 16-byte core matrices, adjacent in K contiguous, 8-row groups ktile*16 bytes apart); hi = bf16(x), lo = bf16(x - hi)
// (n_off, N_total): the N rows packed by this call are rows n_off.. of a tile with N_total rows (several sources side by side)
__global__ void pack_tcb16_kernel(uint16_t* dst, const float* __restrict__ src, int N, int K, int src_ld, int src_off, int src_ks,
                                  int ktile, int Kv, int perm, int n_off, int N_total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n0 = i / K, k = i % K, n = n0 + n_off;
  const int ksrc = tcb_src_index(k, perm);
  const float x = (ksrc >= 0 && ksrc < Kv) ? src[(size_t)n0 * src_ld + src_off + (size_t)ksrc * src_ks] : 0.f;
  const uint32_t xb = __float_as_uint(x);
  // round to nearest even by hand (identical to __float2bfloat16_rn for finite values)
  const uint32_t hb = (xb + 0x7FFFu + ((xb >> 16) & 1u)) >> 16;
  const float r = x - __uint_as_float(hb << 16);
  const uint32_t rb = __float_as_uint(r);
  const uint32_t lb = (rb + 0x7FFFu + ((rb >> 16) & 1u)) >> 16;
  const int kt = k / ktile, kl = k % ktile;
  const size_t base = (size_t)kt * (2 * N_total * ktile);
  const size_t off = (size_t)(n / 8) * (ktile * 8) + (size_t)(kl / 8) * 64 + (n % 8) * 8 + (kl % 8);
  dst[base + off] = (uint16_t)hb;
  dst[base + (size_t)N_total * ktile + off] = (uint16_t)lb;
}

namespace {
struct Packer {
  const float* const* p;
  cudaStream_t st;
  void t(const float* dst, int src, int Kp, int N, int src_ld, int src_off, int Kv, int dst_ld = -1) {
    const int n = Kp * N;
    pack_t_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(dst), p[src], Kp, N, dst_ld < 0 ? N : dst_ld,
                                                  src_ld, src_off, Kv);
  }
  void c(const float* dst, int src, int n, int dst_off = 0) {
    pack_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(dst) + dst_off, p[src], n);
  }
  void tcb16(const float* dst, int src, int N, int K, int src_ld, int src_off, int src_ks, int ktile, int Kv = -1, int perm = 0,
             int n_off = 0, int N_total = -1) {
    const int n = N * K;
    pack_tcb16_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<uint16_t*>(const_cast<float*>(dst)), p[src], N, K, src_ld,
                                                       src_off, src_ks, ktile, Kv < 0 ? K : Kv, perm, n_off, N_total < 0 ? N : N_total);
  }
  void conv(const float* dst, int src, int Cin, int Cout, int ntaps, int t0, int t1, int t2, bool tr) {
    const int n = ntaps * Cin * Cout;
    pack_conv_kernel<<<(n + 255) / 256, 256, 0, st>>>(const_cast<float*>(dst), p[src], Cin, Cout, ntaps, t0, t1, t2, tr);
  }
};
}  // namespace

int render_weights_pack(const float* const* params, int n_params, int S, float* packed, size_t packed_floats,
                        cudaStream_t st) {
  if (n_params != N_PARAMS) return set_error("render_weights_pack: expected 100 parameter tensors (params.py order)");
  if (S < 0 || S % 8 != 0) return set_error("render_weights_pack: S must be a multiple of 8 (RayUnet pools three times)");
  if (packed_floats < render_weights_floats(S)) return set_error("render_weights_pack: packed buffer too small");
  for (int i = 0; i < n_params; ++i)
    if (!params[i]) return set_error("render_weights_pack: null parameter pointer");
  cudaError_t e = cudaMemsetAsync(packed, 0, render_weights_floats(S) * sizeof(float), st);
  if (e != cudaSuccess) return set_error(cudaGetErrorString(e));
  const RenderW w = render_weights_view(packed, S);
  Packer k{params, st};
  // --- visibility decoder: heads in the order mean, var, aw, vis ---
  for (int h = 0; h < 4; ++h) {
    const int b = DEC + 6 * h;
    k.t(w.dec1 + 32 * h, b + 0, 32, 32, 32, 0, 32, 128);   // dst[k][32h+n] = W0[n][k]
    k.c(w.dec1_b, b + 1, 32, 32 * h);
    k.t(w.dec2 + 32 * 32 * h, b + 2, 32, 32, 32, 0, 32);   // [k][n] = W2[n][k]
    k.c(w.dec2_b, b + 3, 32, 32 * h);
  }
  // dec3 rows: mean0, mean1, var0, var1, aw, vis (natural [out][32])
  k.c(w.dec3, DEC + 4, 64, 0);       k.c(w.dec3_b, DEC + 5, 2, 0);
  k.c(w.dec3, DEC + 6 + 4, 64, 64);  k.c(w.dec3_b, DEC + 6 + 5, 2, 2);
  k.c(w.dec3, DEC + 12 + 4, 32, 128); k.c(w.dec3_b, DEC + 12 + 5, 1, 4);
  k.c(w.dec3, DEC + 18 + 4, 32, 160); k.c(w.dec3_b, DEC + 18 + 5, 1, 5);
  k.t(w.fc1, FC0_W, 416, 64, 393, 0, 393);  k.c(w.fc1_b, FC0_B, 64);
  k.t(w.fc2, FC2_W, 64, 128, 64, 0, 64);    k.c(w.fc2_b, FC2_B, 128);
  // --- colour blend: input = [feature_agg 128 | rgb_feat 195 | vis 1 | ray_diff 4] ---
  k.t(w.bl1v, BL0_W, 224, 32, 328, 128, 200);  k.c(w.bl1_b, BL0_B, 32);
  k.t(w.bl1a, BL0_W, 128, 32, 328, 0, 128);
  k.c(w.bl2, BL2_W, 16 * 32);  k.c(w.bl2_b, BL2_B, 16);
  k.c(w.bl3, BL4_W, 16);       k.c(w.bl3_b, BL4_B, 1);
  // --- neighbour MLP: input = [support feature 195 | PE 63 | ray_diff_fc 27] ---
  k.t(w.w1a, BM0_W, 224, 128, 285, 0, 195);  k.c(w.b1, BM0_B, 128);
  k.c(w.rd1, RD0_W, 64);  k.c(w.rd1_b, RD0_B, 16);
  k.c(w.rd2, RD2_W, 27 * 16);  k.c(w.rd2_b, RD2_B, 27);
  k.c(w.b2, BM2_B, 128);
  k.c(w.b3, BM4_B, 128);
  k.c(w.ln_g, AT_LNG, 128);  k.c(w.ln_b, AT_LNB, 128);
  // --- RayUnet ---
  if (S > 0) {
    for (int l = 0; l < 7; ++l) {
      const int b = UNET + 4 * l, ci = UN_CIN[l], co = UN_COUT[l], sl = S / UN_SDIV[l];
      if (!UN_TR[l]) {
        k.conv(w.u[l].w, b, ci, co, 3, 0, 1, 2, false);
      } else {
        k.conv(w.u[l].w, b, ci, co, 1, 1, 0, 0, true);                       // even outputs: tap 1
        k.conv(w.u[l].w + (size_t)ci * co, b, ci, co, 2, 2, 0, 0, true);     // odd outputs: tap 2 (row j), tap 0 (row j+1)
      }
      k.c(w.u[l].b, b + 1, co);
      k.t(w.u[l].g, b + 2, sl, co, sl, 0, sl);   // [C][S] -> [S][C]
      k.t(w.u[l].be, b + 3, sl, co, sl, 0, sl);
      pack_ln_kernel<<<(sl * co + 255) / 256, 256, 0, st>>>(const_cast<float*>(w.u[l].g2), params[b + 2], sl, co);
      pack_ln_kernel<<<(sl * co + 255) / 256, 256, 0, st>>>(const_cast<float*>(w.u[l].be2), params[b + 3], sl, co);
    }
  }
  if (S > 0) {
    // K-tile extents as render_ray2.cu's GEMM list uses them (a tile is at most 16 KB and never straddles two source tiles)
    static const int KT16[7] = {64, 32, 32, 32, 64, 64, 32};
    for (int l = 0; l < 7; ++l) {
      const int b = UNET + 4 * l, ci = UN_CIN[l], co = UN_COUT[l];
      for (int t = 0; t < 3; ++t) {
        if (UN_TR[l]) k.tcb16(w.tb_u[l][t], b, co, ci, 3, t, co * 3, KT16[l]);    // ConvTranspose1d weight [ci][co][3]
        else k.tcb16(w.tb_u[l][t], b, co, ci, ci * 3, t, 3, KT16[l]);             // Conv1d weight [co][ci][3]
      }
    }
  }
  for (int h = 0; h < 4; ++h) {
    k.tcb16(w.tb_dec1, DEC + 6 * h + 0, 32, 32, 32, 0, 1, 32, -1, 0, 32 * h, 128);   // rows 32 h .. of the [128 x 32] tile
    k.tcb16(w.tb_dec2 + 32 * 32 * h, DEC + 6 * h + 2, 32, 32, 32, 0, 1, 32);
  }
  k.tcb16(w.tb_fc1, FC0_W, 64, 416, 393, 0, 1, 32, 393, 2);
  k.tcb16(w.tb_fc2, FC2_W, 128, 64, 64, 0, 1, 32);
  k.tcb16(w.tb_w1b, BM0_W, 128, 96, 285, 195, 1, 32, 90, 1);
  k.tcb16(w.tb_w2, BM2_W, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_w3, BM4_W, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_wq, AT_Q, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_wk, AT_K, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_wv, AT_V, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_wfc, AT_FC, 128, 128, 128, 0, 1, 32);
  k.tcb16(w.tb_bl1a, BL0_W, 32, 128, 328, 0, 1, 128);
  k.tcb16(w.tb_ft1, FT0_W, 128, 128, 128, 0, 1, 32);
  k.c(w.sig_w, SIG_W, 128);  k.c(w.sig_b, SIG_B, 1);
  k.t(w.ft1, FT0_W, 128, 128, 128, 0, 128);  k.c(w.ft1_b, FT0_B, 128);
  k.t(w.ft2, FT2_W, 128, 192, 128, 0, 128);  k.c(w.ft2_b, FT2_B, 192);
  k.t(w.cf1, CF0_W, 128, 64, 128, 0, 128);   k.c(w.cf1_b, CF0_B, 64);
  k.c(w.cf2, CF2_W, 64);  k.c(w.cf2_b, CF2_B, 1);
  k.t(w.pj_c, PJC_W, 352, 192, 323, 0, 323);  k.c(w.pj_c_b, PJC_B, 192);
  k.t(w.pj_f, PJF_W, 352, 192, 323, 0, 323);  k.c(w.pj_f_b, PJF_B, 192);
  return check_launch("render_weights_pack");
}

}  // namespace nlb
