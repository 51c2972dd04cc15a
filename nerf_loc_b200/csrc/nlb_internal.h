// Internal (C++) declarations shared by the translation units of libnerfloc_b200.so.
// The public surface is include/nerfloc_b200.h; nothing here is exported.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace nlb {

int set_error(const char* msg);           // records the message, returns a non-zero code
int check_launch(const char* what);       // cudaGetLastError() -> set_error
void prof_mark(const char* name);         // per-kernel timing of nlb_render_rays (no-op unless nlb_profile_enable(1))

// ---- exact KNN (knn.cu) ------------------------------------------------------------------------------------
size_t knn_index_bytes(int64_t M);
int knn_build(const float* xyz, int64_t M, void* buf, size_t bytes, cudaStream_t st);
int knn_query(const void* index, const float* p1, int64_t N, int K, int64_t* idx64, int* idx32, float* dist2,
              cudaStream_t st);
int knn_query_rays(const void* index, const float* rays_o, const float* rays_d, const float* z_vals, int64_t zs,
                   const float* sup_geo, int64_t R, int S, int* idx32, float* dist2, cudaStream_t st);

// ---- packed weights of the render path (pack.cu) ------------------------------------------------------------
// All matrices are stored transposed ("Wt": row = input feature k, column = output feature n), zero padded so
// that K is a multiple of 32, as the tile GEMM wants them.  Offsets are in floats from the buffer start.
constexpr int C_FEAT = 192;   // backbone2d_fpn_dim
constexpr int C_RGBF = 195;   // rgb + feature
constexpr int W_HID = 128;    // model_3d_hidden_dim
constexpr int C_VIS = 32;     // DepthFusionNet output channels
constexpr int KNN_K = 8;

// Layout of the per-(sample, view) blend partials [rows][32] between aggregate_kernel and the ray kernels: groups of 32 rows,
// the eight 16-byte pieces of a row 512 bytes apart, so that 32 consecutive rows read (or written) one piece at a time by 32
// lanes are one contiguous 512-byte run.  (Row-major, the thread-per-row loads of the ray kernel were 32 separate 128-byte
// lines per instruction and stalled its load issue for 10 k clk per batch.)  Float offset of 4-float piece q of row r:
__host__ __device__ __forceinline__ int64_t partial_off(int64_t r, int q) { return (r >> 5) * 1024 + (int64_t)q * 128 + (r & 31) * 4; }

struct UnetLayer {
  const float* w;      // conv: [3*Cin][Cout]; transposed conv: even [Cin][Cout] followed by odd [2*Cin][Cout]
  const float* b;      // [Cout]
  const float* g;      // LayerNorm gain, transposed to [S_level][Cout]
  const float* be;     // LayerNorm bias,  transposed to [S_level][Cout]
  const float* g2;     // the same two in the piece-major layout of ln_off (pair ray kernel: thread-per-row reads, coalesced)
  const float* be2;
};

// [N][128] fp32 matrices that are written and read by thread-per-row epilogues (aggregated, the attention query): the same
// piece-major layout with 32 pieces per row - groups of 32 rows (4096 floats), piece q of row r at q * 128 + (r & 31) * 4.  A
// row-major 16-byte access per lane touches 32 different 128-byte lines per instruction; here it is one 512-byte run.  Buffers
// hold whole groups (N rounded up to 32 rows).
__host__ __device__ __forceinline__ int64_t pm128_off(int64_t r, int c) { return (r >> 5) * 4096 + (int64_t)(c >> 2) * 128 + (r & 31) * 4 + (c & 3); }
__host__ __device__ __forceinline__ size_t pm128_floats(int64_t N) { return (size_t)((N + 31) / 32) * 4096; }

// LayerNorm affine [rows][C] for thread-per-row readers: groups of 32 rows, the 4-float pieces of a row 512 bytes apart (see
// partial_off).  Float offset of element (r, c):
__host__ __device__ __forceinline__ int64_t ln_off(int r, int c, int C) { return (int64_t)(r >> 5) * (32 * C) + (c >> 2) * 128 + (r & 31) * 4 + (c & 3); }

struct RenderW {
  int S;  // samples per ray the RayUnet LayerNorms were built for (0: no RayUnet packed)
  // aggregator (multiview_aggregator.py, visibility_decoder.py)
  const float *dec1, *dec1_b;   // [32][128] (heads mean|var|aw|vis), [128]
  const float *dec2, *dec2_b;   // 4 x [32][32], [128]
  const float *dec3, *dec3_b;   // [6][32] natural (mean0,mean1,var0,var1,aw,vis), [6]
  const float *fc1, *fc1_b;     // [416][64], [64]
  const float *fc2, *fc2_b;     // [64][128], [128]
  // colour blend (model.py:90-96)
  const float *bl1v, *bl1_b;    // per-view part [224][32]: rows 0..194 rgb_feat, 195 vis, 196..199 ray_diff; bias [32]
  const float *bl1a;            // feature_agg part [128][32]
  const float *bl2, *bl2_b;     // [16][32] natural, [16]
  const float *bl3, *bl3_b;     // [16], [1]
  // neighbour MLP (model.py:36-39,63-77)
  const float *w1a, *b1;        // [224][128] support-feature part of base_mlp.0 (+bias) -> per-frame precompute
  const float *rd1, *rd1_b;     // [16][4] natural, [16]
  const float *rd2, *rd2_b;     // [27][16] natural, [27]
  const float *b2, *b3;         // [128] biases of base_mlp.2 / .4 (the matrices only exist as tensor-core tiles, below)
  const float *ln_g, *ln_b;     // [128]
  // ray stage
  UnetLayer u[7];               // conv1, conv2, conv3, trans_conv3, trans_conv2, trans_conv1, conv_out
  // bf16x3 copies for the pair kernel (render_ray2.cu): per layer and TAP (0, 1, 2) the [Cout x Cin] operand as K-tiles of
  // [hi | lo] weight tiles (tc_bf16.cuh); conv_out keeps its 160 input channels in one block (K-tiles 0-3: x, 4: x2)
  const float* tb_u[7][3];
  const float *tb_bl1a, *tb_ft1;
  // bf16x3 copies for the neighbour kernels (neighbor2.cu): [N = 128][K] operands as K-tiles of 32 ([hi | lo], 16 KB each);
  // tb_w1b in the K order neighbor2_kernel writes its layer-1 operand (pack.cu::tcb_src_index)
  const float *tb_w1b, *tb_w2, *tb_w3, *tb_wq, *tb_wk, *tb_wv, *tb_wfc;
  // visibility decoder (visibility.cu): layer 1 of the four heads as one [128 x 32] tile, layer 2 as four [32 x 32] tiles
  const float *tb_dec1, *tb_dec2;
  // out_fc (fc_tail.cu): layer 1 [64 x 416] in the column order of aggregate_kernel's statistics vector, layer 2 [128 x 64]
  const float *tb_fc1, *tb_fc2;
  const float *sig_w, *sig_b;   // [128], [1]
  const float *ft1, *ft1_b;     // [128][128], [128]
  const float *ft2, *ft2_b;     // [128][192], [192]
  // per-frame setup / descriptor heads
  const float *cf1, *cf1_b, *cf2, *cf2_b;   // confidence_mlp: [128][64], [64], [64], [1]
  const float *pj_c, *pj_c_b, *pj_f, *pj_f_b;  // proj_layer_3d_*: [352][192] (rows 0..127 feature_agg, 128..322 feature), [192]
};

size_t render_weights_floats(int S);
// params: host array of device pointers in the order of nerf_loc_b200/params.py::conditional_nerf_shapes(S)
int render_weights_pack(const float* const* params, int n_params, int S, float* packed, size_t packed_floats,
                        cudaStream_t st);
RenderW render_weights_view(const float* packed, int S);

// ---- scene (per-frame) -----------------------------------------------------------------------------------------
struct SceneDev {
  int V, H, W, h, w, vh, vw;
  const float* images;   // [V][H][W][4]  rgb + pad
  const float* feat;     // [V][h][w][192]
  const float* vis;      // [V][vh][vw][32]
  const float* cams;     // [V][32]: P=K_hom*w2c rows 0..2 (12) | K*Rt (12) | camera centre (3) | pad
  float near_, far_;
  int64_t M;             // support points
  const float* sup_pre;  // [M][128]
  const float* sup_geo;  // [M][8]: xyz, dir, conf, pad
  const void* knn;       // KNN index
  float qc[3];           // query camera centre (colour-blend ray_diff); unused by query()
  const float* featb;    // [V][h][w][32] feature maps pre-projected through rgb_blending_mlp.0 (render only; null for query)
};

}  // namespace nlb
