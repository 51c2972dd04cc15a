// fp32 FMA pipe peak on this GPU: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 16 independent accumulator chains per thread.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp32_peak tools/fp32_peak.cu ; prints MAC/clk/SM and TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float& c0, float& c1, float a, float b0, float b1) {
  asm volatile("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %2};\nmov.b64 rb, {%3, %4};\nmov.b64 rc, {%0, %1};\n"
               "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}" : "+f"(c0), "+f"(c1) : "f"(a), "f"(b0), "f"(b1));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* __restrict__ in, int iters) {
  // operands in registers, as in a register-tiled GEMM (no immediate / constant-bank forms)
  const float a = in[threadIdx.x], b0 = in[256 + threadIdx.x], b1 = in[512 + threadIdx.x];
  float c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(a, (i & 1) ? b1 : b0, c[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) fma2(c[i], c[i + 1], a, b0, b1);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float *out, *in;
  const int blocks = sms * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaMalloc(&in, 768 * sizeof(float));
  cudaMemset(in, 0, 768 * sizeof(float));
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, 256>>>(out, in, iters);
      else k<1><<<blocks, 256>>>(out, in, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double mac = (double)blocks * 256 * iters * 8 * 16;
      if (rep == 1)
        printf("{\"mode\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f, \"mac_per_clk_per_sm_at_max_clock\": %.1f, \"sms\": %d, \"max_khz\": %d}\n",
               mode ? "FFMA2" : "FFMA", ms, 2 * mac / (ms * 1e-3) / 1e12, mac / (ms * 1e-3) / sms / (khz * 1e3), sms, khz);
    }
  }
  return 0;
}
