// C-ABI shim around the REFERENCE's own CUDA KNN (nerf_loc/models/ops/knn/src/knn.cu, the vendored pytorch3d kernels) so that
// bench.py can time it on the same box as this repo's search.  The reference translation unit is compiled from where it lies by
// oracle/Makefile (target ref_cuda); nothing is copied.  Comparator only: never on the product path.
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <tuple>

std::tuple<at::Tensor, at::Tensor> KNearestNeighborIdxCuda(
    const at::Tensor& p1, const at::Tensor& p2, const at::Tensor& lengths1, const at::Tensor& lengths2, int K, int version);

// p1 [n1, D], p2 [n2, D]: DEVICE pointers (fp32); idx [n1, K] int64, dist [n1, K] fp32: DEVICE pointers
extern "C" void knn_ref_cuda(const float* p1, int64_t n1, const float* p2, int64_t n2, int D, int K, int version,
                             int64_t* idx, float* dist) {
  auto opt = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
  auto a = torch::from_blob(const_cast<float*>(p1), {1, n1, D}, opt);
  auto b = torch::from_blob(const_cast<float*>(p2), {1, n2, D}, opt);
  auto l1 = torch::full({1}, n1, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA));
  auto l2 = torch::full({1}, n2, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA));
  auto r = KNearestNeighborIdxCuda(a, b, l1, l2, K, version);
  auto i = std::get<0>(r).contiguous();
  auto d = std::get<1>(r).contiguous();
  cudaMemcpyAsync(idx, i.data_ptr<int64_t>(), sizeof(int64_t) * n1 * K, cudaMemcpyDeviceToDevice, 0);
  cudaMemcpyAsync(dist, d.data_ptr<float>(), sizeof(float) * n1 * K, cudaMemcpyDeviceToDevice, 0);
  cudaStreamSynchronize(0);
}
