// C-ABI shim around the REFERENCE's own CPU KNN so tests can call it through ctypes.
// The reference translation unit is compiled from where it lies
// (/root/reference/nerf_loc/models/ops/knn/src/knn_cpu.cpp) by oracle/Makefile; nothing is copied.
#include <torch/extension.h>
#include <tuple>

std::tuple<at::Tensor, at::Tensor> KNearestNeighborIdxCpu(
    const at::Tensor& p1, const at::Tensor& p2, const at::Tensor& lengths1, const at::Tensor& lengths2, int K);

extern "C" void knn_ref(const float* p1, int64_t n1, const float* p2, int64_t n2, int D, int K,
                        int64_t* idx, float* dist) {
  auto a = torch::from_blob(const_cast<float*>(p1), {1, n1, D}, torch::kFloat32);
  auto b = torch::from_blob(const_cast<float*>(p2), {1, n2, D}, torch::kFloat32);
  auto l1 = torch::full({1}, n1, torch::kInt64);
  auto l2 = torch::full({1}, n2, torch::kInt64);
  auto r = KNearestNeighborIdxCpu(a, b, l1, l2, K);
  auto i = std::get<0>(r).contiguous();
  auto d = std::get<1>(r).contiguous();
  std::memcpy(idx, i.data_ptr<int64_t>(), sizeof(int64_t) * n1 * K);
  std::memcpy(dist, d.data_ptr<float>(), sizeof(float) * n1 * K);
}
