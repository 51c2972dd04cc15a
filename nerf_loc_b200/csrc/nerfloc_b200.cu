// Single translation unit of libnerfloc_b200.so: the kernels share __global__ helpers (pack kernels) and inline
// device code, so they are compiled together instead of with relocatable device code.
#include "pack.cu"
#include "knn.cu"
#include "neighbor_tc.cu"
#include "render_point.cu"
#include "render_ray.cu"
#include "render_ray2.cu"
#include "render_ray_long.cu"
#include "hier_sample.cu"
#include "match.cu"
#include "pnp.cu"
#include "tc_test.cu"
#include "cabi.cu"
