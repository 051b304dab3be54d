// Integer window bounds of Adaptive{Avg,Max}Pool2d as ATen computes them (start_index / end_index):
//   [floor(i * in / out), ceil((i + 1) * in / out))
// Shared by the pooling kernels (kernels_cutouts.cu) and the bookkeeping test hook (test_hooks.cu), so the test
// reads exactly the arithmetic the kernels run.
#pragma once
#include <cuda_runtime.h>

namespace pxr {
__host__ __device__ __forceinline__ int pool_start(int i, int in, int out) { return (int)(((long long)i * in) / out); }
__host__ __device__ __forceinline__ int pool_end(int i, int in, int out) {
  return (int)((((long long)(i + 1)) * in + out - 1) / out);
}
}  // namespace pxr
