// Engine internals shared by engine_*.cu: device allocation, weight store, op lists.
#pragma once
#include "../../include/pixray_b200.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include <cuda_runtime.h>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pxr {

struct HostWeight {
  std::vector<float> data;
  std::vector<int64_t> dims;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
};

struct EngineError : std::runtime_error {
  int code;
  EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PXR_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw pxr::EngineError(-200, std::string(#expr) + " failed: " + cudaGetErrorString(_e));      \
  } while (0)

// NHWC fp16 activation with its gradient buffer
struct Act {
  act_t* p = nullptr;
  act_t* g = nullptr;
  int H = 0, W = 0, C = 0;
  int pixels() const { return H * W; }
};

struct ConvW {
  act_t* w = nullptr;       // forward  [taps][cout_pad][cin]      (K-major in cin)
  act_t* wd = nullptr;      // dgrad    [taps][cin_rows][cout_k]   (flipped taps, K-major in cout)
  float* bias = nullptr;    // [cout]
  int cin = 0, cout = 0, cout_pad = 0, cout_k = 0, cin_rows = 0, ks = 1;
};

struct NormW {
  float* gamma = nullptr;
  float* beta = nullptr;
};

typedef std::function<void()> Op;
struct OpList {
  std::vector<Op> ops;
  std::vector<int> launches;
  std::vector<double> flops;  // > 0: a gemm_tc_kernel launch with this many algorithmic FLOPs
  std::vector<std::string> names;  // label for the per-op profile dump (PXR_PROFILE_DUMP)
  std::vector<double> bytes;  // algorithmic (compulsory) operand + result bytes of a tensor-core launch, 0 otherwise
  void add(int n_launch, Op f, double fl = 0.0, std::string name = std::string(), double by = 0.0) {
    ops.push_back(std::move(f));
    launches.push_back(n_launch);
    flops.push_back(fl);
    names.push_back(std::move(name));
    bytes.push_back(by);
  }
  void append(const OpList& o) {
    ops.insert(ops.end(), o.ops.begin(), o.ops.end());
    launches.insert(launches.end(), o.launches.begin(), o.launches.end());
    flops.insert(flops.end(), o.flops.begin(), o.flops.end());
    names.insert(names.end(), o.names.begin(), o.names.end());
    bytes.insert(bytes.end(), o.bytes.begin(), o.bytes.end());
  }
};

class Engine;

}  // namespace pxr
