// FFT drawer (fftdrawer.py:78-84 -> aphantasia fft_image + to_valid_rgb [UPSTREAM, un-vendored]):
//   image = sigmoid( M_color * ( irfft2(scale * spectrum, ortho) * contrast / std ) )
// The 2-D real FFT itself is a cuFFT call (not a contraction this engine owns); everything around it is
// fused pointwise / reduction kernels here.  cuFFT is resolved at run time (dlopen) like NCCL.
#include "kernels.cuh"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdlib>

namespace pxr {
namespace {

typedef int (*PlanManyFn)(int*, int, int*, int*, int, int, int*, int, int, int, int);
typedef int (*SetStreamFn)(int, cudaStream_t);
typedef int (*ExecC2RFn)(int, void*, float*);
typedef int (*ExecR2CFn)(int, float*, void*);
typedef int (*DestroyFn)(int);

struct CufftApi {
  void* lib = nullptr;
  PlanManyFn plan_many = nullptr;
  SetStreamFn set_stream = nullptr;
  ExecC2RFn exec_c2r = nullptr;
  ExecR2CFn exec_r2c = nullptr;
  DestroyFn destroy = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {getenv("PXR_CUFFT_LIB"), "libcufft.so.11", "libcufft.so.12", "libcufft.so",
                           "/usr/local/cuda/lib64/libcufft.so.11"};
    for (const char* n : names) {
      if (!n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    plan_many = (PlanManyFn)dlsym(lib, "cufftPlanMany");
    set_stream = (SetStreamFn)dlsym(lib, "cufftSetStream");
    exec_c2r = (ExecC2RFn)dlsym(lib, "cufftExecC2R");
    exec_r2c = (ExecR2CFn)dlsym(lib, "cufftExecR2C");
    destroy = (DestroyFn)dlsym(lib, "cufftDestroy");
    return plan_many && set_stream && exec_c2r && exec_r2c && destroy;
  }
};
CufftApi g_cufft;
constexpr int CUFFT_R2C = 0x2a, CUFFT_C2R = 0x2c;

// scaled[c, y, x2] = scale[y, x2] * spectrum[c, y, x2] (complex)
__global__ void fft_scale_kernel(const float2* __restrict__ spec, const float* __restrict__ scale, long long n_plane,
                                 long long n, float2* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = scale[i % n_plane];
    float2 v = spec[i];
    out[i] = make_float2(v.x * s, v.y * s);
  }
}

// partial sums {sum x, sum x^2} (double) per block; x scaled by `norm` (cuFFT transforms are unnormalised)
__global__ void __launch_bounds__(256) fft_moments_kernel(const float* __restrict__ x, float norm, long long n,
                                                          double* __restrict__ part) {
  double s = 0, q = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)(x[i] * norm);
    s += v;
    q += v * v;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = sh[0][0];
    part[2 * blockIdx.x + 1] = sh[1][0];
  }
}
// stats = {mean, std (unbiased, torch .std()), 1/std}
__global__ void fft_moments_final_kernel(const double* __restrict__ part, int nblk, double n, float* __restrict__ stats) {
  double s = 0, q = 0;
  for (int i = 0; i < nblk; ++i) {
    s += part[2 * i];
    q += part[2 * i + 1];
  }
  const double mean = s / n;
  double var = (q - n * mean * mean) / (n - 1.0);
  if (var < 1e-30) var = 1e-30;
  stats[0] = (float)mean;
  stats[1] = (float)sqrt(var);
  stats[2] = (float)(1.0 / sqrt(var));
}

// img[c', p] = sigmoid( sum_c M[c', c] * x[c, p] * norm * contrast / std );  x kept for backward
__global__ void fft_finish_kernel(const float* __restrict__ x, float norm, float contrast,
                                  const float* __restrict__ stats, const float* __restrict__ M, long long px,
                                  float* __restrict__ img) {
  const float k = norm * contrast * stats[2];
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < px; p += (long long)gridDim.x * blockDim.x) {
    const float x0 = x[p] * k, x1 = x[px + p] * k, x2 = x[2 * px + p] * k;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = M[c * 3] * x0 + M[c * 3 + 1] * x1 + M[c * 3 + 2] * x2;
      img[c * px + p] = 1.f / (1.f + __expf(-t));
    }
  }
}

// g1[c, p] = sum_c' M[c', c] * g_img[c', p] * s'(.) ; also partial A = sum g1 * x (double), x already * norm
__global__ void __launch_bounds__(256) fft_finish_bwd1_kernel(const float* __restrict__ g_img,
                                                              const float* __restrict__ img,
                                                              const float* __restrict__ x, float norm,
                                                              const float* __restrict__ M, long long px,
                                                              float* __restrict__ g1, double* __restrict__ part) {
  double a = 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < px; p += (long long)gridDim.x * blockDim.x) {
    float g2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = img[c * px + p];
      g2[c] = g_img[c * px + p] * s * (1.f - s);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = M[c] * g2[0] + M[3 + c] * g2[1] + M[6 + c] * g2[2];
      g1[c * px + p] = v;
      a += (double)v * (double)(x[c * px + p] * norm);
    }
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
// gx = (k/s) g1 - (k/s^2) A (x - mean) / ((N-1) s), in place into g1;  y = k x / s with k = contrast
__global__ void fft_finish_bwd2_kernel(float* __restrict__ g1, const float* __restrict__ x, float norm, float contrast,
                                       const float* __restrict__ stats, const double* __restrict__ part, int nblk,
                                       long long n) {
  __shared__ float sA;
  if (threadIdx.x == 0) {
    double a = 0;
    for (int i = 0; i < nblk; ++i) a += part[i];
    sA = (float)a;
  }
  __syncthreads();
  const float mean = stats[0], inv = stats[2];
  const float c1 = contrast * inv, c2 = contrast * inv * inv * inv * sA / (float)(n - 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g1[i] = c1 * g1[i] - c2 * (x[i] * norm - mean);
}

// z_grad (re, im) = scale * c_k * norm * G * inv_scale, c_k = 2 for interior columns, 1 for kx in {0, W/2}
__global__ void fft_spectrum_grad_kernel(const float2* __restrict__ G, const float* __restrict__ scale, int W2,
                                         int W_even, long long n_plane, long long n, float k, float2* __restrict__ zg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long pi = i % n_plane;
    const int kx = (int)(pi % W2);
    const float c = (kx == 0 || (W_even && kx == W2 - 1)) ? 1.f : 2.f;
    const float s = scale[pi] * c * k;
    float2 g = G[i];
    zg[i] = make_float2(g.x * s, g.y * s);
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g > 1184 ? 1184 : (g < 1 ? 1 : g));
}
constexpr int FFT_RED_BLOCKS = 296;

}  // namespace

struct FftPlans {
  int c2r = 0, r2c = 0;
  int H = 0, W = 0;
};

FftPlans* fft_plans_create(int H, int W, cudaStream_t st, const char** err) {
  if (!g_cufft.load()) {
    *err = "cannot load libcufft (set PXR_CUFFT_LIB)";
    return nullptr;
  }
  FftPlans* p = new FftPlans();
  p->H = H;
  p->W = W;
  int n[2] = {H, W};
  if (g_cufft.plan_many(&p->c2r, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, 3) != 0 ||
      g_cufft.plan_many(&p->r2c, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, 3) != 0) {
    *err = "cufftPlanMany failed";
    delete p;
    return nullptr;
  }
  g_cufft.set_stream(p->c2r, st);
  g_cufft.set_stream(p->r2c, st);
  return p;
}
void fft_plans_destroy(FftPlans* p) {
  if (!p) return;
  g_cufft.destroy(p->c2r);
  g_cufft.destroy(p->r2c);
  delete p;
}

// spectrum [3,H,W2,2] -> image [3,H,W]; keeps x (unnormalised irfft output), stats, for backward.
void fft_synth_forward(FftPlans* pl, const float* spectrum, const float* scale, const float* M, float contrast,
                       float* scaled, float* x, double* part, float* stats, float* img, cudaStream_t st) {
  const int H = pl->H, W = pl->W, W2 = W / 2 + 1;
  const long long n_plane = (long long)H * W2, nc = 3 * n_plane, px = (long long)H * W, n = 3 * px;
  const float norm = 1.f / sqrtf((float)px);  // norm="ortho"
  fft_scale_kernel<<<grid_for(nc), 256, 0, st>>>(reinterpret_cast<const float2*>(spectrum), scale, n_plane, nc,
                                                  reinterpret_cast<float2*>(scaled));
  g_cufft.exec_c2r(pl->c2r, scaled, x);  // note: C2R may overwrite its input (scaled is scratch)
  fft_moments_kernel<<<FFT_RED_BLOCKS, 256, 0, st>>>(x, norm, n, part);
  fft_moments_final_kernel<<<1, 1, 0, st>>>(part, FFT_RED_BLOCKS, (double)n, stats);
  fft_finish_kernel<<<grid_for(px), 256, 0, st>>>(x, norm, contrast, stats, M, px, img);
}

// g_img [3,H,W] (scaled by grad_scale) -> z_grad [3,H,W2,2] (unscaled)
void fft_synth_backward(FftPlans* pl, const float* g_img, const float* img, const float* x, const float* scale,
                        const float* M, float contrast, const float* stats, float* g1, float* G, double* part,
                        float inv_scale, float* z_grad, cudaStream_t st) {
  const int H = pl->H, W = pl->W, W2 = W / 2 + 1;
  const long long n_plane = (long long)H * W2, nc = 3 * n_plane, px = (long long)H * W, n = 3 * px;
  const float norm = 1.f / sqrtf((float)px);
  fft_finish_bwd1_kernel<<<FFT_RED_BLOCKS, 256, 0, st>>>(g_img, img, x, norm, M, px, g1, part);
  fft_finish_bwd2_kernel<<<grid_for(n), 256, 0, st>>>(g1, x, norm, contrast, stats, part, FFT_RED_BLOCKS, n);
  g_cufft.exec_r2c(pl->r2c, g1, G);
  fft_spectrum_grad_kernel<<<grid_for(nc), 256, 0, st>>>(reinterpret_cast<const float2*>(G), scale, W2,
                                                         (W % 2) == 0, n_plane, nc, norm * inv_scale,
                                                         reinterpret_cast<float2*>(z_grad));
}

}  // namespace pxr
