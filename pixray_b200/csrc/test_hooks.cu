// C-ABI test hooks: expose single kernels of the engine so tests/ can check each against the oracle / torch.
// These are not part of the reference-facing surface (include/pixray_b200.h lists them under "test hooks").
#include "../../include/pixray_b200.h"
#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "color_jitter.cuh"
#include "pool_bounds.cuh"
#include "kernels.cuh"
#include <cstdio>
#include <cstring>

using namespace pxr;

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static void fill_epilogue(GemmEpilogue& e, const pxr_test_gemm_desc* d) {
  e.alpha = d->alpha;
  e.bias = d->bias;
  e.bias_per_row = d->bias_per_row;
  e.act = d->act;
  e.aux_in = static_cast<const __half*>(d->aux_in);
  e.aux_out = static_cast<__half*>(d->aux_out);
  e.res_f32 = d->res_f32;
  e.res_f16 = static_cast<const __half*>(d->res_f16);
  e.out_f32 = d->out_f32;
  e.out_f16 = static_cast<__half*>(d->out_f16);
  e.ldc = d->ldc;
  e.bs0 = d->c_bs0;
  e.bs1 = d->c_bs1;
  e.n_store = d->n_store;
  e.cta_group = d->cta_group;
  e.tma_epi = d->tma_epi;
}

extern "C" int pxr_test_gemm(const pxr_test_gemm_desc* d, char* err, int errlen) {
  GemmOperand A, B;
  A.ptr = d->a;
  A.mode = d->a_mode;
  A.ld = d->lda;
  A.mn_extent = d->a_mn_extent;
  A.k_extent = d->a_k_extent;
  A.nb0 = d->nb0;
  A.nb1 = d->nb1;
  A.bs0 = d->a_bs0;
  A.bs1 = d->a_bs1;
  B.ptr = d->b;
  B.mode = d->b_mode;
  B.ld = d->ldb;
  B.mn_extent = d->b_mn_extent;
  B.k_extent = d->b_k_extent;
  B.nb0 = d->b_batched ? d->nb0 : 1;
  B.nb1 = d->b_batched ? d->nb1 : 1;
  B.bs0 = d->b_bs0;
  B.bs1 = d->b_bs1;
  GemmEpilogue e;
  fill_epilogue(e, d);
  GemmPlan plan;
  int rc = gemm_plan_make(&plan, A, B, d->M, d->N, d->K, e, d->block_n, d->fmt, num_sms(), err, errlen);
  if (rc) return rc;
  for (int i = 0; i < (d->repeat > 0 ? d->repeat : 1); ++i) gemm_launch(plan, static_cast<cudaStream_t>(d->stream));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    if (err) snprintf(err, errlen, "launch failed: %s", cudaGetErrorString(ce));
    return -100;
  }
  return 0;
}

extern "C" int pxr_test_conv(const pxr_test_gemm_desc* d, int batch, int H, int W, int c_in, int cout_pad, int ksize,
                             char* err, int errlen) {
  GemmEpilogue e;
  fill_epilogue(e, d);
  GemmPlan plan;
  int rc = conv_plan_make(&plan, d->a, d->lda, batch, H, W, c_in, d->b, cout_pad, d->N, ksize, e, d->block_n, d->fmt,
                          num_sms(), err, errlen);
  if (rc) return rc;
  for (int i = 0; i < (d->repeat > 0 ? d->repeat : 1); ++i) gemm_launch(plan, static_cast<cudaStream_t>(d->stream));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    if (err) snprintf(err, errlen, "launch failed: %s", cudaGetErrorString(ce));
    return -100;
  }
  return 0;
}

// Fused attention forward (+ backward when d_o != NULL) on caller tensors: qkv [B*T, 3W] fp16, o [B*T, W] fp16,
// lse [B*H*T] fp32, d_o [B*T, W] fp16, gqkv [B*T, 3W] fp16.
extern "C" int pxr_test_attention(const void* qkv, void* o, float* lse, const void* d_o, void* gqkv, int B, int T, int H,
                                  int W, float scale, int repeat, char* err, int errlen) {
  AttnPlan plan;
  int rc = attn_plan_make(&plan, static_cast<const __half*>(qkv), static_cast<__half*>(o),
                          static_cast<const __half*>(d_o), static_cast<__half*>(gqkv), lse, B, T, H, W, scale,
                          num_sms(), err, errlen);
  if (rc) return rc;
  for (int i = 0; i < (repeat > 0 ? repeat : 1); ++i) {
    attn_forward_launch(plan, nullptr);
    if (d_o) attn_backward_launch(plan, nullptr);
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    if (err) snprintf(err, errlen, "launch failed: %s", cudaGetErrorString(ce));
    return -100;
  }
  return 0;
}

// Host evaluation of the per-pixel ColorJitter body the cutout kernels run (color_jitter.cuh is host+device code):
// lets the CPU suite compare the arithmetic and its Jacobian with the oracle without a GPU.  rgb / g_out / out / g_in
// are HOST [n, 3]; g_out and g_in may be NULL.
extern "C" int pxr_test_color_jitter_host(const float* rgb, int n, int code, float saturation, float hue,
                                          const float* g_out, float* out, float* g_in) {
  if (!rgb || !out || n < 0) return -1;
  for (int i = 0; i < n; ++i) {
    float c[3] = {rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
    pxr::cj_apply(c, code, saturation, hue);
    for (int k = 0; k < 3; ++k) out[3 * i + k] = c[k];
    if (g_out && g_in) pxr::cj_vjp(rgb + 3 * i, code, saturation, hue, g_out + 3 * i, g_in + 3 * i);
  }
  return 0;
}

// The same body on the device (what cutout_fwd / cutout_bwd execute per pixel), for a device-vs-host comparison of the
// forward and of the vector-Jacobian product.  rgb / g_out / out / g_in are DEVICE [n, 3].
namespace {
__global__ void color_jitter_device_kernel(const float* rgb, int n, int code, float saturation, float hue,
                                           const float* g_out, float* out, float* g_in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float in[3] = {rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
  float c[3] = {in[0], in[1], in[2]};
  pxr::cj_apply(c, code, saturation, hue);
  for (int k = 0; k < 3; ++k) out[3 * i + k] = c[k];
  if (g_out && g_in) {
    const float go[3] = {g_out[3 * i], g_out[3 * i + 1], g_out[3 * i + 2]};
    float gi[3];
    pxr::cj_vjp(in, code, saturation, hue, go, gi);
    for (int k = 0; k < 3; ++k) g_in[3 * i + k] = gi[k];
  }
}
}  // namespace

extern "C" int pxr_test_color_jitter_device(const float* rgb, int n, int code, float saturation, float hue,
                                            const float* g_out, float* out, float* g_in) {
  if (!rgb || !out || n < 0) return -1;
  if (n == 0) return 0;
  color_jitter_device_kernel<<<(n + 255) / 256, 256>>>(rgb, n, code, saturation, hue, g_out, out, g_in);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// The adaptive-pool window bounds as the DEVICE computes them (pool_bounds.cuh, the functions pool_fwd / pool_bwd call):
// starts / ends are DEVICE int [out_size].  Integer bookkeeping: the test asserts equality with ATen's formula.
namespace {
__global__ void pool_bounds_kernel(int in_size, int out_size, int* starts, int* ends) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  starts[i] = pxr::pool_start(i, in_size, out_size);
  ends[i] = pxr::pool_end(i, in_size, out_size);
}
}  // namespace

extern "C" int pxr_test_pool_bounds(int in_size, int out_size, int* starts, int* ends) {
  if (in_size < 1 || out_size < 1 || !starts || !ends) return -1;
  pool_bounds_kernel<<<(out_size + 255) / 256, 256>>>(in_size, out_size, starts, ends);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// GroupNorm(32, C, eps 1e-6)(+ swish) forward, and backward when dy (or ws_dy) is given, on caller DEVICE tensors
// (NHWC fp16 [pixels, C]).  variant 0: the single-kernel grid-barrier GroupNorm; 1: one cluster per group
// (kernels_gn_group.cu).  variant 1 only: ws != NULL makes the forward the epilogue of a split-K convolution
// (x_out = fp16(sum_s ws[s] + bias + res), then normalised); ws_dy != NULL feeds the backward from split-K partials.
// scratch: >= 64 * (num_sms + 2) floats + one zero-initialised u64 at its END (variant 0's grid barrier).
extern "C" int pxr_test_groupnorm(int variant, const void* x, const float* ws, int splits, const float* bias, const void* res,
                                  void* x_out, const float* gamma, const float* beta, int pixels, int C, int swish, void* y,
                                  float* stats, const void* dy, const float* ws_dy, int splits_dy, const void* dres, void* dx,
                                  float* scratch, int repeat) {
  const act_t* xin = static_cast<const act_t*>(x);
  const int nsm = num_sms();
  GridBarrier gb;  // the caller hands in a zeroed counter with every call
  gb.counter = reinterpret_cast<unsigned long long*>(scratch + 64 * (nsm + 2));
  if (variant == 0) {
    if (ws || ws_dy || !gn_coop_supported(pixels, C, nsm)) return -1;
  } else {
    if (!gn_group_possible(pixels, C)) return -2;
  }
  for (int it = 0; it < (repeat > 0 ? repeat : 1); ++it) {
    if (variant == 0) {
      gn_forward_coop(xin, gamma, beta, pixels, C, swish, 1e-6f, scratch, stats, static_cast<act_t*>(y), nsm, &gb, nullptr);
      if (dy)
        gn_backward_coop(static_cast<const act_t*>(dy), xin, stats, gamma, beta, pixels, C, swish,
                         static_cast<const act_t*>(dres), scratch, static_cast<act_t*>(dx), nsm, &gb, nullptr);
    } else {
      GnSplitK sk, skd;
      sk.ws = ws;
      sk.splits = splits;
      sk.ld_ws = C;
      sk.bias = bias;
      sk.res = static_cast<const act_t*>(res);
      sk.out = static_cast<act_t*>(x_out);
      skd.ws = ws_dy;
      skd.splits = splits_dy;
      skd.ld_ws = C;
      gn_forward_group(xin, ws ? &sk : nullptr, gamma, beta, pixels, C, swish, 1e-6f, stats, static_cast<act_t*>(y), nullptr);
      const act_t* xs = ws ? static_cast<const act_t*>(x_out) : xin;
      if (dy || ws_dy)
        gn_backward_group(static_cast<const act_t*>(dy), ws_dy ? &skd : nullptr, xs, stats, gamma, beta, pixels, C, swish,
                          static_cast<const act_t*>(dres), static_cast<act_t*>(dx), nullptr);
    }
  }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce == cudaSuccess) ce = cudaGetLastError();
  return ce == cudaSuccess ? 0 : -100;
}
