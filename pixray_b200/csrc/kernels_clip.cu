// CLIP-side row kernels: LayerNorm fwd/bwd on the fp32 residual stream, attention softmax fwd/bwd, the
// class-token head (ln_post + proj), the Prompt / spherical-distance loss with its gradient, Adam + clip_z.
#include "kernels.cuh"
#include <cfloat>
#include <cmath>

namespace pxr {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

constexpr int LN_MAX_PER_LANE = 32;  // W <= 1024

// one warp per row; W % 32 == 0
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                            int T, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int rows, int W, float eps,
                                                            act_t* __restrict__ y16, float* __restrict__ y32,
                                                            float* __restrict__ stats) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int per = W / 32;
  float v[LN_MAX_PER_LANE];
  const float* xr = x + (size_t)row * W;
  const float* pr = pos ? pos + (size_t)(row % T) * W : nullptr;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) {
      float t = xr[lane + 32 * i];
      if (pr) t += pr[lane + 32 * i];
      v[i] = t;
      s += t;
    }
  const float mean = warp_sum(s) / W;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) {
      float d = v[i] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / W + eps);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) {
      int c = lane + 32 * i;
      float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (y16) y16[(size_t)row * W + c] = __float2half_rn(o);
      if (y32) y32[(size_t)row * W + c] = o;
    }
}

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const act_t* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ pos, int T,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, int rows, int W,
                                                            int accumulate, float* __restrict__ gx,
                                                            act_t* __restrict__ gx16) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int per = W / 32;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const float* xr = x + (size_t)row * W;
  const float* pr = pos ? pos + (size_t)(row % T) * W : nullptr;
  float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) {
      int c = lane + 32 * i;
      float t = xr[c];
      if (pr) t += pr[c];
      xh[i] = (t - mean) * rstd;
      g[i] = __half2float(dy[(size_t)row * W + c]) * gamma[c];
      s1 += g[i];
      s2 += g[i] * xh[i];
    }
  s1 = warp_sum(s1) / W;
  s2 = warp_sum(s2) / W;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) {
      int c = lane + 32 * i;
      float d = rstd * (g[i] - s1 - xh[i] * s2);
      size_t o = (size_t)row * W + c;
      if (accumulate) d += gx[o];
      gx[o] = d;
      if (gx16) gx16[o] = __float2half_rn(d);
    }
}

// one warp per row, row length <= 1024 (cols), elements strided by lane
__global__ void __launch_bounds__(256) softmax_fwd_kernel(act_t* __restrict__ s, long long rows, int cols, int ld) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  act_t* r = s + row * ld;
  float v[32];
  float m = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int c = lane + 32 * i;
    v[i] = (c < cols) ? __half2float(r[c]) : -FLT_MAX;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int c = lane + 32 * i;
    v[i] = (c < cols) ? __expf(v[i] - m) : 0.f;
    sum += v[i];
  }
  const float inv = 1.f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int c = lane + 32 * i;
    if (c < ld) r[c] = __float2half_rn(v[i] * inv);
  }
}

__global__ void __launch_bounds__(256) softmax_bwd_kernel(const act_t* __restrict__ p, act_t* __restrict__ dp,
                                                          long long rows, int cols, int ld) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const act_t* pr = p + row * ld;
  act_t* dr = dp + row * ld;
  float pv[32], dv[32];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int c = lane + 32 * i;
    pv[i] = (c < cols) ? __half2float(pr[c]) : 0.f;
    dv[i] = (c < cols) ? __half2float(dr[c]) : 0.f;
    dot += pv[i] * dv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int c = lane + 32 * i;
    if (c < ld) dr[c] = __float2half_rn(pv[i] * (dv[i] - dot));
  }
}

// one block (256 threads) per image: LN of the class-token row, then e = ln @ proj
__global__ void __launch_bounds__(256) clip_head_fwd_kernel(const float* __restrict__ x, int T, int W, int D,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ proj, float eps,
                                                            float* __restrict__ stats, float* __restrict__ e) {
  extern __shared__ float sh[];  // [W] normalised row
  __shared__ float red[8];
  __shared__ float s_mean, s_rstd;
  const int b = blockIdx.x;
  const float* xr = x + (size_t)b * T * W;
  float s = 0.f;
  for (int c = threadIdx.x; c < W; c += 256) s += xr[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_mean = t / W;
  }
  __syncthreads();
  const float mean = s_mean;
  float q = 0.f;
  for (int c = threadIdx.x; c < W; c += 256) {
    float d = xr[c] - mean;
    q += d * d;
  }
  q = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_rstd = rsqrtf(t / W + eps);
    stats[2 * b] = mean;
    stats[2 * b + 1] = s_rstd;
  }
  __syncthreads();
  const float rstd = s_rstd;
  for (int c = threadIdx.x; c < W; c += 256) sh[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 256) {
    float acc = 0.f;
    for (int k = 0; k < W; ++k) acc = fmaf(sh[k], proj[(size_t)k * D + d], acc);
    e[(size_t)b * D + d] = acc;
  }
}

// one block per image: dln = proj @ de ; LN backward for the class row
__global__ void __launch_bounds__(256) clip_head_bwd_kernel(const float* __restrict__ de, const float* __restrict__ x,
                                                            int T, int W, int D, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ proj, float* __restrict__ gx,
                                                            act_t* __restrict__ gx16) {
  extern __shared__ float sh[];  // [D] de row, then [W] g
  float* s_de = sh;
  float* s_g = sh + D;
  __shared__ float red[2][8];
  __shared__ float s_s1, s_s2;
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += 256) s_de[d] = de[(size_t)b * D + d];
  __syncthreads();
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const float* xr = x + (size_t)b * T * W;
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < W; c += 256) {
    float acc = 0.f;
    const float* pr = proj + (size_t)c * D;
    for (int d = 0; d < D; ++d) acc = fmaf(pr[d], s_de[d], acc);
    float g = acc * gamma[c];
    s_g[c] = g;
    float xh = (xr[c] - mean) * rstd;
    s1 += g;
    s2 += g * xh;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      c += red[1][i];
    }
    s_s1 = a / W;
    s_s2 = c / W;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < W; c += 256) {
    float xh = (xr[c] - mean) * rstd;
    float d = rstd * (s_g[c] - s_s1 - xh * s_s2);
    size_t o = (size_t)b * T * W + c;
    gx[o] = d;
    gx16[o] = __float2half_rn(d);
  }
}

// one block (128 threads) per cutout row.  Accumulates losses with atomics (n prompts).
__global__ void __launch_bounds__(128) prompt_loss_kernel(const float* __restrict__ e, int D,
                                                          const float* __restrict__ prompts,
                                                          const float* __restrict__ weights,
                                                          const float* __restrict__ stops, int n, float inv_count,
                                                          float grad_scale, float* __restrict__ e_unit,
                                                          float* __restrict__ losses, float* __restrict__ de) {
  extern __shared__ float sh[];  // en [D], gen [D]
  float* en = sh;
  float* gen = sh + D;
  __shared__ float red[4];
  __shared__ float s_val;
  const int b = blockIdx.x;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) s_val = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return s_val;
  };
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += 128) {
    float t = e[(size_t)b * D + d];
    ss += t * t;
  }
  const float nrm = fmaxf(sqrtf(block_sum(ss)), 1e-12f);  // img_embeddings / norm (slip.py:66), F.normalize eps
  for (int d = threadIdx.x; d < D; d += 128) {
    float t = e[(size_t)b * D + d] / nrm;
    en[d] = t;
    gen[d] = 0.f;
    if (e_unit) e_unit[(size_t)b * D + d] = t;
  }
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    const float* pj = prompts + (size_t)j * D;
    float rr = 0.f;
    for (int d = threadIdx.x; d < D; d += 128) {
      float t = en[d] - pj[d];
      rr += t * t;
    }
    const float r = sqrtf(block_sum(rr));
    const float w = weights[j], stop = stops[j];
    const float sgn = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : 0.f);
    const float a = asinf(fminf(r * 0.5f, 1.f));
    const float dist = 2.f * a * a * sgn;                // pixray.py:278-279
    if (threadIdx.x == 0) atomicAdd(&losses[j], fabsf(w) * dist * inv_count);
    // replace_grad(dists, maximum(dists, stop)): gradient only where dists > stop (pixray.py:280)
    if (dist > stop && r > 0.f) {
      // d dist / d r = sgn * 2 asin(r/2) / sqrt(1 - r^2/4)
      const float ddr = sgn * 2.f * a / sqrtf(fmaxf(1.f - 0.25f * r * r, 1e-12f));
      const float coef = fabsf(w) * inv_count * grad_scale * ddr / r;
      for (int d = threadIdx.x; d < D; d += 128) gen[d] += coef * (en[d] - pj[d]);
    }
    __syncthreads();
  }
  // back through the normalisation: de = (gen - en * <gen, en>) / |e|
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += 128) dot += gen[d] * en[d];
  dot = block_sum(dot);
  for (int d = threadIdx.x; d < D; d += 128) de[(size_t)b * D + d] = (gen[d] - en[d] * dot) / nrm;
}

__global__ void adam_clip_kernel(float* __restrict__ z, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ g, float inv_scale, int n, int per_channel,
                                 const float* __restrict__ zmin, const float* __restrict__ zmax, int clip01,
                                 float step_size, float b1, float b2, float eps, float inv_sqrt_bc2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * inv_scale;
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  float zi = z[i] - step_size * mi / denom;
  if (zmin) {
    int c = i / per_channel;
    zi = fminf(fmaxf(zi, zmin[c]), zmax[c]);  // vqgan.py:202-204
  } else if (clip01) {
    zi = fminf(fmaxf(zi, 0.f), 1.f);  // fast_pixeldrawer.py:98-100
  }
  z[i] = zi;
}

__global__ void fill_kernel(float* p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}
__global__ void cast_kernel(const float* x, act_t* y, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(x[i] * scale);
}

}  // namespace

void layernorm_forward(const float* x, const float* pos, int T, const float* gamma, const float* beta, int rows, int W,
                       float eps, act_t* y16, float* y32, float* stats, cudaStream_t st) {
  layernorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, pos, T, gamma, beta, rows, W, eps, y16, y32, stats);
}
void layernorm_backward(const act_t* dy, const float* x, const float* pos, int T, const float* stats,
                        const float* gamma, int rows, int W, int accumulate, float* gx, act_t* gx16, cudaStream_t st) {
  layernorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(dy, x, pos, T, stats, gamma, rows, W, accumulate, gx, gx16);
}
void softmax_forward(act_t* s, int rows, int cols, int ld, cudaStream_t st) {
  softmax_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s, rows, cols, ld);
}
void softmax_backward(const act_t* p, act_t* dp_to_ds, int rows, int cols, int ld, cudaStream_t st) {
  softmax_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(p, dp_to_ds, rows, cols, ld);
}
void clip_head_forward(const float* x, int T, int W, int D, const float* gamma, const float* beta, const float* proj,
                       int B, float eps, float* stats, float* e, cudaStream_t st) {
  clip_head_fwd_kernel<<<B, 256, W * sizeof(float), st>>>(x, T, W, D, gamma, beta, proj, eps, stats, e);
}
void clip_head_backward(const float* de, const float* x, int T, int W, int D, const float* stats, const float* gamma,
                        const float* proj, int B, float* gx, act_t* gx16, cudaStream_t st) {
  cudaMemsetAsync(gx, 0, (size_t)B * T * W * sizeof(float), st);
  cudaMemsetAsync(gx16, 0, (size_t)B * T * W * sizeof(act_t), st);
  clip_head_bwd_kernel<<<B, 256, (D + W) * sizeof(float), st>>>(de, x, T, W, D, stats, gamma, proj, gx, gx16);
}
void prompt_loss(const float* e, int B, int D, const float* prompts, const float* weights, const float* stops, int n,
                 int cutn_global, float grad_scale, float* e_unit, float* losses, float* de, cudaStream_t st) {
  // Prompt.forward means over [cutn, n_embed=1] per prompt (pixray.py:280)
  prompt_loss_kernel<<<B, 128, 2 * D * sizeof(float), st>>>(e, D, prompts, weights, stops, n, 1.f / cutn_global,
                                                           grad_scale, e_unit, losses, de);
}

void adam_clip_step(float* z, float* m, float* v, const float* g, float inv_scale, int n, int per_channel,
                    const float* zmin, const float* zmax, int clip01, float lr, float b1, float b2, float eps, int t,
                    cudaStream_t st) {
  // torch.optim.Adam: step_size = lr / (1 - b1^t); denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
  adam_clip_kernel<<<(n + 255) / 256, 256, 0, st>>>(z, m, v, g, inv_scale, n, per_channel, zmin, zmax, clip01,
                                                    (float)(lr / bc1), b1, b2, eps, (float)(1.0 / sqrt(bc2)));
}

void fill_f32(float* p, float v, long long n, cudaStream_t st) {
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  fill_kernel<<<(int)g, 256, 0, st>>>(p, v, n);
}
void cast_f32_to_f16(const float* x, act_t* y, long long n, float scale, cudaStream_t st) {
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  cast_kernel<<<(int)g, 256, 0, st>>>(x, y, n, scale);
}

}  // namespace pxr
