// Fused multi-head self-attention of the CLIP ViT (nn.MultiheadAttention inside SLIP/models.py:18-64 residual blocks;
// openai-CLIP VisionTransformer) for sm_100a: one kernel for the forward, one for the backward, both on tcgen05 with
// the accumulators in TMEM.  Replaces, per layer, {QK^T GEMM, softmax, PV GEMM} and {dP GEMM, softmax backward,
// dQ / dK / dV GEMMs} of the unfused path (which stays as the fallback for sequences longer than 256 tokens).
//
// One CTA owns one (image, head) pair at a time: T <= 256 tokens, head width 64, so the whole score matrix of the pair
// (two 128-row accumulator tiles of N = round_up(T, 16) columns) lives in TMEM and is never written to memory.
// Forward keeps only the row log-sum-exp; backward recomputes P = exp(S - lse) from Q, K (flash-attention style) and
// uses D_i = dO_i . O_i for the softmax backward.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pxr {

struct AttnParams {
  CUtensorMap tm_q;   // qkv [B][T][3W] fp16, box 64 x 128 rows
  CUtensorMap tm_kv;  // same tensor, box 64 x n_pad rows
  CUtensorMap tm_do;  // dO [B][T][W] fp16, box 64 x 128 rows (backward only)
  int T, n_pad, H, B, W, n_mt, items;
  float scale;            // 1 / sqrt(64): S = scale * q k^T
  __half* o;              // forward out / backward in: [B*T, W] (head h at columns h*64)
  float* lse;             // [B*H, T] natural-log sum-exp of the scaled scores
  __half* gqkv;           // backward out [B*T, 3W]: dq | dk | dv
};

struct AttnPlan {
  AttnParams p;
  int grid = 0;
  int smem_fwd = 0, smem_bwd = 0, smem_bwd2 = 0;
  double flops_fwd = 0, flops_bwd = 0;
  double bytes_fwd = 0, bytes_bwd = 0;  // compulsory traffic: qkv in, o out (+ lse) / qkv, o, dO in, dqkv out
};

// true when the fused kernels cover this shape (head width 64, T <= 256, 16-byte aligned rows)
bool attn_supported(int T, int head_dim, int W);
// qkv: [B*T, 3W]; o: [B*T, W]; d_o: [B*T, W]; gqkv: [B*T, 3W]; lse: [B*H*T]
int attn_plan_make(AttnPlan* plan, const __half* qkv, __half* o, const __half* d_o, __half* gqkv, float* lse, int B,
                   int T, int H, int W, float scale, int num_sms, char* err, int errlen);
void attn_forward_launch(const AttnPlan& plan, cudaStream_t st);
void attn_backward_launch(const AttnPlan& plan, cudaStream_t st);

}  // namespace pxr
