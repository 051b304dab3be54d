// See gemm_tc.cuh for the contract.  sm_100a only (tcgen05 + TMA + TMEM).
#include "gemm_tc.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include <cstdio>
#include <cstring>
#include <mutex>

namespace pxr {
using namespace ptx;

namespace {

constexpr uint32_t A_TILE_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KiB
constexpr uint32_t MN_ATOM_BYTES = 64 * GEMM_BLOCK_K * 2;           // one 64(MN) x 64(K) MN-major TMA box

struct TileCoord {
  int tn, tm, z, b0, b1, n0, m0, h0, w0;
};

__device__ __forceinline__ TileCoord make_coord(const GemmParams& p, int tn, int tm, int z) {
  TileCoord t;
  t.tn = tn;
  t.tm = tm;
  t.z = z;
  t.b0 = z % p.nb0;
  t.b1 = z / p.nb0;
  t.n0 = tn * p.block_n;
  t.m0 = tm * GEMM_BLOCK_M;
  t.h0 = 0;
  t.w0 = 0;
  if (p.a_mode == OP_CONV) {
    t.h0 = (tm / p.tiles_w) * p.tile_h;
    t.w0 = (tm % p.tiles_w) * p.tile_w;
  }
  return t;
}

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  t.tn = tile % p.tiles_n;
  int r = tile / p.tiles_n;
  t.tm = r % p.tiles_m;
  t.z = r / p.tiles_m;
  t.b0 = t.z % p.nb0;
  t.b1 = t.z / p.nb0;
  t.n0 = t.tn * p.block_n;
  t.m0 = t.tm * GEMM_BLOCK_M;
  t.h0 = 0;
  t.w0 = 0;
  if (p.a_mode == OP_CONV) {
    t.h0 = (t.tm / p.tiles_w) * p.tile_h;
    t.w0 = (t.tm % p.tiles_w) * p.tile_w;
  }
  return t;
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float quickgelu(float x) { return x * fast_sigmoid(1.702f * x); }
__device__ __forceinline__ float quickgelu_grad(float u) {
  float s = fast_sigmoid(1.702f * u);
  return s * (1.f + 1.702f * u * (1.f - s));
}

constexpr int TB_LD = 33;                       // transposer row pitch (floats): conflict-free both ways
constexpr int TB_FLOATS = 32 * TB_LD;           // one 32 x 32 block per epilogue warp
constexpr int TB_WARP_BYTES = TB_FLOATS * 4 + 32 * 8 + 32 * 4;  // + per-row offset / row-index tables
constexpr int GEMM_EPI_WARPS = 8;

// Row <-> column layout exchange through the warp-private smem block.
//   row layout   : thread = accumulator row (TMEM lane), registers = columns     (what tcgen05.ld delivers)
//   store layout : lane = (row_sub, 16-byte chunk of the row)                    (what coalesced global access wants)
struct RowInfo {
  long long off;  // element offset of this thread's row in the output tensors, -1 if the row is out of range
  int row;        // logical row index (bias_per_row)
};

__device__ __forceinline__ RowInfo row_info(const GemmParams& p, const TileCoord& t, int r_tile) {
  RowInfo ri;
  bool valid;
  if (p.a_mode == OP_CONV) {
    const int h = t.h0 + r_tile / p.tile_w, w = t.w0 + r_tile % p.tile_w;
    valid = (h < p.conv_H) && (w < p.conv_W);
    ri.row = h * p.conv_W + w;
    ri.off = (static_cast<long long>(t.z) * p.conv_H * p.conv_W + ri.row) * p.ldc;
  } else {
    ri.row = t.m0 + r_tile;
    valid = ri.row < p.M;
    ri.off = t.b0 * p.out_bs0 + t.b1 * p.out_bs1 + static_cast<long long>(ri.row) * p.ldc;
  }
  if (!valid) ri.off = -1;
  return ri;
}

// Load `width` (16 or 32) accumulator columns of this thread's row, starting at TMEM address `taddr`.
// Both 16-column loads are issued before the single tcgen05.wait::ld (the wait, not the load, is the expensive part).
__device__ __forceinline__ void load_acc(uint32_t taddr, int width, float (&v)[32]) {
  uint32_t u0[16], u1[16];
  tmem_ld_x16(taddr, u0);
  if (width > 16) tmem_ld_x16(taddr + 16, u1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    v[j] = __uint_as_float(u0[j]);
    v[16 + j] = (width > 16) ? __uint_as_float(u1[j]) : 0.f;
  }
}

__device__ __forceinline__ void stage_rows(float* tb, const float (&v)[32], int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) tb[lane * TB_LD + j] = v[j];
  __syncwarp();
}

template <int E>
__device__ __forceinline__ void ld_f16(const __half* p, bool vec, int nv, float (&x)[E]) {
  if (vec) {
    if (E == 8) {
      uint4 u = *reinterpret_cast<const uint4*>(p);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        x[2 * j] = f.x;
        x[2 * j + 1] = f.y;
      }
    } else {
      uint2 u = *reinterpret_cast<const uint2*>(p);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float2 f = __half22float2(h[j]);
        x[2 * j] = f.x;
        x[2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = (j < nv) ? __half2float(p[j]) : 0.f;
  }
}
template <int E>
__device__ __forceinline__ void st_f16(__half* p, bool vec, int nv, const float (&x)[E]) {
  if (vec) {
    if (E == 8) {
      uint4 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      *reinterpret_cast<uint4*>(p) = u;
    } else {
      uint2 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 2; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      *reinterpret_cast<uint2*>(p) = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (j < nv) p[j] = __float2half_rn(x[j]);
  }
}
template <int E>
__device__ __forceinline__ void ld_f32(const float* p, bool vec, int nv, float (&x)[E]) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < E / 4; ++j) {
      float4 f = reinterpret_cast<const float4*>(p)[j];
      x[4 * j] = f.x;
      x[4 * j + 1] = f.y;
      x[4 * j + 2] = f.z;
      x[4 * j + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = (j < nv) ? p[j] : 0.f;
  }
}
template <int E>
__device__ __forceinline__ void st_f32(float* p, bool vec, int nv, const float (&x)[E]) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < E / 4; ++j)
      reinterpret_cast<float4*>(p)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (j < nv) p[j] = x[j];
  }
}

// Geometry of the store layout for E elements per lane (8: all-fp16 tensors, 4: an fp32 tensor takes part).
template <int E>
struct BlockMap {
  static constexpr int CH = 32 / E;    // 16-byte chunks per 32-column row
  static constexpr int RPI = 32 / CH;  // rows covered by one warp instruction
  static constexpr int NIT = 32 / RPI; // iterations per block
};

// Which input tensor (if any) is prefetched for the block: residual fp32, residual fp16 or the saved pre-activation.
__device__ __forceinline__ int prefetch_kind(const GemmParams& p) {
  return p.res_f32 ? 1 : (p.res_f16 ? 2 : (p.act == ACT_QUICKGELU_BWD ? 3 : 0));
}

// Issue the global loads of the block's input tensor BEFORE the TMEM load / smem transpose, so their DRAM latency
// overlaps it (a load-use chain per row made the fp32-residual GEMMs latency-bound).
template <int E>
__device__ __forceinline__ void prefetch_block(const GemmParams& p, const long long* s_off, int lane, int col_base,
                                               int col_limit, int kind, uint4 (&pf)[BlockMap<E>::NIT]) {
  using M = BlockMap<E>;
  const int chunk = lane % M::CH, rsub = lane / M::CH;
  const int col = col_base + chunk * E;
  const bool vec = p.vec_ok && (col_limit - col >= E);
  if (!vec || kind == 0) return;
#pragma unroll
  for (int it = 0; it < M::NIT; ++it) {
    const long long off_r = s_off[it * M::RPI + rsub];
    if (off_r < 0) continue;
    const long long o = off_r + col;
    if (kind == 1) {
      if (E == 4) pf[it] = *reinterpret_cast<const uint4*>(p.res_f32 + o);
    } else {
      const __half* src = (kind == 2) ? p.res_f16 : p.aux_in;
      if (E == 8) pf[it] = *reinterpret_cast<const uint4*>(src + o);
      else {
        uint2 u = *reinterpret_cast<const uint2*>(src + o);
        pf[it].x = u.x;
        pf[it].y = u.y;
      }
    }
  }
}

template <int E>
__device__ __forceinline__ void unpack_pf(const uint4& u, int kind, float (&x)[E]) {
  if (kind == 1) {
    x[0] = __uint_as_float(u.x);
    x[1] = __uint_as_float(u.y);
    x[2] = __uint_as_float(u.z);
    x[3] = __uint_as_float(u.w);
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < E / 2; ++j) {
      float2 f = __half22float2(h[j]);
      x[2 * j] = f.x;
      x[2 * j + 1] = f.y;
    }
  }
}

// Generic epilogue of one staged 32 x 32 block.  E elements per lane: 8 when every tensor is fp16 (16-byte accesses,
// 8 rows per warp instruction), 4 when an fp32 tensor takes part (16-byte accesses on the fp32 side, 4 rows).
template <int E>
__device__ __forceinline__ void epilogue_block(const GemmParams& p, const float* tb, const long long* s_off,
                                               const int* s_row, int lane, int col_base, int col_limit, int kind,
                                               const uint4 (&pf)[BlockMap<E>::NIT]) {
  using M = BlockMap<E>;
  const int chunk = lane % M::CH, rsub = lane / M::CH;
  const int c0 = chunk * E;
  const int col = col_base + c0;
  int nv = col_limit - col;  // col_limit = min(N, end of this tile)
  nv = nv < 0 ? 0 : (nv > E ? E : nv);
  const bool vec = p.vec_ok && nv == E;
  float bias[E];
#pragma unroll
  for (int j = 0; j < E; ++j) bias[j] = (p.bias && !p.bias_per_row && j < nv) ? p.bias[col + j] : 0.f;
  if (nv == 0) return;
#pragma unroll
  for (int it = 0; it < M::NIT; ++it) {
    const int r = it * M::RPI + rsub;
    const long long off_r = s_off[r];
    if (off_r < 0) continue;
    const long long o = off_r + col;
    float x[E];
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = tb[r * TB_LD + c0 + j] * p.alpha + bias[j];
    if (p.bias && p.bias_per_row) {
      const float b = p.bias[s_row[r]];
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] += b;
    }
    if (p.act == ACT_QUICKGELU) {
      if (p.aux_out) st_f16<E>(p.aux_out + o, vec, nv, x);
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] = quickgelu(x[j]);
    } else if (p.act == ACT_QUICKGELU_BWD) {
      float u[E];
      if (vec && kind == 3) unpack_pf<E>(pf[it], 3, u);
      else ld_f16<E>(p.aux_in + o, vec, nv, u);
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] *= quickgelu_grad(u[j]);
    }
    if (p.res_f32) {
      float t[E];
      if (vec && kind == 1 && E == 4) unpack_pf<E>(pf[it], 1, t);
      else ld_f32<E>(p.res_f32 + o, vec, nv, t);
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] += t[j];
    }
    if (p.res_f16) {
      float t[E];
      if (vec && kind == 2) unpack_pf<E>(pf[it], 2, t);
      else ld_f16<E>(p.res_f16 + o, vec, nv, t);
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] += t[j];
    }
    if (p.out_f32) st_f32<E>(p.out_f32 + o, vec, nv, x);
    if (p.out_f16) st_f16<E>(p.out_f16 + o, vec, nv, x);
  }
}

// fp16 store of a staged block with zero fill up to n_store (softmax epilogues), lane = (8 rows, 4 chunks of 8)
__device__ __forceinline__ void store_block_f16(const GemmParams& p, const float* tb, const long long* s_off, int lane,
                                                int col_base) {
  const int chunk = lane & 3, rsub = lane >> 2;
  const int col = col_base + chunk * 8;
  int nv = p.n_store - col;
  nv = nv < 0 ? 0 : (nv > 8 ? 8 : nv);
  if (nv == 0) return;
  const bool vec = p.vec_ok && nv == 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + rsub;
    const long long off_r = s_off[r];
    if (off_r < 0) continue;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = tb[r * TB_LD + chunk * 8 + j];
    st_f16<8>(p.out_f16 + off_r + col, vec, nv, x);
  }
}

// Fused row softmax (attention probabilities straight out of the QK^T accumulator, nn.MultiheadAttention /
// taming AttnBlock): requires the whole row in one tile (tiles_n == 1).  Thread = row.  Columns [N, n_store) are
// written as zeros so the result can be consumed as a K-major operand with a padded row pitch.
__device__ __forceinline__ void epilogue_softmax_fwd(const GemmParams& p, float* tb, const long long* s_off,
                                                     uint32_t taddr, int lane) {
  const int nslab = (p.block_n + 31) / 32;
  float v[32];
  float mx = -3.0e38f;
  for (int s = 0; s < nslab; ++s) {
    load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (s * 32 + j < p.N) mx = fmaxf(mx, v[j] * p.alpha);
  }
  float sum = 0.f;
  for (int s = 0; s < nslab; ++s) {
    load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (s * 32 + j < p.N) sum += __expf(v[j] * p.alpha - mx);
  }
  const float inv = __fdividef(1.f, sum);
  for (int s = 0; s < nslab; ++s) {
    load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (s * 32 + j < p.N) ? __expf(v[j] * p.alpha - mx) * inv : 0.f;
    stage_rows(tb, v, lane);
    store_block_f16(p, tb, s_off, lane, s * 32);
    __syncwarp();
  }
}

// Fused softmax backward: accumulator = dP row, aux_in = P (fp16, same indexing as the output);
// dS = alpha * P * (dP - sum_j P_j dP_j).
__device__ __forceinline__ void epilogue_softmax_bwd(const GemmParams& p, float* tb, const long long* s_off,
                                                     uint32_t taddr, int lane) {
  const int nslab = (p.block_n + 31) / 32;
  const int chunk = lane & 3, rsub = lane >> 2;
  float v[32];
  float dot = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    for (int s = 0; s < nslab; ++s) {
      // P block: coalesced 16-byte loads (store layout) -> smem -> row layout
      const int col = s * 32 + chunk * 8;
      int nv = p.N - col;
      nv = nv < 0 ? 0 : (nv > 8 ? 8 : nv);
      const bool vec = p.vec_ok && nv == 8;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + rsub;
        const long long off_r = s_off[r];
        float x[8];
        if (off_r >= 0 && nv > 0) {
          ld_f16<8>(p.aux_in + off_r + col, vec, nv, x);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tb[r * TB_LD + chunk * 8 + j] = x[j];
      }
      __syncwarp();
      load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
      if (pass == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dot += tb[lane * TB_LD + j] * v[j];
        __syncwarp();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = p.alpha * tb[lane * TB_LD + j] * (v[j] - dot);
        __syncwarp();
        stage_rows(tb, v, lane);
        store_block_f16(p, tb, s_off, lane, s * 32);
        __syncwarp();
      }
    }
  }
}

// TMA producer (one lane): walks this CTA's tiles and fills the stage ring.
// This single thread paces the whole kernel (one k-block of MMA work is ~512 cycles at block_n = 256): ncu showed the MMA
// thread finding its stage not yet full on 93 % of the k-blocks while the producer never waited for a free slot -- the
// loop body was ~160 dependent integer instructions (runtime divisions for the tap / k offset, operand-mode branches,
// generic->shared address conversions).  The body below is specialised on the operand modes, carries (tap, kk) as
// running counters and works on 32-bit shared addresses: ~25 instructions per k-block.
template <int AM, int BM>
__device__ __forceinline__ void producer_tiles(const GemmParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                               uint32_t stage_bytes) {
  const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const int k_per_tap = p.k_blocks_per_tap * GEMM_BLOCK_K;
  const int n_atoms = p.block_n / 64;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const TileCoord t = decode_tile(p, tile);
    const int bb0 = p.b_batched ? t.b0 : 0, bb1 = p.b_batched ? t.b1 : 0;
    const int kb_begin = p.k_splits > 1 ? t.z * p.kb_per_split : 0;
    const int kb_end = p.k_splits > 1 ? min(p.num_k_blocks, kb_begin + p.kb_per_split) : p.num_k_blocks;
    const int img = p.k_splits > 1 ? 0 : t.z;  // split-K: z is the split, there is one image
    const int ab0 = p.k_splits > 1 ? 0 : t.b0, ab1 = p.k_splits > 1 ? 0 : t.b1;  // ... and one (unbatched) A
    int tap = 0, kk = 0, ty = -1, tx = -1;     // running tap index, k offset inside the tap, tap offset (dy, dx)
    if (AM == OP_CONV) {
      tap = kb_begin / p.k_blocks_per_tap;
      kk = (kb_begin - tap * p.k_blocks_per_tap) * GEMM_BLOCK_K;
      if (p.num_taps == 9) {
        ty = tap / 3 - 1;
        tx = tap % 3 - 1;
      } else {
        ty = tx = 0;
      }
    } else {
      kk = kb_begin * GEMM_BLOCK_K;
    }
    int b_row = t.n0 + tap * p.b_tap_rows;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      const uint32_t fb = full0 + stage * 8, sa = smem0 + stage * stage_bytes, sb = sa + A_TILE_BYTES;
      mbar_wait_s(empty0 + stage * 8, phase ^ 1);
      mbar_arrive_expect_tx_s(fb, stage_bytes);
      if (AM == OP_KMAJOR) {
        tma_load_4d_s(&p.tma_a, fb, sa, kk, t.m0, ab0, ab1);
      } else if (AM == OP_MNMAJOR) {
        tma_load_4d_s(&p.tma_a, fb, sa, t.m0, kk, ab0, ab1);
        tma_load_4d_s(&p.tma_a, fb, sa + MN_ATOM_BYTES, t.m0 + 64, kk, ab0, ab1);
      } else {
        tma_load_4d_s(&p.tma_a, fb, sa, kk, t.w0 + tx, t.h0 + ty, img);
      }
      if (BM == OP_KMAJOR) {
        tma_load_4d_s(&p.tma_b, fb, sb, kk, b_row, bb0, bb1);
      } else {
        for (int j = 0; j < n_atoms; ++j)
          tma_load_4d_s(&p.tma_b, fb, sb + j * MN_ATOM_BYTES, t.n0 + 64 * j, kk, bb0, bb1);
      }
      kk += GEMM_BLOCK_K;
      if (AM == OP_CONV && kk == k_per_tap) {  // next filter tap
        kk = 0;
        b_row += p.b_tap_rows;
        if (++tx == 2) {
          tx = -1;
          ++ty;
        }
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

__device__ __forceinline__ void producer_loop(const GemmParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                              uint32_t stage_bytes) {
  if (p.a_mode == OP_CONV) {
    producer_tiles<OP_CONV, OP_KMAJOR>(p, smem, full_bar, empty_bar, stage_bytes);
  } else if (p.a_mode == OP_KMAJOR) {
    if (p.b_mode == OP_KMAJOR) producer_tiles<OP_KMAJOR, OP_KMAJOR>(p, smem, full_bar, empty_bar, stage_bytes);
    else producer_tiles<OP_KMAJOR, OP_MNMAJOR>(p, smem, full_bar, empty_bar, stage_bytes);
  } else {
    if (p.b_mode == OP_KMAJOR) producer_tiles<OP_MNMAJOR, OP_KMAJOR>(p, smem, full_bar, empty_bar, stage_bytes);
    else producer_tiles<OP_MNMAJOR, OP_MNMAJOR>(p, smem, full_bar, empty_bar, stage_bytes);
  }
}

// MMA issuer (one lane): tcgen05.mma over the stage ring into the double-buffered TMEM accumulator.  The shared-memory
// descriptors are built once; a stage / k-step advance is an add on their address field (bits [0,14) = address >> 4,
// and the whole ring sits below 256 KiB, so the add never carries out of the field).
__device__ __forceinline__ void mma_loop(const GemmParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                         uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint32_t tmem_base,
                                         uint32_t stage_bytes) {
  const uint32_t idesc =
      make_idesc_f16(GEMM_BLOCK_M, p.block_n, p.fmt, p.a_mode == OP_MNMAJOR, p.b_mode == OP_MNMAJOR);
  // K-major: 8-row groups 1024 B apart; K advance = 32 B inside the swizzled row.
  // MN-major: 64-wide MN atoms 8192 B apart (LBO), 8-deep K groups 1024 B apart (SBO); K advance = 2048 B.
  const uint32_t a_lbo = (p.a_mode == OP_MNMAJOR) ? MN_ATOM_BYTES : 0;
  const uint32_t b_lbo = (p.b_mode == OP_MNMAJOR) ? MN_ATOM_BYTES : 0;
  const uint64_t a_kstep = ((p.a_mode == OP_MNMAJOR) ? 2048 : 32) >> 4;
  const uint64_t b_kstep = ((p.b_mode == OP_MNMAJOR) ? 2048 : 32) >> 4;
  const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint64_t a_desc0 = make_smem_desc_sw128(smem0, a_lbo, 1024);
  const uint64_t b_desc0 = make_smem_desc_sw128(smem0 + A_TILE_BYTES, b_lbo, 1024);
  const uint64_t stage_step = stage_bytes >> 4;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.block_n);
    int n_kb = p.num_k_blocks;
    if (p.k_splits > 1) {
      const int kb_begin = (tile / tiles_mn) * p.kb_per_split;
      n_kb = min(p.num_k_blocks, kb_begin + p.kb_per_split) - kb_begin;
    }
    uint32_t acc = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait_s(full0 + stage * 8, phase);
      tc_fence_after();
      const uint64_t ad = a_desc0 + stage * stage_step, bd = b_desc0 + stage * stage_step;
      umma_f16(d_tmem, ad, bd, idesc, acc);
      umma_f16(d_tmem, ad + a_kstep, bd + b_kstep, idesc, 1u);
      umma_f16(d_tmem, ad + 2 * a_kstep, bd + 2 * b_kstep, idesc, 1u);
      umma_f16(d_tmem, ad + 3 * a_kstep, bd + 3 * b_kstep, idesc, 1u);
      acc = 1u;
      umma_commit_s(empty0 + stage * 8);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit(&tmem_full_bar[as]);
  }
}

// EPI: 0 generic epilogue, 1 fused softmax forward, 2 fused softmax backward.  HAS_IN: the generic epilogue reads an
// input tensor (residual / saved pre-activation) that is prefetched; kept out of the input-free kernel so that one
// stays lean (the prefetch registers cost the plain GEMMs ~25 %).
template <int EPI, bool HAS_IN>
__device__ __forceinline__ void gemm_tc_body(const GemmParams& p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ symbol (offset arithmetic only) so smem accesses compile to LDS/STS
  const uint32_t smem_base_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((smem_base_u32 + 1023u) & ~1023u) - smem_base_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;
  const uint32_t stage_bytes = A_TILE_BYTES + b_bytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* tbuf_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above touched only parameters, shared memory and TMEM

  if (warp == 0) {
    if (lane == 0) producer_loop(p, smem, full_bar, empty_bar, stage_bytes);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(p, smem, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base, stage_bytes);
  } else {
    // ------------------------------------------------------------ epilogue (TMEM -> regs -> smem transpose -> global)
    const int ew = warp - 2;          // 0..7
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = ew >> 2;         // two warps share a quarter and split the 32-column slabs
    uint8_t* wbase = reinterpret_cast<uint8_t*>(tbuf_base) + static_cast<size_t>(ew) * TB_WARP_BYTES;
    float* tb = reinterpret_cast<float*>(wbase);
    long long* s_off = reinterpret_cast<long long*>(wbase + TB_FLOATS * 4);
    int* s_row = reinterpret_cast<int*>(wbase + TB_FLOATS * 4 + 32 * 8);
    const bool wide = (p.out_f32 != nullptr) || (p.res_f32 != nullptr);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile(p, tile);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const RowInfo ri = row_info(p, t, q * 32 + lane);
      s_off[lane] = ri.off;
      s_row[lane] = ri.row;
      __syncwarp();
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.block_n);
      if (EPI == 1) {
        if (half == 0) epilogue_softmax_fwd(p, tb, s_off, taddr, lane);
      } else if (EPI == 2) {
        if (half == 0) epilogue_softmax_bwd(p, tb, s_off, taddr, lane);
      } else {
        const int nslab = (p.block_n + 31) / 32;
        const int col_limit = min(p.N, t.n0 + p.block_n);
        const int kind = HAS_IN ? prefetch_kind(p) : 0;
        for (int s = half; s < nslab; s += 2) {
          float v[32];
          if (wide) {
            uint4 pf[BlockMap<4>::NIT];
            if (HAS_IN) prefetch_block<4>(p, s_off, lane, t.n0 + s * 32, col_limit, kind, pf);
            load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
            stage_rows(tb, v, lane);
            epilogue_block<4>(p, tb, s_off, s_row, lane, t.n0 + s * 32, col_limit, kind, pf);
          } else {
            uint4 pf[BlockMap<8>::NIT];
            if (HAS_IN) prefetch_block<8>(p, s_off, lane, t.n0 + s * 32, col_limit, kind, pf);
            load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
            stage_rows(tb, v, lane);
            epilogue_block<8>(p, tb, s_off, s_row, lane, t.n0 + s * 32, col_limit, kind, pf);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}


// ------------------------------------------------------------------------------------------------ tensor-map epilogue
// The generic epilogue above is instruction- and latency-bound (ncu: 8 epilogue warps busy ~85 % of the time, tensor pipe
// 43 % on the plain 12608 x 3072 x 768 GEMM, 20 % with the QuickGELU-backward epilogue): transposing every 32 x 32 block
// through shared memory costs ~10 instructions per element and every input tensor is a dependent global load.
// Here each lane keeps its accumulator ROW: it writes 64 contiguous bytes (32 fp16 / 16 fp32 columns) of that row into a
// 32-row x 64-byte box in shared memory (64B-swizzled, conflict-free), and one lane hands the box to a TMA store.
// Input tensors (residual, saved pre-activation) come in through TMA loads into the same box, two blocks ahead, and the
// result overwrites them in place.  ~2 instructions per element, no dependent global loads, clipping by the tensor map.
constexpr int TE_BUF_BYTES = 2048;  // 32 rows x 64 bytes
constexpr int TE_BUFS = 4;          // boxes per epilogue warp
constexpr int TE_WARP_BYTES = TE_BUFS * TE_BUF_BYTES;
constexpr int TE_BAR_BYTES = 1024;  // barrier block (stage ring, TMEM, per-warp input barriers)

// byte offset of 16-byte chunk c (0..3) of row r (0..31) in a CU_TENSOR_MAP_SWIZZLE_64B box (chunk ^= address bits 7..8)
__device__ __forceinline__ uint32_t te_off(int r, int c) {
  return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* x) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
}

// QuickGELU and its derivative on packed fp16 pairs for the tensor-map epilogue (the fp32 form costs ~8 instructions per
// element with two MUFU ops and made the fc1 / fc2-dgrad epilogues ALU-bound: ncu showed the 8 epilogue warps pacing
// the tile loop at 1.6x the MMA time).  sigmoid(z) = 0.5 tanh(z / 2) + 0.5 with tanh.approx.f16x2: one MUFU per pair.
// The operands are already fp16 here (u is stored as fp16; the reference runs its CLIP in fp16 on CUDA, slip.py:176).
__device__ __forceinline__ __half2 tanh_h2(__half2 x) {
  uint32_t r, a = *reinterpret_cast<uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half2 sigmoid1702_h2(__half2 u) {
  const __half2 t = tanh_h2(__hmul2(u, __float2half2_rn(0.851f)));
  return __hfma2(t, __float2half2_rn(0.5f), __float2half2_rn(0.5f));
}
__device__ __forceinline__ __half2 quickgelu_h2(__half2 u) { return __hmul2(u, sigmoid1702_h2(u)); }
__device__ __forceinline__ __half2 quickgelu_grad_h2(__half2 u) {
  const __half2 s = sigmoid1702_h2(u);
  const __half2 a = __hmul2(u, __float2half2_rn(1.702f));
  const __half2 b = __hfma2(__hneg2(s), a, a);  // 1.702 u (1 - s)
  return __hfma2(s, b, s);                       // s (1 + 1.702 u (1 - s))
}

template <int MODE, bool PAIR>
__device__ __forceinline__ void tma_epilogue_loop(const GemmParams& p, uint8_t* wbuf, uint64_t* in_bar,
                                                  uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint32_t tmem_base,
                                                  int warp, int lane) {
  constexpr bool F32 = (MODE == TE_RES32);
  constexpr bool IN = (MODE == TE_GELU_BWD || MODE == TE_RES32 || MODE == TE_RES16);
  constexpr int BW = F32 ? 16 : 32;  // columns per 64-byte box row
  const int q = warp & 3;            // TMEM lane quarter
  const int half = (warp - 2) >> 2;  // the two warps of a quarter alternate over the column blocks
  uint32_t cnt = 0;                  // blocks this warp has processed: box rotation and input-barrier parity
  int it = 0;
  // work items: tiles of this CTA, or (PAIR) this CTA's half of the tile pairs of its cluster
  uint32_t rank = 0;
  int w_begin = blockIdx.x, w_step = gridDim.x, w_end = p.total_tiles, pairs_m = 1;
  if (PAIR) {
    rank = cluster_ctarank();
    w_begin = blockIdx.x >> 1;
    w_step = gridDim.x >> 1;
    pairs_m = (p.tiles_m + 1) / 2;
    w_end = pairs_m * p.tiles_n * (p.total_tiles / (p.tiles_m * p.tiles_n));
  }
  for (int w = w_begin; w < w_end; w += w_step, ++it) {
    TileCoord t;
    bool exists = true;
    if (PAIR) {
      const int tn = w % p.tiles_n, r = w / p.tiles_n;
      const int tm = 2 * (r % pairs_m) + static_cast<int>(rank);
      t = make_coord(p, tn, tm, r / pairs_m);
      exists = tm < p.tiles_m;  // odd tile count: the peer's last tile does not exist
    } else {
      t = decode_tile(p, w);
    }
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    // tensor-map coordinates of this warp's 32 rows: (col, row, b0, b1) or, for conv, (col, w, h, image)
    int c1, c2, c3;
    if (p.a_mode == OP_CONV) {
      c1 = t.w0;
      c2 = t.h0 + (q * 32) / p.tile_w;
      c3 = t.z;
    } else {
      c1 = t.m0 + q * 32;
      c2 = t.b0;
      c3 = t.b1;
    }
    // blocks with at least one column inside N; this warp takes s = half, half + 2, ...
    const int ncols = min(p.N - t.n0, p.block_n);
    const int nblk = (ncols + BW - 1) / BW;
    const int my = exists ? (nblk - half + 1) / 2 : 0;
    if (IN && lane == 0) {
      bulk_wait_group_read<1>();
      for (int j = 0; j < 2 && j < my; ++j) {
        const uint32_t b = (cnt + j) & 3;
        mbar_arrive_expect_tx(&in_bar[b], TE_BUF_BYTES);
        tma_load_4d(&p.tma_d, &in_bar[b], wbuf + b * TE_BUF_BYTES, t.n0 + (half + 2 * j) * BW, c1, c2, c3);
      }
    }
    mbar_wait(&tmem_full_bar[as], aphase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.block_n);
    for (int j = 0; j < my; ++j) {
      const int s = half + 2 * j;
      const int col = t.n0 + s * BW;
      const uint32_t k = cnt + j;
      uint8_t* buf = wbuf + ((MODE == TE_GELU) ? (k & 1) * 2 : (k & 3)) * TE_BUF_BYTES;
      if (lane == 0) {
        if (IN) {
          if (j + 2 < my) {
            const uint32_t b = (k + 2) & 3;
            bulk_wait_group_read<1>();  // the box's previous store (two blocks back) has drained
            mbar_arrive_expect_tx(&in_bar[b], TE_BUF_BYTES);
            tma_load_4d(&p.tma_d, &in_bar[b], wbuf + b * TE_BUF_BYTES, t.n0 + (s + 4) * BW, c1, c2, c3);
          }
        } else if (MODE == TE_GELU) {
          bulk_wait_group_read<1>();
        } else {
          bulk_wait_group_read<3>();
        }
      }
      __syncwarp();
      // accumulator columns of this row
      float x[BW];
      {
        uint32_t a0[16], a1[16];
        const bool second = !F32 && (p.block_n - s * BW > 16);  // block_n % 32 == 16: the last block is half wide
        tmem_ld_x16(taddr + s * BW, a0);
        if (second) tmem_ld_x16(taddr + s * BW + 16, a1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(a0[i]) * p.alpha;
        if constexpr (!F32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[16 + i] = second ? __uint_as_float(a1[i]) * p.alpha : 0.f;
        }
      }
      if (p.bias) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int i = 0; i < BW / 4; ++i) {
          const float4 b = __ldg(bp + i);
          x[4 * i] += b.x;
          x[4 * i + 1] += b.y;
          x[4 * i + 2] += b.z;
          x[4 * i + 3] += b.w;
        }
      }
      if (IN) mbar_wait(&in_bar[k & 3], (k >> 2) & 1);
      if (MODE == TE_F16) {
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(buf + te_off(lane, c)) = pack8(x + 8 * c);
      } else if (MODE == TE_GELU) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u = pack8(x + 8 * c);  // the saved pre-activation (fp16), and the activation of exactly that value
          *reinterpret_cast<uint4*>(buf + te_off(lane, c)) = u;
          __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = quickgelu_h2(h[j]);
          *reinterpret_cast<uint4*>(buf + TE_BUF_BYTES + te_off(lane, c)) = u;
        }
      } else if (MODE == TE_GELU_BWD) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4* ptr = reinterpret_cast<uint4*>(buf + te_off(lane, c));
          uint4 u = *ptr;
          uint4 g = pack8(x + 8 * c);
          const __half2* uh = reinterpret_cast<const __half2*>(&u);
          __half2* gh = reinterpret_cast<__half2*>(&g);
#pragma unroll
          for (int j = 0; j < 4; ++j) gh[j] = __hmul2(gh[j], quickgelu_grad_h2(uh[j]));
          *ptr = g;
        }
      } else if (MODE == TE_RES16) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4* ptr = reinterpret_cast<uint4*>(buf + te_off(lane, c));
          float u[8];
          unpack8(*ptr, u);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (MODE == TE_GELU_BWD) x[8 * c + i] *= quickgelu_grad(u[i]);
            else x[8 * c + i] += u[i];
          }
          *ptr = pack8(x + 8 * c);
        }
      } else {  // TE_RES32
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4* ptr = reinterpret_cast<float4*>(buf + te_off(lane, c));
          const float4 r = *ptr;
          *ptr = make_float4(x[4 * c] + r.x, x[4 * c + 1] + r.y, x[4 * c + 2] + r.z, x[4 * c + 3] + r.w);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (MODE == TE_GELU) {
          tma_store_4d(&p.tma_d, buf, col, c1, c2, c3);
          tma_store_4d(&p.tma_c, buf + TE_BUF_BYTES, col, c1, c2, c3);
        } else {
          tma_store_4d(&p.tma_c, buf, col, c1, c2, c3);
        }
        bulk_commit_group();
      }
    }
    cnt += my;
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (!PAIR || rank == 0) mbar_arrive(&tmem_empty_bar[as]);
      else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
    }
  }
  if (lane == 0) bulk_wait_group<0>();
  __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void gemm_tce_body(const GemmParams& p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((smem_base_u32 + 1023u) & ~1023u) - smem_base_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t stage_bytes = A_TILE_BYTES + static_cast<uint32_t>(p.block_n) * GEMM_BLOCK_K * 2;

  uint8_t* bar_block = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_block);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* in_bars = reinterpret_cast<uint64_t*>(bar_block + 256);  // [GEMM_EPI_WARPS][TE_BUFS]
  uint8_t* boxes = bar_block + TE_BAR_BYTES;                          // 1024-byte aligned

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    prefetch_tmap(&p.tma_c);
    if (MODE != TE_F16) prefetch_tmap(&p.tma_d);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], GEMM_EPI_WARPS);
    }
    for (int i = 0; i < GEMM_EPI_WARPS * TE_BUFS; ++i) mbar_init(&in_bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above touched only parameters, shared memory and TMEM

  if (warp == 0) {
    if (lane == 0) producer_loop(p, smem, full_bar, empty_bar, stage_bytes);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(p, smem, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base, stage_bytes);
  } else {
    const int ew = warp - 2;
    tma_epilogue_loop<MODE, false>(p, boxes + static_cast<size_t>(ew) * TE_WARP_BYTES, in_bars + ew * TE_BUFS,
                                   tmem_full_bar, tmem_empty_bar, tmem_base, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}


// ------------------------------------------------------------------------------------------------ cta_group::2
// Two CTAs of a cluster (one TPC) cooperate on a 256 x block_n tile pair: each CTA holds its own 128 rows of A and
// HALF of the B tile, one tcgen05.mma.cta_group::2 (issued by the rank-0 CTA) consumes both CTAs' shared memory and
// writes each CTA's 128 x block_n accumulator into its own TMEM.  Per k-block a CTA loads 16 KB + block_n/2 * 128 B
// instead of 16 KB + block_n * 128 B: with 128 x 256 tiles per SM the single-CTA kernel needs ~96 B/clk/SM of L2->SM
// traffic at tensor peak, more than the ~43 B/clk/SM the L2 can deliver, so it is L2-bound at < 50 % of peak.
// Pair geometry shared by the three roles of the cta_group::2 kernels.
struct PairInfo {
  uint32_t rank;
  bool leader;
  int cluster_id, num_clusters, pairs_m, total_pairs, half_n;
};
__device__ __forceinline__ PairInfo pair_info(const GemmParams& p) {
  PairInfo pi;
  pi.rank = cluster_ctarank();
  pi.leader = pi.rank == 0;
  pi.cluster_id = blockIdx.x >> 1;
  pi.num_clusters = gridDim.x >> 1;
  pi.pairs_m = (p.tiles_m + 1) / 2;
  pi.total_pairs = pi.pairs_m * p.tiles_n * (p.total_tiles / (p.tiles_m * p.tiles_n));
  pi.half_n = p.block_n / 2;
  return pi;
}

template <int AM, int BM>
__device__ __forceinline__ void producer2_tiles(const GemmParams& p, const PairInfo& pi, uint8_t* smem, uint64_t* full_bar,
                                                uint64_t* empty_bar, uint32_t stage_bytes) {
  const uint32_t rank = pi.rank;
  const bool leader = pi.leader;
  const int half_n = pi.half_n, n_atoms = half_n / 64;
  const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const int k_per_tap = p.k_blocks_per_tap * GEMM_BLOCK_K;
  int stage = 0;
  uint32_t phase = 0;
  for (int tp = pi.cluster_id; tp < pi.total_pairs; tp += pi.num_clusters) {
    const int tn = tp % p.tiles_n, r = tp / p.tiles_n;
    const TileCoord t = make_coord(p, tn, 2 * (r % pi.pairs_m) + (int)rank, r / pi.pairs_m);
    const int bb0 = p.b_batched ? t.b0 : 0, bb1 = p.b_batched ? t.b1 : 0;
    const int nb = t.n0 + (int)rank * half_n;  // this CTA's half of the B tile
    int kk = 0, ty = p.num_taps == 9 ? -1 : 0, tx = ty, b_row = nb;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      const uint32_t fb = full0 + stage * 8, sa = smem0 + stage * stage_bytes, sb = sa + A_TILE_BYTES;
      mbar_wait_s(empty0 + stage * 8, phase ^ 1);
      if (leader) mbar_arrive_expect_tx_s(fb, 2 * stage_bytes);
      if (AM == OP_KMAJOR) {
        tma_load_4d_2sm_s(&p.tma_a, fb, sa, kk, t.m0, t.b0, t.b1);
      } else if (AM == OP_MNMAJOR) {
        tma_load_4d_2sm_s(&p.tma_a, fb, sa, t.m0, kk, t.b0, t.b1);
        tma_load_4d_2sm_s(&p.tma_a, fb, sa + MN_ATOM_BYTES, t.m0 + 64, kk, t.b0, t.b1);
      } else {
        tma_load_4d_2sm_s(&p.tma_a, fb, sa, kk, t.w0 + tx, t.h0 + ty, t.z);
      }
      if (BM == OP_KMAJOR) {
        tma_load_4d_2sm_s(&p.tma_b, fb, sb, kk, b_row, bb0, bb1);
      } else {
        for (int j = 0; j < n_atoms; ++j)
          tma_load_4d_2sm_s(&p.tma_b, fb, sb + j * MN_ATOM_BYTES, nb + 64 * j, kk, bb0, bb1);
      }
      kk += GEMM_BLOCK_K;
      if (AM == OP_CONV && kk == k_per_tap) {
        kk = 0;
        b_row += p.b_tap_rows;
        if (++tx == 2) {
          tx = -1;
          ++ty;
        }
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

__device__ __forceinline__ void producer2_loop(const GemmParams& p, const PairInfo& pi, uint8_t* smem, uint64_t* full_bar,
                                               uint64_t* empty_bar, uint32_t stage_bytes) {
  if (p.a_mode == OP_CONV) {
    producer2_tiles<OP_CONV, OP_KMAJOR>(p, pi, smem, full_bar, empty_bar, stage_bytes);
  } else if (p.a_mode == OP_KMAJOR) {
    if (p.b_mode == OP_KMAJOR) producer2_tiles<OP_KMAJOR, OP_KMAJOR>(p, pi, smem, full_bar, empty_bar, stage_bytes);
    else producer2_tiles<OP_KMAJOR, OP_MNMAJOR>(p, pi, smem, full_bar, empty_bar, stage_bytes);
  } else {
    if (p.b_mode == OP_KMAJOR) producer2_tiles<OP_MNMAJOR, OP_KMAJOR>(p, pi, smem, full_bar, empty_bar, stage_bytes);
    else producer2_tiles<OP_MNMAJOR, OP_MNMAJOR>(p, pi, smem, full_bar, empty_bar, stage_bytes);
  }
}

__device__ __forceinline__ void mma2_loop(const GemmParams& p, const PairInfo& pi, uint8_t* smem, uint64_t* full_bar,
                                          uint64_t* empty_bar, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                          uint32_t tmem_base, uint32_t stage_bytes) {
  const int cluster_id = pi.cluster_id, num_clusters = pi.num_clusters, total_pairs = pi.total_pairs;
  const uint32_t idesc = make_idesc_f16(2 * GEMM_BLOCK_M, p.block_n, p.fmt, p.a_mode == OP_MNMAJOR,
                                        p.b_mode == OP_MNMAJOR);
  const uint32_t a_lbo = (p.a_mode == OP_MNMAJOR) ? MN_ATOM_BYTES : 0;
  const uint32_t b_lbo = (p.b_mode == OP_MNMAJOR) ? MN_ATOM_BYTES : 0;
  const uint64_t a_kstep = ((p.a_mode == OP_MNMAJOR) ? 2048 : 32) >> 4;
  const uint64_t b_kstep = ((p.b_mode == OP_MNMAJOR) ? 2048 : 32) >> 4;
  const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint64_t a_desc0 = make_smem_desc_sw128(smem0, a_lbo, 1024);
  const uint64_t b_desc0 = make_smem_desc_sw128(smem0 + A_TILE_BYTES, b_lbo, 1024);
  const uint64_t stage_step = stage_bytes >> 4;
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  for (int tp = cluster_id; tp < total_pairs; tp += num_clusters, ++it) {
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.block_n);
    uint32_t acc = 0;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      mbar_wait_s(full0 + stage * 8, phase);
      tc_fence_after();
      const uint64_t ad = a_desc0 + stage * stage_step, bd = b_desc0 + stage * stage_step;
      umma_f16_2sm(d_tmem, ad, bd, idesc, acc);
      umma_f16_2sm(d_tmem, ad + a_kstep, bd + b_kstep, idesc, 1u);
      umma_f16_2sm(d_tmem, ad + 2 * a_kstep, bd + 2 * b_kstep, idesc, 1u);
      umma_f16_2sm(d_tmem, ad + 3 * a_kstep, bd + 3 * b_kstep, idesc, 1u);
      acc = 1u;
      umma_commit_2sm_s(empty0 + stage * 8, 3);  // frees the smem slot in both CTAs
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit_2sm(&tmem_full_bar[as], 3);
  }
}

template <bool HAS_IN>
__device__ __forceinline__ void gemm_tc2_body(const GemmParams& p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((smem_base_u32 + 1023u) & ~1023u) - smem_base_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int half_n = p.block_n / 2;
  const uint32_t b_bytes = static_cast<uint32_t>(half_n) * GEMM_BLOCK_K * 2;
  const uint32_t stage_bytes = A_TILE_BYTES + b_bytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* tbuf_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's copy is the live one: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[i], 1);  // one multicast tcgen05.commit per CTA
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * GEMM_EPI_WARPS);  // leader's copy: epilogue warps of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  const PairInfo pinfo = pair_info(p);
  const int cluster_id = pinfo.cluster_id, num_clusters = pinfo.num_clusters, pairs_m = pinfo.pairs_m,
            total_pairs = pinfo.total_pairs;

  if (warp == 0) {
    if (lane == 0) producer2_loop(p, pinfo, smem, full_bar, empty_bar, stage_bytes);
  } else if (warp == 1) {
    if (lane == 0 && leader)
      mma2_loop(p, pinfo, smem, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base, stage_bytes);
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* wbase = reinterpret_cast<uint8_t*>(tbuf_base) + static_cast<size_t>(ew) * TB_WARP_BYTES;
    float* tb = reinterpret_cast<float*>(wbase);
    long long* s_off = reinterpret_cast<long long*>(wbase + TB_FLOATS * 4);
    int* s_row = reinterpret_cast<int*>(wbase + TB_FLOATS * 4 + 32 * 8);
    const bool wide = (p.out_f32 != nullptr) || (p.res_f32 != nullptr);
    int it = 0;
    for (int tp = cluster_id; tp < total_pairs; tp += num_clusters, ++it) {
      const int tn = tp % p.tiles_n, r = tp / p.tiles_n;
      const int tm = 2 * (r % pairs_m) + (int)rank;
      const TileCoord t = make_coord(p, tn, tm, r / pairs_m);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      RowInfo ri = row_info(p, t, q * 32 + lane);
      if (tm >= p.tiles_m) ri.off = -1;  // odd tile count: the peer's last tile does not exist
      s_off[lane] = ri.off;
      s_row[lane] = ri.row;
      __syncwarp();
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.block_n);
      const int nslab = (p.block_n + 31) / 32;
      const int col_limit = min(p.N, t.n0 + p.block_n);
      const int kind = HAS_IN ? prefetch_kind(p) : 0;
      for (int s = half; s < nslab; s += 2) {
        float v[32];
        if (wide) {
          uint4 pf[BlockMap<4>::NIT];
          if (HAS_IN) prefetch_block<4>(p, s_off, lane, t.n0 + s * 32, col_limit, kind, pf);
          load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
          stage_rows(tb, v, lane);
          epilogue_block<4>(p, tb, s_off, s_row, lane, t.n0 + s * 32, col_limit, kind, pf);
        } else {
          uint4 pf[BlockMap<8>::NIT];
          if (HAS_IN) prefetch_block<8>(p, s_off, lane, t.n0 + s * 32, col_limit, kind, pf);
          load_acc(taddr + s * 32, min(32, p.block_n - s * 32), v);
          stage_rows(tb, v, lane);
          epilogue_block<8>(p, tb, s_off, s_row, lane, t.n0 + s * 32, col_limit, kind, pf);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[as]);
        else mbar_arrive_cluster(&tmem_empty_bar[as], 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's smem: nobody leaves before everything is consumed
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// CTA pairs with the tensor-map epilogue: the pair halves the L2 -> SM operand traffic per FLOP (the single-CTA 128 x 256
// tile is capped near 45 % of tensor peak by the ~43 B/clk/SM the L2 delivers), the epilogue keeps out of its way.
template <int MODE>
__device__ __forceinline__ void gemm_tce2_body(const GemmParams& p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((smem_base_u32 + 1023u) & ~1023u) - smem_base_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const PairInfo pinfo = pair_info(p);
  const uint32_t stage_bytes = A_TILE_BYTES + static_cast<uint32_t>(pinfo.half_n) * GEMM_BLOCK_K * 2;

  uint8_t* bar_block = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_block);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* in_bars = reinterpret_cast<uint64_t*>(bar_block + 256);
  uint8_t* boxes = bar_block + TE_BAR_BYTES;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    prefetch_tmap(&p.tma_c);
    if (MODE != TE_F16) prefetch_tmap(&p.tma_d);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * GEMM_EPI_WARPS);
    }
    for (int i = 0; i < GEMM_EPI_WARPS * TE_BUFS; ++i) mbar_init(&in_bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) producer2_loop(p, pinfo, smem, full_bar, empty_bar, stage_bytes);
  } else if (warp == 1) {
    if (lane == 0 && pinfo.leader)
      mma2_loop(p, pinfo, smem, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, tmem_base, stage_bytes);
  } else {
    const int ew = warp - 2;
    tma_epilogue_loop<MODE, true>(p, boxes + static_cast<size_t>(ew) * TE_WARP_BYTES, in_bars + ew * TE_BUFS,
                                  tmem_full_bar, tmem_empty_bar, tmem_base, warp, lane);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc2_body<false>(p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc2_in_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc2_body<true>(p);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc_body<0, false>(p);
}
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_in_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc_body<0, true>(p);
}
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_softmax_fwd_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc_body<1, false>(p);
}
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_softmax_bwd_kernel(const __grid_constant__ GemmParams p) {
  gemm_tc_body<2, false>(p);
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tce_kernel(const __grid_constant__ GemmParams p) {
  gemm_tce_body<MODE>(p);
}
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tce2_kernel(const __grid_constant__ GemmParams p) {
  gemm_tce2_body<MODE>(p);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

void set_err(char* err, int errlen, const char* msg) {
  if (err && errlen > 0) snprintf(err, errlen, "%s", msg);
}

// 4-D tensor map, zero fill out of bounds.  fmt 0 = fp16, 1 = bf16, 2 = fp32; operands use the 128-byte swizzle, the
// epilogue boxes (64-byte rows) the 64-byte swizzle.
int encode_tmap4(CUtensorMap* out, const void* ptr, int fmt, const uint64_t dims[4], const uint64_t strides_elems[3],
                 const uint32_t box[4], char* err, int errlen, bool swizzle64 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_err(err, errlen, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -1;
  }
  cuuint64_t gdims[4], gstrides[3];
  cuuint32_t gbox[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
  }
  for (int i = 0; i < 3; ++i) {
    gstrides[i] = strides_elems[i] * (fmt == 2 ? 4 : 2);
    if (gstrides[i] % 16 != 0 || gstrides[i] == 0) {
      set_err(err, errlen, "TMA global stride must be a non-zero multiple of 16 bytes");
      return -2;
    }
  }
  if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0) {
    set_err(err, errlen, "TMA global address must be 16-byte aligned");
    return -3;
  }
  const CUtensorMapDataType dt = fmt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                          : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUresult r = fn(out, dt, 4, const_cast<void*>(ptr), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    set_err(err, errlen, buf);
    return -4;
  }
  return 0;
}

int encode_operand(CUtensorMap* out, const GemmOperand& op, int rows_box, int fmt, char* err, int errlen) {
  // batch strides for extent-1 dims still need to be valid multiples of 16 bytes
  long long bs0 = op.nb0 > 1 ? op.bs0 : 0, bs1 = op.nb1 > 1 ? op.bs1 : 0;
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  if (op.mode == OP_KMAJOR) {
    dims[0] = op.k_extent;
    dims[1] = op.mn_extent;
    box[0] = GEMM_BLOCK_K;
    box[1] = rows_box;
  } else {
    dims[0] = op.mn_extent;
    dims[1] = op.k_extent;
    box[0] = 64;
    box[1] = GEMM_BLOCK_K;
  }
  dims[2] = op.nb0;
  dims[3] = op.nb1;
  box[2] = box[3] = 1;
  strides[0] = op.ld;
  strides[1] = bs0 ? bs0 : op.ld * dims[1];
  strides[2] = bs1 ? bs1 : strides[1] * dims[2];
  return encode_tmap4(out, op.ptr, fmt, dims, strides, box, err, errlen);
}

bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace

int tmap_encode_4d(CUtensorMap* out, const void* ptr, int fmt, const uint64_t dims[4], const uint64_t strides_elems[3],
                   const uint32_t box[4], char* err, int errlen) {
  return encode_tmap4(out, ptr, fmt, dims, strides_elems, box, err, errlen);
}

// Measured on B200 (tests/test_gemm_gpu.py::test_gemm_throughput_report): CTA pairs are within +-8% of single-CTA tiles
// on the CLIP GEMM shapes (ahead only at K = 3072), so single-CTA tiles stay the default; PXR_GEMM_CTA_GROUP=2 turns
// pairs on wherever the tile shape allows it.
int gemm_default_cta_group() {
  static const int env_cg = [] {
    const char* e = getenv("PXR_GEMM_CTA_GROUP");
    return e && atoi(e) == 2 ? 2 : 1;
  }();
  return env_cg;
}

namespace {
// CTA pairs need (see gemm_tc2_body): a generic epilogue, an even split of the B tile, at least two M tiles.
int decide_cta_group(const GemmParams& p, const GemmEpilogue& epi, int block_n, int num_sms) {
  const bool softmax_epi = epi.act == ACT_SOFTMAX || epi.act == ACT_SOFTMAX_BWD;
  const bool can_pair = p.k_splits <= 1 && !softmax_epi && block_n % 32 == 0 && (p.b_mode != OP_MNMAJOR || block_n % 128 == 0) &&
                        p.tiles_m >= 2 && num_sms >= 2;
  const int want = epi.cta_group ? epi.cta_group : gemm_default_cta_group();
  return (want == 1 || !can_pair) ? 1 : 2;
}

// Tensor-map epilogue (see tma_epilogue_loop) when the epilogue is one of its shapes and every tensor can be described
// by a tensor map; anything else keeps the generic epilogue.  PXR_GEMM_TMA_EPI=0 disables it globally.
int decide_tma_epi(const GemmParams& p, const GemmEpilogue& e, int block_n) {
  static const bool enabled = [] {
    const char* v = getenv("PXR_GEMM_TMA_EPI");
    return !(v && atoi(v) == 0);
  }();
  if (!enabled || e.tma_epi < 0 || e.bias_per_row || p.fmt != 0 || p.k_splits > 1) return TE_NONE;
  int mode = TE_NONE;
  const bool no_res = !e.res_f32 && !e.res_f16;
  if (e.act == ACT_NONE && e.out_f16 && !e.out_f32 && !e.res_f32 && !e.aux_out) mode = e.res_f16 ? TE_RES16 : TE_F16;
  else if (e.act == ACT_QUICKGELU && e.aux_out && e.out_f16 && !e.out_f32 && no_res) mode = TE_GELU;
  else if (e.act == ACT_QUICKGELU_BWD && e.aux_in && e.out_f16 && !e.out_f32 && no_res) mode = TE_GELU_BWD;
  else if (e.act == ACT_NONE && e.out_f32 && !e.out_f16 && e.res_f32 && !e.res_f16 && !e.aux_out) mode = TE_RES32;
  if (mode == TE_NONE) return TE_NONE;
  const int bw = mode == TE_RES32 ? 16 : 32;
  const int eb = mode == TE_RES32 ? 4 : 2;
  if (block_n % bw && p.tiles_n != 1) return TE_NONE;  // a partial last block would spill into the next tile's columns
  if (e.bias && (!aligned16(e.bias) || p.N % bw)) return TE_NONE;
  if ((e.ldc * eb) % 16 || (e.bs0 * eb) % 16 || (e.bs1 * eb) % 16) return TE_NONE;
  if (!aligned16(e.out_f16) || !aligned16(e.out_f32) || !aligned16(e.aux_in) || !aligned16(e.aux_out) ||
      !aligned16(e.res_f32) || !aligned16(e.res_f16))
    return TE_NONE;
  if (p.a_mode == OP_CONV && (32 % p.tile_w)) return TE_NONE;
  // TMA clips the innermost dimension at 16-byte granularity: a ragged N overwrites the columns up to the next 16-byte
  // boundary (with alpha * 0 when B's out-of-range rows are zero-filled), so they must be padding inside the row pitch
  const int gran = 16 / eb;
  if (p.N % gran && (e.ldc < (p.N + gran - 1) / gran * gran || e.bias)) return TE_NONE;
  return mode;
}

int encode_epilogue_map(CUtensorMap* out, const void* ptr, bool f32, const GemmParams& p, const GemmEpilogue& e,
                        char* err, int errlen) {
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  const uint64_t ldc = static_cast<uint64_t>(e.ldc);
  dims[0] = p.N;
  box[0] = f32 ? 16 : 32;
  if (p.a_mode == OP_CONV) {
    dims[1] = p.conv_W;
    dims[2] = p.conv_H;
    dims[3] = p.nb0;
    strides[0] = ldc;
    strides[1] = ldc * p.conv_W;
    strides[2] = ldc * p.conv_W * p.conv_H;
    box[1] = p.tile_w;
    box[2] = 32 / p.tile_w;
    box[3] = 1;
  } else {
    const int nb1 = p.total_tiles / (p.tiles_m * p.tiles_n * p.nb0);
    dims[1] = p.M;
    dims[2] = p.nb0;
    dims[3] = nb1;
    strides[0] = ldc;
    strides[1] = p.nb0 > 1 ? static_cast<uint64_t>(e.bs0) : ldc * p.M;
    strides[2] = nb1 > 1 ? static_cast<uint64_t>(e.bs1) : strides[1] * p.nb0;
    box[1] = 32;
    box[2] = box[3] = 1;
  }
  return encode_tmap4(out, ptr, f32 ? 2 : 0, dims, strides, box, err, errlen, true);
}

// bytes the epilogue must move for `mn` output elements: every output written once, every input tensor read once
double epilogue_bytes(const GemmEpilogue& e, double mn) {
  double b = 0;
  if (e.out_f16) b += 2 * mn;
  if (e.out_f32) b += 4 * mn;
  if (e.aux_out) b += 2 * mn;
  if (e.aux_in) b += 2 * mn;
  if (e.res_f16) b += 2 * mn;
  if (e.res_f32) b += 4 * mn;
  return b;
}

int finish_plan(GemmPlan* plan, const GemmEpilogue& epi, int block_n, int num_sms, char* err, int errlen) {
  GemmParams& p = plan->p;
  if (block_n < 16 || block_n > 256 || block_n % 16) {
    set_err(err, errlen, "block_n must be a multiple of 16 in [16, 256]");
    return -10;
  }
  if (p.b_mode == OP_MNMAJOR && block_n % 64) {
    set_err(err, errlen, "MN-major B needs block_n % 64 == 0");
    return -11;
  }
  p.block_n = block_n;
  p.tma_epi = decide_tma_epi(p, epi, block_n);
  if (p.tma_epi != TE_NONE) {
    const bool f32 = p.tma_epi == TE_RES32;
    int rc = encode_epilogue_map(&p.tma_c, f32 ? static_cast<const void*>(epi.out_f32) : epi.out_f16, f32, p, epi, err,
                                 errlen);
    const void* second = p.tma_epi == TE_GELU       ? static_cast<const void*>(epi.aux_out)
                         : p.tma_epi == TE_GELU_BWD ? static_cast<const void*>(epi.aux_in)
                         : p.tma_epi == TE_RES32    ? static_cast<const void*>(epi.res_f32)
                         : p.tma_epi == TE_RES16    ? static_cast<const void*>(epi.res_f16)
                                                    : nullptr;
    if (!rc && second) rc = encode_epilogue_map(&p.tma_d, second, f32, p, epi, err, errlen);
    if (rc) p.tma_epi = TE_NONE;  // a stride the tensor map cannot express: generic epilogue
  }
  const int stage_bytes = A_TILE_BYTES + (p.cta_group == 2 ? block_n / 2 : block_n) * GEMM_BLOCK_K * 2;
  const int epi_smem = p.tma_epi != TE_NONE ? TE_BAR_BYTES + GEMM_EPI_WARPS * TE_WARP_BYTES
                                             : 256 + GEMM_EPI_WARPS * TB_WARP_BYTES;
  int stages = (227 * 1024 - 1024 - epi_smem) / stage_bytes;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  int cols = 32;
  while (cols < 2 * block_n) cols <<= 1;
  p.tmem_cols = cols;
  p.alpha = epi.alpha;
  p.n_store = epi.n_store > 0 ? epi.n_store : p.N;
  if ((epi.act == ACT_SOFTMAX || epi.act == ACT_SOFTMAX_BWD) && (p.tiles_n != 1 || !epi.out_f16)) {
    set_err(err, errlen, "fused softmax epilogue needs the whole row in one tile and an fp16 output");
    return -12;
  }
  p.bias = epi.bias;
  p.bias_per_row = epi.bias_per_row;
  p.act = epi.act;
  p.aux_in = epi.aux_in;
  p.aux_out = epi.aux_out;
  p.res_f32 = epi.res_f32;
  p.res_f16 = epi.res_f16;
  p.out_f32 = epi.out_f32;
  p.out_f16 = epi.out_f16;
  p.ldc = epi.ldc;
  p.out_bs0 = epi.bs0;
  p.out_bs1 = epi.bs1;
  p.vec_ok = (epi.ldc % 8 == 0) && (epi.bs0 % 8 == 0) && (epi.bs1 % 8 == 0) && aligned16(epi.aux_in) &&
             aligned16(epi.aux_out) && aligned16(epi.res_f32) && aligned16(epi.res_f16) && aligned16(epi.out_f32) &&
             aligned16(epi.out_f16);
  if (p.cta_group == 2) {
    const int batches = p.total_tiles / (p.tiles_m * p.tiles_n);
    const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n * batches;
    const int max_clusters = num_sms / 2;
    plan->grid = 2 * (pairs < max_clusters ? pairs : max_clusters);
  } else {
    plan->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  }
  plan->smem_bytes = stages * stage_bytes + 1024 + epi_smem;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tc_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tc2_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tc_softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tc_softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce_kernel<TE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce_kernel<TE_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce_kernel<TE_GELU_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce_kernel<TE_RES32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce_kernel<TE_RES16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce2_kernel<TE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce2_kernel<TE_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce2_kernel<TE_GELU_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce2_kernel<TE_RES32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tce2_kernel<TE_RES16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  return 0;
}

}  // namespace

int gemm_plan_make(GemmPlan* plan, const GemmOperand& A, const GemmOperand& B, int M, int N, int K,
                   const GemmEpilogue& epi, int block_n, int fmt, int num_sms, char* err, int errlen) {
  *plan = GemmPlan{};
  GemmParams& p = plan->p;
  if (A.mode == OP_CONV || B.mode == OP_CONV) {
    set_err(err, errlen, "use conv_plan_make for conv operands");
    return -20;
  }
  p.M = M;
  p.N = N;
  p.fmt = fmt;
  p.a_mode = A.mode;
  p.b_mode = B.mode;
  p.num_taps = 1;
  p.k_blocks_per_tap = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  p.num_k_blocks = p.k_blocks_per_tap;
  p.tiles_m = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  p.tiles_n = (N + block_n - 1) / block_n;
  const int nb0 = A.nb0, nb1 = A.nb1;
  p.nb0 = nb0;
  p.b_batched = (B.nb0 > 1 || B.nb1 > 1) ? 1 : 0;
  if (p.b_batched && (B.nb0 != nb0 || B.nb1 != nb1)) {
    set_err(err, errlen, "batched B must have the same batch extents as A");
    return -21;
  }
  p.total_tiles = p.tiles_m * p.tiles_n * nb0 * nb1;
  p.k_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  GemmEpilogue epi_s = epi;
  if (epi.k_splits > 1) {  // split-K: the batch index of a tile becomes the split (fp32 partial sums, slab z of out_f32)
    if (nb0 != 1 || nb1 != 1 || !epi.out_f32 || epi.out_f16 || epi.bias || epi.res_f16 || epi.res_f32 || epi.act != ACT_NONE) {
      set_err(err, errlen, "split-K GEMM: unbatched operands, plain fp32 partial sums only");
      return -22;
    }
    p.kb_per_split = (p.num_k_blocks + epi.k_splits - 1) / epi.k_splits;
    p.k_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.nb0 = p.k_splits;
    p.total_tiles = p.tiles_m * p.tiles_n * p.k_splits;
    epi_s.bs0 = (long long)M * epi.ldc;
  }
  int rc = encode_operand(&p.tma_a, A, GEMM_BLOCK_M, fmt, err, errlen);
  if (rc) return rc;
  p.cta_group = decide_cta_group(p, epi, block_n, num_sms);
  rc = encode_operand(&p.tma_b, B, p.cta_group == 2 ? block_n / 2 : block_n, fmt, err, errlen);
  if (rc) return rc;
  plan->flops = 2.0 * M * N * K * nb0 * nb1;
  {
    const double nb = (double)nb0 * nb1, mn = (double)M * N * nb;
    plan->bytes = 2.0 * M * K * nb + 2.0 * N * K * (p.b_batched ? nb : 1.0) + epilogue_bytes(epi, mn);
  }
  return finish_plan(plan, epi_s, block_n, num_sms, err, errlen);
}

int conv_plan_make(GemmPlan* plan, const void* in, long long in_ld, int batch, int H, int W, int c_in, const void* wt,
                   int cout_pad, int n_out, int ksize, const GemmEpilogue& epi, int block_n, int fmt, int num_sms,
                   char* err, int errlen) {
  *plan = GemmPlan{};
  GemmParams& p = plan->p;
  if (c_in % 64 || (ksize != 1 && ksize != 3)) {
    set_err(err, errlen, "conv: c_in must be a multiple of 64 and ksize 1 or 3");
    return -30;
  }
  if (W < 1 || H < 1) {
    set_err(err, errlen, "conv: empty image");
    return -31;
  }
  // M tile = tile_h x tile_w pixels (8 x 16 or 16 x 8).  Any W works: columns past the image are zero-filled by the
  // TMA loads and clipped by the stores (tensor-map epilogue) / predicated off (generic epilogue); pick the width that
  // wastes the fewest columns (pixray's presets give latents such as 9 x 9, 18 x 18, 12 x 6: pixray.py:1864-1878)
  const int waste16 = (W + 15) / 16 * 16, waste8 = (W + 7) / 8 * 8;
  int tile_w = (waste16 <= waste8) ? 16 : 8;
  int tile_h = GEMM_BLOCK_M / tile_w;
  p.M = H * W;
  p.N = n_out;
  p.fmt = fmt;
  p.a_mode = OP_CONV;
  p.b_mode = OP_KMAJOR;
  p.num_taps = ksize * ksize;
  p.k_blocks_per_tap = c_in / 64;
  p.num_k_blocks = p.num_taps * p.k_blocks_per_tap;
  p.conv_H = H;
  p.conv_W = W;
  p.tile_h = tile_h;
  p.tile_w = tile_w;
  p.tiles_w = (W + tile_w - 1) / tile_w;
  p.tiles_m = p.tiles_w * ((H + tile_h - 1) / tile_h);
  p.tiles_n = (n_out + block_n - 1) / block_n;
  p.nb0 = batch;
  p.b_batched = 0;
  p.b_tap_rows = cout_pad;
  p.total_tiles = p.tiles_m * p.tiles_n * batch;
  p.k_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  if (epi.k_splits > 1) {
    if (batch != 1 || !epi.out_f32 || epi.out_f16 || epi.bias || epi.res_f16 || epi.res_f32 || epi.act != ACT_NONE) {
      set_err(err, errlen, "split-K conv: one image, plain fp32 partial sums only");
      return -32;
    }
    p.kb_per_split = (p.num_k_blocks + epi.k_splits - 1) / epi.k_splits;
    p.k_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.nb0 = p.k_splits;  // decode_tile: z = split
    p.total_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  }
  {
    uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)W, (uint64_t)H, (uint64_t)batch};
    uint64_t strides[3] = {(uint64_t)in_ld, (uint64_t)in_ld * W, (uint64_t)in_ld * W * H};
    uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
    int rc = encode_tmap4(&p.tma_a, in, fmt, dims, strides, box, err, errlen);
    if (rc) return rc;
  }
  {
    GemmOperand B;
    B.ptr = wt;
    B.mode = OP_KMAJOR;
    B.ld = c_in;
    B.mn_extent = (long long)p.num_taps * cout_pad;
    B.k_extent = c_in;
    p.cta_group = decide_cta_group(p, epi, block_n, num_sms);
    int rc = encode_operand(&p.tma_b, B, p.cta_group == 2 ? block_n / 2 : block_n, fmt, err, errlen);
    if (rc) return rc;
  }
  plan->flops = 2.0 * H * W * (double)n_out * c_in * p.num_taps * batch;
  plan->bytes = 2.0 * H * W * (double)c_in * batch + 2.0 * (double)n_out * c_in * p.num_taps +
                epilogue_bytes(epi, (double)H * W * n_out * batch * (epi.k_splits > 1 ? epi.k_splits : 1));
  GemmEpilogue e = epi;
  return finish_plan(plan, e, block_n, num_sms, err, errlen);
}

namespace {
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long pixels,
                                                            int N, int ld_ws, const float* __restrict__ bias,
                                                            const __half* __restrict__ res, __half* __restrict__ out,
                                                            int ld_out) {
  pdl_prologue();
  const int vecs = N / 8;
  const long long n = pixels * vecs;
  const size_t slab = static_cast<size_t>(pixels) * ld_ws;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const long long px = i / vecs;
    const int c0 = static_cast<int>(i % vecs) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[c0 + j] : 0.f;
    const float* src = ws + static_cast<size_t>(px) * ld_ws + c0;
    uint4 r16 = make_uint4(0, 0, 0, 0);
    if (res) r16 = *reinterpret_cast<const uint4*>(res + static_cast<size_t>(px) * ld_out + c0);
    // the partial sums are independent loads: issue four splits' worth before the (fixed-order) adds, so the loop pays
    // one L2 round trip per four splits instead of one per split
    int s0 = 0;
    for (; s0 + 4 <= splits; s0 += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldcg(reinterpret_cast<const float4*>(src + (s0 + u) * slab));
        b[u] = __ldcg(reinterpret_cast<const float4*>(src + (s0 + u) * slab + 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // fixed order: deterministic
        acc[0] += a[u].x;
        acc[1] += a[u].y;
        acc[2] += a[u].z;
        acc[3] += a[u].w;
        acc[4] += b[u].x;
        acc[5] += b[u].y;
        acc[6] += b[u].z;
        acc[7] += b[u].w;
      }
    }
    for (; s0 < splits; ++s0) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(src + s0 * slab)), b = __ldcg(reinterpret_cast<const float4*>(src + s0 * slab + 4));
      acc[0] += a.x;
      acc[1] += a.y;
      acc[2] += a.z;
      acc[3] += a.w;
      acc[4] += b.x;
      acc[5] += b.y;
      acc[6] += b.z;
      acc[7] += b.w;
    }
    if (res) {
      float r[8];
      unpack8(r16, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += r[j];
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(px) * ld_out + c0) = pack8(acc);
  }
}
}  // namespace

void splitk_reduce(const float* ws, int splits, long long pixels, int N, int ld_ws, const float* bias,
                   const __half* res, __half* out, int ld_out, cudaStream_t stream) {
  const long long n = pixels * (N / 8);
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  launch_pdl(splitk_reduce_kernel, dim3(static_cast<int>(g)), dim3(256), 0, stream, ws, splits, pixels, N, ld_ws, bias, res, out, ld_out);
}

#define PXR_LAUNCH_TCE2(MODE) \
  launch_pdl(gemm_tce2_kernel<MODE>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p)

void gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  if (plan.p.tma_epi != TE_NONE && plan.p.cta_group == 2) {
    switch (plan.p.tma_epi) {
      case TE_F16: PXR_LAUNCH_TCE2(TE_F16); break;
      case TE_GELU: PXR_LAUNCH_TCE2(TE_GELU); break;
      case TE_GELU_BWD: PXR_LAUNCH_TCE2(TE_GELU_BWD); break;
      case TE_RES32: PXR_LAUNCH_TCE2(TE_RES32); break;
      default: PXR_LAUNCH_TCE2(TE_RES16); break;
    }
    return;
  }
  if (plan.p.tma_epi == TE_F16)
    launch_pdl(gemm_tce_kernel<TE_F16>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.tma_epi == TE_GELU)
    launch_pdl(gemm_tce_kernel<TE_GELU>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.tma_epi == TE_GELU_BWD)
    launch_pdl(gemm_tce_kernel<TE_GELU_BWD>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.tma_epi == TE_RES32)
    launch_pdl(gemm_tce_kernel<TE_RES32>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.tma_epi == TE_RES16)
    launch_pdl(gemm_tce_kernel<TE_RES16>, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.act == ACT_SOFTMAX)
    launch_pdl(gemm_tc_softmax_fwd_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.act == ACT_SOFTMAX_BWD)
    launch_pdl(gemm_tc_softmax_bwd_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.cta_group == 2 && (plan.p.res_f32 || plan.p.res_f16 || plan.p.act == ACT_QUICKGELU_BWD))
    launch_pdl(gemm_tc2_in_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.cta_group == 2)
    launch_pdl(gemm_tc2_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else if (plan.p.res_f32 || plan.p.res_f16 || plan.p.act == ACT_QUICKGELU_BWD)
    launch_pdl(gemm_tc_in_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
  else
    launch_pdl(gemm_tc_kernel, dim3(plan.grid), dim3(GEMM_THREADS), plan.smem_bytes, stream, plan.p);
}

}  // namespace pxr
