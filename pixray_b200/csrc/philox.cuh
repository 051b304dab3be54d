// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), host + device.  Keyed by (seed, iteration, stream) and
// counted by the GLOBAL element index, so the cutout noise is independent of how cutouts are sharded over ranks.
#pragma once
#include <stdint.h>
#include <math.h>

namespace pxr {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ inline uint32_t philox_mulhi(uint32_t a, uint32_t b) {
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
}

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = philox_mulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = philox_mulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0;
  o.v[1] = c1;
  o.v[2] = c2;
  o.v[3] = c3;
  return o;
}

__host__ __device__ inline float u32_to_unit(uint32_t x) {  // (0, 1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// uniform in (0,1] for (seed, iter, stream, index)
__host__ __device__ inline float philox_uniform(uint64_t seed, uint32_t iter, uint32_t stream, uint64_t index) {
  Philox4 r = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), iter, stream, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
  return u32_to_unit(r.v[0]);
}

// four standard normals (Box-Muller) for counter `group`
__host__ __device__ inline void philox_normal4(uint64_t seed, uint32_t iter, uint32_t stream, uint64_t group,
                                               float (&out)[4]) {
  Philox4 r = philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), iter, stream, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
  float u0 = u32_to_unit(r.v[0]), u1 = u32_to_unit(r.v[1]), u2 = u32_to_unit(r.v[2]), u3 = u32_to_unit(r.v[3]);
  float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  const float two_pi = 6.283185307179586f;
  out[0] = ra * cosf(two_pi * u1);
  out[1] = ra * sinf(two_pi * u1);
  out[2] = rb * cosf(two_pi * u3);
  out[3] = rb * sinf(two_pi * u3);
}

}  // namespace pxr
