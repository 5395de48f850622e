// Shared device/host helpers for libpsvi_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "psvi_b200.h"

// ------------------------------------------------------------------------------------------------ error plumbing
void psvi_set_error(const char* fmt, ...);

#define PSVI_CUDA_CHECK(call)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      psvi_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return PSVI_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

#define PSVI_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      psvi_set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// ------------------------------------------------------------------------------------------------ math
__device__ __forceinline__ float softplus_f(float x) {
  // F.softplus(beta=1, threshold=20): reference psvi/models/neural_net.py:131
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Counter layout used by every PSVI_NOISE_PHILOX consumer: (idx/4, sample, slab, domain), key = 64-bit seed.
// The four outputs become four standard normals (two Box-Muller pairs) for TL indices 4*(idx/4) .. +3.
struct Philox4 {
  uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;  // (a + 0.5) / 2^32 in (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
  const float r = sqrtf(-2.f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t domain, uint32_t slab, uint32_t sample,
                                               uint32_t idx4, float out[4]) {
  const Philox4 p = philox4x32_10(idx4, sample, slab, domain, (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller(p.x, p.y, out[0], out[1]);
  box_muller(p.z, p.w, out[2], out[3]);
}

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
