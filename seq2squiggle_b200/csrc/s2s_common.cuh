// Shared device/host helpers for the s2s_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2s_b200.h"

#define S2S_L_ENC 16
#define S2S_L_DEC 250
#define S2S_L_DEC_PAD 256  // decoder rows per chunk in HBM: 250 real + 6 finite pad rows (2 M-tiles of 128)
#define S2S_D 64
#define S2S_DFF_ 256
#define S2S_H 8
#define S2S_DK 8

namespace s2s {

// ----------------------------------------------------------------------------------------------
// error handling + launch accounting
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern long long g_launch_count;

#define S2S_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      s2s::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

#define S2S_LAUNCH_CHECK()                                                                 \
  do {                                                                                     \
    ++s2s::g_launch_count;                                                                 \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      s2s::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; Salmon et al. 2011).  key = run seed, counter = (stream, index...).
// ----------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
  }
  __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ ka, n1 = lo1, n2 = hi0 ^ c3 ^ kb, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// uniform in (0,1]: never 0 so log() is safe
__host__ __device__ inline float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

#ifdef __CUDACC__
// Two standard normals from two 32-bit words (Box-Muller, fp32).
__device__ inline float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = u01(a), u2 = u01(b);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// torch.nn.Softplus(beta=1, threshold=20): x > 20 ? x : log1p(exp(x))
__device__ inline float softplus_torch(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
#endif

}  // namespace s2s
