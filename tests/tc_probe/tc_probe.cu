// Standalone probe of the sm_100a primitives in seq2squiggle_b200/csrc/tc_prims.cuh (TEST INFRASTRUCTURE).
// Each case launches one CTA with bounded barrier waits, compares against a host fp32 product of the same
// fp16-rounded operands and prints one PASS/FAIL line.  Run on the GPU box:  tests/tc_probe/tc_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../seq2squiggle_b200/csrc/tc_host.h"
#include "../../seq2squiggle_b200/csrc/tc_prims.cuh"

using namespace s2s::tc;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

struct Params {
  int n;        // MMA N (multiple of 16, <= 256)
  int ka, kb;   // K extent (elements) of the A / B global matrices (multiples of 64)
  int a0, b0;   // first 16-element k-step used from A / B
  int steps;    // number of k-steps
  int ts;       // 1: A operand from TMEM
  int tma;      // 1: operands staged by TMA (SS only), 0: generic-proxy swizzled stores
};

// ---- P1: TMEM store / load lane+column mapping --------------------------------------------------
__global__ void __launch_bounds__(128) k_tmem_roundtrip(uint32_t* out, int* status) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<64>(&s_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = s_base;
  uint32_t v[16];
  for (int c0 = 0; c0 < 64; c0 += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (threadIdx.x << 16) | (c0 + i);
    tmem_st_32x16(tmem_addr(base, warp * 32, c0), v);
  }
  tmem_wait_st();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32(tmem_addr(base, warp * 32, c0), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[threadIdx.x * 64 + c0 + i] = r[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(base);
  (void)lane; (void)status;
}

// ---- P2..P5: one 128 x N x (16*steps) UMMA ------------------------------------------------------
__global__ void __launch_bounds__(128) k_mma_probe(const __half* __restrict__ A, const __half* __restrict__ B,
                                                   float* __restrict__ D, Params p,
                                                   const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmB, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int slabsA = p.ka / 64, slabsB = p.kb / 64;
  uint8_t* sA = smem;                                  // slabsA x [128 x 128 B]
  uint8_t* sB = smem + (size_t)slabsA * 128 * 128;     // slabsB x [n x 128 B]
  if (warp == 0) tmem_alloc<512>(&s_base);
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = s_base;
  const uint32_t d_col = 0, a_col = 256;  // D: columns [0,n); TMEM A operand: columns [256, 256+ka/2)
  bool ok = true;

  if (p.tma) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_load, (uint32_t)(slabsA * 128 * 128 + slabsB * p.n * 128));
      for (int s = 0; s < slabsA; ++s) tma_load_2d(sA + (size_t)s * 128 * 128, &tmA, &bar_load, s * 64, 0);
      for (int s = 0; s < slabsB; ++s) tma_load_2d(sB + (size_t)s * p.n * 128, &tmB, &bar_load, s * 64, 0);
    }
    ok = mbar_wait(&bar_load, 0, status, 101);
  } else {
    if (!p.ts) {
      for (int i = tid; i < 128 * (p.ka / 8); i += 128) {  // 16-byte chunks of A
        int row = i / (p.ka / 8), ck = i % (p.ka / 8);
        int slab = ck / 8, c = ck % 8;
        uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)row * p.ka + ck * 8);
        *reinterpret_cast<uint4*>(sA + (size_t)slab * 128 * 128 + sw128_offset(row, c)) = v;
      }
    }
    for (int i = tid; i < p.n * (p.kb / 8); i += 128) {
      int row = i / (p.kb / 8), ck = i % (p.kb / 8);
      int slab = ck / 8, c = ck % 8;
      uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)row * p.kb + ck * 8);
      *reinterpret_cast<uint4*>(sB + (size_t)slab * p.n * 128 + sw128_offset(row, c)) = v;
    }
    fence_proxy_async_smem();
  }
  if (p.ts) {  // thread = row: pack pairs of consecutive K into 32-bit columns
    for (int c0 = 0; c0 < p.ka / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<const uint32_t*>(A + (size_t)tid * p.ka + 2 * (c0 + i));
      tmem_st_32x16(tmem_addr(base, warp * 32, a_col + c0), v);
    }
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (tid == 0 && ok) {
    const uint32_t idesc = umma_idesc(128, p.n, kFmtF16);
    for (int i = 0; i < p.steps; ++i) {
      const int sb = p.b0 + i;
      const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sB) + (sb / 4) * p.n * 128 + (sb % 4) * 32);
      if (p.ts) {
        umma_f16_ts(tmem_addr(base, 0, d_col), tmem_addr(base, 0, a_col + (p.a0 + i) * 8), bdesc, idesc, i > 0);
      } else {
        const int sa = p.a0 + i;
        const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA) + (sa / 4) * 128 * 128 + (sa % 4) * 32);
        umma_f16_ss(tmem_addr(base, 0, d_col), adesc, bdesc, idesc, i > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  if (ok) ok = mbar_wait(&bar_mma, 0, status, 102);
  tcgen05_fence_after();
  if (ok) {
    for (int c0 = 0; c0 < p.n; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_addr(base, warp * 32, d_col + c0), r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[(size_t)tid * p.n + c0 + i] = __uint_as_float(r[i]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(base);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run_case(const char* name, Params p, EncodeTiledFn enc) {
  std::vector<__half> hA((size_t)128 * p.ka), hB((size_t)p.n * p.kb);
  for (auto& x : hA) x = __float2half_rn(frand());
  for (auto& x : hB) x = __float2half_rn(frand());
  __half *dA, *dB;
  float* dD;
  int* dS;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)128 * p.n * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)128 * p.n * 4));
  CK(cudaMemset(dS, 0, 4));
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  if (p.tma) {
    if (!enc || !make_tmap_2d(enc, &tmA, dA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 128, p.ka, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_tmap_2d(enc, &tmB, dB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p.n, p.kb, p.n, 64, CU_TENSOR_MAP_SWIZZLE_128B)) {
      printf("FAIL %-28s tensor map encode failed\n", name);
      return 1;
    }
  }
  size_t smem = (size_t)(p.ka / 64) * 128 * 128 + (size_t)(p.kb / 64) * p.n * 128 + 1024;
  CK(cudaFuncSetAttribute(k_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mma_probe<<<1, 128, smem>>>(dA, dB, dD, p, tmA, tmB, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  std::vector<float> hD((size_t)128 * p.n);
  if (e == cudaSuccess) {
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  }
  double max_err = 0, max_ref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < p.n; ++n) {
      double ref = 0;
      for (int i = 0; i < p.steps; ++i)
        for (int e2 = 0; e2 < 16; ++e2)
          ref += (double)__half2float(hA[(size_t)m * p.ka + (p.a0 + i) * 16 + e2]) *
                 (double)__half2float(hB[(size_t)n * p.kb + (p.b0 + i) * 16 + e2]);
      double got = hD[(size_t)m * p.n + n];
      double err = fabs(got - ref);
      if (!(err <= max_err)) max_err = err;  // NaN-propagating max
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
  bool pass = e == cudaSuccess && st == 0 && max_err < 1e-3 * (max_ref + 1);
  printf("%s %-28s n=%d steps=%d ts=%d tma=%d  cuda=%s status=%d max_err=%.3g (max_ref=%.3g)\n", pass ? "PASS" : "FAIL",
         name, p.n, p.steps, p.ts, p.tma, cudaGetErrorString(e), st, max_err, max_ref);
  if (!pass && e == cudaSuccess && st == 0) {  // dump a corner to help decode layout mistakes
    printf("   got[0][0..7]:");
    for (int n = 0; n < 8 && n < p.n; ++n) printf(" %.4f", hD[n]);
    printf("\n   got[1][0..7]:");
    for (int n = 0; n < 8 && n < p.n; ++n) printf(" %.4f", hD[p.n + n]);
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  if (e != cudaSuccess) { cudaDeviceReset(); }
  return pass ? 0 : 1;
}

int main() {
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  EncodeTiledFn enc = get_encode_tiled();
  printf("cuTensorMapEncodeTiled entry point: %s\n", enc ? "ok" : "MISSING");
  int fails = 0;
  {  // P1
    uint32_t* d;
    int* dS;
    CK(cudaMalloc(&d, 128 * 64 * 4));
    CK(cudaMalloc(&dS, 4));
    CK(cudaMemset(d, 0, 128 * 64 * 4));
    k_tmem_roundtrip<<<1, 128>>>(d, dS);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint32_t> h(128 * 64);
    int bad = 0;
    if (e == cudaSuccess) {
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      for (int t = 0; t < 128; ++t)
        for (int c = 0; c < 64; ++c) bad += (h[t * 64 + c] != (((uint32_t)t << 16) | (uint32_t)c));
    }
    printf("%s tmem_st/ld roundtrip            cuda=%s mismatches=%d\n", (e == cudaSuccess && !bad) ? "PASS" : "FAIL",
           cudaGetErrorString(e), bad);
    fails += !(e == cudaSuccess && !bad);
    cudaFree(d); cudaFree(dS);
    if (e != cudaSuccess) cudaDeviceReset();
  }
  //                         n   ka   kb  a0 b0 steps ts tma
  fails += run_case("ss_n64_k64_manual", {64, 64, 64, 0, 0, 4, 0, 0}, enc);
  fails += run_case("ss_n64_k64_tma", {64, 64, 64, 0, 0, 4, 0, 1}, enc);
  fails += run_case("ss_n192_k64_tma (QKV)", {192, 64, 64, 0, 0, 4, 0, 1}, enc);
  fails += run_case("ss_n256_k64_tma (FFN1)", {256, 64, 64, 0, 0, 4, 0, 1}, enc);
  fails += run_case("ss_n64_k256_tma (FFN2)", {64, 256, 256, 0, 0, 16, 0, 1}, enc);
  fails += run_case("ss_n256_kstep_a1_b3 (QK^T)", {256, 64, 64, 1, 3, 1, 0, 1}, enc);
  fails += run_case("ss_n256_kstep_a3_b0 (QK^T)", {256, 64, 64, 3, 0, 1, 0, 0}, enc);
  fails += run_case("ts_n16_k256 (P.V)", {16, 256, 256, 0, 0, 16, 1, 0}, enc);
  fails += run_case("ts_n64_k256 (FFN2 from TMEM)", {64, 256, 256, 0, 0, 16, 1, 0}, enc);
  fails += run_case("ts_n16_k256_tail (a0=4)", {16, 256, 256, 4, 4, 12, 1, 0}, enc);
  printf("tc_probe: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
