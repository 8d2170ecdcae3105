// Micro-benchmarks of the sm_100a primitives the attention kernel is built from (TEST / DESIGN INFRASTRUCTURE).
// One CTA per launch; operands are whatever is in shared memory (timing only, no numerics).  Prints clk numbers:
//   * tcgen05.mma issue cost and issue->completion latency for the three shapes used (S: SS N=64/256 K=16,
//     PV: TS N=16 K=16, QKV: SS N=96) as a function of the batch size between commits;
//   * tcgen05.ld / st throughput per warp; MUFU.EX2 throughput for 1 and 2 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_timing mma_timing.cu ; run on the GPU box.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../seq2squiggle_b200/csrc/tc_prims.cuh"

using namespace s2s::tc;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split with the 1.5*2^23 magic constant, degree-3 minimax
// polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below the fp16 rounding of P), exponent add by LEA.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -30.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05517167f, f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// the exp pass with K of every 32 exponentials computed by ex2_poly
template <int K>
__device__ __forceinline__ void exp_step_mixed(const uint32_t (&r)[32], uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = fmaf(__uint_as_float(r[2 * i]), 0.5f, -1.f), x1 = fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, -1.f);
    const bool poly0 = ((2 * i) * K) % 32 < K, poly1 = ((2 * i + 1) * K) % 32 < K;
    pk[i] = pack_half2(poly0 ? ex2_poly(x0) : ex2f(x0), poly1 ? ex2_poly(x1) : ex2f(x1));
  }
  tmem_st_32x16(taddr, pk);
}

// 2^x for a PAIR of scores in packed fp16 arithmetic (HFMA2/HADD2 on the FMA pipe, no MUFU): x rounded to fp16,
// n = rint(x) via the 1536+15 magic constant (its low 5 mantissa bits are n+15 = the fp16 exponent field of 2^n),
// degree-3 polynomial on f = x - n, result = p * 2^n by one HMUL2.  Clamped to [-15, 16] (2^16 overflows to inf).
__device__ __forceinline__ uint32_t ex2_poly_h2(float x0, float x1) {
  const __half2 kLo = __float2half2_rn(-15.0f), kHi = __float2half2_rn(16.0f), kMagic = __float2half2_rn(1551.0f);
  __half2 x = __hmin2(__hmax2(__floats2half2_rn(x0, x1), kLo), kHi);
  const __half2 t = __hadd2(x, kMagic);
  const __half2 f = __hsub2(x, __hsub2(t, kMagic));
  __half2 p = __hfma2(__float2half2_rn(0.05517167f), f, __float2half2_rn(0.24261113f));
  p = __hfma2(p, f, __float2half2_rn(0.69326097f));
  p = __hfma2(p, f, __float2half2_rn(0.99992806f));
  const uint32_t sc = (*reinterpret_cast<const uint32_t*>(&t) << 10) & 0x7C007C00u;
  const __half2 r = __hmul2(p, *reinterpret_cast<const __half2*>(&sc));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// exp pass with K of every 16 PAIRS computed by ex2_poly_h2 (H2 = 1) instead of MUFU
template <int K>
__device__ __forceinline__ void exp_step_mixed_h2(const uint32_t (&r)[32], uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = fmaf(__uint_as_float(r[2 * i]), 0.5f, -1.f), x1 = fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, -1.f);
    pk[i] = ((i * K) % 16 < K) ? ex2_poly_h2(x0, x1) : pack_half2(ex2f(x0), ex2f(x1));
  }
  tmem_st_32x16(taddr, pk);
}

template <int K, int H2>
__device__ __forceinline__ void exp_step_sel(const uint32_t (&r)[32], uint32_t taddr) {
  if constexpr (H2) exp_step_mixed_h2<K>(r, taddr); else exp_step_mixed<K>(r, taddr);
}

template <int K, int H2 = 0>
__global__ void __launch_bounds__(512) k_exp_mixed(int reps, long long* out, float* sink) {
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&s_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // every warp owns 128 columns (4 chunks of 32) of its lane quarter: up to 4 warps per scheduler
  const uint32_t lane_addr = tmem_addr(s_base, (warp & 3) * 32, (warp >> 2) * 128);
  uint32_t ra[32], rb[32];
  __syncthreads();
  long long t0 = clock64();
  for (int k = 0; k < reps; ++k) {  // double-buffered loads like the kernel
    tmem_ld_32x32(lane_addr, ra);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      tmem_ld_32x32(lane_addr + ((c + 1) & 3) * 32, rb);
      tmem_wait_ld();   // (waits for both; the second has the whole step to land in the real kernel)
      exp_step_sel<K, H2>(ra, lane_addr + (c & 3) * 16);
      if (c + 2 < 8) tmem_ld_32x32(lane_addr + ((c + 2) & 3) * 32, ra);
      exp_step_sel<K, H2>(rb, lane_addr + ((c + 1) & 3) * 16);
    }
    tmem_wait_st();
  }
  long long t1 = clock64();
  if ((tid & 31) == 0) out[warp] = t1 - t0;
  sink[tid] = __uint_as_float(ra[0]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(s_base);
}

template <int K, int H2 = 0>
void run_exp_mixed(long long* out, float* sink) {
  long long h[16];
  for (int threads : {128, 256, 384, 512}) {
    const int reps = 200;
    k_exp_mixed<K, H2><<<1, threads>>>(reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (reps * 8);
    printf("exp pass, %2d/%d polynomial%s, %d warps/scheduler: %6.1f clk per 32-column step per warp  -> %5.2f clk per exp row per scheduler [%s]\n",
           K, H2 ? 16 : 32, H2 ? " (half2 pairs)" : "", threads / 128, per, per / 32 / (threads / 128), cudaGetErrorString(e));
  }
}

// mode 0: SS (A,B smem), mode 1: TS (A tmem).  n = MMA N.  batch MMAs then one commit, repeated `reps` times
// back to back; out[0] = clocks spent issuing, out[1] = clocks until the last commit's barrier flips.
__global__ void __launch_bounds__(128) k_mma_timing(int mode, int n, int batch, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar[64];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&s_base);
  if (tid == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_base;
  if (tid == 0) {
    const uint64_t dA = umma_desc_k_sw128(smem_u32(smem)), dB = umma_desc_k_sw128(smem_u32(smem + 32768));
    const uint32_t idesc = umma_idesc(128, n, kFmtF16);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int b = 0; b < batch; ++b) {
        if (mode == 0) umma_f16_ss(tmem + (b & 1) * 256, dA + (uint64_t)((b & 3) * 2), dB + (uint64_t)((b & 3) * 2), idesc, b > 1);
        else umma_f16_ts(tmem + 256, tmem + (b & 15) * 8, dB + (uint64_t)((b & 3) * 2), idesc, b > 0);
      }
      umma_commit(&bar[r & 63]);
    }
    long long t1 = clock64();
    uint32_t ok = 0;
    for (uint32_t i = 0; i < (1u << 24) && !ok; ++i) ok = mbar_try_wait(&bar[(reps - 1) & 63], ((reps - 1) >> 6) & 1);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    out[2] = ok;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ping-pong: issue one batch + commit, wait for it, repeat: the full round trip an unpipelined consumer sees
__global__ void __launch_bounds__(128) k_mma_roundtrip(int mode, int n, int batch, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&s_base);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_base;
  if (tid == 0) {
    const uint64_t dA = umma_desc_k_sw128(smem_u32(smem)), dB = umma_desc_k_sw128(smem_u32(smem + 32768));
    const uint32_t idesc = umma_idesc(128, n, kFmtF16);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int b = 0; b < batch; ++b) {
        if (mode == 0) umma_f16_ss(tmem, dA + (uint64_t)((b & 3) * 2), dB + (uint64_t)((b & 3) * 2), idesc, b > 0);
        else umma_f16_ts(tmem + 256, tmem + (b & 15) * 8, dB + (uint64_t)((b & 3) * 2), idesc, b > 0);
      }
      umma_commit(&bar);
      uint32_t ok = 0;
      for (uint32_t i = 0; i < (1u << 22) && !ok; ++i) ok = mbar_try_wait(&bar, r & 1);
      tcgen05_fence_after();
    }
    out[0] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// TMEM ld/st + MUFU throughput: `warps` warps per CTA, each loops `reps` times over 8 x (ld x32) [+ 32 ex2 each]
__global__ void __launch_bounds__(256) k_tmem_mufu(int what, int reps, long long* out, float* sink) {
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&s_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t lane_addr = tmem_addr(s_base, (warp & 3) * 32, (warp >> 2) * 256);
  uint32_t r[32];
  float acc = 0.f;
  for (int i = 0; i < 32; ++i) r[i] = tid * 32 + i;
  __syncthreads();
  long long t0 = clock64();
  if (what == 0) {          // ld only
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int c = 0; c < 8; ++c) { tmem_ld_32x32(lane_addr + c * 32, r); tmem_wait_ld(); acc += __uint_as_float(r[0]) + __uint_as_float(r[31]); }
    }
  } else if (what == 1) {   // st only (x16)
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = r[i];
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int c = 0; c < 8; ++c) tmem_st_32x16(lane_addr + c * 16, v);
      tmem_wait_st();
    }
  } else if (what == 2) {   // ex2 only: 256 per thread per rep
    float x[32];
    for (int i = 0; i < 32; ++i) x[i] = -0.001f * (tid + i);
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = ex2f(x[i]) - 1.0f;
    }
    for (int i = 0; i < 32; ++i) acc += x[i];
  } else if (what == 4) {   // MUFU.EX2.F16 (ex2.approx.f16x2 = two MUFU.EX2.F16): 512 exponentials per thread per rep
    uint32_t x[32];
    for (int i = 0; i < 32; ++i) x[i] = 0xB800B800u + tid + i;
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) { asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i])); x[i] ^= 0x80008000u; }
    }
    for (int i = 0; i < 32; ++i) acc += __uint_as_float(x[i]);
  } else {                  // ld + ex2 + pack + st, the exp pass of the kernel
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(lane_addr + c * 32, r);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_half2(ex2f(fmaf(__uint_as_float(r[2 * i]), 0.5f, -1.f)), ex2f(fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, -1.f)));
        tmem_st_32x16(lane_addr + c * 16, pk);
      }
      tmem_wait_st();
    }
  }
  long long t1 = clock64();
  if ((tid & 31) == 0) out[warp] = t1 - t0;
  sink[tid] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(s_base);
}

// The attention kernel's per-quarter tensor work, issued by an elected lane of warp 4 with clean uniform code:
// [4 dependent TS N=16 MMAs + commit] and [1 SS N=64 MMA + commit]; measures issue time and issue->complete latency of
// each, optionally while warps 0-3 hammer TMEM / MUFU with the exp pass (busy = 1) to expose interference.
__global__ void __launch_bounds__(160) k_quarter_pattern(int busy, int reps, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar_pv, bar_s;
  __shared__ volatile int s_stop;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 16; i += 160) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&s_base);
  if (tid == 0) { mbar_init(&bar_pv, 1); mbar_init(&bar_s, 1); fence_mbar_init(); s_stop = 0; }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_base;
  if (warp == 4) {
    const uint64_t dA = umma_desc_k_sw128(smem_u32(smem)), dB = umma_desc_k_sw128(smem_u32(smem + 32768));
    const uint32_t idesc_o = umma_idesc(128, 16, kFmtF16), idesc_s = umma_idesc(128, 64, kFmtF16);
    long long t_issue_pv = 0, t_lat_pv = 0, t_issue_s = 0, t_lat_s = 0;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16_ts(tmem + 448, tmem + 256 + 8 * ks, dB + (uint64_t)(ks * 2), idesc_o, ks > 0);
        umma_commit(&bar_pv);
      }
      long long t1 = clock64();
      while (!mbar_try_wait(&bar_pv, r & 1)) {}
      long long t2 = clock64();
      if (elect_one()) {
        umma_f16_ss(tmem + 320, dA, dB, idesc_s, 0);
        umma_commit(&bar_s);
      }
      long long t3 = clock64();
      while (!mbar_try_wait(&bar_s, r & 1)) {}
      long long t4 = clock64();
      t_issue_pv += t1 - t0; t_lat_pv += t2 - t0; t_issue_s += t3 - t2; t_lat_s += t4 - t2;
    }
    if ((tid & 31) == 0) { out[0] = t_issue_pv; out[1] = t_lat_pv; out[2] = t_issue_s; out[3] = t_lat_s; }
    s_stop = 1;
  } else if (busy) {
    const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
    uint32_t r[32];
    float acc = 0.f;
    while (!s_stop) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(lane_addr + c * 32, r);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_half2(ex2f(fmaf(__uint_as_float(r[2 * i]), 0.5f, -1.f)), ex2f(fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, -1.f)));
        tmem_st_32x16(lane_addr + c * 16, pk);
      }
      tmem_wait_st();
      acc += __uint_as_float(r[0]);
    }
    sink[tid] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 64 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(float));
  long long h[16];
  cudaFuncSetAttribute(k_mma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct { const char* name; int mode, n; } shapes[] = {{"SS N=64  K=16 (S quarter)", 0, 64}, {"SS N=256 K=16 (S full)   ", 0, 256},
                                                        {"SS N=96  K=16 (QKV step) ", 0, 96}, {"TS N=16  K=16 (P.V step) ", 1, 16},
                                                        {"TS N=64  K=16 (FFN2 step)", 1, 64}};
  for (auto& s : shapes) {
    for (int batch : {1, 4, 16}) {
      const int reps = 64;
      k_mma_timing<<<1, 128, 100 * 1024>>>(s.mode, s.n, batch, reps, out);
      cudaDeviceSynchronize();
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      long long issue = h[0], total = h[1];
      k_mma_roundtrip<<<1, 128, 100 * 1024>>>(s.mode, s.n, batch, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      printf("%s batch %2d: issue %6.1f clk/MMA  pipelined %6.1f clk/MMA  round trip (issue+commit+wait) %6.0f clk/batch  [%s]\n",
             s.name, batch, (double)issue / (reps * batch), (double)total / (reps * batch), (double)h[0] / reps, cudaGetErrorString(e));
    }
  }
  cudaFuncSetAttribute(k_quarter_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int busy = 0; busy < 2; ++busy) {
    const int reps = 200;
    k_quarter_pattern<<<1, 160, 100 * 1024>>>(busy, reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("quarter pattern (softmax warps %s): P.V 4xTS N=16 + commit: issue %5.0f clk, issue->complete %5.0f clk | S 1xSS N=64 + commit: "
           "issue %5.0f clk, issue->complete %5.0f clk [%s]\n", busy ? "busy" : "idle", (double)h[0] / reps, (double)h[1] / reps,
           (double)h[2] / reps, (double)h[3] / reps, cudaGetErrorString(e));
  }
  run_exp_mixed<0>(out, sink); run_exp_mixed<4>(out, sink); run_exp_mixed<8>(out, sink); run_exp_mixed<10>(out, sink);
  run_exp_mixed<12>(out, sink); run_exp_mixed<16>(out, sink);
  run_exp_mixed<4, 1>(out, sink); run_exp_mixed<6, 1>(out, sink); run_exp_mixed<7, 1>(out, sink); run_exp_mixed<8, 1>(out, sink);
  run_exp_mixed<10, 1>(out, sink); run_exp_mixed<16, 1>(out, sink);
  const char* names[] = {"tcgen05.ld x32 (+wait) ", "tcgen05.st x16         ", "ex2.approx             ", "ld+ffma+ex2+pack+st    ",
                         "ex2.approx.f16x2 (2/op)"};
  for (int what = 0; what < 5; ++what) {
    for (int threads : {128, 256}) {
      const int reps = 200;
      k_tmem_mufu<<<1, threads>>>(what, reps, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double per = (double)h[0] / (reps * 8);
      printf("%s %d warps/scheduler: %7.1f clk per 32-column step per warp (%.2f clk per element-row) [%s]\n", names[what],
             threads / 128, per, per / 32, cudaGetErrorString(e));
    }
  }
  return 0;
}
