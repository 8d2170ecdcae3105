// Micro-benchmarks of the sm_100a primitives the attention kernel is built from (TEST / DESIGN INFRASTRUCTURE).
// One CTA per launch; operands are whatever is in shared memory (timing only, no numerics).  Prints clk numbers:
//   * tcgen05.mma issue cost and issue->completion latency for the three shapes used (S: SS N=64/256 K=16,
//     PV: TS N=16 K=16, QKV: SS N=96) as a function of the batch size between commits;
//   * tcgen05.ld / st throughput per warp; MUFU.EX2 throughput for 1 and 2 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_timing mma_timing.cu ; run on the GPU box.
#include <stdio.h>
#include <stdlib.h>

#include "../../seq2squiggle_b200/csrc/tc_prims.cuh"

using namespace s2s::tc;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
      for (int c = 0; c < 8; ++c) { tmem_ld_32x32(lane_addr + c * 32, r); tmem_wait_ld(); acc += __uint_as_float(r[k & 31]); }
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
  const char* names[] = {"tcgen05.ld x32 (+wait) ", "tcgen05.st x16         ", "ex2.approx             ", "ld+ffma+ex2+pack+st    "};
  for (int what = 0; what < 4; ++what) {
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
