// Micro-benchmark (TEST / DESIGN INFRASTRUCTURE): sustained rate of small tcgen05.mma instructions on one SM as a
// function of the shape (M, N), the operand source (SS: A from shared memory, TS: A from TMEM), the number of
// independent accumulators the instructions rotate over, and whether they accumulate.  The decoder attention issues
// 20 small MMAs per (head, query tile); this measures what the tensor pipe charges for each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu ; run on the GPU box.
#include <stdio.h>
#include <stdlib.h>

#include "../../seq2squiggle_b200/csrc/tc_prims.cuh"

using namespace s2s::tc;

__host__ __device__ constexpr uint32_t idesc_mn(uint32_t m, uint32_t n) { return umma_idesc(m, n, kFmtF16); }

// mode 0: SS, 1: TS.  `nacc` accumulators of `n` columns each (rotating), `acc` = accumulate flag of all but the
// first MMA on an accumulator.  total = reps * batch MMAs, one commit per batch.
__global__ void __launch_bounds__(128) k_rate(int mode, int m, int n, int nacc, int acc, int batch, int reps, long long* out) {
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
    const uint32_t idesc = idesc_mn(m, n);
    long long t0 = clock64();
    int k = 0;
    for (int r = 0; r < reps; ++r) {
      for (int b = 0; b < batch; ++b, ++k) {
        const uint32_t d = tmem + (uint32_t)(k & (nacc - 1)) * (uint32_t)n;   // accumulators in columns [0, nacc*n)
        const uint32_t a = k >= nacc ? (uint32_t)acc : 0u;
        if (mode == 0) umma_f16_ss(d, dA + (uint64_t)((b & 3) * 2), dB + (uint64_t)((b & 3) * 2), idesc, a);
        else umma_f16_ts(d, tmem + 384 + (b & 15) * 8, dB + (uint64_t)((b & 3) * 2), idesc, a);
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


// Same measurement with every operand a compile-time constant offset from a uniform base (fully unrolled batch of 16),
// so that the issuing thread's own integer / R2UR work cannot be what is measured.
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128) k_rate_c(int reps, long long* out) {
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
  if (warp == 0) {
    const uint64_t dA = umma_desc_k_sw128(smem_u32(smem)), dB = umma_desc_k_sw128(smem_u32(smem + 32768));
    constexpr uint32_t idesc = idesc_mn(128, N);
    const uint32_t bar0 = smem_u32(&bar[0]);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          if (MODE == 0) umma_f16_ss(tmem + (b % NACC) * N, dA + (uint64_t)((b & 3) * 2), dB + (uint64_t)((b & 3) * 2), idesc, 1u);
          else umma_f16_ts(tmem + (b % NACC) * N, tmem + 384 + (b & 15) * 8, dB + (uint64_t)((b & 3) * 2), idesc, 1u);
        }
        umma_commit_a(bar0 + 8 * (r & 63));
      }
      __syncwarp();
    }
    long long t1 = clock64();
    uint32_t ok = 0;
    for (uint32_t i = 0; i < (1u << 24) && !ok; ++i) ok = mbar_try_wait(&bar[(reps - 1) & 63], ((reps - 1) >> 6) & 1);
    long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = ok; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE, int N, int NACC>
void run_c(long long* out) {
  long long h[4];
  const int reps = 32;
  cudaFuncSetAttribute(k_rate_c<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_rate_c<MODE, N, NACC><<<1, 128, 100 * 1024>>>(reps, out);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  printf("const-operand %s M=128 N=%3d accumulators=%d: issue %6.1f clk/MMA  sustained %6.1f clk/MMA (math %5.1f clk) [%s%s]\n",
         MODE ? "TS" : "SS", N, NACC, (double)h[0] / (reps * 16), (double)h[1] / (reps * 16), 2.0 * 128 * N * 16 / 8192.0,
         cudaGetErrorString(e), h[2] ? "" : " TIMEOUT");
}

int main() {
  long long* out;
  cudaMalloc(&out, 64 * sizeof(long long));
  run_c<1, 16, 1>(out); run_c<1, 16, 2>(out); run_c<1, 16, 4>(out); run_c<1, 64, 1>(out); run_c<1, 64, 2>(out);
  run_c<0, 16, 1>(out); run_c<0, 48, 2>(out); run_c<0, 64, 1>(out); run_c<0, 64, 2>(out); run_c<0, 128, 2>(out); run_c<0, 256, 1>(out);
  if (getenv("RATE_ONLY_CONST")) return 0;
  long long h[4];
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 32, batch = 16;
  for (int mode = 0; mode < 2; ++mode)
    for (int m : {128, 64})
      for (int n : {16, 48, 64, 96, 128, 256})
        for (int nacc : {1, 2, 4})
          for (int acc : {1}) {
            if (nacc * n > 384) continue;
            if (m == 64 && n % 8) continue;
            k_rate<<<1, 128, 100 * 1024>>>(mode, m, n, nacc, acc, batch, reps, out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
            printf("%s M=%3d N=%3d K=16  accumulators=%d accumulate=%d : issue %6.1f clk/MMA  sustained %6.1f clk/MMA  (math at 8192 FLOP/clk: %5.1f clk) [%s%s]\n",
                   mode ? "TS" : "SS", m, n, nacc, acc, (double)h[0] / (reps * batch), (double)h[1] / (reps * batch),
                   2.0 * m * n * 16 / 8192.0, cudaGetErrorString(e), h[2] ? "" : " TIMEOUT");
          }
  return 0;
}
