// Probe for the two attention candidates of DESIGN.md §8 (TEST / DESIGN INFRASTRUCTURE — written at the end of round 1
// WITHOUT a GPU at hand: it compiles for sm_100a, it has not been run yet).
//
//   A. Does tcgen05.mma kind::f16 honour DIFFERENT A and B formats?  The exp pass could hand the MUFU-path
//      probabilities to the P.V MMA as bf16 (the upper halves of the two fp32 results, one PRMT on the ALU pipe) instead
//      of fp16 (one F2FP.PACK_AB on the XU pipe, which is the pipe the MUFU.EX2 saturate), if an MMA with A = bf16 in
//      TMEM and B = fp16 (the V^T slabs) in shared memory is legal.  Part A runs D[128 x 16] = A[128 x 32] B[16 x 32]^T
//      three ways — A fp16 / B fp16 (harness check), A bf16 / B fp16 (the question), A bf16 / B bf16 (control) — and
//      compares with a host reference built from the same rounded operands.
//   B. What does the exp pass cost per exponential with (i) PRMT instead of F2FP on the MUFU pairs, (ii) the
//      polynomial's 2^n factor inserted by one integer LEA + constant instead of shift + mask + HMUL2, (iii) both —
//      at 0, 5 and 8 polynomial pairs of 16, two warps per scheduler like the kernel.  Baseline = the shipped mix.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_variants exp_variants.cu ; run on the GPU box.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../seq2squiggle_b200/csrc/tc_prims.cuh"

using namespace s2s::tc;

// instruction descriptor with separate A / B formats (tc_prims' umma_idesc sets both to the same value)
__host__ __device__ constexpr uint32_t idesc_ab(uint32_t m, uint32_t n, uint32_t fmt_a, uint32_t fmt_b) {
  return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------------
// Part A
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline float a_value(int m, int k) { return 0.03125f * (float)((m * 7 + k * 3) % 61) - 0.9f; }
__host__ __device__ inline float b_value(int n, int k) { return 0.0625f * (float)((n * 5 + k * 11) % 29) - 0.85f; }

// fmt_a / fmt_b: kFmtF16 or kFmtBF16.  d_out: [128][16] fp32.
__global__ void __launch_bounds__(128) k_mixed_fmt(uint32_t fmt_a, uint32_t fmt_b, float* d_out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<64>(&s_base);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  // B: [16 rows (n) x 32 k] K-major SW128 tile (64 of the 128 bytes of a row used): thread t < 64 writes one 16-byte chunk
  if (tid < 64) {
    const int n = tid >> 2, ck = tid & 3;  // 4 chunks of 8 k per row
    uint32_t w[4];
    for (int j = 0; j < 4; ++j) {
      const float v0 = b_value(n, ck * 8 + 2 * j), v1 = b_value(n, ck * 8 + 2 * j + 1);
      if (fmt_b == kFmtF16) {
        w[j] = pack_half2(v0, v1);
      } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(smem + sw128_offset(n, ck)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_base;
  const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
  // A: row m = thread, 32 k as 16 packed columns at TMEM columns [0,16)
  uint32_t pk[16];
  for (int j = 0; j < 16; ++j) {
    const float v0 = a_value(tid, 2 * j), v1 = a_value(tid, 2 * j + 1);
    if (fmt_a == kFmtF16) pk[j] = pack_half2(v0, v1);
    else pk[j] = __byte_perm(__float_as_uint(v0), __float_as_uint(v1), 0x7632);  // bf16 by truncation: what the kernel would do
  }
  tmem_st_32x16(lane_addr, pk);
  tmem_wait_st();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_ab(128, 16, fmt_a, fmt_b);
    const uint64_t dB = umma_desc_k_sw128(smem_u32(smem));
    for (int ks = 0; ks < 2; ++ks) umma_f16_ts(tmem + 32, tmem + 8 * ks, dB + (uint64_t)((ks * 32) >> 4), idesc, ks > 0);
    umma_commit(&bar);
  }
  if (!mbar_wait(&bar, 0, status, 1)) return;
  tcgen05_fence_after();
  uint32_t d[16];
  tmem_ld_32x16(lane_addr + 32, d);
  tmem_wait_ld();
  for (int n = 0; n < 16; ++n) d_out[tid * 16 + n] = __uint_as_float(d[n]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

static float round_to(float v, uint32_t fmt, bool truncate_bf16) {
  if (fmt == kFmtF16) return __half2float(__float2half_rn(v));
  if (truncate_bf16) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xFFFF0000u;
    memcpy(&v, &u, 4);
    return v;
  }
  return __bfloat162float(__float2bfloat16_rn(v));
}

static void run_mixed(const char* name, uint32_t fa, uint32_t fb, float* d_dev, int* status) {
  cudaMemset(status, 0, 4);
  k_mixed_fmt<<<1, 128, 24 * 1024>>>(fa, fb, d_dev, status);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 16];
  int st = 0;
  cudaMemcpy(h, d_dev, sizeof h, cudaMemcpyDeviceToHost);
  cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
  double worst = 0, worst_swapped = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 16; ++n) {
      double ref = 0, ref_swapped = 0;
      for (int k = 0; k < 32; ++k) {
        ref += (double)round_to(a_value(m, k), fa, true) * (double)round_to(b_value(n, k), fb, false);
        // what a unit that ignored the A format and read the bf16 bits as fp16 would compute (only meaningful for fa = bf16)
        uint32_t u;
        float av = a_value(m, k);
        memcpy(&u, &av, 4);
        __half_raw hr;
        hr.x = (unsigned short)(u >> 16);
        ref_swapped += (double)__half2float(__half(hr)) * (double)round_to(b_value(n, k), fb, false);
      }
      worst = fmax(worst, fabs(ref - h[m * 16 + n]));
      worst_swapped = fmax(worst_swapped, fabs(ref_swapped - h[m * 16 + n]));
    }
  printf("mixed formats, %-22s: max |D - ref| = %.3g   (max |D - ref if A were read as fp16| = %.3g)  status %d [%s]\n", name, worst,
         worst_swapped, st, cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------------------------
// Part B
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// the shipped polynomial pair (k_tc.cu: ex2_poly_h2)
__device__ __forceinline__ uint32_t poly_h2_shipped(float x0, float x1) {
  const __half2 kLo = __float2half2_rn(-15.0f), kHi = __float2half2_rn(16.0f), kMagic = __float2half2_rn(1551.0f);
  const __half2 x = __hmin2(__hmax2(__floats2half2_rn(x0, x1), kLo), kHi);
  const __half2 t = __hadd2(x, kMagic);
  const __half2 f = __hsub2(x, __hsub2(t, kMagic));
  __half2 p = __hfma2(__float2half2_rn(0.05517167f), f, __float2half2_rn(0.24261113f));
  p = __hfma2(p, f, __float2half2_rn(0.69326097f));
  p = __hfma2(p, f, __float2half2_rn(0.99992806f));
  const uint32_t sc = (*reinterpret_cast<const uint32_t*>(&t) << 10) & 0x7C007C00u;
  const __half2 r = __hmul2(p, *reinterpret_cast<const __half2*>(&sc));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// 2^n by integer arithmetic on both lanes at once.  t = x + 1551 has the bit pattern 0x6600 + (n + 15) per lane, and
// p in [0.707, 1.414] has exponent field 14 or 15, so  bits(p * 2^n) = bits(p) + ((n + 15) << 10) - (15 << 10)  per lane
// as long as the lane result stays inside [0, 0xFFFF]: n + 15 >= 1 (clamp at -14 instead of -15) keeps it non-negative,
// n = 16 gives exponent field 31 (NaN / inf: the overflow flag, as before).  In one 32-bit word
//   (bits(t) << 10) mod 2^32 = ((n_hi + 15) << 26) + (0x6600 << 10) + ((n_lo + 15) << 10)
// (the high lane's 0x6600 is shifted out, the low lane's lands in the high lane as a constant), hence
//   bits(r) = bits(p) + (bits(t) << 10) - 0x01980000 - 0x3C003C00 :  one LEA and one IADD with a constant.
// Checked by a NumPy fp16 emulation at the end of round 1: bit-identical to the shipped polynomial for every finite
// result on [-14, 15.5); on [15.5, 16) it returns the true value where the shipped one already returns inf; x <= -14.5
// returns 2^-14 p (6.1e-5) where the shipped one returns 2^-15 p or exactly 0 — at most 6.1e-5 per far-tail key against
// a row maximum >= 1, to be weighed (sharp rows) before it ships.
__device__ __forceinline__ uint32_t poly_h2_lea(float x0, float x1) {
  const __half2 kLo = __float2half2_rn(-14.0f), kHi = __float2half2_rn(16.0f), kMagic = __float2half2_rn(1551.0f);
  const __half2 x = __hmin2(__hmax2(__floats2half2_rn(x0, x1), kLo), kHi);
  const __half2 t = __hadd2(x, kMagic);
  const __half2 f = __hsub2(x, __hsub2(t, kMagic));
  __half2 p = __hfma2(__float2half2_rn(0.05517167f), f, __float2half2_rn(0.24261113f));
  p = __hfma2(p, f, __float2half2_rn(0.69326097f));
  p = __hfma2(p, f, __float2half2_rn(0.99992806f));
  const uint32_t tb = *reinterpret_cast<const uint32_t*>(&t), pb = *reinterpret_cast<const uint32_t*>(&p);
  return pb + (tb << 10) - (0x01980000u + 0x3C003C00u);
}

// kPoly pairs of 16 by the polynomial; kPrmt: MUFU pairs packed as bf16 by PRMT; kLea: polynomial with the integer 2^n
// kDirect: the accumulator already holds the exponent (reference folded into the S MMA): no scaling FFMA per score
template <int kPoly, bool kPrmt, bool kLea, bool kDirect = false>
__device__ __forceinline__ void exp_step(const uint32_t (&r)[32], uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = kDirect ? __uint_as_float(r[2 * i]) : fmaf(__uint_as_float(r[2 * i]), 0.5f, -1.f);
    const float x1 = kDirect ? __uint_as_float(r[2 * i + 1]) : fmaf(__uint_as_float(r[2 * i + 1]), 0.5f, -1.f);
    if ((i * kPoly) % 16 < kPoly) {
      pk[i] = kLea ? poly_h2_lea(x0, x1) : poly_h2_shipped(x0, x1);
    } else {
      const float p0 = ex2f(x0), p1 = ex2f(x1);
      pk[i] = kPrmt ? __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632) : pack_half2(p0, p1);
    }
  }
  tmem_st_32x16(taddr, pk);
}

template <int kPoly, bool kPrmt, bool kLea, bool kDirect = false>
__global__ void __launch_bounds__(512) k_exp(int reps, long long* out, float* sink) {
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&s_base);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t lane_addr = tmem_addr(s_base, (warp & 3) * 32, (warp >> 2) * 128);
  uint32_t ra[32], rb[32];
  __syncthreads();
  long long t0 = clock64();
  for (int k = 0; k < reps; ++k) {  // double-buffered loads like the kernel
    tmem_ld_32x32(lane_addr, ra);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      tmem_ld_32x32(lane_addr + ((c + 1) & 3) * 32, rb);
      tmem_wait_ld();
      exp_step<kPoly, kPrmt, kLea, kDirect>(ra, lane_addr + (c & 3) * 16);
      if (c + 2 < 8) tmem_ld_32x32(lane_addr + ((c + 2) & 3) * 32, ra);
      exp_step<kPoly, kPrmt, kLea, kDirect>(rb, lane_addr + ((c + 1) & 3) * 16);
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

template <int kPoly, bool kPrmt, bool kLea, bool kDirect = false>
void run_exp(long long* out, float* sink) {
  long long h[16];
  for (int threads : {128, 256, 384}) {
    const int reps = 200;
    k_exp<kPoly, kPrmt, kLea, kDirect><<<1, threads>>>(reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (reps * 8);
    printf("exp pass%s, %2d/16 polynomial pairs, MUFU pack by %s, 2^n by %s, %d warps/scheduler: %6.1f clk per 32-column step "
           "per warp -> %5.2f clk per exp row per scheduler [%s]\n", kDirect ? " (direct, no FFMA)" : "", kPoly, kPrmt ? "PRMT (bf16)" : "F2FP (fp16)",
           kLea ? "LEA + const  " : "shift/mask/HMUL2", threads / 128, per, per / 32 / (threads / 128), cudaGetErrorString(e));
  }
}

// numerics of the integer 2^n against the shipped polynomial over the whole clamped range (device, one thread per x)
__global__ void k_poly_check(float* worst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // x = -15 + i / 1024
  const float x = -15.0f + (float)i * (1.0f / 1024.0f);
  if (x > 15.9f) return;
  const uint32_t a = poly_h2_shipped(x, x + 0.013f), b = poly_h2_lea(x, x + 0.013f);
  const __half2 ha = *reinterpret_cast<const __half2*>(&a), hb = *reinterpret_cast<const __half2*>(&b);
  const float ref0 = exp2f(x), ref1 = exp2f(x + 0.013f);
  float e = 0.f;
  if (x >= -14.0f) {   // below the new clamp the two differ by design (both are < 2^-14 of the row maximum)
    e = fmaxf(fabsf(__low2float(ha) - __low2float(hb)) / ref0, fabsf(__high2float(ha) - __high2float(hb)) / ref1);
  }
  const float e_abs = fmaxf(fabsf(__low2float(hb) - ref0) / ref0, fabsf(__high2float(hb) - ref1) / ref1);
  atomicMax(reinterpret_cast<int*>(worst), __float_as_int(e));          // non-negative floats order like ints
  atomicMax(reinterpret_cast<int*>(worst + 1), __float_as_int(e_abs));
}

int main(int argc, char** argv) {
  // every part in its own process: an illegal-instruction fault (part A2 on B200) is sticky for the whole context
  const char* part = argc > 1 ? argv[1] : "B";
  long long* out;
  float *sink, *d_dev;
  int* status;
  cudaMalloc(&out, 64 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(float));
  cudaMalloc(&d_dev, 128 * 16 * sizeof(float));
  cudaMalloc(&status, 4);
  cudaFuncSetAttribute(k_mixed_fmt, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024);
  if (!strcmp(part, "A1")) { run_mixed("A fp16 / B fp16", kFmtF16, kFmtF16, d_dev, status); return 0; }
  if (!strcmp(part, "A2")) { run_mixed("A bf16 / B fp16 (?)", kFmtBF16, kFmtF16, d_dev, status); return 0; }
  if (!strcmp(part, "A3")) { run_mixed("A bf16 / B bf16", kFmtBF16, kFmtBF16, d_dev, status); return 0; }

  float* worst;
  cudaMalloc(&worst, 8);
  cudaMemset(worst, 0, 8);
  k_poly_check<<<(31 * 1024 + 255) / 256, 256>>>(worst);
  float hw[2];
  cudaMemcpy(hw, worst, 8, cudaMemcpyDeviceToHost);
  printf("integer 2^n vs shipped polynomial on [-14, 15.9]: max relative difference %.3g; max relative error vs exp2f %.3g [%s]\n", hw[0],
         hw[1], cudaGetErrorString(cudaDeviceSynchronize()));

  run_exp<0, false, false>(out, sink); run_exp<0, true, false>(out, sink);
  run_exp<5, false, false>(out, sink); run_exp<5, true, false>(out, sink);
  run_exp<5, false, true>(out, sink);  run_exp<5, true, true>(out, sink);
  run_exp<8, false, false>(out, sink); run_exp<8, true, false>(out, sink);
  run_exp<8, false, true>(out, sink);  run_exp<8, true, true>(out, sink);
  // round 2: exponent straight from the accumulator (k_tc_attn3: reference folded into the S MMA)
  run_exp<0, false, false, true>(out, sink);
  run_exp<6, false, false, true>(out, sink); run_exp<7, false, false, true>(out, sink);
  run_exp<8, false, false, true>(out, sink); run_exp<9, false, false, true>(out, sink);
  run_exp<10, false, false, true>(out, sink); run_exp<11, false, false, true>(out, sink);
  run_exp<8, false, true, true>(out, sink); run_exp<9, false, true, true>(out, sink); run_exp<10, false, true, true>(out, sink);
  run_exp<8, true, false, true>(out, sink); run_exp<10, true, true, true>(out, sink);
  return 0;
}
