// Decoder FFT blocks on the 5th-generation tensor cores (S2S_PREC_FP16_TC).
//
// layers.py:44-142 at L = 250 (256 rows per chunk in HBM).  fp16 operands (the reference's own GPU dtype:
// Lightning "16-mixed", inference.py:404), fp32 accumulation in TMEM, fp32 softmax / LayerNorm / residual.
// Four kernels per layer, all tcgen05.mma (kind::f16, M=128) fed by TMA into 128-byte-swizzled K-major
// shared-memory tiles, accumulators in TMEM, epilogues by tcgen05.ld with one thread per row:
//   k_tc_qkv        Q|K|V = X Wqkv^T + b          -> Q16 [rows,64], Kmask [rows,128], V^T [chunk,8,16,256]
//   k_tc_attention  per (chunk, 4 heads): S = Q_h K_h^T (one N=256,K=16 MMA), row softmax in registers,
//                   P (fp16) written back to TMEM over S, O = P V_h (16 TS MMAs, A operand from TMEM)
//   k_tc_fc_ln      Y = LayerNorm(O Wfc^T + b + X)
//   k_tc_ffn_ln     X' = LayerNorm(relu(Y W1^T + b1) W2^T + b2 + Y); the 256-wide hidden never leaves TMEM
//
// d_k = 8 but the fp16 MMA has K = 16: K is stored "masked" (head h's 8 values in the half of a 32-byte slot
// that lines up with head h inside the 32-byte Q slice of heads {2i,2i+1}; the other half is zero), so one
// K=16 MMA computes exactly q_h . k_h.  d_v = 8 but N >= 16 for M = 128: V_h^T is padded to 16 rows with a row
// of ones, so column 8 of O is the softmax denominator (sum of the ROUNDED probabilities) for free.
#include <stdlib.h>
#include <string.h>

#include "s2s_tc.h"
#include "tc_host.h"
#include "tc_prims.cuh"

namespace s2s {
using namespace tc;

namespace {

constexpr uint32_t kWaitLimit = 1u << 22;
constexpr uint32_t kStreamNoiseTc = 0x5D0003u;   // the Philox stream of the noise draw (k_epilogue.cu: kStreamNoise)
constexpr int kSlab = 128 * 128;  // bytes of one [128 rows x 128 B] tile

// status codes written on a barrier timeout
enum { kErrQkvLoad = 11, kErrQkvMma = 12, kErrAttLoad = 21, kErrAttS = 22, kErrAttO = 23, kErrAttTmem = 24, kErrFcLoad = 31,
       kErrFcMma = 32, kErrFfnLoad = 41, kErrFfnMma1 = 42, kErrFfnMma2 = 43 };

// Optional phase timing of k_tc_attn (compile with -DS2S_PHASE_TIMING): clock64() deltas of thread 0 of every CTA,
// summed into g_phase[] and read back through s2s_debug_counters().
__device__ unsigned long long g_phase[16];
#if defined(S2S_PHASE_TIMING) && S2S_PHASE_TIMING == 1
// register accumulators (constant indices): a clock read + one 64-bit add per PHASE(), nothing else
#define PHASE_DECL long long ph_acc[16]; _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) ph_acc[i_] = 0; long long ph_t = clock64();
#define PHASE(i) do { long long n_ = clock64(); ph_acc[i] += n_ - ph_t; ph_t = n_; } while (0)
#define PHASE_FLUSH do { if (threadIdx.x == 0 || threadIdx.x == 128) { _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) if (ph_acc[i_]) atomicAdd(&g_phase[i_], (unsigned long long)ph_acc[i_]); } } while (0)
#define PHASE_COUNT(i) do { ph_acc[i] += 1; } while (0)
#else
#define PHASE_DECL
#define PHASE(i) do {} while (0)
#define PHASE_FLUSH do {} while (0)
#define PHASE_COUNT(i) do {} while (0)
#endif

__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, int* status, volatile int* s_abort, int code) {
  if (*s_abort) return false;
  for (uint32_t i = 0; i < kWaitLimit; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  *s_abort = 1;
  atomicExch(status, code);
  return false;
}

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =================================================================================================
// ATT: fused QKV projection + attention for one (chunk, group of 4 heads).  128 threads = 128 rows of a tile,
// 2 CTAs / SM (TMEM 256 columns and ~109 KB of shared memory each) so one CTA's MMA round trips and row-max
// pass overlap the other CTA's exponentials.
//   1. TMA: X16 chunk (2 tiles [128 x 64]) ; the group's weight block [96 x 64] (Wq | Wk | Wv rows) once per CTA
//   2. UMMA: [128 x 96] = X_tile Wg^T for both tiles (accumulators at TMEM columns 0 and 128)
//   3. epilogue: + bias, fp16, written straight into the operand layouts in shared memory (never to HBM):
//        Q  -> bytes [0,64) of the rows of the (now dead) X tiles      (A operand of S, K-slices of 32 B)
//        K  -> masked 32-byte slots, [256 keys x 128 B]                 (B operand of S)
//        V  -> transposed and padded, 4 key slabs x [4 heads x 16 rows x 64 keys]   (B operand of P.V)
//   4. per head and query tile: S MMA (N=256, K=16) -> two-pass softmax on tcgen05.ld double buffers ->
//      P (fp16) over S in TMEM -> 16 TS MMAs (N=16) -> O / rowsum -> fp16 -> O16 in HBM
// =================================================================================================
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int kValid>
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], float m) {
  static_assert(kValid % 2 == 0, "pairs");
#pragma unroll
  for (int i = 0; i < kValid; i += 2) m = max3(m, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
  return m;
}

template <int kValid>
__device__ __forceinline__ void chunk_exp_store(const uint32_t (&r)[32], float scale, float mneg, uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float p0 = 2 * i < kValid ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), scale, mneg)) : 0.f;
    float p1 = 2 * i + 1 < kValid ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), scale, mneg)) : 0.f;
    pk[i] = pack_half2(p0, p1);
  }
  tmem_st_32x16(taddr, pk);
}

// 2^x for a PAIR of scores in packed fp16 arithmetic (HFMA2 / HADD2 on the FMA pipe, no MUFU).  P is rounded to fp16
// for the P.V MMA anyway; here the argument is rounded to fp16 first (|x| < 16: absolute error <= 2^-8, i.e. a relative
// error of P of at most 0.27 %, 0.07 % for |x| < 4).  n = rint(x) comes from the magic constant 1536 + 15: the low five
// mantissa bits of t = x + 1551 are n + 15, which is the fp16 exponent field of 2^n, so 2^n is one shift + mask and the
// result is p(f) * 2^n by one HMUL2.  Degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5).  x is clamped to [-15, 16]:
// 2^-15 * p rounds to <= 3.1e-5 (the row's largest P is >= 1), and n = 16 gives the exponent field 31 = inf, which the
// denominator check turns into the exact-kernel fallback exactly like an overflowing MUFU result.
__device__ __forceinline__ uint32_t ex2_poly_h2(float x0, float x1) {
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

__global__ void __launch_bounds__(128, 2) k_tc_attn(const __grid_constant__ CUtensorMap tmX,
                                                    const __grid_constant__ CUtensorMap tmWg,
                                                    const float* __restrict__ bias_g, __half* __restrict__ o16,
                                                    int n_units, const int* __restrict__ unit_flags,
                                                    const int* __restrict__ n_flagged, int* status,
                                                    int* __restrict__ hint_out = nullptr) {
  // As the exact fallback of k_tc_attn3 (unit_flags != nullptr) only the flagged units are recomputed.
  if (unit_flags) {
    const int nf = *n_flagged;
    // n_flagged counts flagging WARPS (up to 4 per unit).  More than half of them: this layer's attention is too sharp
    // for the single-reference kernel; the hint makes the next sub-batches skip it (see k_tc_attn3).
    if (hint_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *hint_out = nf > 2 * n_units ? 1 : 0;
    if (nf == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_phase[12], (unsigned long long)nf);
  }
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_w, bar_s, bar_o;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[2][96];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sXQ = smem;                 // 2 x [128 x 128 B]: X tiles, then Q (bytes [0,64) of each row)
  uint8_t* sK = smem + 2 * kSlab;      // [256 keys x 128 B]  masked K of this head group
  uint8_t* sV = smem + 4 * kSlab;      // 4 key slabs x [64 rows (4 heads x 16) x 128 B]
  uint8_t* sW = smem + 6 * kSlab;      // 2 groups... one group at a time: [96 x 128 B]
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) s_go = (*status == 0);
  __syncthreads();
  if (!s_go) return;
  if (warp == 0) tmem_alloc<256>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bar_load, 1); mbar_init(&bar_w, 1); mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWg);
  }
  for (int i = tid; i < 192; i += 128) s_bias[i / 96][i % 96] = bias_g[i];
  // V^T padding rows are constant: row 8 of every head = ones (softmax denominator), rows 9..15 = 0
  for (int i = tid; i < 2 * kSlab / 16; i += 128) reinterpret_cast<uint4*>(sV)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < 4 * 4 * 8; i += 128) {  // (slab, head, 16-byte chunk of 8 keys)
    const int slab = i >> 5, hh = (i >> 3) & 3, ck = i & 7;
    const uint32_t one2 = 0x3C003C00u;  // two fp16 ones
    *reinterpret_cast<uint4*>(sV + slab * 8192 + sw128_offset(hh * 16 + 8, ck)) = make_uint4(one2, one2, one2, one2);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
  const uint32_t idesc_qkv = umma_idesc(128, 96, kFmtF16), idesc_s = umma_idesc(128, 256, kFmtF16),
                 idesc_o = umma_idesc(128, 16, kFmtF16);
  const float kScale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e) / sqrt(d_k)
  uint32_t ph_load = 0, ph_w = 0, ph_s = 0, ph_o = 0;
  int cur_g = -1;
  PHASE_DECL
  PHASE(9);  // prologue
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    if (unit_flags && !unit_flags[unit]) continue;
    const int chunk = unit >> 1, g = unit & 1;
    PHASE_COUNT(15);
    if (tid == 0) {
      if (g != cur_g) {  // with an even grid stride every CTA keeps its head group: loaded once
        mbar_arrive_expect_tx(&bar_w, 96 * 128);
        tma_load_2d(sW, &tmWg, &bar_w, 0, g * 96);
      }
      mbar_arrive_expect_tx(&bar_load, 2 * kSlab);
      tma_load_2d(sXQ, &tmX, &bar_load, 0, chunk * 256);
      tma_load_2d(sXQ + kSlab, &tmX, &bar_load, 0, chunk * 256 + 128);
    }
    if (g != cur_g) {
      wait_bar(&bar_w, ph_w, status, &s_abort, kErrAttLoad);
      ph_w ^= 1;
      cur_g = g;
    }
    wait_bar(&bar_load, ph_load, status, &s_abort, kErrAttLoad);
    ph_load ^= 1;
    tcgen05_fence_after();
    PHASE(0);
    if (tid == 0) {  // [128 x 96] = X_tile Wg^T, both tiles
      const uint32_t w0 = smem_u32(sW);
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t a0 = smem_u32(sXQ + tile * kSlab);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_f16_ss(tmem + tile * 128, umma_desc_k_sw128(a0 + s * 32), umma_desc_k_sw128(w0 + s * 32), idesc_qkv, s > 0);
      }
      umma_commit(&bar_s);
    }
    wait_bar(&bar_s, ph_s, status, &s_abort, kErrAttS);
    ph_s ^= 1;
    tcgen05_fence_after();
    PHASE(1);
    {  // QKV epilogue: accumulators -> fp16 operands in shared memory
      uint32_t r[32];
      const float* bq = s_bias[g];
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const int t = tile * 128 + tid;  // key / query index inside the chunk
        tmem_ld_32x32(lane_addr + tile * 128, r);
        tmem_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pk[i] = pack_half2(__uint_as_float(r[8 * hh + 2 * i]) + bq[8 * hh + 2 * i],
                               __uint_as_float(r[8 * hh + 2 * i + 1]) + bq[8 * hh + 2 * i + 1]);
          *reinterpret_cast<uint4*>(sXQ + tile * kSlab + sw128_offset(tid, hh)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        tmem_ld_32x32(lane_addr + tile * 128 + 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pk[i] = pack_half2(__uint_as_float(r[8 * hh + 2 * i]) + bq[32 + 8 * hh + 2 * i],
                               __uint_as_float(r[8 * hh + 2 * i + 1]) + bq[32 + 8 * hh + 2 * i + 1]);
          const uint4 data = make_uint4(pk[0], pk[1], pk[2], pk[3]), zero = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(sK + sw128_offset(t, 2 * hh)) = (hh & 1) ? zero : data;
          *reinterpret_cast<uint4*>(sK + sw128_offset(t, 2 * hh + 1)) = (hh & 1) ? data : zero;
        }
        tmem_ld_32x32(lane_addr + tile * 128 + 64, r);
        tmem_wait_ld();
        uint8_t* vslab = sV + (t >> 6) * 8192 + (t & 7) * 2;
        const uint32_t ck = (t & 63) >> 3;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
#pragma unroll
          for (int d = 0; d < 8; ++d)
            *reinterpret_cast<__half*>(vslab + sw128_offset(hh * 16 + d, ck)) =
                __float2half_rn(__uint_as_float(r[8 * hh + d]) + bq[64 + 8 * hh + d]);
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    PHASE(2);
#pragma unroll 1
    for (int hh = 0; hh < 4; ++hh) {
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        if (tid == 0) {  // S = Q_h K_h^T : one MMA, N = 256 keys, K = 16 (8 real + 8 masked)
          umma_f16_ss(tmem, umma_desc_k_sw128(smem_u32(sXQ + tile * kSlab) + (hh >> 1) * 32),
                      umma_desc_k_sw128(smem_u32(sK) + hh * 32), idesc_s, 0);
          umma_commit(&bar_s);
        }
        wait_bar(&bar_s, ph_s, status, &s_abort, kErrAttS);
        ph_s ^= 1;
        tcgen05_fence_after();
        PHASE(3);
        uint32_t ra[32], rb[32];
        // pass 1: row maximum over the 250 real keys (loads double-buffered against the max3 chains)
        float m = -INFINITY;
        tmem_ld_32x32(lane_addr, ra);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_32x32(lane_addr + (c + 1) * 32, rb);
          m = chunk_max<32>(ra, m);
          tmem_wait_ld();
          if (c + 2 < 8) tmem_ld_32x32(lane_addr + (c + 2) * 32, ra);
          m = (c + 1 == 7) ? chunk_max<S2S_L_DEC - 224>(rb, m) : chunk_max<32>(rb, m);
          if (c + 2 < 8) tmem_wait_ld();
        }
        const float mneg = -m * kScale;
        PHASE(4);
        // pass 2: P = exp2(S*c - m*c) -> fp16, written over S (chunk c -> columns [16c,16c+16))
        tmem_ld_32x32(lane_addr, ra);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_32x32(lane_addr + (c + 1) * 32, rb);
          chunk_exp_store<32>(ra, kScale, mneg, lane_addr + c * 16);
          tmem_wait_ld();
          if (c + 2 < 8) tmem_ld_32x32(lane_addr + (c + 2) * 32, ra);
          if (c + 1 == 7) chunk_exp_store<S2S_L_DEC - 224>(rb, kScale, mneg, lane_addr + (c + 1) * 16);
          else chunk_exp_store<32>(rb, kScale, mneg, lane_addr + (c + 1) * 16);
          if (c + 2 < 8) tmem_wait_ld();
        }
        tmem_wait_st();
        PHASE(5);
        tcgen05_fence_before();
        __syncthreads();
        if (tid == 0) {  // O = P V_h : 16 K-steps of 16 keys, A operand straight from TMEM
          tcgen05_fence_after();
          const uint32_t vb = smem_u32(sV) + hh * 2048;
#pragma unroll
          for (int s = 0; s < 16; ++s)
            umma_f16_ts(tmem + 128, tmem + 8 * s, umma_desc_k_sw128(vb + (s >> 2) * 8192 + (s & 3) * 32), idesc_o, s > 0);
          umma_commit(&bar_o);
        }
        wait_bar(&bar_o, ph_o, status, &s_abort, kErrAttO);
        ph_o ^= 1;
        tcgen05_fence_after();
        PHASE(6);
        uint32_t o[16];
        tmem_ld_32x16(lane_addr + 128, o);
        tmem_wait_ld();
        const float inv = 1.0f / __uint_as_float(o[8]);  // column 8 = sum of the rounded probabilities
        const int64_t row = (int64_t)chunk * 256 + tile * 128 + tid;
        *reinterpret_cast<uint4*>(o16 + row * 64 + (g * 4 + hh) * 8) =
            make_uint4(pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv),
                       pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv),
                       pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv),
                       pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv));
        tcgen05_fence_before();
        __syncthreads();  // O has been read everywhere before the next S MMA reuses the columns
        tcgen05_fence_after();
        PHASE(7);
      }
    }
  }
  PHASE_FLUSH;
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// Fused output epilogue of the last decoder block (modules.py:140-141, model.py:221-240): p = ReLU(out_linear), pA = 165 p,
// Philox noise where pA != 0, clamp at 0 — written ONCE as fp32 pA, 250 positions per chunk, plus the chunk's count of
// non-zero samples (what the zero-strip compaction scans); pad rows 250..255 emit nothing.  y: the thread's LayerNorm-2 row.
__device__ __forceinline__ void out_head_epilogue(const float (&y)[64], const FfnParams& P, const OutEpi& E, int64_t row) {
  float acc = P.bout;
#pragma unroll
  for (int i = 0; i < 64; ++i) acc = fmaf(y[i], P.wout[i], acc);
  const int64_t c = row >> 8;
  const int t = (int)(row & 255);
  bool nz = false;
  if (t < S2S_L_DEC) {
    const int64_t idx = c * S2S_L_DEC + t;
    const float p = fmaxf(acc, 0.f);
    if (E.p_tap) E.p_tap[idx] = p;
    float v_pa = p * E.scaling;
    if (E.o.noise_mode != S2S_NOISE_OFF && v_pa != 0.f) {
      const Philox ph(E.o.seed);
      const uint64_t gc = E.o.chunk_id_base + (uint64_t)c;
      const uint4 rr = ph((uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)t, kStreamNoiseTc);
      const float z = box_muller(rr.x, rr.y).x;
      const float sd = E.o.noise_mode == S2S_NOISE_SAMPLER
                           ? fmaxf(E.sigma_ext[idx], E.o.min_noise) * E.o.noise_std * E.scaling   // model.py:228-230
                           : E.o.noise_std;                                                        // model.py:236
      v_pa += z * sd;
    }
    v_pa = fmaxf(v_pa, 0.f);
    E.pa[idx] = v_pa;
    nz = v_pa != 0.f;
  }
  if (E.counts) {   // a warp's 32 rows belong to one chunk (256 rows per chunk)
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(E.counts + c, __popc(m));
  }
}

// Adaptive fallback gate, launched in front of k_tc_attn3.  When the previous sub-batch of this layer sent most of its
// units to the exact kernel (hint set by k_tc_attn), running the single-reference kernel first only wastes its time
// (measured with W_q, W_k scaled x4: 1.21 M chunks/s with both kernels, 1.95 M with the exact kernel alone): every
// unit is flagged here and status[1] tells k_tc_attn3 to return at once.  The host re-probes every 16th call.
__global__ void k_attn_gate(const int* __restrict__ hint, int probe, int n_units, int* __restrict__ unit_flags,
                            int* __restrict__ n_flagged, int* __restrict__ status) {
  const bool skip = !probe && *hint != 0;
  if (skip)
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) unit_flags[u] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    status[1] = skip ? 1 : 0;
    if (skip) *n_flagged = 4 * n_units;
  }
}
#include "k_tc_attn3.cuh"

#include "k_tc_ffn4.cuh"
#include "k_tc_enc.cuh"

constexpr int kSmemAtt = 6 * kSlab + 96 * 128 + 1024;   // 109 KB -> 2 CTAs / SM

}  // namespace

// -------------------------------------------------------------------------------------------------
void tc_carve(TcBuffers& b, char* base, int64_t& off, int64_t bc) {
  auto take = [&](int64_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  const int64_t rows = bc * S2S_L_DEC_PAD;
  b.x16 = reinterpret_cast<__half*>(take(rows * 64 * 2));
  b.o16 = reinterpret_cast<__half*>(take(rows * 64 * 2));
  const int64_t erows = align_up(bc * S2S_L_ENC, 128);
  b.xe16 = reinterpret_cast<__half*>(take(erows * 64 * 2));
  b.oe16 = reinterpret_cast<__half*>(take(erows * 64 * 2));
  b.flags = reinterpret_cast<int32_t*>(take((2 * bc + 1) * 4));
}

int tc_init(TcState& s, const DevWeights& w, int device) {
  s.device = device;
  cudaDeviceProp prop;
  S2S_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  s.sm_count = prop.multiProcessorCount;
  s.encode_tiled = reinterpret_cast<void*>(get_encode_tiled());
  if (!s.encode_tiled) {
    set_error("cuTensorMapEncodeTiled driver entry point not found");
    return -1;
  }
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_enc_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEncAttn));
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAtt));
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_attn3, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAtt3));
  if (const char* env = getenv("S2S_ATTN_EXACT")) s.attn_exact = atoi(env) != 0;   // tests: the exact kernel on its own
  if (const char* env = getenv("S2S_ATTN_V1")) s.attn_exact = atoi(env) != 0;      // (older spelling)
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_fc_ffn4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFfn4));
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_fc_ffn4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFfn4));
  S2S_CUDA_OK(cudaFuncSetAttribute(k_tc_fc_ffn4<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFfn4));
  (void)w;
  S2S_CUDA_OK(cudaMalloc(&s.d_status, 256));
  S2S_CUDA_OK(cudaMemset(s.d_status, 0, 256));
  return 0;
}

void tc_destroy(TcState& s) {
  if (s.d_status) cudaFree(s.d_status);
  s.d_status = nullptr;
}

int tc_decoder(TcState& s, const DevWeights& w, const TcBuffers& b, const OutEpi& epi, int64_t n_chunks, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(s.encode_tiled);
  const uint64_t rows = (uint64_t)n_chunks * S2S_L_DEC_PAD;
  const int n_tiles = (int)(rows / 128);
  const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tmX, tmO;
  bool ok = make_tmap_2d(enc, &tmX, b.x16, f16, 2, rows, 64, 128, 64, sw) &&
            make_tmap_2d(enc, &tmO, b.o16, f16, 2, rows, 64, 128, 64, sw);
  if (!ok) {
    set_error("cuTensorMapEncodeTiled failed for an activation tensor");
    return -1;
  }
  const int n_units = (int)(2 * n_chunks);
  int32_t* d_flags = b.flags;  // [0] = number of flagged units, [1..] = per-unit overflow flags (workspace)
  const int grid_att = n_units < 2 * s.sm_count ? n_units : 2 * s.sm_count;  // even stride: a CTA keeps its head group
  ++s.attn_calls;
  for (int l = 0; l < w.cfg.decoder_layers; ++l) {
    const BlockDev& bl = w.dec[l];
    CUtensorMap tmWg, tmWfc, tmW1, tmW2;
    ok = make_tmap_2d(enc, &tmWg, bl.wg_h, f16, 2, 192, 64, 96, 64, sw) &&
         make_tmap_2d(enc, &tmWfc, bl.fc_h, f16, 2, 64, 64, 64, 64, sw) &&
         make_tmap_2d(enc, &tmW1, bl.w1_h, f16, 2, 256, 64, 256, 64, sw) &&
         make_tmap_2d(enc, &tmW2, bl.w2_h, f16, 2, 64, 256, 64, 64, sw);
    if (!ok) {
      set_error("cuTensorMapEncodeTiled failed for a weight tensor");
      return -1;
    }
    cudaEvent_t e0 = prof_begin(s, st);
    if (s.attn_exact) {
      k_tc_attn<<<grid_att, 128, kSmemAtt, st>>>(tmX, tmWg, bl.bg, b.o16, n_units, nullptr, nullptr, s.d_status);
    } else {
      // fast kernel (reference = the row's own score, folded into the S MMA) + exact recomputation of the units whose
      // fp16 P overflowed.  Per-layer hint words live behind the status word: when most units of the previous
      // sub-batch overflowed, k_attn_gate sends everything straight to the exact kernel; every 16th call probes again.
      S2S_CUDA_OK(cudaMemsetAsync(d_flags, 0, (size_t)(n_units + 1) * sizeof(int), st));
      int* hint = s.d_status + 8 + l;
      k_attn_gate<<<64, 256, 0, st>>>(hint, s.attn_calls % 16 == 15, n_units, d_flags + 1, d_flags, s.d_status);
      S2S_LAUNCH_CHECK();
      // one 864-thread CTA per SM; an even grid keeps every CTA on one head group
      int grid3 = s.sm_count & ~1;
      if (grid3 > n_units) grid3 = n_units;
      k_tc_attn3<<<grid3, kAttn3Threads, kSmemAtt3, st>>>(tmX, tmWg, bl.bg, b.o16, n_units, d_flags + 1, d_flags, s.d_status);
      S2S_LAUNCH_CHECK();
      k_tc_attn<<<grid_att, 128, kSmemAtt, st>>>(tmX, tmWg, bl.bg, b.o16, n_units, d_flags + 1, d_flags, s.d_status, hint);
    }
    S2S_LAUNCH_CHECK();
    prof_end(s, PROF_ATTN, e0, n_chunks, st);
    e0 = prof_begin(s, st);
    // one 512-thread CTA per SM running four tile pipelines over one copy of the weights (k_tc_ffn4.cuh)
    const int g4 = n_tiles < s.sm_count ? n_tiles : s.sm_count;
    if (l + 1 < w.cfg.decoder_layers)
      k_tc_fc_ffn4<false><<<g4, kFfn4Threads, kSmemFfn4, st>>>(tmO, tmWfc, tmW1, tmW2, tmX, bl.ffn, b.x16, nullptr, epi, n_tiles, s.d_status);
    else
      k_tc_fc_ffn4<true><<<g4, kFfn4Threads, kSmemFfn4, st>>>(tmO, tmWfc, tmW1, tmW2, tmX, bl.ffn, b.x16, nullptr, epi, n_tiles, s.d_status);
    S2S_LAUNCH_CHECK();
    prof_end(s, PROF_FFN, e0, n_chunks, st);
  }
  return 0;
}


// Encoder FFT blocks (modules.py:82-87) on the tensor cores: rows = 16 per chunk.  x32/x16 in place.
int tc_encoder(TcState& s, const DevWeights& w, const TcBuffers& b, float* x32, __half* x16, __half* o16,
               int64_t n_chunks, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(s.encode_tiled);
  const uint64_t rows = (uint64_t)n_chunks * S2S_L_ENC;
  const int n_tiles = (int)((rows + 127) / 128);
  const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tmX, tmO;
  // the buffers are padded to whole 128-row tiles by the workspace carver
  bool ok = make_tmap_2d(enc, &tmX, x16, f16, 2, (uint64_t)n_tiles * 128, 64, 128, 64, sw) &&
            make_tmap_2d(enc, &tmO, o16, f16, 2, (uint64_t)n_tiles * 128, 64, 128, 64, sw);
  if (!ok) {
    set_error("cuTensorMapEncodeTiled failed for an encoder activation tensor");
    return -1;
  }
  const int grid2 = n_tiles < 2 * s.sm_count ? n_tiles : 2 * s.sm_count;
  for (int l = 0; l < w.cfg.encoder_layers; ++l) {
    const BlockDev& bl = w.enc[l];
    CUtensorMap tmWqkv, tmWfc, tmW1, tmW2;
    ok = make_tmap_2d(enc, &tmWqkv, bl.wqkv_h, f16, 2, 192, 64, 192, 64, sw) &&
         make_tmap_2d(enc, &tmWfc, bl.fc_h, f16, 2, 64, 64, 64, 64, sw) &&
         make_tmap_2d(enc, &tmW1, bl.w1_h, f16, 2, 256, 64, 256, 64, sw) &&
         make_tmap_2d(enc, &tmW2, bl.w2_h, f16, 2, 64, 256, 64, 64, sw);
    if (!ok) {
      set_error("cuTensorMapEncodeTiled failed for a weight tensor");
      return -1;
    }
    // q | k | v projection + the 16-key attention (k_tc_enc.cuh), then fc + LN + FFN + LN on the fp32 residual stream
    k_tc_enc_attn<<<grid2, 256, kSmemEncAttn, st>>>(tmX, tmWqkv, bl.bqkv, o16, (int64_t)rows, s.d_status);
    S2S_LAUNCH_CHECK();
    const int g4 = n_tiles < s.sm_count ? n_tiles : s.sm_count;
    k_tc_fc_ffn4<false, true><<<g4, kFfn4Threads, kSmemFfn4, st>>>(tmO, tmWfc, tmW1, tmW2, tmX, bl.ffn, x16, x32, OutEpi{}, n_tiles, s.d_status);
    S2S_LAUNCH_CHECK();
  }
  (void)b;
  return 0;
}

cudaEvent_t prof_begin(TcState& s, cudaStream_t st) {
  if (!s.prof_on) return nullptr;
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  return e;
}

void prof_end(TcState& s, int kind, cudaEvent_t e0, int64_t chunks, cudaStream_t st) {
  if (!s.prof_on || !e0) return;
  cudaEvent_t e1 = nullptr;
  cudaEventCreate(&e1);
  cudaEventRecord(e1, st);
  s.prof_recs.push_back({kind, e0, e1, chunks});
}

int tc_profile_kind(TcState& s, int kind, double* ms_total, int64_t* launches, int64_t* chunks) {
  S2S_CUDA_OK(cudaDeviceSynchronize());
  double tot = 0;
  int64_t n = 0, ch = 0;
  for (const auto& r : s.prof_recs) {
    if (r.kind != kind) continue;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    tot += ms; ++n; ch += r.chunks;
  }
  if (ms_total) *ms_total = tot;
  if (launches) *launches = n;
  if (chunks) *chunks = ch;
  return 0;
}

int tc_profile(TcState& s, int enable, double* ms_total, int64_t* launches, int64_t* chunks) {
  if (enable) {
    s.prof_on = true;
    return 0;
  }
  s.prof_on = false;
  if (tc_profile_kind(s, PROF_ATTN, ms_total, launches, chunks)) return -1;
  for (auto& r : s.prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  s.prof_recs.clear();
  return 0;
}

int tc_debug_counters(int64_t* out, int n, int reset) {
  unsigned long long h[16];
  S2S_CUDA_OK(cudaDeviceSynchronize());
  S2S_CUDA_OK(cudaMemcpyFromSymbol(h, g_phase, sizeof h));
  for (int i = 0; i < n && i < 16; ++i) out[i] = (int64_t)h[i];
  if (reset) {
    memset(h, 0, sizeof h);
    S2S_CUDA_OK(cudaMemcpyToSymbol(g_phase, h, sizeof h));
  }
  return 0;
}

int tc_check_status(TcState& s, cudaStream_t st) {
  int32_t v = 0;
  S2S_CUDA_OK(cudaStreamSynchronize(st));
  S2S_CUDA_OK(cudaMemcpy(&v, s.d_status, 4, cudaMemcpyDeviceToHost));
  if (v != 0) {
    cudaMemset(s.d_status, 0, 4);
    set_error("tensor-core decoder: mbarrier wait timed out (kernel code %d); results are invalid", v);
    return -1;
  }
  return 0;
}

}  // namespace s2s
