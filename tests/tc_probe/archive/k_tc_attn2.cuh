// k_tc_attn2 — pipelined decoder attention (included by k_tc.cu inside namespace s2s::{anonymous}).
//
// Same operand layouts as k_tc_attn (fused QKV projection, masked K, padded V^T with a ones row, fp16 P in TMEM,
// TS-mode P.V), restructured so that a softmax warp does almost nothing but exponentials:
//   * 5 warps: warps 0-3 are the softmax warps (one query row per thread, TMEM lanes 32w..32w+31); warp 4 issues every
//     TMA load and tcgen05.mma and talks to the others only through mbarriers — no __syncthreads in the steady state;
//   * the 250 keys of a (head, query tile) are processed as four 64-key quarters whose scores live in a ring of
//     three 64-column TMEM buffers [S_q, overwritten by P_q]; S(j+2) and P.V(j) run on the tensor pipe while the
//     softmax warps work on quarter j+1, so no MMA round trip (~500 clk) is ever waited for;
//   * ONE reference maximum per row: m_ref = max of the row's first 32 scores.  Softmax is invariant to the reference,
//     m_ref <= the true maximum so the largest probability is >= 1 (nothing underflows), and P = exp2((s - m_ref) c)
//     only misbehaves if it overflows fp16 (s - m_ref > ~31).  Then the row's denominator — accumulated by the
//     tensor core from the same fp16 P through the ones row of V^T — is inf/NaN: the unit is flagged and recomputed
//     by the exact two-pass kernel k_tc_attn.  With a fixed reference all four quarters accumulate into one TMEM
//     accumulator (no per-quarter rescale, no max pass, one tcgen05.ld per score instead of two);
//   * two O accumulators (columns 192.. and 208..) alternate between consecutive (head, tile)s; O is read one quarter
//     after its last P.V was issued.
// Ring barriers complete once per use of a buffer (parity = use count & 1, tracked identically by both roles):
//   bar_S[b]  MMA warp -> softmax : S ready (tcgen05.commit)     bar_P[b]  softmax -> MMA warp: P written (4 warps)
//   bar_PV[b] MMA warp -> both    : P.V of that buffer done -> buffer reusable; for a row's last quarter: O ready
//   bar_OF[a] softmax  -> MMA warp: accumulator a has been read (4 warps)
#pragma once

constexpr int kAttn2Threads = 160;
#ifndef S2S_POLY_EXP
#define S2S_POLY_EXP 0
#endif
constexpr int kPolyExp = S2S_POLY_EXP;  // exponentials per 32 computed on the FMA pipe (fp32 polynomial) instead of MUFU
#ifndef S2S_POLY_EXP_H2
#define S2S_POLY_EXP_H2 5
#endif
constexpr int kPolyExpH2 = S2S_POLY_EXP_H2;  // pairs per 16 computed by the packed-fp16 polynomial instead of MUFU
#ifndef S2S_POLY_BOUND_H2
#define S2S_POLY_BOUND_H2 8
#endif
constexpr int kPolyBoundH2 = S2S_POLY_BOUND_H2;  // the same for the kBound kernel (no scaling FFMAs: more fit)

// kBound exp pass: the accumulator already holds (s - m) c, P = 2^x directly
__device__ __forceinline__ uint32_t ex2_poly_h2_neg(float x0, float x1) {  // x <= ~0: only the lower clamp
  const __half2 kLo = __float2half2_rn(-15.0f), kMagic = __float2half2_rn(1551.0f);
  const __half2 x = __hmax2(__floats2half2_rn(x0, x1), kLo);
  const __half2 t = __hadd2(x, kMagic);
  const __half2 f = __hsub2(x, __hsub2(t, kMagic));
  __half2 p = __hfma2(__float2half2_rn(0.05517167f), f, __float2half2_rn(0.24261113f));
  p = __hfma2(p, f, __float2half2_rn(0.69326097f));
  p = __hfma2(p, f, __float2half2_rn(0.99992806f));
  const uint32_t sc = (*reinterpret_cast<const uint32_t*>(&t) << 10) & 0x7C007C00u;
  const __half2 r = __hmul2(p, *reinterpret_cast<const __half2*>(&sc));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <int kValid, int kPolyH>
__device__ __forceinline__ void chunk_exp_store_direct(const uint32_t (&r)[32], uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (kPolyH > 0 && 2 * i + 1 < kValid && (i * kPolyH) % 16 < kPolyH) {
      pk[i] = ex2_poly_h2_neg(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    } else {
      const float p0 = 2 * i < kValid ? ex2_approx(__uint_as_float(r[2 * i])) : 0.f;
      const float p1 = 2 * i + 1 < kValid ? ex2_approx(__uint_as_float(r[2 * i + 1])) : 0.f;
      pk[i] = pack_half2(p0, p1);
    }
  }
  tmem_st_32x16(taddr, pk);
}

__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

struct Ring3 {  // which of the three S/P buffers the next quarter uses, and the parity of that use
  uint32_t b = 0, bits = 0;
  __device__ __forceinline__ void next(uint32_t& ob, uint32_t& opar) {
    ob = b;
    opar = (bits >> b) & 1u;
    bits ^= 1u << b;
    b = (b == 2u) ? 0u : b + 1u;
  }
};

// kBound = true: the softmax reference of a row is an upper bound known BEFORE the scores exist,
//   m_i = |c q_i| * max_j |k_j|  (Cauchy-Schwarz; c = log2(e)/sqrt(d_k), per head),
// and it is subtracted by the S MMA itself: every head owns a whole K=16 operand slice, A = [c q_i (8) | -m_i, 0 x 7],
// B = [k_j (8) | 1, 0 x 7], so the accumulator already holds (s_ij - m_i) c and the exp pass is one MUFU (or the
// packed polynomial) per score with NO scaling FFMA and no row-max pass: the FMA pipe, which the scaling FFMAs and the
// polynomial share, is what limits how many exponentials can be taken off the MUFU pipe.  P <= 1 always (no overflow);
// if the bound is so loose that the probabilities sink towards the fp16 subnormals the row's denominator (ones row
// of V^T) falls under 2^-9 and the unit is recomputed by the exact kernel, like an overflow in the kBound = false
// scheme (reference = max of the row's first 32 scores, subtracted in the exp pass).
template <bool kBound>
__global__ void __launch_bounds__(kAttn2Threads, 2) k_tc_attn2(const __grid_constant__ CUtensorMap tmX,
                                                               const __grid_constant__ CUtensorMap tmWg,
                                                               const float* __restrict__ bias_g, __half* __restrict__ o16,
                                                               int n_units, int* __restrict__ unit_flags,
                                                               int* __restrict__ n_flagged, int* status) {
  extern __shared__ uint8_t smem_raw[];
  // all barriers in one array: a barrier is addressed as (32-bit shared address of bars) + 8 * index
  enum { B_LOAD = 0, B_W, B_QKV, B_KV, B_UNIT, B_S, B_P = B_S + 3, B_PV = B_P + 3, B_OF = B_PV + 3, B_COUNT = B_OF + 2 };
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[2][96];
  __shared__ float s_kmax[2][4];       // kBound: max_j |k_j|^2 per head of the unit (double-buffered by unit parity)
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sXQ = smem;                 // 2 x [128 x 128 B]: X tiles, then Q (bytes [0,64) of each row; kBound: all 128)
  uint8_t* sK = smem + 2 * kSlab;      // [256 keys x 128 B]  masked K of this head group (4 quarters of 8 KB)
  uint8_t* sV = smem + 4 * kSlab;      // 4 key quarters x [64 rows (4 heads x 16) x 128 B]
  uint8_t* sW = smem + 6 * kSlab;      // [96 x 128 B] weight block of the CTA's head group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // status[0]: a barrier wait timed out somewhere; status[1]: k_attn_gate sent this launch's units to the exact kernel
  if (tid == 0) s_go = (status[0] == 0 && status[1] == 0);
  __syncthreads();
  if (!s_go) return;
  if (warp == 0) tmem_alloc<256>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bars[B_LOAD], 1); mbar_init(&bars[B_W], 1); mbar_init(&bars[B_QKV], 1); mbar_init(&bars[B_KV], 4);
    mbar_init(&bars[B_UNIT], 4);
    for (int b = 0; b < 3; ++b) { mbar_init(&bars[B_S + b], 1); mbar_init(&bars[B_PV + b], 1); mbar_init(&bars[B_P + b], 4); }
    mbar_init(&bars[B_OF], 4); mbar_init(&bars[B_OF + 1], 4);
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWg);
  }
  for (int i = tid; i < 192; i += kAttn2Threads) s_bias[i / 96][i % 96] = bias_g[i];
  // V^T padding rows are constant: row 8 of every head = ones (softmax denominator), rows 9..15 = 0
  for (int i = tid; i < 2 * kSlab / 16; i += kAttn2Threads) {
    reinterpret_cast<uint4*>(sV)[i] = make_uint4(0u, 0u, 0u, 0u);
    // kBound = false: the masked (zero) half of every K slot never changes.  kBound = true: the second half of every
    // 32-byte slot is the constant (1, 0 x 7) that picks up -m_i from the A operand (16-byte chunk index is odd; the
    // SW128 swizzle only permutes chunks within a row, so "odd chunk" can be tested on the physical index XOR row).
    uint32_t w0 = 0u;
    if (kBound) {
      const uint32_t row = (uint32_t)i >> 3, phys = (uint32_t)i & 7u;
      if (((phys ^ (row & 7u)) & 1u) != 0u) w0 = 0x00003C00u;  // fp16 1.0 in element 0
    }
    reinterpret_cast<uint4*>(sK)[i] = make_uint4(w0, 0u, 0u, 0u);
  }
  if (tid < 8) s_kmax[tid >> 2][tid & 3] = 0.f;
  __syncthreads();
  for (int i = tid; i < 4 * 4 * 8; i += kAttn2Threads) {  // (quarter, head, 16-byte chunk of 8 keys)
    const int slab = i >> 5, hh = (i >> 3) & 3, ck = i & 7;
    const uint32_t one2 = 0x3C003C00u;  // two fp16 ones
    *reinterpret_cast<uint4*>(sV + slab * 8192 + sw128_offset(hh * 16 + 8, ck)) = make_uint4(one2, one2, one2, one2);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t bar0 = smem_u32(&bars[0]), abort_a = smem_u32(&s_abort);
  auto BAR = [&](uint32_t idx) { return bar0 + 8u * idx; };
  // bounded wait on a barrier address; the abort flag is only consulted on the slow path
  auto wait_a = [&](uint32_t a, uint32_t parity, int code) -> bool {
    for (uint32_t i = 0; i < kWaitLimit; ++i) {
      if (mbar_try_wait_a(a, parity)) return true;
      if ((i & 255u) == 255u && lds_u32(abort_a)) return false;
    }
    sts_u32(abort_a, 1u);
    atomicExch(status, code);
    return false;
  };
  auto warp_arrive_a = [&](uint32_t a) {
    __syncwarp();
    if (lane == 0) mbar_arrive_a(a);
  };
  const float kScale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e) / sqrt(d_k)
  constexpr uint32_t kOaccCol = 192;
  Ring3 ring;
  PHASE_DECL

  if (warp == 4) {
    // =============================== TMA + MMA issue warp =======================================
    const uint32_t idesc_qkv = umma_idesc(128, 96, kFmtF16), idesc_s = umma_idesc(128, 64, kFmtF16),
                   idesc_o = umma_idesc(128, 16, kFmtF16);
    const uint32_t aXQ = smem_u32(sXQ), aW = smem_u32(sW);
    const uint64_t dXQ = umma_desc_k_sw128(aXQ), dK = umma_desc_k_sw128(smem_u32(sK)), dV = umma_desc_k_sw128(smem_u32(sV));
    uint32_t it = 0, ph_w = 0;
    int cur_g = -1;
    bool x_prefetched = false;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int chunk = unit >> 1, g = unit & 1;
      const uint32_t upar = it & 1;
      if (elect_one()) {
        if (g != cur_g) {  // with an even grid stride every CTA keeps its head group: loaded once
          mbar_arrive_expect_tx(&bars[B_W], 96 * 128);
          tma_load_2d(sW, &tmWg, &bars[B_W], 0, g * 96);
        }
        if (!x_prefetched) {
          mbar_arrive_expect_tx(&bars[B_LOAD], 2 * kSlab);
          tma_load_2d(sXQ, &tmX, &bars[B_LOAD], 0, chunk * 256);
          tma_load_2d(sXQ + kSlab, &tmX, &bars[B_LOAD], 0, chunk * 256 + 128);
        }
      }
      x_prefetched = false;
      if (g != cur_g) {
        wait_a(BAR(B_W), ph_w, kErrAttLoad);
        ph_w ^= 1;
        cur_g = g;
      }
      wait_a(BAR(B_LOAD), upar, kErrAttLoad);
      tcgen05_fence_after();
      if (elect_one()) {  // [128 x 96] = X_tile Wg^T, both tiles (accumulators at columns 0 and 128)
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int s = 0; s < 4; ++s)
            umma_f16_ss(tmem + tile * 128, umma_desc_k_sw128(aXQ + tile * kSlab + s * 32), umma_desc_k_sw128(aW + s * 32),
                        idesc_qkv, s > 0);
        umma_commit_a(BAR(B_QKV));
      }
      wait_a(BAR(B_KV), upar, kErrAttS);  // Q / K / V^T operands are in shared memory
      tcgen05_fence_after();
      // S for quarter j: A = Q slice of (tile, head pair), B = masked-K slot of the head, rows of key quarter q
      auto issue_S = [&](int j, uint32_t buf) {
        const int M = j >> 2, q = j & 3, hh = M >> 1, tile = M & 1;
        umma_f16_ss(tmem + 64 * buf, dXQ + (uint64_t)((tile * kSlab + (kBound ? hh : (hh >> 1)) * 32) >> 4),
                    dK + (uint64_t)((q * 8192 + hh * 32) >> 4), idesc_s, 0);
        umma_commit_a(BAR(B_S + buf));
      };
      {  // all three ring buffers are free at the start of a unit: S of quarters 0, 1, 2
        const uint32_t b0 = ring.b, b1 = (b0 == 2u) ? 0u : b0 + 1u, b2 = (b1 == 2u) ? 0u : b1 + 1u;
        if (elect_one()) { issue_S(0, b0); issue_S(1, b1); issue_S(2, b2); }
      }
#pragma unroll 1
      for (int j = 0; j < 32; ++j) {
        const int M = j >> 2, q = j & 3, hh = M >> 1, acc_i = M & 1;
        uint32_t b, par;
        ring.next(b, par);
        if (q == 0 && (it > 0 || M >= 2)) {  // accumulator acc_i still holds the O of two (head, tile)s ago until it is read
          const uint32_t use = 4u * it + (uint32_t)(M >> 1);
          wait_a(BAR(B_OF + acc_i), (use - 1u) & 1u, kErrAttO);
        }
        wait_a(BAR(B_P + b), par, kErrAttO);
        tcgen05_fence_after();
        PHASE(10);
        if (elect_one()) {  // O += P_q V_h over the quarter's 64 keys: 4 K-steps, A operand straight from TMEM
          const uint64_t dVq = dV + (uint64_t)((q * 8192 + hh * 2048) >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16_ts(tmem + kOaccCol + 16 * acc_i, tmem + 64 * b + 8 * ks, dVq + (uint64_t)((ks * 32) >> 4), idesc_o,
                        (q > 0 || ks > 0) ? 1u : 0u);
          umma_commit_a(BAR(B_PV + b));
        }
        PHASE(11);
        if (j + 3 < 32) {  // this buffer is free again as soon as its P.V has completed: S three quarters ahead goes in
          wait_a(BAR(B_PV + b), par, kErrAttS);
          tcgen05_fence_after();
          PHASE(12);
          if (elect_one()) issue_S(j + 3, b);
          PHASE(13);
          if (j + 3 == 31) {
            // The unit's last S has been issued.  Once it completes nothing reads the Q tiles any more, so the next
            // unit's X tiles can stream into sXQ during the last three quarters (hides the ~1.5k clk TMA latency).
            const int next_unit = unit + (int)gridDim.x;
            if (next_unit < n_units && (next_unit & 1) == g) {
              wait_a(BAR(B_S + b), par ^ 1u, kErrAttLoad);
              if (elect_one()) {
                const int nchunk = next_unit >> 1;
                mbar_arrive_expect_tx(&bars[B_LOAD], 2 * kSlab);
                tma_load_2d(sXQ, &tmX, &bars[B_LOAD], 0, nchunk * 256);
                tma_load_2d(sXQ + kSlab, &tmX, &bars[B_LOAD], 0, nchunk * 256 + 128);
              }
              x_prefetched = true;
            }
          }
        }
      }
      // every softmax warp has read its last O: shared-memory operands and TMEM may be overwritten
      wait_a(BAR(B_UNIT), upar, kErrAttO);
      tcgen05_fence_after();
    }
  } else {
    // =============================== softmax warps ===============================================
    const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int chunk = unit >> 1, g = unit & 1;
      const uint32_t upar = it & 1;
      PHASE_COUNT(15);
      wait_a(BAR(B_QKV), upar, kErrAttS);
      tcgen05_fence_after();
      PHASE(1);
      if constexpr (!kBound) {
        // QKV epilogue: accumulators -> fp16 operands in shared memory.  Only Q needs its bias here: the K bias adds a
        // per-row constant q.b_k to every score (softmax-invariant), and the V bias is added once to the normalised
        // output (sum_j p_j (v_j + b_v) = sum_j p_j v_j + b_v).  The zero halves of the masked K slots are static.
        const float* bq = s_bias[g];
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const int t = tile * 128 + tid;  // key / query index inside the chunk
          uint32_t rq[32], rk[32], rv[32];
          tmem_ld_32x32(lane_addr + tile * 128, rq);
          tmem_ld_32x32(lane_addr + tile * 128 + 32, rk);
          tmem_ld_32x32(lane_addr + tile * 128 + 64, rv);
          tmem_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            uint32_t pq[4], pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pq[i] = pack_half2(__uint_as_float(rq[8 * hh + 2 * i]) + bq[8 * hh + 2 * i],
                                 __uint_as_float(rq[8 * hh + 2 * i + 1]) + bq[8 * hh + 2 * i + 1]);
              pk[i] = pack_half2(__uint_as_float(rk[8 * hh + 2 * i]), __uint_as_float(rk[8 * hh + 2 * i + 1]));
            }
            *reinterpret_cast<uint4*>(sXQ + tile * kSlab + sw128_offset(tid, hh)) = make_uint4(pq[0], pq[1], pq[2], pq[3]);
            *reinterpret_cast<uint4*>(sK + sw128_offset(t, 2 * hh + (hh & 1))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          uint8_t* vslab = sV + (t >> 6) * 8192 + (t & 7) * 2;
          const uint32_t ck = (t & 63) >> 3;
#pragma unroll
          for (int hh = 0; hh < 4; ++hh)
#pragma unroll
            for (int d = 0; d < 8; ++d)
              *reinterpret_cast<__half*>(vslab + sw128_offset(hh * 16 + d, ck)) = __float2half_rn(__uint_as_float(rv[8 * hh + d]));
        }
      } else {
        // kBound epilogue, two passes.  Pass 1 (both tiles): K (first half of each head's 32-byte slot; no bias, see
        // above) and V^T to shared memory, and max_j |k_j|^2 per head over the chunk's 256 keys (warp max, then one
        // shared atomicMax per warp; non-negative floats order like their bit patterns).
        const uint32_t kp = upar;
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const int t = tile * 128 + tid;
          uint32_t rk[32], rv[32];
          tmem_ld_32x32(lane_addr + tile * 128 + 32, rk);
          tmem_ld_32x32(lane_addr + tile * 128 + 64, rv);
          tmem_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            uint32_t pk[4];
            float n2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float k0 = __uint_as_float(rk[8 * hh + 2 * i]), k1 = __uint_as_float(rk[8 * hh + 2 * i + 1]);
              pk[i] = pack_half2(k0, k1);
              n2 = fmaf(k0, k0, fmaf(k1, k1, n2));
            }
            *reinterpret_cast<uint4*>(sK + sw128_offset(t, 2 * hh)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(n2));
            if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(&s_kmax[kp][hh]), wmax);
          }
          uint8_t* vslab = sV + (t >> 6) * 8192 + (t & 7) * 2;
          const uint32_t ck = (t & 63) >> 3;
#pragma unroll
          for (int hh = 0; hh < 4; ++hh)
#pragma unroll
            for (int d = 0; d < 8; ++d)
              *reinterpret_cast<__half*>(vslab + sw128_offset(hh * 16 + d, ck)) = __float2half_rn(__uint_as_float(rv[8 * hh + d]));
        }
        if (tid < 4) s_kmax[kp ^ 1u][tid] = 0.f;             // the other parity's slots: free since the previous unit
        asm volatile("bar.sync 1, 128;" ::: "memory");       // the four softmax warps only (warp 4 never joins)
        // Pass 2: Q scaled by c = log2(e)/sqrt(d_k), and -m_i = -|c q_i| max_j|k_j| in element 8 of the head's slice
        const float* bq = s_bias[g];
        float kmax2[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) kmax2[hh] = s_kmax[kp][hh] * 1.002f;  // fp16 rounding of q and k: keep it a bound
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          uint32_t rq[32];
          tmem_ld_32x32(lane_addr + tile * 128, rq);
          tmem_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            uint32_t pq[4];
            float n2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float q0 = (__uint_as_float(rq[8 * hh + 2 * i]) + bq[8 * hh + 2 * i]) * kScale;
              const float q1 = (__uint_as_float(rq[8 * hh + 2 * i + 1]) + bq[8 * hh + 2 * i + 1]) * kScale;
              pq[i] = pack_half2(q0, q1);
              n2 = fmaf(q0, q0, fmaf(q1, q1, n2));
            }
            const float m = sqrtf(n2 * kmax2[hh]);
            *reinterpret_cast<uint4*>(sXQ + tile * kSlab + sw128_offset(tid, 2 * hh)) = make_uint4(pq[0], pq[1], pq[2], pq[3]);
            *reinterpret_cast<uint4*>(sXQ + tile * kSlab + sw128_offset(tid, 2 * hh + 1)) =
                make_uint4(pack_half2(-m, 0.f), 0u, 0u, 0u);
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      warp_arrive_a(BAR(B_KV));
      PHASE(2);

      bool overflow = false;
      uint32_t ob = 0, opar = 0;  // ring slot of the pending row's last quarter: its bar_PV is "O ready"
      uint32_t o[16];
      // O of (head, tile) M: wait for its last P.V, start the TMEM read ...
      auto take_O_issue = [&]() {
        wait_a(BAR(B_PV + ob), opar, kErrAttO);
        tcgen05_fence_after();
      };
      // ... and, once a tcgen05.wait::ld has covered it, release the accumulator, normalise, add the V bias, store.
      auto take_O_finish = [&](int M) {
        const int hh = M >> 1, tile = M & 1, acc_i = M & 1;
        tcgen05_fence_before();
        warp_arrive_a(BAR(B_OF + acc_i));
        const float den = __uint_as_float(o[8]);  // sum of the rounded probabilities (ones row of V^T)
        overflow |= !(den < 1e30f);                // inf / NaN: some P overflowed fp16 -> exact kernel redoes the unit
        if (kBound) overflow |= den < 0.001953125f;  // loose bound: probabilities near the fp16 subnormals
        const float inv = 1.0f / den;
        const float* bv = s_bias[g] + 64 + 8 * hh;
        const int64_t row = (int64_t)chunk * 256 + tile * 128 + tid;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(__uint_as_float(o[i]), inv, bv[i]);
        *reinterpret_cast<uint4*>(o16 + row * 64 + (g * 4 + hh) * 8) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
      };
      // Software-pipelined quarter loop: S(j+1) is probed (non-blocking) and its first 32 columns are loaded in the
      // middle of quarter j, the other 32 right after the last exponential of quarter j, so that a quarter starts
      // with its scores already in registers.
      uint32_t cb, cpar;
      ring.next(cb, cpar);
      wait_a(BAR(B_S + cb), cpar, kErrAttS);
      tcgen05_fence_after();
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(lane_addr + 64 * cb, ra);
      tmem_ld_32x32(lane_addr + 64 * cb + 32, rb);
      PHASE(3);
#pragma unroll 1
      for (int M = 0; M < 8; ++M) {
        float mneg = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool has_next = !(M == 7 && q == 3);
          uint32_t nb = 0, npar = 0;
          if (has_next) ring.next(nb, npar);
          const uint32_t col = lane_addr + 64 * cb, ncol = lane_addr + 64 * nb;
          tmem_wait_ld();
          if (q == 1 && M > 0) take_O_finish(M - 1);  // its tcgen05.ld was issued in the middle of the previous quarter
          if (!kBound && q == 0) mneg = -chunk_max<32>(ra, -INFINITY) * kScale;  // reference: max of the first 32 scores
          PHASE(4);
          if (kBound) chunk_exp_store_direct<32, kPolyBoundH2>(ra, col);
          else chunk_exp_store_mixed<32, kPolyExp, kPolyExpH2>(ra, kScale, mneg, col);
          if (q == 0 && M > 0) {  // O of the previous (head, tile): its last P.V was issued most of a quarter ago
            take_O_issue();
            tmem_ld_32x16(lane_addr + kOaccCol + 16 * ((M - 1) & 1), o);
          }
          bool ready = false;
          if (has_next) {
            ready = __all_sync(0xffffffffu, mbar_test_wait_a(BAR(B_S + nb), npar));
            if (ready) {
              tcgen05_fence_after();
              tmem_ld_32x32(ncol, ra);
            }
          }
          if (kBound) {
            if (q == 3) chunk_exp_store_direct<S2S_L_DEC - 224, kPolyBoundH2>(rb, col + 16);
            else chunk_exp_store_direct<32, kPolyBoundH2>(rb, col + 16);
          } else {
            if (q == 3) chunk_exp_store_mixed<S2S_L_DEC - 224, kPolyExp, kPolyExpH2>(rb, kScale, mneg, col + 16);
            else chunk_exp_store_mixed<32, kPolyExp, kPolyExpH2>(rb, kScale, mneg, col + 16);
          }
          PHASE(5);
          if (has_next) {
            if (!ready) {
              wait_a(BAR(B_S + nb), npar, kErrAttS);
              tcgen05_fence_after();
              tmem_ld_32x32(ncol, ra);
            }
            tmem_ld_32x32(ncol + 32, rb);
          }
          PHASE(3);
          tmem_wait_st();
          tcgen05_fence_before();
          warp_arrive_a(BAR(B_P + cb));
          if (q == 3) { ob = cb; opar = cpar; }
          cb = nb; cpar = npar;
          PHASE(6);
        }
      }
      take_O_issue();
      tmem_ld_32x16(lane_addr + kOaccCol + 16, o);
      tmem_wait_ld();
      take_O_finish(7);
      if (__any_sync(0xffffffffu, overflow) && lane == 0) {
        unit_flags[unit] = 1;
        atomicAdd(n_flagged, 1);
      }
      tcgen05_fence_before();
      warp_arrive_a(BAR(B_UNIT));
      PHASE(7);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  PHASE_FLUSH;
  if (warp == 0) tmem_dealloc<256>(tmem);
}
