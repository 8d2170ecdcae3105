// k_tc_attn4 — decoder attention with FOUR softmax warps per scheduler (included by k_tc.cu inside s2s::{anonymous}).
//
// Same unit (chunk x group of 4 heads), operand layouts, single-reference softmax and overflow fallback as k_tc_attn2;
// what changes is the occupancy of the MUFU pipe, which bounds this kernel (d_k = 8: 32 MMA-FLOPs per exponential):
//   * 9 warps per CTA, 2 CTAs per SM: warps 0-3 = softmax group 0 (query tile 0 of the unit), warps 4-7 = softmax
//     group 1 (query tile 1), warp 8 = TMA + tcgen05.mma issue.  Both groups share the unit's K / V^T operands in
//     shared memory, so 16 softmax warps per SM (4 per scheduler) cost the shared memory of 8: when a warp waits for
//     a tcgen05.ld, an MMA round trip or the QKV epilogue, three others keep the scheduler's MUFU and FMA pipes busy;
//   * TMEM (256 columns per CTA): group g owns columns [128g, 128g+128): a ring of THREE 32-key score buffers
//     [S, overwritten by fp16 P] at +0, +32, +64, two O accumulators at +96 and +112.  The 250 keys of a (head, tile)
//     are eight 32-key steps (the last with 26 valid keys); S(j+3) is issued as soon as P.V(j) has released its
//     buffer, i.e. two whole steps before the softmax warps need it (a ring of two 48-key buffers was measured: the
//     P(j) -> P.V(j) done -> S(j+2) chain of ~1000 clk is as long as a step, and the softmax warps starved).  The ring
//     restarts at buffer 0 in every unit; the QKV accumulators of the next unit ([128 x 96] per tile) reuse the ring
//     columns of the group that owns the tile;
//   * one MMA issue warp per group (warps 8, 9), fully unrolled over (head, step) so that every MMA operand is an
//     immediate offset from a uniform base; warp 8 also owns the TMA loads and the QKV projection;
//   * every softmax group converts its own tile in the QKV epilogue (half the time of k_tc_attn2's epilogue).
// Barriers (parity = use count & 1; per unit buffers 0 and 1 are used 11 times and buffer 2 ten times, so the parity of
// use u of buffer b in the it-th unit of a CTA is (u + (b < 2 ? it : 0)) & 1):
//   B_S[g][b]  MMA warp -> group g : S in buffer b ready           B_P[g][b]  group g -> MMA warp: P written (4 warps)
//   B_PV[g][b] MMA warp -> both    : P.V of buffer b complete       B_OF[g][a] group g -> MMA warp: accumulator a read
//   B_QKV      MMA warp -> all     : projection accumulators ready  B_KV       8 softmax warps -> MMA warp: operands in smem
#pragma once

constexpr int kAttn4Threads = 320;
#if defined(S2S_PHASE_TIMING) && S2S_PHASE_TIMING == 1
// lane 0 of softmax warp 0 and of the MMA warp add clock64() deltas to shared counters (flushed to g_phase at exit)
#define PH4_DECL __shared__ unsigned long long s_ph[16]; if (threadIdx.x < 16) s_ph[threadIdx.x] = 0; \
  const bool ph_on = (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == 8); long long ph_t = clock64();
#define PH4(i) do { if (ph_on) { long long n_ = clock64(); atomicAdd(&s_ph[i], (unsigned long long)(n_ - ph_t)); ph_t = n_; } } while (0)
#define PH4_COUNT(i) do { if (ph_on) atomicAdd(&s_ph[i], 1ull); } while (0)
#define PH4_FLUSH do { if (threadIdx.x < 16 && s_ph[threadIdx.x]) atomicAdd(&g_phase[threadIdx.x], s_ph[threadIdx.x]); } while (0)
#else
#define PH4_DECL
#define PH4(i) do {} while (0)
#define PH4_COUNT(i) do {} while (0)
#define PH4_FLUSH do {} while (0)
#endif
#ifndef S2S_ATTN4_SLEEP
#define S2S_ATTN4_SLEEP 20
#endif
#ifndef S2S_POLY4_H2
#define S2S_POLY4_H2 6
#endif
constexpr int kPoly4H2 = S2S_POLY4_H2;  // pairs per 16 computed by the packed-fp16 polynomial instead of MUFU

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// P = 2^(s * scale + mneg) for kN score columns (the first kValid are real keys), fp16 pairs -> TMEM at taddr
template <int kN, int kValid, int kPolyH>
__device__ __forceinline__ void exp_store_n(const uint32_t (&r)[kN], float scale, float mneg, uint32_t taddr) {
  static_assert(kN == 32 || kN == 16, "32- or 16-column pieces");
  uint32_t pk[kN / 2];
#pragma unroll
  for (int i = 0; i < kN / 2; ++i) {
    if (kPolyH > 0 && 2 * i + 1 < kValid && (i * kPolyH) % 16 < kPolyH) {
      pk[i] = ex2_poly_h2(fmaf(__uint_as_float(r[2 * i]), scale, mneg), fmaf(__uint_as_float(r[2 * i + 1]), scale, mneg));
    } else {
      const float p0 = 2 * i < kValid ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), scale, mneg)) : 0.f;
      const float p1 = 2 * i + 1 < kValid ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), scale, mneg)) : 0.f;
      pk[i] = pack_half2(p0, p1);
    }
  }
  if constexpr (kN == 32) tmem_st_32x16(taddr, pk); else tmem_st_32x8(taddr, pk);
}

__global__ void __launch_bounds__(kAttn4Threads, 2) k_tc_attn4(const __grid_constant__ CUtensorMap tmX,
                                                               const __grid_constant__ CUtensorMap tmWg,
                                                               const float* __restrict__ bias_g, __half* __restrict__ o16,
                                                               int n_units, int* __restrict__ unit_flags,
                                                               int* __restrict__ n_flagged, int* status) {
  extern __shared__ uint8_t smem_raw[];
  // all barriers in one array: a barrier is addressed as (32-bit shared address of bars) + 8 * index
  enum { B_LOAD = 0, B_W, B_QKV, B_KV, B_SDONE, B_GDONE, B_S, B_P = B_S + 6, B_PV = B_P + 6, B_OF = B_PV + 6, B_COUNT = B_OF + 4 };
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[2][96];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sXQ = smem;                 // 2 x [128 x 128 B]: X tiles, then Q (bytes [0,64) of each row)
  uint8_t* sK = smem + 2 * kSlab;      // [256 keys x 128 B]  masked K of this head group
  uint8_t* sV = smem + 4 * kSlab;      // 4 slabs of 64 keys x [64 rows (4 heads x 16) x 128 B]
  uint8_t* sW = smem + 6 * kSlab;      // [96 x 128 B] weight block of the CTA's head group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_go = (*status == 0);
  __syncthreads();
  if (!s_go) return;
  PH4_DECL
  if (warp == 0) tmem_alloc<256>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bars[B_LOAD], 1); mbar_init(&bars[B_W], 1); mbar_init(&bars[B_QKV], 1); mbar_init(&bars[B_KV], 8);
    mbar_init(&bars[B_SDONE], 2); mbar_init(&bars[B_GDONE], 2);
    for (int i = 0; i < 6; ++i) { mbar_init(&bars[B_S + i], 1); mbar_init(&bars[B_PV + i], 1); mbar_init(&bars[B_P + i], 4); }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[B_OF + i], 4);
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWg);
  }
  for (int i = tid; i < 192; i += kAttn4Threads) s_bias[i / 96][i % 96] = bias_g[i];
  // V^T padding rows are constant: row 8 of every head = ones (softmax denominator), rows 9..15 = 0; the masked (zero)
  // half of every K slot never changes either
  for (int i = tid; i < 2 * kSlab / 16; i += kAttn4Threads) {
    reinterpret_cast<uint4*>(sV)[i] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(sK)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int i = tid; i < 4 * 4 * 8; i += kAttn4Threads) {  // (slab, head, 16-byte chunk of 8 keys)
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
  auto test_all = [&](uint32_t a, uint32_t parity) -> bool {
    return __all_sync(0xffffffffu, mbar_test_wait_a(a, parity));
  };
  const float kScale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e) / sqrt(d_k)

  if (warp >= 8) {
    // =============================== MMA issue warps (one per softmax group) =====================
    // warp 8 serves group 0 and also owns the TMA loads and the QKV projection; warp 9 serves group 1.  Every step index
    // below is a compile-time constant after unrolling, so descriptors are immediate adds on a uniform base.
    const int g = warp - 8;
    const uint32_t idesc_qkv = umma_idesc(128, 96, kFmtF16), idesc_s = umma_idesc(128, 32, kFmtF16),
                   idesc_o = umma_idesc(128, 16, kFmtF16);
    const uint32_t aXQ = smem_u32(sXQ), aW = smem_u32(sW);
    const uint64_t dQ = umma_desc_k_sw128(aXQ + g * kSlab), dK = umma_desc_k_sw128(smem_u32(sK)),
                   dV = umma_desc_k_sw128(smem_u32(sV));
    const uint32_t tg = tmem + 128 * g;
    const uint32_t bS = B_S + 3 * g, bP = B_P + 3 * g, bPV = B_PV + 3 * g, bOF = B_OF + 2 * g;
    // S of (head hh, key step s) into ring buffer b: A = Q slice of (tile g, head pair), B = masked-K slots of 32 keys
    auto issue_S = [&](int hh, int s, int b) {
      umma_f16_ss(tg + 32 * b, dQ + (uint64_t)((hh >> 1) * 2), dK + (uint64_t)(s * 256 + hh * 2), idesc_s, 0);
      umma_commit_a(BAR(bS + b));
    };
    uint32_t it = 0, ph_w = 0;
    int cur_g = -1;
    bool x_prefetched = false;
    bool alive = true;
    for (int unit = blockIdx.x; unit < n_units && alive; unit += gridDim.x, ++it) {
      const int chunk = unit >> 1, hg = unit & 1;
      const uint32_t upar = it & 1;
      const int next_unit = unit + (int)gridDim.x;
      const bool want_prefetch = next_unit < n_units && (next_unit & 1) == hg;
      bool pf_pending = false;
      if (g == 0) {
        if (elect_one()) {
          if (hg != cur_g) {  // with an even grid stride every CTA keeps its head group: loaded once
            mbar_arrive_expect_tx(&bars[B_W], 96 * 128);
            tma_load_2d(sW, &tmWg, &bars[B_W], 0, hg * 96);
          }
          if (!x_prefetched) {
            mbar_arrive_expect_tx(&bars[B_LOAD], 2 * kSlab);
            tma_load_2d(sXQ, &tmX, &bars[B_LOAD], 0, chunk * 256);
            tma_load_2d(sXQ + kSlab, &tmX, &bars[B_LOAD], 0, chunk * 256 + 128);
          }
        }
        x_prefetched = false;
        if (hg != cur_g) {
          alive = alive && wait_a(BAR(B_W), ph_w, kErrAttLoad);
          ph_w ^= 1;
          cur_g = hg;
        }
        alive = alive && wait_a(BAR(B_LOAD), upar, kErrAttLoad);
        // both groups' P.V of the previous unit are complete: ring columns and shared-memory operands are free
        if (it > 0) alive = alive && wait_a(BAR(B_GDONE), (it - 1) & 1, kErrAttO);
        tcgen05_fence_after();
        if (elect_one()) {  // [128 x 96] = X_tile Wg^T, both tiles (accumulators in the ring columns of the tile's group)
#pragma unroll
          for (int tile = 0; tile < 2; ++tile)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ss(tmem + tile * 128, umma_desc_k_sw128(aXQ + tile * kSlab + ks * 32),
                          umma_desc_k_sw128(aW + ks * 32), idesc_qkv, ks > 0);
          umma_commit_a(BAR(B_QKV));
        }
      }
      alive = alive && wait_a(BAR(B_KV), upar, kErrAttS);  // Q / K / V^T operands are in shared memory
      tcgen05_fence_after();
      PH4(14);
      if (elect_one()) { issue_S(0, 0, 0); issue_S(0, 1, 1); issue_S(0, 2, 2); }
      // fully unrolled over (head, step): every MMA operand is an immediate offset from a uniform base.  With run-time
      // head indices the issuing thread spends ~150 clk per MMA on integer math and R2UR moves (measured), which made
      // the MMA warp, not the MUFU pipe, the bottleneck; constant operands issue in ~15 (TS) / ~45 (SS) clk.
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int j = 8 * hh + s, b = j % 3, u = j / 3;
          const uint32_t par = ((uint32_t)u + (b < 2 ? it : 0u)) & 1u;
          // first P.V of a head: the accumulator still holds the O of two heads ago until the group has read it
          if (s == 0 && (it > 0 || hh >= 2)) alive = alive && wait_a(BAR(bOF + (hh & 1)), (uint32_t)((hh >> 1) + 1) & 1u, kErrAttO);
          alive = alive && wait_a(BAR(bP + b), par, kErrAttO);
          tcgen05_fence_after();
          PH4(10);
          if (elect_one()) {  // O += P V_h over the 32 keys of the step: two TS MMAs, A (fp16 P) straight from TMEM
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const int kb = 2 * s + ks;  // 16-key block of the chunk
              umma_f16_ts(tg + 96 + 16 * (hh & 1), tg + 32 * b + 8 * ks,
                          dV + (uint64_t)((kb >> 2) * 512 + (kb & 3) * 2 + hh * 128), idesc_o, (s > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit_a(BAR(bPV + b));
            if (j == 31) umma_commit_a(BAR(B_GDONE));  // the unit's last P.V of this group
          }
          PH4(11);
          if (j + 3 < 32) {  // buffer b is free again once its P.V has completed: S three steps ahead goes in
            alive = alive && wait_a(BAR(bPV + b), par, kErrAttS);
            tcgen05_fence_after();
            PH4(12);
            if (elect_one()) {
              issue_S((j + 3) / 8, (j + 3) % 8, b);
              if (j + 3 == 31) umma_commit_a(BAR(B_SDONE));  // the unit's last S of this group
            }
            PH4(13);
            if (g == 0 && j + 3 == 31) pf_pending = want_prefetch;
          } else if (g == 0 && pf_pending) {
            // Both groups' last S have been issued.  Once they complete nothing reads the Q tiles any more, so the next
            // unit's X tiles can stream into sXQ during the last steps (hides the ~1.5k clk TMA latency).
            const bool ok = j == 31 ? wait_a(BAR(B_SDONE), upar, kErrAttLoad) : test_all(BAR(B_SDONE), upar);
            if (ok) {
              if (elect_one()) {
                const int nchunk = next_unit >> 1;
                mbar_arrive_expect_tx(&bars[B_LOAD], 2 * kSlab);
                tma_load_2d(sXQ, &tmX, &bars[B_LOAD], 0, nchunk * 256);
                tma_load_2d(sXQ + kSlab, &tmX, &bars[B_LOAD], 0, nchunk * 256 + 128);
              }
              x_prefetched = true;
              pf_pending = false;
            }
          }
        }
      }
    }
  } else {
    // =============================== softmax warps ===============================================
    const int g = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;              // row inside the group's query tile
    const int t = g * 128 + r;                 // key / query index inside the chunk
    const uint32_t lane_addr = tmem_addr(tmem, wq * 32, 128 * g);
    const uint32_t bS = B_S + 3 * g, bP = B_P + 3 * g, bPV = B_PV + 3 * g, bOF = B_OF + 2 * g;
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int chunk = unit >> 1, hg = unit & 1;
      const uint32_t upar = it & 1;
      PH4_COUNT(15);
      if (!wait_a(BAR(B_QKV), upar, kErrAttS)) break;
      tcgen05_fence_after();
      PH4(1);
      {  // QKV epilogue of the group's own tile: accumulators -> fp16 operands in shared memory.  Only Q needs its bias
         // here: the K bias adds a per-row constant q.b_k to every score (softmax-invariant), and the V bias is added
         // once to the normalised output.  The zero halves of the masked K slots are static.
        const float* bq = s_bias[hg];
        uint32_t rr[32];
        tmem_ld_32x32(lane_addr, rr);
        tmem_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t pq[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pq[i] = pack_half2(__uint_as_float(rr[8 * hh + 2 * i]) + bq[8 * hh + 2 * i],
                               __uint_as_float(rr[8 * hh + 2 * i + 1]) + bq[8 * hh + 2 * i + 1]);
          *reinterpret_cast<uint4*>(sXQ + g * kSlab + sw128_offset(r, hh)) = make_uint4(pq[0], pq[1], pq[2], pq[3]);
        }
        tmem_ld_32x32(lane_addr + 32, rr);
        tmem_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pk[i] = pack_half2(__uint_as_float(rr[8 * hh + 2 * i]), __uint_as_float(rr[8 * hh + 2 * i + 1]));
          *reinterpret_cast<uint4*>(sK + sw128_offset(t, 2 * hh + (hh & 1))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        tmem_ld_32x32(lane_addr + 64, rr);
        tmem_wait_ld();
        uint8_t* vslab = sV + (t >> 6) * 8192 + (t & 7) * 2;
        const uint32_t ck = (t & 63) >> 3;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
#pragma unroll
          for (int d = 0; d < 8; ++d)
            *reinterpret_cast<__half*>(vslab + sw128_offset(hh * 16 + d, ck)) = __float2half_rn(__uint_as_float(rr[8 * hh + d]));
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      warp_arrive_a(BAR(B_KV));
      PH4(2);

      bool overflow = false, alive = true;
      uint32_t ob = 0, opar = 0;  // ring slot / parity of the pending head's last step: its B_PV is "O ready"
      uint32_t o[16];
      auto take_O_issue = [&]() {
        alive = alive && wait_a(BAR(bPV + ob), opar, kErrAttO);
        tcgen05_fence_after();
      };
      // once a tcgen05.wait::ld has covered the O load: release the accumulator, normalise, add the V bias, store
      auto take_O_finish = [&](int hh) {
        tcgen05_fence_before();
        warp_arrive_a(BAR(bOF + (hh & 1)));
        const float den = __uint_as_float(o[8]);  // sum of the rounded probabilities (ones row of V^T)
        overflow |= !(den < 1e30f);                // inf / NaN: some P overflowed fp16 -> exact kernel redoes the unit
        const float inv = 1.0f / den;
        const float* bv = s_bias[hg] + 64 + 8 * hh;
        const int64_t row = (int64_t)chunk * 256 + t;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(__uint_as_float(o[i]), inv, bv[i]);
        *reinterpret_cast<uint4*>(o16 + row * 64 + (hg * 4 + hh) * 8) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
      };
      // the ring restarts at buffer 0 in every unit; use u of buffer b has parity (u + (b < 2 ? it : 0)) & 1
      auto parity = [&](uint32_t b, uint32_t u) { return (u + (b < 2u ? it : 0u)) & 1u; };
      uint32_t cb = 0, cu = 0;
      uint32_t ra[16], rb[16];
      alive = alive && wait_a(BAR(bS), parity(0u, 0u), kErrAttS);
      tcgen05_fence_after();
      tmem_ld_32x16(lane_addr, ra);
      tmem_ld_32x16(lane_addr + 16, rb);
      PH4(3);
#pragma unroll 1
      for (int hh = 0; hh < 4 && alive; ++hh) {
        float mneg = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const uint32_t nb = cb == 2u ? 0u : cb + 1u, nu = cu + (cb == 2u ? 1u : 0u);
          const uint32_t npar = parity(nb, nu);
          const bool has_next = !(hh == 3 && s == 7);
          const uint32_t col = lane_addr + 32 * cb, ncol = lane_addr + 32 * nb;
          tmem_wait_ld();
          if (s == 1 && hh > 0) take_O_finish(hh - 1);  // its tcgen05.ld was issued in the middle of the previous step
          if (s == 0) {  // the row's reference: max of its first 32 scores
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; i += 2) m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
#pragma unroll
            for (int i = 0; i < 16; i += 2) m = max3(m, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
            mneg = -m * kScale;
          }
          PH4(4);
          exp_store_n<16, 16, kPoly4H2>(ra, kScale, mneg, col);
          if (s == 0 && hh > 0) {  // O of the previous head: its last P.V was issued most of a step ago
            take_O_issue();
            tmem_ld_32x16(lane_addr + 96 + 16 * ((hh - 1) & 1), o);
          }
          bool ready = false;
          if (has_next) {  // S of the next step was issued two steps ago: normally complete by now
            ready = test_all(BAR(bS + nb), npar);
            if (ready) {
              tcgen05_fence_after();
              tmem_ld_32x16(ncol, ra);
            }
          }
          if (s < 7) exp_store_n<16, 16, kPoly4H2>(rb, kScale, mneg, col + 8);
          else exp_store_n<16, S2S_L_DEC - 240, kPoly4H2>(rb, kScale, mneg, col + 8);  // keys 240..249 of the tail step
          PH4(5);
          if (has_next) {
            if (!ready) {
              alive = alive && wait_a(BAR(bS + nb), npar, kErrAttS);
              tcgen05_fence_after();
              tmem_ld_32x16(ncol, ra);
            }
            tmem_ld_32x16(ncol + 16, rb);
          }
          PH4(3);
          tmem_wait_st();
          tcgen05_fence_before();
          warp_arrive_a(BAR(bP + cb));
          if (s == 7) { ob = cb; opar = parity(cb, cu); }
          cb = nb; cu = nu;
          PH4(6);
        }
      }
      if (!alive) break;
      take_O_issue();
      tmem_ld_32x16(lane_addr + 96 + 16, o);
      tmem_wait_ld();
      take_O_finish(3);
      if (__any_sync(0xffffffffu, overflow) && lane == 0) {
        unit_flags[unit] = 1;
        atomicAdd(n_flagged, 1);
      }
      PH4(7);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  PH4_FLUSH;
  if (warp == 0) tmem_dealloc<256>(tmem);
}
