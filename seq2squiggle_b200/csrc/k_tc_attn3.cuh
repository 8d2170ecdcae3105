// k_tc_attn3 — warp-specialised decoder attention (included by k_tc.cu inside namespace s2s::{anonymous}).
//
// layers.py:19-41, 64-88 for one (chunk, group of 4 heads) per "unit", ONE 864-thread CTA per SM, whole TMEM (512 columns).
// The d_k = 8 attention is bound by the softmax exponentials, so everything else is moved off the softmax warps and the
// exponent arrives in TMEM ready to use:
//   warps 0-15        four softmax warpgroups = (query tile 0 / 1 of the chunk) x (half of every 64-key quarter), one query
//                     row per thread: tcgen05.ld 32 scores -> 2^x (MUFU.EX2 or packed-fp16 polynomial) -> fp16 P over S in
//                     TMEM, nothing else; four of them per scheduler cover each other's barrier and TMEM round trips;
//   warps 16-19       producer warpgroup: QKV-projection epilogue of the NEXT unit (accumulator -> fp16 Q / K / V^T in the
//                     UMMA operand layouts, double-buffered in shared memory) while the softmax warps work on this one;
//   warps 20-23       output warpgroup: O accumulator -> normalise by the denominator -> + V bias -> fp16 rows in HBM;
//   warps 24 / 25     MMA issue warps, one per query tile: S quarters (SS, N=64) and P.V steps (TS, N=16) from a
//                     fully unrolled 16-quarter schedule, every descriptor a compile-time offset from a uniform base
//                     (an MMA whose operands are computed at run time costs 116-130 clk of issue, profiles/r01_mma_rate.txt);
//   warp 26           TMA loads of the X tiles / weight block and the QKV-projection MMAs.
// The softmax reference is folded into the S MMA: every head owns a K=16 operand slice,
//   A_i = [ c q_i (8) | -m_i, -30000, 0 x 6 ],  B_j = [ k_j (8) | 1, pad_j, 0 x 6 ],   c = log2(e)/sqrt(d_k),
// so the accumulator holds x_ij = c q_i.k_j - m_i (and about -30000 for the six pad keys 250..255, whose P is exactly 0)
// and P_ij = 2^x_ij with no scaling FFMA, no row-max pass and no tail special case.  m_i = max of the row's own
// (diagonal) score c q_i.k_i and of its scores against 16 keys spread over the chunk (a [128 x 16] MMA per head and tile
// against a compact copy of those keys): softmax is invariant to the reference, P_ii <= 1 <= max P so the denominator is >= 1, and P only misbehaves
// when some score exceeds the reference by more than 16 (11 nats): then the fp16 P overflows, the row's denominator (accumulated by the tensor core from the same rounded P
// through a ones row of V^T) is inf/NaN, and the unit is flagged and recomputed by the exact two-pass kernel k_tc_attn.
// K bias is dropped (adds a per-row constant to the scores) and the V bias is added to the normalised output.
// TMEM columns: query tile g: ring of three 64-column S/P buffers at 208 g + {0, 64, 128}, O accumulator at 208 g + 192
// (16 columns: 8 values, the denominator, 7 unused); QKV accumulator (one 128-row tile, 96 columns) at 416.
// Barriers complete once per use; both sides derive the parity from the same use counts (the ring restarts at slot 0
// every unit; slot 0 is used 6 times per unit, slots 1 and 2 five times).
#pragma once

constexpr int kAttn3Threads = 864;   // 27 warps, 72 registers per thread; no setmaxnreg (the pool of a CTA is what it was launched with)
#ifndef S2S_POLY3_H2
#define S2S_POLY3_H2 6
#endif
constexpr int kPoly3H2 = S2S_POLY3_H2;  // pairs of every 16 computed by the packed-fp16 polynomial instead of MUFU.EX2
constexpr int kA3Unit = 6 * kSlab;      // bytes of one operand buffer: Q (2 tiles) | K (256 keys) | V^T (4 quarters)
constexpr int kA3RefBytes = 16 * 128;   // compact K tile of the 16 reference keys (one per buffer)
constexpr int kSmemAtt3 = 2 * kA3Unit + 96 * 128 + 2 * kA3RefBytes + 1024;
constexpr uint32_t kA3QkvCol = 416;


// 32 exponents (already x = (s - m) log2e/sqrt(d_k)) -> 16 packed fp16 probabilities -> TMEM
template <int kPolyH>
__device__ __forceinline__ void exp32_store(const uint32_t (&r)[32], uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
    if (kPolyH > 0 && (i * kPolyH) % 16 < kPolyH) pk[i] = ex2_poly_h2(x0, x1);
    else pk[i] = pack_half2(ex2_approx(x0), ex2_approx(x1));
  }
  tmem_st_32x16(taddr, pk);
}

// Optional phase timing (-DS2S_PHASE_TIMING=3): clock() deltas of lane 0 of softmax warp 0, producer warp 8 and MMA warp
// 12, summed into g_phase[] (tools/attn3_phase_timing.py).
#if defined(S2S_PHASE_TIMING) && S2S_PHASE_TIMING == 3
#define A3PH_DECL uint32_t a3_acc[8]; _Pragma("unroll") for (int i_ = 0; i_ < 8; ++i_) a3_acc[i_] = 0u; uint32_t a3_t = (uint32_t)clock();
#define A3PH(i) do { const uint32_t n_ = (uint32_t)clock(); a3_acc[i] += n_ - a3_t; a3_t = n_; } while (0)
#define A3PH_FLUSH(base, n) do { if ((threadIdx.x & 31) == 0) { _Pragma("unroll") for (int i_ = 0; i_ < n; ++i_) atomicAdd(&g_phase[base + i_], (unsigned long long)a3_acc[i_]); } } while (0)
#else
#define A3PH_DECL
#define A3PH(i) do {} while (0)
#define A3PH_FLUSH(base, n) do {} while (0)
#endif

struct A3Bars {  // indices into the barrier array
  enum { W = 0, X = 1, QKV = 3, ACC = 4, KV = 5, DONE = 7, S = 9, P = 15, PV = 21, OF = 27, OR = 29, KQ = 31, COUNT = 32 };
};

__global__ void __launch_bounds__(kAttn3Threads, 1) k_tc_attn3(const __grid_constant__ CUtensorMap tmX,
                                                               const __grid_constant__ CUtensorMap tmWg,
                                                               const float* __restrict__ bias_g, __half* __restrict__ o16,
                                                               int n_units, int* __restrict__ unit_flags,
                                                               int* __restrict__ n_flagged, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[A3Bars::COUNT];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[2][96];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW = smem + 2 * kA3Unit;   // [96 x 128 B] weight block (Wq | Wk | Wv rows) of the CTA's head group
  uint8_t* sRef = sW + 96 * 128;      // 2 x [16 x 128 B]: K rows of the reference keys (second chunks stay zero)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform role index
  if (tid == 0) s_go = (status[0] == 0 && status[1] == 0);
  __syncthreads();
  if (!s_go) return;
  if (warp == 0) tmem_alloc<512>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bars[A3Bars::W], 1);
    mbar_init(&bars[A3Bars::QKV], 1);
    mbar_init(&bars[A3Bars::ACC], 4);
    mbar_init(&bars[A3Bars::KQ], 4);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[A3Bars::X + b], 1);
      mbar_init(&bars[A3Bars::KV + b], 4);
      mbar_init(&bars[A3Bars::DONE + b], 4);
      mbar_init(&bars[A3Bars::OF + b], 4);
      mbar_init(&bars[A3Bars::OR + b], 1);
    }
    for (int i = 0; i < 6; ++i) {
      mbar_init(&bars[A3Bars::S + i], 1);
      mbar_init(&bars[A3Bars::P + i], 8);
      mbar_init(&bars[A3Bars::PV + i], 1);
    }
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWg);
  }
  for (int i = tid; i < 192; i += kAttn3Threads) s_bias[i / 96][i % 96] = bias_g[i];
  // Static parts of the operand buffers (both): zero everything, then
  //   Q rows : nothing static (the X tile lands in the Q buffer and the producer rewrites every chunk of it)
  //   K rows : second chunk = (1, pad_j, 0 x 6)
  //   V^T    : row 8 of every head = ones for the 250 real keys (softmax denominator), rows 9..15 = 0
  for (int i = tid; i < 2 * kA3Unit / 16; i += kAttn3Threads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 2 * kA3RefBytes / 16; i += kAttn3Threads) reinterpret_cast<uint4*>(sRef)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < 2 * 256 * 4; i += kAttn3Threads) {  // (buffer, row, head)
    const int b = i >> 10, row = (i >> 2) & 255, hh = i & 3;
    uint8_t* base = smem + b * kA3Unit;
    // fp16 1.0 = 0x3C00
    *reinterpret_cast<uint4*>(base + 2 * kSlab + sw128_offset(row, 2 * hh + 1)) =
        make_uint4(row < S2S_L_DEC ? 0x00003C00u : 0x3C003C00u, 0u, 0u, 0u);
  }
  for (int i = tid; i < 2 * 4 * 4 * 8; i += kAttn3Threads) {  // (buffer, quarter, head, 16-byte chunk of 8 keys)
    const int b = i >> 7, slab = (i >> 5) & 3, hh = (i >> 3) & 3, ck = i & 7;
    const uint32_t one2 = 0x3C003C00u;
    uint4 v = make_uint4(one2, one2, one2, one2);
    if (slab == 3 && ck == 7) v = make_uint4(one2, 0u, 0u, 0u);   // keys 248, 249 real; 250..255 pad
    *reinterpret_cast<uint4*>(smem + b * kA3Unit + 4 * kSlab + slab * 8192 + sw128_offset(hh * 16 + 8, ck)) = v;
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t bar0 = smem_u32(&bars[0]), abort_a = smem_u32(&s_abort);
  auto BAR = [&](uint32_t idx) { return bar0 + 8u * idx; };
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
  // All 512 columns are allocated by the one CTA of this SM, so the base is column 0 of lane 0; the MMA warps rely on
  // that to keep every TMEM operand an immediate.
  if (tmem != 0u) {
    if (tid == 0) atomicExch(status, kErrAttTmem);
    sts_u32(abort_a, 1u);
  }
  const int n_it = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // units of this CTA
  const int g = blockIdx.x & 1;   // even grid stride: a CTA keeps its head group

  if (warp < 16) {
    // =============================== softmax warpgroups ===========================================
    // Four warpgroups: (query tile, half of every 64-key quarter).  Both halves of a tile work on the same ring slot: half
    // h loads S columns [32 h, 32 h + 32) and writes its 16 packed P columns at [32 h, 32 h + 16) of the slot (inside its
    // own S columns, so it cannot overwrite scores the other half has not loaded yet).  Four softmax warps per scheduler
    // cover each other's barrier / TMEM round trips.
    const int wg = warp >> 3, half = (warp >> 2) & 1;
    const uint32_t lane_addr = tmem_addr(0u, (warp & 3) * 32, 208 * wg + 32 * half);
    const uint32_t barS = BAR(A3Bars::S + 3 * wg), barP = BAR(A3Bars::P + 3 * wg);
    uint32_t bits = 0;                       // parity bits of the ring slots (bit s = parity of the slot's next use)
    A3PH_DECL
    for (int it = 0; it < n_it; ++it) {
      if (lds_u32(abort_a)) break;
      A3PH(5);
      // 16 quarters over the ring of three slots, unrolled by three so that slot addresses and barrier offsets are
      // immediates: five rounds of three and the last quarter (slot 0) on its own.  The four softmax warps of a scheduler are
      // issue-bound while they are in their exponentials, so every instruction around them counts: an early probe of the next
      // quarter's barrier (to hide its ~100 clk round trip) cost more in bookkeeping than it hid (3.451 vs 3.425 ms).
#define A3_QUARTER(cb)                                                                                   \
      {                                                                                                    \
        const uint32_t cpar = (bits >> (cb)) & 1u;                                                         \
        bits ^= 1u << (cb);                                                                                \
        wait_a(barS + 8u * (cb), cpar, kErrAttS);                                                          \
        A3PH(0);                                                                                           \
        tcgen05_fence_after();                                                                             \
        uint32_t ra[32];                                                                                   \
        tmem_ld_32x32(lane_addr + 64 * (cb), ra);                                                          \
        tmem_wait_ld();                                                                                    \
        A3PH(1);                                                                                           \
        exp32_store<kPoly3H2>(ra, lane_addr + 64 * (cb));                                                  \
        A3PH(2);                                                                                           \
        tmem_wait_st();                                                                                    \
        tcgen05_fence_before();                                                                            \
        warp_arrive_a(barP + 8u * (cb));                                                                   \
        A3PH(4);                                                                                           \
      }
#pragma unroll 1
      for (int j0 = 0; j0 < 15; j0 += 3) {
        A3_QUARTER(0)
        A3_QUARTER(1)
        A3_QUARTER(2)
      }
      A3_QUARTER(0)
#undef A3_QUARTER
    }
    if (warp == 0) A3PH_FLUSH(0, 6);
  } else if (warp < 20) {
    // =============================== producer warpgroup ===========================================
    const int lq = warp & 3, r = lq * 32 + lane;
    const uint32_t lane_addr = tmem_addr(0u, lq * 32, kA3QkvCol);
    const float kScale = 0.35355339059327373f * 1.4426950408889634f;   // log2(e) / sqrt(d_k)
    const float* bq = s_bias[g];
    A3PH_DECL
    for (int it = 0; it < n_it; ++it) {
      if (lds_u32(abort_a)) break;
      uint8_t* ub = smem + (it & 1) * kA3Unit;
      uint8_t* krow = ub + 2 * kSlab;
      uint8_t* kref = sRef + (it & 1) * kA3RefBytes;
      float md[2][4];   // the rows' own (diagonal) scores, fp32
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t n = 4u * (uint32_t)it + (uint32_t)tile;
        A3PH(1);
        wait_a(BAR(A3Bars::QKV), n & 1u, kErrAttS);
        A3PH(0);
        tcgen05_fence_after();
        uint32_t rq[32], rk[32];
        tmem_ld_32x32(lane_addr, rq);
        tmem_ld_32x32(lane_addr + 32, rk);
        tmem_wait_ld();
        const int t = tile * 128 + r;      // key index inside the chunk
        const bool is_ref = (t & 15) == 8;  // keys 8, 24, .., 248: the reference keys (one per k-mer position, roughly)
        uint8_t* qrow = ub + tile * kSlab;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          uint32_t pq[4], pk[4];
          float m = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float q0 = (__uint_as_float(rq[8 * hh + 2 * i]) + bq[8 * hh + 2 * i]) * kScale;
            const float q1 = (__uint_as_float(rq[8 * hh + 2 * i + 1]) + bq[8 * hh + 2 * i + 1]) * kScale;
            const float k0 = __uint_as_float(rk[8 * hh + 2 * i]), k1 = __uint_as_float(rk[8 * hh + 2 * i + 1]);
            pq[i] = pack_half2(q0, q1);
            pk[i] = pack_half2(k0, k1);
            m = fmaf(q0, k0, fmaf(q1, k1, m));
          }
          md[tile][hh] = m;
          *reinterpret_cast<uint4*>(qrow + sw128_offset(r, 2 * hh)) = make_uint4(pq[0], pq[1], pq[2], pq[3]);
          *reinterpret_cast<uint4*>(krow + sw128_offset(t, 2 * hh)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (is_ref) *reinterpret_cast<uint4*>(kref + sw128_offset(t >> 4, 2 * hh)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        uint32_t rv[32];   // V after Q / K: 64 + 32 live accumulator registers instead of 96
        tmem_ld_32x32(lane_addr + 64, rv);
        tmem_wait_ld();
        tcgen05_fence_before();
        warp_arrive_a(BAR(A3Bars::ACC));   // the accumulator may be overwritten (next tile's projection / reference scores)
        uint8_t* vslab = ub + 4 * kSlab + (t >> 6) * 8192 + (t & 7) * 2;
        const uint32_t ck = (t & 63) >> 3;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
#pragma unroll
          for (int d = 0; d < 8; ++d)
            *reinterpret_cast<__half*>(vslab + sw128_offset(hh * 16 + d, ck)) = __float2half_rn(__uint_as_float(rv[8 * hh + d]));
      }
      // The softmax reference of a row: max of its own score and of its scores against the 16 reference keys, the latter
      // by the tensor core ([128 x 16] per (tile, head), issued by the load warp into the projection accumulator's
      // columns once Q and the reference keys are in shared memory).  Any reference is exact for the softmax; it only
      // has to be close enough to the row maximum that no probability overflows fp16 (2^16).  The diagonal alone is a poor
      // guess: with W_q, W_k scaled x3 every unit had a row whose maximum beat its own score by more than 11 nats.
      fence_proxy_async_smem();
      warp_arrive_a(BAR(A3Bars::KQ));
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t n = 4u * (uint32_t)it + 2u + (uint32_t)tile;
        wait_a(BAR(A3Bars::QKV), n & 1u, kErrAttS);
        tcgen05_fence_after();
        uint32_t rs[32], rt[32];
        tmem_ld_32x32(lane_addr, rs);        // heads 0, 1: 16 reference scores each
        tmem_ld_32x32(lane_addr + 32, rt);   // heads 2, 3
        tmem_wait_ld();
        tcgen05_fence_before();
        warp_arrive_a(BAR(A3Bars::ACC));
        uint8_t* qrow = ub + tile * kSlab;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          float m = md[tile][hh];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const uint32_t a = hh < 2 ? rs[16 * hh + j] : rt[16 * (hh - 2) + j];
            const uint32_t b2 = hh < 2 ? rs[16 * hh + j + 1] : rt[16 * (hh - 2) + j + 1];
            m = max3(m, __uint_as_float(a), __uint_as_float(b2));
          }
          // (-m_i, -30000): subtracted by the S MMA through the ones column of K; pad keys get -30000 on top
          *reinterpret_cast<uint4*>(qrow + sw128_offset(r, 2 * hh + 1)) = make_uint4(pack_half2(-m, -30000.0f), 0u, 0u, 0u);
        }
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      warp_arrive_a(BAR(A3Bars::KV + (it & 1)));
    }
    A3PH(1);
    if (warp == 16) A3PH_FLUSH(14, 2);
  } else if (warp < 24) {
    // =============================== output warpgroup =============================================
    // O of every (head, query tile): wait for the head's last P.V, read the accumulator, release it to the MMA warp,
    // normalise by the denominator (column 8), add the V bias, store 16 bytes per row.  Kept off the softmax warps: they
    // run in lockstep per quarter, so anything one of them does besides exponentials is paid by the whole tile.
    const int lq = warp & 3, r = lq * 32 + lane;
    A3PH_DECL
    for (int it = 0; it < n_it; ++it) {
      if (lds_u32(abort_a)) break;
      const int unit = blockIdx.x + it * gridDim.x;
      const int chunk = unit >> 1;
      bool overflow = false;
#pragma unroll 1
      for (int hh = 0; hh < 4; ++hh) {
        const float* bv = s_bias[g] + 64 + 8 * hh;
#pragma unroll 1
        for (int wg = 0; wg < 2; ++wg) {
          A3PH(1);
          wait_a(BAR(A3Bars::OR + wg), (uint32_t)hh & 1u, kErrAttO);   // 4 completions per unit: parity = hh & 1
          A3PH(0);
          tcgen05_fence_after();
          uint32_t o[16];
          tmem_ld_32x16(tmem_addr(0u, lq * 32, 208 * wg + 192), o);
          tmem_wait_ld();
          tcgen05_fence_before();
          warp_arrive_a(BAR(A3Bars::OF + wg));
          const float den = __uint_as_float(o[8]);   // sum of the rounded probabilities; P_ii = 1, so den >= 1
          overflow |= !(den < 1e30f) || !(den > 0.25f);
          const float inv = __frcp_rn(den);
          const int64_t row = (int64_t)chunk * 256 + wg * 128 + r;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaf(__uint_as_float(o[i]), inv, bv[i]);
          *reinterpret_cast<uint4*>(o16 + row * 64 + (g * 4 + hh) * 8) =
              make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
      }
      if (__any_sync(0xffffffffu, overflow) && lane == 0) {
        unit_flags[unit] = 1;
        atomicAdd(n_flagged, 1);
      }
      // every accumulator of the unit has been read: all its MMAs are complete, the operand buffer may be refilled
      warp_arrive_a(BAR(A3Bars::DONE + (it & 1)));
    }
    A3PH(1);
    if (warp == 20) A3PH_FLUSH(12, 2);
  } else {
    if (warp == 24 || warp == 25) {
      // =============================== MMA issue warps ============================================
      const uint32_t wg = (uint32_t)(warp - 24);
      const uint32_t idesc_s = umma_idesc(128, 64, kFmtF16), idesc_o = umma_idesc(128, 16, kFmtF16);
      const uint32_t tS = 208u * wg, tO = 208u * wg + 192u;
      const uint64_t d0 = umma_desc_k_sw128(smem_u32(smem));
      A3PH_DECL
      for (int it = 0; it < n_it; ++it) {
        if (lds_u32(abort_a)) break;
        const uint32_t buf = (uint32_t)it & 1u;
        const uint32_t p12 = (uint32_t)it & 1u;   // parity base of ring slots 1 and 2 (5 uses per unit); slot 0: 6 uses -> 0
        const uint64_t dQ = d0 + (uint64_t)((buf * kA3Unit + wg * kSlab) >> 4);
        const uint64_t dK = d0 + (uint64_t)((buf * kA3Unit + 2 * kSlab) >> 4);
        const uint64_t dV = d0 + (uint64_t)((buf * kA3Unit + 4 * kSlab) >> 4);
        A3PH(4);
        wait_a(BAR(A3Bars::KV + buf), ((uint32_t)it >> 1) & 1u, kErrAttS);
        A3PH(5);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {   // all three ring slots are free at the start of a unit
            umma_f16_ss(tS + 64u * j, dQ, dK + (uint64_t)((j * 8192) >> 4), idesc_s, 0u);
            umma_commit_a(BAR(A3Bars::S + 3 * wg + j));
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int hh = j >> 2, q = j & 3, slot = j % 3, use = j / 3;
          const uint32_t par = ((slot == 0 ? 0u : p12) + (uint32_t)use) & 1u;
          A3PH(3);
          if (q == 0 && (j > 0 || it > 0)) {   // the O accumulator still holds the previous head until it has been read
            wait_a(BAR(A3Bars::OF + wg), (uint32_t)(hh + 1) & 1u, kErrAttO);
          }
          A3PH(4);
          wait_a(BAR(A3Bars::P + 3 * wg + slot), par, kErrAttO);
          A3PH(0);
          tcgen05_fence_after();
          if (elect_one()) {   // O += P_q V_h over the quarter's 64 keys: 4 K-steps, A operand straight from TMEM
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tO, tS + 64u * slot + 32u * (ks >> 1) + 8u * (ks & 1), dV + (uint64_t)((q * 8192 + hh * 2048 + ks * 32) >> 4), idesc_o,
                          (q > 0 || ks > 0) ? 1u : 0u);
            umma_commit_a(BAR(A3Bars::PV + 3 * wg + slot));
            if (q == 3) umma_commit_a(BAR(A3Bars::OR + wg));   // the head's O is complete: output warpgroup
          }
          __syncwarp();
          A3PH(1);
          if (j + 3 < 16) {   // the slot is free again once its P.V has completed: S three quarters ahead goes in
            wait_a(BAR(A3Bars::PV + 3 * wg + slot), par, kErrAttS);
            A3PH(2);
            tcgen05_fence_after();
            if (elect_one()) {
              const int j3 = j + 3, h3 = j3 >> 2, q3 = j3 & 3;
              umma_f16_ss(tS + 64u * slot, dQ + (uint64_t)((h3 * 32) >> 4), dK + (uint64_t)((q3 * 8192 + h3 * 32) >> 4), idesc_s, 0u);
              umma_commit_a(BAR(A3Bars::S + 3 * wg + slot));
            }
            __syncwarp();
          } else {
            // last use of the slot in this unit: the next unit's first S quarters overwrite the ring, so its P.V must
            // have completed (the softmax warps only wait for the very last one)
            wait_a(BAR(A3Bars::PV + 3 * wg + slot), par, kErrAttS);
            A3PH(2);
          }
        }
      }
      if (warp == 24) A3PH_FLUSH(6, 6);
    } else if (warp == 26) {
      // =============================== TMA + QKV projection warp ==================================
      const uint32_t idesc_qkv = umma_idesc(128, 96, kFmtF16);
      const uint64_t d0 = umma_desc_k_sw128(smem_u32(smem));
      const uint64_t dW = umma_desc_k_sw128(smem_u32(sW));
      const uint64_t dRef = umma_desc_k_sw128(smem_u32(sRef));
      const uint32_t idesc_ref = umma_idesc(128, 16, kFmtF16);
      if (n_it > 0 && elect_one()) {
        mbar_arrive_expect_tx(&bars[A3Bars::W], 96 * 128);
        tma_load_2d(sW, &tmWg, &bars[A3Bars::W], 0, g * 96);
      }
      __syncwarp();
      for (int it = 0; it < n_it; ++it) {
        if (lds_u32(abort_a)) break;
        const uint32_t buf = (uint32_t)it & 1u;
        const int chunk = (blockIdx.x + it * gridDim.x) >> 1;
        uint8_t* ub = smem + buf * kA3Unit;
        // the unit that used this buffer two iterations ago has been finished by both softmax warpgroups
        if (it >= 2) wait_a(BAR(A3Bars::DONE + buf), (((uint32_t)it >> 1) - 1u) & 1u, kErrAttLoad);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[A3Bars::X + buf], 2 * kSlab);
          tma_load_2d(ub, &tmX, &bars[A3Bars::X + buf], 0, chunk * 256);
          tma_load_2d(ub + kSlab, &tmX, &bars[A3Bars::X + buf], 0, chunk * 256 + 128);
        }
        __syncwarp();
        if (it == 0) wait_a(BAR(A3Bars::W), 0u, kErrAttLoad);
        wait_a(BAR(A3Bars::X + buf), ((uint32_t)it >> 1) & 1u, kErrAttLoad);
        tcgen05_fence_after();
        const uint64_t dX = d0 + (uint64_t)((buf * kA3Unit) >> 4);
        // Four accumulator uses per unit, handed over through the same two barriers (QKV: written, ACC: read):
        // projection of tile 0, of tile 1, reference scores of tile 0, of tile 1.
#pragma unroll
        for (int step = 0; step < 4; ++step) {
          const uint32_t n = 4u * (uint32_t)it + (uint32_t)step;
          if (n >= 1u) {   // the producer has read the previous contents of the accumulator columns
            wait_a(BAR(A3Bars::ACC), (n - 1u) & 1u, kErrAttLoad);
            tcgen05_fence_after();
          }
          if (step == 2) {   // Q (without the reference) and the reference keys of the unit are in shared memory
            wait_a(BAR(A3Bars::KQ), (uint32_t)it & 1u, kErrAttLoad);
            tcgen05_fence_after();
          }
          if (elect_one()) {
            if (step < 2) {   // [128 x 96] = X_tile Wg^T
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_f16_ss(kA3QkvCol, dX + (uint64_t)((step * kSlab + s * 32) >> 4), dW + (uint64_t)((s * 32) >> 4), idesc_qkv, s > 0);
            } else {          // [128 x 16] = c Q_h Kref_h^T per head: the reference scores of tile (step - 2)
#pragma unroll
              for (int hh = 0; hh < 4; ++hh)
                umma_f16_ss(kA3QkvCol + 16u * hh, dX + (uint64_t)(((step - 2) * kSlab + hh * 32) >> 4),
                            dRef + (uint64_t)((buf * kA3RefBytes + hh * 32) >> 4), idesc_ref, 0u);
            }
            umma_commit_a(BAR(A3Bars::QKV));
          }
          __syncwarp();
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
