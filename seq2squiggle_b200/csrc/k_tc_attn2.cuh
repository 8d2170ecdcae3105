// k_tc_attn2 — pipelined decoder attention (included by k_tc.cu inside namespace s2s::{anonymous}).
//
// Same math and operand layouts as k_tc_attn (fused QKV projection, masked K, padded V^T with a ones row, fp16
// P in TMEM, TS-mode P.V), restructured so the MUFU pipe never waits for a tensor-core round trip:
//   * 5 warps: warps 0-3 are the softmax warps (one query row per thread, TMEM lanes 32w..32w+31); warp 4 issues every
//     TMA load and tcgen05.mma and only talks to the others through mbarriers — there is no __syncthreads in the
//     steady state, so warps drift freely and the four schedulers always have an exp-ready warp;
//   * the 250 keys of a (head, query tile) are processed as FOUR 64-key quarters with a flash-style running
//     max / rescale, so a quarter's scores need only 64 TMEM columns: the CTA's 256 columns are a ring of four
//     buffers [S_q | P_q over S_q | O_q at +32].  S(j+2) and P.V(j) run on the tensor pipe while the softmax warps
//     are busy with quarter j+1, i.e. MMA latency (~500 clk) is fully hidden;
//   * a quarter's 64 scores stay in registers between the max and the exp (one tcgen05.ld per score instead of two).
// Micro-iteration j = ((hh*2 + tile)*4 + q) uses ring buffer q.  Barriers (each completes once per (hh,tile)):
//   bar_S[q]  MMA warp  -> softmax : S_q ready          (tcgen05.commit)
//   bar_P[q]  softmax   -> MMA warp: P_q written        (4 warp arrivals)
//   bar_O[q]  MMA warp  -> softmax : O_q = P_q V ready  (tcgen05.commit)
//   bar_F[q]  softmax   -> MMA warp: O_q read, buffer q free (4 warp arrivals)
#pragma once

constexpr int kAttn2Threads = 160;

__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// acc (8 dims + rowsum) <- acc * 2^((m_run - m_new) c) + o * 2^((m_q - m_new) c)
__device__ __forceinline__ void online_combine(float (&acc)[9], float& m_run, const uint32_t (&o)[16], float m_q, float c) {
  const float m_new = fmaxf(m_run, m_q);
  const float a = ex2_approx((m_run - m_new) * c), b = ex2_approx((m_q - m_new) * c);
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = fmaf(acc[i], a, __uint_as_float(o[i]) * b);
  m_run = m_new;
}

__global__ void __launch_bounds__(kAttn2Threads, 2) k_tc_attn2(const __grid_constant__ CUtensorMap tmX,
                                                               const __grid_constant__ CUtensorMap tmWg,
                                                               const float* __restrict__ bias_g, __half* __restrict__ o16,
                                                               int n_units, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_w, bar_qkv, bar_kv, bar_unit, bar_S[4], bar_P[4], bar_O[4], bar_F[4];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[2][96];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sXQ = smem;                 // 2 x [128 x 128 B]: X tiles, then Q (bytes [0,64) of each row)
  uint8_t* sK = smem + 2 * kSlab;      // [256 keys x 128 B]  masked K of this head group (4 quarters of 8 KB)
  uint8_t* sV = smem + 4 * kSlab;      // 4 key quarters x [64 rows (4 heads x 16) x 128 B]
  uint8_t* sW = smem + 6 * kSlab;      // [96 x 128 B] weight block of the CTA's head group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_go = (*status == 0);
  __syncthreads();
  if (!s_go) return;
  if (warp == 0) tmem_alloc<256>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bar_load, 1); mbar_init(&bar_w, 1); mbar_init(&bar_qkv, 1); mbar_init(&bar_kv, 4); mbar_init(&bar_unit, 4);
    for (int q = 0; q < 4; ++q) { mbar_init(&bar_S[q], 1); mbar_init(&bar_O[q], 1); mbar_init(&bar_P[q], 4); mbar_init(&bar_F[q], 4); }
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWg);
  }
  for (int i = tid; i < 192; i += kAttn2Threads) s_bias[i / 96][i % 96] = bias_g[i];
  // V^T padding rows are constant: row 8 of every head = ones (softmax denominator), rows 9..15 = 0
  for (int i = tid; i < 2 * kSlab / 16; i += kAttn2Threads) reinterpret_cast<uint4*>(sV)[i] = make_uint4(0u, 0u, 0u, 0u);
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
  const float kScale = 0.35355339059327373f * 1.4426950408889634f;  // log2(e) / sqrt(d_k)
  PHASE_DECL

  if (warp == 4) {
    // =============================== TMA + MMA issue warp =======================================
    const uint32_t idesc_qkv = umma_idesc(128, 96, kFmtF16), idesc_s = umma_idesc(128, 64, kFmtF16),
                   idesc_o = umma_idesc(128, 16, kFmtF16);
    const uint32_t aXQ = smem_u32(sXQ), aW = smem_u32(sW);
    const uint64_t dXQ = umma_desc_k_sw128(aXQ), dK = umma_desc_k_sw128(smem_u32(sK)), dV = umma_desc_k_sw128(smem_u32(sV));
    const bool elected = lane == 0;
    uint32_t mg0 = 0, it = 0, ph_w = 0;
    int cur_g = -1;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it, mg0 += 8) {
      const int chunk = unit >> 1, g = unit & 1;
      const uint32_t upar = it & 1;
      if (elected) {
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
      wait_bar(&bar_load, upar, status, &s_abort, kErrAttLoad);
      tcgen05_fence_after();
      if (elected) {  // [128 x 96] = X_tile Wg^T, both tiles (accumulators at columns 0 and 128)
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int s = 0; s < 4; ++s)
            umma_f16_ss(tmem + tile * 128, umma_desc_k_sw128(aXQ + tile * kSlab + s * 32), umma_desc_k_sw128(aW + s * 32),
                        idesc_qkv, s > 0);
        umma_commit(&bar_qkv);
      }
      wait_bar(&bar_kv, upar, status, &s_abort, kErrAttS);  // Q / K / V^T operands are in shared memory
      tcgen05_fence_after();
      // Descriptors are base + (byte offset >> 4) with compile-time offsets inside the unrolled quarter loop, so an
      // MMA costs one or two uniform-datapath adds to issue (the issue warp must cycle faster than a softmax quarter).
      if (elected) {
        umma_f16_ss(tmem, dXQ, dK, idesc_s, 0);                       // S(M=0, q=0)
        umma_commit(&bar_S[0]);
        umma_f16_ss(tmem + 64, dXQ, dK + (8192 >> 4), idesc_s, 0);    // S(M=0, q=1)
        umma_commit(&bar_S[1]);
      }
#pragma unroll 1
      for (int M = 0; M < 8; ++M) {
        const int hh = M >> 1, tile = M & 1, Mn = M + 1;
        const uint32_t par = (mg0 + M) & 1;
        const uint64_t dA = dXQ + (uint64_t)((tile * kSlab + (hh >> 1) * 32) >> 4);             // Q slice of (tile, head pair)
        const uint64_t dB = dK + (uint64_t)((hh * 32) >> 4);                                     // masked K slot of head hh
        const uint64_t dAn = dXQ + (uint64_t)(((Mn & 1) * kSlab + ((Mn >> 1) >> 1) * 32) >> 4);  // same for (M+1)
        const uint64_t dBn = dK + (uint64_t)(((Mn >> 1) * 32) >> 4);
        const uint64_t dVh = dV + (uint64_t)((hh * 2048) >> 4);                                  // V^T rows of head hh
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          wait_bar(&bar_P[q], par, status, &s_abort, kErrAttO);
          tcgen05_fence_after();
          PHASE(10);
          if (elected) {  // O_q = P_q V_h over the quarter's 64 keys: 4 K-steps, A operand straight from TMEM
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tmem + 64 * q + 32, tmem + 64 * q + 8 * ks, dVh + (uint64_t)((q * 8192 + ks * 32) >> 4), idesc_o, ks > 0);
            umma_commit(&bar_O[q]);
          }
          PHASE(11);
          // S two quarters ahead goes into ring buffer (q+2)&3, last used two quarters ago: its O must have been read
          const int q2 = (q + 2) & 3;
          if (q < 2) {
            if (M > 0) {
              wait_bar(&bar_F[q2], par ^ 1, status, &s_abort, kErrAttS);
              tcgen05_fence_after();
            }
            PHASE(12);
            if (elected) {
              umma_f16_ss(tmem + 64 * q2, dA, dB + (uint64_t)((q2 * 8192) >> 4), idesc_s, 0);
              umma_commit(&bar_S[q2]);
            }
          } else if (M < 7) {
            wait_bar(&bar_F[q2], par, status, &s_abort, kErrAttS);
            tcgen05_fence_after();
            PHASE(12);
            if (elected) {
              umma_f16_ss(tmem + 64 * q2, dAn, dBn + (uint64_t)((q2 * 8192) >> 4), idesc_s, 0);
              umma_commit(&bar_S[q2]);
            }
          }
          PHASE(13);
        }
      }
      // every softmax warp has drained its last O: shared-memory operands and TMEM may be overwritten
      wait_bar(&bar_unit, upar, status, &s_abort, kErrAttO);
      tcgen05_fence_after();
    }
  } else {
    // =============================== softmax warps ===============================================
    const uint32_t lane_addr = tmem_addr(tmem, warp * 32, 0);
    uint32_t mg = 0, it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int chunk = unit >> 1, g = unit & 1;
      const uint32_t upar = it & 1;
      PHASE_COUNT(15);
      wait_bar(&bar_qkv, upar, status, &s_abort, kErrAttS);
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
      warp_arrive(&bar_kv);
      PHASE(2);

      float acc[9], m_run = -INFINITY, mq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 9; ++i) acc[i] = 0.f;
      auto take_O = [&](int b, uint32_t parity, float m_b) {
        wait_bar(&bar_O[b], parity, status, &s_abort, kErrAttO);
        tcgen05_fence_after();
        uint32_t o[16];
        tmem_ld_32x16(lane_addr + 64 * b + 32, o);
        tmem_wait_ld();
        tcgen05_fence_before();
        warp_arrive(&bar_F[b]);
        online_combine(acc, m_run, o, m_b, kScale);
      };
      auto finalize = [&](int M) {
        const int hh = M >> 1, tile = M & 1;
        const float inv = 1.0f / acc[8];  // sum of the rounded probabilities, rescaled like the numerators
        const int64_t row = (int64_t)chunk * 256 + tile * 128 + tid;
        *reinterpret_cast<uint4*>(o16 + row * 64 + (g * 4 + hh) * 8) =
            make_uint4(pack_half2(acc[0] * inv, acc[1] * inv), pack_half2(acc[2] * inv, acc[3] * inv),
                       pack_half2(acc[4] * inv, acc[5] * inv), pack_half2(acc[6] * inv, acc[7] * inv));
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
        m_run = -INFINITY;
      };
#pragma unroll 1
      for (int M = 0; M < 8; ++M, ++mg) {
        const uint32_t par = mg & 1;
        const float mp2 = mq[2], mp3 = mq[3];  // maxima of the previous (head, tile)'s quarters 2 and 3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          wait_bar(&bar_S[q], par, status, &s_abort, kErrAttS);
          tcgen05_fence_after();
          PHASE(3);
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(lane_addr + 64 * q, ra);
          tmem_ld_32x32(lane_addr + 64 * q + 32, rb);
          tmem_wait_ld();
          float m = chunk_max<32>(ra, -INFINITY);
          m = (q == 3) ? chunk_max<S2S_L_DEC - 224>(rb, m) : chunk_max<32>(rb, m);
          mq[q] = m;
          const float mneg = -m * kScale;
          PHASE(4);
          chunk_exp_store<32>(ra, kScale, mneg, lane_addr + 64 * q);
          if (q == 3) chunk_exp_store<S2S_L_DEC - 224>(rb, kScale, mneg, lane_addr + 64 * q + 16);
          else chunk_exp_store<32>(rb, kScale, mneg, lane_addr + 64 * q + 16);
          tmem_wait_st();
          tcgen05_fence_before();
          warp_arrive(&bar_P[q]);
          PHASE(5);
          // the O of two quarters ago has had a whole quarter of softmax time to finish
          if (q >= 2) {
            take_O(q - 2, par, mq[q - 2]);
          } else if (M > 0) {
            take_O(q + 2, par ^ 1, q == 0 ? mp2 : mp3);
            if (q == 1) finalize(M - 1);
          }
          PHASE(6);
        }
      }
      take_O(2, (mg - 1) & 1, mq[2]);
      take_O(3, (mg - 1) & 1, mq[3]);
      finalize(7);
      tcgen05_fence_before();
      warp_arrive(&bar_unit);
      PHASE(7);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  PHASE_FLUSH;
  if (warp == 0) tmem_dealloc<256>(tmem);
}
