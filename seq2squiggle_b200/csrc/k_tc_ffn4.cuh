// k_tc_fc_ffn4 — decoder fc + LN + FFN + LN with FOUR tiles in flight per SM (included by k_tc.cu inside
// namespace s2s::{anonymous}).
//
// layers.py:82-86, 108-113 for 128-row tiles of the decoder's fp16 residual stream (fp16 operands, fp32 accumulate /
// LayerNorm / residual adds).  The round-1 kernel k_tc_fc_ffn ran two 128-thread CTAs per SM (256 TMEM columns and a private
// 105 KB copy of the weights per tile): two tiles in flight, 0.365 of the tensor peak, and one tile alone on an SM takes
// 8.3 k clk against 5.1 k per tile with two (profiles/r02_experiments_not_kept.txt) — the kernel is bound by how many
// independent tiles cover each other's MMA, TMEM and memory latencies.  Here ONE 512-thread CTA per SM runs four tile
// pipelines over ONE copy of the weights:
//   warpgroup p = warps 4p..4p+3 (thread = row of the pipeline's current tile): residual, LayerNorms, ReLU, output; the
//   pipeline's tcgen05.mma are issued by one elected lane of warp 4p with uniform descriptors (dedicated MMA warps cost
//   the row threads a quarter of their registers: 96 instead of 128, and the spills that followed made it slower).
// The hidden layer is processed in four 64-column quarters so that a pipeline needs 128 TMEM columns:
//   [128p, 128p+64)    ACC: fc accumulator, later D2 (b2 + the four quarters)
//   [128p+64, 128p+128) D1 quarter (fp32) -> H quarter (ReLU, packed fp16 over its first 32 columns, A operand of W2)
// Per tile: fc -> residual + LN1 (fp16 Y over the O tile's buffer) -> 4 x [W1 quarter -> ReLU/pack -> W2 quarter]
// -> residual + LN2 -> TMA store, or the fused output epilogue of the last block.
// Instruction diet (the row threads' issue slots are the second bound after latency):
//   * every bias goes through the tensor core: one extra K = 16 MMA of a constant (1, 1, 0 ...) A tile against a bias B
//     tile (bias as fp16 hi + lo, exact to ~22 bits in the fp32 accumulator) instead of 384 FADD per row;
//   * LayerNorm in one pass (sum and sum of squares while the accumulator is added) and two FFMA per element;
//   * the residual row by four 256-bit loads issued ahead of the fc MMA, its tile prefetched into L2 by TMA a tile earlier.
// Shared memory: W1 32 KB | W2 32 KB | Wfc 8 KB | 4 pipelines x 2 x 16 KB (O / Y / output rows, double-buffered) | ones +
// bias tile 16 KB = 216 KB.  Measured: 1.142 ms (k_tc_fc_ffn) -> 0.81 ms per 32768-chunk layer, 0.52 of the measured bf16
// tensor peak at 60 % of the HBM copy peak (3.2 GB per launch); profiles/r02_ffn4_ncu.txt.
#pragma once

constexpr int kFfn4Threads = 4 * 128;
constexpr int kSmemFfn4 = 4 * kSlab + 8192 + 8 * kSlab + kSlab + 1024;

struct F4Bars {  // per pipeline p: index = base + p (O: base + 2 p + buf)
  enum { W = 0, O = 1, FC = 9, M1 = 13, M2 = 17, D2 = 21, COUNT = 25 };
};

// LayerNorm over a 64-wide row held in registers, from its running sum and sum of squares (one pass: the rows are O(1)
// activations, fp32 cancellation in E[y^2] - mean^2 is ~1e-6 relative); two FFMA per element.
__device__ __forceinline__ void ln_apply(float (&y)[64], float sum, float sq, const float* g, const float* b) {
  const float mean = sum * (1.f / 64.f);
  const float var = fmaxf(fmaf(-mean, mean, sq * (1.f / 64.f)), 0.f);
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  const float shift = -mean * rstd;
#pragma unroll
  for (int i = 0; i < 64; ++i) y[i] = fmaf(fmaf(y[i], rstd, shift), g[i], b[i]);
}

// kRes32 (encoder): the residual stream is fp32 in HBM (x32 read and written, its fp16 copy x16 written by TMA as the next
// block's GEMM operand); otherwise (decoder) the fp16 tensor x16 is the only copy of the stream.
template <bool kOutHead, bool kRes32 = false>
__global__ void __launch_bounds__(kFfn4Threads, 1) k_tc_fc_ffn4(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmWfc,
                                                                const __grid_constant__ CUtensorMap tmW1,
                                                                const __grid_constant__ CUtensorMap tmW2,
                                                                const __grid_constant__ CUtensorMap tmXout,
                                                                const __grid_constant__ FfnParams P,
                                                                const __half* __restrict__ x16, float* __restrict__ x32,
                                                                const __grid_constant__ OutEpi E, int n_tiles, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[F4Bars::COUNT];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW1 = smem;                       // [256 x 128 B]
  uint8_t* sW2 = smem + 2 * kSlab;           // 4 K-slabs x [64 x 128 B]
  uint8_t* sWfc = smem + 4 * kSlab;          // [64 x 128 B]
  uint8_t* sA = smem + 4 * kSlab + 8192;     // [pipeline][buffer] x [128 x 128 B]
  // Bias tile [128 x 128 B], SW128 like the others, read 32 bytes (one K = 16 step) at a time:
  //   K-step 0, 128 rows: (1, 1, 0 ...)                     the A operand of every bias MMA
  //   K-step 1, rows r:   b1[r]        K-step 2: b1[128 + r]  (two 64-row B operands each: the four quarters of b1)
  //   K-step 3, rows 0..63: bfc ; rows 64..127: b2
  // each bias as fp16 (hi, lo) in k = 0, 1, so the product 1*hi + 1*lo carries it to ~22 bits into the fp32 accumulator.
  uint8_t* sB = smem + 4 * kSlab + 8192 + 8 * kSlab;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) s_go = (*status == 0);
  for (int i = tid; i < (int)(kSlab / 16); i += kFfn4Threads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (!s_go) return;
  if (tid < 128) {
    auto hilo = [](float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
    };
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 0)) = 0x3C003C00u;
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 2)) = hilo(P.b1[tid]);
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 4)) = hilo(P.b1[128 + tid]);
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 6)) = hilo(tid < 64 ? P.bfc[tid] : P.b2[tid - 64]);
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc<512>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bars[F4Bars::W], 1);
    for (int p = 0; p < 4; ++p) {
      mbar_init(&bars[F4Bars::O + 2 * p], 1); mbar_init(&bars[F4Bars::O + 2 * p + 1], 1);
      mbar_init(&bars[F4Bars::FC + p], 1); mbar_init(&bars[F4Bars::M1 + p], 1); mbar_init(&bars[F4Bars::M2 + p], 1);
      mbar_init(&bars[F4Bars::D2 + p], 1);
    }
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmWfc); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmXout);
  }
  fence_proxy_async_smem();   // the bias tile (zero fill by every thread) is read by the tensor core's async proxy
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
  if (tmem != 0u) {   // the one CTA of the SM owns all 512 columns: TMEM operands below are immediates
    if (tid == 0) atomicExch(status, kErrFfnLoad);
    sts_u32(abort_a, 1u);
  }
  // pipeline p processes the tiles blockIdx.x + (4 it + p) gridDim.x, it = 0, 1, ...
  auto n_tiles_of = [&](int p) {
    const int first = (int)blockIdx.x + p * (int)gridDim.x, stride = 4 * (int)gridDim.x;
    return first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
  };

  {
    // =============================== warpgroup of pipeline p ============================================
    constexpr bool kTmaStore = !kOutHead;
    const int p = warp >> 2, r = tid & 127;
    const bool issuer = (warp & 3) == 0;     // warp 4p issues the pipeline's MMAs (one elected lane, uniform descriptors)
    const uint32_t lane_addr = tmem_addr(0u, (warp & 3) * 32, 128 * p);
    const uint32_t tACC = 128u * p, tD1 = 128u * p + 64u;
    uint8_t* sAp = sA + p * 2 * kSlab;
    const uint32_t idesc64 = umma_idesc(128, 64, kFmtF16);
    const uint64_t dW1 = umma_desc_k_sw128(smem_u32(sW1)), dW2 = umma_desc_k_sw128(smem_u32(sW2)),
                   dWfc = umma_desc_k_sw128(smem_u32(sWfc)), dAp = umma_desc_k_sw128(smem_u32(sAp)),
                   dOne = umma_desc_k_sw128(smem_u32(sB));
    // bias operand: 64-row half `h` of the bias tile, K-step `ks`
    auto dBias = [&](int h, int ks) { return dOne + (uint64_t)((h * 8192) >> 4) + (uint64_t)(2 * ks); };
    const int n_it = n_tiles_of(p);
    const int stride = 4 * (int)gridDim.x;
    int tile = (int)blockIdx.x + p * (int)gridDim.x;
    auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + p) : "memory"); };
    if (tid == 0) {   // the weights, once per CTA
      mbar_arrive_expect_tx(&bars[F4Bars::W], 4 * kSlab + 8192);
      tma_load_2d(sW1, &tmW1, &bars[F4Bars::W], 0, 0);
#pragma unroll
      for (int s = 0; s < 4; ++s) tma_load_2d(sW2 + s * 8192, &tmW2, &bars[F4Bars::W], s * 64, 0);
      tma_load_2d(sWfc, &tmWfc, &bars[F4Bars::W], 0, 0);
    }
    if (r == 0 && n_it > 0) {
      mbar_arrive_expect_tx(&bars[F4Bars::O + 2 * p], kSlab);
      tma_load_2d(sAp, &tmA, &bars[F4Bars::O + 2 * p], 0, tile * 128);
    }
    if (issuer) wait_a(BAR(F4Bars::W), 0u, kErrFfnLoad);
    for (int it = 0; it < n_it; ++it, tile += stride) {
      if (lds_u32(abort_a)) break;
      const int buf = it & 1;
      const uint32_t ph = (uint32_t)it & 1u;
      uint8_t* sAb = sAp + buf * kSlab;
      const uint32_t sAb_a = smem_u32(sAb);
      const uint64_t dA = dAp + (uint64_t)((buf * kSlab) >> 4);
      const int64_t row = (int64_t)tile * 128 + r;
      if (r == 0 && it + 1 < n_it) {   // the other buffer: its last reader is the TMA store of the previous tile's rows
        if (kTmaStore && it > 0) tma_store_wait_read();
        mbar_arrive_expect_tx(&bars[F4Bars::O + 2 * p + (buf ^ 1)], kSlab);
        tma_load_2d(sAp + (buf ^ 1) * kSlab, &tmA, &bars[F4Bars::O + 2 * p + (buf ^ 1)], 0, (tile + stride) * 128);
        // the next tile's residual rows (x16 = the tensor tmXout describes): into L2 now, so that the row loads a tile
        // later are not DRAM round trips
        if (!kRes32) tma_prefetch_l2_2d(&tmXout, 0, (tile + stride) * 128);
      }
      // first residual (the block input): the 32-byte sectors of the thread's row, in flight while the fc MMA is issued
      // and runs -- four of fp16 or eight of fp32
      uint32_t xr[kRes32 ? 8 : 4][8];
#pragma unroll
      for (int i = 0; i < (kRes32 ? 8 : 4); ++i) {
        if constexpr (kRes32) ldg_256(x32 + row * 64 + 8 * i, xr[i]);
        else ldg_256(x16 + row * 64 + 16 * i, xr[i]);
      }
      if (issuer) {   // attention output projection: ACC = O Wfc^T (the previous tile's D2 was read before its closing sync)
        wait_a(BAR(F4Bars::O + 2 * p + buf), ((uint32_t)it >> 1) & 1u, kErrFfnLoad);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_f16_ss(tACC, dA + (uint64_t)(2 * s), dWfc + (uint64_t)(2 * s), idesc64, s > 0);
          umma_f16_ss(tACC, dOne, dBias(0, 3), idesc64, 1u);   // + bfc
          umma_commit_a(BAR(F4Bars::FC + p));
        }
        __syncwarp();
      }
      float y[64];
      if constexpr (kRes32) {
#pragma unroll
        for (int i = 0; i < 64; ++i) y[i] = __uint_as_float(xr[i >> 3][i & 7]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xr[i][j]));
            y[16 * i + 2 * j] = f.x;
            y[16 * i + 2 * j + 1] = f.y;
          }
        }
      }
      wait_a(BAR(F4Bars::FC + p), ph, kErrFcMma);
      tcgen05_fence_after();
      uint32_t rr[16];
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld_32x16(lane_addr + c0, rr);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          y[c0 + i] += __uint_as_float(rr[i]);
          sum += y[c0 + i];
          sq = fmaf(y[c0 + i], y[c0 + i], sq);
        }
      }
      ln_apply(y, sum, sq, P.g1, P.be1);   // LayerNorm 1 (slf_attn.layer_norm); Y stays in registers
#pragma unroll
      for (int c = 0; c < 8; ++c)   // fp16 copy -> swizzled A tile over the O tile
        sts_u4(sAb_a + sw128_offset(r, c),
            make_uint4(pack_half2(y[8 * c], y[8 * c + 1]), pack_half2(y[8 * c + 2], y[8 * c + 3]),
                       pack_half2(y[8 * c + 4], y[8 * c + 5]), pack_half2(y[8 * c + 6], y[8 * c + 7])));
      fence_proxy_async_smem();
      tcgen05_fence_before();
      wg_sync();   // Y is in shared memory, the fc accumulator has been read
      // the hidden layer, a 64-column quarter at a time: D1 = Y W1_q^T ; relu(D1 + b1) -> fp16, packed over the quarter's
      // first 32 columns ; ACC (+)= H_q W2_q^T
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (issuer) {
          tcgen05_fence_after();
          if (q > 0) {   // W2 of the previous quarter reads H out of the columns this D1 quarter is written to
            wait_a(BAR(F4Bars::M2 + p), (uint32_t)(q - 1) & 1u, kErrFfnMma2);
            tcgen05_fence_after();
          }
          if (elect_one()) {
            const uint64_t b = dW1 + (uint64_t)((q * 8192) >> 4);
#pragma unroll
            for (int s = 0; s < 4; ++s) umma_f16_ss(tD1, dA + (uint64_t)(2 * s), b + (uint64_t)(2 * s), idesc64, s > 0);
            umma_f16_ss(tD1, dOne, dBias(q & 1, 1 + (q >> 1)), idesc64, 1u);   // + b1[64 q ..]
            umma_commit_a(BAR(F4Bars::M1 + p));
          }
          __syncwarp();
        }
        wait_a(BAR(F4Bars::M1 + p), (uint32_t)q & 1u, kErrFfnMma1);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x16(lane_addr + 64 + 16 * c, rr);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_half2_relu(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
          tmem_st_32x8(lane_addr + 64 + 8 * c, pk);   // over columns this thread has already loaded
        }
        tmem_wait_st();
        tcgen05_fence_before();
        wg_sync();   // H_q is in TMEM
        if (issuer) {
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t b = dW2 + (uint64_t)((q * 8192) >> 4);
            if (q == 0) umma_f16_ss(tACC, dOne, dBias(1, 3), idesc64, 0u);   // b2 opens the accumulator
#pragma unroll
            for (int s = 0; s < 4; ++s) umma_f16_ts(tACC, tD1 + 8u * s, b + (uint64_t)(2 * s), idesc64, 1u);
            umma_commit_a(BAR(F4Bars::M2 + p));
            if (q == 3) umma_commit_a(BAR(F4Bars::D2 + p));
          }
          __syncwarp();
        }
      }
      wait_a(BAR(F4Bars::D2 + p), ph, kErrFfnMma2);
      tcgen05_fence_after();
      sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld_32x16(lane_addr + c0, rr);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          y[c0 + i] += __uint_as_float(rr[i]);
          sum += y[c0 + i];
          sq = fmaf(y[c0 + i], y[c0 + i], sq);
        }
      }
      tcgen05_fence_before();
      ln_apply(y, sum, sq, P.g2, P.be2);   // LayerNorm 2 in registers
      if constexpr (kTmaStore) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts_u4(sAb_a + sw128_offset(r, c),
              make_uint4(pack_half2(y[8 * c], y[8 * c + 1]), pack_half2(y[8 * c + 2], y[8 * c + 3]),
                         pack_half2(y[8 * c + 4], y[8 * c + 5]), pack_half2(y[8 * c + 6], y[8 * c + 7])));
        fence_proxy_async_smem();
        if constexpr (kRes32) {   // the fp32 stream: eight full sectors per row
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(x32 + row * 64 + 8 * i),
                         "f"(y[8 * i]), "f"(y[8 * i + 1]), "f"(y[8 * i + 2]), "f"(y[8 * i + 3]), "f"(y[8 * i + 4]),
                         "f"(y[8 * i + 5]), "f"(y[8 * i + 6]), "f"(y[8 * i + 7])
                         : "memory");
        }
      } else {
        out_head_epilogue(y, P, E, row);
      }
      wg_sync();   // D2 has been read everywhere (the next fc may overwrite it); the output rows are staged
      if (kTmaStore && r == 0) {
        tma_store_2d(&tmXout, sAb, 0, tile * 128);
        tma_store_commit();
      }
    }
    if (kTmaStore && r == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
