// k_tc_fc_ffn3 — decoder fc + LN + FFN + LN, THREE tile pipelines per SM whose MMA round trips are off the critical path
// (included by k_tc.cu inside namespace s2s::{anonymous}; same arithmetic as k_tc_fc_ffn4, layers.py:82-86, 108-113).
//
// k_tc_fc_ffn4 (four pipelines, 128 TMEM columns each) leaves every pipeline waiting ~850 clk per hidden-layer quarter for
// "W2 of this quarter, then W1 of the next" (profiles/r02_ffn4_ncu.txt: 21 % of the warp time at the M1 waits, 16 % at
// the residual row load).  With 160 columns per pipeline the hidden layer gets TWO D1 buffers, so the W1 MMAs of piece
// k+1 run while the threads are still packing piece k, and with 384 threads (168 registers each) the next tile's residual
// row is prefetched into registers a tile ahead:
//   [160p, +64)      ACC: fc accumulator, later D2 (b2 + the five W2 pieces)
//   [160p+64, +64)   D1 buffer a: hidden pieces 0, 2, 4 (64 units: [0,64) [96,160) [192,256))
//   [160p+128, +32)  D1 buffer b: hidden pieces 1, 3    (32 units: [64,96) [160,192))
// Each D1 piece (fp32) is overwritten in place by its ReLU'd fp16 copy H (packed, first half of the piece's columns), the A
// operand of that piece's W2 MMAs.  After the warpgroup's sync on "H_k stored", warp 4p issues W2(k) and then, WITHOUT
// waiting, W1(k+2) into the same buffer: tcgen05.mma of one thread execute in issue order, so W1(k+2) cannot overwrite H_k
// before W2(k) has read it (checked bit for bit against the waiting variant on 6 x 16 launches, tools notes in
// profiles/r02_experiments_not_kept.txt; CUTLASS's Blackwell attention relies on the same ordering for S/P).  The threads
// then wait for W1(k+1), which was issued a whole piece earlier.  Every tcgen05.commit covers all MMAs issued before it, so
// "W1(k+2) complete" also means "W2(k) complete" before H_{k+2} is stored over the buffer.
// Shared memory: W1 32 KB | W2 32 KB | Wfc 8 KB | 3 pipelines x 2 x 16 KB (O / Y / output rows) | ones + bias tile 16 KB.
#pragma once

constexpr int kFfn3Threads = 3 * 128;
constexpr int kSmemFfn3 = 4 * kSlab + 8192 + 6 * kSlab + kSlab + 1024;

struct F3Bars {  // per pipeline p
  enum { W = 0, O = 1 /* + 2p + buf */, FC = 7 /* + p */, M1 = 10 /* + 2p + D1 buffer */, D2 = 16 /* + p */,
         EV = 19 /* + 2p + (point & 1): the warpgroup's four warps have reached an issue point */, COUNT = 25 };
};

// hidden piece k = 0..4: first unit, width, bias rows / K-step in the bias tile
__host__ __device__ constexpr int f3_h0(int k) { return 96 * (k >> 1) + 64 * (k & 1); }
__host__ __device__ constexpr int f3_n(int k) { return (k & 1) ? 32 : 64; }
__host__ __device__ constexpr int f3_bias_row(int k) { return k == 0 || k == 2 ? 0 : (k == 3 ? 96 : 64); }
__host__ __device__ constexpr int f3_bias_ks(int k) { return k == 2 || k == 4 ? 2 : 1; }

template <bool kOutHead>
__global__ void __launch_bounds__(kFfn3Threads, 1) k_tc_fc_ffn3(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmWfc,
                                                                const __grid_constant__ CUtensorMap tmW1,
                                                                const __grid_constant__ CUtensorMap tmW2,
                                                                const __grid_constant__ CUtensorMap tmXout,
                                                                const __grid_constant__ FfnParams P,
                                                                const __half* __restrict__ x16,
                                                                const __grid_constant__ OutEpi E, int n_tiles, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[F3Bars::COUNT];
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW1 = smem;                       // [256 x 128 B]
  uint8_t* sW2 = smem + 2 * kSlab;           // 4 K-slabs x [64 x 128 B]
  uint8_t* sWfc = smem + 4 * kSlab;          // [64 x 128 B]
  uint8_t* sA = smem + 4 * kSlab + 8192;     // [pipeline][buffer] x [128 x 128 B]
  // Bias tile [128 x 128 B], SW128 like the others, read 32 bytes (one K = 16 step) at a time:
  //   K-step 0, 128 rows: (1, 1, 0 ...)            the A operand of every bias MMA
  //   K-step 1: rows 0..63 b1 of piece 0, 64..95 piece 1, 96..127 piece 3 ; K-step 2: rows 0..63 piece 2, 64..127 piece 4
  //   K-step 3: rows 0..63 bfc ; rows 64..127 b2
  // each bias as fp16 (hi, lo) in k = 0, 1, so the product 1*hi + 1*lo carries it to ~22 bits into the fp32 accumulator.
  uint8_t* sB = smem + 4 * kSlab + 8192 + 6 * kSlab;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) s_go = (*status == 0);
  for (int i = tid; i < (int)(kSlab / 16); i += kFfn3Threads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (!s_go) return;
  if (tid < 128) {
    auto hilo = [](float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
    };
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 0)) = 0x3C003C00u;
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 2)) = hilo(P.b1[tid < 96 ? tid : tid + 64]);
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 4)) = hilo(P.b1[tid < 64 ? 96 + tid : 128 + tid]);
    *reinterpret_cast<uint32_t*>(sB + sw128_offset(tid, 6)) = hilo(tid < 64 ? P.bfc[tid] : P.b2[tid - 64]);
  }
  if (warp == 0) tmem_alloc<512>(&s_tmem);
  if (tid == 0) {
    for (int i = 0; i < F3Bars::COUNT; ++i) mbar_init(&bars[i], i >= F3Bars::EV ? 4 : 1);
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmWfc); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmXout);
  }
  fence_proxy_async_smem();   // the bias tile is read by the tensor core's async proxy
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
  {
    constexpr bool kTmaStore = !kOutHead;
    const int p = warp >> 2, r = tid & 127;
    // Issue points (six per tile: "Y stored", "H_k stored" k = 0..4): every warp arrives on the pipeline's EV barrier and goes
    // on; ONE warp (rotating, so that the time spent issuing is spread evenly) waits for all four and issues the MMAs that
    // the point releases, by one elected lane with uniform descriptors.  No warpgroup-wide bar.sync inside a tile.
    const int wq = warp & 3;
    const uint32_t tACC = 160u * p;
    const uint32_t lane_addr = tmem_addr(0u, (warp & 3) * 32, tACC);
    uint8_t* sAp = sA + p * 2 * kSlab;
    const uint32_t idesc64 = umma_idesc(128, 64, kFmtF16), idesc32 = umma_idesc(128, 32, kFmtF16);
    const uint64_t dW1 = umma_desc_k_sw128(smem_u32(sW1)), dW2 = umma_desc_k_sw128(smem_u32(sW2)),
                   dWfc = umma_desc_k_sw128(smem_u32(sWfc)), dAp = umma_desc_k_sw128(smem_u32(sAp)),
                   dOne = umma_desc_k_sw128(smem_u32(sB));
    // bias operand: rows row0.. of the bias tile, K-step ks
    auto dBias = [&](int row0, int ks) { return dOne + (uint64_t)((row0 * 128) >> 4) + (uint64_t)(2 * ks); };
    // pipeline p processes the tiles blockIdx.x + (3 it + p) gridDim.x, it = 0, 1, ...
    const int stride = 3 * (int)gridDim.x;
    int tile = (int)blockIdx.x + p * (int)gridDim.x;
    const int n_it = tile < n_tiles ? (n_tiles - tile + stride - 1) / stride : 0;
    auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + p) : "memory"); };
    if (tid == 0) {   // the weights, once per CTA
      mbar_arrive_expect_tx(&bars[F3Bars::W], 4 * kSlab + 8192);
      tma_load_2d(sW1, &tmW1, &bars[F3Bars::W], 0, 0);
#pragma unroll
      for (int s = 0; s < 4; ++s) tma_load_2d(sW2 + s * 8192, &tmW2, &bars[F3Bars::W], s * 64, 0);
      tma_load_2d(sWfc, &tmWfc, &bars[F3Bars::W], 0, 0);
    }
    if (r == 0 && n_it > 0) {
      mbar_arrive_expect_tx(&bars[F3Bars::O + 2 * p], kSlab);
      tma_load_2d(sAp, &tmA, &bars[F3Bars::O + 2 * p], 0, tile * 128);
    }
    // W1 of hidden piece k -> its D1 buffer (+ b1), one commit
    auto issue_w1 = [&](int k, uint64_t dA) {
      const uint32_t d = tACC + 64u + 64u * (uint32_t)(k & 1);
      const uint32_t idesc = (k & 1) ? idesc32 : idesc64;
      const uint64_t b = dW1 + (uint64_t)((f3_h0(k) * 128) >> 4);
#pragma unroll
      for (int s = 0; s < 4; ++s) umma_f16_ss(d, dA + (uint64_t)(2 * s), b + (uint64_t)(2 * s), idesc, s > 0);
      umma_f16_ss(d, dOne, dBias(f3_bias_row(k), f3_bias_ks(k)), idesc, 1u);
      umma_commit_a(BAR(F3Bars::M1 + 2 * p + (k & 1)));
    };
    // ACC += H_k W2[:, piece k]^T, H_k packed fp16 at the start of the piece's D1 buffer
    auto issue_w2 = [&](int k) {
      const uint32_t a = tACC + 64u + 64u * (uint32_t)(k & 1);
#pragma unroll
      for (int j = 0; j < f3_n(k) / 16; ++j) {
        const int h = f3_h0(k) + 16 * j;
        umma_f16_ts(tACC, a + 8u * j, dW2 + (uint64_t)(((h >> 6) * 8192) >> 4) + (uint64_t)(2 * ((h & 63) >> 4)), idesc64, 1u);
      }
    };
    uint32_t xr[4][8];   // residual row (the block input, fp16) of the NEXT tile to start: four 32-byte sectors
    if (n_it > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) ldg_256(x16 + ((int64_t)tile * 128 + r) * 64 + 16 * i, xr[i]);
    }
    uint32_t n_a = 0, n_b = 0;   // completed waits on the two D1 barriers (phase parity)
    // attention output projection of the tile in buffer `b`: D1a = O Wfc^T + bfc (its consumer reads it before W1(0) is released)
    auto issue_fc = [&](int b, uint32_t o_parity) {
      wait_a(BAR(F3Bars::O + 2 * p + b), o_parity, kErrFfnLoad);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t dAo = dAp + (uint64_t)((b * kSlab) >> 4);
#pragma unroll
        for (int s = 0; s < 4; ++s) umma_f16_ss(tACC + 64u, dAo + (uint64_t)(2 * s), dWfc + (uint64_t)(2 * s), idesc64, s > 0);
        umma_f16_ss(tACC + 64u, dOne, dBias(0, 3), idesc64, 1u);
        umma_commit_a(BAR(F3Bars::FC + p));
      }
      __syncwarp();
    };
    if (wq == 0) {
      wait_a(BAR(F3Bars::W), 0u, kErrFfnLoad);
      if (n_it > 0) issue_fc(0, 0u);
    }
    for (int it = 0; it < n_it; ++it, tile += stride) {
      if (lds_u32(abort_a)) break;
      const int buf = it & 1;
      const uint32_t ph = (uint32_t)it & 1u;
      uint8_t* sAb = sAp + buf * kSlab;
      const uint64_t dA = dAp + (uint64_t)((buf * kSlab) >> 4);
      const int64_t row = (int64_t)tile * 128 + r;
      const int rot = (2 * it) & 3;   // six issue points per tile
      // issue point e of this tile: arrive; the point's warp waits for the four arrivals and runs `fn`.  Two barriers, for
      // even and odd points: a warp can be one point ahead of the slowest warp of its group (what it needs for point e + 1
      // was released at point e - 1), never two (point e + 2 needs the MMAs released at point e), so arrivals of
      // different points never mix in one barrier phase, and three phases per tile and barrier give the parity below.
      auto issue_point = [&](int e, auto&& fn) {
        __syncwarp();
        if (lane == 0) mbar_arrive_a(BAR(F3Bars::EV + 2 * p + (e & 1)));
        if (wq == ((rot + e) & 3)) {
          wait_a(BAR(F3Bars::EV + 2 * p + (e & 1)), (uint32_t)(it + (e >> 1)) & 1u, kErrFfnMma1);
          tcgen05_fence_after();
          fn();
        }
      };
      if (r == 0 && it + 1 < n_it) {   // the other buffer: its last reader is the TMA store of the previous tile's rows
        if (kTmaStore && it > 0) tma_store_wait_read();
        mbar_arrive_expect_tx(&bars[F3Bars::O + 2 * p + (buf ^ 1)], kSlab);
        tma_load_2d(sAp + (buf ^ 1) * kSlab, &tmA, &bars[F3Bars::O + 2 * p + (buf ^ 1)], 0, (tile + stride) * 128);
      }
      float y[64];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xr[i][j]));
          y[16 * i + 2 * j] = f.x;
          y[16 * i + 2 * j + 1] = f.y;
        }
      }
      wait_a(BAR(F3Bars::FC + p), ph, kErrFcMma);
      tcgen05_fence_after();
      uint32_t rr[16];
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld_32x16(lane_addr + 64 + c0, rr);   // the fc result sits in D1 buffer a
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
        *reinterpret_cast<uint4*>(sAb + sw128_offset(r, c)) =
            make_uint4(pack_half2(y[8 * c], y[8 * c + 1]), pack_half2(y[8 * c + 2], y[8 * c + 3]),
                       pack_half2(y[8 * c + 4], y[8 * c + 5]), pack_half2(y[8 * c + 6], y[8 * c + 7]));
      fence_proxy_async_smem();
      tcgen05_fence_before();
      issue_point(0, [&]() {   // Y is in shared memory, the fc result has been read
        if (elect_one()) {
          issue_w1(0, dA);
          issue_w1(1, dA);
        }
        __syncwarp();
      });
      if (it + 1 < n_it) {   // the next tile's residual row, a tile ahead
#pragma unroll
        for (int i = 0; i < 4; ++i) ldg_256(x16 + (row + (int64_t)stride * 128) * 64 + 16 * i, xr[i]);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint32_t d1 = lane_addr + 64u + 64u * (uint32_t)(k & 1);
        if (k & 1) {
          wait_a(BAR(F3Bars::M1 + 2 * p + 1), n_b & 1u, kErrFfnMma1);
          ++n_b;
        } else {
          wait_a(BAR(F3Bars::M1 + 2 * p), n_a & 1u, kErrFfnMma1);
          ++n_a;
        }
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < f3_n(k) / 16; ++c) {
          tmem_ld_32x16(d1 + 16 * c, rr);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_half2_relu(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
          tmem_st_32x8(d1 + 8 * c, pk);   // over columns this thread has already loaded
        }
        tmem_wait_st();
        tcgen05_fence_before();
        issue_point(1 + k, [&]() {   // H_k is in TMEM
          if (elect_one()) {
            if (k == 0) umma_f16_ss(tACC, dOne, dBias(64, 3), idesc64, 0u);   // b2 opens the accumulator
            issue_w2(k);
            if (k + 2 < 5) issue_w1(k + 2, dA);   // same D1 buffer: executes after W2(k) has read H_k (issue order)
            if (k == 4) umma_commit_a(BAR(F3Bars::D2 + p));
          }
          __syncwarp();
          // the next tile's fc, into D1 buffer a behind W2(4): its round trip is over long before the tile starts
          if (k == 4 && it + 1 < n_it) issue_fc(buf ^ 1, ((uint32_t)(it + 1) >> 1) & 1u);
        });
      }
      wait_a(BAR(F3Bars::D2 + p), ph, kErrFfnMma2);
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
          *reinterpret_cast<uint4*>(sAb + sw128_offset(r, c)) =
              make_uint4(pack_half2(y[8 * c], y[8 * c + 1]), pack_half2(y[8 * c + 2], y[8 * c + 3]),
                         pack_half2(y[8 * c + 4], y[8 * c + 5]), pack_half2(y[8 * c + 6], y[8 * c + 7]));
        fence_proxy_async_smem();
      } else {
        out_head_epilogue(y, P, E, row);
      }
      if (kTmaStore) wg_sync();   // the output rows are staged
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
