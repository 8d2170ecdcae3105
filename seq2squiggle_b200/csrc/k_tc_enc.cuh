// k_tc_enc_attn — encoder block, first half: QKV projection + the 16-key self-attention in one kernel (included by
// k_tc.cu inside namespace s2s::{anonymous}).
//
// layers.py:64-86 for the encoder rows (16 per chunk, 8 chunks per 128-row tile): q | k | v = X Wqkv^T + b on the tensor
// core ([128 x 192] accumulator in TMEM), then softmax(q k^T / sqrt(8)) v per head over the row's OWN chunk on the CUDA
// cores, straight from the fp32 accumulators: K and V go to shared memory as fp32 (never to HBM, never rounded to fp16), q
// stays in the thread's registers.  Round 1 wrote q | k | v as fp16 (384 B per row) with k_tc_qkv_plain and read it back
// in k_attention_enc_f32, one CTA per chunk.
//   two threads per row of the tile (warps 0-3: heads 0-3 and the K half of the accumulator, warps 4-7: heads 4-7 and the V
//   half), 16 warps per SM; keys of a row = the 16 rows of its chunk, so the 16 threads of a half-warp read the same K / V
//   address (broadcast) and the two chunks of a warp are skewed by 16 bytes so that they sit in different banks.
//   The next tile's MMA is issued as soon as every thread has read the accumulator, and runs under the attention math.
// Shared memory: Wqkv 24 KB | X tile 16 KB | K,V fp32 [128 rows x 132 floats + skew] 66 KB = 107 KB -> 2 CTAs / SM.
#pragma once

constexpr int kEncKvStride = 132;                                    // floats per row: K [0,64) | V [64,128) | 4 pad
constexpr int kEncKvBytes = 128 * kEncKvStride * 4 + 8 * 16;         // + 16 bytes of skew per chunk
constexpr int kSmemEncAttn = 192 * 128 + kSlab + kEncKvBytes + 1024;

__global__ void __launch_bounds__(256, 2) k_tc_enc_attn(const __grid_constant__ CUtensorMap tmX,
                                                        const __grid_constant__ CUtensorMap tmW,
                                                        const float* __restrict__ bias, __half* __restrict__ o16,
                                                        int64_t n_rows, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_mma;
  __shared__ uint32_t s_tmem;
  __shared__ int s_abort, s_go;
  __shared__ float s_bias[192];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW = smem;                                       // [192 x 128 B]
  uint8_t* sA = smem + 192 * 128;                           // [128 x 128 B]
  float* sKV = reinterpret_cast<float*>(smem + 192 * 128 + kSlab);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r_t = tid & 127, g = tid >> 7;   // row of the tile, head group (heads 4g .. 4g+3)
  const int n_tiles = (int)((n_rows + 127) / 128);
  if (tid == 0) s_go = (*status == 0);
  __syncthreads();
  if (!s_go) return;
  if (warp == 0) tmem_alloc<256>(&s_tmem);
  if (tid == 0) {
    mbar_init(&bar_w, 1); mbar_init(&bar_a, 1); mbar_init(&bar_mma, 1);
    fence_mbar_init();
    s_abort = 0;
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW);
  }
  if (tid < 192) s_bias[tid] = bias[tid];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = umma_idesc(128, 192, kFmtF16);
  const uint32_t lane_addr = tmem_addr(tmem, (warp & 3) * 32, 0);
  // row r of the tile: chunk r >> 4, skewed by 4 floats per chunk
  const uint32_t sKV_a = smem_u32(sKV);
  auto kv_row = [&](int r) { return sKV_a + 4u * (uint32_t)(r * kEncKvStride + 4 * (r >> 4)); };   // shared address
  auto issue_mma = [&]() {   // thread 0: QKV of the tile in sA
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sW);
#pragma unroll
    for (int s = 0; s < 4; ++s) umma_f16_ss(tmem, umma_desc_k_sw128(a0 + s * 32), umma_desc_k_sw128(b0 + s * 32), idesc, s > 0);
    umma_commit(&bar_mma);
  };
  int tile = blockIdx.x;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_w, 192 * 128);
    tma_load_2d(sW, &tmW, &bar_w, 0, 0);
    if (tile < n_tiles) {
      mbar_arrive_expect_tx(&bar_a, kSlab);
      tma_load_2d(sA, &tmX, &bar_a, 0, tile * 128);   // rows past n_rows are zero-filled by TMA
    }
  }
  if (tid == 0 && tile < n_tiles) {
    wait_bar(&bar_w, 0, status, &s_abort, kErrQkvLoad);
    wait_bar(&bar_a, 0, status, &s_abort, kErrQkvLoad);
    tcgen05_fence_after();
    issue_mma();
  }
  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int next = tile + gridDim.x;
    wait_bar(&bar_mma, it & 1, status, &s_abort, kErrQkvMma);
    tcgen05_fence_after();
    if (tid == 0 && next < n_tiles) {   // the X tile has been consumed by the MMA: the next one may land
      mbar_arrive_expect_tx(&bar_a, kSlab);
      tma_load_2d(sA, &tmX, &bar_a, 0, next * 128);
    }
    float q[32];
    {
      uint32_t r[32];
      const uint32_t dst = kv_row(r_t) + 256u * g;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {   // k (g = 0) or v (g = 1) of this row -> shared memory, fp32
        const int c0 = 64 + 64 * g + c;
        tmem_ld_32x32(lane_addr + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts_f4(dst + 4u * (c + 4 * i),
              make_float4(__uint_as_float(r[4 * i]) + s_bias[c0 + 4 * i], __uint_as_float(r[4 * i + 1]) + s_bias[c0 + 4 * i + 1],
                          __uint_as_float(r[4 * i + 2]) + s_bias[c0 + 4 * i + 2], __uint_as_float(r[4 * i + 3]) + s_bias[c0 + 4 * i + 3]));
      }
      tmem_ld_32x32(lane_addr + 32 * g, r);   // q of the thread's four heads, scaled by 1 / sqrt(d_k) (layers.py:21)
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) q[i] = (__uint_as_float(r[i]) + s_bias[32 * g + i]) * 0.35355339059327373f;
    }
    tcgen05_fence_before();
    __syncthreads();   // K, V of the tile are in shared memory; the accumulator has been read
    if (tid == 0 && next < n_tiles) {   // the next tile's projection runs under this tile's attention
      wait_bar(&bar_a, (it + 1) & 1, status, &s_abort, kErrQkvLoad);
      tcgen05_fence_after();
      issue_mma();
    }
    const uint32_t kv0 = kv_row(r_t & ~15) + 128u * g;   // first row of this row's chunk, the thread's four heads
    uint32_t out[16];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float sc[S2S_L_ENC];
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < S2S_L_ENC; ++j) {
        const float4 ka = lds_f4(kv0 + 4u * (j * kEncKvStride + 8 * h));
        const float4 kb = lds_f4(kv0 + 4u * (j * kEncKvStride + 8 * h + 4));
        float a = q[8 * h] * ka.x;
        a = fmaf(q[8 * h + 1], ka.y, a); a = fmaf(q[8 * h + 2], ka.z, a); a = fmaf(q[8 * h + 3], ka.w, a);
        a = fmaf(q[8 * h + 4], kb.x, a); a = fmaf(q[8 * h + 5], kb.y, a); a = fmaf(q[8 * h + 6], kb.z, a);
        a = fmaf(q[8 * h + 7], kb.w, a);
        sc[j] = a;
        m = fmaxf(m, a);
      }
      float sum = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < S2S_L_ENC; ++j) {
        const float p = __expf(sc[j] - m);
        sum += p;
        const float4 va = lds_f4(kv0 + 4u * (j * kEncKvStride + 64 + 8 * h));
        const float4 vb = lds_f4(kv0 + 4u * (j * kEncKvStride + 64 + 8 * h + 4));
        o[0] = fmaf(p, va.x, o[0]); o[1] = fmaf(p, va.y, o[1]); o[2] = fmaf(p, va.z, o[2]); o[3] = fmaf(p, va.w, o[3]);
        o[4] = fmaf(p, vb.x, o[4]); o[5] = fmaf(p, vb.y, o[5]); o[6] = fmaf(p, vb.z, o[6]); o[7] = fmaf(p, vb.w, o[7]);
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int d = 0; d < 4; ++d) out[4 * h + d] = pack_half2(o[2 * d] * inv, o[2 * d + 1] * inv);
    }
    const int64_t row = (int64_t)tile * 128 + r_t;
    if (row < n_rows) {
      __half* op = o16 + row * 64 + 32 * g;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + 16 * i), "r"(out[8 * i]),
                     "r"(out[8 * i + 1]), "r"(out[8 * i + 2]), "r"(out[8 * i + 3]), "r"(out[8 * i + 4]), "r"(out[8 * i + 5]),
                     "r"(out[8 * i + 6]), "r"(out[8 * i + 7])
                     : "memory");
    }
    __syncthreads();   // K, V may be overwritten by the next tile
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}
