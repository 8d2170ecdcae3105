// K-D length regulator (modules.py:344-392 LengthRegulator.LR + modules.py:136 decoder position add).
//
// The reference builds a 0/1 alignment matrix M[b,j,t] = [t < cum_j] - [t < cum_{j-1}] and multiplies
// (bmm) it with the encoder output; that is a gather: out[t] = x[j(t)] with j(t) = #{i : cum_i <= t} for
// t < cum_15, zero rows after, cropped / zero-padded to 250.  Here: warp-shuffle inclusive scan of the 16
// integer durations, the 16x64 encoder tile staged in shared memory, and 128-bit coalesced row stores.
// HBM bytes per chunk: 4 KB + 128 B read, 64 KB (fp32 rows) + 1 KB written.
#include "s2s_kernels.h"

namespace s2s {

__global__ void __launch_bounds__(256) k_length_regulate(const float* __restrict__ enc_out,
                                                         const float* __restrict__ sigma,
                                                         const int32_t* __restrict__ dur,
                                                         const float* __restrict__ dec_pos, float* __restrict__ x_dec,
                                                         __half* __restrict__ x_dec16, int rows_out, float* __restrict__ sigma_ext,
                                                         int32_t* __restrict__ total, float* __restrict__ lr_tap) {
  __shared__ __align__(16) float s_x[S2S_L_ENC][S2S_D];
  __shared__ float s_sig[S2S_L_ENC];
  __shared__ int s_cum[S2S_L_ENC];
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  reinterpret_cast<float4*>(&s_x[0][0])[tid] = reinterpret_cast<const float4*>(enc_out + c * S2S_L_ENC * S2S_D)[tid];
  if (tid < 32) {
    int d = tid < S2S_L_ENC ? dur[c * S2S_L_ENC + tid] : 0;
    d = d < 0 ? 0 : (d > 4 * S2S_L_DEC ? 4 * S2S_L_DEC : d);  // keeps the scan in int range; >=250 already fills the chunk
#pragma unroll
    for (int o = 1; o < S2S_L_ENC; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, d, o);
      if (tid >= o) d += n;
    }
    if (tid < S2S_L_ENC) {
      s_cum[tid] = d;
      s_sig[tid] = sigma ? sigma[c * S2S_L_ENC + tid] : 0.f;
    }
    if (tid == S2S_L_ENC - 1 && total) total[c] = d < S2S_L_DEC ? d : S2S_L_DEC;
  }
  __syncthreads();
  const int q = tid & 15;  // float4 column of the 64-wide row
#pragma unroll 4
  for (int t = tid >> 4; t < rows_out; t += 16) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < S2S_L_DEC) {
      int j = 0;
#pragma unroll
      for (int i = 0; i < S2S_L_ENC; ++i) j += (s_cum[i] <= t);
      float sg = 0.f;
      if (j < S2S_L_ENC) {
        v = *reinterpret_cast<const float4*>(&s_x[j][4 * q]);
        sg = s_sig[j];
      }
      if (lr_tap) *reinterpret_cast<float4*>(lr_tap + ((size_t)c * S2S_L_DEC + t) * S2S_D + 4 * q) = v;
      if (q == 0 && sigma_ext) sigma_ext[c * S2S_L_DEC + t] = sg;
      if (dec_pos) {
        float4 p = *reinterpret_cast<const float4*>(dec_pos + t * S2S_D + 4 * q);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
    }
    if (x_dec) *reinterpret_cast<float4*>(x_dec + ((size_t)c * rows_out + t) * S2S_D + 4 * q) = v;
    if (x_dec16) {
      __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(x_dec16 + ((size_t)c * rows_out + t) * S2S_D + 4 * q) =
          make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
  }
}

// Tensor-core path variant: fp16 rows only (the decoder's residual stream), persistent CTAs.  Thread (r = tid / 8,
// q = tid % 8) owns the 16-byte column group q of rows r, r+64, r+128, r+192 of EVERY chunk its CTA processes, so the
// decoder position encoding of its 8 x 8 elements stays in registers (the per-chunk-CTA kernel above re-reads the
// 64 KB table from L2 for every chunk) and every store is a full 16-byte vector (a warp writes 4 consecutive 128-byte
// rows).  Software pipeline with ONE block barrier per chunk: while the rows of chunk c are stored from one half of the
// double-buffered shared tile, the encoder tile of the CTA's next chunk is loaded into the other half and its row -> k-mer
// map j(t) is computed — every thread scans the 16 durations itself (64 bytes, the same for the whole CTA), so the map
// needs no barrier of its own.  (Round 1: three barriers per chunk and no overlap between the loads of chunk c+1 and the
// stores of chunk c: barrier stalls 9.9 per issued instruction, 3.4 TB/s.)  Algorithmic HBM bytes per chunk: 4 KB + 128 B
// read, 256 x 128 B + 1 KB written.
constexpr int kLr16Threads = 512, kLr16Rows = 4;   // thread (r = tid / 8, q = tid % 8): rows r + 64 i, i < 4

__global__ void __launch_bounds__(kLr16Threads, 2) k_length_regulate16(const float* __restrict__ enc_out, const float* __restrict__ sigma,
                                                           const int32_t* __restrict__ dur, const float* __restrict__ dec_pos,
                                                           __half* __restrict__ x_dec16, float* __restrict__ sigma_ext,
                                                           int32_t* __restrict__ total, int64_t n_chunks) {
  __shared__ __align__(16) float s_x[2][S2S_L_ENC + 1][S2S_D];  // row 16 = zeros (positions past the last k-mer)
  __shared__ uint8_t s_j[2][S2S_L_DEC_PAD];
  const int tid = threadIdx.x, r = tid >> 3, q = tid & 7;
  float pos[kLr16Rows][8];
#pragma unroll
  for (int i = 0; i < kLr16Rows; ++i) {
    const int t = r + 64 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) pos[i][e] = t < S2S_L_DEC ? dec_pos[t * S2S_D + 8 * q + e] : 0.f;
  }
  if (tid < S2S_D) s_x[0][S2S_L_ENC][tid] = s_x[1][S2S_L_ENC][tid] = 0.f;
  // stage chunk c into buffer b: encoder tile -> shared, j(t) of row t = tid -> shared, sigma_ext / total -> global
  auto stage = [&](int64_t c, int b) {
    if (tid < 256) {
      reinterpret_cast<float4*>(&s_x[b][0][0])[tid] = reinterpret_cast<const float4*>(enc_out + c * S2S_L_ENC * S2S_D)[tid];
    } else {
      const int t = tid - 256;   // the row whose k-mer index this thread computes
      const int4* dp = reinterpret_cast<const int4*>(dur + c * S2S_L_ENC);
      int cum = 0, j = 0;
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const int4 d4 = dp[g4];
        const int dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = dv[e] < 0 ? 0 : (dv[e] > 4 * S2S_L_DEC ? 4 * S2S_L_DEC : dv[e]);  // keeps the scan in int range
          cum += d;
          j += (cum <= t);      // j(t) = #{i : cum_i <= t}
        }
      }
      if (t >= S2S_L_DEC) j = S2S_L_ENC;   // pad rows 250..255: zero features
      s_j[b][t] = (uint8_t)j;
      if (t < S2S_L_DEC) sigma_ext[c * S2S_L_DEC + t] = j < S2S_L_ENC ? sigma[c * S2S_L_ENC + j] : 0.f;
      if (t == 0) total[c] = cum < S2S_L_DEC ? cum : S2S_L_DEC;
    }
  };
  int64_t c = blockIdx.x;
  if (c < n_chunks) stage(c, 0);
  __syncthreads();
  for (int b = 0; c < n_chunks; c += gridDim.x, b ^= 1) {
    const int64_t cn = c + gridDim.x;
    if (cn < n_chunks) stage(cn, b ^ 1);   // loads of the next chunk are in flight under this chunk's stores
    __half* out = x_dec16 + (size_t)c * S2S_L_DEC_PAD * S2S_D + 8 * q;
#pragma unroll
    for (int i = 0; i < kLr16Rows; ++i) {
      const int t = r + 64 * i;
      const int j = s_j[b][t];
      const float4 a = *reinterpret_cast<const float4*>(&s_x[b][j][8 * q]);
      const float4 bb = *reinterpret_cast<const float4*>(&s_x[b][j][8 * q + 4]);
      const __half2 h0 = __floats2half2_rn(a.x + pos[i][0], a.y + pos[i][1]), h1 = __floats2half2_rn(a.z + pos[i][2], a.w + pos[i][3]);
      const __half2 h2 = __floats2half2_rn(bb.x + pos[i][4], bb.y + pos[i][5]), h3 = __floats2half2_rn(bb.z + pos[i][6], bb.w + pos[i][7]);
      *reinterpret_cast<uint4*>(out + (size_t)t * S2S_D) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                     *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
    __syncthreads();   // buffer b ^ 1 is complete; buffer b may be overwritten by the next iteration's stage()
  }
}

int launch_length_regulate(const float* enc_out, const float* sigma, const int32_t* dur, int64_t n_chunks,
                           const float* dec_pos, float* x_dec, __half* x_dec16, int rows_per_chunk_out,
                           float* sigma_ext, int32_t* total, float* lr_tap, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  if (x_dec16 && !x_dec && !lr_tap && dec_pos && sigma && sigma_ext && total && rows_per_chunk_out == S2S_L_DEC_PAD) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t grid = n_chunks < 2LL * sms ? n_chunks : 2LL * sms;   // persistent: two 512-thread CTAs per SM
    k_length_regulate16<<<(unsigned)grid, kLr16Threads, 0, st>>>(enc_out, sigma, dur, dec_pos, x_dec16, sigma_ext, total, n_chunks);
    S2S_LAUNCH_CHECK();
    return 0;
  }
  k_length_regulate<<<(unsigned)n_chunks, 256, 0, st>>>(enc_out, sigma, dur, dec_pos, x_dec, x_dec16, rows_per_chunk_out,
                                                        sigma_ext, total, lr_tap);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
