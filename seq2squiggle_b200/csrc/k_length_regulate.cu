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
// q = tid % 8) owns the 16-byte column group q of rows r, r+32, ..., r+224 of EVERY chunk its CTA processes, so the
// decoder position encoding of its 8 x 8 elements stays in registers (the per-chunk-CTA kernel above re-reads the
// 64 KB table from L2 for every chunk), the row -> k-mer map j(t) is computed once per chunk into shared memory, and
// every store is a full 16-byte vector (a warp writes 4 consecutive 128-byte rows).  Algorithmic HBM bytes per chunk:
// 4 KB + 128 B read, 256 x 128 B + 1 KB written.
__global__ void __launch_bounds__(256) k_length_regulate16(const float* __restrict__ enc_out, const float* __restrict__ sigma,
                                                           const int32_t* __restrict__ dur, const float* __restrict__ dec_pos,
                                                           __half* __restrict__ x_dec16, float* __restrict__ sigma_ext,
                                                           int32_t* __restrict__ total, int64_t n_chunks) {
  __shared__ __align__(16) float s_x[S2S_L_ENC + 1][S2S_D];  // row 16 = zeros (positions past the last k-mer)
  __shared__ float s_sig[S2S_L_ENC + 1];
  __shared__ int s_cum[S2S_L_ENC];
  __shared__ uint8_t s_j[S2S_L_DEC_PAD];
  const int tid = threadIdx.x, r = tid >> 3, q = tid & 7;
  float pos[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = r + 32 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) pos[i][e] = t < S2S_L_DEC ? dec_pos[t * S2S_D + 8 * q + e] : 0.f;
  }
  if (tid < S2S_D) s_x[S2S_L_ENC][tid] = 0.f;
  if (tid == 0) s_sig[S2S_L_ENC] = 0.f;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    __syncthreads();  // the previous chunk's readers are done with the shared tile
    reinterpret_cast<float4*>(&s_x[0][0])[tid] = reinterpret_cast<const float4*>(enc_out + c * S2S_L_ENC * S2S_D)[tid];
    if (tid < 32) {
      int d = tid < S2S_L_ENC ? dur[c * S2S_L_ENC + tid] : 0;
      d = d < 0 ? 0 : (d > 4 * S2S_L_DEC ? 4 * S2S_L_DEC : d);
#pragma unroll
      for (int o = 1; o < S2S_L_ENC; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, d, o);
        if (tid >= o) d += n;
      }
      if (tid < S2S_L_ENC) {
        s_cum[tid] = d;
        s_sig[tid] = sigma[c * S2S_L_ENC + tid];
      }
      if (tid == S2S_L_ENC - 1) total[c] = d < S2S_L_DEC ? d : S2S_L_DEC;
    }
    __syncthreads();
    {  // j(t) = #{i : cum_i <= t}; 16 = "past the end" (zero features, zero sigma); pad rows 250..255 likewise
      int j = S2S_L_ENC;
      if (tid < S2S_L_DEC) {
        j = 0;
#pragma unroll
        for (int i = 0; i < S2S_L_ENC; ++i) j += (s_cum[i] <= tid);
        sigma_ext[c * S2S_L_DEC + tid] = s_sig[j];
      }
      s_j[tid] = (uint8_t)j;
    }
    __syncthreads();
    __half* out = x_dec16 + (size_t)c * S2S_L_DEC_PAD * S2S_D + 8 * q;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = r + 32 * i;
      const int j = s_j[t];
      const float4 a = *reinterpret_cast<const float4*>(&s_x[j][8 * q]);
      const float4 b = *reinterpret_cast<const float4*>(&s_x[j][8 * q + 4]);
      const __half2 h0 = __floats2half2_rn(a.x + pos[i][0], a.y + pos[i][1]), h1 = __floats2half2_rn(a.z + pos[i][2], a.w + pos[i][3]);
      const __half2 h2 = __floats2half2_rn(b.x + pos[i][4], b.y + pos[i][5]), h3 = __floats2half2_rn(b.z + pos[i][6], b.w + pos[i][7]);
      *reinterpret_cast<uint4*>(out + (size_t)t * S2S_D) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                     *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    }
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
    const int64_t grid = n_chunks < 8LL * sms ? n_chunks : 8LL * sms;
    k_length_regulate16<<<(unsigned)grid, 256, 0, st>>>(enc_out, sigma, dur, dec_pos, x_dec16, sigma_ext, total, n_chunks);
    S2S_LAUNCH_CHECK();
    return 0;
  }
  k_length_regulate<<<(unsigned)n_chunks, 256, 0, st>>>(enc_out, sigma, dur, dec_pos, x_dec, x_dec16, rows_per_chunk_out,
                                                        sigma_ext, total, lr_tap);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
