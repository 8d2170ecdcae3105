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

int launch_length_regulate(const float* enc_out, const float* sigma, const int32_t* dur, int64_t n_chunks,
                           const float* dec_pos, float* x_dec, __half* x_dec16, int rows_per_chunk_out,
                           float* sigma_ext, int32_t* total, float* lr_tap, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  k_length_regulate<<<(unsigned)n_chunks, 256, 0, st>>>(enc_out, sigma, dur, dec_pos, x_dec, x_dec16, rows_per_chunk_out,
                                                        sigma_ext, total, lr_tap);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
