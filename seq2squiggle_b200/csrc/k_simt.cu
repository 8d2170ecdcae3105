// fp32 CUDA-core kernels of the FFT block (layers.py:44-142): used for the encoder (L=16) and as the
// full-precision "parity" decoder path (S2S_PREC_FP32).  Same op order as the reference:
//   q,k,v = Linear(x); S = q k^T / sqrt(d_k); P = softmax(S); O = P v; y = LN(fc(O) + x);
//   z = LN(W2 relu(W1 y + b1) + b2 + y).
#include "s2s_kernels.h"

namespace s2s {

// ---------------------------------------------------------------------------------------------
// Row-wise linear layer, 32 rows per CTA, 256 threads: thread = (column tx + 64 j, rows ty*8 .. ty*8+7).
// Wt is the transposed weight [K][N] so the 64 tx lanes read consecutive floats.
// ---------------------------------------------------------------------------------------------
template <int K, int N, int EPI>
__global__ void __launch_bounds__(256) k_linear_f32(const float* __restrict__ X, const float* __restrict__ Wt,
                                                    const float* __restrict__ bias, const float* __restrict__ R,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float* __restrict__ Y, int64_t M, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;  // fallback launch of the k-mer table path: nothing to recompute
  constexpr int TM = 32, NJ = N / 64;
  static_assert(N % 64 == 0 && K % 4 == 0, "shape");
  static_assert(EPI != EPI_BIAS_RES_LN || N == 64, "LayerNorm epilogue needs the whole row in one tile");
  __shared__ __align__(16) float sX[TM][K];
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const int64_t row0 = (int64_t)blockIdx.x * TM;

  for (int i = tid; i < TM * K / 4; i += 256) {
    int r = i / (K / 4), c4 = i - r * (K / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < M) v = *reinterpret_cast<const float4*>(X + (row0 + r) * K + 4 * c4);
    *reinterpret_cast<float4*>(&sX[r][4 * c4]) = v;
  }
  __syncthreads();

  float acc[8][NJ];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;

#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    float4 xv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) xv[r] = *reinterpret_cast<const float4*>(&sX[ty * 8 + r][k0]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) w[j] = __ldg(Wt + (size_t)(k0 + kk) * N + tx + 64 * j);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float x = kk == 0 ? xv[r].x : kk == 1 ? xv[r].y : kk == 2 ? xv[r].z : xv[r].w;
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(x, w[j], acc[r][j]);
      }
    }
  }

  if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float b = bias[tx + 64 * j];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        int64_t row = row0 + ty * 8 + r;
        if (row < M) {
          float v = acc[r][j] + b;
          if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
          Y[row * N + tx + 64 * j] = v;
        }
      }
    }
  } else {
    __syncthreads();  // everyone is done reading sX; reuse it as the [32][64] pre-LN tile
    const float b = bias[tx];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int64_t row = row0 + ty * 8 + r;
      float res = row < M ? R[row * 64 + tx] : 0.f;
      sX[ty * 8 + r][tx] = acc[r][0] + b + res;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const float g0 = gamma[lane], g1 = gamma[lane + 32], be0 = beta[lane], be1 = beta[lane + 32];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      int r = warp * 4 + rr;
      int64_t row = row0 + r;
      float a = sX[r][lane], c = sX[r][lane + 32];
      float mean = warp_sum(a + c) * (1.f / 64.f);
      float da = a - mean, dc = c - mean;
      float var = warp_sum(da * da + dc * dc) * (1.f / 64.f);
      float rstd = 1.0f / sqrtf(var + 1e-5f);
      if (row < M) {
        Y[row * 64 + lane] = da * rstd * g0 + be0;
        Y[row * 64 + lane + 32] = dc * rstd * g1 + be1;
      }
    }
  }
}

template <int K, int N, int EPI>
static int launch_linear_t(const float* X, const float* Wt, const float* b, const float* R, const float* g,
                           const float* beta, float* Y, int64_t M, cudaStream_t st, const int* run_if) {
  if (M == 0) return 0;
  k_linear_f32<K, N, EPI><<<(unsigned)ceil_div(M, 32), 256, 0, st>>>(X, Wt, b, R, g, beta, Y, M, run_if);
  S2S_LAUNCH_CHECK();
  return 0;
}

int launch_linear_f32(const float* X, const float* Wt, const float* b, const float* R, const float* g,
                      const float* beta, float* Y, int64_t M, int K, int N, int epi, cudaStream_t st, const int* run_if) {
  if (K == 64 && N == 192 && epi == EPI_BIAS) return launch_linear_t<64, 192, EPI_BIAS>(X, Wt, b, R, g, beta, Y, M, st, run_if);
  if (K == 64 && N == 192 && epi == EPI_BIAS_RELU) return launch_linear_t<64, 192, EPI_BIAS_RELU>(X, Wt, b, R, g, beta, Y, M, st, run_if);
  if (K == 64 && N == 256 && epi == EPI_BIAS_RELU) return launch_linear_t<64, 256, EPI_BIAS_RELU>(X, Wt, b, R, g, beta, Y, M, st, run_if);
  if (K == 64 && N == 64 && epi == EPI_BIAS_RES_LN) return launch_linear_t<64, 64, EPI_BIAS_RES_LN>(X, Wt, b, R, g, beta, Y, M, st, run_if);
  if (K == 256 && N == 64 && epi == EPI_BIAS_RES_LN) return launch_linear_t<256, 64, EPI_BIAS_RES_LN>(X, Wt, b, R, g, beta, Y, M, st, run_if);
  set_error("launch_linear_f32: unsupported shape K=%d N=%d epi=%d", K, N, epi);
  return -1;
}

// ---------------------------------------------------------------------------------------------
// Attention, decoder length (250 keys), one CTA per (chunk, head), one thread per query row.
// Two passes over the keys (row max, then exp / sum / PV) like torch.softmax + bmm; mask=None
// (model.py:217): all 250 positions, including the zero-filled tail, attend and are attended to.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attention_dec_f32(const float* __restrict__ qkv, float* __restrict__ out) {
  __shared__ __align__(16) float sK[S2S_L_DEC][S2S_DK];
  __shared__ __align__(16) float sV[S2S_L_DEC][S2S_DK];
  const int64_t c = blockIdx.x >> 3;
  const int h = blockIdx.x & 7;
  const int t = threadIdx.x;
  const float* base = qkv + c * S2S_L_DEC_PAD * 192;
  if (t < S2S_L_DEC) {
    const float4* kp = reinterpret_cast<const float4*>(base + (size_t)t * 192 + 64 + 8 * h);
    const float4* vp = reinterpret_cast<const float4*>(base + (size_t)t * 192 + 128 + 8 * h);
    *reinterpret_cast<float4*>(&sK[t][0]) = kp[0];
    *reinterpret_cast<float4*>(&sK[t][4]) = kp[1];
    *reinterpret_cast<float4*>(&sV[t][0]) = vp[0];
    *reinterpret_cast<float4*>(&sV[t][4]) = vp[1];
  }
  __syncthreads();
  float* op = out + (c * S2S_L_DEC_PAD + t) * 64 + 8 * h;
  if (t >= S2S_L_DEC) {  // keep the 6 pad rows finite
    *reinterpret_cast<float4*>(op) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(op + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float scale = 0.35355339059327373f;  // 1/sqrt(d_k)
  float q[8];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + (size_t)t * 192 + 8 * h);
    float4 a = qp[0], b = qp[1];
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
  }
  float m = -INFINITY;
  for (int j = 0; j < S2S_L_DEC; ++j) {
    float4 a = *reinterpret_cast<const float4*>(&sK[j][0]), b = *reinterpret_cast<const float4*>(&sK[j][4]);
    float s = q[0] * a.x;
    s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
    s = fmaf(q[4], b.x, s); s = fmaf(q[5], b.y, s); s = fmaf(q[6], b.z, s); s = fmaf(q[7], b.w, s);
    m = fmaxf(m, s * scale);
  }
  float sum = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < S2S_L_DEC; ++j) {
    float4 a = *reinterpret_cast<const float4*>(&sK[j][0]), b = *reinterpret_cast<const float4*>(&sK[j][4]);
    float s = q[0] * a.x;
    s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
    s = fmaf(q[4], b.x, s); s = fmaf(q[5], b.y, s); s = fmaf(q[6], b.z, s); s = fmaf(q[7], b.w, s);
    float p = __expf(s * scale - m);
    sum += p;
    float4 va = *reinterpret_cast<const float4*>(&sV[j][0]), vb = *reinterpret_cast<const float4*>(&sV[j][4]);
    o[0] = fmaf(p, va.x, o[0]); o[1] = fmaf(p, va.y, o[1]); o[2] = fmaf(p, va.z, o[2]); o[3] = fmaf(p, va.w, o[3]);
    o[4] = fmaf(p, vb.x, o[4]); o[5] = fmaf(p, vb.y, o[5]); o[6] = fmaf(p, vb.z, o[6]); o[7] = fmaf(p, vb.w, o[7]);
  }
  const float inv = 1.0f / sum;
  *reinterpret_cast<float4*>(op) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
  *reinterpret_cast<float4*>(op + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
}

// Encoder length (16 keys), fp32 path: one CTA of 128 threads per chunk, thread = (head, query).  (The fp16 path does this
// inside k_tc_enc_attn, csrc/k_tc_enc.cuh.)
__global__ void __launch_bounds__(128) k_attention_enc_f32(const float* __restrict__ qkv, float* __restrict__ out) {
  __shared__ __align__(16) float s[S2S_L_ENC][192];
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  const float4* src = reinterpret_cast<const float4*>(qkv + c * S2S_L_ENC * 192);
  for (int i = tid; i < S2S_L_ENC * 192 / 4; i += 128) reinterpret_cast<float4*>(&s[0][0])[i] = src[i];
  __syncthreads();
  const int h = tid >> 4, qi = tid & 15;
  const float scale = 0.35355339059327373f;
  float q[8], sc[S2S_L_ENC];
#pragma unroll
  for (int d = 0; d < 8; ++d) q[d] = s[qi][8 * h + d];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < S2S_L_ENC; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) a = fmaf(q[d], s[j][64 + 8 * h + d], a);
    sc[j] = a * scale;
    m = fmaxf(m, sc[j]);
  }
  float sum = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < S2S_L_ENC; ++j) {
    float p = __expf(sc[j] - m);
    sum += p;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = fmaf(p, s[j][128 + 8 * h + d], o[d]);
  }
  const float inv = 1.0f / sum;
  float* op = out + (c * S2S_L_ENC + qi) * 64 + 8 * h;
  *reinterpret_cast<float4*>(op) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
  *reinterpret_cast<float4*>(op + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
}

int launch_attention_f32(const float* qkv, float* out, int64_t n_chunks, int L, int rows_per_chunk, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  if (L == S2S_L_ENC && rows_per_chunk == S2S_L_ENC) {
    k_attention_enc_f32<<<(unsigned)n_chunks, 128, 0, st>>>(qkv, out);
  } else if (L == S2S_L_DEC && rows_per_chunk == S2S_L_DEC_PAD) {
    k_attention_dec_f32<<<(unsigned)(n_chunks * 8), 256, 0, st>>>(qkv, out);
  } else {
    set_error("launch_attention_f32: unsupported L=%d rows_per_chunk=%d", L, rows_per_chunk);
    return -1;
  }
  S2S_LAUNCH_CHECK();
  return 0;
}


}  // namespace s2s
