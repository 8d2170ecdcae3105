// K-C: the 64->1 heads of the three sampler MLPs, Softplus, clamps, duration draw and rounding.
//
//   modules.py:216-225  conc, rate = clamp(Softplus(MLP(emb_out)), 1e-8); Gamma(conc, rate).sample();
//                       clamp(min=1.0)
//   modules.py:275-278  sigma = Softplus(MLP(emb_out))
//   modules.py:410-437  duration_sampling: clamp(min=min_length) | dwell_std<=0: full(dwell_mean), NO clamp |
//                       normal(dwell_mean, dwell_std).clamp(min=min_length); then torch.round (half-even).int()
// The first layers (three Linear(64,64)+ReLU, concatenated to N=192) are computed by the row-wise
// linear kernel; this kernel consumes h3 [n_kmers,192].
// Random draws: Philox4x32-10 keyed by the run seed, counter = (global k-mer index, attempt, stream tag), so
// a k-mer's duration does not depend on batch boundaries or on how reads are sharded over GPUs.
#include "s2s_kernels.h"

namespace s2s {

constexpr uint32_t kStreamGamma = 0x5D0001u;
constexpr uint32_t kStreamDwellNormal = 0x5D0002u;

// Marsaglia & Tsang (2000) with the alpha<1 boost — the algorithm behind torch._standard_gamma.
__device__ float sample_standard_gamma(const Philox& ph, uint64_t idx, float alpha) {
  const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  float boost = 1.0f;
  float a = alpha;
  uint4 r = ph(lo, hi, 0u, kStreamGamma);
  if (alpha < 1.0f) {
    boost = powf(u01(r.w), 1.0f / alpha);  // u^(1/alpha); u in (0,1]
    a = alpha + 1.0f;
  }
  const float d = a - (1.0f / 3.0f);
  const float c = 1.0f / sqrtf(9.0f * d);
  for (uint32_t attempt = 0; attempt < 64u; ++attempt) {
    if (attempt) r = ph(lo, hi, attempt, kStreamGamma);
    float x = box_muller(r.x, r.y).x;
    float v = 1.0f + c * x;
    if (v <= 0.0f) continue;
    v = v * v * v;
    float u = u01(r.z);
    float x2 = x * x;
    if (u < 1.0f - 0.0331f * x2 * x2 || logf(u) < 0.5f * x2 + d * (1.0f - v + logf(v))) return boost * d * v;
  }
  return boost * d;  // unreachable in practice (acceptance > 95% per attempt)
}

// clamps, duration draw and rounding of one k-mer given its (conc, rate, sigma)
__device__ __forceinline__ void finish_sampler(int64_t row, float conc, float rate, float sg, const s2s_run_opts& o,
                                               float* __restrict__ sigma, int32_t* __restrict__ dur_int,
                                               float* __restrict__ conc_tap, float* __restrict__ rate_tap,
                                               float* __restrict__ dur_float_tap) {
  sigma[row] = sg;
  if (conc_tap) conc_tap[row] = conc;
  if (rate_tap) rate_tap[row] = rate;
  const uint64_t gidx = o.chunk_id_base * S2S_L_ENC + (uint64_t)row;
  Philox ph(o.seed);
  float d;
  if (o.duration_mode == S2S_DUR_SAMPLER) {
    d = sample_standard_gamma(ph, gidx, conc) / rate;
    d = fmaxf(d, 1.17549435e-38f);  // torch.distributions.Gamma.rsample clamps to finfo.tiny
    d = fmaxf(d, 1.0f);             // modules.py:223
    d = fmaxf(d, o.min_duration);   // modules.py:414
  } else if (o.duration_mode == S2S_DUR_NORMAL) {
    uint4 r = ph((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, kStreamDwellNormal);
    d = o.dwell_mean + o.dwell_std * box_muller(r.x, r.y).x;
    d = fmaxf(d, o.min_duration);   // modules.py:430
  } else {
    d = o.dwell_mean;               // modules.py:420: no clamp in the constant branch
  }
  if (dur_float_tap) dur_float_tap[row] = d;
  dur_int[row] = (int32_t)rintf(d);  // torch.round: half to even
}

__global__ void __launch_bounds__(256) k_sampler_lookup(const float4* __restrict__ tab_smp, const int32_t* __restrict__ kidx,
                                                        int64_t n_kmers, s2s_run_opts o, float* __restrict__ sigma,
                                                        int32_t* __restrict__ dur_int, float* __restrict__ conc_tap,
                                                        float* __restrict__ rate_tap, float* __restrict__ dur_float_tap) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_kmers) return;
  const int32_t idx = kidx[row];
  if (idx < 0) return;
  const float4 t = tab_smp[idx];
  finish_sampler(row, t.x, t.y, t.z, o, sigma, dur_int, conc_tap, rate_tap, dur_float_tap);
}

__global__ void k_pack_smp_table(const float* __restrict__ conc, const float* __restrict__ rate,
                                 const float* __restrict__ sigma, int64_t n, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(conc[i], rate[i], sigma[i], 0.f);
}

__global__ void __launch_bounds__(256) k_sampler_heads(const float* __restrict__ h3, const float* __restrict__ w3,
                                                       const float* __restrict__ b3, int64_t n_kmers, s2s_run_opts o,
                                                       float* __restrict__ sigma, int32_t* __restrict__ dur_int,
                                                       float* __restrict__ conc_tap, float* __restrict__ rate_tap,
                                                       float* __restrict__ dur_float_tap, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t row0 = warp_global * 32;
  if (row0 >= n_kmers) return;
  // lane keeps weights for columns lane and lane+32 of each head
  float w[3][2];
#pragma unroll
  for (int m = 0; m < 3; ++m) { w[m][0] = w3[m * 64 + lane]; w[m][1] = w3[m * 64 + lane + 32]; }
  float mine[3] = {0.f, 0.f, 0.f};
  const int n_here = (int)((n_kmers - row0) < 32 ? (n_kmers - row0) : 32);
  for (int i = 0; i < n_here; ++i) {
    const float* hp = h3 + (row0 + i) * 192;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float s = warp_sum(hp[m * 64 + lane] * w[m][0] + hp[m * 64 + lane + 32] * w[m][1]);
      if (lane == i) mine[m] = s;
    }
  }
  if (lane >= n_here) return;
  const int64_t row = row0 + lane;
  const float conc = fmaxf(softplus_torch(mine[0] + b3[0]), 1e-8f);
  const float rate = fmaxf(softplus_torch(mine[1] + b3[1]), 1e-8f);
  const float sg = softplus_torch(mine[2] + b3[2]);
  finish_sampler(row, conc, rate, sg, o, sigma, dur_int, conc_tap, rate_tap, dur_float_tap);
}

int launch_sampler_heads(const DevWeights& w, const float* h3, int64_t n_kmers, const s2s_run_opts& o, float* sigma,
                         int32_t* dur_int, float* conc_tap, float* rate_tap, float* dur_float_tap, cudaStream_t st,
                         const int* run_if) {
  if (n_kmers == 0) return 0;
  int64_t warps = ceil_div(n_kmers, 32);
  k_sampler_heads<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(h3, w.smp3_w, w.smp3_b, n_kmers, o, sigma, dur_int,
                                                                conc_tap, rate_tap, dur_float_tap, run_if);
  S2S_LAUNCH_CHECK();
  return 0;
}

int launch_sampler_lookup(const KmerTables& tab, const int32_t* kidx, int64_t n_kmers, const s2s_run_opts& o,
                          float* sigma, int32_t* dur_int, float* conc_tap, float* rate_tap, float* dur_float_tap,
                          cudaStream_t st) {
  if (n_kmers == 0) return 0;
  k_sampler_lookup<<<(unsigned)ceil_div(n_kmers, 256), 256, 0, st>>>(tab.smp, kidx, n_kmers, o, sigma, dur_int, conc_tap,
                                                                     rate_tap, dur_float_tap);
  S2S_LAUNCH_CHECK();
  return 0;
}

int launch_pack_smp_table(const float* conc, const float* rate, const float* sigma, int64_t n, float4* out, cudaStream_t st) {
  k_pack_smp_table<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(conc, rate, sigma, n, out);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
