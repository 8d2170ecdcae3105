// K-A front end: chunk map, on-device tokeniser, k-mer embedding + prenet.
//
// Reference semantics reproduced (paths relative to /root/reference/src/seq2squiggle):
//   utils.py:334-356  extract_kmers / add_remainder / split_sequence: a read of n bases gives n-k+1
//                     overlapping k-mers, padded with "_"*k k-mers to a multiple of 16, cut into
//                     windows of 16 k-mers;
//   utils.py:56-89    one_hot_encode: "_ACGT" -> 0..4, any other byte (lower case, N, ...) -> zero row;
//   modules.py:70-80  ReLU(src_emb(onehot)), ReLU(pre_net(.)) = emb_out, enc input = emb_out + position_enc.
// The one-hot GEMM [16,5k]x[5k,64] is a gather-sum of <= k columns of src_emb.weight, so no one-hot tensor
// is ever materialised: 24 bytes of bases per chunk are the only input traffic.
#include "s2s_kernels.h"

namespace s2s {

__global__ void k_chunk_map(const int64_t* __restrict__ read_offsets, const int64_t* __restrict__ chunk_offsets,
                            int64_t n_reads, int64_t n_chunks, int k, int32_t* __restrict__ chunk_read,
                            int64_t* __restrict__ chunk_base, int32_t* __restrict__ chunk_nk) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  // largest r with chunk_offsets[r] <= c  (reads with zero chunks are skipped by the upper bound)
  int64_t lo = 0, hi = n_reads;  // invariant: chunk_offsets[lo] <= c < chunk_offsets[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (chunk_offsets[mid] <= c) lo = mid; else hi = mid;
  }
  int64_t ci = c - chunk_offsets[lo];
  int64_t len = read_offsets[lo + 1] - read_offsets[lo];
  int64_t n_kmers = len - k + 1;
  chunk_read[c] = (int32_t)lo;
  chunk_base[c] = read_offsets[lo] + ci * S2S_L_ENC;
  int64_t nk = n_kmers - ci * S2S_L_ENC;
  chunk_nk[c] = (int32_t)(nk > S2S_L_ENC ? S2S_L_ENC : nk);
}

__device__ __forceinline__ int letter_code(uint8_t ch) {
  // utils.py:74 letter_to_int, case-sensitive; everything else has no one-hot column
  switch (ch) {
    case '_': return 0;
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'T': return 4;
    default: return -1;
  }
}

// One CTA (256 threads) per chunk: thread = (k-mer j = tid/16, 4 channels cg = tid%16).
__global__ void __launch_bounds__(256) k_embed(const float* __restrict__ src_t, const float* __restrict__ src_b,
                                               const float* __restrict__ pre_t, const float* __restrict__ pre_b,
                                               const float* __restrict__ enc_pos, int k,
                                               const uint8_t* __restrict__ bases, const int64_t* __restrict__ chunk_base,
                                               const int32_t* __restrict__ chunk_nk, const int8_t* __restrict__ codes,
                                               int64_t n_chunks, float* __restrict__ emb_out, float* __restrict__ x_enc,
                                               __half* __restrict__ x_enc16, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;
  __shared__ int8_t s_code[S2S_L_ENC][12];
  __shared__ __align__(16) float s_e1[S2S_L_ENC][S2S_D];
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < S2S_L_ENC * k) {
    int j = tid / k, i = tid - j * k;
    int code;
    if (codes != nullptr) {
      code = codes[(c * S2S_L_ENC + j) * k + i];
    } else {
      code = (j < chunk_nk[c]) ? letter_code(bases[chunk_base[c] + j + i]) : 0;  // "_"*k padding k-mers
    }
    s_code[j][i] = (int8_t)code;
  }
  __syncthreads();
  const int j = tid >> 4, cg = tid & 15;
  float4 acc = *reinterpret_cast<const float4*>(src_b + 4 * cg);
  for (int i = 0; i < k; ++i) {
    int code = s_code[j][i];
    if (code >= 0) {
      float4 w = *reinterpret_cast<const float4*>(src_t + (size_t)(5 * i + code) * S2S_D + 4 * cg);
      acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
    }
  }
  acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
  *reinterpret_cast<float4*>(&s_e1[j][4 * cg]) = acc;
  __syncthreads();
  float4 o = *reinterpret_cast<const float4*>(pre_b + 4 * cg);
#pragma unroll 8
  for (int kk = 0; kk < S2S_D; ++kk) {
    float e = s_e1[j][kk];
    float4 w = *reinterpret_cast<const float4*>(pre_t + (size_t)kk * S2S_D + 4 * cg);
    o.x = fmaf(e, w.x, o.x); o.y = fmaf(e, w.y, o.y); o.z = fmaf(e, w.z, o.z); o.w = fmaf(e, w.w, o.w);
  }
  o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  const size_t off = ((size_t)c * S2S_L_ENC + j) * S2S_D + 4 * cg;
  *reinterpret_cast<float4*>(emb_out + off) = o;
  float4 p = *reinterpret_cast<const float4*>(enc_pos + j * S2S_D + 4 * cg);
  o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
  *reinterpret_cast<float4*>(x_enc + off) = o;
  if (x_enc16) {
    __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
    *reinterpret_cast<uint2*>(x_enc16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
}

int launch_chunk_map(const int64_t* read_offsets, const int64_t* chunk_offsets, int64_t n_reads, int64_t n_chunks,
                     int32_t k, int32_t* chunk_read, int64_t* chunk_base, int32_t* chunk_nk, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  k_chunk_map<<<(unsigned)ceil_div(n_chunks, 256), 256, 0, st>>>(read_offsets, chunk_offsets, n_reads, n_chunks, k,
                                                                 chunk_read, chunk_base, chunk_nk);
  S2S_LAUNCH_CHECK();
  return 0;
}

int launch_embed(const DevWeights& w, const uint8_t* bases, const int64_t* chunk_base, const int32_t* chunk_nk,
                 const int8_t* codes, int64_t n_chunks, float* emb_out, float* x_enc, __half* x_enc16, cudaStream_t st,
                 const int* run_if) {
  if (n_chunks == 0) return 0;
  k_embed<<<(unsigned)n_chunks, 256, 0, st>>>(w.src_t, w.src_b, w.pre_t, w.pre_b, w.enc_pos, w.cfg.seq_kmer, bases,
                                              chunk_base, chunk_nk, codes, n_chunks, emb_out, x_enc, x_enc16, run_if);
  S2S_LAUNCH_CHECK();
  return 0;
}

// ---- per-k-mer tables ------------------------------------------------------------------------------
__global__ void k_all_kmer_codes(int k, int64_t n_rows, int64_t n_kmers, int8_t* __restrict__ codes) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int64_t v = r;
  for (int i = k - 1; i >= 0; --i) {  // most significant letter first
    codes[r * k + i] = r < n_kmers ? (int8_t)(1 + (v & 3)) : (int8_t)0;
    v >>= 2;
  }
}

int launch_all_kmer_codes(int k, int64_t n_rows, int8_t* codes, cudaStream_t st) {
  k_all_kmer_codes<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(k, n_rows, (int64_t)1 << (2 * k), codes);
  S2S_LAUNCH_CHECK();
  return 0;
}

// One CTA (256 threads) per chunk like k_embed: thread = (k-mer j = tid/16, 4 channels cg = tid%16).  The k-mer's
// table index is the base-4 number of its letters (A,C,G,T = 0..3); "_"*k is entry 4^k; anything else (a letter
// outside "_ACGT", or "_" mixed with bases, which only hand-made code tensors can contain) is not in the table.
__global__ void __launch_bounds__(256) k_embed_lookup(const float* __restrict__ tab_emb, int64_t n_kmers, int* __restrict__ flag,
                                                      const float* __restrict__ enc_pos, int k,
                                                      const uint8_t* __restrict__ bases, const int64_t* __restrict__ chunk_base,
                                                      const int32_t* __restrict__ chunk_nk, const int8_t* __restrict__ codes,
                                                      float* __restrict__ emb_out, float* __restrict__ x_enc,
                                                      __half* __restrict__ x_enc16, int32_t* __restrict__ kidx) {
  __shared__ int32_t s_idx[S2S_L_ENC];
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < S2S_L_ENC) {
    const int j = tid;
    int64_t idx = 0;
    int n_pad = 0;
    bool bad = false;
    const bool is_pad_kmer = codes == nullptr && j >= chunk_nk[c];
    for (int i = 0; i < k; ++i) {
      int code;
      if (codes != nullptr) code = codes[(c * S2S_L_ENC + j) * k + i];
      else code = is_pad_kmer ? 0 : letter_code(bases[chunk_base[c] + j + i]);
      if (code < 0) bad = true;
      else if (code == 0) ++n_pad;
      else idx = idx * 4 + (code - 1);
    }
    int32_t out;
    if (bad || (n_pad != 0 && n_pad != k)) out = -1;
    else out = n_pad == k ? (int32_t)n_kmers : (int32_t)idx;
    if (out < 0) atomicOr(flag, 1);
    s_idx[j] = out;
    kidx[c * S2S_L_ENC + j] = out;
  }
  __syncthreads();
  const int j = tid >> 4, cg = tid & 15;
  const int32_t idx = s_idx[j];
  if (idx < 0) return;  // the fallback launches recompute the whole sub-batch
  float4 o = *reinterpret_cast<const float4*>(tab_emb + (size_t)idx * S2S_D + 4 * cg);
  const size_t off = ((size_t)c * S2S_L_ENC + j) * S2S_D + 4 * cg;
  *reinterpret_cast<float4*>(emb_out + off) = o;
  float4 p = *reinterpret_cast<const float4*>(enc_pos + j * S2S_D + 4 * cg);
  o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
  *reinterpret_cast<float4*>(x_enc + off) = o;
  if (x_enc16) {
    __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
    *reinterpret_cast<uint2*>(x_enc16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
}

int launch_embed_lookup(const DevWeights& w, const KmerTables& tab, const uint8_t* bases, const int64_t* chunk_base,
                        const int32_t* chunk_nk, const int8_t* codes, int64_t n_chunks, float* emb_out, float* x_enc,
                        __half* x_enc16, int32_t* kidx, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  k_embed_lookup<<<(unsigned)n_chunks, 256, 0, st>>>(tab.emb, tab.n_kmers, tab.flag, w.enc_pos, w.cfg.seq_kmer, bases,
                                                     chunk_base, chunk_nk, codes, emb_out, x_enc, x_enc16, kidx);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
