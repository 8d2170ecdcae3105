// Internal launcher declarations (C++; the C-ABI lives in include/s2s_b200.h / s2s_api.cu).
#pragma once
#include "s2s_common.cuh"
#include "s2s_weights.h"

namespace s2s {

// Per-column vectors of one FFT block's fc + LayerNorm + FFN + LayerNorm tail, passed to k_tc_fc_ffn BY VALUE as a
// __grid_constant__ kernel parameter (constant-bank operands).  wout / bout: out_linear (modules.py:140) — every
// decoder block carries a copy, only the last block's kernel reads it.
struct FfnParams {
  float bfc[64], g1[64], be1[64], b1[256], b2[64], g2[64], be2[64], wout[64];
  float bout;
};

// Device-resident derived weights, built once by s2s_create().
struct BlockDev {
  FfnParams ffn;               // host copy (kernel parameter)
  // fp32, transposed to [K][N] so a warp reads consecutive output columns (SIMT path)
  const float *wqkv_t, *bqkv;  // [64][192], [192]   (q | k | v)
  const float *fc_t, *fc_b;    // [64][64]
  const float *w1_t, *b1;      // [64][256]
  const float *w2_t, *b2;      // [256][64]
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  // fp16, K-major [N][K] (PyTorch's own [out][in] order) for the tcgen05 path
  const __half *wqkv_h;        // [192][64]
  const __half *fc_h;          // [64][64]
  const __half *w1_h;          // [256][64]
  const __half *w2_h;          // [64][256]
  const __half *wg_h;          // [2][96][64]: per group of 4 heads, rows = Wq(32) | Wk(32) | Wv(32)
  const float *bg;             // [2][96]
};

struct DevWeights {
  s2s_config cfg;
  const float *enc_pos, *dec_pos;     // [16][64], [250][64]
  const float *src_t, *src_b;         // [5k][64]
  const float *pre_t, *pre_b;         // [64][64]
  const float *smp0_t, *smp0_b;       // [64][192], [192]: first layers of conc | rate | noise MLPs
  const float *smp3_w, *smp3_b;       // [3][64], [3]
  const float *out_w, *out_b;         // [64], [1]
  BlockDev enc[4], dec[4];
};

// ---- k_frontend.cu -------------------------------------------------------------------------
// chunk -> (read, first base offset, valid k-mers) by binary search in chunk_offsets.
int launch_chunk_map(const int64_t* read_offsets, const int64_t* chunk_offsets, int64_t n_reads, int64_t n_chunks,
                     int32_t k, int32_t* chunk_read, int64_t* chunk_base, int32_t* chunk_nk, cudaStream_t st);
// Tokenise (bases or codes) + src_emb + ReLU + prenet + ReLU -> emb_out; x_enc = emb_out + pos.
// run_if (here and below): optional device flag; the launch is a no-op when *run_if == 0 (fallback launches of the
// k-mer table path, which only have work when a sub-batch contains letters outside "_ACGT").
int launch_embed(const DevWeights& w, const uint8_t* bases, const int64_t* chunk_base, const int32_t* chunk_nk,
                 const int8_t* codes, int64_t n_chunks, float* emb_out, float* x_enc, __half* x_enc16, cudaStream_t st,
                 const int* run_if = nullptr);

// Per-k-mer tables: emb_out, conc, rate and sigma are functions of ONE k-mer (modules.py:70-78,
// 216-219, 276), so they are computed once per checkpoint for all 4^k k-mers plus the "_"*k padding k-mer by the
// kernels above and looked up afterwards: bit-identical values, no per-row GEMMs in the front end.
struct KmerTables {
  const float* emb = nullptr;   // [4^k + 1][64]  emb_out; entry 4^k is the padding k-mer
  const float4* smp = nullptr;  // [4^k + 1]      (conc, rate, sigma, 0)
  int64_t n_kmers = 0;          // 4^k
  int* flag = nullptr;          // device int: set by a lookup launch that met a letter outside "_ACGT"
};
// all k-mers in table order as int8 letter codes [n][k] (A,C,G,T = 1..4, most significant letter first; row 4^k = zeros)
int launch_all_kmer_codes(int k, int64_t n_rows, int8_t* codes, cudaStream_t st);
// emb_out / x_enc / x_enc16 by lookup; kidx[row] = table index (-1: not in the table, *tab.flag set)
int launch_embed_lookup(const DevWeights& w, const KmerTables& tab, const uint8_t* bases, const int64_t* chunk_base,
                        const int32_t* chunk_nk, const int8_t* codes, int64_t n_chunks, float* emb_out, float* x_enc,
                        __half* x_enc16, int32_t* kidx, cudaStream_t st);

// ---- k_simt.cu (fp32 CUDA-core path) --------------------------------------------------------
enum { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_RES_LN = 2 };
// Y[M,N] = epi(X[M,K] @ Wt[K,N] + b); RES_LN: Y = LayerNorm(. + R) * g + beta (N must be 64)
int launch_linear_f32(const float* X, const float* Wt, const float* b, const float* R, const float* g,
                      const float* beta, float* Y, int64_t M, int K, int N, int epi, cudaStream_t st,
                      const int* run_if = nullptr);
// softmax(QK^T/sqrt(8))V per (chunk, head); qkv is [rows,192] (q|k|v), out [rows,64].
// L = 16 (rows_per_chunk 16) or 250 (rows_per_chunk 256: pad rows are neither keys nor written... they are zeroed)
int launch_attention_f32(const float* qkv, float* out, int64_t n_chunks, int L, int rows_per_chunk, cudaStream_t st);
// encoder attention (fp32 math) on fp16 q|k|v with fp16 output: the A operand of the tensor-core fc GEMM

// ---- k_samplers.cu ---------------------------------------------------------------------------
// h3 [M,192] = ReLU(first layers) already computed; this applies the 64->1 heads, Softplus, clamps,
// draws durations and rounds.  Outputs per k-mer: sigma, dur_int (+ optional conc, rate, dur_float taps).
int launch_sampler_heads(const DevWeights& w, const float* h3, int64_t n_kmers, const s2s_run_opts& o,
                         float* sigma, int32_t* dur_int, float* conc_tap, float* rate_tap, float* dur_float_tap,
                         cudaStream_t st, const int* run_if = nullptr);
// the same outputs from the per-k-mer table (rows with kidx < 0 are left to the fallback launch)
int launch_sampler_lookup(const KmerTables& tab, const int32_t* kidx, int64_t n_kmers, const s2s_run_opts& o,
                          float* sigma, int32_t* dur_int, float* conc_tap, float* rate_tap, float* dur_float_tap,
                          cudaStream_t st);
int launch_pack_smp_table(const float* conc, const float* rate, const float* sigma, int64_t n, float4* out, cudaStream_t st);

// ---- k_length_regulate.cu ---------------------------------------------------------------------
// x_dec[c, t, :] = (t < total ? enc_out[c, j(t), :] : 0) + dec_pos[t]  for t < 250, 0 for the 6 pad rows.
// dec_pos == nullptr gives the plain LR output with row stride `rows_per_chunk_out` (stage entry point).
int launch_length_regulate(const float* enc_out, const float* sigma, const int32_t* dur, int64_t n_chunks,
                           const float* dec_pos, float* x_dec, __half* x_dec16, int rows_per_chunk_out,
                           float* sigma_ext, int32_t* total, float* lr_tap, cudaStream_t st);

// ---- k_epilogue.cu ----------------------------------------------------------------------------
// p = ReLU(y . w_out + b); pA = clamp(165 p + noise, 0).  y has 256 rows per chunk, outputs 250 per chunk.
int launch_out_epilogue(const DevWeights& w, const float* y, const float* sigma_ext, int64_t n_chunks,
                        const s2s_run_opts& o, float* p_tap, float* pa, cudaStream_t st);
// Arguments of the fused output epilogue of the last tensor-core FFN kernel (model.py:221-240 inside k_tc_fc_ffn):
// pA = clamp(165 p + noise, 0) written once, 250 floats per chunk, plus the chunk's number of non-zero samples.
struct OutEpi {
  float* pa;                 // [chunks][250]
  float* p_tap;              // optional: p = ReLU(out_linear(.)) per position
  const float* sigma_ext;    // [chunks][250] expanded noise std
  int32_t* counts;           // optional: per-chunk non-zero count (zeroed by the caller), input of the compaction
  float scaling;             // scaling_max_value (165)
  s2s_run_opts o;            // noise mode / std / seed; chunk_id_base = global index of the sub-batch's first chunk
};
int launch_digitise(const float* pa, int64_t n, float dig, float range, float offset, int16_t* raw, cudaStream_t st);
int64_t compact_workspace_bytes(int64_t n_chunks);
// counts_ready: the per-chunk non-zero counts (compact_counts(ws)) were already written by the fused decoder epilogue
int launch_compact(const float* pa, const int64_t* chunk_offsets, int64_t n_reads, int64_t n_chunks, float dig,
                   float range, float offset, int rna_reverse, void* ws, int64_t ws_bytes, int16_t* raw,
                   int64_t* raw_offsets, cudaStream_t st, bool counts_ready = false);
inline int32_t* compact_counts(void* ws) { return static_cast<int32_t*>(ws); }   // first region of the workspace

}  // namespace s2s
