// tcgen05 / TMEM / TMA decoder path (S2S_PREC_FP16_TC): declarations shared with s2s_api.cu.
#pragma once
#include <vector>

#include "s2s_kernels.h"

namespace s2s {

struct TcState {
  int device = 0;
  int sm_count = 148;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled entry point
  int32_t* d_status = nullptr;   // device word set by a kernel whose mbarrier wait timed out
  // optional CUDA-event bracketing of the attention launches (bench.py roofline leg)
  bool attn_exact = false;       // S2S_ATTN_EXACT=1: the exact two-pass attention kernel on its own (tests)
  // optional CUDA-event bracketing of kernel groups on the launching stream (bench.py roofline leg)
  bool prof_on = false;
  struct ProfRec { int kind; cudaEvent_t e0, e1; int64_t chunks; };
  std::vector<ProfRec> prof_recs;
  int64_t attn_calls = 0;        // decoder calls so far: every 16th ignores the "attention too sharp" hints (re-probe)
};

// Per-sub-batch device buffers of the tensor-core path (carved from the caller's workspace).
struct TcBuffers {
  __half* x16 = nullptr;    // [rows,64]  fp16 copy of the residual stream (GEMM A operand)
  __half* o16 = nullptr;    // [rows,64]  attention output (A operand of fc)
  __half* xe16 = nullptr;   // encoder: [chunks*16 (padded to 128), 64] fp16 residual-stream copy
  __half* oe16 = nullptr;   // encoder attention output
  int32_t* flags = nullptr; // [0] = number of flagged units, [1..2*chunks] = per-unit overflow flags of k_tc_attn3
};

void tc_carve(TcBuffers& b, char* base, int64_t& off, int64_t batch_chunks);
int tc_init(TcState& s, const DevWeights& w, int device);
void tc_destroy(TcState& s);
// Runs all decoder layers in place on x32 ([chunks*256,64] fp32 residual stream).
enum ProfKind { PROF_ATTN = 0, PROF_FFN = 1, PROF_LR = 2, PROF_COMPACT = 3, PROF_ENCODER = 4, PROF_FRONT = 5, PROF_KINDS = 6 };
// prof_begin/prof_end bracket a group of launches (no-ops unless profiling is enabled)
cudaEvent_t prof_begin(TcState& s, cudaStream_t st);
void prof_end(TcState& s, int kind, cudaEvent_t e0, int64_t chunks, cudaStream_t st);
int tc_profile(TcState& s, int enable, double* ms_total, int64_t* launches, int64_t* chunks);
// after tc_profile(enable = 1) ... run ...: per-kind sums (synchronises); tc_profile(enable = 0) releases the events
int tc_profile_kind(TcState& s, int kind, double* ms_total, int64_t* launches, int64_t* chunks);
// Phase-timing counters of k_tc_attn (all zero unless built with -DS2S_PHASE_TIMING).
int tc_debug_counters(int64_t* out, int n, int reset);
// Synchronises the stream and reports a device-side barrier timeout, if any.
int tc_check_status(TcState& s, cudaStream_t st);
// Runs all decoder layers on the fp16 residual stream b.x16 ([chunks*256,64], written by the length regulator) and
// writes pA = clamp(165 ReLU(out_linear(.)) + noise, 0) (and the per-chunk non-zero counts) through the fused epilogue.
int tc_decoder(TcState& s, const DevWeights& w, const TcBuffers& b, const OutEpi& epi, int64_t n_chunks, cudaStream_t st);
// Runs all encoder layers in place on x32/x16 ([chunks*16 rows, padded to 128]); qkv32 [rows,192], o16 [rows,64] scratch.
int tc_encoder(TcState& s, const DevWeights& w, const TcBuffers& b, float* x32, __half* x16, __half* o16,
               int64_t n_chunks, cudaStream_t st);

}  // namespace s2s
