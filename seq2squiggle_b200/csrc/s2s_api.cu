// C-ABI entry points (include/s2s_b200.h) and the host-side pipeline that strings the kernels together.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "s2s_kernels.h"
#include "s2s_tc.h"

namespace s2s {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace s2s

using namespace s2s;

struct s2s_engine {
  int device = 0;
  int sm_count = 148;
  s2s_config cfg{};
  float* d_f32 = nullptr;   // derived fp32 weights
  __half* d_f16 = nullptr;  // fp16 operand copies (tcgen05 path)
  DevWeights dw{};
  int64_t batch_chunks = 65536;  // chunks per sub-batch of the tensor-core path (S2S_BATCH_CHUNKS): 4.6 GB of workspace;
                                 // 32768 -> 65536 is +1.0 % (fewer kernel heads and tails), 131072 another +0.2 %
  int64_t batch_chunks_f32 = 1024;  // fp32 parity path: its fp32 scratch is 2.6 MB per chunk
  TcState tc{};
  // per-k-mer tables of the front end (S2S_KMER_TABLES=0 disables them): built by s2s_create for k <= 9
  KmerTables tab{};
  float* d_tab_emb = nullptr;
  float4* d_tab_smp = nullptr;
  int* d_tab_flag = nullptr;
};

// ------------------------------------------------------------------------------------------------
// derived weights
// ------------------------------------------------------------------------------------------------
namespace {

struct Packer {
  std::vector<float> f;
  std::vector<__half> h;
  int64_t put(const float* src, int64_t n) {  // returns offset (floats), 64-float aligned
    int64_t off = (int64_t)f.size();
    f.insert(f.end(), src, src + n);
    while (f.size() % 64) f.push_back(0.f);
    return off;
  }
  // transpose [rows][cols] -> [cols][rows]
  int64_t put_t(const float* src, int rows, int cols) {
    std::vector<float> t((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = src[(size_t)r * cols + c];
    return put(t.data(), (int64_t)rows * cols);
  }
  int64_t put_h(const float* src, int64_t n) {
    int64_t off = (int64_t)h.size();
    for (int64_t i = 0; i < n; ++i) h.push_back(__float2half_rn(src[i]));
    while (h.size() % 64) h.push_back(__float2half_rn(0.f));
    return off;
  }
};

struct BlockOff {
  int64_t wqkv_t, bqkv, fc_t, fc_b, w1_t, b1, w2_t, b2, ln1_w, ln1_b, ln2_w, ln2_b, wqkv_h, fc_h, w1_h, w2_h, wg_h, bg;
};

BlockOff pack_block(Packer& pk, const BlockW& b) {
  BlockOff o{};
  // concatenated [192][64] (q rows, then k, then v), transposed to [64][192]
  std::vector<float> cat(192 * 64), bias(192);
  memcpy(cat.data(), b.wq, 4096 * 4);
  memcpy(cat.data() + 4096, b.wk, 4096 * 4);
  memcpy(cat.data() + 8192, b.wv, 4096 * 4);
  memcpy(bias.data(), b.bq, 64 * 4);
  memcpy(bias.data() + 64, b.bk, 64 * 4);
  memcpy(bias.data() + 128, b.bv, 64 * 4);
  o.wqkv_t = pk.put_t(cat.data(), 192, 64);
  o.bqkv = pk.put(bias.data(), 192);
  o.fc_t = pk.put_t(b.fc_w, 64, 64);
  o.fc_b = pk.put(b.fc_b, 64);
  o.w1_t = pk.put_t(b.w1, 256, 64);
  o.b1 = pk.put(b.b1, 256);
  o.w2_t = pk.put_t(b.w2, 64, 256);
  o.b2 = pk.put(b.b2, 64);
  o.ln1_w = pk.put(b.ln1_w, 64);
  o.ln1_b = pk.put(b.ln1_b, 64);
  o.ln2_w = pk.put(b.ln2_w, 64);
  o.ln2_b = pk.put(b.ln2_b, 64);
  o.wqkv_h = pk.put_h(cat.data(), 192 * 64);
  o.fc_h = pk.put_h(b.fc_w, 64 * 64);
  o.w1_h = pk.put_h(b.w1, 256 * 64);
  o.w2_h = pk.put_h(b.w2, 64 * 256);
  // per head group g (4 heads): [96][64] = Wq rows 32g.., Wk rows 32g.., Wv rows 32g..  (+ bias [96])
  std::vector<float> wg(2 * 96 * 64), bgv(2 * 96);
  for (int g = 0; g < 2; ++g)
    for (int part = 0; part < 3; ++part) {
      memcpy(wg.data() + ((size_t)g * 96 + part * 32) * 64, cat.data() + ((size_t)part * 64 + g * 32) * 64, 32 * 64 * 4);
      memcpy(bgv.data() + g * 96 + part * 32, bias.data() + part * 64 + g * 32, 32 * 4);
    }
  o.wg_h = pk.put_h(wg.data(), 2 * 96 * 64);
  o.bg = pk.put(bgv.data(), 2 * 96);
  return o;
}

void bind_block(BlockDev& d, const BlockOff& o, const float* f, const __half* h) {
  d.wqkv_t = f + o.wqkv_t; d.bqkv = f + o.bqkv; d.fc_t = f + o.fc_t; d.fc_b = f + o.fc_b;
  d.w1_t = f + o.w1_t; d.b1 = f + o.b1; d.w2_t = f + o.w2_t; d.b2 = f + o.b2;
  d.ln1_w = f + o.ln1_w; d.ln1_b = f + o.ln1_b; d.ln2_w = f + o.ln2_w; d.ln2_b = f + o.ln2_b;
  d.wqkv_h = h + o.wqkv_h; d.fc_h = h + o.fc_h; d.w1_h = h + o.w1_h; d.w2_h = h + o.w2_h;
  d.wg_h = h + o.wg_h; d.bg = f + o.bg;
}

int check_config(const s2s_config* c) {
  if (!c) { set_error("null config"); return -1; }
  if (c->dmodel != 64 || c->dff != 256 || c->heads != 8 || c->max_dna_len != 16 || c->max_signal_len != 250 ||
      c->pre_layers != 1) {
    set_error("unsupported architecture: dmodel=%d dff=%d heads=%d max_dna_len=%d max_signal_len=%d pre_layers=%d "
              "(compiled for 64/256/8/16/250/1)", c->dmodel, c->dff, c->heads, c->max_dna_len, c->max_signal_len,
              c->pre_layers);
    return -1;
  }
  if (c->seq_kmer < 1 || c->seq_kmer > 12 || c->encoder_layers < 1 || c->encoder_layers > 4 ||
      c->decoder_layers < 1 || c->decoder_layers > 4) {
    set_error("unsupported seq_kmer=%d / encoder_layers=%d / decoder_layers=%d", c->seq_kmer, c->encoder_layers,
              c->decoder_layers);
    return -1;
  }
  return 0;
}

// ---- workspace carving ---------------------------------------------------------------------------
struct Carver {
  char* base;
  int64_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(int64_t n) {
    T* p = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += align_up(n * (int64_t)sizeof(T), 256);
    return p;
  }
};

struct Workspace {
  // whole call
  int32_t *chunk_read, *chunk_nk;
  int64_t* chunk_base;
  float* pa;
  void* compact_ws;
  int64_t compact_bytes;
  // per sub-batch
  float *emb, *xe, *qkv_e, *att_e, *ye, *he, *h3, *sigma;
  int32_t *dur, *total, *kidx;
  float *xd, *qkv_d, *att_d, *yd, *hd, *sigma_ext;
  TcBuffers tcb;
};

int64_t carve(Workspace& w, void* base, const s2s_engine* h, int64_t n_chunks, int64_t n_reads, bool need_pa) {
  Carver cv(base);
  const int64_t bc = n_chunks < h->batch_chunks ? n_chunks : h->batch_chunks;
  const int64_t bc32 = n_chunks < h->batch_chunks_f32 ? n_chunks : h->batch_chunks_f32;  // fp32-only buffers
  const int64_t me = bc * S2S_L_ENC, md = bc * S2S_L_DEC_PAD, md32 = bc32 * S2S_L_DEC_PAD;
  w.chunk_read = cv.take<int32_t>(n_chunks);
  w.chunk_nk = cv.take<int32_t>(n_chunks);
  w.chunk_base = cv.take<int64_t>(n_chunks);
  w.pa = need_pa ? cv.take<float>(n_chunks * S2S_L_DEC) : nullptr;
  w.compact_bytes = compact_workspace_bytes(n_chunks);
  w.compact_ws = cv.take<char>(w.compact_bytes);
  w.emb = cv.take<float>(me * 64);
  w.xe = cv.take<float>(align_up(me, 128) * 64);   // whole 128-row tiles for the tensor-core encoder
  w.qkv_e = cv.take<float>(align_up(me, 128) * 192);
  w.att_e = cv.take<float>(me * 64);
  w.ye = cv.take<float>(me * 64);
  w.he = cv.take<float>(me * 256);
  w.h3 = cv.take<float>(me * 192);
  w.sigma = cv.take<float>(me);
  w.dur = cv.take<int32_t>(me);
  w.kidx = cv.take<int32_t>(me);
  w.total = cv.take<int32_t>(bc);
  w.xd = cv.take<float>(md32 * 64);
  w.qkv_d = cv.take<float>(md32 * 192);
  w.att_d = cv.take<float>(md32 * 64);
  w.yd = cv.take<float>(md32 * 64);
  w.hd = cv.take<float>(md32 * 256);
  w.sigma_ext = cv.take<float>(bc * S2S_L_DEC);
  tc_carve(w.tcb, cv.base, cv.off, bc);
  (void)n_reads;
  return cv.off;
}

int tap_copy(void* dst, const void* src, int64_t bytes, cudaStream_t st) {
  if (dst && bytes > 0) S2S_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// One FFT block in fp32 on CUDA cores.  x (input and residual) -> x (output); y, qkv, att, hbuf scratch.
int fft_block_f32(const BlockDev& b, float* x, float* y, float* qkv, float* att, float* hbuf, int64_t n_chunks, int L,
                  int rows_per_chunk, cudaStream_t st) {
  const int64_t M = n_chunks * rows_per_chunk;
  if (launch_linear_f32(x, b.wqkv_t, b.bqkv, nullptr, nullptr, nullptr, qkv, M, 64, 192, EPI_BIAS, st)) return -1;
  if (launch_attention_f32(qkv, att, n_chunks, L, rows_per_chunk, st)) return -1;
  if (launch_linear_f32(att, b.fc_t, b.fc_b, x, b.ln1_w, b.ln1_b, y, M, 64, 64, EPI_BIAS_RES_LN, st)) return -1;
  if (launch_linear_f32(y, b.w1_t, b.b1, nullptr, nullptr, nullptr, hbuf, M, 64, 256, EPI_BIAS_RELU, st)) return -1;
  if (launch_linear_f32(hbuf, b.w2_t, b.b2, y, b.ln2_w, b.ln2_b, x, M, 256, 64, EPI_BIAS_RES_LN, st)) return -1;
  return 0;
}

// The whole path for chunks [0, n_chunks): writes pA rows into pa_out[n_chunks*250].
// counts (optional, tensor-core path): per-chunk non-zero sample counts for the compaction, written by the fused decoder
// epilogue; *counts_done tells the caller whether they were.
int run_pipeline(s2s_engine* h, const uint8_t* bases, const int8_t* codes, const Workspace& w, int64_t n_chunks,
                 const s2s_run_opts& opts, float* pa_out, const s2s_taps* taps, cudaStream_t st,
                 int32_t* counts = nullptr, bool* counts_done = nullptr) {
  const DevWeights& dw = h->dw;
  const int k = h->cfg.seq_kmer;
  const bool tc_path = opts.precision == S2S_PREC_FP16_TC;
  const int64_t step = tc_path ? h->batch_chunks : (h->batch_chunks_f32 < h->batch_chunks ? h->batch_chunks_f32 : h->batch_chunks);
  if (counts_done) *counts_done = tc_path && counts != nullptr;
  if (tc_path && counts && n_chunks > 0) S2S_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)n_chunks * sizeof(int32_t), st));
  for (int64_t c0 = 0; c0 < n_chunks; c0 += step) {
    const int64_t bc = (n_chunks - c0) < step ? (n_chunks - c0) : step;
    const int64_t me = bc * S2S_L_ENC;
    s2s_run_opts o = opts;
    o.chunk_id_base = opts.chunk_id_base + (uint64_t)c0;
    // K-A: by per-k-mer table when there is one; the direct kernels then only run (device-side flag) for a sub-batch
    // that contains a k-mer outside the table, and recompute that whole sub-batch with identical arithmetic
    const bool use_tab = h->tab.emb != nullptr;
    const int* run_if = use_tab ? h->tab.flag : nullptr;
    cudaEvent_t pe = prof_begin(h->tc, st);
    if (use_tab) {
      S2S_CUDA_OK(cudaMemsetAsync(h->tab.flag, 0, sizeof(int), st));
      if (launch_embed_lookup(dw, h->tab, bases, bases ? w.chunk_base + c0 : nullptr, bases ? w.chunk_nk + c0 : nullptr,
                              codes ? codes + c0 * S2S_L_ENC * k : nullptr, bc, w.emb, w.xe,
                              tc_path ? w.tcb.xe16 : nullptr, w.kidx, st)) return -1;
    }
    if (launch_embed(dw, bases, bases ? w.chunk_base + c0 : nullptr, bases ? w.chunk_nk + c0 : nullptr,
                     codes ? codes + c0 * S2S_L_ENC * k : nullptr, bc, w.emb, w.xe, tc_path ? w.tcb.xe16 : nullptr, st,
                     run_if))
      return -1;
    prof_end(h->tc, PROF_FRONT, pe, bc, st);
    pe = prof_begin(h->tc, st);
    // encoder (modules.py:82-87)
    if (tc_path) {
      if (tc_encoder(h->tc, dw, w.tcb, w.xe, w.tcb.xe16, w.tcb.oe16, bc, st)) return -1;
    } else {
      for (int l = 0; l < h->cfg.encoder_layers; ++l)
        if (fft_block_f32(dw.enc[l], w.xe, w.ye, w.qkv_e, w.att_e, w.he, bc, S2S_L_ENC, S2S_L_ENC, st)) return -1;
    }
    prof_end(h->tc, PROF_ENCODER, pe, bc, st);
    // samplers (K-C)
    float* conc_tap = taps && taps->conc_dev ? taps->conc_dev + c0 * 16 : nullptr;
    float* rate_tap = taps && taps->rate_dev ? taps->rate_dev + c0 * 16 : nullptr;
    float* durf_tap = taps && taps->dur_float_dev ? taps->dur_float_dev + c0 * 16 : nullptr;
    if (use_tab && launch_sampler_lookup(h->tab, w.kidx, me, o, w.sigma, w.dur, conc_tap, rate_tap, durf_tap, st)) return -1;
    if (launch_linear_f32(w.emb, dw.smp0_t, dw.smp0_b, nullptr, nullptr, nullptr, w.h3, me, 64, 192, EPI_BIAS_RELU, st,
                          run_if))
      return -1;
    if (launch_sampler_heads(dw, w.h3, me, o, w.sigma, w.dur, conc_tap, rate_tap, durf_tap, st, run_if)) return -1;
    // K-D
    // fp32 path: fp32 residual stream xd; tensor-core path: the fp16 stream x16 is the only copy
    pe = prof_begin(h->tc, st);
    if (launch_length_regulate(w.xe, w.sigma, w.dur, bc, dw.dec_pos, tc_path ? nullptr : w.xd, tc_path ? w.tcb.x16 : nullptr, S2S_L_DEC_PAD,
                               w.sigma_ext, w.total,
                               taps && taps->lr_out_dev ? taps->lr_out_dev + c0 * S2S_L_DEC * 64 : nullptr, st)) return -1;
    prof_end(h->tc, PROF_LR, pe, bc, st);
    if (taps) {
      if (tap_copy(taps->emb_out_dev ? taps->emb_out_dev + c0 * 16 * 64 : nullptr, w.emb, me * 64 * 4, st)) return -1;
      if (tap_copy(taps->enc_out_dev ? taps->enc_out_dev + c0 * 16 * 64 : nullptr, w.xe, me * 64 * 4, st)) return -1;
      if (tap_copy(taps->sigma_dev ? taps->sigma_dev + c0 * 16 : nullptr, w.sigma, me * 4, st)) return -1;
      if (tap_copy(taps->dur_int_dev ? taps->dur_int_dev + c0 * 16 : nullptr, w.dur, me * 4, st)) return -1;
      if (tap_copy(taps->sigma_ext_dev ? taps->sigma_ext_dev + c0 * S2S_L_DEC : nullptr, w.sigma_ext,
                   bc * S2S_L_DEC * 4, st)) return -1;
    }
    // decoder (modules.py:138-139)
    // decoder + K-F.  Tensor-core path: out_linear, x165, noise, clamp and the non-zero count live in the last FFN kernel.
    float* pa_b = pa_out + c0 * S2S_L_DEC;
    float* p_tap = taps && taps->p_dev ? taps->p_dev + c0 * S2S_L_DEC : nullptr;
    if (opts.precision == S2S_PREC_FP32) {
      for (int l = 0; l < h->cfg.decoder_layers; ++l)
        if (fft_block_f32(dw.dec[l], w.xd, w.yd, w.qkv_d, w.att_d, w.hd, bc, S2S_L_DEC, S2S_L_DEC_PAD, st)) return -1;
      if (launch_out_epilogue(dw, w.xd, w.sigma_ext, bc, o, p_tap, pa_b, st)) return -1;
    } else {
      OutEpi epi;
      epi.pa = pa_b; epi.p_tap = p_tap; epi.sigma_ext = w.sigma_ext; epi.counts = counts ? counts + c0 : nullptr;
      epi.scaling = dw.cfg.scaling_max_value; epi.o = o;
      if (tc_decoder(h->tc, dw, w.tcb, epi, bc, st)) return -1;
    }
    if (taps && tap_copy(taps->pa_dev ? taps->pa_dev + c0 * S2S_L_DEC : nullptr, pa_b, bc * S2S_L_DEC * 4, st)) return -1;
  }
  return 0;
}

// emb_out, conc, rate, sigma of every k-mer (and of the "_"*k padding k-mer), computed ONCE with the same kernels the
// direct path runs per row (k_embed on letter codes, the first sampler layers, the sampler heads), so a looked-up value
// is bit-identical to a recomputed one.  k = 9: 262,145 entries, 67 MB + 4 MB of HBM, ~1 ms to build.
int build_kmer_tables(s2s_engine* h) {
  const int k = h->cfg.seq_kmer;
  const int64_t nk = (int64_t)1 << (2 * k), rows = align_up(nk + 1, S2S_L_ENC), chunks = rows / S2S_L_ENC;
  int8_t* codes = nullptr;
  float *xe = nullptr, *h3 = nullptr, *vals = nullptr;
  int32_t* dur = nullptr;
  auto fail = [&](const char* what) {
    const std::string msg(what);  // `what` may be g_err itself
    set_error("s2s_create: k-mer tables: %s", msg.c_str());
    cudaFree(codes); cudaFree(xe); cudaFree(h3); cudaFree(vals); cudaFree(dur);
    return -1;
  };
  if (cudaMalloc(&h->d_tab_emb, rows * 64 * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&h->d_tab_smp, rows * sizeof(float4)) != cudaSuccess ||
      cudaMalloc(&h->d_tab_flag, sizeof(int)) != cudaSuccess || cudaMalloc(&codes, rows * k) != cudaSuccess ||
      cudaMalloc(&xe, rows * 64 * sizeof(float)) != cudaSuccess || cudaMalloc(&h3, rows * 192 * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&vals, rows * 3 * sizeof(float)) != cudaSuccess || cudaMalloc(&dur, rows * sizeof(int32_t)) != cudaSuccess)
    return fail("cudaMalloc failed");
  cudaStream_t st = nullptr;
  s2s_run_opts o{};
  o.duration_mode = S2S_DUR_CONSTANT;  // only conc / rate / sigma are kept; no draw is made
  o.dwell_mean = 1.f;
  if (launch_all_kmer_codes(k, rows, codes, st) ||
      launch_embed(h->dw, nullptr, nullptr, nullptr, codes, chunks, h->d_tab_emb, xe, nullptr, st) ||
      launch_linear_f32(h->d_tab_emb, h->dw.smp0_t, h->dw.smp0_b, nullptr, nullptr, nullptr, h3, rows, 64, 192,
                        EPI_BIAS_RELU, st) ||
      launch_sampler_heads(h->dw, h3, rows, o, vals + 2 * rows, dur, vals, vals + rows, nullptr, st) ||
      launch_pack_smp_table(vals, vals + rows, vals + 2 * rows, rows, h->d_tab_smp, st))
    return fail(g_err);
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("kernel failure");
  cudaMemset(h->d_tab_flag, 0, sizeof(int));
  cudaFree(codes); cudaFree(xe); cudaFree(h3); cudaFree(vals); cudaFree(dur);
  h->tab.emb = h->d_tab_emb;
  h->tab.smp = h->d_tab_smp;
  h->tab.n_kmers = nk;
  h->tab.flag = h->d_tab_flag;
  return 0;
}

int check_opts(const s2s_run_opts* o) {
  if (!o) { set_error("null run options"); return -1; }
  if (o->duration_mode < 0 || o->duration_mode > 2 || o->noise_mode < 0 || o->noise_mode > 2 ||
      (o->precision != S2S_PREC_FP16_TC && o->precision != S2S_PREC_FP32)) {
    set_error("bad run options: duration_mode=%d noise_mode=%d precision=%d", o->duration_mode, o->noise_mode, o->precision);
    return -1;
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* s2s_last_error(void) { return g_err; }
int s2s_abi_version(void) { return S2S_ABI_VERSION; }
int64_t s2s_launch_count(void) { return g_launch_count; }

int64_t s2s_weights_count(const s2s_config* cfg) {
  if (check_config(cfg)) return -1;
  return weights_count(*cfg);
}

int64_t s2s_chunks_of_read(int64_t read_len, int32_t seq_kmer) {
  int64_t n = read_len - seq_kmer + 1;
  return n <= 0 ? 0 : (n + S2S_L_ENC - 1) / S2S_L_ENC;
}

int s2s_create(const float* weights_host, int64_t n_weights, const s2s_config* cfg, int device, s2s_handle* out) {
  if (!out) { set_error("null out handle"); return -1; }
  *out = nullptr;
  if (check_config(cfg)) return -1;
  if (!weights_host || n_weights != weights_count(*cfg)) {
    set_error("weight blob has %lld floats, expected %lld", (long long)n_weights, (long long)weights_count(*cfg));
    return -1;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_error("no CUDA device available (%s): the s2s_b200 path has no CPU fallback", cudaGetErrorString(e));
    return -2;
  }
  S2S_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  S2S_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return -2;
  }
  s2s_engine* h = new s2s_engine();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cfg = *cfg;
  if (const char* env = getenv("S2S_BATCH_CHUNKS")) {
    long v = atol(env);
    if (v >= 1) h->batch_chunks = v;
  }
  const Weights W = map_weights(weights_host, *cfg);
  Packer pk;
  const int k5 = 5 * cfg->seq_kmer;
  int64_t o_enc_pos = pk.put(W.enc_pos, 16 * 64), o_dec_pos = pk.put(W.dec_pos, 250 * 64);
  int64_t o_src_t = pk.put_t(W.src_w, 64, k5), o_src_b = pk.put(W.src_b, 64);
  int64_t o_pre_t = pk.put_t(W.pre_w, 64, 64), o_pre_b = pk.put(W.pre_b, 64);
  std::vector<float> s0(192 * 64), s0b(192), s3(3 * 64), s3b(3);
  const MlpW* mlps[3] = {&W.conc, &W.rate, &W.noise};
  for (int m = 0; m < 3; ++m) {
    memcpy(s0.data() + m * 4096, mlps[m]->w0, 4096 * 4);
    memcpy(s0b.data() + m * 64, mlps[m]->b0, 64 * 4);
    memcpy(s3.data() + m * 64, mlps[m]->w3, 64 * 4);
    s3b[m] = mlps[m]->b3[0];
  }
  int64_t o_s0t = pk.put_t(s0.data(), 192, 64), o_s0b = pk.put(s0b.data(), 192);
  int64_t o_s3 = pk.put(s3.data(), 192), o_s3b = pk.put(s3b.data(), 3);
  int64_t o_out_w = pk.put(W.out_w, 64), o_out_b = pk.put(W.out_b, 1);
  BlockOff eo[4], dofs[4];
  for (int i = 0; i < cfg->encoder_layers; ++i) eo[i] = pack_block(pk, W.enc[i]);
  for (int i = 0; i < cfg->decoder_layers; ++i) dofs[i] = pack_block(pk, W.dec[i]);

  if (cudaMalloc(&h->d_f32, pk.f.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&h->d_f16, pk.h.size() * sizeof(__half)) != cudaSuccess) {
    set_error("cudaMalloc of the weight buffers failed");
    s2s_destroy(h);
    return -1;
  }
  cudaMemcpy(h->d_f32, pk.f.data(), pk.f.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_f16, pk.h.data(), pk.h.size() * sizeof(__half), cudaMemcpyHostToDevice);
  const float* f = h->d_f32;
  DevWeights& d = h->dw;
  d.cfg = *cfg;
  d.enc_pos = f + o_enc_pos; d.dec_pos = f + o_dec_pos;
  d.src_t = f + o_src_t; d.src_b = f + o_src_b; d.pre_t = f + o_pre_t; d.pre_b = f + o_pre_b;
  d.smp0_t = f + o_s0t; d.smp0_b = f + o_s0b; d.smp3_w = f + o_s3; d.smp3_b = f + o_s3b;
  d.out_w = f + o_out_w; d.out_b = f + o_out_b;
  for (int i = 0; i < cfg->encoder_layers; ++i) bind_block(d.enc[i], eo[i], f, h->d_f16);
  for (int i = 0; i < cfg->decoder_layers; ++i) bind_block(d.dec[i], dofs[i], f, h->d_f16);
  auto fill_ffn = [&](FfnParams& p, const BlockW& b) {   // host copy of the per-column vectors (kernel parameter)
    memcpy(p.bfc, b.fc_b, 64 * 4); memcpy(p.g1, b.ln1_w, 64 * 4); memcpy(p.be1, b.ln1_b, 64 * 4);
    memcpy(p.b1, b.b1, 256 * 4); memcpy(p.b2, b.b2, 64 * 4); memcpy(p.g2, b.ln2_w, 64 * 4); memcpy(p.be2, b.ln2_b, 64 * 4);
    memcpy(p.wout, W.out_w, 64 * 4);
    p.bout = W.out_b[0];
  };
  for (int i = 0; i < cfg->encoder_layers; ++i) fill_ffn(d.enc[i].ffn, W.enc[i]);
  for (int i = 0; i < cfg->decoder_layers; ++i) fill_ffn(d.dec[i].ffn, W.dec[i]);
  if (tc_init(h->tc, h->dw, device)) {
    s2s_destroy(h);
    return -1;
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("s2s_create: %s", cudaGetErrorString(e));
    s2s_destroy(h);
    return -1;
  }
  const char* tab_env = getenv("S2S_KMER_TABLES");
  if (cfg->seq_kmer <= 9 && !(tab_env && atoi(tab_env) == 0) && build_kmer_tables(h)) {
    s2s_destroy(h);
    return -1;
  }
  *out = h;
  return 0;
}

void s2s_destroy(s2s_handle h) {
  if (!h) return;
  tc_destroy(h->tc);
  if (h->d_f32) cudaFree(h->d_f32);
  if (h->d_f16) cudaFree(h->d_f16);
  if (h->d_tab_emb) cudaFree(h->d_tab_emb);
  if (h->d_tab_smp) cudaFree(h->d_tab_smp);
  if (h->d_tab_flag) cudaFree(h->d_tab_flag);
  delete h;
}

int64_t s2s_workspace_bytes(s2s_handle h, int64_t n_chunks, int64_t n_reads) {
  if (!h || n_chunks < 0) return -1;
  Workspace w;
  return carve(w, nullptr, h, n_chunks, n_reads, true) + 256;
}

int s2s_forward_reads(s2s_handle h, const uint8_t* bases_dev, const int64_t* read_offsets_dev,
                      const int64_t* chunk_offsets_dev, int64_t n_reads, int64_t n_chunks, const s2s_run_opts* opts,
                      void* workspace_dev, int64_t workspace_bytes, int16_t* raw_out_dev, int64_t* raw_offsets_dev,
                      const s2s_taps* taps, s2s_stream stream) {
  if (!h) { set_error("null handle"); return -1; }
  if (check_opts(opts)) return -1;
  if (n_reads < 0 || n_chunks < 0) { set_error("negative sizes"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace w;
  // 256-byte align the caller's pointer
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<int64_t>(workspace_dev), 256));
  int64_t need = carve(w, base, h, n_chunks, n_reads, true);
  if (!workspace_dev || need + (base - static_cast<char*>(workspace_dev)) > workspace_bytes) {
    set_error("workspace too small: need %lld bytes", (long long)(need + 256));
    return -1;
  }
  if (launch_chunk_map(read_offsets_dev, chunk_offsets_dev, n_reads, n_chunks, h->cfg.seq_kmer, w.chunk_read,
                       w.chunk_base, w.chunk_nk, st)) return -1;
  bool counted = false;
  if (run_pipeline(h, bases_dev, nullptr, w, n_chunks, *opts, w.pa, taps, st, compact_counts(w.compact_ws), &counted)) return -1;
  cudaEvent_t pe = prof_begin(h->tc, st);
  const int rc = launch_compact(w.pa, chunk_offsets_dev, n_reads, n_chunks, opts->digitisation, opts->range,
                                opts->offset_mean, opts->rna_reverse, w.compact_ws, w.compact_bytes, raw_out_dev,
                                raw_offsets_dev, st, counted);
  prof_end(h->tc, PROF_COMPACT, pe, n_chunks, st);
  return rc;
}

int s2s_forward_chunks(s2s_handle h, const int8_t* codes_dev, int64_t n_chunks, const s2s_run_opts* opts,
                       void* workspace_dev, int64_t workspace_bytes, float* pa_out_dev, const s2s_taps* taps,
                       s2s_stream stream) {
  if (!h) { set_error("null handle"); return -1; }
  if (check_opts(opts)) return -1;
  if (n_chunks == 0) return 0;  // an empty DataLoader batch is a no-op (its tensors have no storage to point to)
  if (n_chunks < 0 || !codes_dev || !pa_out_dev) { set_error("bad arguments"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace w;
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<int64_t>(workspace_dev), 256));
  int64_t need = carve(w, base, h, n_chunks, 0, false);
  if (!workspace_dev || need + (base - static_cast<char*>(workspace_dev)) > workspace_bytes) {
    set_error("workspace too small: need %lld bytes", (long long)(need + 256));
    return -1;
  }
  return run_pipeline(h, nullptr, codes_dev, w, n_chunks, *opts, pa_out_dev, taps, st);
}

int s2s_check(s2s_handle h, s2s_stream stream) {
  if (!h) { set_error("null handle"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tc_check_status(h->tc, st)) return -1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("CUDA error: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int s2s_profile_kernel(s2s_handle h, int enable, double* ms_total, int64_t* launches, int64_t* chunks) {
  if (!h) { set_error("null handle"); return -1; }
  return tc_profile(h->tc, enable, ms_total, launches, chunks);
}

int s2s_profile_kernel_group(s2s_handle h, const char* group, double* ms_total, int64_t* launches, int64_t* chunks) {
  if (!h || !group) { set_error("null argument"); return -1; }
  static const char* names[PROF_KINDS] = {"attention", "ffn", "length_regulate", "compact", "encoder", "front_end"};
  for (int k = 0; k < PROF_KINDS; ++k)
    if (!strcmp(group, names[k])) return tc_profile_kind(h->tc, k, ms_total, launches, chunks);
  set_error("unknown kernel group '%s'", group);
  return -1;
}

int s2s_debug_counters(int64_t* out, int32_t n, int32_t reset) {
  if (!out || n < 0) { set_error("bad arguments"); return -1; }
  return tc_debug_counters(out, n, reset);
}

int s2s_length_regulate(const float* x_dev, const float* sigma_dev, const int32_t* dur_dev, int64_t n_chunks,
                        float* out_dev, float* sigma_ext_dev, int32_t* total_dev, s2s_stream stream) {
  return launch_length_regulate(x_dev, sigma_dev, dur_dev, n_chunks, nullptr, out_dev, nullptr, S2S_L_DEC, sigma_ext_dev,
                                total_dev, nullptr, static_cast<cudaStream_t>(stream));
}

int s2s_digitise(const float* pa_dev, int64_t n, float digitisation, float range, float offset_mean, int16_t* raw_dev,
                 s2s_stream stream) {
  return launch_digitise(pa_dev, n, digitisation, range, offset_mean, raw_dev, static_cast<cudaStream_t>(stream));
}

int s2s_compact_reads(const float* pa_dev, const int64_t* chunk_offsets_dev, int64_t n_reads, int64_t n_chunks,
                      float digitisation, float range, float offset_mean, int32_t rna_reverse, void* workspace_dev,
                      int64_t workspace_bytes, int16_t* raw_out_dev, int64_t* raw_offsets_dev, s2s_stream stream) {
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<int64_t>(workspace_dev), 256));
  return launch_compact(pa_dev, chunk_offsets_dev, n_reads, n_chunks, digitisation, range, offset_mean, rna_reverse,
                        base, workspace_bytes - (base - static_cast<char*>(workspace_dev)), raw_out_dev,
                        raw_offsets_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
