// Packed fp32 weight blob: order and sizes.  Mirrored by seq2squiggle_b200/checkpoint.py:pack_weights
// (state_dict keys of the reference model, model.py:47-50 / SURVEY §5 "Checkpoint").
#pragma once
#include <stdint.h>

#include "../../include/s2s_b200.h"

namespace s2s {

// One FFT block (layers.py:116-142): offsets in floats relative to the block start.
struct BlockW {
  const float *wq, *bq, *wk, *bk, *wv, *bv;  // [64,64],[64] each (w_qs, w_ks, w_vs)
  const float *ln1_w, *ln1_b;                // slf_attn.layer_norm
  const float *fc_w, *fc_b;                  // [64,64],[64]
  const float *w1, *b1;                      // pos_ffn.w_1 [256,64],[256]
  const float *w2, *b2;                      // pos_ffn.w_2 [64,256],[64]
  const float *ln2_w, *ln2_b;                // pos_ffn.layer_norm
};
constexpr int64_t kBlockFloats = 3 * (64 * 64 + 64) + 128 + (64 * 64 + 64) + (256 * 64 + 256) + (64 * 256 + 64) + 128;

struct MlpW {  // Linear(64,64) - ReLU - Linear(64,1) - Softplus (modules.py:180-193, 266-272)
  const float *w0, *b0, *w3, *b3;
};
constexpr int64_t kMlpFloats = 64 * 64 + 64 + 64 + 1;

struct Weights {
  const float* enc_pos;           // [16,64]
  const float *src_w, *src_b;     // [64,5k],[64]
  const float *pre_w, *pre_b;     // [64,64],[64]
  BlockW enc[4];
  MlpW conc, rate;
  const float* dec_pos;           // [250,64]
  const float *out_w, *out_b;     // [1,64],[1]
  BlockW dec[4];
  MlpW noise;
};

inline int64_t weights_count(const s2s_config& c) {
  return 16 * 64 + (64 * 5 * c.seq_kmer + 64) + (64 * 64 + 64) + c.encoder_layers * kBlockFloats + 2 * kMlpFloats +
         250 * 64 + (64 + 1) + c.decoder_layers * kBlockFloats + kMlpFloats;
}

inline const float* take(const float*& p, int64_t n) {
  const float* r = p;
  p += n;
  return r;
}
inline void map_block(const float*& p, BlockW& b) {
  b.wq = take(p, 4096); b.bq = take(p, 64);
  b.wk = take(p, 4096); b.bk = take(p, 64);
  b.wv = take(p, 4096); b.bv = take(p, 64);
  b.ln1_w = take(p, 64); b.ln1_b = take(p, 64);
  b.fc_w = take(p, 4096); b.fc_b = take(p, 64);
  b.w1 = take(p, 256 * 64); b.b1 = take(p, 256);
  b.w2 = take(p, 64 * 256); b.b2 = take(p, 64);
  b.ln2_w = take(p, 64); b.ln2_b = take(p, 64);
}
inline void map_mlp(const float*& p, MlpW& m) {
  m.w0 = take(p, 4096); m.b0 = take(p, 64); m.w3 = take(p, 64); m.b3 = take(p, 1);
}
inline Weights map_weights(const float* base, const s2s_config& c) {
  Weights w{};
  const float* p = base;
  w.enc_pos = take(p, 16 * 64);
  w.src_w = take(p, 64 * 5 * c.seq_kmer); w.src_b = take(p, 64);
  w.pre_w = take(p, 4096); w.pre_b = take(p, 64);
  for (int i = 0; i < c.encoder_layers; ++i) map_block(p, w.enc[i]);
  map_mlp(p, w.conc);
  map_mlp(p, w.rate);
  w.dec_pos = take(p, 250 * 64);
  w.out_w = take(p, 64); w.out_b = take(p, 1);
  for (int i = 0; i < c.decoder_layers; ++i) map_block(p, w.dec[i]);
  map_mlp(p, w.noise);
  return w;
}

}  // namespace s2s
