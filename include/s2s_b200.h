/*
 * s2s_b200 — C-ABI of the B200-native seq2squiggle predict hot path.
 *
 * Plain C, no torch / C++ types.  Every pointer named *_dev is a CUDA device pointer owned by the
 * caller (PyTorch allocates them and passes tensor.data_ptr()); kernels never allocate after
 * s2s_create().  All calls are asynchronous on the caller's stream unless stated otherwise and return
 * 0 on success, <0 on error (s2s_last_error() gives the message).  One handle per device; a handle is
 * not thread-safe, different handles are independent (one host thread / process per GPU).
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/seq2squiggle):
 *   s2s_forward_reads   : model.py:195-250 predict_step + model.py:253-302 export_and_clear_results
 *                         + signal_io.py:134-141 digitisation, fed by utils.py:350-356 split_sequence
 *   s2s_forward_chunks  : model.py:195-240 predict_step on a DataLoader batch (one-hot k-mer chunks)
 *   s2s_length_regulate : modules.py:344-392 LengthRegulator.LR
 *   s2s_digitise        : signal_io.py:134-141 / 247-254
 *   s2s_create          : inference.py:386-397 load_from_checkpoint (weights already unpacked by the host)
 */
#ifndef S2S_B200_H
#define S2S_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2S_ABI_VERSION 1

#define S2S_MAX_DNA_LEN 16     /* config.yaml: max_dna_len   */
#define S2S_MAX_SIGNAL_LEN 250 /* config.yaml: max_signal_len */
#define S2S_DMODEL 64
#define S2S_DFF 256
#define S2S_HEADS 8

typedef struct s2s_engine* s2s_handle;
typedef void* s2s_stream; /* cudaStream_t */

/* Architecture of the checkpoint (config.yaml:14-33).  Only the default architecture family is
 * compiled: dmodel 64, dff 256, 8 heads, max_dna_len 16, max_signal_len 250, pre_layers 1;
 * seq_kmer and the layer counts are free. */
typedef struct {
  int32_t seq_kmer;        /* 9 (dna-r10, rna-004) or 6 (dna-r9); src_emb is [64, 5*seq_kmer] */
  int32_t encoder_layers;  /* 1..4 */
  int32_t decoder_layers;  /* 1..4 */
  int32_t pre_layers;      /* must be 1 */
  int32_t dmodel, dff, heads, max_dna_len, max_signal_len; /* must be 64,256,8,16,250 */
  float scaling_max_value; /* 165.0 */
} s2s_config;

enum { S2S_DUR_CONSTANT = 0, S2S_DUR_NORMAL = 1, S2S_DUR_SAMPLER = 2 };   /* modules.py:410-432 */
enum { S2S_NOISE_OFF = 0, S2S_NOISE_STATIC = 1, S2S_NOISE_SAMPLER = 2 };  /* model.py:224-238   */
enum { S2S_PREC_FP16_TC = 0, S2S_PREC_FP32 = 1 };

typedef struct {
  int32_t duration_mode;   /* S2S_DUR_* : duration_sampling / dwell_std>0 / constant */
  float dwell_mean;        /* sample_rate / bps unless overridden (inference.py:358-359) */
  float dwell_std;
  float min_duration;      /* clamp in SAMPLER and NORMAL modes only (modules.py:414-432) */
  int32_t noise_mode;      /* S2S_NOISE_* ; OFF when noise_std <= 0 */
  float noise_std;
  float min_noise;
  float digitisation, range, offset_mean; /* signal_io.py:79-85 */
  int32_t rna_reverse;     /* profile starts with "rna": reverse each read's signal */
  uint64_t seed;           /* Philox key; draws are indexed by (global chunk id, position) */
  uint64_t chunk_id_base;  /* global id of this call's first chunk (shard-count invariance) */
  int32_t precision;       /* S2S_PREC_* */
} s2s_run_opts;

/* Optional per-stage taps (device pointers, may be NULL) used by the parity tests. */
typedef struct {
  float* emb_out_dev;     /* [C,16,64]  Encoder emb_out  (modules.py:70-78) */
  float* enc_out_dev;     /* [C,16,64]  Encoder output   (modules.py:80-89) */
  float* sigma_dev;       /* [C,16]     NoiseSampler     (modules.py:275-278) */
  float* conc_dev;        /* [C,16]     DurationSampler conc (modules.py:216-217) */
  float* rate_dev;        /* [C,16]     DurationSampler rate (modules.py:218-219) */
  float* dur_float_dev;   /* [C,16]     durations before rounding */
  int32_t* dur_int_dev;   /* [C,16]     rounded durations (modules.py:436-437) */
  float* lr_out_dev;      /* [C,250,64] length-regulated features (modules.py:366-388) */
  float* sigma_ext_dev;   /* [C,250]    expanded noise std */
  float* p_dev;           /* [C,250]    decoder output before x165 (modules.py:140-141) */
  float* pa_dev;          /* [C,250]    pA after noise and clamp (model.py:221-240) */
} s2s_taps;

const char* s2s_last_error(void);
int s2s_abi_version(void);

/* Number of fp32 values in the packed weight blob for `cfg` (layout: s2s_weights.h order, mirrored by
 * seq2squiggle_b200/checkpoint.py). */
int64_t s2s_weights_count(const s2s_config* cfg);

/* Copies the HOST weight blob to the device, builds the fp16 operand copies.  Synchronous. */
int s2s_create(const float* weights_host, int64_t n_weights, const s2s_config* cfg, int device, s2s_handle* out);
void s2s_destroy(s2s_handle h);

/* Bytes of device scratch s2s_forward_* needs for a call of n_chunks chunks / n_reads reads. */
int64_t s2s_workspace_bytes(s2s_handle h, int64_t n_chunks, int64_t n_reads);

/* Number of chunks a read of `read_len` bases produces: ceil((len-k+1)/16), 0 if len<k
 * (utils.py:334-356). */
int64_t s2s_chunks_of_read(int64_t read_len, int32_t seq_kmer);

/*
 * Reads -> digitised signal.
 *   bases_dev        : uint8 [read_offsets[n_reads]] ASCII bases of all reads, concatenated
 *   read_offsets_dev : int64 [n_reads+1] prefix offsets into bases_dev
 *   chunk_offsets_dev: int64 [n_reads+1] prefix sum of s2s_chunks_of_read() (host computes it)
 *   n_chunks         : chunk_offsets[n_reads]
 *   raw_out_dev      : int16 [n_chunks*250] capacity; read r occupies [raw_offsets[r], raw_offsets[r+1])
 *   raw_offsets_dev  : int64 [n_reads+1] (written)
 *   taps             : optional stage outputs, indexed by chunk in read order
 */
int s2s_forward_reads(s2s_handle h, const uint8_t* bases_dev, const int64_t* read_offsets_dev,
                      const int64_t* chunk_offsets_dev, int64_t n_reads, int64_t n_chunks,
                      const s2s_run_opts* opts, void* workspace_dev, int64_t workspace_bytes,
                      int16_t* raw_out_dev, int64_t* raw_offsets_dev, const s2s_taps* taps, s2s_stream stream);

/*
 * DataLoader-batch form of predict_step: k-mer letter codes (argmax of the one-hot; -1 = all-zero row)
 *   codes_dev : int8 [n_chunks,16,seq_kmer]
 *   pa_out_dev: float [n_chunks,250] pA after noise + clamp (what predict_step appends to results)
 */
int s2s_forward_chunks(s2s_handle h, const int8_t* codes_dev, int64_t n_chunks, const s2s_run_opts* opts,
                       void* workspace_dev, int64_t workspace_bytes, float* pa_out_dev, const s2s_taps* taps,
                       s2s_stream stream);

/* Synchronises `stream` and returns <0 if any kernel of this handle raised its device-side error word
 * (a bounded mbarrier wait that timed out) or a CUDA error is pending.  Call before trusting results. */
int s2s_check(s2s_handle h, s2s_stream stream);

/* Stage entry points (stage-isolated parity tests; same kernels the forward calls use). */
int s2s_length_regulate(const float* x_dev /*[C,16,64]*/, const float* sigma_dev /*[C,16]*/,
                        const int32_t* dur_dev /*[C,16]*/, int64_t n_chunks, float* out_dev /*[C,250,64]*/,
                        float* sigma_ext_dev /*[C,250]*/, int32_t* total_dev /*[C] min(sum,250)*/, s2s_stream stream);
int s2s_digitise(const float* pa_dev, int64_t n, float digitisation, float range, float offset_mean,
                 int16_t* raw_dev, s2s_stream stream);
/* Zero-strip + digitise + per-read compaction of dense pA rows (model.py:284-286, signal_io.py:134-141). */
int s2s_compact_reads(const float* pa_dev /*[C,250]*/, const int64_t* chunk_offsets_dev, int64_t n_reads,
                      int64_t n_chunks, float digitisation, float range, float offset_mean, int32_t rna_reverse,
                      void* workspace_dev, int64_t workspace_bytes, int16_t* raw_out_dev, int64_t* raw_offsets_dev,
                      s2s_stream stream);

/* Measurement hook (bench.py roofline leg): enable=1 starts bracketing every launch of the dominant kernel
 * (k_tc_attention) with CUDA events on the launching stream; enable=0 stops, synchronises and returns the
 * summed device time, the number of launches and the chunks they covered. */
int s2s_profile_kernel(s2s_handle h, int enable, double* ms_total, int64_t* launches, int64_t* chunks);
/* Between s2s_profile_kernel(h, 1, ..) and s2s_profile_kernel(h, 0, ..): the same three figures for one kernel group of
 * the path: "attention" (k_attn_gate + k_tc_attn3 + exact fallback), "ffn" (k_tc_fc_ffn4 incl. the fused output epilogue),
 * "length_regulate", "compact" (scan + k_compact), "encoder", "front_end" (tokeniser / table lookups).  Synchronises. */
int s2s_profile_kernel_group(s2s_handle h, const char* group, double* ms_total, int64_t* launches, int64_t* chunks);

/* Developer hook: per-phase clock64() sums of the attention kernel (thread 0 of every CTA); all zero unless the
 * library was built with -DS2S_PHASE_TIMING.  out[0..n): see csrc/k_tc.cu PHASE() indices.  Synchronises the device. */
int s2s_debug_counters(int64_t* out, int32_t n, int32_t reset);

/* Launch counter: number of kernels this library has launched since load (bench.py gpu_launches). */
int64_t s2s_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* S2S_B200_H */
