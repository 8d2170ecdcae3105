/*
 * s2s_blow5 — C-ABI of the native SLOW5/BLOW5 record writer (host code, no CUDA).
 *
 * Replaces the pyslow5 calls of the reference writer (paths relative to /root/reference/src/seq2squiggle):
 *   s2s_blow5_open        : signal_io.py:98-119  pyslow5.Open(filename, 'w'|'a') + get_empty_header + write_header
 *   s2s_blow5_write_batch : signal_io.py:143-171 get_empty_record(aux=True) per read + write_record_batch(threads=..)
 *   s2s_blow5_close       : signal_io.py:172     s5.close()  (writes the end-of-file marker)
 *
 * File layout follows the public SLOW5 specification v0.2.0 (hasindu2008/slow5specs): 64-byte binary file header
 * (magic "BLOW5\1", version, record compression, number of read groups, signal compression), uint32 size +
 * ASCII attribute / column header, length-prefixed records, "5WOLB" end marker.  Record compression: none or
 * zlib; signal compression: none or svb-zd (zigzag-delta + StreamVByte, slow5lib's default).  slow5lib / pyslow5 are
 * not available in the build image, so byte parity with pyslow5's output is unpinned; the format is pinned by an
 * independent reader (incl. an svb-zd decoder) in tests/blow5_reader.py.
 *
 * All pointers are HOST pointers owned by the caller.  Return 0 on success, <0 on error
 * (s2s_blow5_last_error()).  A handle is not thread-safe; write_batch itself fans out over n_threads.
 */
#ifndef S2S_BLOW5_H
#define S2S_BLOW5_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s2s_blow5_writer* s2s_blow5_handle;

enum { S2S_BLOW5_BINARY = 0, S2S_SLOW5_ASCII = 1 };            /* by file extension: .blow5 / .slow5 */
enum { S2S_BLOW5_COMPRESS_NONE = 0, S2S_BLOW5_COMPRESS_ZLIB = 1 };   /* record compression (file header byte 9) */
enum { S2S_BLOW5_SIGNAL_NONE = 0, S2S_BLOW5_SIGNAL_SVB_ZD = 1 };     /* signal compression (file header byte 14) */
/* Every `record_compression` argument below carries both: record method | (signal method << 8).  pyslow5's own default
 * (signal_io.py:98-102 opens with the defaults) is zlib records + svb-zd signal = S2S_BLOW5_COMPRESS_ZLIB |
 * (S2S_BLOW5_SIGNAL_SVB_ZD << 8). */

const char* s2s_blow5_last_error(void);

/* append != 0: the file must exist; its end marker is removed and records are appended (pyslow5 mode 'a').
 * header_attrs: the run's header attributes as "name\tvalue\n" lines (read group 0), ignored when appending. */
int s2s_blow5_open(const char* path, int format, int append, int record_compression, const char* header_attrs,
                   s2s_blow5_handle* out);

/* n_reads records.  read_ids: n_reads NUL-terminated strings back to back.  signal: int16 samples of all reads,
 * read r = signal[sig_offsets[r] .. sig_offsets[r+1]).  Per-read arrays: offset, median_before, read_number,
 * start_time.  Per-call scalars: digitisation, range, sampling_rate.  Fixed fields the reference writes:
 * read_group 0, channel_number "0", start_mux 0. */
int s2s_blow5_write_batch(s2s_blow5_handle h, int64_t n_reads, const char* read_ids, const int16_t* signal,
                          const int64_t* sig_offsets, const double* offset, const double* median_before,
                          const int32_t* read_number, const uint64_t* start_time, double digitisation, double range,
                          double sampling_rate, int32_t n_threads);

/* The same bytes WITHOUT a file: the file header, and the records of a batch (what s2s_blow5_write_batch appends), into a
 * malloc'ed buffer the caller releases with s2s_blow5_free.  Multi-GPU predict (one process per GPU) writes one shared
 * file this way: every rank encodes its batches and pwrite()s them at the offsets the ranks hand each other in read order,
 * so there are no part files to splice afterwards (inference.py; the reference has no sharded predict). */
int s2s_blow5_header(int format, int record_compression, const char* header_attrs, char** out, int64_t* out_bytes);
int s2s_blow5_encode_batch(int format, int record_compression, int64_t n_reads, const char* read_ids,
                           const int16_t* signal, const int64_t* sig_offsets, const double* offset,
                           const double* median_before, const int32_t* read_number, const uint64_t* start_time,
                           double digitisation, double range, double sampling_rate, int32_t n_threads, char** out,
                           int64_t* out_bytes);
void s2s_blow5_free(char* buf);

/* Bytes written so far (header + records). */
int64_t s2s_blow5_bytes_written(s2s_blow5_handle h);

/* Writes the end marker (binary) and closes the file.  The handle is freed. */
int s2s_blow5_close(s2s_blow5_handle h);

#ifdef __cplusplus
}
#endif
#endif /* S2S_BLOW5_H */
