// Native SLOW5 / BLOW5 record writer (include/s2s_blow5.h).  Host code: g++ -O3 -shared -fPIC -lz -pthread.
//
// Layout (SLOW5 specification v0.2.0):
//   file header  "BLOW5\1" | major minor patch (u8 x3) | record compression (u8) | num_read_groups (u32) |
//                signal compression (u8) | zero padding up to byte 64 | header size (u32) | ASCII header
//   ASCII header "@attr\tvalue\n" per attribute (sorted by name), then the column type line and the column name line
//   record       record_size (u64) | body.  body (zlib-deflated as a whole when record compression is zlib):
//                read_id_len (u16) read_id | read_group (u32) | digitisation offset range sampling_rate (f64 x4) |
//                len_raw_signal (u64) | raw_signal (i16 x len) | aux: channel_number (u64 len + chars) |
//                median_before (f64) | read_number (i32) | start_mux (u8) | start_time (u64)
//   end marker   "5WOLB"
// Records of a batch are laid out (or deflated) by a pool of threads into one buffer and written with a single
// fwrite, so the caller's compute stream never waits on per-record I/O.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/s2s_blow5.h"

namespace {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char kMagic[6] = {'B', 'L', 'O', 'W', '5', '\1'};
const char kEof[5] = {'5', 'W', 'O', 'L', 'B'};
const uint8_t kVersion[3] = {0, 2, 0};
const char* kTypes = "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\tchar*\tdouble\tint32_t\tuint8_t\tuint64_t\n";
const char* kNames = "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\t"
                     "channel_number\tmedian_before\tread_number\tstart_mux\tstart_time\n";

template <typename T>
inline void put(char*& p, T v) {
  memcpy(p, &v, sizeof(T));
  p += sizeof(T);
}

// shortest decimal that round-trips a double (what a text reader needs)
std::string fmt_double(double v) {
  char buf[40];
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, v);
    if (strtod(buf, nullptr) == v) break;
  }
  return buf;
}

// ---- svb-zd: the signal compression pyslow5 / slow5lib apply by default (slow5lib src/slow5_press.c, "svb-zd") ----------
// zigzag-delta of the int16 samples widened to int32 (d_i = x_i - x_{i-1}, x_{-1} = 0, z = (d << 1) ^ (d >> 31)), then
// StreamVByte (Lemire & Kurz) of the uint32 stream: ceil(n / 4) control bytes (2 bits per value: bytes - 1, value i of a
// group in bits 2 (i % 4)), then the values' low 1..4 bytes, little endian.  The compressed blob is
//   uint32 n | control bytes | data bytes,
// and a record stores it as  uint64 blob_bytes | blob  in place of the raw samples (len_raw_signal still counts samples).
// slow5lib is not available offline: written from its published algorithm; byte parity with slow5lib is unpinned.
size_t svb_zd_bound(size_t n) { return 4 + (n + 3) / 4 + 4 * n; }

size_t svb_zd_encode(const int16_t* x, size_t n, unsigned char* out) {
  const uint32_t n32 = (uint32_t)n;
  memcpy(out, &n32, 4);
  unsigned char* ctrl = out + 4;
  unsigned char* data = ctrl + (n + 3) / 4;
  int32_t prev = 0;
  unsigned char key = 0;
  for (size_t i = 0; i < n; ++i) {
    const int32_t d = (int32_t)x[i] - prev;
    prev = x[i];
    const uint32_t z = ((uint32_t)d << 1) ^ (uint32_t)(d >> 31);
    unsigned code;
    if (z < (1u << 8)) { code = 0; data[0] = (unsigned char)z; data += 1; }
    else if (z < (1u << 16)) { code = 1; data[0] = (unsigned char)z; data[1] = (unsigned char)(z >> 8); data += 2; }
    else if (z < (1u << 24)) { code = 2; data[0] = (unsigned char)z; data[1] = (unsigned char)(z >> 8); data[2] = (unsigned char)(z >> 16); data += 3; }
    else { code = 3; memcpy(data, &z, 4); data += 4; }
    key |= (unsigned char)(code << (2 * (i & 3)));
    if ((i & 3) == 3) { *ctrl++ = key; key = 0; }
  }
  if (n & 3) *ctrl++ = key;
  return (size_t)(data - out);
}

std::string sorted_attrs(const char* header_attrs) {
  std::vector<std::pair<std::string, std::string>> kv;
  const char* p = header_attrs ? header_attrs : "";
  while (*p) {
    const char* nl = strchr(p, '\n');
    std::string line = nl ? std::string(p, nl) : std::string(p);
    p = nl ? nl + 1 : p + line.size();
    size_t tab = line.find('\t');
    if (line.empty() || tab == std::string::npos) continue;
    kv.emplace_back(line.substr(0, tab), line.substr(tab + 1));
  }
  std::sort(kv.begin(), kv.end());
  std::string out;
  for (auto& e : kv) out += "@" + e.first + "\t" + e.second + "\n";
  return out;
}

}  // namespace

struct s2s_blow5_writer {
  FILE* fp = nullptr;
  int format = S2S_BLOW5_BINARY;
  int compression = S2S_BLOW5_COMPRESS_NONE;
  int64_t bytes = 0;
};

namespace {

std::string header_bytes(int format, int compression, const char* header_attrs) {
  const std::string attrs = sorted_attrs(header_attrs);
  std::string hdr;
  if (format == S2S_BLOW5_BINARY) {
    const std::string ascii = attrs + kTypes + kNames;
    hdr.assign(64, '\0');
    memcpy(&hdr[0], kMagic, 6);
    memcpy(&hdr[6], kVersion, 3);
    hdr[9] = (char)(compression & 0xFF);          // record compression: 0 none, 1 zlib
    const uint32_t n_groups = 1;
    memcpy(&hdr[10], &n_groups, 4);
    hdr[14] = (char)((compression >> 8) & 0xFF);  // signal compression: 0 none, 1 svb-zd
    const uint32_t hsize = (uint32_t)ascii.size();
    hdr.append(reinterpret_cast<const char*>(&hsize), 4);
    hdr += ascii;
  } else {
    hdr = "#slow5_version\t0.2.0\n#num_read_groups\t1\n" + attrs + kTypes + kNames;
  }
  return hdr;
}

// The records of a batch, laid out (or deflated) by a pool of threads: one string per thread, in record order.
int encode_records(int format, int compression, int64_t n_reads, const char* read_ids, const int16_t* signal,
                   const int64_t* sig_offsets, const double* offset, const double* median_before,
                   const int32_t* read_number, const uint64_t* start_time, double digitisation, double range,
                   double sampling_rate, int32_t n_threads, std::vector<std::string>& chunks);

}  // namespace

namespace {

int encode_records(int format, int compression, int64_t n_reads, const char* read_ids, const int16_t* signal,
                   const int64_t* sig_offsets, const double* offset, const double* median_before,
                   const int32_t* read_number, const uint64_t* start_time, double digitisation, double range,
                   double sampling_rate, int32_t n_threads, std::vector<std::string>& chunks) {
  if (!read_ids || !sig_offsets || !offset || !median_before || !read_number || !start_time || (!signal && sig_offsets[n_reads] > 0)) {
    set_error("null argument");
    return -1;
  }
  std::vector<const char*> ids((size_t)n_reads);
  std::vector<uint32_t> id_len((size_t)n_reads);
  const char* p = read_ids;
  for (int64_t r = 0; r < n_reads; ++r) {
    ids[r] = p;
    size_t l = strlen(p);
    if (l > 65535) { set_error("read id longer than 65535 bytes"); return -1; }
    id_len[r] = (uint32_t)l;
    p += l + 1;
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if ((int64_t)n_threads > n_reads) n_threads = (int32_t)n_reads;

  chunks.assign((size_t)n_threads, std::string());
  const int rec_comp = compression & 0xFF, sig_comp = (compression >> 8) & 0xFF;
  // upper bound of a record body (exact when the signal is stored raw)
  auto body_size = [&](int64_t r) -> size_t {
    const size_t n = (size_t)(sig_offsets[r + 1] - sig_offsets[r]);
    const size_t sig_bytes = sig_comp == S2S_BLOW5_SIGNAL_SVB_ZD ? 8 + svb_zd_bound(n) : 2 * n;
    return 2 + id_len[r] + 4 + 32 + 8 + sig_bytes + (8 + 1) + 8 + 4 + 1 + 8;
  };
  // writes the body at q, returns its size
  auto fill_body = [&](int64_t r, char* q) -> size_t {
    char* const q0 = q;
    const uint64_t n = (uint64_t)(sig_offsets[r + 1] - sig_offsets[r]);
    put<uint16_t>(q, (uint16_t)id_len[r]);
    memcpy(q, ids[r], id_len[r]); q += id_len[r];
    put<uint32_t>(q, 0u);
    put<double>(q, digitisation); put<double>(q, offset[r]); put<double>(q, range); put<double>(q, sampling_rate);
    put<uint64_t>(q, n);
    if (sig_comp == S2S_BLOW5_SIGNAL_SVB_ZD) {
      const uint64_t cb = (uint64_t)svb_zd_encode(signal + sig_offsets[r], (size_t)n, reinterpret_cast<unsigned char*>(q + 8));
      put<uint64_t>(q, cb);
      q += cb;
    } else {
      memcpy(q, signal + sig_offsets[r], 2 * n); q += 2 * n;
    }
    put<uint64_t>(q, 1ull); *q++ = '0';           // channel_number "0"
    put<double>(q, median_before[r]);
    put<int32_t>(q, read_number[r]);
    put<uint8_t>(q, 0);                           // start_mux
    put<uint64_t>(q, start_time[r]);
    return (size_t)(q - q0);
  };
  std::vector<int> status((size_t)n_threads, 0);
  auto work = [&](int t) {
    const int64_t lo = n_reads * t / n_threads, hi = n_reads * (t + 1) / n_threads;
    std::string& out = chunks[t];
    if (format == S2S_SLOW5_ASCII) {
      for (int64_t r = lo; r < hi; ++r) {
        const int64_t n = sig_offsets[r + 1] - sig_offsets[r];
        out.append(ids[r], id_len[r]);
        out += "\t0\t" + fmt_double(digitisation) + "\t" + fmt_double(offset[r]) + "\t" + fmt_double(range) + "\t" +
               fmt_double(sampling_rate) + "\t" + std::to_string(n) + "\t";
        char num[8];
        for (int64_t i = 0; i < n; ++i) {
          int len = snprintf(num, sizeof num, "%d", (int)signal[sig_offsets[r] + i]);
          if (i) out.push_back(',');
          out.append(num, (size_t)len);
        }
        out += "\t0\t" + fmt_double(median_before[r]) + "\t" + std::to_string(read_number[r]) + "\t0\t" +
               std::to_string((unsigned long long)start_time[r]) + "\n";
      }
      return;
    }
    if (rec_comp == S2S_BLOW5_COMPRESS_NONE) {
      size_t total = 0;
      for (int64_t r = lo; r < hi; ++r) total += 8 + body_size(r);
      out.resize(total);
      char* q = &out[0];
      for (int64_t r = lo; r < hi; ++r) {
        const size_t bs = fill_body(r, q + 8);
        put<uint64_t>(q, (uint64_t)bs);
        q += bs;
      }
      out.resize((size_t)(q - &out[0]));
      return;
    }
    std::vector<char> body;
    std::vector<unsigned char> comp;
    for (int64_t r = lo; r < hi; ++r) {
      body.resize(body_size(r));
      const size_t bs = fill_body(r, body.data());
      uLongf clen = compressBound((uLong)bs);
      comp.resize(clen);
      if (compress2(comp.data(), &clen, reinterpret_cast<const Bytef*>(body.data()), (uLong)bs, Z_DEFAULT_COMPRESSION) != Z_OK) {
        status[t] = -1;
        return;
      }
      const uint64_t sz = (uint64_t)clen;
      out.append(reinterpret_cast<const char*>(&sz), 8);
      out.append(reinterpret_cast<const char*>(comp.data()), clen);
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (int t = 0; t < n_threads; ++t)
    if (status[t]) { set_error("zlib compression failed"); return -1; }
  return 0;
}

}  // namespace

extern "C" {

const char* s2s_blow5_last_error(void) { return g_err; }

int s2s_blow5_open(const char* path, int format, int append, int record_compression, const char* header_attrs,
                   s2s_blow5_handle* out) {
  if (!out || !path) { set_error("null argument"); return -1; }
  *out = nullptr;
  if (format != S2S_BLOW5_BINARY && format != S2S_SLOW5_ASCII) { set_error("unknown format %d", format); return -1; }
  if ((record_compression & 0xFF) > S2S_BLOW5_COMPRESS_ZLIB || ((record_compression >> 8) & 0xFF) > S2S_BLOW5_SIGNAL_SVB_ZD ||
      record_compression < 0 || record_compression > 0xFFFF) {
    set_error("unknown record / signal compression 0x%x", record_compression);
    return -1;
  }
  s2s_blow5_writer* w = new s2s_blow5_writer();
  w->format = format;
  w->compression = format == S2S_BLOW5_BINARY ? record_compression : S2S_BLOW5_COMPRESS_NONE;
  if (append) {
    w->fp = fopen(path, "r+b");
    if (!w->fp) { set_error("cannot open %s for appending", path); delete w; return -1; }
    if (format == S2S_BLOW5_BINARY) {
      // adopt the file's record compression, drop its end marker
      unsigned char head[16];
      if (fread(head, 1, 16, w->fp) != 16 || memcmp(head, kMagic, 6) != 0) {
        set_error("%s is not a BLOW5 file", path); fclose(w->fp); delete w; return -1;
      }
      w->compression = head[9] | (head[14] << 8);
      if (head[9] > S2S_BLOW5_COMPRESS_ZLIB || head[14] > S2S_BLOW5_SIGNAL_SVB_ZD) {
        set_error("%s uses record / signal compression %d / %d, which this writer cannot append to", path, head[9], head[14]);
        fclose(w->fp); delete w; return -1;
      }
      char tail[5];
      if (fseek(w->fp, -5, SEEK_END) != 0 || fread(tail, 1, 5, w->fp) != 5 || memcmp(tail, kEof, 5) != 0) {
        set_error("%s has no BLOW5 end marker (truncated file?)", path); fclose(w->fp); delete w; return -1;
      }
      fseek(w->fp, -5, SEEK_END);
    } else {
      fseek(w->fp, 0, SEEK_END);
    }
    w->bytes = ftell(w->fp);
  } else {
    w->fp = fopen(path, "wb");
    if (!w->fp) { set_error("cannot create %s", path); delete w; return -1; }
    const std::string hdr = header_bytes(format, w->compression, header_attrs);
    if (fwrite(hdr.data(), 1, hdr.size(), w->fp) != hdr.size()) {
      set_error("short write of the header to %s", path); fclose(w->fp); delete w; return -1;
    }
    w->bytes = (int64_t)hdr.size();
  }
  *out = w;
  return 0;
}

int s2s_blow5_write_batch(s2s_blow5_handle h, int64_t n_reads, const char* read_ids, const int16_t* signal,
                          const int64_t* sig_offsets, const double* offset, const double* median_before,
                          const int32_t* read_number, const uint64_t* start_time, double digitisation, double range,
                          double sampling_rate, int32_t n_threads) {
  if (!h || !h->fp) { set_error("writer is not open"); return -1; }
  if (n_reads <= 0) return 0;
  std::vector<std::string> chunks;
  if (encode_records(h->format, h->compression, n_reads, read_ids, signal, sig_offsets, offset, median_before, read_number,
                     start_time, digitisation, range, sampling_rate, n_threads, chunks))
    return -1;
  for (auto& c : chunks) {
    if (!c.empty() && fwrite(c.data(), 1, c.size(), h->fp) != c.size()) {
      set_error("short write (disk full?)");
      return -1;
    }
    h->bytes += (int64_t)c.size();
  }
  return 0;
}

int s2s_blow5_header(int format, int record_compression, const char* header_attrs, char** out, int64_t* out_bytes) {
  if (!out || !out_bytes) { set_error("null argument"); return -1; }
  const std::string hdr = header_bytes(format, format == S2S_BLOW5_BINARY ? record_compression : S2S_BLOW5_COMPRESS_NONE,
                                       header_attrs);
  *out = static_cast<char*>(malloc(hdr.size() ? hdr.size() : 1));
  if (!*out) { set_error("out of memory"); return -1; }
  memcpy(*out, hdr.data(), hdr.size());
  *out_bytes = (int64_t)hdr.size();
  return 0;
}

int s2s_blow5_encode_batch(int format, int record_compression, int64_t n_reads, const char* read_ids,
                           const int16_t* signal, const int64_t* sig_offsets, const double* offset,
                           const double* median_before, const int32_t* read_number, const uint64_t* start_time,
                           double digitisation, double range, double sampling_rate, int32_t n_threads, char** out,
                           int64_t* out_bytes) {
  if (!out || !out_bytes) { set_error("null argument"); return -1; }
  *out = nullptr;
  *out_bytes = 0;
  if (n_reads <= 0) return 0;
  std::vector<std::string> chunks;
  if (encode_records(format, format == S2S_BLOW5_BINARY ? record_compression : S2S_BLOW5_COMPRESS_NONE, n_reads, read_ids,
                     signal, sig_offsets, offset, median_before, read_number, start_time, digitisation, range,
                     sampling_rate, n_threads, chunks))
    return -1;
  size_t total = 0;
  for (auto& c : chunks) total += c.size();
  char* buf = static_cast<char*>(malloc(total ? total : 1));
  if (!buf) { set_error("out of memory (%zu bytes)", total); return -1; }
  size_t pos = 0;
  for (auto& c : chunks) { memcpy(buf + pos, c.data(), c.size()); pos += c.size(); }
  *out = buf;
  *out_bytes = (int64_t)total;
  return 0;
}

void s2s_blow5_free(char* buf) { free(buf); }

int64_t s2s_blow5_bytes_written(s2s_blow5_handle h) { return h ? h->bytes : -1; }

int s2s_blow5_close(s2s_blow5_handle h) {
  if (!h) return 0;
  int rc = 0;
  if (h->fp) {
    if (h->format == S2S_BLOW5_BINARY) {
      if (fwrite(kEof, 1, 5, h->fp) != 5) { set_error("short write of the end marker"); rc = -1; }
      // an append into a longer file cannot happen (we only ever extend), but keep the file exact
      fflush(h->fp);
    }
    if (fclose(h->fp) != 0) { set_error("fclose failed"); rc = -1; }
  }
  delete h;
  return rc;
}

}  // extern "C"
