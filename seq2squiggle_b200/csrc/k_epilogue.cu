// K-F / K-G: decoder output head, noise, clamp, zero-strip, digitisation and per-read compaction.
//
// (fp32 parity path and the stand-alone stage entry points; on the tensor-core path p, x165, noise, clamp and the per-chunk
//  non-zero count are fused into the last FFN kernel's epilogue, k_tc.cu, and only the compaction below remains)
//   modules.py:140-141   p = ReLU(out_linear(y))
//   model.py:221-240     pA = 165 p; where pA != 0: += N(0, clamp(sigma_ext,min_noise)*noise_std*165)  (sampler)
//                                                     or N(0, noise_std)                                  (static)
//                        pA = clamp(pA, min=0)
//   model.py:284-286     per read: concatenate chunk rows, drop every exact 0.0
//   signal_io.py:134-141 raw = int16(round_half_even(float32(pA) * digitisation / range - offset_mean)),
//                        RNA profiles reversed
#include "s2s_kernels.h"

namespace s2s {

constexpr uint32_t kStreamNoise = 0x5D0003u;

__device__ __forceinline__ int16_t digitise_one(float pa, float dig, float range, float offset) {
  // NumPy evaluates the expression in float32, left to right, one rounding per operation.
  float v = __fsub_rn(__fdiv_rn(__fmul_rn(pa, dig), range), offset);
  return (int16_t)(int32_t)rintf(v);  // np.round = half-to-even; astype(int16) wraps
}

// 16 lanes per row: float4 each of the 64-wide decoder output row.
__global__ void __launch_bounds__(256) k_out_epilogue(const float* __restrict__ y, const float* __restrict__ w_out,
                                                      const float* __restrict__ b_out,
                                                      const float* __restrict__ sigma_ext, int64_t n_pos, float scaling,
                                                      s2s_run_opts o, float* __restrict__ p_tap, float* __restrict__ pa) {
  const int q = threadIdx.x & 15;
  const float4 w = *reinterpret_cast<const float4*>(w_out + 4 * q);
  const float b = b_out[0];
  const Philox ph(o.seed);
  for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; idx < n_pos;
       idx += ((int64_t)gridDim.x * blockDim.x) >> 4) {
    const int64_t c = idx / S2S_L_DEC;
    const int t = (int)(idx - c * S2S_L_DEC);
    float4 v = *reinterpret_cast<const float4*>(y + ((size_t)c * S2S_L_DEC_PAD + t) * S2S_D + 4 * q);
    float s = v.x * w.x;
    s = fmaf(v.y, w.y, s); s = fmaf(v.z, w.z, s); s = fmaf(v.w, w.w, s);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (q == 0) {
      float p = fmaxf(s + b, 0.f);
      if (p_tap) p_tap[idx] = p;
      float v_pa = p * scaling;
      if (o.noise_mode != S2S_NOISE_OFF && v_pa != 0.f) {
        const uint64_t gc = o.chunk_id_base + (uint64_t)c;
        uint4 r = ph((uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)t, kStreamNoise);
        float z = box_muller(r.x, r.y).x;
        float sd = o.noise_mode == S2S_NOISE_SAMPLER
                       ? fmaxf(sigma_ext[idx], o.min_noise) * o.noise_std * scaling  // model.py:228-230
                       : o.noise_std;                                                 // model.py:236
        v_pa += z * sd;
      }
      pa[idx] = fmaxf(v_pa, 0.f);
    }
  }
}

__global__ void k_digitise(const float* __restrict__ pa, int64_t n, float dig, float range, float offset,
                           int16_t* __restrict__ raw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    raw[i] = digitise_one(pa[i], dig, range, offset);
}

// ---- compaction -------------------------------------------------------------------------------
// warp per chunk: number of surviving (non-zero) samples
__global__ void __launch_bounds__(256) k_count_nonzero(const float* __restrict__ pa, int64_t n_chunks,
                                                       int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_chunks) return;
  int n = 0;
  for (int t = lane; t < S2S_L_DEC; t += 32) n += (pa[c * S2S_L_DEC + t] != 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) counts[c] = n;
}

constexpr int kScanTile = 2048;  // counts per CTA (256 threads x 8)

__device__ __forceinline__ int64_t block_exclusive_scan_256(int64_t v, int64_t* s_warp, int64_t& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int64_t warp_off = 0, tot = 0;
#pragma unroll
  for (int wi = 0; wi < 8; ++wi) {
    int64_t sw = s_warp[wi];
    if (wi < warp) warp_off += sw;
    tot += sw;
  }
  block_total = tot;
  return warp_off + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_tile_sums(const int32_t* __restrict__ counts, int64_t n,
                                                        int64_t* __restrict__ tile_sums) {
  __shared__ int64_t s_warp[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  int64_t v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) if (base + i < n) v += counts[base + i];
  int64_t tot;
  block_exclusive_scan_256(v, s_warp, tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void k_scan_tiles_serial(int64_t* __restrict__ tile_sums, int64_t n_tiles) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t run = 0;
    for (int64_t i = 0; i < n_tiles; ++i) { int64_t v = tile_sums[i]; tile_sums[i] = run; run += v; }
    tile_sums[n_tiles] = run;
  }
}

__global__ void __launch_bounds__(256) k_scan_apply(const int32_t* __restrict__ counts, int64_t n,
                                                    const int64_t* __restrict__ tile_sums,
                                                    int64_t* __restrict__ chunk_out /*[n+1]*/, int64_t n_tiles) {
  __shared__ int64_t s_warp[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  int32_t loc[8];
  int64_t v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { loc[i] = (base + i < n) ? counts[base + i] : 0; v += loc[i]; }
  int64_t tot;
  int64_t off = block_exclusive_scan_256(v, s_warp, tot) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (base + i < n) chunk_out[base + i] = off;
    off += loc[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) chunk_out[n] = tile_sums[n_tiles];
}

__global__ void k_read_offsets(const int64_t* __restrict__ chunk_offsets, const int64_t* __restrict__ chunk_out,
                               int64_t n_reads, int64_t* __restrict__ raw_offsets) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_reads) raw_offsets[r] = chunk_out[chunk_offsets[r]];
}

// warp per chunk: ordered scatter of the surviving samples, digitised on the way out.
__global__ void __launch_bounds__(256) k_compact(const float* __restrict__ pa, const int64_t* __restrict__ chunk_offsets,
                                                 const int64_t* __restrict__ chunk_out,
                                                 const int64_t* __restrict__ raw_offsets, int64_t n_reads,
                                                 int64_t n_chunks, float dig, float range, float offset, int rna_reverse,
                                                 int16_t* __restrict__ raw) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_chunks) return;
  int64_t dst = chunk_out[c];
  int64_t rd_lo = 0, rd_hi = 0;
  if (rna_reverse) {  // find the read of this chunk: largest r with chunk_offsets[r] <= c
    int64_t lo = 0, hi = n_reads;
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (chunk_offsets[mid] <= c) lo = mid; else hi = mid;
    }
    rd_lo = raw_offsets[lo];
    rd_hi = raw_offsets[lo + 1];
  }
  // all eight loads of the chunk's 250 floats first (independent, 1000 contiguous bytes), then the ordered scatter
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = 32 * i + lane;
    v[i] = t < S2S_L_DEC ? __ldg(pa + c * S2S_L_DEC + t) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool keep = v[i] != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      int64_t pos = dst + __popc(m & ((1u << lane) - 1u));
      if (rna_reverse) pos = rd_hi - 1 - (pos - rd_lo);
      raw[pos] = digitise_one(v[i], dig, range, offset);
    }
    dst += __popc(m);
  }
}

int launch_out_epilogue(const DevWeights& w, const float* y, const float* sigma_ext, int64_t n_chunks,
                        const s2s_run_opts& o, float* p_tap, float* pa, cudaStream_t st) {
  if (n_chunks == 0) return 0;
  const int64_t n_pos = n_chunks * S2S_L_DEC;
  int64_t blocks = ceil_div(n_pos, 16);
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_out_epilogue<<<(unsigned)blocks, 256, 0, st>>>(y, w.out_w, w.out_b, sigma_ext, n_pos, w.cfg.scaling_max_value, o,
                                                   p_tap, pa);
  S2S_LAUNCH_CHECK();
  return 0;
}

int launch_digitise(const float* pa, int64_t n, float dig, float range, float offset, int16_t* raw, cudaStream_t st) {
  if (n == 0) return 0;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_digitise<<<(unsigned)blocks, 256, 0, st>>>(pa, n, dig, range, offset, raw);
  S2S_LAUNCH_CHECK();
  return 0;
}

int64_t compact_workspace_bytes(int64_t n_chunks) {
  int64_t n_tiles = ceil_div(n_chunks, kScanTile);
  return align_up(n_chunks * 4, 256) + align_up((n_chunks + 1) * 8, 256) + align_up((n_tiles + 1) * 8, 256);
}

int launch_compact(const float* pa, const int64_t* chunk_offsets, int64_t n_reads, int64_t n_chunks, float dig,
                   float range, float offset, int rna_reverse, void* ws, int64_t ws_bytes, int16_t* raw,
                   int64_t* raw_offsets, cudaStream_t st, bool counts_ready) {
  if (ws_bytes < compact_workspace_bytes(n_chunks)) {
    set_error("launch_compact: workspace too small (%lld < %lld)", (long long)ws_bytes,
              (long long)compact_workspace_bytes(n_chunks));
    return -1;
  }
  if (n_chunks == 0) {
    if (n_reads >= 0) S2S_CUDA_OK(cudaMemsetAsync(raw_offsets, 0, (n_reads + 1) * 8, st));
    return 0;
  }
  const int64_t n_tiles = ceil_div(n_chunks, kScanTile);
  char* p = static_cast<char*>(ws);
  int32_t* counts = reinterpret_cast<int32_t*>(p); p += align_up(n_chunks * 4, 256);
  int64_t* chunk_out = reinterpret_cast<int64_t*>(p); p += align_up((n_chunks + 1) * 8, 256);
  int64_t* tile_sums = reinterpret_cast<int64_t*>(p);
  const unsigned warp_blocks = (unsigned)ceil_div(n_chunks, 8);
  if (!counts_ready) {   // fp32 path / stand-alone compaction: count here; the tensor-core decoder counts in its epilogue
    k_count_nonzero<<<warp_blocks, 256, 0, st>>>(pa, n_chunks, counts);
    S2S_LAUNCH_CHECK();
  }
  k_scan_tile_sums<<<(unsigned)n_tiles, 256, 0, st>>>(counts, n_chunks, tile_sums);
  S2S_LAUNCH_CHECK();
  k_scan_tiles_serial<<<1, 32, 0, st>>>(tile_sums, n_tiles);
  S2S_LAUNCH_CHECK();
  k_scan_apply<<<(unsigned)n_tiles, 256, 0, st>>>(counts, n_chunks, tile_sums, chunk_out, n_tiles);
  S2S_LAUNCH_CHECK();
  k_read_offsets<<<(unsigned)ceil_div(n_reads + 1, 256), 256, 0, st>>>(chunk_offsets, chunk_out, n_reads, raw_offsets);
  S2S_LAUNCH_CHECK();
  k_compact<<<warp_blocks, 256, 0, st>>>(pa, chunk_offsets, chunk_out, raw_offsets, n_reads, n_chunks, dig, range,
                                         offset, rna_reverse, raw);
  S2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace s2s
