// reassign.cu — the operators either side of the embedding kernels in the dense passes of
// /root/reference/anti_stick_diarize.py:
//
//   sd_scd_peaks            scd_split_segments :102-116 — adjacent cosine distance of a segment's sliding-window
//                           embeddings, z-score, scipy.signal.find_peaks(z, height=thr), for ALL segments in one launch
//   sd_speaker_centroids    speaker_centroids :333-349 — per-speaker mean embedding, / (norm + 1e-8)
//   sd_label_runs           _labels_to_segments :370-386 — run-length encoding of the window labels (-1 = no speech)
//   sd_merge_adjacent       merge_adjacent :464-475 — same-speaker neighbours closer than `gap` seconds merge
//
// All four are tiny scans / reductions; they exist so that the dense pass (embed every 0.1 s -> score -> labels ->
// segments) never leaves the device between the audio upload and the final segment list.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "sd_ptx.cuh"
#include "sd_status.h"

using namespace sd;

namespace {

// ------------------------------------------------------------------------------------------ SCD
// One CTA per segment.  seg_off[s] .. seg_off[s+1] are the rows of `emb` holding the segment's windows (in time
// order).  dist[i] = 1 - cos(e_i, e_{i+1}) (denominator + 1e-8, :102-104); z = (dist - mean) / std when
// std > 1e-6 else dist (:106-109, numpy's population std); peak[i] = 1 where scipy's find_peaks(z, height=thr)
// reports a peak: a strict local maximum, or the middle (floor) of a flat top, never the first / last sample.
// zbuf: [total] f32 scratch + output of z (row seg_off[s] + i holds z_i; the last row of a segment is unused).
__global__ void __launch_bounds__(256)
scd_peaks_kernel(const float* __restrict__ emb, int D, const int32_t* __restrict__ seg_off, float thr,
                 float* __restrict__ zbuf, uint8_t* __restrict__ peak) {
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = seg_off[s], n = seg_off[s + 1] - r0;   // windows
  const int m = n - 1;                                  // distances
  for (int i = tid; i < n; i += 256) peak[r0 + i] = 0;
  if (m < 1) return;
  float* z = zbuf + r0;
  for (int i = warp; i < m; i += 8) {
    const float* a = emb + static_cast<size_t>(r0 + i) * D;
    const float* b = a + D;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float x = a[c], y = b[c];
      ab = fmaf(x, y, ab);
      aa = fmaf(x, x, aa);
      bb = fmaf(y, y, bb);
    }
    ab = warp_sum(ab);
    aa = warp_sum(aa);
    bb = warp_sum(bb);
    if (lane == 0) z[i] = 1.0f - ab / (sqrtf(aa) * sqrtf(bb) + 1e-8f);
  }
  __syncthreads();
  // mean and population std in f64 (numpy reduces the f32 array pairwise; the two agree to ~1e-7 relative)
  double sm = 0.0;
  for (int i = tid; i < m; i += 256) sm += z[i];
  __shared__ double dred[8];
  for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
  if (lane == 0) dred[warp] = sm;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < 8; ++w) tot += dred[w];
  const double mean = tot / m;
  __syncthreads();
  double sq = 0.0;
  for (int i = tid; i < m; i += 256) {
    const double d = z[i] - mean;
    sq += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) dred[warp] = sq;
  __syncthreads();
  double tq = 0.0;
  for (int w = 0; w < 8; ++w) tq += dred[w];
  const float stdv = static_cast<float>(sqrt(tq / m));
  const float meanf = static_cast<float>(mean);
  __syncthreads();
  if (stdv > 1e-6f)
    for (int i = tid; i < m; i += 256) z[i] = (z[i] - meanf) / stdv;
  __syncthreads();
  // peaks: thread i owns a possible LEFT edge of a flat top
  for (int i = tid + 1; i < m - 1; i += 256) {
    const float v = z[i];
    if (!(z[i - 1] < v)) continue;
    int j = i;
    while (j + 1 < m && z[j + 1] == v) ++j;
    if (j + 1 >= m || !(z[j + 1] < v)) continue;      // runs into the right border, or rises: not a peak
    if (v >= thr) peak[r0 + (i + j) / 2] = 1;
  }
}

// --------------------------------------------------------------------------- speaker centroids
// One CTA per speaker id: mean over the rows whose label equals it (f64 accumulation), divided by
// (its f32 norm + 1e-8).  A speaker without rows gets a zero row (cannot happen for ids taken from the labels).
__global__ void __launch_bounds__(256)
speaker_centroids_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels, int N, int D,
                         const int32_t* __restrict__ spk_ids, float* __restrict__ out) {
  __shared__ double part[8][256];
  __shared__ int cnt[8];
  __shared__ float nrm[8];
  const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int id = spk_ids[k];
  // warp w scans rows w, w+8, ...; lane covers columns lane, lane+32, ... (D <= 256)
  double acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.0;
  int c = 0;
  for (int r = warp; r < N; r += 8) {
    if (labels[r] != id) continue;
    ++c;
    const float* p = emb + static_cast<size_t>(r) * D;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (lane + 32 * q < D) acc[q] += p[lane + 32 * q];
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) part[warp][lane + 32 * q] = acc[q];
  if (lane == 0) cnt[warp] = c;
  __syncthreads();
  float v = 0.f;
  if (tid < D) {
    double t = 0.0;
    int n = 0;
    for (int w = 0; w < 8; ++w) {
      t += part[w][tid];
      n += cnt[w];
    }
    v = n > 0 ? static_cast<float>(t / n) : 0.f;
  }
  float ss = warp_sum(v * v);
  if (lane == 0) nrm[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += nrm[w];
  if (tid < D) out[static_cast<size_t>(k) * D + tid] = v / (sqrtf(tot) + 1e-8f);
}

// ------------------------------------------------------------------------------ block scan
__device__ __forceinline__ int block_scan_incl_1024(int v, int* wt /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wt[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int t = wt[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    wt[lane] = t;
  }
  __syncthreads();
  const int r = x + (warp > 0 ? wt[warp - 1] : 0);
  __syncthreads();
  return r;
}

// Run-length encoding of labels[0, n) (anti_stick_diarize.py:374-386): a run starts where the label changes
// (position 0 always starts one: np.diff(..., prepend=nan)); runs of -1 are dropped; a run [i, j) of speaker k
// becomes (start = ws[i] / sr, end = j < n ? ws[j] / sr : max_t) and is kept when end > start.
// Out: run_idx [3 * count] = (i, j, k), run_t [2 * count] f64 = (start, end).  Single CTA, 1024 threads.
__global__ void __launch_bounds__(1024)
label_runs_kernel(const int32_t* __restrict__ labels, int n, const int64_t* __restrict__ ws, double sr, double max_t,
                  int32_t* __restrict__ start_idx /*[n] scratch*/, int32_t* __restrict__ run_idx, double* __restrict__ run_t,
                  int32_t* __restrict__ count) {
  __shared__ int wt[32];
  __shared__ int n_runs, n_kept;
  const int tid = threadIdx.x;
  if (tid == 0) { n_runs = 0; n_kept = 0; }
  __syncthreads();
  // pass 1: compact the run starts
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int first = i < n && (i == 0 || labels[i] != labels[i - 1]);
    const int p = block_scan_incl_1024(first, wt);
    const int b0 = n_runs;
    if (first) start_idx[b0 + p - 1] = i;
    __syncthreads();
    if (tid == 1023) n_runs = b0 + p;
    __syncthreads();
  }
  const int runs = n_runs;
  // pass 2: keep speaker runs of positive duration
  for (int base = 0; base < runs; base += 1024) {
    const int r = base + tid;
    int keep = 0, i = 0, j = 0, k = -1;
    double t0 = 0.0, t1 = 0.0;
    if (r < runs) {
      i = start_idx[r];
      j = r + 1 < runs ? start_idx[r + 1] : n;
      k = labels[i];
      t0 = static_cast<double>(ws[i]) / sr;
      t1 = j < n ? static_cast<double>(ws[j]) / sr : max_t;
      keep = k != -1 && t1 > t0;
    }
    const int p = block_scan_incl_1024(keep, wt);
    const int b0 = n_kept;
    if (keep) {
      const int o = b0 + p - 1;
      run_idx[3 * o] = i;
      run_idx[3 * o + 1] = j;
      run_idx[3 * o + 2] = k;
      run_t[2 * o] = t0;
      run_t[2 * o + 1] = t1;
    }
    __syncthreads();
    if (tid == 1023) n_kept = b0 + p;
    __syncthreads();
  }
  if (tid == 0) *count = n_kept;
}

// merge_adjacent (:464-475): segment k opens a new group unless spk[k] == spk[k-1] and start[k] - end[k-1] <= gap
// (the running merged segment always ends where segment k-1 ends).  Out: group [2 * count] = (first, last) segment
// index of each merged group.  n is read from *n_dev when n_dev != nullptr (chained behind label_runs_kernel).
__global__ void __launch_bounds__(1024)
merge_adjacent_kernel(const double* __restrict__ seg_t /*[2n]*/, const int32_t* __restrict__ spk, int spk_stride,
                      int n_host, const int32_t* __restrict__ n_dev, double gap, int32_t* __restrict__ group,
                      int32_t* __restrict__ count) {
  __shared__ int wt[32];
  __shared__ int n_grp;
  const int tid = threadIdx.x;
  const int n = n_dev != nullptr ? *n_dev : n_host;
  if (tid == 0) n_grp = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int k = base + tid;
    int first = 0, last = 0;
    if (k < n) {
      first = k == 0 || spk[k * spk_stride] != spk[(k - 1) * spk_stride] || (seg_t[2 * k] - seg_t[2 * k - 1]) > gap;
      last = k == n - 1 || spk[(k + 1) * spk_stride] != spk[k * spk_stride] || (seg_t[2 * k + 2] - seg_t[2 * k + 1]) > gap;
    }
    const int p = block_scan_incl_1024(first, wt);
    const int b0 = n_grp;
    if (first) group[2 * (b0 + p - 1)] = k;
    if (last) group[2 * (b0 + p - 1) + 1] = k;
    __syncthreads();
    if (tid == 1023) n_grp = b0 + p;
    __syncthreads();
  }
  if (tid == 0) *count = n_grp;
}

// full[valid[i]] = labels[i]  (full pre-filled with -1 by the caller): anti_stick_diarize.py:374-375
__global__ void __launch_bounds__(256)
scatter_labels_kernel(const int32_t* __restrict__ valid, const int32_t* __restrict__ labels,
                      const int32_t* __restrict__ label_map /*nullable: spk id of cluster index*/, int m,
                      int32_t* __restrict__ full) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int l = labels[i];
  full[valid[i]] = label_map != nullptr ? label_map[l] : l;
}

// Zero-padded batch of variable-length snippets (embed_segments :162-166): out[b, i] = audio[start[b] + i] for
// i < len[b], else 0.  grid (ceil(max_len / 1024), B).
__global__ void __launch_bounds__(256)
gather_pad_kernel(const float* __restrict__ audio, const int64_t* __restrict__ start, const int32_t* __restrict__ len,
                  int max_len, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (i0 >= max_len) return;
  const int n = len[b];
  const float* src = audio + start[b];
  float* dst = out + static_cast<size_t>(b) * max_len + i0;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (i0 + e < max_len) dst[e] = i0 + e < n ? __ldg(src + i0 + e) : 0.f;
}

}  // namespace

extern "C" int sd_gather_pad_f32(const float* audio_dev, const int64_t* start_dev, const int32_t* len_dev, int B,
                                 int max_len, float* out_dev, void* stream) {
  if (B < 0 || max_len < 0 || (B > 0 && max_len > 0 && (!audio_dev || !start_dev || !len_dev || !out_dev)) || B > 65535)
    return fail(SD_ERR_ARG, "sd_gather_pad_f32: bad arguments B=%d max_len=%d", B, max_len);
  if (B == 0 || max_len == 0) return SD_OK;
  gather_pad_kernel<<<dim3((max_len + 1023) / 1024, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(audio_dev, start_dev, len_dev,
                                                                                                  max_len, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_scd_peaks(const float* emb_dev, int D, const int32_t* seg_off_dev, int n_segments, float thr,
                            float* z_dev, uint8_t* peak_dev, void* stream) {
  if (n_segments < 0 || D < 1 || (n_segments > 0 && (!emb_dev || !seg_off_dev || !z_dev || !peak_dev)))
    return fail(SD_ERR_ARG, "sd_scd_peaks: bad arguments");
  if (n_segments == 0) return SD_OK;
  scd_peaks_kernel<<<n_segments, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb_dev, D, seg_off_dev, thr, z_dev, peak_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_speaker_centroids(const float* emb_dev, const int32_t* labels_dev, int N, int D,
                                    const int32_t* spk_ids_dev, int K, float* out_dev, void* stream) {
  if (N < 0 || K < 0 || D < 1 || D > 256 || (K > 0 && (!spk_ids_dev || !out_dev)) || (N > 0 && (!emb_dev || !labels_dev)))
    return fail(SD_ERR_ARG, "sd_speaker_centroids: bad arguments N=%d D=%d K=%d", N, D, K);
  if (K == 0) return SD_OK;
  speaker_centroids_kernel<<<K, 256, 0, static_cast<cudaStream_t>(stream)>>>(emb_dev, labels_dev, N, D, spk_ids_dev, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_scatter_labels(const int32_t* valid_dev, const int32_t* labels_dev, const int32_t* label_map_dev,
                                 int m, int32_t* full_dev, void* stream) {
  if (m < 0 || (m > 0 && (!valid_dev || !labels_dev || !full_dev))) return fail(SD_ERR_ARG, "sd_scatter_labels: bad arguments");
  if (m == 0) return SD_OK;
  scatter_labels_kernel<<<(m + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(valid_dev, labels_dev, label_map_dev, m, full_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_label_runs(const int32_t* labels_dev, int n, const int64_t* window_starts_dev, double sr, double max_t,
                             int32_t* scratch_dev, int32_t* run_idx_dev, double* run_t_dev, int32_t* count_dev, void* stream) {
  if (n < 0 || !count_dev || (n > 0 && (!labels_dev || !window_starts_dev || !scratch_dev || !run_idx_dev || !run_t_dev)) ||
      !(sr > 0.0))
    return fail(SD_ERR_ARG, "sd_label_runs: bad arguments");
  label_runs_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(labels_dev, n, window_starts_dev, sr, max_t, scratch_dev,
                                                                     run_idx_dev, run_t_dev, count_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_merge_adjacent(const double* seg_t_dev, const int32_t* spk_dev, int spk_stride, int n,
                                 const int32_t* n_dev, double gap, int32_t* group_dev, int32_t* count_dev, void* stream) {
  if (n < 0 || spk_stride < 1 || !count_dev || (n > 0 && (!seg_t_dev || !spk_dev || !group_dev)))
    return fail(SD_ERR_ARG, "sd_merge_adjacent: bad arguments");
  merge_adjacent_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(seg_t_dev, spk_dev, spk_stride, n, n_dev, gap,
                                                                         group_dev, count_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}
