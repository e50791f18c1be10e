// post.cu — score post-processing and VAD mask operators: the small stages
// either side of the embedding/clustering hot path that the reference runs as numpy / numba / scipy.
//
//   sd_viterbi_hmm        diar_diag.viterbi_hmm       (/root/reference/diar_diag.py:231-247)
//   sd_asnorm_scores      diar_diag.asnorm_scores     (diar_diag.py:196-208; call site :389)
//   sd_hysteresis_u8      vad.hysteresis_binarize     (/root/reference/vad.py:59-74)
//   sd_morph_open_close_u8 vad.morph_open_close       (vad.py:77-87; scipy binary_opening / binary_closing)
//   sd_mask_segments_i32  vad.mask_to_segments        (vad.py:90-163, the frame-index part)
//
// Integer / boolean results are bit-exact with the reference; Viterbi reproduces the reference's
// float32 recursion operation for operation (one rounded add per candidate, first maximum wins), so
// paths are identical, ties included.  AS-norm is floating point: cohort similarities come from the
// split-f16 tensor-core affinity kernel (|err| < 2e-6), statistics are accumulated in f64.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include "gemm_host.cuh"
#include "sd_status.h"

using namespace sd;

namespace {

// ================================================================================ Viterbi
constexpr int VT_CHUNK = 128;   // frames per back-tracking chunk
constexpr int VT_PF = 32;       // score rows prefetched ahead of the recursion (covers ~1 us of DRAM latency)

template <typename ScoreT>
__device__ __forceinline__ float vt_emit(float prev_best, ScoreT s);
// numpy: prev[ptr, k] (f32) + scores[t] -> f32 add for f32 scores; f64 add rounded on store for f64 scores
template <> __device__ __forceinline__ float vt_emit<float>(float p, float s) { return __fadd_rn(p, s); }
template <> __device__ __forceinline__ float vt_emit<double>(float p, double s) {
  return __double2float_rn(__dadd_rn(static_cast<double>(p), s));
}

// Forward recursion: ONE warp, lane j = state j (K <= 32).  The T steps are strictly dependent, so the
// kernel is a latency chain (K shuffles, K adds, K compare-selects per step).  Score rows are fetched a
// BLOCK of VT_PF steps ahead into a second register buffer, all loads issued together at the top of a block,
// so no load sits between the steps (single loads interleaved with the steps share scoreboard slots and
// stalled every step: 250 cycles/step measured).  K <= 8 is unrolled at compile time (KT; KT = 0 = generic
// loop).  Back-pointers of four consecutive steps are packed into one u32 per state: ptr4[t / 4][j].
template <typename ScoreT, int KT>
__device__ __forceinline__ float vt_step(float dp, ScoreT s, int K, int lane, float log_stay, float log_move, int& arg) {
  float best = __fadd_rn(__shfl_sync(0xffffffffu, dp, 0), 0 == lane ? log_stay : log_move);
  arg = 0;
  if (KT > 0) {
#pragma unroll
    for (int i = 1; i < KT; ++i) {
      const float v = __fadd_rn(__shfl_sync(0xffffffffu, dp, i), i == lane ? log_stay : log_move);
      if (v > best) { best = v; arg = i; }       // strict '>' : first maximum, as np.argmax
    }
  } else {
    for (int i = 1; i < K; ++i) {
      const float v = __fadd_rn(__shfl_sync(0xffffffffu, dp, i), i == lane ? log_stay : log_move);
      if (v > best) { best = v; arg = i; }
    }
  }
  return vt_emit<ScoreT>(best, s);
}

template <typename ScoreT, int KT>
__global__ void __launch_bounds__(32)
viterbi_forward_kernel(const ScoreT* __restrict__ scores, int T, int K_rt, float log_stay, float log_move,
                       uint32_t* __restrict__ ptr4, int* __restrict__ last_state) {
  const int K = KT > 0 ? KT : K_rt;
  const int lane = threadIdx.x;
  const bool act = lane < K;
  const int col = act ? lane : 0;
  ScoreT cur[VT_PF], nxt[VT_PF];
  // block b covers steps t = b * VT_PF .. b * VT_PF + VT_PF - 1 (t = 0 is the initialisation)
#pragma unroll
  for (int u = 0; u < VT_PF; ++u) cur[u] = u < T ? scores[static_cast<size_t>(u) * K + col] : ScoreT(0);
  float dp = static_cast<float>(cur[0]);   // dp[0] = scores[0] (stored as f32)
  for (int t0 = 0; t0 < T; t0 += VT_PF) {
#pragma unroll
    for (int u = 0; u < VT_PF; ++u) {
      const int tn = t0 + VT_PF + u;
      nxt[u] = tn < T ? scores[static_cast<size_t>(tn) * K + col] : ScoreT(0);
    }
#pragma unroll
    for (int u4 = 0; u4 < VT_PF; u4 += 4) {
      uint32_t packed = 0;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int t = t0 + u4 + v;
        if (t >= 1 && t < T) {
          int arg;
          dp = vt_step<ScoreT, KT>(dp, cur[u4 + v], K, lane, log_stay, log_move, arg);
          packed |= static_cast<uint32_t>(arg) << (8 * v);
        }
      }
      if (act && t0 + u4 < T) ptr4[static_cast<size_t>((t0 + u4) >> 2) * K + lane] = packed;
    }
#pragma unroll
    for (int u = 0; u < VT_PF; ++u) cur[u] = nxt[u];
  }
  // path[-1] = argmax(dp[-1]) (first maximum)
  float best = __shfl_sync(0xffffffffu, dp, 0);
  int arg = 0;
  for (int i = 1; i < K; ++i) {
    const float v = __shfl_sync(0xffffffffu, dp, i);
    if (v > best) { best = v; arg = i; }
  }
  if (lane == 0) *last_state = arg;
}

template <typename ScoreT>
void launch_viterbi_forward(const void* scores, int T, int K, float ls, float lm, uint32_t* ptr, int* last, cudaStream_t st) {
  const ScoreT* sc = static_cast<const ScoreT*>(scores);
  switch (K) {
    case 2: viterbi_forward_kernel<ScoreT, 2><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 3: viterbi_forward_kernel<ScoreT, 3><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 4: viterbi_forward_kernel<ScoreT, 4><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 5: viterbi_forward_kernel<ScoreT, 5><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 6: viterbi_forward_kernel<ScoreT, 6><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 7: viterbi_forward_kernel<ScoreT, 7><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    case 8: viterbi_forward_kernel<ScoreT, 8><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
    default: viterbi_forward_kernel<ScoreT, 0><<<1, 32, 0, st>>>(sc, T, K, ls, lm, ptr, last); break;
  }
}

__device__ __forceinline__ int vt_ptr(const uint32_t* __restrict__ ptr4, int t, int K, int st) {
  return (ptr4[static_cast<size_t>(t >> 2) * K + st] >> (8 * (t & 3))) & 0xff;
}

// Back-tracking is a composition of the maps  s_{t} = ptr[t+1][s_{t+1}],  which is associative:
// (1) every chunk composes its maps for all K possible end states, (2) one warp chains the chunk maps,
// (3) every chunk re-walks with its now-known end state and writes the path.
__global__ void __launch_bounds__(32)
viterbi_chunk_maps_kernel(const uint32_t* __restrict__ ptr4, int T, int K, int32_t* __restrict__ maps) {
  const int c = blockIdx.x, s = threadIdx.x;
  if (s >= K) return;
  const int lo = c * VT_CHUNK, hi = min(lo + VT_CHUNK, T - 1);   // states at times lo .. hi
  int st = s;
  for (int t = hi; t > lo; --t) st = vt_ptr(ptr4, t, K, st);
  maps[c * K + s] = st;   // state at time lo given state s at time hi
}

constexpr int VT_CHAIN_SMEM = 8192;   // chunk-map entries staged in shared memory (T*K <= 1 M)
__global__ void __launch_bounds__(256)
viterbi_chain_kernel(const int32_t* __restrict__ maps, int n_chunks, int K, const int* __restrict__ last_state,
                     int32_t* __restrict__ end_state) {
  __shared__ int32_t sm[VT_CHAIN_SMEM];
  const bool staged = n_chunks * K <= VT_CHAIN_SMEM;
  if (staged) {
    for (int i = threadIdx.x; i < n_chunks * K; i += 256) sm[i] = maps[i];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  int st = *last_state;
  for (int c = n_chunks - 1; c >= 0; --c) {
    end_state[c] = st;
    st = staged ? sm[c * K + st] : maps[c * K + st];
  }
}

__global__ void __launch_bounds__(32)
viterbi_walk_kernel(const uint32_t* __restrict__ ptr4, int T, int K, const int32_t* __restrict__ end_state,
                    int32_t* __restrict__ path) {
  if (threadIdx.x != 0) return;
  const int c = blockIdx.x;
  const int lo = c * VT_CHUNK, hi = min(lo + VT_CHUNK, T - 1);
  int st = end_state[c];
  if (hi == T - 1) path[hi] = st;
  for (int t = hi; t > lo; --t) {
    st = vt_ptr(ptr4, t, K, st);
    path[t - 1] = st;
  }
}

// ================================================================================ AS-norm
// x / (||x|| + eps), split into hi | lo' f16 halves for the three-pass tensor-core product
// (affinity.cu explains the split); also the f32 normalised rows when `xn` is given.
__global__ void __launch_bounds__(256)
l2n_split_kernel(const float* __restrict__ x, int N, int D, float eps, __half* __restrict__ xs,
                 float* __restrict__ xn) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* p = x + static_cast<size_t>(row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  const float den = sqrtf(s) + eps;
  __half* o = xs + static_cast<size_t>(row) * 2 * D;
  for (int i = lane; i < D; i += 32) {
    const float v = p[i] / den;
    const __half hi = __float2half_rn(v);
    o[i] = hi;
    o[D + i] = __float2half_rn((v - __half2float(hi)) * 2048.0f);
    if (xn) xn[static_cast<size_t>(row) * D + i] = v;
  }
}

__device__ __forceinline__ uint32_t f32_key(float f) {   // monotone: larger float -> larger key
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Cohort statistics of one row: mean and population std of its `topk` LARGEST similarities, given the
// row of cosine DISTANCES d = 1 - s (so the topk smallest d).  One CTA per row; the row is staged in
// shared memory when it fits, the k-th value is found by a 4-pass 8-bit radix select, ties at the
// threshold are counted exactly.  Sums in f64.  out: mu[row], sigma[row] (+1e-6, diar_diag.py:204-205).
constexpr int TK_THREADS = 512;
__global__ void __launch_bounds__(TK_THREADS)
cohort_topk_stats_kernel(const float* __restrict__ dist, int ld, int nc, int topk, int stage_elems,
                         float* __restrict__ mu, float* __restrict__ sigma) {
  extern __shared__ float srow[];
  __shared__ unsigned hist[256];
  __shared__ unsigned sel_prefix, sel_remaining;
  __shared__ double red[2][TK_THREADS / 32];
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* g = dist + static_cast<size_t>(row) * ld;
  const bool staged = nc <= stage_elems;
  if (staged) {
    for (int i = tid; i < nc; i += TK_THREADS) srow[i] = 1.0f - g[i];
    __syncthreads();
  }
  auto sim = [&](int i) -> float { return staged ? srow[i] : 1.0f - g[i]; };
  // radix select of the topk-th largest key
  if (tid == 0) { sel_prefix = 0; sel_remaining = static_cast<unsigned>(topk); }
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += TK_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = sel_prefix;
    const unsigned mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < nc; i += TK_THREADS) {
      const uint32_t k = f32_key(sim(i));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned rem = sel_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      sel_prefix = prefix | (static_cast<unsigned>(b) << shift);
      sel_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t kth = sel_prefix;          // key of the topk-th largest value
  const unsigned n_ties = sel_remaining;    // how many copies of it belong to the top k
  // mean
  double s = 0.0;
  float vth = 0.f;
  for (int i = tid; i < nc; i += TK_THREADS) {
    const float v = sim(i);
    const uint32_t k = f32_key(v);
    if (k > kth) s += static_cast<double>(v);
  }
  {
    const uint32_t u = (kth & 0x80000000u) ? (kth & 0x7fffffffu) : ~kth;
    vth = __uint_as_float(u);
  }
  auto block_sum = [&](double v, int slot) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[slot][tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < TK_THREADS / 32; ++w) t += red[slot][w];
    return t;
  };
  const double total = block_sum(s, 0) + static_cast<double>(n_ties) * static_cast<double>(vth);
  const double mean = total / static_cast<double>(topk);
  double q = 0.0;
  for (int i = tid; i < nc; i += TK_THREADS) {
    const float v = sim(i);
    if (f32_key(v) > kth) {
      const double d = static_cast<double>(v) - mean;
      q += d * d;
    }
  }
  const double dt = static_cast<double>(vth) - mean;
  const double var = (block_sum(q, 1) + static_cast<double>(n_ties) * dt * dt) / static_cast<double>(topk);
  if (tid == 0) {
    mu[row] = static_cast<float>(mean);
    sigma[row] = static_cast<float>(sqrt(var)) + 1e-6f;
  }
}

// raw = Qn . Rn^T and the symmetric z-norm (diar_diag.py:199,206-208); one warp per query row.
__global__ void __launch_bounds__(256)
asnorm_combine_kernel(const float* __restrict__ qn, const float* __restrict__ rn, int nq, int nr, int D,
                      const float* __restrict__ q_mu, const float* __restrict__ q_sig,
                      const float* __restrict__ r_mu, const float* __restrict__ r_sig, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < nq; row += gridDim.x * 8) {
    const float* p = qn + static_cast<size_t>(row) * D;
    for (int k = 0; k < nr; ++k) {
      const float* c = rn + static_cast<size_t>(k) * D;
      float s = 0.f;
      for (int i = lane; i < D; i += 32) s = fmaf(p[i], c[i], s);
      s = warp_sum(s);
      if (lane == 0) {
        const float zq = (s - q_mu[row]) / q_sig[row];
        const float zr = (s - r_mu[k]) / r_sig[k];
        out[static_cast<size_t>(row) * nr + k] = 0.5f * (zq + zr);
      }
    }
  }
}

// scores[i][k] = <x_i, c_k>  (diar_diag.py:386, `embs @ centers.T`); one warp per row.
__global__ void __launch_bounds__(256)
dot_scores_kernel(const float* __restrict__ x, const float* __restrict__ cent, int N, int K, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < N; row += gridDim.x * 8) {
    const float* p = x + static_cast<size_t>(row) * D;
    for (int k = 0; k < K; ++k) {
      const float* c = cent + static_cast<size_t>(k) * D;
      float s = 0.f;
      for (int i = lane; i < D; i += 32) s = fmaf(p[i], c[i], s);
      s = warp_sum(s);
      if (lane == 0) out[static_cast<size_t>(row) * K + k] = s;
    }
  }
}

int cohort_distances(const __half* a_split, int rows_total, int row0, int rows, const __half* c_split, int nc, int D,
                     float* out, cudaStream_t st) {
  GemmParams P;
  init_params(P);
  SD_TRY(make_tmap_f16(&P.tmapA, a_split, rows_total, 2 * D, 2 * D, BM));
  SD_TRY(make_tmap_f16(&P.tmapB, c_split, nc, 2 * D, 2 * D, 128));
  P.n_tile = 128;
  P.acc_slots = 3;
  P.a_row_base = row0;
  P.num_m_blocks = (rows + BM - 1) / BM;
  P.num_n_blocks = (nc + 127) / 128;
  P.idesc = make_idesc_f16(128, 0);
  int ki = 0;
  for (int slot = 0; slot < 3; ++slot)
    for (int c = 0; c < D / 64; ++c, ++ki) {
      P.kit[ki].a_col = (slot == 2 ? D : 0) + c * 64;
      P.kit[ki].b_col = (slot == 1 ? D : 0) + c * 64;
      P.kit[ki].slot = slot;
      P.kit[ki].accum = c > 0;
    }
  P.num_kiters = ki;
  P.epi.M_rows = rows;
  P.epi.N_cols = nc;
  P.epi.out = out;
  P.epi.ld_out = nc;
  return launch_gemm<EPI_AFF>(P, st);
}

constexpr int AS_ROW_BLOCK = 4096;   // query rows per distance block (bounds the workspace)
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Unit-norm cluster centres (diar_diag.py:377-383): centre k = mean of the rows labelled k, divided by
// (norm + 1e-9).  One CTA per cluster, thread = column, f64 accumulation in row order (as numpy's mean(0)).
__global__ void __launch_bounds__(256)
cluster_centers_kernel(const double* __restrict__ x, const int32_t* __restrict__ labels, int N, int D,
                       double* __restrict__ out64, float* __restrict__ out32) {
  __shared__ double red[8];
  __shared__ int lab[1024];
  const int k = blockIdx.x, d = threadIdx.x;
  double acc = 0.0;
  int cnt = 0;
  for (int base = 0; base < N; base += 1024) {
    __syncthreads();
    for (int i = d; i < 1024; i += 256) lab[i] = base + i < N ? labels[base + i] : -1;
    __syncthreads();
    const int n = min(1024, N - base);
    for (int i = 0; i < n; ++i)
      if (lab[i] == k) {
        ++cnt;
        if (d < D) acc += x[static_cast<size_t>(base + i) * D + d];
      }
  }
  const double m = cnt > 0 ? acc / static_cast<double>(cnt) : 0.0;
  double q = d < D ? m * m : 0.0;
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((d & 31) == 0) red[d >> 5] = q;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 8; ++w) t += red[w];
  if (d < D) {
    const double v = m / (sqrt(t) + 1e-9);
    if (out64) out64[static_cast<size_t>(k) * D + d] = v;
    if (out32) out32[static_cast<size_t>(k) * D + d] = static_cast<float>(v);
  }
}

// ================================================================================ VAD mask ops
// Block-wide inclusive scan over 1024 threads (int).
__device__ __forceinline__ int block_scan_incl(int v, int* warp_tot /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += n;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const int add = w > 0 ? warp_tot[w - 1] : 0;
  __syncthreads();
  return v + add;
}

// Hysteresis is a scan over state maps {0,1} -> {0,1}: element i maps  not-talking -> (p >= on),
// talking -> !(p < off).  Maps are 2-bit codes (bit0 = image of 0, bit1 = image of 1); composition is
// associative, so a block scans 1024 elements at a time and carries the state across chunks.
__device__ __forceinline__ int map_apply(int m, int s) { return (m >> s) & 1; }
__device__ __forceinline__ int map_then(int f, int g) {   // first f, then g
  return map_apply(g, map_apply(f, 0)) | (map_apply(g, map_apply(f, 1)) << 1);
}

template <typename T>
__global__ void __launch_bounds__(1024)
hysteresis_kernel(const T* __restrict__ probs, int n, double on, double off, uint8_t* __restrict__ mask) {
  __shared__ int wmap[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    int m = 2;  // identity
    if (i < n) {
      const double p = static_cast<double>(probs[i]);
      m = (p >= on ? 1 : 0) | ((p < off) ? 0 : 2);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int f = __shfl_up_sync(0xffffffffu, m, o);
      if (lane >= o) m = map_then(f, m);
    }
    if (lane == 31) wmap[w] = m;
    __syncthreads();
    if (w == 0) {
      int t = wmap[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int f = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t = map_then(f, t);
      }
      wmap[lane] = t;
    }
    __syncthreads();
    const int pre = w > 0 ? map_then(wmap[w - 1], m) : m;
    const int s = map_apply(pre, carry);
    if (i < n) mask[i] = static_cast<uint8_t>(s);
    __syncthreads();
    if (tid == 1023) carry = s;   // (threads past n hold the identity, so s is the last real state)
    __syncthreads();
  }
}

// out[i] = AND (erode) / OR (dilate) of in[i+lo .. i+hi], samples outside [0, n) read as 0
// (scipy.ndimage binary_erosion / binary_dilation with border_value = 0).
__global__ void __launch_bounds__(256)
morph1d_kernel(const uint8_t* __restrict__ in, int n, int lo, int hi, int erode, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int acc = erode ? 1 : 0;
  for (int j = i + lo; j <= i + hi; ++j) {
    const int v = (j >= 0 && j < n) ? (in[j] != 0) : 0;
    if (erode) { acc &= v; if (!acc) break; }
    else { acc |= v; if (acc) break; }
  }
  out[i] = static_cast<uint8_t>(acc);
}

// mask -> runs -> (drop runs shorter than min_speech) -> (merge gaps <= min_gap): one CTA walks the mask in
// 1024-element chunks with block scans; results are frame indices [start, end) (vad.py:121-151).
// seg: [2*max_segs] (start,end pairs); count: number of segments.
__global__ void __launch_bounds__(1024)
mask_segments_kernel(const uint8_t* __restrict__ mask, int n, int min_speech, int min_gap,
                     int32_t* __restrict__ run_s, int32_t* __restrict__ run_e, int32_t* __restrict__ seg,
                     int32_t* __restrict__ count) {
  __shared__ int wt[32];
  __shared__ int n_start, n_end, n_kept, n_seg;
  const int tid = threadIdx.x;
  if (tid == 0) n_start = n_end = n_kept = n_seg = 0;
  __syncthreads();
  // pass 1: run starts (mask[i] && !mask[i-1]) and ends (first index after a run)
  for (int base = 0; base <= n; base += 1024) {
    const int i = base + tid;
    const int cur = (i < n) ? (mask[i] != 0) : 0;
    const int prev = (i > 0 && i <= n) ? (mask[i - 1] != 0) : 0;
    const int fs = (i <= n) && cur && !prev, fe = (i <= n) && !cur && prev;
    const int ps = block_scan_incl(fs, wt);
    const int b0 = n_start;
    if (fs) run_s[b0 + ps - 1] = i;
    __syncthreads();
    if (tid == 1023) n_start = b0 + ps;
    const int pe = block_scan_incl(fe, wt);
    const int b1 = n_end;
    if (fe) run_e[b1 + pe - 1] = i;
    __syncthreads();
    if (tid == 1023) n_end = b1 + pe;
    __syncthreads();
  }
  const int n_runs = n_start;
  // pass 2: keep runs with (end - start) >= min_speech, compacted in place (kept index <= run index)
  for (int base = 0; base < n_runs; base += 1024) {
    const int k = base + tid;
    int s = 0, e = 0, keep = 0;
    if (k < n_runs) { s = run_s[k]; e = run_e[k]; keep = (e - s) >= min_speech; }
    const int p = block_scan_incl(keep, wt);
    const int b0 = n_kept;
    __syncthreads();           // every thread has read its run before any compacted write lands
    if (keep) { run_s[b0 + p - 1] = s; run_e[b0 + p - 1] = e; }
    __syncthreads();
    if (tid == 1023) n_kept = b0 + p;
    __syncthreads();
  }
  const int kept = n_kept;
  // pass 3: a segment starts at kept run k when k == 0 or start[k] - end[k-1] > min_gap; it ends at the
  // end of the last run before the next segment start.
  for (int base = 0; base < kept; base += 1024) {
    const int k = base + tid;
    int first = 0, last = 0;
    if (k < kept) {
      first = (k == 0) || (run_s[k] - run_e[k - 1] > min_gap);
      last = (k == kept - 1) || (run_s[k + 1] - run_e[k] > min_gap);
    }
    const int p = block_scan_incl(first, wt);   // segment index of run k = b0 + p - 1
    const int b0 = n_seg;
    if (first) seg[2 * (b0 + p - 1)] = run_s[k];
    if (last) seg[2 * (b0 + p - 1) + 1] = run_e[k];
    __syncthreads();
    if (tid == 1023) n_seg = b0 + p;
    __syncthreads();
  }
  if (tid == 0) *count = n_seg;
}

}  // namespace

// ================================================================================ C ABI
extern "C" size_t sd_viterbi_workspace_bytes(int T, int K) {
  if (T < 1 || K < 1) return 0;
  const size_t n_chunks = (static_cast<size_t>(T) + VT_CHUNK - 1) / VT_CHUNK;
  return align256((static_cast<size_t>(T) / 4 + 1) * K * 4) + align256(n_chunks * K * 4) + align256(n_chunks * 4) + 256;
}

extern "C" int sd_viterbi_hmm(const void* scores_dev, int scores_f64, int T, int K, float log_stay, float log_move,
                              int32_t* path_dev, void* workspace_dev, void* stream) {
  if (!scores_dev || !path_dev || !workspace_dev || T < 1 || K < 1 || K > 32)
    return fail(SD_ERR_ARG, "sd_viterbi_hmm: bad arguments T=%d K=%d (1 <= K <= 32)", T, K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_chunks = (T + VT_CHUNK - 1) / VT_CHUNK;
  uint8_t* base = static_cast<uint8_t*>(workspace_dev);
  uint32_t* ptr = reinterpret_cast<uint32_t*>(base);
  int32_t* maps = reinterpret_cast<int32_t*>(base + align256((static_cast<size_t>(T) / 4 + 1) * K * 4));
  int32_t* end_state = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(maps) + align256(static_cast<size_t>(n_chunks) * K * 4));
  int* last_state = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(end_state) + align256(static_cast<size_t>(n_chunks) * 4));
  if (scores_f64) launch_viterbi_forward<double>(scores_dev, T, K, log_stay, log_move, ptr, last_state, st);
  else launch_viterbi_forward<float>(scores_dev, T, K, log_stay, log_move, ptr, last_state, st);
  viterbi_chunk_maps_kernel<<<n_chunks, 32, 0, st>>>(ptr, T, K, maps);
  viterbi_chain_kernel<<<1, 256, 0, st>>>(maps, n_chunks, K, last_state, end_state);
  viterbi_walk_kernel<<<n_chunks, 32, 0, st>>>(ptr, T, K, end_state, path_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(4);
  return SD_OK;
}

extern "C" size_t sd_asnorm_workspace_bytes(int nq, int nr, int nc, int D) {
  if (nq < 0 || nr < 0 || nc < 1 || D < 1) return 0;
  const size_t rows = static_cast<size_t>(nq) + nr + nc;
  const size_t blk = static_cast<size_t>(nq < AS_ROW_BLOCK ? (nq > nr ? nq : nr) : AS_ROW_BLOCK);
  return align256(rows * 2 * D * sizeof(__half)) + align256((static_cast<size_t>(nq) + nr) * D * 4) +
         align256((blk > static_cast<size_t>(nr) ? blk : nr) * nc * 4) + 4 * align256((static_cast<size_t>(nq) + nr) * 4) + 512;
}

extern "C" int sd_asnorm_scores(const float* q_dev, const float* r_dev, const float* c_dev, int nq, int nr, int nc,
                                int D, int topk, float* out_dev, void* workspace_dev, void* stream) {
  if (!q_dev || !r_dev || !c_dev || !out_dev || !workspace_dev || nq < 0 || nr < 1 || nc < 1 || D < 64 || D % 64 ||
      D > 512 || topk < 1)
    return fail(SD_ERR_ARG, "sd_asnorm_scores: bad arguments nq=%d nr=%d nc=%d D=%d topk=%d", nq, nr, nc, D, topk);
  if (nq == 0) return SD_OK;
  if (topk > nc) topk = nc;   // diar_diag.py:201  [-min(topk, C.shape[0]):]
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  __half* qs = reinterpret_cast<__half*>(w);
  __half* rs = qs + static_cast<size_t>(nq) * 2 * D;
  __half* cs = rs + static_cast<size_t>(nr) * 2 * D;
  w += align256((static_cast<size_t>(nq) + nr + nc) * 2 * D * sizeof(__half));
  float* qn = reinterpret_cast<float*>(w);
  float* rn = qn + static_cast<size_t>(nq) * D;
  w += align256((static_cast<size_t>(nq) + nr) * D * 4);
  const int blk = nq < AS_ROW_BLOCK ? (nq > nr ? nq : nr) : AS_ROW_BLOCK;
  float* dist = reinterpret_cast<float*>(w);
  w += align256(static_cast<size_t>(blk > nr ? blk : nr) * nc * 4);
  float* q_mu = reinterpret_cast<float*>(w);
  float* q_sig = reinterpret_cast<float*>(w + align256((static_cast<size_t>(nq) + nr) * 4));
  float* r_mu = q_mu + nq;
  float* r_sig = q_sig + nq;

  l2n_split_kernel<<<(nq + 7) / 8, 256, 0, st>>>(q_dev, nq, D, 1e-9f, qs, qn);
  l2n_split_kernel<<<(nr + 7) / 8, 256, 0, st>>>(r_dev, nr, D, 1e-9f, rs, rn);
  l2n_split_kernel<<<(nc + 7) / 8, 256, 0, st>>>(c_dev, nc, D, 1e-9f, cs, nullptr);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(3);

  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int max_stage = 200 * 1024;
  if (!attr[dev & 63]) {
    SD_CUDA_OK(cudaFuncSetAttribute(cohort_topk_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stage));
    attr[dev & 63] = true;
  }
  const int stage_elems = static_cast<size_t>(nc) * 4 <= static_cast<size_t>(max_stage) ? nc : 0;
  const size_t smem = static_cast<size_t>(stage_elems) * 4;
  for (int r0 = 0; r0 < nq; r0 += blk) {
    const int rows = nq - r0 < blk ? nq - r0 : blk;
    SD_TRY(cohort_distances(qs, nq, r0, rows, cs, nc, D, dist, st));
    cohort_topk_stats_kernel<<<rows, TK_THREADS, smem, st>>>(dist, nc, nc, topk, stage_elems, q_mu + r0, q_sig + r0);
    SD_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  SD_TRY(cohort_distances(rs, nr, 0, nr, cs, nc, D, dist, st));
  cohort_topk_stats_kernel<<<nr, TK_THREADS, smem, st>>>(dist, nc, nc, topk, stage_elems, r_mu, r_sig);
  int grid = (nq + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  asnorm_combine_kernel<<<grid, 256, 0, st>>>(qn, rn, nq, nr, D, q_mu, q_sig, r_mu, r_sig, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return SD_OK;
}

extern "C" int sd_hysteresis_u8(const void* probs_dev, int probs_f64, int n, double on, double off,
                                uint8_t* mask_dev, void* stream) {
  if (n < 0 || (n > 0 && (!probs_dev || !mask_dev))) return fail(SD_ERR_ARG, "sd_hysteresis_u8: bad arguments");
  if (n == 0) return SD_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (probs_f64) hysteresis_kernel<double><<<1, 1024, 0, st>>>(static_cast<const double*>(probs_dev), n, on, off, mask_dev);
  else hysteresis_kernel<float><<<1, 1024, 0, st>>>(static_cast<const float*>(probs_dev), n, on, off, mask_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_morph_open_close_u8(const uint8_t* mask_dev, int n, int open_w, int close_w, uint8_t* out_dev,
                                      uint8_t* tmp_dev, void* stream) {
  if (n < 0 || open_w < 0 || close_w < 0 || (n > 0 && (!mask_dev || !out_dev || !tmp_dev)))
    return fail(SD_ERR_ARG, "sd_morph_open_close_u8: bad arguments");
  if (n == 0) return SD_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (n + 255) / 256;
  // scipy, 1-D structure of w ones, origin 0: erosion reads in[i - w/2 .. i - w/2 + w - 1];
  // dilation (reflected structure, origin shifted by one for even w) reads in[i + w/2 - w + 1 .. i + w/2].
  const uint8_t* cur = mask_dev;
  uint8_t* bufs[2] = {out_dev, tmp_dev};
  int which = 0, launches = 0;
  auto pass = [&](int w, int erode) {
    const int lo = erode ? -(w / 2) : (w / 2) - w + 1;
    const int hi = lo + w - 1;
    morph1d_kernel<<<grid, 256, 0, st>>>(cur, n, lo, hi, erode, bufs[which]);
    cur = bufs[which];
    which ^= 1;
    ++launches;
  };
  if (open_w > 0) { pass(open_w, 1); pass(open_w, 0); }
  if (close_w > 0) { pass(close_w, 0); pass(close_w, 1); }
  if (cur != out_dev) {
    if (cur == mask_dev) SD_CUDA_OK(cudaMemcpyAsync(out_dev, mask_dev, n, cudaMemcpyDeviceToDevice, st));
    else SD_CUDA_OK(cudaMemcpyAsync(out_dev, cur, n, cudaMemcpyDeviceToDevice, st));
  }
  SD_CUDA_OK(cudaGetLastError());
  count_launch(launches);
  return SD_OK;
}

extern "C" size_t sd_mask_segments_workspace_bytes(int n) {
  return n < 0 ? 0 : 2 * align256((static_cast<size_t>(n) / 2 + 2) * 4) + 256;
}

extern "C" int sd_mask_segments_i32(const uint8_t* mask_dev, int n, int min_speech_frames, int min_gap_frames,
                                    int32_t* seg_dev, int32_t* count_dev, void* workspace_dev, void* stream) {
  if (n < 0 || !seg_dev || !count_dev || !workspace_dev || (n > 0 && !mask_dev))
    return fail(SD_ERR_ARG, "sd_mask_segments_i32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* run_s = static_cast<int32_t*>(workspace_dev);
  int32_t* run_e = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(workspace_dev) + align256((static_cast<size_t>(n) / 2 + 2) * 4));
  mask_segments_kernel<<<1, 1024, 0, st>>>(mask_dev, n, min_speech_frames, min_gap_frames, run_s, run_e, seg_dev, count_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_cluster_centers_f64(const double* x_dev, const int32_t* labels_dev, int N, int D, int K,
                                      double* out_f64_dev, float* out_f32_dev, void* stream) {
  if (!x_dev || !labels_dev || (!out_f64_dev && !out_f32_dev) || N < 1 || D < 1 || D > 256 || K < 1)
    return fail(SD_ERR_ARG, "sd_cluster_centers_f64: bad arguments N=%d D=%d K=%d", N, D, K);
  cluster_centers_kernel<<<K, 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, labels_dev, N, D, out_f64_dev, out_f32_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_dot_scores(const float* x_dev, const float* cent_dev, int N, int K, int D, float* out_dev, void* stream) {
  if (!x_dev || !cent_dev || !out_dev || N < 0 || K < 1 || D < 1) return fail(SD_ERR_ARG, "sd_dot_scores: bad arguments");
  if (N == 0) return SD_OK;
  int grid = (N + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  dot_scores_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, cent_dev, N, K, D, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}
