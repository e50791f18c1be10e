// ahc.cu — average-linkage agglomerative clustering with a distance threshold on the GPU
// (SURVEY.md §2.1 K11).
//
// Replaces  AgglomerativeClustering(n_clusters=None, linkage="average", metric="precomputed",
//           distance_threshold=1-cos_thr).fit_predict(D)      (/root/reference/diar_diag.py:221-226),
// which runs scipy's single-threaded nn_chain in float64 on the f32 distance matrix.
//
// Algorithm: average linkage (UPGMA) is REDUCIBLE, so every pair of reciprocal nearest
// neighbours (RNN) is a merge of the final dendrogram at exactly its current distance, no
// matter what is merged elsewhere first.  Instead of N-1 dependent merges we therefore run
// ROUNDS; each round
//   A  recomputes the nearest neighbour (argmin over the row) of every "dirty" cluster,
//   B  collects all RNN pairs whose distance is < threshold,
//   C  applies the Lance-Williams update  d(k, i+j) = (n_i d(k,i) + n_j d(k,j)) / (n_i + n_j)
//      for all those pairs at once (rows first, then the pair x pair corners, then the
//      mirrored columns, which keeps the matrix exactly symmetric),
//   D  retires the absorbed clusters and marks as dirty only the merged rows and the rows
//      whose cached nearest neighbour was merged (reducibility keeps every other cache valid).
// It stops when no RNN pair is below the threshold, which — average linkage being monotone —
// is the flat clustering sklearn cuts at `distance_threshold`.  The whole loop is ONE
// persistent cooperative kernel (grid-wide barriers between phases — three per round: A | B | C+D — and no
// host round trips).
// Arithmetic is f64 on the f32-rounded input, as in scipy.  The matrix is HBM-resident
// (8 N^2 bytes: 3.2 GB at N = 20k, 20 GB at 50k).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include "sd_ptx.cuh"
#include "sd_status.h"

namespace cg = cooperative_groups;
using namespace sd;

namespace {

constexpr int AHC_THREADS = 256;

struct AhcState {
  void* D;          // [N, N] working matrix, f64 or f32 (ahc_rounds_kernel<U, MT>)
  double* nn_dist;  // [N]
  int* nn_idx;      // [N]
  int* size;        // [N]
  int* active;      // [N] 1 = live cluster
  int* parent;      // [N] merge forest (root = smallest member)
  int* role;        // [N] -1 none, else 2*pair (absorbing i) / 2*pair+1 (absorbed j) in this round
  int* dirty_list;  // [N]
  int* pair_i;      // [N/2]
  int* pair_j;      // [N/2]
  int* pair_ni;     // [N/2] sizes of the two clusters when the pair was found (phase C reads these, phase D updates `size`)
  int* pair_nj;     // [N/2]
  int* act_list;    // [2][N] compact list of live clusters (ping-pong)
  int* counters;    // [0],[1] n_dirty ping-pong [2] n_pairs [3] rounds [4] merges [5] list length [6] list index
                    // [7] live clusters [8] staged new list length
  long long* prof;  // debug (SD_AHC_PROF): ns per phase as seen by thread 0 of CTA 0, printed at the end
  int N;
  double thr;
};

__device__ __forceinline__ long long ahc_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// f32 [N,N] -> working matrix [N,N] (f64 or f32), symmetrised from the upper triangle (sklearn reads D[i,j], i<j).
template <typename MT>
__global__ void __launch_bounds__(256)
ahc_init_matrix_kernel(const float* __restrict__ dist, int N, MT* __restrict__ D) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < N && j < N) ? dist[static_cast<size_t>(i) * N + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < N && j < N && (bi < bj || i <= j)) D[static_cast<size_t>(i) * N + j] = tile[r][tx];
    // mirrored element: row j-block, column i-block
    const int mi = bj * 32 + r, mj = bi * 32 + tx;  // D[mi, mj] = tile[tx][r]
    if (mi < N && mj < N && (bi < bj || mj < mi)) D[static_cast<size_t>(mi) * N + mj] = tile[tx][r];
  }
}

__global__ void ahc_init_state_kernel(AhcState S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S.N) {
    S.size[i] = 1;
    S.active[i] = 1;
    S.parent[i] = i;
    S.role[i] = -1;
    S.dirty_list[i] = i;
    S.nn_idx[i] = -1;
    S.nn_dist[i] = 1e300;
    S.act_list[i] = i;
  }
  if (i == 0) {
    S.counters[0] = S.N;
    S.counters[1] = 0;
    S.counters[2] = 0;
    S.counters[3] = 0;
    S.counters[4] = 0;
    S.counters[5] = S.N;
    S.counters[6] = 0;
    S.counters[7] = S.N;
    S.counters[8] = 0;
    for (int q = 9; q < 64; ++q) S.counters[q] = 0;   // [12] grid barrier arrivals; phase timers (AhcState::prof)
  }
}

__device__ __forceinline__ void argmin_combine(double& d, int& i, double od, int oi) {
  if (od < d || (od == d && oi < i)) { d = od; i = oi; }
}

// AHC_U = independent row entries in flight per thread (phases A, C1, C3).  MT = storage type of the working matrix:
// double (default: scipy's arithmetic exactly) or float (SD_AHC_F32=1: every Lance-Williams update is still computed in
// f64 but STORED rounded to f32 — half the bytes per round; merge heights then differ from scipy's by ~1e-7 relative,
// which can only matter for ties closer than that)
template <int AHC_U, typename MT>
__global__ void __launch_bounds__(AHC_THREADS)
ahc_rounds_kernel(AhcState S) {
  MT* const DM = static_cast<MT*>(S.D);
  cg::grid_group grid = cg::this_grid();
  __shared__ double red_d[AHC_THREADS / 32];
  __shared__ int red_i[AHC_THREADS / 32];
  __shared__ int scan_s[AHC_THREADS];
  const int N = S.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = blockIdx.x * AHC_THREADS + tid;
  const int gthreads = gridDim.x * AHC_THREADS;
  // (a hand-rolled one-arrival-per-CTA barrier instead of grid.sync() was measured: 24.0 vs 24.05 ms at N = 20k)
  auto grid_sync = [&]() { grid.sync(); };
  int cur = 0;  // which dirty counter is current
  long long t_prev = S.prof ? ahc_now() : 0;
  auto stamp = [&](int phase) {
    if (S.prof != nullptr && gtid == 0) {
      const long long t = ahc_now();
      S.prof[phase] += t - t_prev;
      t_prev = t;
    }
  };

  // every thread tracks these identically (all decisions below are grid-uniform), so no phase is needed to publish them
  int list_idx = 0, n_list = N, alive = N, prev_pairs = 0;
  bool rebuilt = false;

  for (int round = 0; round < N; ++round) {
    // the compact list of (mostly) live clusters: late rounds touch a few hundred columns, not N
    if (rebuilt) { list_idx ^= 1; n_list = S.counters[8]; }   // block 0 staged the new list in the previous round
    const int* list = S.act_list + static_cast<size_t>(list_idx) * N;
    // ---- A: nearest neighbour of every dirty row (one CTA per row)
    const int n_dirty = S.counters[cur];
    // roles of the previous round's pairs: their last readers (C3 / D) are behind the barrier that ended that round,
    // the next writer (B) and readers (C3 / D) are behind the barrier that ends this phase
    for (int p = gtid; p < prev_pairs; p += gthreads) {
      S.role[S.pair_i[p]] = -1;
      S.role[S.pair_j[p]] = -1;
    }
    // the pair counter of the previous round: every thread read it right after that round's B barrier, and a whole
    // barrier (the one that ended the round) lies between that read and this reset; B of this round, which counts
    // again, is behind the barrier that ends this phase.  (Resetting it at the end of the round, as before phase C3
    // lost its barrier, let a slow CTA read 0 and leave the loop: wrong partitions at N = 50k.)
    if (gtid == 0) S.counters[2] = 0;
    for (int q = blockIdx.x; q < n_dirty; q += gridDim.x) {
      const int r = S.dirty_list[q];
      const MT* row = DM + static_cast<size_t>(r) * N;
      double bd = 1e300;
      int bi = 0x7fffffff;
      // AHC_U independent list -> (distance, liveness) chains in flight per thread: one entry at a time is a chain
      // of three dependent loads (list, active, the DRAM-resident row) per iteration, n_list / 256 iterations long
      for (int i0 = tid; i0 < n_list; i0 += AHC_THREADS * AHC_U) {
        int k[AHC_U], a[AHC_U];
        double v[AHC_U];
#pragma unroll
        for (int u = 0; u < AHC_U; ++u) {
          const int i = i0 + u * AHC_THREADS;
          k[u] = i < n_list ? list[i] : -1;
        }
#pragma unroll
        for (int u = 0; u < AHC_U; ++u) {
          v[u] = k[u] >= 0 ? row[k[u]] : 1e300;
          a[u] = k[u] >= 0 ? S.active[k[u]] : 0;
        }
#pragma unroll
        for (int u = 0; u < AHC_U; ++u)
          if (a[u] && k[u] != r) argmin_combine(bd, bi, v[u], k[u]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmin_combine(bd, bi, od, oi);
      }
      if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < AHC_THREADS / 32; ++w) argmin_combine(bd, bi, red_d[w], red_i[w]);
        S.nn_dist[r] = bd;
        S.nn_idx[r] = bi == 0x7fffffff ? -1 : bi;
      }
      __syncthreads();
    }
    grid_sync();
    stamp(0);

    // ---- B: reciprocal nearest neighbours below the threshold
    if (gtid == 0) S.counters[cur] = 0;  // every thread has read n_dirty; this counter collects the round after next's
    for (int i = gtid; i < n_list; i += gthreads) {
      const int r = list[i];
      if (!S.active[r]) continue;
      const int j = S.nn_idx[r];
      if (j > r && S.nn_dist[r] < S.thr && S.nn_idx[j] == r) {
        const int p = atomicAdd(&S.counters[2], 1);
        S.pair_i[p] = r;
        S.pair_j[p] = j;
        S.pair_ni[p] = S.size[r];
        S.pair_nj[p] = S.size[j];
        S.role[r] = 2 * p;
        S.role[j] = 2 * p + 1;
      }
    }
    grid_sync();
    stamp(1);
    const int n_pairs = S.counters[2];
    if (n_pairs == 0) break;

    // ---- C: the Lance-Williams update of every pair, rows, corners and mirrored column by the SAME CTA, no grid
    // barrier in between (it used to be rows + upper corners | barrier | mirror).  Who writes what:
    //   row i_p, columns k that do not merge this round ........ CTA p (C1), then mirrored into D[k, i_p] (C3)
    //   corner (i_p, i_q), i_p < i_q .............................. CTA p: C1 on columns i_q and j_q, C2 combines them,
    //                                                              and CTA p also writes the mirror image D[i_q, i_p]
    //   row i_p, columns i_q < i_p ................................ never touched by CTA p (CTA q writes them)
    // so no entry has two writers, and every value a CTA combines comes from rows i_p / j_p, which only CTA p writes.
    // The cluster sizes are the snapshot phase B took (pair_ni / pair_nj): phase D below updates `size` concurrently.
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const int i = S.pair_i[p], j = S.pair_j[p];
      const double ni = S.pair_ni[p], nj = S.pair_nj[p], inv = ni + nj;
      MT* ri = DM + static_cast<size_t>(i) * N;
      const MT* rj = DM + static_cast<size_t>(j) * N;
      for (int q0 = tid; q0 < n_list; q0 += AHC_THREADS * AHC_U) {
        int k[AHC_U], rk[AHC_U], act[AHC_U];
        double a[AHC_U], b[AHC_U];
#pragma unroll
        for (int u = 0; u < AHC_U; ++u) {
          const int q = q0 + u * AHC_THREADS;
          k[u] = q < n_list ? list[q] : -1;
        }
#pragma unroll
        for (int u = 0; u < AHC_U; ++u) {
          act[u] = 0;
          rk[u] = -1;
          a[u] = b[u] = 0.0;
          if (k[u] >= 0 && k[u] != i) { act[u] = S.active[k[u]]; rk[u] = S.role[k[u]]; a[u] = ri[k[u]]; b[u] = rj[k[u]]; }
        }
#pragma unroll
        for (int u = 0; u < AHC_U; ++u) {
          // live columns; a column that merges this round counts as live whatever `active` says (phase D of another
          // CTA may already have retired an absorbed cluster whose entry C2 below still needs)
          if (!(act[u] || rk[u] >= 0) || k[u] < 0 || k[u] == i) continue;
          if (rk[u] >= 0 && !(rk[u] & 1) && k[u] < i) continue;   // corner owned by the CTA of row k
          const MT v = static_cast<MT>((ni * a[u] + nj * b[u]) / inv);
          ri[k[u]] = v;
          if (rk[u] < 0) DM[static_cast<size_t>(k[u]) * N + i] = v;   // C3: column i of a row that does not merge
        }
      }
      __syncthreads();
      // C2: corners with the pairs of larger row index, combined from this row's two freshly updated entries,
      // written to both triangles
      for (int q = tid; q < n_pairs; q += AHC_THREADS) {
        const int iq = S.pair_i[q], jq = S.pair_j[q];
        if (i < iq) {
          const double nq = S.pair_ni[q], mq = S.pair_nj[q];
          const MT v = static_cast<MT>((nq * static_cast<double>(ri[iq]) + mq * static_cast<double>(ri[jq])) / (nq + mq));
          ri[iq] = v;
          DM[static_cast<size_t>(iq) * N + i] = v;
        }
      }
      __syncthreads();
    }
    stamp(2);
    // ---- D (same phase: touches size/parent/dirty only): retire absorbed clusters, mark dirty rows
    const int nxt = cur ^ 1;
    for (int q = gtid; q < n_list; q += gthreads) {
      const int r = list[q];
      if (!S.active[r]) continue;
      const int rr = S.role[r];
      if (rr >= 0 && (rr & 1)) continue;  // absorbed: handled by its partner below
      bool dirty = false;
      if (rr >= 0) {
        const int j = S.pair_j[rr >> 1];
        S.size[r] += S.size[j];
        S.parent[j] = r;
        S.active[j] = 0;   // benign for this phase's other readers: they skip j through its role as well
        dirty = true;
      } else {
        const int nn = S.nn_idx[r];
        dirty = nn >= 0 && S.role[nn] >= 0;
      }
      if (dirty) S.dirty_list[atomicAdd(&S.counters[nxt], 1)] = r;
    }
    // ---- compaction of the live list (block 0, only when a quarter of it is dead)
    const int n_alive_after = alive - n_pairs;
    const bool rebuild = (n_list - n_alive_after) * 4 > n_list;
    if (rebuild && blockIdx.x == 0) {
      int* out = S.act_list + static_cast<size_t>(list_idx ^ 1) * N;
      const int chunk = (n_list + AHC_THREADS - 1) / AHC_THREADS;
      const int lo = tid * chunk, hi = min(n_list, lo + chunk);
      int cnt = 0;
      for (int q = lo; q < hi; ++q) {
        const int k = list[q];
        const int rk = S.role[k];
        cnt += (S.active[k] && !(rk >= 0 && (rk & 1))) ? 1 : 0;
      }
      scan_s[tid] = cnt;
      __syncthreads();
      for (int o = 1; o < AHC_THREADS; o <<= 1) {
        const int v = tid >= o ? scan_s[tid - o] : 0;
        __syncthreads();
        scan_s[tid] += v;
        __syncthreads();
      }
      int w = scan_s[tid] - cnt;
      for (int q = lo; q < hi; ++q) {
        const int k = list[q];
        const int rk = S.role[k];
        if (S.active[k] && !(rk >= 0 && (rk & 1))) out[w++] = k;
      }
      if (tid == AHC_THREADS - 1) S.counters[8] = scan_s[tid];  // new length, picked up by every thread at the top of the next round
    }
    if (gtid == 0) {   // statistics for sd_ahc_read_stats (the pair counter itself is reset in the next round's phase A)
      S.counters[3] = round + 1;
      S.counters[4] += n_pairs;
      S.counters[7] = n_alive_after;
    }
    alive = n_alive_after;
    prev_pairs = n_pairs;
    rebuilt = rebuild;
    cur = nxt;
    grid_sync();
    stamp(4);
    if (S.prof != nullptr && gtid == 0 && (round == 9 || round == 29 || round == 99)) {
      const int c = round == 9 ? 0 : round == 29 ? 1 : 2;   // cumulative snapshots: where the time goes over the rounds
      for (int q = 0; q < 6; ++q) S.prof[6 + c * 6 + q] = S.prof[q];
    }
  }
  if (S.prof != nullptr && gtid == 0) {
    for (int c = 0; c < 3; ++c)
      printf("  after %3d rounds: A %.0f  B %.0f  C1 %.0f  C2 %.0f  C3+D %.0f  cleanup %.0f\n", c == 0 ? 10 : c == 1 ? 30 : 100,
             S.prof[6 + c * 6] * 1e-3, S.prof[7 + c * 6] * 1e-3, S.prof[8 + c * 6] * 1e-3, S.prof[9 + c * 6] * 1e-3,
             S.prof[10 + c * 6] * 1e-3, S.prof[11 + c * 6] * 1e-3);
  }
  if (S.prof != nullptr && gtid == 0)
    printf("ahc phases us: A %.0f  B %.0f  C1 %.0f  C2 %.0f  C3+D %.0f  cleanup %.0f  (grid %d CTAs)\n", S.prof[0] * 1e-3,
           S.prof[1] * 1e-3, S.prof[2] * 1e-3, S.prof[3] * 1e-3, S.prof[4] * 1e-3, S.prof[5] * 1e-3, gridDim.x);
}

// labels[x] = rank of root(x) among the roots in ascending order (root = smallest member).
__global__ void __launch_bounds__(1024)
ahc_labels_kernel(AhcState S, int* __restrict__ labels, int* __restrict__ n_clusters, int* scratch) {
  __shared__ int part[1024];
  const int N = S.N, tid = threadIdx.x;
  const int chunk = (N + 1023) / 1024;
  const int lo = tid * chunk, hi = min(N, lo + chunk);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += S.active[i] ? 1 : 0;
  part[tid] = cnt;
  __syncthreads();
  // inclusive scan (Hillis-Steele) over 1024 partial counts
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int base = part[tid] - cnt;
  for (int i = lo; i < hi; ++i) {
    scratch[i] = base;
    base += S.active[i] ? 1 : 0;
  }
  if (tid == 1023) *n_clusters = part[1023];
  __syncthreads();
  for (int x = tid; x < N; x += 1024) {
    int r = x;
    while (S.parent[r] != r) r = S.parent[r];
    labels[x] = scratch[r];
  }
}

struct Layout {
  size_t off_D, off_nn_dist, off_ints, total;
};
Layout ahc_layout(int N) {
  Layout L;
  size_t o = 0;
  L.off_D = o;
  o += static_cast<size_t>(N) * N * 8;
  L.off_nn_dist = o;
  o += static_cast<size_t>(N) * 8;
  L.off_ints = o;
  o += (static_cast<size_t>(N) * 10 + 64) * 4;  // 7 int arrays + scratch + counters + 2 live lists
  L.total = o + 256;
  return L;
}

}  // namespace

extern "C" size_t sd_ahc_workspace_bytes(int N) { return N < 1 ? 0 : ahc_layout(N).total; }

extern "C" int sd_ahc_average_f32(const float* dist_dev, int N, double threshold, int32_t* labels_dev,
                                  int32_t* n_clusters_dev, void* workspace_dev, void* stream) {
  if (!dist_dev || !labels_dev || !n_clusters_dev || !workspace_dev || N < 1)
    return fail(SD_ERR_ARG, "sd_ahc_average_f32: bad arguments (N=%d)", N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  const Layout L = ahc_layout(N);
  AhcState S;
  S.N = N;
  S.thr = threshold;  // f64, as sklearn compares the f64 linkage distances with a Python float
  S.D = base + L.off_D;
  S.nn_dist = reinterpret_cast<double*>(base + L.off_nn_dist);
  int* ip = reinterpret_cast<int*>(base + L.off_ints);
  S.nn_idx = ip;
  S.size = ip + N;
  S.active = ip + 2 * static_cast<size_t>(N);
  S.parent = ip + 3 * static_cast<size_t>(N);
  S.role = ip + 4 * static_cast<size_t>(N);
  S.dirty_list = ip + 5 * static_cast<size_t>(N);
  S.pair_i = ip + 6 * static_cast<size_t>(N);
  S.pair_j = S.pair_i + (N + 1) / 2;  // at most floor(N/2) pairs per round
  int* scratch = ip + 7 * static_cast<size_t>(N);
  S.pair_ni = scratch;                 // the label kernel's scratch is free while the rounds run
  S.pair_nj = scratch + (N + 1) / 2;
  S.counters = ip + 8 * static_cast<size_t>(N);
  S.act_list = ip + 8 * static_cast<size_t>(N) + 64;
  S.prof = getenv("SD_AHC_PROF") ? reinterpret_cast<long long*>(S.counters + 16) : nullptr;

  // SD_AHC_F32=1: f32 working matrix (see ahc_rounds_kernel)
  static const bool f32_store = [] { const char* e = getenv("SD_AHC_F32"); return e && atoi(e) != 0; }();
  const int nb = (N + 31) / 32;
  if (f32_store) ahc_init_matrix_kernel<float><<<dim3(nb, nb), 256, 0, st>>>(dist_dev, N, static_cast<float*>(S.D));
  else ahc_init_matrix_kernel<double><<<dim3(nb, nb), 256, 0, st>>>(dist_dev, N, static_cast<double*>(S.D));
  ahc_init_state_kernel<<<(N + 255) / 256, 256, 0, st>>>(S);
  SD_CUDA_OK(cudaGetLastError());

  int dev = 0, sms = 0, occ = 0;
  SD_CUDA_OK(cudaGetDevice(&dev));
  SD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // SD_AHC_U = 1 | 4 | 8 (A/B switch): row entries in flight per thread; more means more registers, fewer CTAs
  // (64 / 80 / 128 registers -> 4 / 3 / 2 CTAs per SM).  Measured at N = 5k / 20k / 50k: U = 1 4.06 / 33.6 / 206 ms,
  // U = 4 3.37 / 25.5 / 149 ms, U = 8 3.48 / 27.3 / 160 ms.
  static const int unroll = [] { const char* e = getenv("SD_AHC_U"); return e ? atoi(e) : 4; }();
  void (*kern)(AhcState) = f32_store ? (unroll <= 1 ? ahc_rounds_kernel<1, float> : unroll <= 4 ? ahc_rounds_kernel<4, float> : ahc_rounds_kernel<8, float>)
                                     : (unroll <= 1 ? ahc_rounds_kernel<1, double> : unroll <= 4 ? ahc_rounds_kernel<4, double> : ahc_rounds_kernel<8, double>);
  SD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, AHC_THREADS, 0));
  if (occ < 1) return fail(SD_ERR_CUDA, "ahc_rounds_kernel cannot be made resident");
  if (occ > 4) occ = 4;
  void* args[] = {&S};
  SD_CUDA_OK(cudaLaunchCooperativeKernel((void*)kern, dim3(sms * occ), dim3(AHC_THREADS), args, 0, st));
  ahc_labels_kernel<<<1, 1024, 0, st>>>(S, labels_dev, n_clusters_dev, scratch);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(4);
  return SD_OK;
}

extern "C" int sd_ahc_read_stats(const void* workspace_dev, int N, int32_t* rounds, int32_t* merges) {
  if (!workspace_dev || N < 1 || !rounds || !merges) return fail(SD_ERR_ARG, "sd_ahc_read_stats: bad arguments");
  const uint8_t* base =
      reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  const Layout L = ahc_layout(N);
  const int* counters = reinterpret_cast<const int*>(base + L.off_ints) + 8 * static_cast<size_t>(N);
  int h[5];
  SD_CUDA_OK(cudaMemcpy(h, counters, sizeof(h), cudaMemcpyDeviceToHost));
  *rounds = h[3];
  *merges = h[4];
  return SD_OK;
}
