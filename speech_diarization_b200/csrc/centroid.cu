// centroid.cu — centroid-linkage agglomerative clustering on the GPU.
//
// Replaces  scipy.cluster.hierarchy.linkage(X, method="centroid", metric="euclidean")  as used by
// pyannote.audio's AgglomerativeClustering, which is what /root/reference/diarization_baseline.py:176-180,
// 252-257 runs (`pipeline.clustering.threshold = 0.70`, method "centroid" in the 3.1 pipeline config;
// SURVEY Appendix B, recalled — pyannote is not on disk).
//
// Centroid linkage is NOT reducible (merges can invert), so the reciprocal-nearest-neighbour rounds of
// ahc.cu do not apply: the N-1 merges are taken strictly in order of the global minimum distance.  The
// A cluster is its f64 centroid and size, d(a, b) = ||c_a - c_b|| is recomputed from the centroids (no
// Lance-Williams drift) and cached in an N x N f64 matrix so that nearest-neighbour rescans are row scans;
// each cluster caches its nearest neighbour.  One persistent cooperative kernel runs all N-1 steps:
//   S1  block-partial arg-min over the nearest-neighbour cache
//   S2  block 0 reduces the partials to the global pair (i, j) and merges it
//       (c_i <- (n_i c_i + n_j c_j) / (n_i + n_j), records the scipy linkage row, publishes (i, j))
//   S3  every live cluster k computes d(k, new) (one warp each), stores it in the matrix and takes it as its
//       neighbour if closer; clusters whose cached neighbour was i or j become "orphans" (in clustered,
//       high-dimensional data one point is often the neighbour of many); block-partial minima for row `new`
//   S4  orphans rescan their matrix row (one CTA each); block 0 finishes the merged cluster's neighbour
// The output is scipy's (N-1) x 4 linkage matrix; the flat cut (fcluster "distance", inversion-safe) and
// pyannote's small-cluster reassignment are O(N) host code (speech_diarization_b200/diarization_baseline.py).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "sd_ptx.cuh"
#include "sd_status.h"

namespace cg = cooperative_groups;
using namespace sd;

namespace {

constexpr int CL_THREADS = 256;

struct ClState {
  double* Dm;       // [N, N] cached centroid distances (rows/columns of dead clusters are stale)
  double* C;        // [N, D] centroids (slot of a merged cluster = slot of its lower-indexed member)
  double* nn_dist;  // [N]
  double* part_d;   // [grid] block partial minima (S1) ; [grid .. 2 grid) partial minima of the new row (S3)
  double* Z;        // [N-1, 4] output
  int* nn_idx;      // [N]
  int* size;        // [N]
  int* active;      // [N]
  int* cid;         // [N] scipy cluster id currently held by the slot
  int* part_i;      // [2 grid]
  int* orphans;     // [N]
  int* counters;    // [0] n_orphans [1] merged slot i [2] absorbed slot j
  int N, D;
};

__device__ __forceinline__ void amin(double& d, int& i, double od, int oi) {
  if (od < d || (od == d && oi < i)) { d = od; i = oi; }
}

// squared distance between two centroids, lanes of one warp stride over D
__device__ __forceinline__ double warp_dist2(const double* a, const double* b, int D, int lane) {
  double s = 0.0;
  for (int c = lane; c < D; c += 32) {
    const double t = a[c] - b[c];
    s = fma(t, t, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

__global__ void cl_init_kernel(ClState S, const float* __restrict__ x) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long>(S.N) * S.D) S.C[i] = static_cast<double>(x[i]);
  if (i < S.N) {
    S.size[i] = 1;
    S.active[i] = 1;
    S.cid[i] = static_cast<int>(i);
    S.nn_idx[i] = -1;
    S.nn_dist[i] = 1e300;
    S.orphans[i] = static_cast<int>(i);  // every point needs its first nearest neighbour
  }
  if (i == 0) S.counters[0] = S.N;
}

// nearest live neighbour of cluster r from its cached distance row: one CTA
__device__ void rescan(const ClState& S, int r, double* red_d, int* red_i) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* row = S.Dm + static_cast<size_t>(r) * S.N;
  double bd = 1e300;
  int bi = 0x7fffffff;
  // eight independent (distance, liveness) loads in flight per thread: the row comes from DRAM (the matrix is
  // 8 N^2 bytes) and a load-test-load loop of N / 256 dependent round trips made one rescan ~80 us at N = 20k —
  // the longest pole of a step, since a step has only a handful of orphans and each is one CTA's job
  constexpr int RU = 8;
  for (int k0 = tid; k0 < S.N; k0 += CL_THREADS * RU) {
    double v[RU];
    int a[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int k = k0 + u * CL_THREADS;
      const bool in = k < S.N;
      v[u] = in ? row[k] : 1e300;
      a[u] = in ? S.active[k] : 0;
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int k = k0 + u * CL_THREADS;
      if (a[u] && k != r) amin(bd, bi, v[u], k);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    amin(bd, bi, od, oi);
  }
  if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < CL_THREADS / 32; ++w) amin(bd, bi, red_d[w], red_i[w]);
    S.nn_dist[r] = bi == 0x7fffffff ? 1e300 : bd;
    S.nn_idx[r] = bi == 0x7fffffff ? -1 : bi;
  }
  __syncthreads();
}

// all pairwise centroid distances of the initial points: 16 x 16 outputs per CTA, D staged through smem
__global__ void __launch_bounds__(256)
cl_pairwise_kernel(ClState S) {
  __shared__ double sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int c0 = 0; c0 < S.D; c0 += 16) {
    const int ia = blockIdx.y * 16 + ty, jb = blockIdx.x * 16 + ty;   // rows loaded by this thread
    sa[ty][tx] = (ia < S.N && c0 + tx < S.D) ? S.C[static_cast<size_t>(ia) * S.D + c0 + tx] : 0.0;
    sb[ty][tx] = (jb < S.N && c0 + tx < S.D) ? S.C[static_cast<size_t>(jb) * S.D + c0 + tx] : 0.0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const double t = sa[ty][c] - sb[tx][c];
      acc = fma(t, t, acc);
    }
    __syncthreads();
  }
  if (i < S.N && j < S.N) S.Dm[static_cast<size_t>(i) * S.N + j] = sqrt(acc);
}

__global__ void __launch_bounds__(CL_THREADS)
cl_steps_kernel(ClState S) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red_d[CL_THREADS / 32];
  __shared__ int red_i[CL_THREADS / 32];
  __shared__ double sh_d;
  __shared__ int sh_i, sh_j;
  const int N = S.N, D = S.D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x;
  const int gwarp = blockIdx.x * (CL_THREADS / 32) + warp, nwarps = nblk * (CL_THREADS / 32);

  // first nearest neighbours: every point is an "orphan"
  for (int q = blockIdx.x; q < N; q += nblk) rescan(S, S.orphans[q], red_d, red_i);
  grid.sync();

  for (int step = 0; step < N - 1; ++step) {
    // ---- S1: block-partial arg-min of the nearest-neighbour distances
    {
      double bd = 1e300;
      int bi = 0x7fffffff;
      for (int k = blockIdx.x * CL_THREADS + tid; k < N; k += nblk * CL_THREADS)
        if (S.active[k] && S.nn_idx[k] >= 0) amin(bd, bi, S.nn_dist[k], k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        amin(bd, bi, od, oi);
      }
      if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < CL_THREADS / 32; ++w) amin(bd, bi, red_d[w], red_i[w]);
        S.part_d[blockIdx.x] = bd;
        S.part_i[blockIdx.x] = bi;
      }
    }
    grid.sync();
    // ---- S2: block 0 finds the global pair, merges it and publishes (i, j)
    if (blockIdx.x == 0) {
      if (warp == 0) {
        double bd = 1e300;
        int bi = 0x7fffffff;
        for (int q = lane; q < nblk; q += 32) amin(bd, bi, S.part_d[q], S.part_i[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          amin(bd, bi, od, oi);
        }
        if (lane == 0) {
          const int a = bi, b = S.nn_idx[bi];
          sh_d = bd;
          sh_i = a < b ? a : b;  // the merged cluster lives in the lower slot
          sh_j = a < b ? b : a;
        }
      }
      __syncthreads();
      const int mi = sh_i, mj = sh_j;
      const int ni = S.size[mi], nj = S.size[mj];
      double* ci = S.C + static_cast<size_t>(mi) * D;
      const double* cj = S.C + static_cast<size_t>(mj) * D;
      const double wi = static_cast<double>(ni), wj = static_cast<double>(nj);
      for (int c = tid; c < D; c += CL_THREADS) ci[c] = (wi * ci[c] + wj * cj[c]) / (wi + wj);
      if (tid == 0) {
        const int ia = S.cid[mi], ib = S.cid[mj];
        double* z = S.Z + static_cast<size_t>(step) * 4;
        z[0] = static_cast<double>(ia < ib ? ia : ib);
        z[1] = static_cast<double>(ia < ib ? ib : ia);
        z[2] = sh_d;
        z[3] = static_cast<double>(ni + nj);
        S.size[mi] = ni + nj;
        S.cid[mi] = N + step;
        S.active[mj] = 0;
        S.nn_idx[mi] = -1;   // set in S4
        S.nn_dist[mi] = 1e300;
        S.counters[0] = 0;   // orphan count of this step
        S.counters[1] = mi;
        S.counters[2] = mj;
      }
    }
    grid.sync();
    const int mi = S.counters[1], mj = S.counters[2];
    // ---- S3: distance of every live cluster to the merged one (its minimum = the merged cluster's neighbour)
    {
      const double* cn = S.C + static_cast<size_t>(mi) * D;
      double bd = 1e300;
      int bi = 0x7fffffff;
      for (int k = gwarp; k < N; k += nwarps) {
        if (k == mi || !S.active[k]) continue;
        const double d = sqrt(warp_dist2(cn, S.C + static_cast<size_t>(k) * D, D, lane));
        amin(bd, bi, d, k);
        if (lane == 0) {
          S.Dm[static_cast<size_t>(k) * N + mi] = d;
          S.Dm[static_cast<size_t>(mi) * N + k] = d;
          const int nk = S.nn_idx[k];
          if (nk == mi || nk == mj) {          // its neighbour no longer exists as such: rescan its row
            S.orphans[atomicAdd(&S.counters[0], 1)] = k;
          } else if (d < S.nn_dist[k] || (d == S.nn_dist[k] && mi < nk)) {
            S.nn_dist[k] = d;
            S.nn_idx[k] = mi;
          }
        }
      }
      if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < CL_THREADS / 32; ++w) amin(bd, bi, red_d[w], red_i[w]);
        S.part_d[nblk + blockIdx.x] = bd;
        S.part_i[nblk + blockIdx.x] = bi;
      }
    }
    grid.sync();
    // ---- S4: orphans rescan their rows (one CTA each); block 0 finishes the merged cluster's neighbour
    {
      const int n_orph = S.counters[0];
      if (blockIdx.x == 0 && warp == 0) {
        double bd = 1e300;
        int bi = 0x7fffffff;
        for (int q = lane; q < nblk; q += 32) amin(bd, bi, S.part_d[nblk + q], S.part_i[nblk + q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          amin(bd, bi, od, oi);
        }
        if (lane == 0) {
          S.nn_dist[mi] = bi == 0x7fffffff ? 1e300 : bd;
          S.nn_idx[mi] = bi == 0x7fffffff ? -1 : bi;
        }
      }
      for (int q = blockIdx.x; q < n_orph; q += nblk) rescan(S, S.orphans[q], red_d, red_i);
    }
    grid.sync();
  }
}

struct ClLayout {
  size_t off_Dm, off_C, off_nnd, off_part, off_ints, total;
};
ClLayout cl_layout(int N, int D) {
  ClLayout L;
  size_t o = 0;
  L.off_Dm = o;
  o += static_cast<size_t>(N) * N * 8;
  L.off_C = o;
  o += static_cast<size_t>(N) * D * 8;
  L.off_nnd = o;
  o += static_cast<size_t>(N) * 8;
  L.off_part = o;
  o += 8192 * 8;   // 2 x grid (<= 4096) partial minima
  L.off_ints = o;
  o += (static_cast<size_t>(N) * 5 + 8192 + 64) * 4;
  L.total = o + 256;
  return L;
}

}  // namespace

extern "C" size_t sd_centroid_linkage_workspace_bytes(int N, int D) {
  return (N < 1 || D < 1) ? 0 : cl_layout(N, D).total;
}

extern "C" int sd_centroid_linkage_f64(const float* x_dev, int N, int D, double* Z_dev, void* workspace_dev,
                                       void* stream) {
  if (!x_dev || !Z_dev || !workspace_dev || N < 2 || D < 1)
    return fail(SD_ERR_ARG, "sd_centroid_linkage_f64: bad arguments (N=%d D=%d)", N, D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  const ClLayout L = cl_layout(N, D);
  ClState S;
  S.N = N;
  S.D = D;
  S.Dm = reinterpret_cast<double*>(base + L.off_Dm);
  S.C = reinterpret_cast<double*>(base + L.off_C);
  S.nn_dist = reinterpret_cast<double*>(base + L.off_nnd);
  S.part_d = reinterpret_cast<double*>(base + L.off_part);
  S.Z = Z_dev;
  int* ip = reinterpret_cast<int*>(base + L.off_ints);
  S.nn_idx = ip;
  S.size = ip + N;
  S.active = ip + 2 * static_cast<size_t>(N);
  S.cid = ip + 3 * static_cast<size_t>(N);
  S.orphans = ip + 4 * static_cast<size_t>(N);
  S.part_i = ip + 5 * static_cast<size_t>(N);
  S.counters = ip + 5 * static_cast<size_t>(N) + 8192;

  const long tot = static_cast<long>(N) * D;
  cl_init_kernel<<<static_cast<int>((tot + 255) / 256), 256, 0, st>>>(S, x_dev);
  cl_pairwise_kernel<<<dim3((N + 15) / 16, (N + 15) / 16), 256, 0, st>>>(S);
  SD_CUDA_OK(cudaGetLastError());
  int dev = 0, sms = 0, occ = 0;
  SD_CUDA_OK(cudaGetDevice(&dev));
  SD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cl_steps_kernel, CL_THREADS, 0));
  if (occ < 1) return fail(SD_ERR_CUDA, "cl_steps_kernel cannot be made resident");
  if (occ > 2) occ = 2;   // grid barriers dominate a step: keep the grid small
  int gridn = sms * occ;
  if (gridn > 4096) gridn = 4096;
  void* args[] = {&S};
  SD_CUDA_OK(cudaLaunchCooperativeKernel((void*)cl_steps_kernel, dim3(gridn), dim3(CL_THREADS), args, 0, st));
  count_launch(3);
  return SD_OK;
}
