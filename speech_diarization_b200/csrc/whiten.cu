// whiten.cu — ZCA whitening + L2 normalisation of the segment embeddings:
//   whiten_l2(embs)  (/root/reference/diar_diag.py:187-194; call site :352, between embedding and clustering)
//     X  = embs - embs.mean(0)                     (float32, like the input)
//     C  = np.cov(X.T)                             (float64, D x D)
//     U, S, _ = svd(C);  W = U diag(1/sqrt(S + 1e-6)) U^T
//     Xw = X @ W;  Xw /= ||Xw|| + 1e-9             (float64 result)
// C is symmetric positive semi-definite, so its SVD is its eigendecomposition and W = (C + 1e-6 I)^(-1/2)
// is a matrix function of C — independent of the sign / order ambiguities of U.  Here:
//   column sums and the D x D second-moment matrix in f64 (one CTA per row chunk, fixed-order reduction of
//   the per-CTA partials, so results are reproducible), a one-sided cyclic Jacobi eigen-solver on C in ONE
//   CTA (D/2 disjoint column pairs per round, one warp per pair, the D x D f64 work matrix lives in L2),
//   W = f(0) I + sum_k (f(lambda_k) - f(0)) v_k v_k^T, and a final row-block product + normalisation.  All arithmetic after the centring is
//   f64, as in the reference.  D <= 192, D % 32 == 0.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "sd_ptx.cuh"
#include "sd_status.h"

using namespace sd;

namespace {

constexpr int WH_CHUNKS = 148;      // row chunks (= partial matrices)
constexpr int WH_MAXD = 192;     // ECAPA / ERes2NetV2 / CAM++ embeddings are 192-d

inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

// partial[c][d] = sum over the rows of chunk c of (x[r][d] - sub[d])  (f64; sub may be null)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int N, int D, const float* __restrict__ sub, double* __restrict__ partial) {
  const int c = blockIdx.x, d = threadIdx.x;
  if (d >= D) return;
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const int r0 = c * per, r1 = min(N, r0 + per);
  const float s = sub ? sub[d] : 0.f;
  double acc = 0.0;
  for (int r = r0; r < r1; ++r) acc += static_cast<double>(__fsub_rn(x[static_cast<size_t>(r) * D + d], s));
  partial[static_cast<size_t>(c) * D + d] = acc;
}

// mean32[d] = f32(sum of partials / N)
__global__ void __launch_bounds__(256)
mean_finish_kernel(const double* __restrict__ partial, int n_chunks, int N, int D, float* __restrict__ mean32) {
  const int d = threadIdx.x;
  if (d >= D) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += partial[static_cast<size_t>(c) * D + d];
  mean32[d] = static_cast<float>(s / static_cast<double>(N));
}

// Second moments of the centred rows: part[c][i][j] = sum_r xc[r][i] * xc[r][j], xc = f64(f32(x - mean32)).
// The matrix is symmetric: each thread owns one TILE x TILE block of the upper triangle (D = 192, TILE = 6:
// 528 blocks -> 528 active threads with 36 f64 accumulators each) and writes it and its mirror image; rows
// are staged in shared memory 8 at a time.
constexpr int GRAM_THREADS = 544;
template <int TILE>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_kernel(const float* __restrict__ x, int N, int D, const float* __restrict__ mean32, double* __restrict__ part) {
  __shared__ double rows[8][WH_MAXD];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const int r0 = c * per, r1 = min(N, r0 + per);
  const int tiles = D / TILE;                       // tiles per side; tiles * (tiles + 1) / 2 <= GRAM_THREADS
  // tid -> (a <= b) in row-major order of the upper triangle
  int a = 0, rem = tid;
  while (a < tiles && rem >= tiles - a) { rem -= tiles - a; ++a; }
  const bool act = a < tiles;
  const int ti = a * TILE, tj = (a + rem) * TILE;
  double acc[TILE][TILE];
#pragma unroll
  for (int p = 0; p < TILE; ++p)
#pragma unroll
    for (int q = 0; q < TILE; ++q) acc[p][q] = 0.0;
  for (int r = r0; r < r1; r += 8) {
    __syncthreads();
    for (int i = tid; i < 8 * D; i += GRAM_THREADS) {
      const int rr = i / D, d = i - rr * D;
      rows[rr][d] = (r + rr < r1)
                        ? static_cast<double>(__fsub_rn(x[static_cast<size_t>(r + rr) * D + d], mean32[d])) : 0.0;
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        double u[TILE], v[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) { u[k] = rows[rr][ti + k]; v[k] = rows[rr][tj + k]; }
#pragma unroll
        for (int p = 0; p < TILE; ++p)
#pragma unroll
          for (int q = 0; q < TILE; ++q) acc[p][q] = fma(u[p], v[q], acc[p][q]);
      }
    }
  }
  if (act) {
    double* o = part + static_cast<size_t>(c) * D * D;
#pragma unroll
    for (int p = 0; p < TILE; ++p)
#pragma unroll
      for (int q = 0; q < TILE; ++q) {
        o[static_cast<size_t>(ti + p) * D + tj + q] = acc[p][q];
        o[static_cast<size_t>(tj + q) * D + ti + p] = acc[p][q];
      }
  }
}

// C = (sum_c part[c] - N mu mu^T) / (N - 1), mu = (sum_c colpart[c]) / N   -> G (column-major == row-major,
// C is symmetric).
__global__ void __launch_bounds__(256)
cov_finish_kernel(const double* __restrict__ part, const double* __restrict__ colpart, int n_chunks, int N, int D,
                  double* __restrict__ G) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= D * D) return;
  const int i = idx / D, j = idx - i * D;
  double s = 0.0, mi = 0.0, mj = 0.0;
  for (int c = 0; c < n_chunks; ++c) {
    s += part[static_cast<size_t>(c) * D * D + idx];
    mi += colpart[static_cast<size_t>(c) * D + i];
    mj += colpart[static_cast<size_t>(c) * D + j];
  }
  mi /= static_cast<double>(N);
  mj /= static_cast<double>(N);
  G[idx] = (s - static_cast<double>(N) * mi * mj) / static_cast<double>(N - 1);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One-sided (Hestenes) cyclic Jacobi on the columns of G (one column per contiguous D doubles; G = C
// initially): plane rotations make the columns mutually orthogonal.  With V the accumulated rotations,
// G = C V = V Lambda on exit, i.e. column k is lambda_k v_k: eigenvalue = its norm, eigenvector = its direction.
// V itself is never formed (that halves the traffic): directions of columns with lambda ~ 0 are noise, but
// whiten_matrix_kernel below weights direction k by f(lambda_k) - f(0), which vanishes with lambda_k.
// Round-robin ordering: D - 1 rounds of D / 2 disjoint pairs per sweep, one warp per pair (lanes stride the
// column), __syncthreads between rounds.  The two columns of every pair cross the SM <-> L2 path each round
// (D^2 * 16 B per round): that single-SM bandwidth, not arithmetic, sets the ~1 ms per sweep.
// Stops when a whole sweep rotated nothing (|g_i . g_j| <= 1e-15 ||g_i|| ||g_j|| for every pair) or after 40 sweeps.
__global__ void __launch_bounds__(1024)
jacobi_kernel(double* __restrict__ G, int D, double* __restrict__ lambda, int* __restrict__ sweeps_out) {
  __shared__ int rotated;
  __shared__ int perm[WH_MAXD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = D / 2;
  for (int i = tid; i < D; i += 1024) perm[i] = i;
  int sweep = 0;
  for (; sweep < 40; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < D - 1; ++round) {
      for (int p = warp; p < half; p += 32) {
        double* gi = G + static_cast<size_t>(perm[p]) * D;
        double* gj = G + static_cast<size_t>(perm[D - 1 - p]) * D;
        double a = 0.0, b = 0.0, g = 0.0;
        double xi[WH_MAXD / 32], xj[WH_MAXD / 32];
#pragma unroll
        for (int k = 0; k < WH_MAXD / 32; ++k) {
          const int e = lane + 32 * k;
          xi[k] = e < D ? gi[e] : 0.0;
          xj[k] = e < D ? gj[e] : 0.0;
          a = fma(xi[k], xi[k], a); b = fma(xj[k], xj[k], b); g = fma(xi[k], xj[k], g);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (g * g > 1e-30 * (a * b) && a * b > 0.0) {
          const double zeta = (b - a) / (2.0 * g);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
          for (int k = 0; k < WH_MAXD / 32; ++k) {
            const int e = lane + 32 * k;
            if (e < D) {
              gi[e] = cs * xi[k] - sn * xj[k];
              gj[e] = sn * xi[k] + cs * xj[k];
            }
          }
          if (lane == 0) rotated = 1;
        }
      }
      __syncthreads();
      // round-robin tournament: position 0 fixed, the others rotate by one
      int nv = 0;
      if (tid > 0 && tid < D) nv = perm[tid == 1 ? D - 1 : tid - 1];
      __syncthreads();
      if (tid > 0 && tid < D) perm[tid] = nv;
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  // lambda_k = ||g_k||; normalise the column to the unit eigenvector (left as zeros when lambda_k == 0)
  for (int c = warp; c < D; c += 32) {
    double* gc = G + static_cast<size_t>(c) * D;
    double a = 0.0;
    for (int e = lane; e < D; e += 32) a = fma(gc[e], gc[e], a);
    a = warp_sum_f64(a);
    const double nrm = sqrt(a);
    for (int e = lane; e < D; e += 32) gc[e] = nrm > 0.0 ? gc[e] / nrm : 0.0;
    if (lane == 0) lambda[c] = nrm;
  }
  if (tid == 0) *sweeps_out = sweep;
}

// W = f(0) I + sum_k (f(lambda_k) - f(0)) v_k v_k^T,  f(l) = 1 / sqrt(l + 1e-6)  — equal to V f(Lambda) V^T for an
// orthonormal eigenbasis, and insensitive to the (undetermined) eigenvectors of the null space.
__global__ void __launch_bounds__(256)
whiten_matrix_kernel(const double* __restrict__ Vn, const double* __restrict__ lambda, int D, double* __restrict__ W) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= D * D) return;
  const int i = idx / D, j = idx - i * D;
  const double f0 = 1.0 / sqrt(1e-6);
  double s = i == j ? f0 : 0.0;
  for (int k = 0; k < D; ++k)
    s = fma(Vn[static_cast<size_t>(k) * D + i] * (1.0 / sqrt(lambda[k] + 1e-6) - f0), Vn[static_cast<size_t>(k) * D + j], s);
  W[idx] = s;
}

// out[r] = xw / (||xw|| + 1e-9), xw = f64(f32(x[r] - mean32)) @ W.  8 rows per CTA, 256 threads (thread = column).
__global__ void __launch_bounds__(256)
whiten_apply_kernel(const float* __restrict__ x, int N, int D, const float* __restrict__ mean32,
                    const double* __restrict__ W, double* __restrict__ out) {
  __shared__ double rows[8][WH_MAXD];
  __shared__ double red[8][8];
  const int r0 = blockIdx.x * 8, tid = threadIdx.x;
  for (int i = tid; i < 8 * D; i += 256) {
    const int rr = i / D, d = i - rr * D;
    rows[rr][d] = (r0 + rr < N) ? static_cast<double>(__fsub_rn(x[static_cast<size_t>(r0 + rr) * D + d], mean32[d])) : 0.0;
  }
  __syncthreads();
  double acc[8];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) acc[rr] = 0.0;
  if (tid < D) {
    for (int k = 0; k < D; ++k) {
      const double w = W[static_cast<size_t>(k) * D + tid];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) acc[rr] = fma(rows[rr][k], w, acc[rr]);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const double q = warp_sum_f64(tid < D ? acc[rr] * acc[rr] : 0.0);
    if ((tid & 31) == 0) red[rr][tid >> 5] = q;
  }
  __syncthreads();
  if (tid < D) {
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      double q = 0.0;
      for (int w = 0; w < 8; ++w) q += red[rr][w];
      if (r0 + rr < N) out[static_cast<size_t>(r0 + rr) * D + tid] = acc[rr] / (sqrt(q) + 1e-9);
    }
  }
}

}  // namespace

extern "C" size_t sd_whiten_workspace_bytes(int N, int D) {
  if (N < 2 || D < 32 || D > WH_MAXD) return 0;
  const size_t dd = static_cast<size_t>(D) * D * 8;
  return al256(static_cast<size_t>(WH_CHUNKS) * dd) + 2 * al256(static_cast<size_t>(WH_CHUNKS) * D * 8) + 2 * al256(dd) +
         al256(static_cast<size_t>(D) * 8) + al256(static_cast<size_t>(D) * 4) + 512;
}

extern "C" int sd_whiten_l2_f64(const float* x_dev, int N, int D, double* out_dev, void* workspace_dev, int32_t* sweeps_host,
                                void* stream) {
  if (!x_dev || !out_dev || !workspace_dev || N < 2 || D < 32 || D > WH_MAXD || D % 32)
    return fail(SD_ERR_ARG, "sd_whiten_l2_f64: bad arguments N=%d D=%d (N >= 2, D %% 32 == 0, D <= %d)", N, D, WH_MAXD);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t dd = static_cast<size_t>(D) * D * 8;
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  double* part = reinterpret_cast<double*>(w);              w += al256(static_cast<size_t>(WH_CHUNKS) * dd);
  double* colpart0 = reinterpret_cast<double*>(w);          w += al256(static_cast<size_t>(WH_CHUNKS) * D * 8);
  double* colpart1 = reinterpret_cast<double*>(w);          w += al256(static_cast<size_t>(WH_CHUNKS) * D * 8);
  double* G = reinterpret_cast<double*>(w);                 w += al256(dd);
  double* W = reinterpret_cast<double*>(w);                 w += al256(dd);
  double* lambda = reinterpret_cast<double*>(w);            w += al256(static_cast<size_t>(D) * 8);
  float* mean32 = reinterpret_cast<float*>(w);              w += al256(static_cast<size_t>(D) * 4);
  int* sweeps = reinterpret_cast<int*>(w);
  const int chunks = N < WH_CHUNKS ? N : WH_CHUNKS;

  colsum_kernel<<<chunks, 256, 0, st>>>(x_dev, N, D, nullptr, colpart0);
  mean_finish_kernel<<<1, 256, 0, st>>>(colpart0, chunks, N, D, mean32);
  colsum_kernel<<<chunks, 256, 0, st>>>(x_dev, N, D, mean32, colpart1);
  auto fits = [&](int tile) { return D % tile == 0 && (D / tile) * (D / tile + 1) / 2 <= GRAM_THREADS; };
  if (fits(6)) gram_kernel<6><<<chunks, GRAM_THREADS, 0, st>>>(x_dev, N, D, mean32, part);
  else if (fits(4)) gram_kernel<4><<<chunks, GRAM_THREADS, 0, st>>>(x_dev, N, D, mean32, part);
  else return fail(SD_ERR_UNSUPPORTED, "sd_whiten_l2_f64: no tiling of the %d x %d moment matrix", D, D);
  cov_finish_kernel<<<(D * D + 255) / 256, 256, 0, st>>>(part, colpart1, chunks, N, D, G);
  jacobi_kernel<<<1, 1024, 0, st>>>(G, D, lambda, sweeps);
  whiten_matrix_kernel<<<(D * D + 255) / 256, 256, 0, st>>>(G, lambda, D, W);
  whiten_apply_kernel<<<(N + 7) / 8, 256, 0, st>>>(x_dev, N, D, mean32, W, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(8);
  if (sweeps_host) {   // diagnostics: synchronous read of the number of Jacobi sweeps
    SD_CUDA_OK(cudaStreamSynchronize(st));
    SD_CUDA_OK(cudaMemcpy(sweeps_host, sweeps, sizeof(int), cudaMemcpyDeviceToHost));
  }
  return SD_OK;
}
