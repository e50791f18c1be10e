// ecapa_kernels.cuh — the non-GEMM kernels of the ECAPA-TDNN trunk:
// squeeze-excitation (time mean -> 2-layer MLP -> scale + residual), the global
// statistics of attentive pooling, and the split-K reductions of the per-utterance
// dense layers.  All are bandwidth-bound reductions over the f16 channels-last
// activation tensors [B*Tp, C] written by the GEMM epilogues (gemm_tc.cuh).
//
// Reference arithmetic: speechbrain SEBlock / AttentiveStatisticsPooling /
// asp_bn + fc inside ECAPA_TDNN.forward, reached from speech_encode.py:77 and
// ecapa_annote.py:22.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "sd_ptx.cuh"

namespace sd {

// mean over the T interior frames of utterance b, per channel.
// x: [B*Tp, ld] f16.  grid (C/256, B), block 128: one half2 (2 channels) per thread.
__global__ void __launch_bounds__(128)
time_mean_kernel(const __half* __restrict__ x, int ld, int Tp, int T, int H, int C,
                 float* __restrict__ mean_out /*[B, C]*/) {
  pdl_trigger();
  pdl_wait();
  // utterances last to first: the producer GEMM wrote them first to last, so the most recently written
  // (still L2-resident) rows are read first instead of being evicted by this kernel's own misses
  const int b = gridDim.y - 1 - blockIdx.y;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const __half2* p = reinterpret_cast<const __half2*>(x + (static_cast<size_t>(b) * Tp + H) * ld + c);
  const size_t step = static_cast<size_t>(ld) / 2;
  float s0 = 0.f, s1 = 0.f;
  int t = 0;
  for (; t + 8 <= T; t += 8) {
    __half2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[(t + j) * step];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = __half22float2(v[j]);
      s0 += f.x;
      s1 += f.y;
    }
  }
  for (; t < T; ++t) {
    const float2 f = __half22float2(p[t * step]);
    s0 += f.x;
    s1 += f.y;
  }
  const float inv = 1.0f / static_cast<float>(T);
  mean_out[static_cast<size_t>(b) * C + c] = s0 * inv;
  mean_out[static_cast<size_t>(b) * C + c + 1] = s1 * inv;
}

// mean and std (uniform weights 1/T, eps 1e-12) over interior frames; stats[b] = [mean(C), std(C)].
// ONE pass over the activations: sums of (x - k) and (x - k)^2 with k = the channel's first
// frame, which removes the cancellation of the naive E[x^2] - E[x]^2 form
// (var = (S2 - S1^2 / T) / T is shift-invariant).
__global__ void __launch_bounds__(128)
time_mean_std_kernel(const __half* __restrict__ x, int ld, int Tp, int T, int H, int C,
                     float* __restrict__ stats /*[B, 2C]*/, __half* __restrict__ stats_h /*[B, 2C] or null*/) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const __half2* p = reinterpret_cast<const __half2*>(x + (static_cast<size_t>(b) * Tp + H) * ld + c);
  const size_t step = static_cast<size_t>(ld) / 2;
  const float2 k = __half22float2(p[0]);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  int t = 0;
  for (; t + 8 <= T; t += 8) {
    __half2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[(t + j) * step];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = __half22float2(v[j]);
      const float d0 = f.x - k.x, d1 = f.y - k.y;
      s0 += d0;
      s1 += d1;
      q0 = fmaf(d0, d0, q0);
      q1 = fmaf(d1, d1, q1);
    }
  }
  for (; t < T; ++t) {
    const float2 f = __half22float2(p[t * step]);
    const float d0 = f.x - k.x, d1 = f.y - k.y;
    s0 += d0;
    s1 += d1;
    q0 = fmaf(d0, d0, q0);
    q1 = fmaf(d1, d1, q1);
  }
  const float inv = 1.0f / static_cast<float>(T);
  const float m0 = k.x + s0 * inv, m1 = k.y + s1 * inv;
  const float v0 = (q0 - s0 * s0 * inv) * inv, v1 = (q1 - s1 * s1 * inv) * inv;
  float* o = stats + static_cast<size_t>(b) * 2 * C;
  o[c] = m0;
  o[c + 1] = m1;
  const float d0 = sqrtf(fmaxf(v0, 1e-12f)), d1 = sqrtf(fmaxf(v1, 1e-12f));
  o[C + c] = d0;
  o[C + c + 1] = d1;
  if (stats_h) {  // f16 copy: A operand of the context-bias GEMM
    __half* oh = stats_h + static_cast<size_t>(b) * 2 * C;
    *reinterpret_cast<__half2*>(oh + c) = __floats2half2_rn(m0, m1);
    *reinterpret_cast<__half2*>(oh + C + c) = __floats2half2_rn(d0, d1);
  }
}

// Finishes the column statistics the GEMM write-out accumulated per (m block, window slot)
// (EpiParams::colsum): window b spans m blocks (b*Tp)/128 .. (b*Tp + Tp - 1)/128; its slot in block m is
// b - (m*128)/Tp.  mean = k + S/T, var = (Q - S^2/T)/T (shift-invariant), std = sqrt(max(var, 1e-12)).
// grid (C/256, B), block 256.  std_out / out_h may be null; out_h gets [mean | std] as f16 when std is wanted.
__global__ void __launch_bounds__(256)
colstats_finish_kernel(const float* __restrict__ colsum, const float* __restrict__ colsq,
                       const float* __restrict__ shift, int C, int Tp, int T, int num_m_blocks,
                       float* __restrict__ mean_out, int ld_out, float* __restrict__ std_out,
                       __half* __restrict__ out_h) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int m_lo = (b * Tp) / 128;
  const int m_hi = min((b * Tp + Tp - 1) / 128, num_m_blocks - 1);
  float S = 0.f, Q = 0.f;
  for (int m = m_lo; m <= m_hi; ++m) {
    const int slot = b - (m * 128) / Tp;
    const size_t o = (static_cast<size_t>(m) * 2 + slot) * C + c;
    S += colsum[o];
    if (colsq != nullptr) Q += colsq[o];
  }
  const float k = shift != nullptr ? __half2float(__float2half_rn(shift[c])) : 0.f;
  const float inv = 1.0f / static_cast<float>(T);
  const float mean = k + S * inv;
  mean_out[static_cast<size_t>(b) * ld_out + c] = mean;
  if (std_out != nullptr) {
    const float sd = sqrtf(fmaxf((Q - S * S * inv) * inv, 1e-12f));
    std_out[static_cast<size_t>(b) * ld_out + c] = sd;
    if (out_h != nullptr) {
      out_h[static_cast<size_t>(b) * ld_out + c] = __float2half_rn(mean);
      out_h[static_cast<size_t>(b) * ld_out + C + c] = __float2half_rn(sd);
    }
  }
}

// SE excitation, layer 1: hid[b, j] = relu(W1[j, :] . mean[b, :] + b1[j]).   W1 [S][C] row-major.
// One CTA = SE_U utterances x 32 hidden units (4 per warp).  (SE_U = 16 and a tcgen05 version of these two
// layers were measured: neither beats this — both kernels sit at their launch / latency floor.)
// grid (ceil(B/SE_U), S/32), block 256, dynamic smem SE_U*C floats.
constexpr int SE_U = 4;
__global__ void __launch_bounds__(256)
se_hidden_kernel(const float* __restrict__ mean, const float* __restrict__ W1,
                 const float* __restrict__ b1, int B, int C, int S, float* __restrict__ hid) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];  // [SE_U][C]
  const int b0 = blockIdx.x * SE_U, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = min(SE_U, B - b0);
  for (int i = tid * 4; i < SE_U * C; i += 256 * 4) {
    const int u = i / C;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < nb) v = *reinterpret_cast<const float4*>(mean + static_cast<size_t>(b0 + u) * C + (i - u * C));
    *reinterpret_cast<float4*>(sm + i) = v;
  }
  __syncthreads();
  // The warp's 4 hidden units together, 4 column steps at a time: 16 independent 128-bit weight loads in
  // flight per lane (one unit / two steps at a time left the kernel waiting on L2 round trips: 19 us).
  const int j0 = blockIdx.y * 32 + warp * 4;
  float acc[4][SE_U];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
#pragma unroll
    for (int u = 0; u < SE_U; ++u) acc[jj][u] = 0.f;
  const int n4 = C / 4;
  for (int i0 = lane; i0 < n4; i0 += 128) {
    float4 wv[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int i = i0 + 32 * s;
        wv[s][jj] = (i < n4 && j0 + jj < S) ? __ldg(reinterpret_cast<const float4*>(W1 + static_cast<size_t>(j0 + jj) * C) + i)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int i = i0 + 32 * s;
      if (i < n4) {
#pragma unroll
        for (int u = 0; u < SE_U; ++u) {
          const float4 mv = *reinterpret_cast<const float4*>(sm + u * C + 4 * i);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            acc[jj][u] += wv[s][jj].x * mv.x + wv[s][jj].y * mv.y + wv[s][jj].z * mv.z + wv[s][jj].w * mv.w;
        }
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = j0 + jj;
    if (j >= S) break;
    const float bj = b1[j];
#pragma unroll
    for (int u = 0; u < SE_U; ++u) {
      const float a = warp_sum(acc[jj][u]);
      if (lane == 0 && u < nb) hid[static_cast<size_t>(b0 + u) * S + j] = fmaxf(a + bj, 0.f);
    }
  }
}

// SE excitation, layer 2: scale[b, c] = sigmoid(W2[c, :] . hid[b, :] + b2[c]).   W2t [S][C] (transposed
// conv2 weight, so consecutive threads read consecutive addresses).
// grid (ceil(B/SE_U), C/256), block 256.  S <= 128, S % 4 == 0.
__global__ void __launch_bounds__(256)
se_scale_kernel(const float* __restrict__ hid, const float* __restrict__ W2t,
                const float* __restrict__ b2, int B, int C, int S, float* __restrict__ scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float h[SE_U * 128];
  const int b0 = blockIdx.x * SE_U, tid = threadIdx.x;
  const int nb = min(SE_U, B - b0);
  for (int i = tid; i < SE_U * S; i += 256) {
    const int u = i / S;
    h[i] = u < nb ? hid[static_cast<size_t>(b0 + u) * S + (i - u * S)] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.y * 256 + tid;
  if (c >= C) return;
  float acc[SE_U];
  const float bc = b2[c];
#pragma unroll
  for (int u = 0; u < SE_U; ++u) acc[u] = bc;
  // 16 weight rows per step: 16 independent loads in flight per thread (S % 16 == 0 for S = 128)
  for (int j = 0; j < S; j += 16) {
    float w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = j + q < S ? __ldg(W2t + static_cast<size_t>(j + q) * C + c) : 0.f;
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
      if (j + q < S) {
#pragma unroll
        for (int u = 0; u < SE_U; ++u) {
          const float4 hv = *reinterpret_cast<const float4*>(h + u * S + j + q);
          acc[u] = fmaf(hv.x, w[q], acc[u]);
          acc[u] = fmaf(hv.y, w[q + 1], acc[u]);
          acc[u] = fmaf(hv.z, w[q + 2], acc[u]);
          acc[u] = fmaf(hv.w, w[q + 3], acc[u]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < SE_U; ++u)
    if (u < nb) scale[static_cast<size_t>(b0 + u) * C + c] = 1.0f / (1.0f + __expf(-acc[u]));
}

// out[r, c] = w[r, c] * scale[b(r), c] + res[r, c]  over ALL rows (halo rows included, so the
// reflect halo stays valid); 8 channels per thread.
__global__ void __launch_bounds__(256)
se_apply_kernel(const __half* __restrict__ w, int ld_w, const float* __restrict__ scale,
                const __half* __restrict__ res, int ld_res, __half* __restrict__ out, int ld_out,
                long rows, int Tp, int C, int reverse) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = C / 8;
  const long total = rows * vec_per_row;
  // from the last row to the first when `reverse`: tdnn2 wrote w in ascending row order, so its last ~100 MB are
  // still in L2 when this kernel starts
  for (long i0 = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i0 < total;
       i0 += static_cast<long>(gridDim.x) * 256) {
    const long i = reverse ? total - 1 - i0 : i0;
    const long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row) * 8;
    const int b = static_cast<int>(r / Tp);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + r * ld_w + c);
    const uint4 rv = *reinterpret_cast<const uint4*>(res + r * ld_res + c);
    const float4 s0 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(b) * C + c);
    const float4 s1 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(b) * C + c + 4);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint4 ov;
    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = __half22float2(wh[e]);
      const float2 q = __half22float2(rh[e]);
      oh[e] = __floats2half2_rn(fmaf(a.x, sc[2 * e], q.x), fmaf(a.y, sc[2 * e + 1], q.y));
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + c) = ov;
  }
}

// out[i,:] = x[i,:] / (||x[i,:]|| + eps); one warp per row.
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ x, int N, int D, float eps, float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* p = x + static_cast<size_t>(row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  const float inv = 1.0f / (sqrtf(s) + eps);
  for (int i = lane; i < D; i += 32) out[static_cast<size_t>(row) * D + i] = p[i] * inv;
}

// out[i] = sum_s partial[s * stride + i]  (split-K partial sums, fixed order)
__global__ void __launch_bounds__(256)
sum_splits_kernel(const float* __restrict__ partial, int n_splits, long stride, long count,
                  float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count) return;
  float acc = 0.f;
  for (int s = 0; s < n_splits; ++s) acc += partial[s * stride + i];
  out[i] = acc;
}

// Final step of the embedding: e = sum of the FC's split-K partials (bias already inside split 0),
// optionally e / (||e|| + eps).  One warp per utterance.
__global__ void __launch_bounds__(256)
fc_finish_kernel(const float* __restrict__ partial, int n_splits, long stride, int B, int D,
                 int l2_normalize, float eps, float* __restrict__ emb) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float v[8];  // D <= 256
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    float acc = 0.f;
    if (c < D)
      for (int s = 0; s < n_splits; ++s) acc += partial[s * stride + static_cast<long>(row) * D + c];
    v[q] = acc;
    ss = fmaf(acc, acc, ss);
  }
  float inv = 1.f;
  if (l2_normalize) inv = 1.0f / (sqrtf(warp_sum(ss)) + eps);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    if (c < D) emb[static_cast<long>(row) * D + c] = v[q] * inv;
  }
}

// test hook: interior frames of a padded f16 tensor -> f32 [B, T, C]
__global__ void __launch_bounds__(256)
fetch_interior_kernel(const __half* __restrict__ x, int ld, int col_off, int Tp, int T, int H,
                      int C, long total, float* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    const long bt = i / C;
    const int t = static_cast<int>(bt % T);
    const long b = bt / T;
    out[i] = __half2float(x[(b * Tp + H + t) * ld + col_off + c]);
  }
}

}  // namespace sd
