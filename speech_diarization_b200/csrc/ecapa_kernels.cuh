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
    *reinterpret_cast<uint32_t*>(oh + c) = pack_half2(m0, m1);
    *reinterpret_cast<uint32_t*>(oh + C + c) = pack_half2(d0, d1);
  }
}

// Finishes the column statistics the GEMM write-out accumulated per group of G = gcd(128, Tp) rows
// (EpiParams::colsum): window b owns the Tp / G groups that start at rows b*Tp + j*G; the group of row r sits
// at index (r / 128) * (128 / G) + (r % 128) / G.  The groups are added in ascending j, so the result does not
// depend on which batch slot the window occupies.  mean = k + S/T, var = (Q - S^2/T)/T (shift-invariant),
// std = sqrt(max(var, 1e-12)).
// grid B, block 256.  std_out / out_h may be null; out_h gets [mean | std] as f16 when std is wanted.
__device__ __forceinline__ size_t cs_group_index(int b, int j, int Tp, int G) {
  const int r = b * Tp + j * G;
  return static_cast<size_t>(r >> 7) * (128 / G) + (r & 127) / G;
}

__global__ void __launch_bounds__(256)
colstats_finish_kernel(const float* __restrict__ colsum, const float* __restrict__ colsq,
                       const float* __restrict__ shift, int C, int Tp, int T, int G,
                       float* __restrict__ mean_out, int ld_out, float* __restrict__ std_out,
                       __half* __restrict__ out_h) {
  pdl_trigger();
  pdl_wait();
  // grid = B (one CTA per window), four channels per thread and pass: every group's 16-byte loads of both
  // statistics are in flight together (the first version — one channel per thread, one load at a time — spent 31 us
  // on 63 MB)
  const int b = blockIdx.x;
  const int ng = Tp / G;
  const float inv = 1.0f / static_cast<float>(T);
  for (int c = threadIdx.x * 4; c < C; c += 1024) {
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < ng; j0 += 8) {     // eight groups at a time, added in ascending order (slot-invariant)
      float4 vs[8], vq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool ok = j0 + e < ng;
        const size_t o = ok ? cs_group_index(b, j0 + e, Tp, G) * C + c : 0;
        vs[e] = ok ? *reinterpret_cast<const float4*>(colsum + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        vq[e] = ok && colsq != nullptr ? *reinterpret_cast<const float4*>(colsq + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (j0 + e < ng) {
          S.x += vs[e].x; S.y += vs[e].y; S.z += vs[e].z; S.w += vs[e].w;
          Q.x += vq[e].x; Q.y += vq[e].y; Q.z += vq[e].z; Q.w += vq[e].w;
        }
    }
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shift != nullptr) {
      const float4 sh = *reinterpret_cast<const float4*>(shift + c);
      k = make_float4(__half2float(__float2half_rn(sh.x)), __half2float(__float2half_rn(sh.y)),
                      __half2float(__float2half_rn(sh.z)), __half2float(__float2half_rn(sh.w)));
    }
    const float4 mean = make_float4(k.x + S.x * inv, k.y + S.y * inv, k.z + S.z * inv, k.w + S.w * inv);
    *reinterpret_cast<float4*>(mean_out + static_cast<size_t>(b) * ld_out + c) = mean;
    if (std_out != nullptr) {
      const float4 sd = make_float4(sqrtf(fmaxf((Q.x - S.x * S.x * inv) * inv, 1e-12f)), sqrtf(fmaxf((Q.y - S.y * S.y * inv) * inv, 1e-12f)),
                                    sqrtf(fmaxf((Q.z - S.z * S.z * inv) * inv, 1e-12f)), sqrtf(fmaxf((Q.w - S.w * S.w * inv) * inv, 1e-12f)));
      *reinterpret_cast<float4*>(std_out + static_cast<size_t>(b) * ld_out + c) = sd;
      if (out_h != nullptr) {
        __half* oh = out_h + static_cast<size_t>(b) * ld_out;
        *reinterpret_cast<uint2*>(oh + c) = make_uint2(pack_half2(mean.x, mean.y), pack_half2(mean.z, mean.w));
        *reinterpret_cast<uint2*>(oh + C + c) = make_uint2(pack_half2(sd.x, sd.y), pack_half2(sd.z, sd.w));
      }
    }
  }
}

// The whole squeeze-excitation gate of one SERes2Net block in ONE launch (was: statistics finish, hidden layer
// and gate as three latency-bound launches, ~59 us per block of which < 10 us was work):
//   mean[b, c]  = k[c] + (sum of window b's column-sum groups) / T        (or read from `mean_in`)
//   hid[b, j]   = relu(W1[j, :] . mean[b, :] + b1[j])                      j < 128
//   scale[b, c] = sigmoid(W2[c, :] . hid[b, :] + b2[c])
// One CTA = SEG windows; the two weight matrices are f16 ([128][C] each, the second one transposed so that
// consecutive threads read consecutive channels), 512 KB per CTA out of L2.  speechbrain SEBlock (App. A.2).
// grid ceil(B / SEG), block 256, static smem.  C = 1024, S = 128 (the only ECAPA configuration).
constexpr int SEG = 4;
constexpr int SE_C = 1024;
constexpr int SE_S = 128;
constexpr int SE_THREADS = 512;
__global__ void __launch_bounds__(SE_THREADS)
se_gate_kernel(const float* __restrict__ colsum, const float* __restrict__ shift, int Tp, int T, int G,
               const float* __restrict__ mean_in, const __half* __restrict__ W1h, const float* __restrict__ b1,
               const __half* __restrict__ W2th, const float* __restrict__ b2, int B,
               float* __restrict__ mean_out, float* __restrict__ scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float sm_mean[SEG * SE_C];
  __shared__ __align__(16) float sm_hid[SEG * SE_S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * SEG;
  const int nb = min(SEG, B - b0);
  // The kernel is a chain of L2 round trips (partial sums -> W1 -> W2), so every stage requests the next stage's
  // first operands before it starts its own arithmetic.
  // ---- squeeze: thread -> (window u = tid / 128 .. , 4 channels); all of a window's groups requested together
  {
    const int c = (tid & 255) * 4;
    const float inv = 1.0f / static_cast<float>(T);
    const int ng = colsum != nullptr ? Tp / G : 0;
    for (int u = tid >> 8; u < SEG; u += SE_THREADS / 256) {
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < nb) {
        if (colsum != nullptr) {
          float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j0 = 0; j0 < ng; j0 += 8) {     // eight groups' loads in flight (a 1.5 s window has five)
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = j0 + j < ng ? *reinterpret_cast<const float4*>(colsum + cs_group_index(b0 + u, j0 + j, Tp, G) * SE_C + c)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j0 + j < ng) { S.x += v[j].x; S.y += v[j].y; S.z += v[j].z; S.w += v[j].w; }   // ascending: slot-invariant
          }
          float4 kk = make_float4(0.f, 0.f, 0.f, 0.f);
          if (shift != nullptr) {
            const float4 sh = *reinterpret_cast<const float4*>(shift + c);
            kk = make_float4(__half2float(__float2half_rn(sh.x)), __half2float(__float2half_rn(sh.y)),
                             __half2float(__float2half_rn(sh.z)), __half2float(__float2half_rn(sh.w)));
          }
          m = make_float4(kk.x + S.x * inv, kk.y + S.y * inv, kk.z + S.z * inv, kk.w + S.w * inv);
        } else {
          m = *reinterpret_cast<const float4*>(mean_in + static_cast<size_t>(b0 + u) * SE_C + c);
        }
        if (mean_out != nullptr) *reinterpret_cast<float4*>(mean_out + static_cast<size_t>(b0 + u) * SE_C + c) = m;
      }
      *reinterpret_cast<float4*>(sm_mean + u * SE_C + c) = m;
    }
  }
  __syncthreads();
  // ---- excitation layer 1: warp w owns hidden units 8w .. 8w+7; a lane covers the channel octets lane + 32 st.
  // Per st the four windows' means of the octet sit in registers (8 LDS.128) and are used for all eight units
  // (256 FMA): the first version read them from shared memory per unit pair and was bound by the LDS pipe
  // (stall_short_sb + stall_mio = 43 % of the samples).
  const int c2 = (tid & 255) * 4;
  const int jhalf = (tid >> 8) * (SE_S / 2);          // threads 0-255 take hidden units 0-63, 256-511 take 64-127
  {
    float acc[8][SEG];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int u = 0; u < SEG; ++u) acc[jj][u] = 0.f;
    const uint4* const wrow = reinterpret_cast<const uint4*>(W1h + static_cast<size_t>(warp * 8) * SE_C) + lane;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      uint4 wv[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) wv[jj] = __ldg(wrow + jj * (SE_C / 8) + 32 * st);
      float4 m0[SEG], m1[SEG];
      const int c8 = (lane + 32 * st) * 8;
#pragma unroll
      for (int u = 0; u < SEG; ++u) {
        m0[u] = *reinterpret_cast<const float4*>(sm_mean + u * SE_C + c8);
        m1[u] = *reinterpret_cast<const float4*>(sm_mean + u * SE_C + c8 + 4);
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const __half2* wh = reinterpret_cast<const __half2*>(&wv[jj]);
        const float2 w0 = __half22float2(wh[0]), w1 = __half22float2(wh[1]);
        const float2 w2f = __half22float2(wh[2]), w3 = __half22float2(wh[3]);
#pragma unroll
        for (int u = 0; u < SEG; ++u) {
          float a = acc[jj][u];
          a = fmaf(w0.x, m0[u].x, a); a = fmaf(w0.y, m0[u].y, a); a = fmaf(w1.x, m0[u].z, a); a = fmaf(w1.y, m0[u].w, a);
          a = fmaf(w2f.x, m1[u].x, a); a = fmaf(w2f.y, m1[u].y, a); a = fmaf(w3.x, m1[u].z, a); a = fmaf(w3.y, m1[u].w, a);
          acc[jj][u] = a;
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float bj = __ldg(b1 + warp * 8 + jj);
#pragma unroll
      for (int u = 0; u < SEG; ++u) {
        const float a = warp_sum(acc[jj][u]);
        if (lane == 0) sm_hid[u * SE_S + warp * 8 + jj] = fmaxf(a + bj, 0.f);
      }
    }
  }
  uint2 w2[16];       // layer 2's first 16 weight rows, requested before the barrier
#pragma unroll
  for (int q = 0; q < 16; ++q)
    w2[q] = __ldg(reinterpret_cast<const uint2*>(W2th + static_cast<size_t>(jhalf + q) * SE_C + c2));
  __syncthreads();
  // ---- excitation layer 2 + sigmoid: 4 channels per thread, the 128 hidden units split between the two thread
  // halves (combined through shared memory, lower half first: a fixed order)
  {
    float acc[SEG][4];
#pragma unroll
    for (int u = 0; u < SEG; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int j = 0; j < SE_S / 2; j += 16) {
      uint2 wn[16];
      if (j + 16 < SE_S / 2) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
          wn[q] = __ldg(reinterpret_cast<const uint2*>(W2th + static_cast<size_t>(jhalf + j + 16 + q) * SE_C + c2));
      }
#pragma unroll
      for (int q4 = 0; q4 < 16; q4 += 4) {
        float4 hv[SEG];       // four hidden units of every window per (broadcast) 16-byte read
#pragma unroll
        for (int u = 0; u < SEG; ++u) hv[u] = *reinterpret_cast<const float4*>(sm_hid + u * SE_S + jhalf + j + q4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 wa = __half22float2(*reinterpret_cast<const __half2*>(&w2[q4 + e].x));
          const float2 wb = __half22float2(*reinterpret_cast<const __half2*>(&w2[q4 + e].y));
#pragma unroll
          for (int u = 0; u < SEG; ++u) {
            const float h = e == 0 ? hv[u].x : e == 1 ? hv[u].y : e == 2 ? hv[u].z : hv[u].w;
            acc[u][0] = fmaf(h, wa.x, acc[u][0]);
            acc[u][1] = fmaf(h, wa.y, acc[u][1]);
            acc[u][2] = fmaf(h, wb.x, acc[u][2]);
            acc[u][3] = fmaf(h, wb.y, acc[u][3]);
          }
        }
      }
      if (j + 16 < SE_S / 2) {
#pragma unroll
        for (int q = 0; q < 16; ++q) w2[q] = wn[q];
      }
    }
    __syncthreads();                       // sm_mean is dead: reuse it for the upper half's partial sums
    if (tid >= 256) {
#pragma unroll
      for (int u = 0; u < SEG; ++u)
        *reinterpret_cast<float4*>(sm_mean + u * SE_C + c2) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
    }
    __syncthreads();
    if (tid < 256) {
      const float4 bc = *reinterpret_cast<const float4*>(b2 + c2);
#pragma unroll
      for (int u = 0; u < SEG; ++u)
        if (u < nb) {
          const float4 o = *reinterpret_cast<const float4*>(sm_mean + u * SE_C + c2);
          const float z0 = bc.x + acc[u][0] + o.x, z1 = bc.y + acc[u][1] + o.y;
          const float z2 = bc.z + acc[u][2] + o.z, z3 = bc.w + acc[u][3] + o.w;
          *reinterpret_cast<float4*>(scale + static_cast<size_t>(b0 + u) * SE_C + c2) =
              make_float4(1.0f / (1.0f + __expf(-z0)), 1.0f / (1.0f + __expf(-z1)),
                          1.0f / (1.0f + __expf(-z2)), 1.0f / (1.0f + __expf(-z3)));
        }
    }
  }
}

// out[r, c] = w[r, c] * scale[b(r), c] + res[r, c]  over ALL rows (halo rows included, so the
// reflect halo stays valid); 8 channels per thread.
__global__ void __launch_bounds__(256)
se_apply_kernel(const __half* __restrict__ w, int ld_w, const float* __restrict__ scale,
                const __half* __restrict__ res, int ld_res, __half* __restrict__ out, int ld_out,
                long rows, int Tp, int C, int reverse, int* __restrict__ oflow) {
  pdl_trigger();
  pdl_wait();
  float amax = 0.f;
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
      const float o0 = fmaf(a.x, sc[2 * e], q.x), o1 = fmaf(a.y, sc[2 * e + 1], q.y);
      amax = fmaxf(amax, fmaxf(fabsf(o0), fabsf(o1)));
      const uint32_t pk = pack_half2(o0, o1);
      oh[e] = *reinterpret_cast<const __half2*>(&pk);
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + c) = ov;
  }
  if (amax > kHalfMax && oflow != nullptr) atomicOr(oflow, 1);
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
                 int l2_normalize, float eps, float* __restrict__ emb, const int* __restrict__ oflow,
                 int* __restrict__ oflow_sticky) {
  pdl_trigger();
  pdl_wait();
  // an activation left the f16 range somewhere in this forward (saturated, so everything stayed finite): the
  // embeddings would be silently wrong — deliver NaN instead and remember it for the host (SD_ERR_RANGE)
  const bool poisoned = oflow != nullptr && *oflow != 0;
  if (poisoned && oflow_sticky != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *oflow_sticky = 1;
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
    if (c < D) emb[static_cast<long>(row) * D + c] = poisoned ? __int_as_float(0x7fc00000) : v[q] * inv;
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
