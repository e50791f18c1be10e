// fbank.cu — fused log-mel filterbank front end.
//
// Replaces, on the device and without materialising frames or spectra in HBM:
//   variant 0: fbank_batch (/root/reference/speech_encode.py:10-38): torchaudio
//              MelSpectrogram(n_fft=win=400, hop=160, f_min=20, f_max=7900, n_mels=80, power=2)
//              = reflect centre pad, periodic Hann, |rFFT|^2, HTK triangles; log(x+1e-6); CMN.
//   variant 1: speechbrain Fbank + InputNormalization inside encode_batch (call sites
//              speech_encode.py:77, ecapa_annote.py:22): zero centre pad, periodic Hamming,
//              |rFFT|^2, speechbrain triangles on [0,8000], 10 log10(max(x,1e-10)), floor at
//              utterance max - 80 dB, minus the per-mel time mean.
//
// Kernel 1 (fbank_frames_kernel): one CTA = 32 frames of one window, 8 warps x 4 frames.
//   The 400-point real DFT of a windowed frame is a 200-point complex FFT of
//   z[n] = x[2n] + i x[2n+1] (200 = 8 x 25, Cooley-Tukey: 8-point butterflies, twiddle,
//   25-point = 5 x 5 in registers) followed by the real-input split; power, then the
//   banded mel sum (<= 16 bins per filter) and the log, all from shared memory.
//   HBM traffic = the samples once (overlapping frames hit L1/L2) + T*80 floats out.
// Kernel 2 (fbank_norm_kernel): one CTA per window: utterance max / per-mel mean
//   (the window's 48 KB of features are L2-resident), then either the f32 [B,T,80]
//   API output, or the f16 channels-last [B,Tp,128] tensor with the reflect halo that
//   the first ECAPA convolution's TMA loads expect.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include "fbank.cuh"
#include "sd_ptx.cuh"
#include "sd_status.h"

namespace sd {

constexpr int NFFT = 400;
constexpr int HOP = 160;
constexpr int NBIN = 201;
constexpr int NMEL = 80;
constexpr int MEL_MAXLEN = 16;
constexpr int FR_PER_CTA = 32;
constexpr int ZSTRIDE = 200;  // float2 per frame: 400 words = 16 mod 32, so the two frames of a half-warp in the
                              // 25-point stage (lane stride 25 float2) hit disjoint bank pairs
constexpr int PSTRIDE = 205;  // floats per frame of power (odd: lane = frame reads in the mel stage are conflict-free)
constexpr int SPAN = (FR_PER_CTA - 1) * HOP + NFFT;  // samples one CTA's frames cover (5360)
static_assert(SPAN % 4 == 0 && SPAN <= FR_PER_CTA * PSTRIDE, "sample span must fit the aliased power buffer");

struct FbankTables {
  float window[NFFT];
  float2 tw200[200];  // W_200^(n2*k1) at [k1*25 + n2] (consecutive lanes = consecutive n2)
  float2 tw400[NBIN]; // W_400^k
  float2 tw25[25];    // W_25^j
  int mel_start[NMEL];
  int mel_len[NMEL];
  float mel_w[MEL_MAXLEN * NMEL];  // [tap j][mel m]: consecutive lanes (mels) read consecutive words
};

struct FbankSmem {
  float2 z[FR_PER_CTA * ZSTRIDE];   // FFT work area; later the [32 frames][81] mel output tile
  float pw[FR_PER_CTA * PSTRIDE];
  float window[NFFT];
  float2 tw200[200];
  float2 tw400[NBIN];
  float2 tw25[25];
};
static_assert(FR_PER_CTA * (NMEL + 1) * 4 <= FR_PER_CTA * ZSTRIDE * 8, "mel output tile aliases the FFT work area");

// banded mel filters of the two variants, read with warp-uniform indices; c_mel_wt is [mel][tap] so that a
// filter's taps are immediate offsets
__constant__ int c_mel_start[2][NMEL];
__constant__ int c_mel_len[2][NMEL];
__constant__ float c_mel_wt[2][NMEL * MEL_MAXLEN];

template <int VARIANT, int STRIDE, int TAPS>
__device__ __forceinline__ float mel_taps(const float* pp, int m) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int j = 0; j < TAPS; ++j) {
    const float t = pp[j * STRIDE] * c_mel_wt[VARIANT][m * MEL_MAXLEN + j];
    if (j & 1) a1 += t; else a0 += t;
  }
  return a0 + a1;
}
// sum_j pp[j * STRIDE] * w[m][j] over the filter's len bins; m and len are warp-uniform (both banks have filters of
// 1..13 bins), so this is a jump to straight-line code
template <int VARIANT, int STRIDE>
__device__ __forceinline__ float mel_filter(const float* pp, int m, int len) {
  switch (len) {
    case 0: return 0.f;
    case 1: return mel_taps<VARIANT, STRIDE, 1>(pp, m);
    case 2: return mel_taps<VARIANT, STRIDE, 2>(pp, m);
    case 3: return mel_taps<VARIANT, STRIDE, 3>(pp, m);
    case 4: return mel_taps<VARIANT, STRIDE, 4>(pp, m);
    case 5: return mel_taps<VARIANT, STRIDE, 5>(pp, m);
    case 6: return mel_taps<VARIANT, STRIDE, 6>(pp, m);
    case 7: return mel_taps<VARIANT, STRIDE, 7>(pp, m);
    case 8: return mel_taps<VARIANT, STRIDE, 8>(pp, m);
    case 9: return mel_taps<VARIANT, STRIDE, 9>(pp, m);
    case 10: return mel_taps<VARIANT, STRIDE, 10>(pp, m);
    case 11: return mel_taps<VARIANT, STRIDE, 11>(pp, m);
    case 12: return mel_taps<VARIANT, STRIDE, 12>(pp, m);
    case 13: return mel_taps<VARIANT, STRIDE, 13>(pp, m);
    default: return mel_taps<VARIANT, STRIDE, MEL_MAXLEN>(pp, m);
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i : (x + iy)(-i) = y - ix
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// forward 5-point DFT, in place
__device__ __forceinline__ void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const float2 m1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
  const float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  a0 = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  // out1 = m1 - i u1 ; out4 = m1 + i u1 ; out2 = m2 - i u2 ; out3 = m2 + i u2
  a1 = make_float2(m1.x + u1.y, m1.y - u1.x);
  a4 = make_float2(m1.x - u1.y, m1.y + u1.x);
  a2 = make_float2(m2.x + u2.y, m2.y - u2.x);
  a3 = make_float2(m2.x - u2.y, m2.y + u2.x);
}

template <int VARIANT>
__device__ __forceinline__ float load_sample(const float* __restrict__ w, int i, int n) {
  if (VARIANT == 0) {  // reflect
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  }
  return (i >= 0 && i < n) ? __ldg(w + i) : 0.f;
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 2)
fbank_frames_kernel(const float* __restrict__ wav, long wav_stride, const long* __restrict__ offsets, int n_samples, int T,
                    int n_windows, const FbankTables* __restrict__ gtab, float* __restrict__ raw) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FbankSmem& S = *reinterpret_cast<FbankSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nc = (T + FR_PER_CTA - 1) / FR_PER_CTA;
  const int total = n_windows * nc;

  // FFT tables -> shared, once per (persistent) CTA
  for (int i = tid; i < NFFT; i += 256) S.window[i] = gtab->window[i];
  for (int i = tid; i < 200; i += 256) S.tw200[i] = gtab->tw200[i];
  for (int i = tid; i < NBIN; i += 256) S.tw400[i] = gtab->tw400[i];
  if (tid < 25) S.tw25[tid] = gtab->tw25[tid];

  // the chunk's sample span (32 frames x hop + 240) as float4 per thread; the centre padding (reflect / zero) is
  // resolved here, once per sample instead of once per FFT input.  Loaded one chunk ahead (under steps 3 and 4 of
  // the previous chunk), so the global latency is off the critical path.
  constexpr int LD4 = (SPAN / 4 + 255) / 256;
  float4 span[LD4];
  auto load_span = [&](int chunk) {
    const int b = chunk / nc;
    const int f0 = (chunk - b * nc) * FR_PER_CTA;
    const float* w = wav + (offsets != nullptr ? __ldg(offsets + b) : static_cast<long>(b) * wav_stride);
    const int gs0 = f0 * HOP - NFFT / 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(w + gs0) & 15) == 0;
#pragma unroll
    for (int e = 0; e < LD4; ++e) {
      const int i4 = tid + 256 * e;
      const int g = gs0 + 4 * i4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i4 < SPAN / 4) {
        if (aligned && g >= 0 && g + 3 < n_samples) {
          v = __ldg(reinterpret_cast<const float4*>(w + g));
        } else {
          v.x = load_sample<VARIANT>(w, g, n_samples);
          v.y = load_sample<VARIANT>(w, g + 1, n_samples);
          v.z = load_sample<VARIANT>(w, g + 2, n_samples);
          v.w = load_sample<VARIANT>(w, g + 3, n_samples);
        }
      }
      span[e] = v;
    }
  };
  if (blockIdx.x < total) load_span(blockIdx.x);

  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int b = chunk / nc;
    const int f0 = (chunk - b * nc) * FR_PER_CTA;

    // ---- step 0: registers -> shared.  The span aliases the power buffer, which is not written before step 3.
    float* xs = S.pw;
#pragma unroll
    for (int e = 0; e < LD4; ++e) {
      const int i4 = tid + 256 * e;
      if (i4 < SPAN / 4) reinterpret_cast<float4*>(xs)[i4] = span[e];
    }
    __syncthreads();   // (also: tables written, previous chunk's output tile read)

    // ---- step 1: 8-point DFTs over n1 for each (frame, n2); twiddle; store Y[k1*25 + n2]
    for (int item = lane; item < 100; item += 32) {
      const int fl = warp * 4 + item / 25;  // local frame
      const int n2 = item % 25;
      const int f = f0 + fl;
      float2 a[8];
      if (f < T) {
        const float* xf = xs + fl * HOP;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int n = 2 * (25 * n1 + n2);
          const float2 wn = *reinterpret_cast<const float2*>(&S.window[n]);
          const float2 xv = *reinterpret_cast<const float2*>(xf + n);
          a[n1].x = wn.x * xv.x;
          a[n1].y = wn.y * xv.y;
        }
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) a[n1] = make_float2(0.f, 0.f);
      }
      const float2 b0 = cadd(a[0], a[4]), b1 = csub(a[0], a[4]), b2 = cadd(a[2], a[6]), b3 = csub(a[2], a[6]);
      const float2 b4 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]), b6 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
      const float2 c0 = cadd(b0, b2), c2 = csub(b0, b2), c1 = cadd(b1, mul_mi(b3)), c3 = csub(b1, mul_mi(b3));
      const float2 c4 = cadd(b4, b6), c6 = csub(b4, b6), c5 = cadd(b5, mul_mi(b7)), c7 = csub(b5, mul_mi(b7));
      const float r = 0.70710678118654752f;
      const float2 w1c5 = make_float2(r * (c5.x + c5.y), r * (c5.y - c5.x));    // (1-i)/sqrt2 * c5
      const float2 w3c7 = make_float2(r * (c7.y - c7.x), -r * (c7.x + c7.y));   // (-1-i)/sqrt2 * c7
      float2 X[8];
      X[0] = cadd(c0, c4); X[4] = csub(c0, c4);
      X[1] = cadd(c1, w1c5); X[5] = csub(c1, w1c5);
      X[2] = cadd(c2, mul_mi(c6)); X[6] = csub(c2, mul_mi(c6));
      X[3] = cadd(c3, w3c7); X[7] = csub(c3, w3c7);
      float2* zf = S.z + fl * ZSTRIDE;
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) zf[k1 * 25 + n2] = cmul(X[k1], S.tw200[k1 * 25 + n2]);
    }
    __syncthreads();  // every warp is done with the staged samples before step 3 overwrites them with power

    // ---- step 2: 25-point DFT over n2 for each (frame, k1): lane = frame_local*8 + k1
    {
      const int fl = warp * 4 + (lane >> 3);
      const int k1 = lane & 7;
      float2* zf = S.z + fl * ZSTRIDE;
      float2 y[25];
#pragma unroll
      for (int i = 0; i < 25; ++i) y[i] = zf[k1 * 25 + i];
      __syncwarp();
      // n2 = 5a + b : DFT5 over a for each b, output index c replaces a
#pragma unroll
      for (int bb = 0; bb < 5; ++bb) dft5(y[bb], y[5 + bb], y[10 + bb], y[15 + bb], y[20 + bb]);
      // twiddle W_25^(b*c); y[5c + b]
#pragma unroll
      for (int c = 1; c < 5; ++c)
#pragma unroll
        for (int bb = 1; bb < 5; ++bb) y[5 * c + bb] = cmul(y[5 * c + bb], S.tw25[(bb * c) % 25]);
      // DFT5 over b for each c -> Z[c + 5d] at y[5c + d]
#pragma unroll
      for (int c = 0; c < 5; ++c) dft5(y[5 * c], y[5 * c + 1], y[5 * c + 2], y[5 * c + 3], y[5 * c + 4]);
      // k2 = c + 5d ; k = k1 + 8*k2
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = 0; d < 5; ++d) zf[k1 + 8 * (c + 5 * d)] = y[5 * c + d];
    }
    __syncwarp();
    if (chunk + gridDim.x < total) load_span(chunk + gridDim.x);   // next chunk's samples, in flight under steps 3 and 4

    // ---- step 3: real-input split + power for the warp's 4 frames, bins k and 200 - k together:
    // with a = Z[k], c = conj(Z[200 - k]):  X[k] = (a + c)/2 + W^k (-i)(a - c)/2  and  X[200 - k] = conj((a + c)/2 - W^k (-i)(a - c)/2),
    // so one (sum, rotated difference) pair gives both powers; the two halvings become one exact * 0.25 on the power.
#pragma unroll
    for (int fi = 0; fi < 4; ++fi) {
      const int fl = warp * 4 + fi;
      const float2* zf = S.z + fl * ZSTRIDE;
      float* pf = S.pw + fl * PSTRIDE;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = lane + 32 * r;
        if (k <= 100) {
          const float2 zk = zf[k];
          float2 zc = zf[k == 0 ? 0 : 200 - k];
          zc.y = -zc.y;
          const float2 e = cadd(zk, zc);
          const float2 o = mul_mi(csub(zk, zc));
          const float2 t = cmul(S.tw400[k], o);
          const float2 xp = cadd(e, t), xm = csub(e, t);
          pf[k] = 0.25f * (xp.x * xp.x + xp.y * xp.y);
          pf[200 - k] = 0.25f * (xm.x * xm.x + xm.y * xm.y);
        }
      }
    }
    __syncthreads();   // all frames' power written; the FFT work area is free

    // ---- step 4: banded mel sums with lane = frame (a warp's mel, its first bin and its length are warp-uniform:
    // straight-line taps with constant-bank weights), log, then through a [32][81] tile so that the
    // [frame][80] rows leave as full lines
    float* otile = reinterpret_cast<float*>(S.z);
#pragma unroll 1
    for (int m = warp; m < NMEL; m += 8) {
      const int k0 = c_mel_start[VARIANT][m], len = c_mel_len[VARIANT][m];
      const float acc = mel_filter<VARIANT, 1>(S.pw + lane * PSTRIDE + k0, m, len);
      float v;
      if (VARIANT == 0) v = logf(acc + 1e-6f);
      else v = 10.0f * log10f(fmaxf(acc, 1e-10f));
      otile[lane * (NMEL + 1) + m] = v;
    }
    __syncthreads();
    {
      const int nfr = min(FR_PER_CTA, T - f0);
      float* out = raw + (static_cast<size_t>(b) * T + f0) * NMEL;
      for (int item = tid; item < nfr * NMEL; item += 256) {
        const int fl = item / NMEL, m = item - fl * NMEL;
        out[item] = otile[fl * (NMEL + 1) + m];
      }
    }
    // the next chunk's staging writes S.pw (all mel reads are behind the barrier above) and its step 1 writes the
    // FFT area only after its own first barrier, which every thread reaches after this copy-out
  }
}

// =====================================================================================================
// Tensor-core front end (default): the 400-point real DFT of every frame as tcgen05 GEMMs.
//
// Folding the real input twice (n <-> 400-n, then n <-> 200-n) splits the DFT into four ~100 x 100 real blocks:
//   even bins  Re X[2m]   = sum_{n=0..100} p[n] cos(pi m n / 100)          p = a[n] + a[200-n]   (p[100] = a[100])
//   odd  bins  Re X[2m+1] = sum_{n=0..99}  q[n] cos(pi (2m+1) n / 200)     q = a[n] - a[200-n]
//   even bins  Im X[2m]   = sum_{n=1..99}  r[n] sin(pi m n / 100)          r = b[n] - b[200-n]
//   odd  bins  Im X[2m+1] = sum_{n=1..100} t[n] sin(pi (2m+1) n / 200)     t = b[n] + b[200-n]   (t[100] = b[100])
// with a[n] = x[n] + x[400-n], b[n] = x[n] - x[400-n] of the windowed frame (the periodic Hann / Hamming windows
// are symmetric about n = 200, so the window is applied after the first fold).  4 x 100 x 100 MACs per frame
// instead of the 400 x 402 of the plain DFT, and every matrix fits an M = 128 tile.
//
// Precision: both operands are split into f16 hi + f16 lo * 2^-11 (lo scaled by 2^11 so that it is a normal f16),
// D = Wh xh + 2^-11 (Wh xl + Wl xh) with the correction in its own TMEM accumulator; each frame is first scaled by
// a power of two to max|x| in [1, 2) (exact, undone on the power).  Measured against an f64 DFT this is as
// accurate as the f32 FFT it replaces (tools/micro/fbank_tc_proto.py).
//
// One CTA per SM, two ROLES: even CTAs hold the two cosine matrices (hi and lo: 104 KB of shared memory, loaded
// once), odd CTAs the two sine matrices; a CTA walks over half-chunks of 32 frames (N = 32) of its share of the
// windows: stage the samples (cp.async, two buffers), fold + split into the K-major 128-byte-swizzled B operand,
// 42 MMAs (M = 128 bins, N = 32 frames, K = 16) from one thread, then two iterations later the epilogue: TMEM ->
// power -> shared memory (aliasing the consumed B operand) -> banded mel sum -> red.add into raw[B, T, 80], which
// therefore receives  sum_k w[k] Re^2  from the cosine CTA and  sum_k w[k] Im^2  from the sine CTA (two addends on
// a zeroed buffer: order-independent, bit-deterministic).  The log is taken by fbank_norm_kernel.
constexpr int FT_N = 32;                       // frames per half-chunk (MMA N)
constexpr int FT_WROWS = 104;                  // stored rows per matrix (13 groups of 8; the MMA reads 128, the rest is never used)
constexpr int FT_KSTEPS = 7;                   // K = 112 >= 101
constexpr int FT_WCHUNK = FT_WROWS * 128;      // bytes of one 64-column K chunk of a matrix
constexpr int FT_WMAT = 2 * FT_WCHUNK;         // one matrix (hi or lo)
constexpr int FT_WROLE = 4 * FT_WMAT;          // [block 2][hi, lo]
constexpr int FT_BCHUNK = FT_N * 128;
constexpr int FT_BBUF = 8 * FT_BCHUNK;         // [vector 2][hi, lo][K chunk 2]
constexpr int FT_SPAN = (FT_N - 1) * HOP + NFFT + 8;   // staged samples per half-chunk (+8: the misaligned reversed reads)
constexpr int FT_SBUF = ((FT_SPAN + 7) / 8) * 8;
constexpr int FT_PSTRIDE = 33;
constexpr int FT_PROWS = NBIN + MEL_MAXLEN - 1;  // rows past bin 200 stay zero: the mel loop reads whole tap groups
constexpr int FT_OSTRIDE = NMEL + 1;
constexpr int FT_GROUP = 256;                  // threads of one compute group
constexpr int FT_GWARPS = FT_GROUP / 32;
constexpr int FT_THREADS = 2 * FT_GROUP + 32;  // two compute groups + the MMA warp
constexpr int FT_LD4 = (FT_SBUF / 4 + FT_GROUP - 1) / FT_GROUP;   // float4 per thread of one staged span
constexpr float FT_LO_SCALE = 2048.f;
static_assert(FT_PROWS * FT_PSTRIDE * 4 <= FT_BBUF, "power staging must fit the consumed B operand");
static_assert(FT_N * FT_OSTRIDE <= FT_SBUF, "mel output tile must fit the span buffer");
static_assert(NMEL % FT_GWARPS == 0, "mel stage: whole mels per warp");

struct FtSmem {
  uint8_t w[FT_WROLE];
  uint8_t b[2][FT_BBUF];        // per group: B operand, then the power staging
  float s[2][FT_SBUF];          // per group: sample span, then the mel output tile
  float win[112], winr[112];
  float wmax[2][FT_GWARPS];
  uint64_t mma_done[2], b_full[2];
  uint32_t tmem_ptr;
};


template <int ID>
__device__ __forceinline__ void ft_group_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(FT_GROUP) : "memory"); }
__device__ __forceinline__ void ft_group_sync(int g) {
  if (g == 0) ft_group_sync<1>(); else ft_group_sync<2>();
}

// the samples of frames [f0, f0 + 32) of one window into registers; centre padding resolved here
template <int VARIANT>
__device__ __forceinline__ void ft_load_span(float4 (&r)[FT_LD4], const float* __restrict__ w, int f0, int n_samples, int gt) {
  const int gs0 = f0 * HOP - NFFT / 2;
  const bool aligned = (reinterpret_cast<uintptr_t>(w + gs0) & 15) == 0;
  if (aligned && gs0 >= 0 && gs0 + FT_SBUF <= n_samples) {   // interior span: no per-sample tests
#pragma unroll
    for (int e = 0; e < FT_LD4; ++e) {
      const int i4 = gt + e * FT_GROUP;
      r[e] = i4 < FT_SBUF / 4 ? __ldg(reinterpret_cast<const float4*>(w + gs0) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < FT_LD4; ++e) {
    const int i4 = gt + e * FT_GROUP;
    const int g = gs0 + 4 * i4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < FT_SBUF / 4) {
      if (aligned && g >= 0 && g + 3 < n_samples) {
        v = __ldg(reinterpret_cast<const float4*>(w + g));
      } else {
        v.x = load_sample<VARIANT>(w, g, n_samples);
        v.y = load_sample<VARIANT>(w, g + 1, n_samples);
        v.z = load_sample<VARIANT>(w, g + 2, n_samples);
        v.w = load_sample<VARIANT>(w, g + 3, n_samples);
      }
    }
    r[e] = v;
  }
}

// 8 values -> hi (f16) and lo = (v - hi) * 2^11 (f16), packed for one 16-byte store each
__device__ __forceinline__ void ft_split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn((v[2 * e] - back.x) * FT_LO_SCALE, (v[2 * e + 1] - back.y) * FT_LO_SCALE);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void ft_ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// v[j] = p[8 - j], j = 0..7: eight floats in descending order that end one float past a 16-byte boundary
__device__ __forceinline__ void ft_ld8_rev(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4),
               c = *reinterpret_cast<const float4*>(p + 8);
  v[0] = c.x; v[1] = b.w; v[2] = b.z; v[3] = b.y; v[4] = b.x; v[5] = a.w; v[6] = a.z; v[7] = a.y;
}


template <int VARIANT>
__global__ void __launch_bounds__(FT_THREADS, 1)
fbank_tc_kernel(const float* __restrict__ wav, long wav_stride, const long* __restrict__ offsets, int n_samples, int T,
                int n_windows, const FbankTables* __restrict__ gtab, const uint8_t* __restrict__ wimg,
                float* __restrict__ part_cos, float* __restrict__ part_sin) {
  extern __shared__ __align__(16) uint8_t smem_raw[];   // aligned up to 1024 by hand (the swizzle works on address bits)
  FtSmem& S = *reinterpret_cast<FtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int role = blockIdx.x & 1;
  const int G = gridDim.x >> 1, gidx = blockIdx.x >> 1;
  const int nc = (T + FT_N - 1) / FT_N;
  const int total = n_windows * nc;
  const int n_mine = gidx < total ? (total - gidx + G - 1) / G : 0;   // half-chunks gidx, gidx + G, ...
  float* __restrict__ part = role == 0 ? part_cos : part_sin;

  // ---- prologue (all warps): matrices, window tables, barriers, TMEM
  {
    const uint4* src = reinterpret_cast<const uint4*>(wimg + static_cast<size_t>(role) * FT_WROLE);
    uint4* dst = reinterpret_cast<uint4*>(S.w);
    for (int i = tid; i < FT_WROLE / 16; i += FT_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < 112; i += FT_THREADS) {
    S.win[i] = gtab->window[i];
    S.winr[i] = gtab->window[200 - i];
  }
  for (int i = tid; i < 2 * FT_BBUF / 16; i += FT_THREADS) reinterpret_cast<uint4*>(S.b[0])[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&S.mma_done[0], 1);
    mbar_init(&S.mma_done[1], 1);
    mbar_init(&S.b_full[0], 1);
    mbar_init(&S.b_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&S.tmem_ptr, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_ptr;

  if (warp == 2 * FT_GWARPS) {
    // ================================================================= MMA warp: one thread issues every MMA
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(FT_N, 0);
      const uint32_t wb = smem_u32(S.w);
      for (int i = 0; i < n_mine; ++i) {
        const int g = i & 1, t = i >> 1;
        mbar_wait(&S.b_full[g], t & 1);
        tc_fence_after();
        const uint32_t bb = smem_u32(S.b[g]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t d_main = tmem_base + g * 128 + j * 64, d_corr = d_main + 32;
          const uint32_t wh = wb + (2 * j) * FT_WMAT, wl = wh + FT_WMAT;
          const uint32_t xh = bb + (4 * j) * FT_BCHUNK, xl = xh + 2 * FT_BCHUNK;
#pragma unroll
          for (int ks = 0; ks < FT_KSTEPS; ++ks) {
            const uint32_t ao = (ks >> 2) * FT_WCHUNK + (ks & 3) * 32, bo = (ks >> 2) * FT_BCHUNK + (ks & 3) * 32;
            umma_f16(d_main, make_smem_desc_sw128(wh + ao), make_smem_desc_sw128(xh + bo), idesc, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < FT_KSTEPS; ++ks) {
            const uint32_t ao = (ks >> 2) * FT_WCHUNK + (ks & 3) * 32, bo = (ks >> 2) * FT_BCHUNK + (ks & 3) * 32;
            umma_f16(d_corr, make_smem_desc_sw128(wh + ao), make_smem_desc_sw128(xl + bo), idesc, ks > 0);
          }
#pragma unroll
          for (int ks = 0; ks < FT_KSTEPS; ++ks) {
            const uint32_t ao = (ks >> 2) * FT_WCHUNK + (ks & 3) * 32, bo = (ks >> 2) * FT_BCHUNK + (ks & 3) * 32;
            umma_f16(d_corr, make_smem_desc_sw128(wl + ao), make_smem_desc_sw128(xh + bo), idesc, 1);
          }
        }
        umma_commit(&S.mma_done[g]);
      }
    }
  } else {
    // ================================================================= two compute groups, alternate half-chunks
    const int g = warp / FT_GWARPS;             // group
    const int gw = warp - g * FT_GWARPS;        // warp within the group
    const int gt = tid - g * FT_GROUP;          // thread within the group
    const int n_grp = (n_mine - g + 1) / 2;     // this group's half-chunks: i = 2 t + g
    float* xs = S.s[g];
    uint8_t* bbuf = S.b[g];
    float* P = reinterpret_cast<float*>(bbuf);  // [FT_PROWS][33] power staging once the MMAs have consumed B
    // position of half-chunk t of this group, advanced by 2 G per step without divisions
    const int dq = (2 * G) / nc, dr = 2 * G - dq * nc;
    int wb_, wc_;   // window, half-chunk within the window, of the NEXT span to load
    {
      const int idx = gidx + g * G;
      wb_ = idx / nc;
      wc_ = idx - wb_ * nc;
    }
    auto advance = [&]() {
      wb_ += dq;
      wc_ += dr;
      if (wc_ >= nc) { wc_ -= nc; wb_ += 1; }
    };
    auto wav_of = [&](int b) {
      return wav + (offsets != nullptr ? __ldg(offsets + b) : static_cast<long>(b) * wav_stride);
    };
    float4 span[FT_LD4];
    int cur_b = wb_, cur_f0 = wc_ * FT_N;       // position of the span held in registers
    if (n_grp > 0) ft_load_span<VARIANT>(span, wav_of(cur_b), cur_f0, n_samples, gt);
    advance();
    int prev_b = 0, prev_f0 = 0;
    float prev_unscale = 1.f;

    for (int t = 0; t <= n_grp; ++t) {
      // -------------------------------------------------------------- epilogue of this group's half-chunk t - 1
      if (t >= 1) {
        const int nfr = min(FT_N, T - prev_f0);
        if (gw == 0) mbar_wait(&S.mma_done[g], (t - 1) & 1);   // one warp polls, the others sleep in the barrier
        ft_group_sync(g);
        tc_fence_after();
        {
          const int q = gw & 3, ch = gw >> 2;      // TMEM lane quarter, 16-column half of every accumulator
          const uint32_t tb = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 128 + ch * 16;
          uint32_t vm[2][16], vc[2][16];
          tmem_ld16(tb, vm[0]);
          tmem_ld16(tb + 32, vc[0]);
          tmem_ld16(tb + 64, vm[1]);
          tmem_ld16(tb + 96, vc[1]);
          tmem_ld_wait();
          const int m = q * 32 + lane;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (m < (j == 0 ? 101 : 100)) {
              float* Pr = P + (2 * m + j) * FT_PSTRIDE + ch * 16;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float v = fmaf(__uint_as_float(vc[j][e]), 1.f / FT_LO_SCALE, __uint_as_float(vm[j][e]));
                Pr[e] = v * v * prev_unscale;
              }
            }
          }
          // rows 201.. stay zero (the fold had this memory as f16 operand data)
          for (int z = gt; z < (FT_PROWS - NBIN) * FT_PSTRIDE; z += FT_GROUP) P[NBIN * FT_PSTRIDE + z] = 0.f;
        }
        tc_fence_before();
        ft_group_sync(g);
        // banded mel sums: lane = frame, a warp's mel is warp-uniform; tap groups of 4 / 8 / 13 with zero weights
        // past the filter's length
        float* otile = xs;                         // the span buffer is free until this iteration's samples go in
#pragma unroll 1
        for (int m = gw; m < NMEL; m += FT_GWARPS) {
          const int k0 = c_mel_start[VARIANT][m], len = c_mel_len[VARIANT][m];
          const float* pp = P + k0 * FT_PSTRIDE + lane;
          const float acc = mel_filter<VARIANT, FT_PSTRIDE>(pp, m, len);
          otile[lane * FT_OSTRIDE + m] = acc;
        }
        ft_group_sync(g);
        float* out = part + (static_cast<size_t>(prev_b) * T + prev_f0) * NMEL;
        for (int item = gt; item < nfr * NMEL; item += FT_GROUP) {
          const int fl = item / NMEL, m = item - fl * NMEL;
          out[item] = otile[fl * FT_OSTRIDE + m];
        }
        ft_group_sync(g);                          // tile read before the samples overwrite it
      }
      if (t >= n_grp) break;
      // -------------------------------------------------------------- samples of half-chunk t: registers -> shared
      const int nfr = min(FT_N, T - cur_f0);
      {
        float mx = 0.f;
#pragma unroll
        for (int e = 0; e < FT_LD4; ++e) {
          const int i4 = gt + e * FT_GROUP;
          const float4 v = span[e];
          mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
          if (i4 < FT_SBUF / 4) reinterpret_cast<float4*>(xs)[i4] = v;
        }
        mx = warp_max(mx);
        if (lane == 0) S.wmax[g][gw] = mx;
      }
      ft_group_sync(g);
      // one power-of-two scale per half-chunk: max|x| into [1, 2) (exact; keeps hi and lo normal f16 for samples
      // down to 2^-14 of the span's peak, and a loud input inside the f16 range); exponent clamped so that the
      // inverse on the power stays a normal f32
      float sc;
      {
        float mx = S.wmax[g][0];
#pragma unroll
        for (int e = 1; e < FT_GWARPS; ++e) mx = fmaxf(mx, S.wmax[g][e]);
        const int ex = static_cast<int>((__float_as_uint(mx) >> 23) & 0xffu);
        int sh = (ex == 0 || ex == 255) ? 0 : 127 - ex;
        sh = max(-60, min(40, sh));
        sc = __uint_as_float(static_cast<uint32_t>(127 + sh) << 23);
        prev_unscale = __uint_as_float(static_cast<uint32_t>(127 - 2 * sh) << 23);
      }
      // -------------------------------------------------------------- fold + split into the B operand
      for (int item = gt; item < nfr * 14; item += FT_GROUP) {
        const int fl = item / 14, kg = item - fl * 14;
        uint8_t* dst = bbuf + (kg >> 3) * FT_BCHUNK + fl * 128 + (((kg & 7) ^ (fl & 7)) << 4);
        if (kg == 13) {   // k = 104..111: zero matrix columns, but the operand must be finite (the power staging aliased it)
          const uint4 z = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dst) = z;
          *reinterpret_cast<uint4*>(dst + 2 * FT_BCHUNK) = z;
          *reinterpret_cast<uint4*>(dst + 4 * FT_BCHUNK) = z;
          *reinterpret_cast<uint4*>(dst + 6 * FT_BCHUNK) = z;
          continue;
        }
        const float* xf = xs + fl * HOP;
        float x0[8], x1[8], x2[8], x3[8], w0[8], w2[8];
        ft_ld8(xf + 8 * kg, x0);                // x[n + j]
        ft_ld8(xf + 200 + 8 * kg, x3);          // x[200 + n + j]
        ft_ld8_rev(xf + 392 - 8 * kg, x1);      // x[400 - n - j]
        ft_ld8_rev(xf + 192 - 8 * kg, x2);      // x[200 - n - j]
        ft_ld8(S.win + 8 * kg, w0);             // w[n + j] = w[400 - n - j]
        ft_ld8(S.winr + 8 * kg, w2);            // w[200 - n - j] = w[200 + n + j]
        if (kg == 0) { x1[0] = 0.f; x3[0] = 0.f; }       // n = 0: x[400] is not part of the frame, x[200] counts once
        if (kg == 12) { x2[4] = 0.f; x3[4] = 0.f; }      // n = 100 is its own mirror image
        float v0[8], v1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float wa = w0[j] * sc, wb2 = w2[j] * sc;
          if (role == 0) {
            const float an = (x0[j] + x1[j]) * wa, am = (x2[j] + x3[j]) * wb2;
            v0[j] = an + am;
            v1[j] = an - am;
          } else {
            const float bn = (x0[j] - x1[j]) * wa, bm = (x2[j] - x3[j]) * wb2;
            v0[j] = bn - bm;
            v1[j] = bn + bm;
          }
        }
        uint4 h0, l0, h1, l1;
        ft_split8(v0, h0, l0);
        ft_split8(v1, h1, l1);
        *reinterpret_cast<uint4*>(dst) = h0;
        *reinterpret_cast<uint4*>(dst + 2 * FT_BCHUNK) = l0;
        *reinterpret_cast<uint4*>(dst + 4 * FT_BCHUNK) = h1;
        *reinterpret_cast<uint4*>(dst + 6 * FT_BCHUNK) = l1;
      }
      fence_proxy_async();
      tc_fence_before();
      ft_group_sync(g);
      if (gt == 0) mbar_arrive(&S.b_full[g]);
      prev_b = cur_b;
      prev_f0 = cur_f0;
      // -------------------------------------------------------------- next span into registers (lands under the MMAs)
      if (t + 1 < n_grp) {
        cur_b = wb_;
        cur_f0 = wc_ * FT_N;
        ft_load_span<VARIANT>(span, wav_of(cur_b), cur_f0, n_samples, gt);
        advance();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------ kernel 2
__device__ __forceinline__ float block_max_256(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < 8; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}

// raw [B,T,80] f32 -> (clamp at max - top_db) -> (minus time mean) -> f32 in place and/or
// f16 channels-last padded [B, Tp, 128] with reflect halo H (padded rows beyond the halo and
// channels 80..127 are zero).
__global__ void __launch_bounds__(256)
fbank_norm_kernel(float* raw, int T, int use_top_db, int mean_norm,
                  float* out_f32, __half* __restrict__ out_f16, int Tp, int H, int from_power,
                  const float* __restrict__ part_sin) {
  __shared__ float red[8];
  __shared__ float part[3][NMEL];
  __shared__ float mean_s[NMEL];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* x = raw + static_cast<size_t>(b) * T * NMEL;
  const int total = T * NMEL;
  float floor_v = -INFINITY;
  if (from_power) {
    // the tensor-core front end leaves the two halves of the mel POWER (sum w Re^2 here, sum w Im^2 in part_sin);
    // add them and take the log in place: 1 = speechbrain 10 log10(max(x, 1e-10)), 2 = torchaudio log(x + 1e-6)
    const float* x2 = part_sin + static_cast<size_t>(b) * T * NMEL;
    float mx = -INFINITY;
    for (int i = tid; i < total; i += 256) {
      const float pw = x[i] + x2[i];
      const float v = from_power == 1 ? 10.0f * log10f(fmaxf(pw, 1e-10f)) : logf(pw + 1e-6f);
      x[i] = v;
      mx = fmaxf(mx, v);
    }
    __syncthreads();
    if (use_top_db) floor_v = block_max_256(mx, red) - 80.0f;
  } else if (use_top_db) {
    float mx = -INFINITY;
    for (int i = tid; i < total; i += 256) mx = fmaxf(mx, x[i]);
    floor_v = block_max_256(mx, red) - 80.0f;
  }
  if (tid < 240) {
    const int g = tid / NMEL, m = tid % NMEL;
    float s = 0.f;
    if (mean_norm) {
      // eight frames' loads in flight, two partial sums (the one-load-at-a-time loop was a chain of ~50 L1/L2 round trips)
      float s2 = 0.f;
      int t = g;
      for (; t + 21 < T; t += 24) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = x[(t + 3 * e) * NMEL + m];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          s += fmaxf(v[e], floor_v);
          s2 += fmaxf(v[e + 1], floor_v);
        }
      }
      for (; t < T; t += 3) s += fmaxf(x[t * NMEL + m], floor_v);
      s += s2;
    }
    part[g][m] = s;
  }
  __syncthreads();
  if (tid < NMEL) mean_s[tid] = mean_norm ? (part[0][tid] + part[1][tid] + part[2][tid]) / static_cast<float>(T) : 0.f;
  __syncthreads();
  if (out_f32 != nullptr) {
    float* o = out_f32 + static_cast<size_t>(b) * T * NMEL;
    for (int i = tid; i < total; i += 256) o[i] = fmaxf(x[i], floor_v) - mean_s[i % NMEL];
  }
  if (out_f16 != nullptr) {
    __half* o = out_f16 + static_cast<size_t>(b) * Tp * 128;
    // one thread per (row, pair of channels): 64 half2 per row
    for (int i = tid; i < Tp * 64; i += 256) {
      const int p = i >> 6, c = (i & 63) * 2;
      int t = p - H;
      if (t < 0) t = -t;
      else if (t >= T) t = 2 * (T - 1) - t;
      float v0 = 0.f, v1 = 0.f;
      if (c < NMEL && t >= 0 && t < T && p < T + 2 * H) {
        v0 = fmaxf(x[t * NMEL + c], floor_v) - mean_s[c];
        v1 = fmaxf(x[t * NMEL + c + 1], floor_v) - mean_s[c + 1];
      }
      reinterpret_cast<__half2*>(o)[i] = __floats2half2_rn(v0, v1);
    }
  }
}

// ------------------------------------------------------------------- host side
static void build_tables(int variant, FbankTables& t) {
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < NFFT; ++n) {
    // torch.hann_window / torch.hamming_window, periodic=True
    const double c = cos(2.0 * PI * n / NFFT);
    t.window[n] = static_cast<float>(variant == 0 ? 0.5 - 0.5 * c : 0.54 - 0.46 * c);
  }
  for (int n2 = 0; n2 < 25; ++n2)
    for (int k1 = 0; k1 < 8; ++k1) {
      const double a = -2.0 * PI * (n2 * k1) / 200.0;
      t.tw200[k1 * 25 + n2] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
  for (int k = 0; k < NBIN; ++k) {
    const double a = -2.0 * PI * k / 400.0;
    t.tw400[k] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
  }
  for (int j = 0; j < 25; ++j) {
    const double a = -2.0 * PI * j / 25.0;
    t.tw25[j] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
  }
  // mel filterbank [201][80]
  std::vector<double> fb(NBIN * NMEL, 0.0);
  const double f_min = variant == 0 ? 20.0 : 0.0, f_max = variant == 0 ? 7900.0 : 8000.0;
  auto to_mel = [](double hz) { return 2595.0 * log10(1.0 + hz / 700.0); };
  auto to_hz = [](double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1.0); };
  double fpts[NMEL + 2];
  for (int i = 0; i < NMEL + 2; ++i) {
    // torch.linspace evaluates in f32; mirror that rounding of the mel grid
    const float mel = static_cast<float>(to_mel(f_min) + (to_mel(f_max) - to_mel(f_min)) * i / (NMEL + 1));
    fpts[i] = static_cast<float>(to_hz(mel));
  }
  for (int k = 0; k < NBIN; ++k) {
    const double freq = 8000.0 * k / (NBIN - 1);
    for (int m = 0; m < NMEL; ++m) {
      double v;
      if (variant == 0) {
        // torchaudio.functional.melscale_fbanks (functional.py:507-513), norm=None
        const double down = (freq - fpts[m]) / (fpts[m + 1] - fpts[m]);
        const double up = (fpts[m + 2] - freq) / (fpts[m + 2] - fpts[m + 1]);
        v = fmax(0.0, fmin(down, up));
      } else {
        // speechbrain Filterbank._triangular_filters: both sides use the LEFT band width
        const double band = fpts[m + 1] - fpts[m];
        const double slope = (freq - fpts[m + 1]) / band;
        v = fmax(0.0, fmin(slope + 1.0, -slope + 1.0));
      }
      fb[k * NMEL + m] = v;
    }
  }
  for (int m = 0; m < NMEL; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBIN; ++k)
      if (static_cast<float>(fb[k * NMEL + m]) > 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    int len = lo < 0 ? 0 : hi - lo + 1;
    if (len > MEL_MAXLEN) len = MEL_MAXLEN;  // cannot happen for these two banks (max 13)
    t.mel_start[m] = lo < 0 ? 0 : lo;
    t.mel_len[m] = len;
    for (int j = 0; j < MEL_MAXLEN; ++j)
      t.mel_w[j * NMEL + m] = j < len ? static_cast<float>(fb[(lo + j) * NMEL + m]) : 0.f;
  }
}

static FbankTables* g_tab[64][2] = {};   // per device, per variant
static std::mutex g_tab_mu;

static int get_tables(int variant, FbankTables** out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!g_tab[dev][variant]) {
    FbankTables* h = new FbankTables;
    build_tables(variant, *h);
    FbankTables* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(FbankTables));
    if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(FbankTables), cudaMemcpyHostToDevice);
    // the tensor-core kernel reads the banded mel filters from the constant bank (this device's copy of the symbols)
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_mel_start, h->mel_start, sizeof(h->mel_start), variant * sizeof(h->mel_start));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_mel_len, h->mel_len, sizeof(h->mel_len), variant * sizeof(h->mel_len));
    if (e == cudaSuccess) {
      std::vector<float> wt(NMEL * MEL_MAXLEN);
      for (int m = 0; m < NMEL; ++m)
        for (int j = 0; j < MEL_MAXLEN; ++j) wt[m * MEL_MAXLEN + j] = h->mel_w[j * NMEL + m];
      e = cudaMemcpyToSymbol(c_mel_wt, wt.data(), wt.size() * sizeof(float), variant * wt.size() * sizeof(float));
    }
    delete h;
    if (e != cudaSuccess) return fail(SD_ERR_CUDA, "fbank tables: %s", cudaGetErrorString(e));
    g_tab[dev][variant] = d;
  }
  *out = g_tab[dev][variant];
  return SD_OK;
}

// The four folded DFT matrices per role, hi / lo split, laid out as the exact shared-memory image the kernel
// wants: K-major rows of 64 f16 with the 128-byte swizzle (16-byte piece index ^= row & 7), two K chunks per matrix.
static void build_wimg(std::vector<uint8_t>& img) {
  const double PI = 3.14159265358979323846;
  img.assign(2 * static_cast<size_t>(FT_WROLE), 0);
  for (int role = 0; role < 2; ++role)
    for (int j = 0; j < 2; ++j)
      for (int m = 0; m < FT_WROWS; ++m)
        for (int k = 0; k < 16 * FT_KSTEPS; ++k) {
          double v = 0.0;
          if (role == 0 && j == 0) { if (m <= 100 && k <= 100) v = cos(PI * ((m * k) % 200) / 100.0); }
          if (role == 0 && j == 1) { if (m <= 99 && k <= 99) v = cos(PI * (((2 * m + 1) * k) % 400) / 200.0); }
          if (role == 1 && j == 0) { if (m <= 100 && k >= 1 && k <= 99) v = sin(PI * ((m * k) % 200) / 100.0); }
          if (role == 1 && j == 1) { if (m <= 99 && k >= 1 && k <= 100) v = sin(PI * (((2 * m + 1) * k) % 400) / 200.0); }
          const __half hi = __float2half_rn(static_cast<float>(v));
          const __half lo = __float2half_rn(static_cast<float>((v - static_cast<double>(__half2float(hi))) * FT_LO_SCALE));
          const size_t at = static_cast<size_t>(role) * FT_WROLE + (k >> 6) * FT_WCHUNK + m * 128 +
                            ((((k & 63) >> 3) ^ (m & 7)) << 4) + (k & 7) * 2;
          memcpy(&img[at + static_cast<size_t>(2 * j) * FT_WMAT], &hi, 2);
          memcpy(&img[at + static_cast<size_t>(2 * j + 1) * FT_WMAT], &lo, 2);
        }
}

static uint8_t* g_wimg[64] = {};

static int get_wimg(uint8_t** out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!g_wimg[dev]) {
    std::vector<uint8_t> img;
    build_wimg(img);
    uint8_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, img.size());
    if (e == cudaSuccess) e = cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(SD_ERR_CUDA, "fbank DFT matrices: %s", cudaGetErrorString(e));
    g_wimg[dev] = d;
  }
  *out = g_wimg[dev];
  return SD_OK;
}

// which >= 0 selects the frames kernel (0 = FFT on the FP32 pipe, 1 = tensor-core DFT); returns the current
// choice.  Default: $SD_FBANK_TC, else 0 (the FFT kernel is the faster one).
int fbank_kernel_choice(int which) {
  static int choice = [] {
    const char* e = getenv("SD_FBANK_TC");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  if (which == 0 || which == 1) choice = which;
  return choice;
}

int fbank_launch(const float* wav, long wav_stride, int B, int n_samples, int variant,
                 int mean_norm, float* raw, float* out_f32, __half* out_f16, int Tp, int H,
                 cudaStream_t stream, const long* offsets, float* raw2) {
  if (!wav || !raw || B < 1 || n_samples < NFFT || (variant != 0 && variant != 1))
    return fail(SD_ERR_ARG, "fbank: bad arguments (B=%d n=%d variant=%d)", B, n_samples, variant);
  if (B > 65535) return fail(SD_ERR_ARG, "fbank: B=%d exceeds 65535 windows per call", B);
  const int T = 1 + n_samples / HOP;
  FbankTables* tab = nullptr;
  SD_TRY(get_tables(variant, &tab));
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63]) {   // per device
    SD_CUDA_OK(cudaFuncSetAttribute(fbank_frames_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(FbankSmem))));
    SD_CUDA_OK(cudaFuncSetAttribute(fbank_frames_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(FbankSmem))));
    attr[dev & 63] = true;
  }
  if (fbank_kernel_choice(-1) == 1) {
    // tensor-core DFT: both roles add their half of the mel power into a zeroed buffer
    uint8_t* wimg = nullptr;
    SD_TRY(get_wimg(&wimg));
    static bool attr_tc[64] = {};
    static int sms[64] = {};
    if (!attr_tc[dev & 63]) {
      SD_CUDA_OK(cudaFuncSetAttribute(fbank_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(sizeof(FtSmem) + 1024)));
      SD_CUDA_OK(cudaFuncSetAttribute(fbank_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(sizeof(FtSmem) + 1024)));
      SD_CUDA_OK(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
      attr_tc[dev & 63] = true;
    }
    const long total = static_cast<long>(B) * ((T + FT_N - 1) / FT_N);
    const int G = static_cast<int>(std::min<long>(std::max(1, sms[dev & 63] / 2), total));
    // the sine CTAs' half of the mel power: the caller's second scratch, else a stream-ordered temporary
    float* part_sin = raw2;
    if (!part_sin) {
      // own pool that keeps its memory across synchronisations (the default pool trims to zero at every sync)
      static cudaMemPool_t pool[64] = {};
      {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        if (!pool[dev & 63]) {
          cudaMemPoolProps props = {};
          props.allocType = cudaMemAllocationTypePinned;
          props.location.type = cudaMemLocationTypeDevice;
          props.location.id = dev;
          SD_CUDA_OK(cudaMemPoolCreate(&pool[dev & 63], &props));
          unsigned long long keep = ~0ull;
          SD_CUDA_OK(cudaMemPoolSetAttribute(pool[dev & 63], cudaMemPoolAttrReleaseThreshold, &keep));
        }
      }
      SD_CUDA_OK(cudaMallocFromPoolAsync(&part_sin, static_cast<size_t>(B) * T * NMEL * sizeof(float), pool[dev & 63],
                                         stream));
    }
    if (variant == 0)
      fbank_tc_kernel<0><<<2 * G, FT_THREADS, sizeof(FtSmem) + 1024, stream>>>(wav, wav_stride, offsets, n_samples, T, B,
                                                                               tab, wimg, raw, part_sin);
    else
      fbank_tc_kernel<1><<<2 * G, FT_THREADS, sizeof(FtSmem) + 1024, stream>>>(wav, wav_stride, offsets, n_samples, T, B,
                                                                               tab, wimg, raw, part_sin);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) {
      fbank_norm_kernel<<<B, 256, 0, stream>>>(raw, T, variant == 1, mean_norm, out_f32, out_f16, Tp, H,
                                               variant == 1 ? 1 : 2, part_sin);
      le = cudaGetLastError();
    }
    if (!raw2) cudaFreeAsync(part_sin, stream);
    SD_CUDA_OK(le);
    count_launch(2);
    return SD_OK;
  }
  static int sms_fft[64] = {};
  if (!sms_fft[dev & 63]) SD_CUDA_OK(cudaDeviceGetAttribute(&sms_fft[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const long chunks = static_cast<long>(B) * ((T + FR_PER_CTA - 1) / FR_PER_CTA);
  const int grid = static_cast<int>(std::min<long>(chunks, 2L * sms_fft[dev & 63]));   // persistent: two CTAs per SM
  if (variant == 0)
    fbank_frames_kernel<0><<<grid, 256, sizeof(FbankSmem), stream>>>(wav, wav_stride, offsets, n_samples, T, B, tab, raw);
  else
    fbank_frames_kernel<1><<<grid, 256, sizeof(FbankSmem), stream>>>(wav, wav_stride, offsets, n_samples, T, B, tab, raw);
  SD_CUDA_OK(cudaGetLastError());
  fbank_norm_kernel<<<B, 256, 0, stream>>>(raw, T, variant == 1, mean_norm, out_f32, out_f16, Tp, H, 0, nullptr);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return SD_OK;
}

int feats_to_padded_f16(const float* feats, int B, int T, __half* out_f16, int Tp, int H,
                        cudaStream_t stream) {
  fbank_norm_kernel<<<B, 256, 0, stream>>>(const_cast<float*>(feats), T, 0, 0, nullptr, out_f16, Tp, H, 0, nullptr);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

}  // namespace sd

extern "C" int sd_fbank_num_frames(int n_samples) { return n_samples < 0 ? 0 : 1 + n_samples / sd::HOP; }

extern "C" int sd_fbank_f32(const float* wav_dev, long wav_stride, int B, int n_samples, int variant,
                            int mean_norm, float* out_dev, void* stream) {
  if (!out_dev) return sd::fail(SD_ERR_ARG, "sd_fbank_f32: out_dev is NULL");
  // raw log-mel goes straight into out_dev; the normalisation pass rewrites it in place
  return sd::fbank_launch(wav_dev, wav_stride, B, n_samples, variant, mean_norm, out_dev, out_dev,
                          nullptr, 0, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int sd_fbank_kernel(int which) { return sd::fbank_kernel_choice(which); }
