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
constexpr int PSTRIDE = 204;  // floats per frame of power
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
  float2 z[FR_PER_CTA * ZSTRIDE];
  float pw[FR_PER_CTA * PSTRIDE];
  FbankTables tab;
};

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
                    const FbankTables* __restrict__ gtab, float* __restrict__ raw) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FbankSmem& S = *reinterpret_cast<FbankSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR_PER_CTA;
  const float* w = wav + (offsets != nullptr ? __ldg(offsets + b) : static_cast<long>(b) * wav_stride);

  // tables -> shared
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gtab);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S.tab);
    for (int i = tid; i < static_cast<int>(sizeof(FbankTables) / 4); i += 256) dst[i] = src[i];
  }
  // ---- step 0: stage the CTA's sample span (32 frames x hop + 240) in shared memory with 128-bit coalesced
  // loads; the centre padding (reflect / zero) is resolved here, once per sample instead of once per FFT input.
  // The span aliases the power buffer, which is not written before step 3.
  float* xs = S.pw;
  {
    const int gs0 = f0 * HOP - NFFT / 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(w + gs0) & 15) == 0;
    for (int i4 = tid; i4 < SPAN / 4; i4 += 256) {
      const int g = gs0 + 4 * i4;
      float4 v;
      if (aligned && g >= 0 && g + 3 < n_samples) {
        v = __ldg(reinterpret_cast<const float4*>(w + g));
      } else {
        v.x = load_sample<VARIANT>(w, g, n_samples);
        v.y = load_sample<VARIANT>(w, g + 1, n_samples);
        v.z = load_sample<VARIANT>(w, g + 2, n_samples);
        v.w = load_sample<VARIANT>(w, g + 3, n_samples);
      }
      reinterpret_cast<float4*>(xs)[i4] = v;
    }
  }
  __syncthreads();

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
        const float2 wn = *reinterpret_cast<const float2*>(&S.tab.window[n]);
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
    for (int k1 = 0; k1 < 8; ++k1) zf[k1 * 25 + n2] = cmul(X[k1], S.tab.tw200[k1 * 25 + n2]);
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
      for (int bb = 1; bb < 5; ++bb) y[5 * c + bb] = cmul(y[5 * c + bb], S.tab.tw25[(bb * c) % 25]);
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

  // ---- step 3: real-input split + power for the warp's 4 frames
  for (int item = lane; item < 4 * NBIN; item += 32) {
    const int fl = warp * 4 + item / NBIN;
    const int k = item % NBIN;
    const float2* zf = S.z + fl * ZSTRIDE;
    const float2 zk = zf[k == 200 ? 0 : k];
    float2 zc = zf[k == 0 ? 0 : 200 - k];
    zc.y = -zc.y;
    const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 dd = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
    const float2 o = mul_mi(dd);
    const float2 x = cadd(e, cmul(S.tab.tw400[k], o));
    S.pw[fl * PSTRIDE + k] = x.x * x.x + x.y * x.y;
  }
  __syncthreads();

  // ---- step 4: banded mel sum + log, coalesced store of [frame][80]
  for (int item = tid; item < FR_PER_CTA * NMEL; item += 256) {
    const int fl = item / NMEL, m = item % NMEL;
    const int f = f0 + fl;
    if (f >= T) break;
    const float* p = S.pw + fl * PSTRIDE + S.tab.mel_start[m];
    const float* mw = S.tab.mel_w + m;
    const int len = S.tab.mel_len[m];
    float acc = 0.f;
    for (int j = 0; j < len; ++j) acc = fmaf(p[j], mw[j * NMEL], acc);
    float v;
    if (VARIANT == 0) v = logf(acc + 1e-6f);
    else v = 10.0f * log10f(fmaxf(acc, 1e-10f));
    raw[(static_cast<size_t>(b) * T + f) * NMEL + m] = v;
  }
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
                  float* out_f32, __half* __restrict__ out_f16, int Tp, int H) {
  __shared__ float red[8];
  __shared__ float part[3][NMEL];
  __shared__ float mean_s[NMEL];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* x = raw + static_cast<size_t>(b) * T * NMEL;
  const int total = T * NMEL;
  float floor_v = -INFINITY;
  if (use_top_db) {
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
    delete h;
    if (e != cudaSuccess) return fail(SD_ERR_CUDA, "fbank tables: %s", cudaGetErrorString(e));
    g_tab[dev][variant] = d;
  }
  *out = g_tab[dev][variant];
  return SD_OK;
}

int fbank_launch(const float* wav, long wav_stride, int B, int n_samples, int variant,
                 int mean_norm, float* raw, float* out_f32, __half* out_f16, int Tp, int H,
                 cudaStream_t stream, const long* offsets) {
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
  dim3 grid((T + FR_PER_CTA - 1) / FR_PER_CTA, B);
  if (variant == 0)
    fbank_frames_kernel<0><<<grid, 256, sizeof(FbankSmem), stream>>>(wav, wav_stride, offsets, n_samples, T, tab, raw);
  else
    fbank_frames_kernel<1><<<grid, 256, sizeof(FbankSmem), stream>>>(wav, wav_stride, offsets, n_samples, T, tab, raw);
  SD_CUDA_OK(cudaGetLastError());
  fbank_norm_kernel<<<B, 256, 0, stream>>>(raw, T, variant == 1, mean_norm, out_f32, out_f16, Tp, H);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return SD_OK;
}

int feats_to_padded_f16(const float* feats, int B, int T, __half* out_f16, int Tp, int H,
                        cudaStream_t stream) {
  fbank_norm_kernel<<<B, 256, 0, stream>>>(const_cast<float*>(feats), T, 0, 0, nullptr, out_f16, Tp, H);
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
