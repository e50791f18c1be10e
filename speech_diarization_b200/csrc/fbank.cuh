// fbank.cuh — internal interface of fbank.cu (used by the ECAPA plan).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

namespace sd {
// raw: [B,T,80] f32 scratch (may alias out_f32).  out_f16: channels-last [B,Tp,128] with
// reflect halo H, or nullptr.  T = 1 + n_samples/160.
// raw2: a second [B,T,80] f32 scratch for the tensor-core kernel's sine half (nullptr: a stream-ordered temporary).
// Window b starts at wav + b * wav_stride, or at wav + offsets[b] (device array, samples) when offsets != nullptr.
int fbank_launch(const float* wav, long wav_stride, int B, int n_samples, int variant,
                 int mean_norm, float* raw, float* out_f32, __half* out_f16, int Tp, int H,
                 cudaStream_t stream, const long* offsets = nullptr, float* raw2 = nullptr);
// 1 = tensor-core DFT frames kernel (default), 0 = FFT on the FP32 pipe; which < 0 only queries
int fbank_kernel_choice(int which);
// precomputed features [B,T,80] f32 -> padded f16 layout (no normalisation applied)
int feats_to_padded_f16(const float* feats, int B, int T, __half* out_f16, int Tp, int H,
                        cudaStream_t stream);
}  // namespace sd
