"""A/B of the two fbank frames kernels (sd_fbank_kernel 0 = FFT on the FP32 pipe, 1 = tensor-core DFT):
max |difference| between them and against the oracle, and CUDA-event times at the BASELINE batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from speech_diarization_b200 import _lib, speech_encode as se
from conftest import synth_wave

lib = _lib.load()
dev = torch.device("cuda:0")


def run(which, w, variant, mean_nor, **kw):
    lib.sd_fbank_kernel(which)
    return se.fbank_batch_device(w, variant=variant, mean_nor=mean_nor, **kw)


for variant in (0, 1):
    for n in (400, 401, 559, 4000, 16000, 24000, 48123):
        for scale in (1.0, 1e-4, 300.0):
            w = torch.from_numpy(synth_wave(3, n, n) * np.float32(scale)).to(dev)
            a = run(0, w, variant, False)
            b = run(1, w, variant, False)
            d = (a - b).abs().max().item()
            print(f"variant {variant} n {n:6d} scale {scale:8.1e}: max|fft - tc| = {d:.3e}  finite {bool(torch.isfinite(b).all())}")
    from oracle import fbank_oracle as fo, ecapa_oracle as eo
    w = synth_wave(4, 24000, 3)
    ref = fo.fbank_batch(w) if variant == 0 else eo.fbank_speechbrain(torch.from_numpy(w), mean_norm=True).numpy()
    for which in (0, 1):
        got = run(which, torch.from_numpy(w).to(dev), variant, True).cpu().numpy()
        print(f"variant {variant} kernel {which}: max|got - oracle| = {np.abs(got - ref).max():.3e}")

# timing at the BASELINE batch
y = (0.1 * torch.randn(511 * 12000 + 24000, device=dev)).clamp(-1, 1)
for which in (0, 1, 0, 1):
    lib.sd_fbank_kernel(which)
    for _ in range(3):
        se.fbank_batch_device(y, variant=1, mean_nor=True, wav_stride=12000, n_windows=512, n_samples=24000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        se.fbank_batch_device(y, variant=1, mean_nor=True, wav_stride=12000, n_windows=512, n_samples=24000)
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel {which}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per 512-window fbank (frames + norm + output alloc)")
lib.sd_fbank_kernel(1)
