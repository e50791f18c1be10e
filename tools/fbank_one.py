"""One 512-window fbank call (speechbrain variant) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_diarization_b200 import speech_encode as se
y = (0.1 * torch.randn(511 * 12000 + 24000, device="cuda:0")).clamp(-1, 1)
for _ in range(2):
    out = se.fbank_batch_device(y, variant=1, mean_nor=True, wav_stride=12000, n_windows=512, n_samples=24000)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
