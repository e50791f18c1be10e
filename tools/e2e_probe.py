"""Wall-clock per call of ecapa_encode_batch on 512 windows of pinned host audio (the bench's e2e leg) under the
switches of sd_ecapa_embed_host: SD_ECAPA_PIPE (front per upload chunk), SD_DEBUG_NOCOPY (no H2D at all)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import speech_encode, vad
from speech_diarization_b200.weights import random_ecapa_state_dict
speech_encode.register_ecapa_state_dict(random_ecapa_state_dict(0))
n = 16000 * 600
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.copy_((0.1 * torch.randn(n)).clamp(-1, 1))
frames = vad.frame_audio(host.numpy(), 16000, 1500.0, 750.0)
B = 512
for i in range(4): speech_encode.ecapa_encode_batch(frames[:B])
torch.cuda.synchronize()
ts = []
for i in range(30):
    b = (i % (len(frames) // B))
    t0 = time.perf_counter(); speech_encode.ecapa_encode_batch(frames[b * B:(b + 1) * B]); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"PIPE={os.environ.get('SD_ECAPA_PIPE','-')} NOCOPY={os.environ.get('SD_DEBUG_NOCOPY','-')}  median {np.median(ts):.3f} ms  min {ts.min():.3f}  mean {ts.mean():.3f}")
