"""Prints per-stage milliseconds of the ECAPA forward at B=512 (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
dev = torch.device("cuda:0"); B = 512
audio = (0.1 * torch.randn((B - 1) * 12000 + 24000, device=dev)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=B, max_samples=24000)
for _ in range(3): enc.embed_device(audio, 12000, B, 24000, l2_normalize=True)
torch.cuda.synchronize(); enc.profile(True)
for _ in range(10): enc.embed_device(audio, 12000, B, 24000, l2_normalize=True)
torch.cuda.synchronize(); ms, n = enc.profile_read()
tot = sum(ms.values()) / n
print(os.environ.get("SD_DEBUG_EPI", "-"), os.environ.get("SD_ECAPA_CHAIN", "-"), f"total {tot:.3f} ms", {k: round(v / n, 4) for k, v in ms.items()})

# eager vs CUDA-graph replay of the same forward (launch-gap estimate)
def timed(fn, n=20):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = torch.empty((B, 192), device=dev)
eager = timed(lambda: enc.embed_device(audio, 12000, B, 24000, l2_normalize=True, out=out))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    enc.embed_device(audio, 12000, B, 24000, l2_normalize=True, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        enc.embed_device(audio, 12000, B, 24000, l2_normalize=True, out=out)
graph = timed(g.replay)
print(f"eager {eager:.3f} ms   graph replay {graph:.3f} ms")
