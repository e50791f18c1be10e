"""Per-call cost of the embedding API on variable-length batches of 32 (the reference's embed_segments pattern)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
dev = torch.device("cuda:0")
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=512, max_samples=24000)
rng = np.random.default_rng(0)
lens = rng.integers(8000, 80000, 60)          # 0.5 .. 5 s
audio = torch.randn(32, 80000, device=dev) * 0.05
torch.cuda.synchronize()
for name, ls in (("first use of each shape", lens), ("same shapes again", lens), ("third pass", lens)):
    t0 = time.perf_counter()
    for n in ls:
        x = audio[:, :int(n)].contiguous()
        e = enc.embed_device(x, x.stride(0), 32, int(n))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {1e3*dt/len(ls):.3f} ms per batch of 32 (mean length {ls.mean()/16000:.2f} s) -> {32*len(ls)/dt:.0f} emb/s")
