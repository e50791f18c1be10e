"""Debug: per-conv clock64 stamps of CTA 0 of res2net_fused_kernel (SD_R2_TRACE).  Columns per conv:
MMA start, MMA issue done, then (t_full seen, epilogue done) for each of the 8 epilogue warps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SD_ECAPA_GRAPH"] = "0"
os.environ["SD_R2_TRACE"] = "gpurun_out/r2_trace.txt"
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
dev = torch.device("cuda:0")
B = 512
audio = (0.1 * torch.randn((B - 1) * 12000 + 24000, device=dev)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=B, max_samples=24000)
for _ in range(2):
    enc.embed_device(audio, 12000, B, 24000)
torch.cuda.synchronize()
rows = [[int(v) for v in l.split()] for l in open("gpurun_out/r2_trace.txt")]
print("conv  mma_start mma_done | t_full(min..max)  epi_done per warp (cycles, relative to first stamp)")
for n, r in enumerate(rows[:14]):
    tf = r[2::2]; ed = r[3::2]
    print(f"{n:3d} {r[0]:9d} {r[1]:9d} | {min(tf):8d}..{max(tf):8d} | " + " ".join(f"{e - min(tf):6d}" for e in ed))
print("warp 4 chunk stamps (conv, chunk): tmem_ld issue -> ld done -> math done -> staged -> A written -> write-out done")
print("MMA thread, per weight box: (wait for box, MMAs issued until next wait)")
for n in range(2, 6):
    r = rows[32 + n]
    print(n, " ".join(f"({r[2*b+1]-r[2*b]:5d},{(r[2*b+2] if b < 5 else rows[n][1]) - r[2*b+1]:5d})" for b in range(6)))
for n in range(2, 8):
    for k in range(4):
        r = rows[64 + n * 4 + k]
        if r[0] < 0: continue
        print(n, k, " ".join(f"{r[j + 1] - r[j]:6d}" for j in range(5)), " total", r[5] - r[0])
