"""Debug: per-job clock64 stamps of CTA 0 of res2net_pipe_kernel (SD_R2_TRACE).  Columns per job: MMA start, MMA
issued, then for the group's first and last warp: wait start, t_full seen, tail done, chunk 0 done, arrivals, end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SD_ECAPA_GRAPH"] = "0"
os.environ["SD_R2_TRACE"] = "gpurun_out/r2p_trace.txt"
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
rows = [[int(v) for v in l.split()] for l in open("gpurun_out/r2p_trace.txt")]
print("job  mma_start mma_issued | w0: wait t_full tail chunk0 arrive end | w7: wait t_full tail chunk0 arrive end")
for n, r in enumerate(rows[:32]):
    print(f"{n:3d} {r[0]:8d} {r[1]:8d} | " + " ".join(f"{v:7d}" for v in r[2:9]) + " | " + " ".join(f"{v:7d}" for v in r[9:16]))
