"""Debug: clock64 stamps of CTA 0 of the cta_group::2 GEMM (SD_GEMM_TRACE) for the tdnn layers (16 k-iterations).
Per tile: MMA thread [wait-for-TMEM start, TMEM free, last MMA issued]; epilogue warps 2 and 9:
[prefetch done, accumulator ready, (barrier) epilogue start, drained+staged, staging barrier, write-out done]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SD_ECAPA_GRAPH"] = "0"
os.environ["SD_GEMM_TRACE"] = "gpurun_out/gemm_trace.txt"
os.environ.setdefault("SD_GEMM_TRACE_K", "16")
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
dev = torch.device("cuda:0"); B = 512
audio = (0.1 * torch.randn((B - 1) * 12000 + 24000, device=dev)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=B, max_samples=24000)
for _ in range(2): enc.embed_device(audio, 12000, B, 24000)
torch.cuda.synchronize()
rows = [[int(v) for v in l.split()] for l in open("gpurun_out/gemm_trace.txt")]
print("tile | MMA: wait_start tmem_free issued(+dur) | warp2: ready  epi_start drained staged_bar written | warp9: ready drained written   (cycles from first stamp)")
for t, r in enumerate(rows[:20]):
    if r[1] < 0: continue
    print(f"{t:3d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} (+{r[2]-r[1]:5d}) | {r[4]:7d} {r[5]:7d} {r[6]:7d} {r[7]:7d} {r[8]:7d} | {r[10]:7d} {r[12]:7d} {r[14]:7d}")
print("per tile: MMA issue span, epilogue drain+math (warp2), barrier wait, write-out; tile period (MMA issued deltas)")
prev = None
for t, r in enumerate(rows[:20]):
    if r[1] < 0: continue
    per = r[2] - prev if prev is not None else -1
    prev = r[2]
    print(f"{t:3d}  mma {r[2]-r[1]:6d}  tmem_wait {r[1]-r[0]:6d} | acc_wait {r[4]-r[3]:6d} epi {r[6]-r[5]:6d} bar {r[7]-r[6]:6d} wout {r[8]-r[7]:6d} | period {per:6d}")
