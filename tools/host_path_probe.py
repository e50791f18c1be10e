"""Per-call wall clock of ecapa_encode_batch (B = 512 x 1.5 s) for pinned / pageable host memory, per setting of
SD_ECAPA_HOST_THREADS / SD_ECAPA_STAGING / SD_ECAPA_UPCHUNKS (each setting in its own process)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, json
sys.path.insert(0, %r)
import numpy as np, torch
from speech_diarization_b200 import speech_encode, vad
from speech_diarization_b200.weights import random_ecapa_state_dict
speech_encode.register_ecapa_state_dict(random_ecapa_state_dict(0))
rng = np.random.default_rng(0)
y = (0.1 * rng.standard_normal(16000 * 1200)).astype(np.float32)
pin = torch.from_numpy(y).pin_memory()
fr_pin = vad.frame_audio(pin.numpy(), 16000, 1500.0, 750.0)
fr_page = vad.frame_audio(y, 16000, 1500.0, 750.0)
batches = [np.ascontiguousarray(fr_page[b * 512:(b + 1) * 512]) for b in range(3)]
def run(get, n=12):
    for i in range(3): speech_encode.ecapa_encode_batch(get(i))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): speech_encode.ecapa_encode_batch(get(i))
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
out = {"pinned": run(lambda i: fr_pin[(i %% 3) * 512:(i %% 3 + 1) * 512]),
       "page_strided": run(lambda i: fr_page[(i %% 3) * 512:(i %% 3 + 1) * 512]),
       "page_batch": run(lambda i: batches[i %% 3])}
t0 = time.perf_counter(); z = np.empty_like(batches[0]); 
for _ in range(5): np.copyto(z, batches[0])
out["memcpy_GBs_1thread"] = 5 * batches[0].nbytes / (time.perf_counter() - t0) / 1e9
out["cpus"] = os.cpu_count()
print(json.dumps({k: round(v, 3) for k, v in out.items()}))
''' % ROOT
for spec in sys.argv[1:] or ["default="]:
    name, rest = spec.split("=", 1)
    env = dict(os.environ)
    for kv in [p for p in rest.split(",") if p]:
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{name:12s}", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-300:], flush=True)
