"""A/B timing of library builds / environment switches on one GPU.
    python tools/ab_probe.py name=path/to/lib.so[,ENV=V,...] ...
Each variant runs in its own process: graph-replayed 512-window step (the bench's `value` path) + eager per-stage ms."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
dev = torch.device("cuda:0"); B = 512
g = torch.Generator(device=dev).manual_seed(0)
audio = (0.1 * torch.randn(57_600_000, device=dev, generator=g)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=B, max_samples=24000)
out = torch.empty((B, 192), device=dev)
def step(i):
    enc.embed_device(audio[(i %% 9) * B * 12000:], 12000, B, 24000, l2_normalize=True, out=out)
for i in range(5): step(i)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): step(i)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
enc.profile(True)
for i in range(10): step(i)
torch.cuda.synchronize(); ms, n = enc.profile_read()
st = {k: round(v / n, 4) for k, v in ms.items()}
agg = {"fbank": st["fbank"], "block0": st["block0"], "tdnn1": round(sum(st[f"b{b}.tdnn1"] for b in (1,2,3)), 4),
       "res2net": round(sum(st[f"b{b}.res2net"] for b in (1,2,3)), 4), "tdnn2": round(sum(st[f"b{b}.tdnn2"] for b in (1,2,3)), 4),
       "se": round(sum(st[f"b{b}.se"] for b in (1,2,3)), 4), "mfa": st["mfa"], "ctx": st["asp.context"], "attn": st["asp.attn"],
       "pool": st["asp.pool"], "fc": st["fc"]}
print(json.dumps({"step_ms": round(best, 4), "eager_sum": round(sum(st.values()), 4), **agg}))
''' % ROOT
for spec in sys.argv[1:]:
    name, rest = spec.split("=", 1)
    parts = rest.split(",")
    env = dict(os.environ)
    if parts[0]:
        env["SD_LIB_PATH"] = os.path.join(ROOT, parts[0])
    for kv in parts[1:]:
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:]
    print(f"{name:14s} {line}", flush=True)
