"""Bit-equality of the embeddings of two builds of the library (each in its own process): python tools/ab_equal.py a.so b.so"""
import os, subprocess, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib
sys.path.insert(0, %r)
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
g = torch.Generator(device="cuda:0").manual_seed(3)
audio = (0.1 * torch.randn(511 * 12000 + 24000, device="cuda:0", generator=g)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device="cuda:0", max_batch=512, max_samples=24000)
e = enc.embed_device(audio, 12000, 512, 24000)
e2 = enc.embed_device(audio[:100 * 8000 + 16000], 8000, 100, 16000)
print(hashlib.sha256(e.cpu().numpy().tobytes()).hexdigest(), hashlib.sha256(e2.cpu().numpy().tobytes()).hexdigest(), float(e.abs().mean()))
''' % ROOT
outs = []
for lib in sys.argv[1:]:
    env = dict(os.environ, SD_LIB_PATH=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-300:]
    print(lib, line, flush=True)
    outs.append(line)
print("EQUAL" if len(set(outs)) == 1 else "DIFFERENT")
