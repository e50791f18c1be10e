"""Host-side timeline of ecapa_encode_batch calls (SD_HOST_TRACE=1) next to the Python-level wall clock per call."""
import os, sys, time
os.environ["SD_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import speech_encode, vad
from speech_diarization_b200.weights import random_ecapa_state_dict
speech_encode.register_ecapa_state_dict(random_ecapa_state_dict(0))
rng = np.random.default_rng(0)
y = (0.1 * rng.standard_normal(16000 * 1200)).astype(np.float32)
pin = torch.from_numpy(y).pin_memory()
fr = vad.frame_audio(y if "--pageable" in sys.argv else pin.numpy(), 16000, 1500.0, 750.0)
for i in range(3): speech_encode.ecapa_encode_batch(fr[i * 512:(i + 1) * 512])
torch.cuda.synchronize()
t_prev = time.perf_counter()
for i in range(6):
    t0 = time.perf_counter()
    speech_encode.ecapa_encode_batch(fr[(i % 3) * 512:(i % 3 + 1) * 512])
    t1 = time.perf_counter()
    print(f"[py] call {i}: {1e6 * (t1 - t0):.0f} us (since previous return {1e6 * (t0 - t_prev):.0f} us)", file=sys.stderr, flush=True)
    t_prev = t1
