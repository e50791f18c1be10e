import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_diarization_b200 import speech_encode
from speech_diarization_b200.weights import random_ecapa_state_dict
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device="cuda:0", max_batch=4, max_samples=8000)
x = torch.randn(2, 26, 80)
e = enc.forward_feats(x)
torch.cuda.synchronize()
print("ok", e.shape, float(e.abs().mean()))
