"""Minimal program for ncu captures: two forwards of the BASELINE batch (512 x 1.5 s), one
affinity + AHC at N = 5000 and the rank-4 operators (whitening, AS-norm, Viterbi, VAD mask chain) at N = 5000.  Kernel order per forward: fbank_frames, fbank_norm, block0 GEMM<1,256>,
3 x [tdnn1 <1,256>, 7 x res2net <1,128>, tdnn2 <1,256>, time_mean, se_mlp, se_apply], MFA <1,256>,
time_mean_std, dense_rows, attention <1,128>, pool <2,256>, dense_rows, l2norm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import speech_encode, clustering, postproc
from speech_diarization_b200.weights import random_ecapa_state_dict

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
audio = (0.1 * torch.randn((B - 1) * 12000 + 24000, device=dev)).clamp(-1, 1)
enc = speech_encode.EcapaEncoderB200(random_ecapa_state_dict(0), device=dev, max_batch=B, max_samples=24000)
for _ in range(2):
    e = enc.embed_device(audio, 12000, B, 24000, l2_normalize=True)
torch.cuda.synchronize()
rng = np.random.default_rng(0)
c = rng.standard_normal((8, 192)); c /= np.linalg.norm(c, axis=1, keepdims=True)
X = (c[rng.integers(0, 8, 5000)] + 0.02 * rng.standard_normal((5000, 192))).astype(np.float32)
lab = clustering.cluster_embeddings_device(torch.from_numpy(X).to(dev), 0.68)
torch.cuda.synchronize()
xd = torch.from_numpy(X).to(dev)
wh = postproc.whiten_l2_device(xd)
sc = postproc.asnorm_device(xd, torch.from_numpy(c.astype(np.float32)).to(dev), xd, 200)
path = postproc.viterbi_device(sc, 0.995)
probs = torch.from_numpy(np.clip(np.convolve(rng.random(36008), np.ones(9) / 9, mode="valid") * 1.6 - 0.3, 0, 1).astype(np.float32)).to(dev)
m = postproc.morph_open_close_device(postproc.hysteresis_device(probs, 0.6, 0.4), 8, 4)
segs = postproc.mask_segments_device(m, 25, 10)
torch.cuda.synchronize()
print("ok", float(e.norm(dim=1).mean()), int(lab.max()) + 1, int(path.max()) + 1, len(segs))
