"""Affinity + AHC at N = 50 000 (BASELINE config 4's largest size): correctness vs the planted partition + timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import clustering
clustering.ahc_keep_stats(True)
dev = torch.device("cuda:0")
def same_partition(a, b):
    fwd, bwd = {}, {}
    for x, y in zip(a.tolist(), b.tolist()):
        if fwd.setdefault(x, y) != y or bwd.setdefault(y, x) != x: return False
    return True
for N in (int(a) for a in (sys.argv[1:] or ["50000"])):
    rng = np.random.default_rng(0)
    c = rng.standard_normal((8, 192)); c /= np.linalg.norm(c, axis=1, keepdims=True)
    lab = rng.integers(0, 8, N)
    X = (c[lab] + 0.02 * rng.standard_normal((N, 192))).astype(np.float32)
    xd = torch.from_numpy(X).to(dev)
    for rep in range(2):
        torch.cuda.synchronize(); e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); dist = clustering.cosine_distance_device(xd); e[1].record()
        labels, ncl = clustering.ahc_average_device(dist, 1 - 0.68); e[2].record(); torch.cuda.synchronize()
        print(N, "affinity ms", round(e[0].elapsed_time(e[1]), 3), "ahc ms", round(e[1].elapsed_time(e[2]), 2), "clusters", int(ncl.item()),
              "match planted", same_partition(labels.cpu().numpy(), lab), clustering.ahc_last_stats(), "mem GB", round(torch.cuda.max_memory_allocated() / 1e9, 1))
        del dist, labels; torch.cuda.empty_cache()
