"""AHC timing alone (CUDA events around sd_ahc_average on a resident distance matrix), 3 repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import clustering
clustering.ahc_keep_stats(True)
dev = torch.device("cuda:0")
for N in (int(a) for a in (sys.argv[1:] or ["5000", "20000"])):
    rng = np.random.default_rng(0)
    c = rng.standard_normal((8, 192)); c /= np.linalg.norm(c, axis=1, keepdims=True)
    X = (c[rng.integers(0, 8, N)] + 0.02 * rng.standard_normal((N, 192))).astype(np.float32)
    dist = clustering.cosine_distance_device(torch.from_numpy(X).to(dev))
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); labels, ncl = clustering.ahc_average_device(dist, 1 - 0.68); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    print("U", os.environ.get("SD_AHC_U", "-"), "N", N, "ahc ms", ts, "clusters", int(ncl.item()), clustering.ahc_last_stats())
