"""Timing of the GPU centroid linkage vs scipy (scipy 2.26 s @5k, 60.3 s @20k)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_diarization_b200 import diarization_baseline as db
for N in (int(a) for a in (sys.argv[1:] or ["2000", "5000", "20000"])):
    rng = np.random.default_rng(0)
    c = rng.standard_normal((8, 192)); c /= np.linalg.norm(c, axis=1, keepdims=True)
    X = (c[rng.integers(0, 8, N)] + 0.02 * rng.standard_normal((N, 192))).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    db.linkage_centroid(X[:64])
    torch.cuda.synchronize(); t0 = time.perf_counter(); Z = db.linkage_centroid(X); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); lab = db.AgglomerativeClustering(threshold=0.70).cluster(X); tc = time.perf_counter() - t0
    line = f"N={N}: GPU linkage {tg*1e3:.1f} ms, full pyannote-style cluster() {tc*1e3:.1f} ms, clusters {len(set(lab.tolist()))}"
    if N <= 5000:
        from scipy.cluster.hierarchy import linkage
        t0 = time.perf_counter(); Zr = linkage(X, "centroid", "euclidean"); ts = time.perf_counter() - t0
        line += f", scipy linkage {ts*1e3:.0f} ms, merges equal {np.array_equal(Z[:, [0, 1, 3]], Zr[:, [0, 1, 3]])}"
    print(line)
