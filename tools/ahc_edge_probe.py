"""Debug: isolate why the sigma=0.045 N=20k case gives 39 clusters on the GPU vs 38 from sklearn."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import synth_emb
from sklearn.metrics.pairwise import cosine_similarity
from scipy.cluster.hierarchy import linkage
from speech_diarization_b200 import clustering as cl
X, _ = synth_emb(20000, 8, 0.045, 7)
D_ref = 1 - cosine_similarity(X)
xd = torch.from_numpy(X).cuda()
D_gpu = cl.cosine_distance_device(xd).cpu().numpy()
diff = D_gpu.astype(np.float64) - D_ref.astype(np.float64)
print("affinity max|d|", np.abs(diff).max(), "mean d", diff.mean(), "mean |d|", np.abs(diff).mean())
for name, D in (("gpuAHC(ref D)", D_ref), ("gpuAHC(gpu D)", D_gpu)):
    lab, ncl = cl.ahc_average_device(torch.from_numpy(np.ascontiguousarray(D)).cuda(), 1 - 0.68)
    print(name, "clusters", int(ncl.item()))
iu = np.triu_indices(20000, 1)
t = time.time()
Z = linkage(D_gpu[iu].astype(np.float64), "average")
h = Z[:, 2]
print("scipy on gpu D: merges<thr", (h < 1 - 0.68).sum(), "closest", np.sort(np.abs(h - (1 - 0.68)))[:3], time.time() - t)
