"""Drop-in for the hot-path pieces of /root/reference/diar_diag.py:
``cluster_embeddings`` (:213-229, "agglo" branch), the ``frame_audio`` duplicate (:48-56), and the score
post-processing of its pipeline: ``whiten_l2`` (:187-194), ``asnorm_scores`` (:196-208) and ``viterbi_hmm``
(:231-247).  Audio loading, plotting and the JSON / SRT / CSV writers are not built."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, postproc
from ._device import require_cuda
from .clustering import cluster_embeddings_device, to_cuda_embeddings


def frame_audio(y: np.ndarray, sr: int, win_ms: float = 30.0, hop_ms: float = 10.0):
    """diar_diag.py:48-56 — pads a too-short signal to one window; returns (frames, hop)."""
    win = int(round(win_ms / 1000.0 * sr))
    hop = int(round(hop_ms / 1000.0 * sr))
    if len(y) < win:
        y = np.pad(y, (0, win - len(y)))
    n = 1 + (len(y) - win) // hop
    idx = np.arange(win)[None, :] + hop * np.arange(n)[:, None]
    return y[idx], hop


def cluster_embeddings(embs: np.ndarray, method="hdbscan", cos_thr: float = 0.68):
    """diar_diag.py:213-229.  method="agglo": average-linkage AHC on 1 - cosine similarity, cut at
    1 - cos_thr, on the GPU (tensor-core affinity + RNN-parallel Lance-Williams).  Labels equal
    sklearn's up to a permutation (numbered by first appearance)."""
    if method == "agglo":
        x = to_cuda_embeddings(embs)
        if x.shape[1] % 64:
            raise _lib.SdError(f"embedding dimension {x.shape[1]} must be a multiple of 64 (ECAPA: 192)")
        return cluster_embeddings_device(x, cos_thr).cpu().numpy().astype(np.int64)
    if method == "hdbscan":
        raise NotImplementedError("method='hdbscan' is outside the B200 hot path (SURVEY.md §2 #9-10); use 'agglo'")
    raise ValueError("method 必须是 hdbscan 或 agglo")      # diar_diag.py:228


def whiten_l2(embs: np.ndarray) -> np.ndarray:
    """diar_diag.py:187-194 — centre, ZCA-whiten with (cov + 1e-6 I)^(-1/2), L2-normalise; float64 [N, D].
    Covariance, eigendecomposition (Jacobi) and the projection all run on the GPU in f64."""
    x = to_cuda_embeddings(embs)
    return postproc.whiten_l2_device(x).cpu().numpy()


def asnorm_scores(query_embs: np.ndarray, ref_centers: np.ndarray, cohort_embs: np.ndarray,
                  topk: int = 200) -> np.ndarray:
    """diar_diag.py:196-208 — adaptive symmetric normalisation of query x centre cosine scores against the
    top-k cohort similarities of each side.  Computed on the GPU in f32 (cohort similarities on the tensor
    cores, statistics in f64); the result has the inputs' result dtype."""
    res_dtype = np.result_type(np.asarray(query_embs).dtype, np.asarray(ref_centers).dtype,
                               np.asarray(cohort_embs).dtype, np.float32)
    q, r, c = (to_cuda_embeddings(x) for x in (query_embs, ref_centers, cohort_embs))
    if q.shape[1] % 64:
        raise _lib.SdError(f"embedding dimension {q.shape[1]} must be a multiple of 64 (ECAPA: 192)")
    return postproc.asnorm_device(q, r, c, topk).cpu().numpy().astype(res_dtype, copy=False)


def viterbi_hmm(scores: np.ndarray, alpha: float = 0.995) -> np.ndarray:
    """diar_diag.py:231-247 — most likely state path of a sticky HMM (stay probability alpha) over per-step
    scores [T, K]; int32 [T].  The float32 recursion of the reference is reproduced exactly on the GPU."""
    s = np.asarray(scores)
    T, K = s.shape
    if s.dtype == np.float16:
        s = s.astype(np.float32)
    elif s.dtype not in (np.float32, np.float64):
        s = s.astype(np.float64)
    d = torch.from_numpy(np.ascontiguousarray(s)).to(require_cuda())
    return postproc.viterbi_device(d, alpha).cpu().numpy().astype(np.int32)
