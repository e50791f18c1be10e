"""Drop-in for the hot-path pieces of /root/reference/diar_diag.py:
``cluster_embeddings`` (:213-229, "agglo" branch) and the ``frame_audio`` duplicate (:48-56).
Whitening / AS-norm / Viterbi / plotting are "next" rows and not built."""
from __future__ import annotations

import numpy as np

from . import _lib
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
