"""Device-side affinity + clustering primitives (thin wrappers over the C ABI) shared by the
reference-facing modules diar_diag.py / anti_stick_diarize.py and by the multi-GPU driver."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._device import require_cuda, to_device_f32


def l2_normalize_device(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """x / (||x|| + eps) row-wise (anti_stick_diarize.py:176,203,430)."""
    lib = _lib.load()
    out = torch.empty_like(x)
    if x.shape[0]:
        with torch.cuda.device(x.device):
            _lib.check(lib.sd_l2norm_f32(x.data_ptr(), x.shape[0], x.shape[1], eps, out.data_ptr(),
                                         _lib.stream_ptr()), "sd_l2norm_f32")
    return out


def cosine_distance_device(emb: torch.Tensor, row0: int = 0, rows: int | None = None,
                           want_f64: bool = False):
    """Rows [row0, row0+rows) of D = 1 - cosine_similarity(emb) as a CUDA f32 [rows, N] tensor
    (diar_diag.py:219).  emb: CUDA f32 [N, D], D a multiple of 64."""
    lib = _lib.load()
    N, D = emb.shape
    rows = N - row0 if rows is None else rows
    out = torch.empty((rows, N), dtype=torch.float32, device=emb.device)
    out64 = torch.empty((rows, N), dtype=torch.float64, device=emb.device) if want_f64 else None
    ws = torch.empty((lib.sd_affinity_workspace_bytes(N, D),), dtype=torch.uint8, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(lib.sd_cosine_distance_rowblock(emb.data_ptr(), N, D, row0, rows, out.data_ptr(),
                                                   out64.data_ptr() if want_f64 else None, ws.data_ptr(),
                                                   _lib.stream_ptr()), "sd_cosine_distance_rowblock")
    return (out, out64) if want_f64 else out


def ahc_average_device(dist: torch.Tensor, threshold: float):
    """Average-linkage AHC with a distance threshold on a CUDA f32 [N, N] matrix
    (diar_diag.py:221-226).  Returns (labels int32 [N] on device, n_clusters tensor [1])."""
    lib = _lib.load()
    N = dist.shape[0]
    assert dist.shape == (N, N) and dist.dtype == torch.float32 and dist.is_contiguous()
    labels = torch.empty((N,), dtype=torch.int32, device=dist.device)
    ncl = torch.zeros((1,), dtype=torch.int32, device=dist.device)
    ws = torch.empty((lib.sd_ahc_workspace_bytes(N),), dtype=torch.uint8, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(lib.sd_ahc_average_f32(dist.data_ptr(), N, float(threshold), labels.data_ptr(),
                                          ncl.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), "sd_ahc_average_f32")
    if _KEEP_STATS:
        # debug / bench only: reading the two counters synchronises.  The 8 N^2-byte workspace itself is NOT kept
        # alive past this call (it used to be, 3.2 GB at N = 20k)
        import ctypes
        r, m = ctypes.c_int32(0), ctypes.c_int32(0)
        _lib.check(lib.sd_ahc_read_stats(ws.data_ptr(), N, ctypes.byref(r), ctypes.byref(m)), "sd_ahc_read_stats")
        ahc_average_device.last_stats = {"rounds": r.value, "merges": m.value}
    return labels, ncl


_KEEP_STATS = False


def ahc_keep_stats(enable: bool) -> None:
    """Record {"rounds", "merges"} of every following ahc_average_device call (costs a synchronisation per call)."""
    global _KEEP_STATS
    _KEEP_STATS = bool(enable)


def ahc_last_stats() -> dict:
    """{"rounds", "merges"} of the most recent ahc_average_device call made while ahc_keep_stats(True)."""
    return dict(getattr(ahc_average_device, "last_stats", {"rounds": -1, "merges": -1}))


def window_argmax_device(x: torch.Tensor, cent: torch.Tensor):
    """argmax_k <x_i, c_k> and the maximum (anti_stick_diarize.py:433-434)."""
    lib = _lib.load()
    N, D = x.shape
    K = cent.shape[0]
    best = torch.empty((N,), dtype=torch.int32, device=x.device)
    score = torch.empty((N,), dtype=torch.float32, device=x.device)
    if N:
        with torch.cuda.device(x.device):
            _lib.check(lib.sd_window_argmax(x.data_ptr(), cent.data_ptr(), N, K, D, best.data_ptr(),
                                            score.data_ptr(), _lib.stream_ptr()), "sd_window_argmax")
    return best, score


def adjacent_cosine_device(x: torch.Tensor) -> torch.Tensor:
    """cos(x_i, x_{i+1}) with the reference's +1e-8 in the denominator (anti_stick_diarize.py:102-104)."""
    lib = _lib.load()
    N, D = x.shape
    out = torch.empty((max(N - 1, 0),), dtype=torch.float32, device=x.device)
    if N > 1:
        with torch.cuda.device(x.device):
            _lib.check(lib.sd_adjacent_cosine(x.data_ptr(), N, D, out.data_ptr(), _lib.stream_ptr()),
                       "sd_adjacent_cosine")
    return out


def cluster_embeddings_device(emb: torch.Tensor, cos_thr: float = 0.68) -> torch.Tensor:
    """cluster_embeddings(method="agglo") entirely on the device; labels int32 [N]."""
    N = emb.shape[0]
    if N == 1:
        # sklearn raises for a single sample; the reference would propagate that ValueError
        raise ValueError("Found array with 1 sample(s) while a minimum of 2 is required by AgglomerativeClustering.")
    dist = cosine_distance_device(emb)
    labels, _ = ahc_average_device(dist, 1 - cos_thr)
    return labels


def to_cuda_embeddings(embs, device=None) -> torch.Tensor:
    return to_device_f32(np.asarray(embs) if not torch.is_tensor(embs) else embs, require_cuda(device))
