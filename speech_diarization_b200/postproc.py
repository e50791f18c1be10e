"""Device-side score post-processing and VAD-mask primitives (thin wrappers over the C ABI, csrc/post.cu)
behind the reference-facing functions in diar_diag.py and vad.py."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def viterbi_device(scores: torch.Tensor, alpha: float = 0.995) -> torch.Tensor:
    """viterbi_hmm (diar_diag.py:231-247) on a CUDA [T, K] f32 / f64 tensor; path int32 [T] on the device."""
    lib = _lib.load()
    T, K = scores.shape
    if T < 1:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # dp[0] = scores[0]
    eps = 1e-8
    log_move = np.float32(np.log((1 - alpha) / (K - 1) + eps))                   # ZeroDivisionError for K == 1, as the reference
    log_stay = np.float32(np.log(alpha + eps))
    if scores.dtype not in (torch.float32, torch.float64):
        scores = scores.to(torch.float64)
    scores = scores.contiguous()
    path = torch.empty((T,), dtype=torch.int32, device=scores.device)
    ws = torch.empty((lib.sd_viterbi_workspace_bytes(T, K),), dtype=torch.uint8, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(lib.sd_viterbi_hmm(scores.data_ptr(), int(scores.dtype == torch.float64), T, K, float(log_stay),
                                      float(log_move), path.data_ptr(), ws.data_ptr(), _lib.stream_ptr()),
                   "sd_viterbi_hmm")
    return path


def asnorm_device(q: torch.Tensor, r: torch.Tensor, c: torch.Tensor, topk: int = 200) -> torch.Tensor:
    """asnorm_scores (diar_diag.py:196-208) on CUDA f32 tensors q [nq, D], r [nr, D], c [nc, D] -> [nq, nr] f32."""
    lib = _lib.load()
    nq, D = q.shape
    nr, nc = r.shape[0], c.shape[0]
    out = torch.empty((nq, nr), dtype=torch.float32, device=q.device)
    if nq == 0:
        return out
    ws = torch.empty((lib.sd_asnorm_workspace_bytes(nq, nr, nc, D),), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.sd_asnorm_scores(q.data_ptr(), r.data_ptr(), c.data_ptr(), nq, nr, nc, D, int(topk),
                                        out.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), "sd_asnorm_scores")
    return out


def whiten_l2_device(x: torch.Tensor, return_sweeps: bool = False):
    """whiten_l2 (diar_diag.py:187-194) on a CUDA f32 [N, D] tensor -> CUDA f64 [N, D] (unit rows)."""
    import ctypes
    lib = _lib.load()
    N, D = x.shape
    out = torch.empty((N, D), dtype=torch.float64, device=x.device)
    ws = torch.empty((lib.sd_whiten_workspace_bytes(N, D),), dtype=torch.uint8, device=x.device)
    sweeps = ctypes.c_int32(0)
    with torch.cuda.device(x.device):
        _lib.check(lib.sd_whiten_l2_f64(x.data_ptr(), N, D, out.data_ptr(), ws.data_ptr(),
                                        ctypes.byref(sweeps) if return_sweeps else None, _lib.stream_ptr()),
                   "sd_whiten_l2_f64")
    return (out, sweeps.value) if return_sweeps else out


def cluster_centers_device(x64: torch.Tensor, labels: torch.Tensor, K: int) -> torch.Tensor:
    """Unit-norm centres of K clusters (diar_diag.py:377-383): CUDA f64 [N, D] + int32 labels -> CUDA f32 [K, D]."""
    lib = _lib.load()
    N, D = x64.shape
    out = torch.empty((K, D), dtype=torch.float32, device=x64.device)
    with torch.cuda.device(x64.device):
        _lib.check(lib.sd_cluster_centers_f64(x64.data_ptr(), labels.data_ptr(), N, D, K, None, out.data_ptr(),
                                              _lib.stream_ptr()), "sd_cluster_centers_f64")
    return out


def dot_scores_device(x: torch.Tensor, cent: torch.Tensor) -> torch.Tensor:
    """embs @ centers.T (diar_diag.py:386) on CUDA f32 tensors -> [N, K] f32."""
    lib = _lib.load()
    N, D = x.shape
    K = cent.shape[0]
    out = torch.empty((N, K), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sd_dot_scores(x.data_ptr(), cent.data_ptr(), N, K, D, out.data_ptr(), _lib.stream_ptr()),
                   "sd_dot_scores")
    return out


def hysteresis_device(probs: torch.Tensor, on: float = 0.6, off: float = 0.4) -> torch.Tensor:
    """hysteresis_binarize (vad.py:59-74) on a CUDA [n] f32 / f64 tensor -> uint8 mask [n]."""
    lib = _lib.load()
    if probs.dtype not in (torch.float32, torch.float64):
        probs = probs.to(torch.float64)
    probs = probs.contiguous()
    n = probs.shape[0]
    mask = torch.empty((n,), dtype=torch.uint8, device=probs.device)
    if n:
        with torch.cuda.device(probs.device):
            _lib.check(lib.sd_hysteresis_u8(probs.data_ptr(), int(probs.dtype == torch.float64), n, float(on),
                                            float(off), mask.data_ptr(), _lib.stream_ptr()), "sd_hysteresis_u8")
    return mask


def morph_open_close_device(mask: torch.Tensor, open_w: int, close_w: int) -> torch.Tensor:
    """binary_opening(ones(open_w)) then binary_closing(ones(close_w)) (vad.py:77-87) on a CUDA uint8 mask."""
    lib = _lib.load()
    n = mask.shape[0]
    out = torch.empty_like(mask)
    if n:
        tmp = torch.empty_like(mask)
        with torch.cuda.device(mask.device):
            _lib.check(lib.sd_morph_open_close_u8(mask.data_ptr(), n, int(open_w), int(close_w), out.data_ptr(),
                                                  tmp.data_ptr(), _lib.stream_ptr()), "sd_morph_open_close_u8")
    return out


def mask_segments_device(mask: torch.Tensor, min_speech_frames: int, min_gap_frames: int) -> np.ndarray:
    """Frame-index segments [count, 2] (start, end exclusive) of a CUDA uint8 mask (vad.py:121-151)."""
    lib = _lib.load()
    n = mask.shape[0]
    seg = torch.empty((n // 2 + 1, 2), dtype=torch.int32, device=mask.device)
    count = torch.zeros((1,), dtype=torch.int32, device=mask.device)
    ws = torch.empty((lib.sd_mask_segments_workspace_bytes(n),), dtype=torch.uint8, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.check(lib.sd_mask_segments_i32(mask.data_ptr(), n, int(min_speech_frames), int(min_gap_frames),
                                            seg.data_ptr(), count.data_ptr(), ws.data_ptr(), _lib.stream_ptr()),
                   "sd_mask_segments_i32")
    k = int(count.item())
    return seg[:k].cpu().numpy().astype(np.int64)
