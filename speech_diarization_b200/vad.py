"""Drop-in for the array helpers of /root/reference/vad.py: ``frame_audio`` (:9-16) on the hot path, and the
VAD-mask operators ``hysteresis_binarize`` (:59-74), ``morph_open_close`` (:77-87), ``mask_to_segments``
(:90-163) as GPU scans.  The Silero model itself is out of scope."""
from __future__ import annotations

import numpy as np
import torch

from . import postproc
from ._device import require_cuda


def frame_audio(y: np.ndarray, sr: int, win_ms: float = 30.0, hop_ms: float = 10.0) -> np.ndarray:
    """[n_frames, win] overlapping frames, no padding, tail dropped — librosa.util.frame(...).T.
    Like librosa it returns a strided VIEW of `y` (no copy); ``ecapa_encode_batch`` recognises
    such views and uploads the underlying samples once instead of the 2x-duplicated windows."""
    win = int(round(win_ms / 1000.0 * sr))
    hop = int(round(hop_ms / 1000.0 * sr))
    y = np.ascontiguousarray(y)
    if y.ndim != 1:
        raise ValueError("frame_audio expects a 1-D signal")
    if len(y) < win:
        raise ValueError(f"Input is too short (n={len(y)}) for frame_length={win}")   # librosa ParameterError
    n = 1 + (len(y) - win) // hop
    return np.lib.stride_tricks.as_strided(y, shape=(n, win), strides=(hop * y.itemsize, y.itemsize),
                                           writeable=False)


def _mask_to_device(mask: np.ndarray) -> torch.Tensor:
    m = np.ascontiguousarray(np.asarray(mask).astype(bool, copy=False)).view(np.uint8)
    return torch.from_numpy(m.copy()).to(require_cuda())


def hysteresis_binarize(probs: np.ndarray, on: float = 0.6, off: float = 0.4) -> np.ndarray:
    """vad.py:59-74 — a frame turns speech on at p >= on and off at p < off; bool mask of probs' shape."""
    p = np.asarray(probs)
    if p.ndim != 1:
        raise ValueError("hysteresis_binarize expects a 1-D probability track")
    if p.dtype not in (np.float32, np.float64):
        p = p.astype(np.float64)
    if p.shape[0] == 0:
        return np.zeros(p.shape, dtype=np.bool_)
    d = torch.from_numpy(np.ascontiguousarray(p)).to(require_cuda())
    return postproc.hysteresis_device(d, on, off).cpu().numpy().astype(np.bool_)


def morph_open_close(mask: np.ndarray, hop_ms: float, open_ms: float = 80.0, close_ms: float = 40.0) -> np.ndarray:
    """vad.py:77-87 — binary opening (removes speech blips shorter than open_ms) then closing (fills gaps
    shorter than close_ms), flat structures of max(1, round(ms / hop_ms)) frames."""
    mask = np.asarray(mask)
    open_w = max(1, int(round(open_ms / hop_ms))) if open_ms > 0 else 0
    close_w = max(1, int(round(close_ms / hop_ms))) if close_ms > 0 else 0
    if mask.shape[0] == 0 or (open_w == 0 and close_w == 0):
        return mask.copy()
    out = postproc.morph_open_close_device(_mask_to_device(mask), open_w, close_w)
    return out.cpu().numpy().astype(np.bool_)


def mask_to_segments(mask: np.ndarray, hop_ms: float, min_speech_ms: float = 250.0, min_gap_ms: float = 100.0,
                     speech_pad_ms: float = 80.) -> list[tuple[float, float]]:
    """vad.py:90-163 — boolean VAD mask -> [(start_s, end_s)], dropping speech shorter than min_speech_ms,
    merging gaps <= min_gap_ms, padding by speech_pad_ms and rounding to milliseconds."""
    mask = np.asarray(mask)
    total_frames = len(mask)
    if total_frames == 0:
        return []
    min_speech_frames = round(min_speech_ms / hop_ms)
    min_gap_frames = round(min_gap_ms / hop_ms)
    hop_s = hop_ms / 1000.0
    speech_pad_frames = round(speech_pad_ms / hop_ms)
    seg = postproc.mask_segments_device(_mask_to_device(mask), min_speech_frames, min_gap_frames)
    final_segments = []
    for s, e in seg:                                   # np.int64 frame indices, as np.where yields in the reference
        s_padded = max(s - speech_pad_frames, 0)
        e_padded = min(e + speech_pad_frames, total_frames)
        final_segments.append((round(s_padded * hop_s, 3), round(e_padded * hop_s, 3)))
    return final_segments
