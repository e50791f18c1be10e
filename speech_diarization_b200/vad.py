"""Drop-in for the one hot-path helper of /root/reference/vad.py: ``frame_audio`` (:9-16).
Silero VAD, hysteresis and mask morphology are out of scope."""
from __future__ import annotations

import numpy as np


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
