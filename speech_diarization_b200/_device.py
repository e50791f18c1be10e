"""Small torch-side helpers shared by the reference-facing modules: device memory,
streams and error checking only (PyTorch is plumbing here, not the arithmetic)."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.SdError("speech_diarization_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None or device == "cuda":
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        return torch.device("cuda", device)
    d = torch.device(device)
    if d.type != "cuda":
        raise _lib.SdError(f"device {device!r} is not a CUDA device; there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def to_device_f32(x, device: torch.device) -> torch.Tensor:
    """numpy / torch, any float dtype -> contiguous f32 CUDA tensor (H2D copy if needed)."""
    if isinstance(x, np.ndarray):
        with warnings.catch_warnings():      # read-only views (frame_audio) are only ever read
            warnings.simplefilter("ignore", UserWarning)
            x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
