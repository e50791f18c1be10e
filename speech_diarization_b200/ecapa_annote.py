"""Drop-in for /root/reference/ecapa_annote.py: the pyannote embedding-model adapter.

The reference subclasses ``pyannote.audio.core.model.Model`` (ecapa_annote.py:2,6); pyannote is
not installable here, so the adapter subclasses ``torch.nn.Module`` and keeps the contract
pyannote's pipeline relies on: ``.dimension`` and ``forward(waveforms[B, n]) -> [B, 192]``.
"""
from __future__ import annotations

import torch

from .speech_encode import using_ecapa_encoder


class ECAPAEncoder(torch.nn.Module):
    def __init__(self, device: str | int = 0):              # ecapa_annote.py:7
        super().__init__()
        self.model = using_ecapa_encoder(device)            # :9
        self.dimension = 192                                # :11

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """(batch_size, num_samples) -> (batch_size, 192), on the encoder's device (:13-22)."""
        return self.model.encode_batch(waveforms).squeeze(1)
