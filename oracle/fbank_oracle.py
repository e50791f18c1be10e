"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of ``fbank_batch`` (/root/reference/speech_encode.py:10-38).  It
executes the SAME third-party call the reference makes
(``torchaudio.transforms.MelSpectrogram``, torchaudio 2.11.0 here) with the
reference's arguments, the only change being that the hard-coded ``'cuda'``
(defect D6) becomes CPU.  Pinned by tests/golden/fbank_ref_*.npz, which
were produced by importing the reference module itself (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch
from torchaudio.transforms import MelSpectrogram


def fbank_batch(wavs: np.ndarray, sr: int = 16000, n_mels: int = 80, mean_nor: bool = True) -> np.ndarray:
    assert wavs.ndim == 2                                      # speech_encode.py:12
    win_length = int(sr * 0.025)                               # :14
    hop_length = int(sr * 0.010)                               # :15
    mel_spectrogram = MelSpectrogram(                          # :17-26
        sample_rate=sr, n_mels=n_mels, n_fft=win_length, win_length=win_length,
        hop_length=hop_length, f_min=20.0, f_max=sr / 2 - 100, power=2.0,
    )
    with torch.inference_mode():
        feat = mel_spectrogram(torch.from_numpy(wavs))         # :28-30  [B, n_mels, T]
        feat = torch.log(feat + 1e-6)                          # :32
        feat = feat.transpose(1, 2)                            # :33
        if mean_nor:
            feat = feat - feat.mean(1, keepdim=True)           # :35-36
    return feat.numpy()                                        # :38  [B, T, n_mels]
