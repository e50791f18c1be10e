"""speech_diarization_b200 — B200 (sm_100a) implementation of the diarization hot path of
hzane/speech-diarization: log-mel fbank -> ECAPA-TDNN embeddings -> cosine affinity -> AHC.

The modules mirror the reference's file names and callables so that
``from speech_encode import ecapa_encode_batch`` becomes
``from speech_diarization_b200.speech_encode import ecapa_encode_batch``.
All arithmetic runs in libsd_b200.so (hand-written CUDA behind a C ABI, include/sd_b200.h);
there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
