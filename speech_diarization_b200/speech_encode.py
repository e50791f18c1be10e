"""Drop-in for /root/reference/speech_encode.py (the ECAPA / fbank entry points).

    fbank_batch(wavs, sr=16000, n_mels=80, mean_nor=True) -> np.ndarray [B, T, 80]   (:10-38)
    using_ecapa_encoder(device="cuda") -> encoder with .encode_batch(Tensor[B,n]) -> [B,1,192]  (:64-70)
    ecapa_encode_batch(wavs) -> np.ndarray [B, 192]                                  (:73-78)

Differences from the reference that are forced by this environment, not by design:
the reference downloads `LanceaKing/spkrec-ecapa-cnceleb` from the HF hub
(speech_encode.py:66-69); here the speechbrain-keyed state dict must be registered with
``register_ecapa_state_dict`` or pointed to by ``$SD_ECAPA_CKPT`` (a torch-saved
``embedding_model.ckpt``).  The ONNX ERes2NetV2 encoder (:42-60) is out of scope
(SURVEY.md §8f rank 3).
"""
from __future__ import annotations

import ctypes
import os
from functools import lru_cache

import numpy as np
import torch

from . import _lib
from ._device import require_cuda, to_device_f32

EMB_DIM = 192
N_MELS = 80
_HOP = 160

_registered_state_dict: dict | None = None


# ----------------------------------------------------------------------------- fbank
def fbank_batch_device(wavs: torch.Tensor, variant: int = 0, mean_nor: bool = True,
                       wav_stride: int | None = None, n_windows: int | None = None,
                       n_samples: int | None = None) -> torch.Tensor:
    """Device-resident fbank: wavs is a CUDA f32 tensor, either [B, n] contiguous or a 1-D audio
    buffer addressed as n_windows windows of n_samples every wav_stride samples."""
    lib = _lib.load()
    if wavs.dim() == 2:
        B, n = wavs.shape
        stride = wavs.stride(0)
    else:
        B, n, stride = int(n_windows), int(n_samples), int(wav_stride)
        if (B - 1) * stride + n > wavs.numel():
            raise ValueError("windows run past the end of the audio buffer")
    T = lib.sd_fbank_num_frames(n)
    out = torch.empty((B, T, N_MELS), dtype=torch.float32, device=wavs.device)
    if B == 0:
        return out
    with torch.cuda.device(wavs.device):
        _lib.check(lib.sd_fbank_f32(wavs.data_ptr(), stride, B, n, variant, int(bool(mean_nor)),
                                    out.data_ptr(), _lib.stream_ptr()), "sd_fbank_f32")
    return out


def fbank_batch(wavs: np.ndarray, sr: int = 16000, n_mels: int = 80, mean_nor: bool = True) -> np.ndarray:
    """speech_encode.py:10-38.  [B, n_samples] -> [B, T, n_mels] log-mel (+ CMN)."""
    assert wavs.ndim == 2                                  # :12
    if sr != 16000 or n_mels != 80:
        raise _lib.SdError("fbank_batch: the CUDA kernel is built for sr=16000, n_mels=80 "
                           "(the only configuration the reference calls it with, speech_encode.py:57)")
    dev = require_cuda()
    x = to_device_f32(wavs, dev)
    return fbank_batch_device(x, variant=0, mean_nor=mean_nor).cpu().numpy()   # :38


# --------------------------------------------------------------------- ECAPA encoder
class EcapaEncoderB200:
    """Stands in for speechbrain's EncoderClassifier on this path: ``encode_batch`` is the only
    method the reference calls (speech_encode.py:77, ecapa_annote.py:22, diar_diag.py:169)."""

    def __init__(self, state_dict: dict, device="cuda", max_batch: int = 512, max_samples: int = 24000):
        self.device = require_cuda(device)
        self._lib = _lib.load()
        self._sd = {k: v.detach().to("cpu", torch.float32).contiguous()
                    for k, v in state_dict.items() if torch.is_tensor(v) and v.dtype.is_floating_point}
        self._plan = ctypes.c_void_p()
        self._cap = (0, 0)
        self._make_plan(max_batch, max_samples)

    # -- plan management
    def _make_plan(self, max_batch: int, max_samples: int) -> None:
        self.close()
        names = list(self._sd.keys())
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*[s.encode() for s in names])
        c_ptrs = (ctypes.c_void_p * n)(*[self._sd[s].data_ptr() for s in names])
        c_numel = (ctypes.c_int64 * n)(*[self._sd[s].numel() for s in names])
        plan = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sd_ecapa_plan_create(c_names, c_ptrs, c_numel, n, max_batch, max_samples,
                                                      ctypes.byref(plan)), "sd_ecapa_plan_create")
        self._plan = plan
        self._cap = (max_batch, max_samples)
        self._max_rows = max_batch * self._tp(1 + max_samples // _HOP)

    @staticmethod
    def _tp(T: int) -> int:
        return ((T + 8 + 15) // 16) * 16

    def close(self) -> None:
        if getattr(self, "_plan", None) is not None and self._plan.value:
            self._lib.sd_ecapa_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the hot call
    def embed_device(self, audio: torch.Tensor, wav_stride: int, n_windows: int, n_samples: int,
                     l2_normalize: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
        """Embeddings of windows addressed in place in a CUDA f32 buffer: window b =
        audio[b*wav_stride : b*wav_stride + n_samples].  Returns [n_windows, 192] f32 on the device.
        No host synchronisation."""
        B, n = int(n_windows), int(n_samples)
        if out is None:
            out = torch.empty((B, EMB_DIM), dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        if n < 400:
            raise ValueError(f"windows of {n} samples are shorter than one 25 ms analysis frame")
        if not (torch.is_tensor(audio) and audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()):
            raise ValueError("embed_device needs a contiguous f32 CUDA tensor")
        if audio.device != self.device:
            raise ValueError(f"audio lives on {audio.device}, the encoder on {self.device}")
        if wav_stride < 1 or (B - 1) * int(wav_stride) + n > audio.numel():
            raise ValueError(f"{B} windows of {n} samples every {wav_stride} run past the {audio.numel()}-sample buffer")
        tp = self._tp(1 + n // _HOP)
        if tp > self._max_rows:                      # one window larger than the whole workspace: grow
            self._make_plan(1, n)
        per_call = max(1, self._max_rows // tp)
        flat = audio.reshape(-1)
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for b0 in range(0, B, per_call):
                nb = min(per_call, B - b0)
                _lib.check(self._lib.sd_ecapa_embed(self._plan, flat.data_ptr() + 4 * b0 * wav_stride,
                                                    wav_stride, nb, n, int(bool(l2_normalize)),
                                                    out.data_ptr() + 4 * EMB_DIM * b0, st), "sd_ecapa_embed")
        return out

    def embed_offsets_device(self, audio: torch.Tensor, offsets, n_samples: int, l2_normalize: bool = False,
                             out: torch.Tensor | None = None) -> torch.Tensor:
        """Embeddings of windows that start at ARBITRARY sample offsets of one CUDA f32 recording: window b =
        audio[offsets[b] : offsets[b] + n_samples].  offsets: int64 array / tensor.  One launch sequence for all
        of them (the speech windows of a recording, the sliding windows of all SCD segments).  No host sync."""
        off = torch.as_tensor(offsets, dtype=torch.int64)
        B, n = int(off.numel()), int(n_samples)
        if out is None:
            out = torch.empty((B, EMB_DIM), dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        if n < 400:
            raise ValueError(f"windows of {n} samples are shorter than one 25 ms analysis frame")
        if not (torch.is_tensor(audio) and audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
                and audio.device == self.device):
            raise ValueError("embed_offsets_device needs a contiguous f32 CUDA tensor on the encoder's device")
        if not off.is_cuda:
            if int(off.min()) < 0 or int(off.max()) + n > audio.numel():
                raise ValueError("a window runs past the audio buffer")
            off = off.to(self.device)
        off = off.contiguous()
        tp = self._tp(1 + n // _HOP)
        if tp > self._max_rows:
            self._make_plan(1, n)
        per_call = max(1, self._max_rows // tp)
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for b0 in range(0, B, per_call):
                nb = min(per_call, B - b0)
                _lib.check(self._lib.sd_ecapa_embed_offsets(self._plan, audio.data_ptr(), off.data_ptr() + 8 * b0, nb, n,
                                                            int(bool(l2_normalize)), out.data_ptr() + 4 * EMB_DIM * b0, st),
                           "sd_ecapa_embed_offsets")
        return out

    def embed_host(self, wavs: np.ndarray, wav_stride: int, n_windows: int, n_samples: int,
                   l2_normalize: bool = False) -> np.ndarray:
        """Embeddings of windows addressed in place in a HOST f32 buffer (numpy, contiguous span): the upload is
        chunked and overlapped with the fbank kernels inside the C call (sd_ecapa_embed_host).  Returns
        [n_windows, 192] f32 numpy; synchronous like the reference's `.cpu().numpy()`."""
        B, n = int(n_windows), int(n_samples)
        out = np.empty((B, EMB_DIM), dtype=np.float32)
        if B == 0:
            return out
        if n < 400:
            raise ValueError(f"windows of {n} samples are shorter than one 25 ms analysis frame")
        if not (isinstance(wavs, np.ndarray) and wavs.dtype == np.float32 and wavs.flags.c_contiguous):
            raise ValueError("embed_host needs a C-contiguous float32 numpy array")
        if wav_stride < 1 or (B - 1) * int(wav_stride) + n > wavs.size:
            raise ValueError(f"{B} windows of {n} samples every {wav_stride} run past the {wavs.size}-sample buffer")
        tp = self._tp(1 + n // _HOP)
        if tp > self._max_rows:
            self._make_plan(1, n)
        per_call = max(1, self._max_rows // tp)
        base = wavs.ctypes.data
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for b0 in range(0, B, per_call):
                nb = min(per_call, B - b0)
                _lib.check(self._lib.sd_ecapa_embed_host(self._plan, base + 4 * b0 * wav_stride, wav_stride, nb, n,
                                                         int(bool(l2_normalize)), out.ctypes.data + 4 * EMB_DIM * b0, st),
                           "sd_ecapa_embed_host")
        return out

    def encode_batch(self, wavs: torch.Tensor, wav_lens=None, normalize: bool = False) -> torch.Tensor:
        """EncoderClassifier.encode_batch: [B, n] (or [n]) waveform tensor -> [B, 1, 192] on the
        encoder's device.  wav_lens is accepted for signature compatibility; the reference never
        passes it, and relative lengths other than 1.0 are not implemented."""
        if wav_lens is not None and not bool(torch.all(torch.as_tensor(wav_lens) == 1.0)):
            raise _lib.SdError("encode_batch: wav_lens != 1 is not supported (the reference never passes wav_lens)")
        if normalize:
            raise _lib.SdError("encode_batch(normalize=True) (speechbrain's embedding mean/var norm) is not used "
                               "by the reference and not implemented")
        if wavs.dim() == 1:
            wavs = wavs.unsqueeze(0)
        x = to_device_f32(wavs, self.device)
        emb = self.embed_device(x, x.stride(0), x.shape[0], x.shape[1])
        return emb.unsqueeze(1)

    def overflowed(self, reset: bool = True) -> bool:
        """True when a forward since the last reset saturated an activation at the f16 range (its embeddings were
        delivered as NaN).  Synchronises the current stream."""
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sd_ecapa_overflow(self._plan, int(bool(reset)), ctypes.byref(flag), _lib.stream_ptr()),
                       "sd_ecapa_overflow")
        return bool(flag.value)

    def forward_feats(self, feats: torch.Tensor, l2_normalize: bool = False) -> torch.Tensor:
        """Trunk only, from [B, T, 80] features (parity tests)."""
        x = to_device_f32(feats, self.device)
        B, T, _ = x.shape
        out = torch.empty((B, EMB_DIM), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sd_ecapa_forward_feats(self._plan, x.data_ptr(), B, T, int(bool(l2_normalize)),
                                                        out.data_ptr(), _lib.stream_ptr()), "sd_ecapa_forward_feats")
        return out

    def profile(self, enable: bool) -> None:
        """Record CUDA events at every stage boundary of subsequent forwards (and reset the record)."""
        _lib.check(self._lib.sd_ecapa_profile(self._plan, int(bool(enable))), "sd_ecapa_profile")

    def profile_read(self) -> tuple[dict, int]:
        """({stage: total milliseconds over the recorded forwards}, number of forwards)."""
        ns = self._lib.sd_ecapa_num_stages()
        names = ctypes.create_string_buffer(32 * ns)
        ms = (ctypes.c_float * ns)()
        nf = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sd_ecapa_profile_read(self._plan, ns, names, ms, ctypes.byref(nf)),
                       "sd_ecapa_profile_read")
        out = {names.raw[32 * i:32 * (i + 1)].split(b"\0")[0].decode(): float(ms[i]) for i in range(ns)}
        return out, nf.value

    def debug_fetch(self, name: str, B: int, T: int) -> torch.Tensor:
        big = torch.empty((B * T * 3072,), dtype=torch.float32, device=self.device)
        C = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.sd_ecapa_debug_fetch(self._plan, name.encode(), big.data_ptr(), ctypes.byref(C),
                                                      _lib.stream_ptr()), "sd_ecapa_debug_fetch")
        c = C.value
        if name in ("b3.se", "asp.stats", "asp.uttbias", "pooled"):
            return big[: B * c].reshape(B, c).clone()
        return big[: B * T * c].reshape(B, T, c).clone()


def register_ecapa_state_dict(state_dict: dict | None) -> None:
    """Provide the speechbrain-keyed ECAPA-TDNN weights ``using_ecapa_encoder`` will load."""
    global _registered_state_dict
    _registered_state_dict = state_dict
    using_ecapa_encoder.cache_clear()


@lru_cache(maxsize=1)                                       # speech_encode.py:64 (process-wide singleton)
def using_ecapa_encoder(device: str | int = "cuda") -> EcapaEncoderB200:
    sd = _registered_state_dict
    if sd is None:
        path = os.environ.get("SD_ECAPA_CKPT")
        if not path:
            raise _lib.SdError(
                "no ECAPA-TDNN weights: call register_ecapa_state_dict(state_dict) or set $SD_ECAPA_CKPT to a "
                "speechbrain embedding_model.ckpt (the reference fetches LanceaKing/spkrec-ecapa-cnceleb from "
                "the HF hub, speech_encode.py:66-69; there is no network here)")
        sd = torch.load(path, map_location="cpu", weights_only=True)
    return EcapaEncoderB200(sd, device=device)


def _overlap_span(wavs: np.ndarray):
    """If `wavs` is a strided view of overlapping windows over one buffer (what
    vad.frame_audio returns), give back (1-D span, hop) so only the span is uploaded."""
    if wavs.ndim != 2 or wavs.dtype != np.float32 or wavs.shape[0] < 2:
        return None
    s0, s1 = wavs.strides
    n = wavs.shape[1]
    if s1 != 4 or s0 % 4 or not (0 < s0 < n * 4):
        return None
    hop = s0 // 4
    span = np.lib.stride_tricks.as_strided(wavs, shape=((wavs.shape[0] - 1) * hop + n,), strides=(4,),
                                           writeable=False)
    return span, hop


def ecapa_encode_batch(wavs: np.ndarray) -> np.ndarray:
    """speech_encode.py:73-78.  [B, n] -> [B, 192] f32 (not L2-normalised)."""
    encoder = using_ecapa_encoder()
    with torch.inference_mode():
        ov = _overlap_span(wavs) if isinstance(wavs, np.ndarray) else None
        host_path = os.environ.get("SD_ECAPA_HOST_PATH", "1") != "0"
        if ov is not None and host_path:
            span, hop = ov                                               # windows addressed in place in host memory
            y = encoder.embed_host(span, hop, wavs.shape[0], wavs.shape[1])
        elif (host_path and isinstance(wavs, np.ndarray) and wavs.ndim == 2 and wavs.dtype == np.float32
              and wavs.flags.c_contiguous and wavs.shape[0] > 0):
            y = encoder.embed_host(wavs, wavs.shape[1], wavs.shape[0], wavs.shape[1])
        elif ov is not None:
            span, hop = ov
            audio = to_device_f32(span, encoder.device)
            y = encoder.embed_device(audio, hop, wavs.shape[0], wavs.shape[1]).cpu().numpy()
        else:
            x = torch.from_numpy(np.ascontiguousarray(wavs)).float()     # :76
            y = encoder.encode_batch(x).squeeze(1).cpu().numpy()        # :77
    return y  # [B, 192]
