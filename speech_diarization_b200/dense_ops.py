"""Device-side operators of the dense passes (thin wrappers over csrc/reassign.cu): change-point peaks,
speaker centroids, label scatter, run-length segments, neighbour merge, zero-padded snippet batches.
Reference semantics: /root/reference/anti_stick_diarize.py:78-127, 130-172, 333-349, 370-386, 464-475."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _st():
    return _lib.stream_ptr()


def scd_peaks_device(emb: torch.Tensor, seg_off: torch.Tensor, thr: float):
    """emb [total, D] f32 CUDA: rows seg_off[s]..seg_off[s+1] = the sliding-window embeddings of segment s.
    Returns (peak uint8 [total], z f32 [total]): peak[seg_off[s] + i] = 1 where find_peaks(z_s, height=thr) fires."""
    lib = _lib.load()
    total, D = emb.shape
    nseg = int(seg_off.numel()) - 1
    z = torch.zeros((total,), dtype=torch.float32, device=emb.device)
    peak = torch.zeros((total,), dtype=torch.uint8, device=emb.device)
    if nseg > 0 and total > 0:
        with torch.cuda.device(emb.device):
            _lib.check(lib.sd_scd_peaks(emb.data_ptr(), D, seg_off.data_ptr(), nseg, float(thr), z.data_ptr(),
                                        peak.data_ptr(), _st()), "sd_scd_peaks")
    return peak, z


def speaker_centroids_device(emb: torch.Tensor, labels: torch.Tensor, spk_ids: torch.Tensor) -> torch.Tensor:
    """Unit-norm (eps 1e-8) mean embedding of every speaker id in spk_ids; [K, D] f32 on the device."""
    lib = _lib.load()
    N, D = emb.shape
    K = int(spk_ids.numel())
    out = torch.empty((K, D), dtype=torch.float32, device=emb.device)
    if K:
        with torch.cuda.device(emb.device):
            _lib.check(lib.sd_speaker_centroids(emb.data_ptr(), labels.data_ptr(), N, D, spk_ids.data_ptr(), K,
                                                out.data_ptr(), _st()), "sd_speaker_centroids")
    return out


def scatter_labels_device(n_full: int, valid: torch.Tensor, labels: torch.Tensor, label_map: torch.Tensor | None = None):
    """full = -1 everywhere; full[valid[i]] = label_map[labels[i]] (or labels[i]).  int32 [n_full] on the device."""
    lib = _lib.load()
    full = torch.full((n_full,), -1, dtype=torch.int32, device=valid.device)
    m = int(valid.numel())
    if m:
        with torch.cuda.device(valid.device):
            _lib.check(lib.sd_scatter_labels(valid.data_ptr(), labels.data_ptr(),
                                             label_map.data_ptr() if label_map is not None else None, m,
                                             full.data_ptr(), _st()), "sd_scatter_labels")
    return full


def label_runs_device(full_labels: torch.Tensor, window_starts: torch.Tensor, sr: float, max_t: float):
    """Run-length encoding on the device.  Returns (run_idx int32 [n, 3], run_t f64 [n, 2], count int32 [1]) with
    capacity n = len(full_labels); the first count[0] rows are valid."""
    lib = _lib.load()
    n = int(full_labels.numel())
    dev = full_labels.device
    run_idx = torch.empty((max(n, 1), 3), dtype=torch.int32, device=dev)
    run_t = torch.empty((max(n, 1), 2), dtype=torch.float64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    scratch = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sd_label_runs(full_labels.data_ptr(), n, window_starts.data_ptr(), float(sr), float(max_t),
                                     scratch.data_ptr(), run_idx.data_ptr(), run_t.data_ptr(), count.data_ptr(), _st()),
                   "sd_label_runs")
    return run_idx, run_t, count


def merge_adjacent_device(seg_t: torch.Tensor, spk: torch.Tensor, spk_stride: int, n: int, gap: float,
                          n_dev: torch.Tensor | None = None):
    """Groups of segments merge_adjacent would fuse.  seg_t f64 [cap, 2]; speaker of segment k at spk.view(-1)[k * spk_stride].
    Returns (group int32 [cap, 2] = first / last segment of each group, count int32 [1])."""
    lib = _lib.load()
    dev = seg_t.device
    group = torch.empty((max(n, 1), 2), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sd_merge_adjacent(seg_t.data_ptr(), spk.data_ptr(), int(spk_stride), int(n),
                                         n_dev.data_ptr() if n_dev is not None else None, float(gap),
                                         group.data_ptr(), count.data_ptr(), _st()), "sd_merge_adjacent")
    return group, count


def gather_pad_device(audio: torch.Tensor, starts: np.ndarray, lens: np.ndarray) -> torch.Tensor:
    """[B, max(lens)] f32 CUDA batch: row b = audio[starts[b] : starts[b] + lens[b]] followed by zeros."""
    lib = _lib.load()
    B = int(len(starts))
    max_len = int(lens.max()) if B else 0
    out = torch.empty((B, max_len), dtype=torch.float32, device=audio.device)
    if B and max_len:
        if int((starts + lens).max()) > audio.numel() or int(starts.min()) < 0:
            raise ValueError("snippet runs past the audio buffer")
        st = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int64)).to(audio.device)
        ln = torch.from_numpy(np.ascontiguousarray(lens, dtype=np.int32)).to(audio.device)
        with torch.cuda.device(audio.device):
            _lib.check(lib.sd_gather_pad_f32(audio.data_ptr(), st.data_ptr(), ln.data_ptr(), B, max_len, out.data_ptr(),
                                             _st()), "sd_gather_pad_f32")
    return out
