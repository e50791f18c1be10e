"""Drop-in for the hot-path pieces of /root/reference/anti_stick_diarize.py: the callers and data formats either
side of the embedding + affinity kernels.

    embed_segments(y, sr, segs, ...)            :130-172   variable-length segments -> [N, 192]
    scd_split_segments(y, sr, segments, ...)    :78-127    sliding-window change detection inside segments
    speaker_centroids(segs, embs)               :333-349   unit-norm mean embedding per speaker
    frame_reassign(y, sr, speech_mask, segs, embs, ...)  :390-460  dense re-labelling pass
    _get_speech_windows / _labels_to_segments / merge_adjacent   :352-386, :464-475

Same names, arguments, return types and numbers as the reference; the organisation is this repo's own: the
recording is uploaded ONCE and windows are addressed in place on the device (by offset lists across all segments /
all speech windows, not per-segment or per-128 Python batches), embeddings never return to the host between
stages, and scoring, arg-max, run-length encoding and neighbour merging are device kernels (csrc/reassign.cu,
csrc/affinity.cu).  VAD, HDBSCAN, conservative_merge and the CLI are out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import dense_ops
from .clustering import (adjacent_cosine_device, cosine_distance_device, l2_normalize_device,
                         to_cuda_embeddings, window_argmax_device)
from .speech_encode import using_ecapa_encoder
from ._device import to_device_f32

def _current_device() -> torch.device:
    """The device the process-wide encoder singleton is created on (speech_encode.using_ecapa_encoder("cuda"))."""
    from ._device import require_cuda
    return require_cuda()


_NO_SPK = -(2 ** 31)          # device stand-in for Segment.spk is None (merge_adjacent compares spk with ==)


@dataclass
class Segment:                                   # anti_stick_diarize.py:21-26
    start: float
    end: float
    spk: int | None = None
    score: float | None = None


# --------------------------------------------------------------------------------- embeddings
def _embed_rows_device(audio: torch.Tensor, offsets: np.ndarray, n_samples: int, l2_normalize: bool = False) -> torch.Tensor:
    """Embeddings of equal-length windows at the given sample offsets of the device-resident recording (the one
    place the dense passes reach the encoder; tests substitute it to inject known embeddings)."""
    return using_ecapa_encoder().embed_offsets_device(audio, offsets, n_samples, l2_normalize=l2_normalize)


def _encode_batch_device(batch: torch.Tensor) -> torch.Tensor:
    """[B, n] zero-padded CUDA batch -> [B, 192] CUDA embeddings, not L2-normalised (what ecapa_encode_batch computes,
    speech_encode.py:73-78, without the host round trip).  Tests substitute it to record the batches."""
    return using_ecapa_encoder().encode_batch(batch).squeeze(1)


def embed_segments(y: np.ndarray, sr: int, segs: list, batch_size: int = 32,
                   min_duration_ms: float = 500.0, pad_duration_ms: float = 150.0) -> np.ndarray:
    """anti_stick_diarize.py:130-172.  One embedding per segment; segments shorter than `min_duration_ms` are
    widened by `pad_duration_ms` on both sides (:157-160); each group of `batch_size` consecutive segments is
    zero-padded to its longest member (the padding is signal: the reference passes no wav_lens).
    The sample ranges are computed in one vectorised step, the recording is uploaded once and the padded batches
    are assembled on the device."""
    if len(segs) == 0:
        return np.empty((0, 192), dtype=np.float32)          # :143-144
    n = len(y)
    lo = np.array([int(s.start * sr) for s in segs], dtype=np.int64)      # :153 (Python slice bounds)
    hi = np.array([int(s.end * sr) for s in segs], dtype=np.int64)
    lo_c, hi_c = np.clip(lo, 0, n), np.clip(hi, 0, n)                     # y[s:e] clips like this for s, e >= 0
    short = np.maximum(hi_c - lo_c, 0) < int(min_duration_ms / 1000.0 * sr)
    pad = int(pad_duration_ms / 1000.0 * sr)
    lo_w = np.where(short, np.minimum(np.maximum(0, lo - pad), n), lo_c)
    hi_w = np.where(short, np.minimum(n, hi + pad), hi_c)
    lens = np.maximum(hi_w - lo_w, 0).astype(np.int32)
    dev = _current_device()
    audio = to_device_f32(y, dev)
    out = torch.empty((len(segs), 192), dtype=torch.float32, device=dev)
    for i in range(0, len(segs), batch_size):
        sl = slice(i, min(i + batch_size, len(segs)))
        out[sl] = _encode_batch_device(dense_ops.gather_pad_device(audio, lo_w[sl], lens[sl]))   # :162-168
    return out.cpu().numpy()


# ------------------------------------------------------------------------------ scoring heads
def cosine_distance(embs: np.ndarray) -> np.ndarray:
    """Head of cluster_hdbscan (anti_stick_diarize.py:176-177): L2-normalise, D = 1 - cos."""
    x = to_cuda_embeddings(embs)
    return cosine_distance_device(l2_normalize_device(x)).cpu().numpy()


def adjacent_cosine(embs: np.ndarray) -> np.ndarray:
    """anti_stick_diarize.py:102-104."""
    return adjacent_cosine_device(to_cuda_embeddings(embs)).cpu().numpy()


# ------------------------------------------------------------------------------------- SCD
def scd_split_segments(y: np.ndarray, sr: int, segments: list, win_ms: float = 1000.0, hop_ms: float = 200.0,
                       thr: float = 1.25, min_speech_ms: float = 1000.0) -> list:
    """anti_stick_diarize.py:78-127.  Inside every segment: sliding windows (frame_audio semantics), embeddings,
    adjacent cosine distance, z-score, peaks above `thr`, cuts at the window mid-points that leave at least
    `min_speech_ms` on both sides.  The reference calls the encoder once per segment; here the windows of ALL
    segments are embedded in one pass over the device-resident recording and one kernel finds every segment's
    peaks (sd_scd_peaks).  A segment shorter than one window is passed through (librosa would raise)."""
    assert y.ndim == 1 and y.dtype == np.float32
    min_speech_s = min_speech_ms / 1000.0
    win = int(round(win_ms / 1000.0 * sr))
    hop = int(round(hop_ms / 1000.0 * sr))
    n_audio = len(y)
    # windows per segment: frame_audio(y[int(start*sr):int(end*sr)]) -> 1 + (len - win) // hop
    a0 = np.clip(np.array([int(s.start * sr) for s in segments], dtype=np.int64), 0, n_audio)
    a1 = np.clip(np.array([int(s.end * sr) for s in segments], dtype=np.int64), 0, n_audio)
    seg_len = np.maximum(a1 - a0, 0)
    n_win = np.where(seg_len >= win, 1 + (seg_len - win) // max(hop, 1), 0)
    active = np.flatnonzero(n_win >= 3)                                    # :98-100
    peaks_of: dict[int, np.ndarray] = {}
    if active.size:
        counts = n_win[active]
        seg_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        offsets = np.concatenate([a0[s] + hop * np.arange(c, dtype=np.int64) for s, c in zip(active, counts)])
        enc_dev = _current_device()
        audio = to_device_f32(y, enc_dev)
        emb = _embed_rows_device(audio, offsets, win)
        peak, _ = dense_ops.scd_peaks_device(emb, torch.from_numpy(seg_off).to(enc_dev), thr)
        peak = peak.cpu().numpy()
        for k, s in enumerate(active):
            peaks_of[int(s)] = np.flatnonzero(peak[seg_off[k]:seg_off[k + 1]])
    out: list = []
    for i, seg in enumerate(segments):
        pk = peaks_of.get(i)
        if pk is None or pk.size == 0:                                      # :99-100, :112-114
            out.append(seg)
            continue
        cuts = np.unique(seg.start + (pk + 0.5) * hop_ms / 1000.0)          # :116-117 sorted(set(...))
        last_cut = seg.start
        for cut_time in cuts:
            if cut_time - last_cut >= min_speech_s:
                out.append(Segment(last_cut, cut_time))
                last_cut = cut_time
        if seg.end - last_cut >= min_speech_s:
            out.append(Segment(last_cut, seg.end))
    return out


# ------------------------------------------------------------------------------ centroids
def speaker_centroids(segs: list, embs: np.ndarray):
    """anti_stick_diarize.py:333-349: (speaker ids, unit-norm centroid per speaker).  Returns the ids as an int
    array in ascending order — the reference's `np.array(centroids.keys())` is a 0-d object array (SURVEY defect
    D4) that frame_reassign then indexes by position; the positions are the same."""
    ids = sorted({s.spk for s in segs if s.spk is not None and s.spk >= 0})
    if not ids:
        return np.empty(0, dtype=int), np.empty((0, 192), dtype=np.float32)
    labels = np.array([_NO_SPK if s.spk is None else s.spk for s in segs], dtype=np.int32)
    x = to_cuda_embeddings(embs)
    cent = dense_ops.speaker_centroids_device(x, torch.from_numpy(labels).to(x.device),
                                              torch.tensor(ids, dtype=torch.int32, device=x.device))
    return np.array(ids, dtype=int), cent.cpu().numpy()


# ------------------------------------------------------------------------- windows and segments
def _get_speech_windows(y: np.ndarray, sr: int, speech_mask: list, win_samples: int, step_samples: int):
    """anti_stick_diarize.py:352-367: start samples of all sliding windows and the indices of those whose centre
    falls on a 10 ms frame covered by `speech_mask`.  Coverage is built from interval end points (a difference
    array), one pass instead of one slice assignment per mask segment."""
    hop_s = 0.01
    n_frames = math.ceil(len(y) / sr / hop_s)
    edges = np.zeros(n_frames + 1, dtype=np.int32)
    for sm in speech_mask:
        a, b = int(sm.start / hop_s), int(sm.end / hop_s)              # smask[a:b] = True with Python slice clipping
        a, b = min(max(a, 0), n_frames), min(max(b, 0), n_frames)
        if b > a:
            edges[a] += 1
            edges[b] -= 1
    covered = np.cumsum(edges[:-1]) > 0
    window_starts = np.arange(0, len(y) - win_samples, step_samples)
    centre_frames = np.clip(((window_starts + win_samples / 2) / sr / hop_s).astype(int), 0, n_frames - 1)
    return window_starts, np.flatnonzero(covered[centre_frames])


def _segments_from_runs(run_idx: np.ndarray, run_t: np.ndarray) -> list:
    return [Segment(float(t0), float(t1), int(k)) for (_, _, k), (t0, t1) in zip(run_idx, run_t)]


def _labels_to_segments(window_starts: np.ndarray, valid_indices: np.ndarray, window_labels: np.ndarray,
                        sr: int, max_t: float) -> list:
    """anti_stick_diarize.py:370-386: labels of the valid windows -> Segment list (run-length encoding on the
    device, sd_label_runs; times are the reference's own float64 expressions)."""
    n = len(window_starts)
    if n == 0:
        return []
    dev = _current_device()
    full = dense_ops.scatter_labels_device(
        n, torch.from_numpy(np.ascontiguousarray(valid_indices, dtype=np.int32)).to(dev),
        torch.from_numpy(np.ascontiguousarray(window_labels, dtype=np.int32)).to(dev))
    ws = torch.from_numpy(np.ascontiguousarray(window_starts, dtype=np.int64)).to(dev)
    run_idx, run_t, count = dense_ops.label_runs_device(full, ws, sr, max_t)
    c = int(count.item())
    return _segments_from_runs(run_idx[:c].cpu().numpy(), run_t[:c].cpu().numpy())


def merge_adjacent(segments: list, gap: float = 0.05) -> list:
    """anti_stick_diarize.py:464-475: consecutive segments of the same speaker no further than `gap` apart fuse
    (sd_merge_adjacent finds the groups).  As in the reference a segment that fuses with nothing is returned as
    the original object, a fused group as a new Segment(first.start, last.end, spk)."""
    n = len(segments)
    if n == 0:
        return []
    dev = _current_device()
    seg_t = torch.tensor([[s.start, s.end] for s in segments], dtype=torch.float64, device=dev)
    spk = torch.tensor([_NO_SPK if s.spk is None else int(s.spk) for s in segments], dtype=torch.int32, device=dev)
    group, count = dense_ops.merge_adjacent_device(seg_t, spk, 1, n, gap)
    g = group[:int(count.item())].cpu().numpy()
    return [segments[a] if a == b else Segment(segments[a].start, segments[b].end, segments[a].spk) for a, b in g]


# ---------------------------------------------------------------------------- dense reassignment
def reassign_windows(y: np.ndarray, sr: int, speech_mask: list, spk_ids, c_matrix: np.ndarray,
                     smooth_step: float = 0.1, win: float = 1.0) -> list:
    """The dense pass of frame_reassign (anti_stick_diarize.py:411-460) given unit-norm speaker centroids: embed
    every `win`-second window whose centre is speech (step `smooth_step`), score against the centroids, arg-max,
    run-length encode and merge.  Everything between the audio upload and the final (small) segment table stays
    on the device: offsets -> fbank/ECAPA -> L2 norm -> window x centroid arg-max -> label scatter -> run-length
    -> neighbour merge."""
    win_samples = int(win * sr)
    step_samples = int(smooth_step * sr)
    window_starts, valid_indices = _get_speech_windows(y, sr, speech_mask, win_samples, step_samples)
    if valid_indices.size == 0 or len(c_matrix) == 0:
        return []
    dev = _current_device()
    audio = to_device_f32(y, dev)
    embs = _embed_rows_device(audio, window_starts[valid_indices].astype(np.int64), win_samples, l2_normalize=True)
    best, _ = window_argmax_device(embs, to_device_f32(np.asarray(c_matrix), dev))
    ids = torch.from_numpy(np.ascontiguousarray(np.asarray(spk_ids), dtype=np.int32)).to(dev)
    n = len(window_starts)
    full = dense_ops.scatter_labels_device(n, torch.from_numpy(valid_indices.astype(np.int32)).to(dev), best, ids)
    ws = torch.from_numpy(window_starts.astype(np.int64)).to(dev)
    run_idx, run_t, count = dense_ops.label_runs_device(full, ws, sr, len(y) / sr)
    group, gcount = dense_ops.merge_adjacent_device(run_t, run_idx.view(-1)[2:], 3, n, 0.05, n_dev=count)
    c, gc = int(count.item()), int(gcount.item())
    rt, ri, g = run_t[:c].cpu().numpy(), run_idx[:c].cpu().numpy(), group[:gc].cpu().numpy()
    return [Segment(float(rt[a, 0]), float(rt[b, 1]), int(ri[a, 2])) for a, b in g]


def frame_reassign(y: np.ndarray, sr: int, speech_mask: list, segs: list, embs: np.ndarray,
                   smooth_step: float = 0.1, win: float = 1.0, batch_size: int = 128) -> list:
    """anti_stick_diarize.py:390-460.  `batch_size` is accepted for signature compatibility: the device pass embeds
    all speech windows from one offset list (the encoder chunks internally by its workspace)."""
    if not segs or np.asarray(embs).size == 0:
        return []
    spk_ids, c_matrix = speaker_centroids(segs, embs)
    if c_matrix.size == 0:
        return segs                                                         # :407-408
    out = reassign_windows(y, sr, speech_mask, spk_ids, c_matrix, smooth_step, win)
    if not out:
        # no speech windows (:417-418 returns the input); windows but no surviving run gives [] in the reference too
        win_samples, step_samples = int(win * sr), int(smooth_step * sr)
        if _get_speech_windows(y, sr, speech_mask, win_samples, step_samples)[1].size == 0:
            return segs
    return out
