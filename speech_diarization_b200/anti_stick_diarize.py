"""Drop-in for the hot-path pieces of /root/reference/anti_stick_diarize.py: the callers and
data formats either side of the embedding + affinity kernels.
VAD, HDBSCAN, conservative_merge and the CLI are out of scope."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .clustering import (adjacent_cosine_device, cosine_distance_device, l2_normalize_device,
                         to_cuda_embeddings, window_argmax_device)
from .speech_encode import ecapa_encode_batch, using_ecapa_encoder
from ._device import to_device_f32


@dataclass
class Segment:                                   # anti_stick_diarize.py:21-26
    start: float
    end: float
    spk: int | None = None
    score: float | None = None


def embed_segments(y: np.ndarray, sr: int, segs: list, batch_size: int = 32,
                   min_duration_ms: float = 500.0, pad_duration_ms: float = 150.0) -> np.ndarray:
    """anti_stick_diarize.py:130-172 — variable-length segments, zero-padded per batch of
    `batch_size` to that batch's longest snippet (padding is signal: no wav_lens)."""
    num_segs = len(segs)
    if num_segs == 0:
        return np.empty((0, 192), dtype=np.float32)          # :143-144
    min_duration_samples = int(min_duration_ms / 1000.0 * sr)
    pad_samples = int(pad_duration_ms / 1000.0 * sr)
    embs = []
    for i in range(0, num_segs, batch_size):
        batch_snippets = []
        for seg in segs[i:i + batch_size]:
            s, e = int(seg.start * sr), int(seg.end * sr)
            snippet = y[s:e]
            if snippet.shape[0] < min_duration_samples:      # :157-160
                snippet = y[max(0, s - pad_samples):min(len(y), e + pad_samples)]
            batch_snippets.append(snippet)
        max_len = max(len(s) for s in batch_snippets)
        wav_batch = np.zeros((len(batch_snippets), max_len), dtype=np.float32)
        for k, s in enumerate(batch_snippets):
            wav_batch[k, :len(s)] = s
        embs.append(ecapa_encode_batch(wav_batch))           # :168
    return np.concatenate(embs, axis=0)


def cosine_distance(embs: np.ndarray) -> np.ndarray:
    """Head of cluster_hdbscan (anti_stick_diarize.py:176-177): L2-normalise, D = 1 - cos."""
    x = to_cuda_embeddings(embs)
    return cosine_distance_device(l2_normalize_device(x)).cpu().numpy()


def adjacent_cosine(embs: np.ndarray) -> np.ndarray:
    """anti_stick_diarize.py:102-104."""
    return adjacent_cosine_device(to_cuda_embeddings(embs)).cpu().numpy()


def _get_speech_windows(y: np.ndarray, sr: int, speech_mask: list, win_samples: int, step_samples: int):
    """anti_stick_diarize.py:352-367."""
    max_t = len(y) / sr
    hop_s = 0.01
    n_frames = math.ceil(max_t / hop_s)
    smask = np.zeros(n_frames, dtype=bool)
    for sm in speech_mask:
        s, e = int(sm.start / hop_s), int(sm.end / hop_s)
        smask[s:e] = True
    window_starts = np.arange(0, len(y) - win_samples, step_samples)
    window_centers_s = (window_starts + win_samples / 2) / sr
    window_center_frames = np.clip((window_centers_s / hop_s).astype(int), 0, n_frames - 1)
    valid_indices = np.where(smask[window_center_frames])[0]
    return window_starts, valid_indices


def _labels_to_segments(window_starts: np.ndarray, valid_indices: np.ndarray, window_labels: np.ndarray,
                        sr: int, max_t: float) -> list:
    """anti_stick_diarize.py:370-386."""
    full_labels = np.full(len(window_starts), -1, dtype=int)
    full_labels[valid_indices] = window_labels
    change_points = np.where(np.diff(full_labels, prepend=np.nan))[0]
    refined_segs = []
    for start_idx, end_idx in zip(change_points, list(change_points[1:]) + [len(full_labels)]):
        spk_id = int(full_labels[start_idx])
        if spk_id != -1:
            start_time = window_starts[start_idx] / sr
            end_time = window_starts[end_idx] / sr if end_idx < len(window_starts) else max_t
            if end_time > start_time:
                refined_segs.append(Segment(start_time, end_time, spk_id))
    return refined_segs


def merge_adjacent(segments: list, gap: float = 0.05) -> list:
    """anti_stick_diarize.py:464-475."""
    if not segments:
        return []
    merged = [segments[0]]
    for next_seg in segments[1:]:
        last_seg = merged[-1]
        if next_seg.spk == last_seg.spk and (next_seg.start - last_seg.end) <= gap:
            merged[-1] = Segment(last_seg.start, next_seg.end, last_seg.spk)
        else:
            merged.append(next_seg)
    return merged


def reassign_windows(y: np.ndarray, sr: int, speech_mask: list, spk_ids, c_matrix: np.ndarray,
                     smooth_step: float = 0.1, win: float = 1.0) -> list:
    """The dense pass of frame_reassign (anti_stick_diarize.py:411-460) given unit-norm speaker
    centroids: embed every `win`-second window whose centre is speech (step `smooth_step`),
    score against the centroids, arg-max, run-length encode and merge.  The audio is uploaded
    once and windows are addressed in place; embeddings never leave the device."""
    win_samples = int(win * sr)
    step_samples = int(smooth_step * sr)
    window_starts, valid_indices = _get_speech_windows(y, sr, speech_mask, win_samples, step_samples)
    if valid_indices.size == 0 or len(c_matrix) == 0:
        return []
    enc = using_ecapa_encoder()
    audio = to_device_f32(y, enc.device)
    # valid windows form runs of consecutive indices: embed each run with stride = step
    embs = torch.empty((len(valid_indices), 192), dtype=torch.float32, device=enc.device)
    run_start = 0
    vi = valid_indices
    while run_start < len(vi):
        run_end = run_start
        while run_end + 1 < len(vi) and vi[run_end + 1] == vi[run_end] + 1:
            run_end += 1
        n_run = run_end - run_start + 1
        off = int(window_starts[vi[run_start]])
        enc.embed_device(audio[off:], step_samples, n_run, win_samples, l2_normalize=True,
                         out=embs[run_start:run_start + n_run])
        run_start = run_end + 1
    best, _ = window_argmax_device(embs, to_device_f32(c_matrix, enc.device))
    window_labels = np.asarray(spk_ids)[best.cpu().numpy()]
    refined = _labels_to_segments(window_starts, valid_indices, window_labels, sr, len(y) / sr)
    return merge_adjacent(refined, gap=0.05)
