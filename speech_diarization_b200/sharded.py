"""Multi-GPU driver for the hot path (new design — the reference is single-device).

One process per GPU.  Windows are independent units: rank r embeds the contiguous window range
``shard_range(n_windows, r, world)`` from its slice of the audio with NO communication.  The only
exchange step is one all-gather of the L2-normalised [N/G, 192] f32 embeddings (<= 30 MB at 8 h of
audio), after which each rank computes its [N/G, N] row block of the cosine-distance matrix; the row
blocks are gathered on rank 0, which runs the AHC (the full matrix fits one B200 up to N ~ 100k)
and broadcasts the labels.

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only; every function takes the
process group and works on whatever device its tensors live on.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of the n units owned by `rank`; every rank gets ceil(n/world)
    except the last non-empty one (later ranks may be empty)."""
    per = math.ceil(n / world) if n else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def window_count(n_samples: int, win: int, hop: int) -> int:
    """frame_audio semantics (vad.py:9-16): no padding, tail dropped."""
    return 0 if n_samples < win else 1 + (n_samples - win) // hop


def audio_slice_for(lo: int, hi: int, win: int, hop: int) -> tuple[int, int]:
    """Sample range a rank needs for windows [lo, hi): its hop-spaced starts plus one window."""
    if hi <= lo:
        return 0, 0
    return lo * hop, (hi - 1) * hop + win


def gather_embeddings(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of the per-rank embedding shards -> [n_total, D] on every rank.  Shards follow
    shard_range (equal size `per`, the tail ranks padded), so one all_gather_into_tensor does it."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n_total]
    D = local.shape[1]
    per = math.ceil(n_total / world)
    padded = torch.zeros((per, D), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * per, D), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]


def gather_row_blocks(block: torch.Tensor, n_total: int, dst: int = 0, group=None):
    """Row blocks [rows_r, N] (rows_r per shard_range) -> the full [N, N] matrix on rank `dst`
    (None elsewhere)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return block
    rank = dist.get_rank(group)
    per = math.ceil(n_total / world)
    padded = torch.zeros((per, n_total), dtype=block.dtype, device=block.device)
    padded[: block.shape[0]] = block
    if rank == dst:
        full = torch.empty((world * per, n_total), dtype=block.dtype, device=block.device)
        parts = list(full.split(per, dim=0))
        dist.gather(padded, parts, dst=dst, group=group)
        return full[:n_total]
    dist.gather(padded, None, dst=dst, group=group)
    return None


class _PhaseTimer:
    """Per-phase device times of the clustering stage: CUDA events on the current stream when the tensors live
    on a GPU, perf_counter in the gloo/CPU tests.  `read()` synchronises once, after the last phase."""

    def __init__(self, device: torch.device, enabled: bool):
        self.cuda = enabled and device.type == "cuda"
        self.enabled = enabled
        self.marks: list = []
        self.names: list[str] = []

    def mark(self, name: str | None = None) -> None:
        if not self.enabled:
            return
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append(ev)
        else:
            import time
            self.marks.append(time.perf_counter())
        if name is not None:
            self.names.append(name)

    def read(self) -> dict:
        if not self.enabled or len(self.marks) < 2:
            return {}
        if self.cuda:
            self.marks[-1].synchronize()
            return {n: self.marks[i].elapsed_time(self.marks[i + 1]) for i, n in enumerate(self.names)}
        return {n: 1e3 * (self.marks[i + 1] - self.marks[i]) for i, n in enumerate(self.names)}


def cluster_sharded(emb_all: torch.Tensor, cos_thr: float, group=None, distance_fn=None, ahc_fn=None,
                    timings: dict | None = None) -> torch.Tensor:
    """Row-block sharded affinity + AHC on rank 0 + label broadcast.  distance_fn(emb_all, row0, rows)
    and ahc_fn(dist, threshold) default to the CUDA kernels; the CPU tests inject stand-ins.
    A rank whose shard is empty (n < world * (world - 1) can leave the last ranks without rows) contributes an
    empty block and still takes part in the gather and the broadcast.  `timings`, when given, receives the
    milliseconds of each phase on this rank: affinity_rowblock, gather_rowblocks, ahc, broadcast_labels."""
    if distance_fn is None or ahc_fn is None:
        from . import clustering
        distance_fn = distance_fn or clustering.cosine_distance_device
        ahc_fn = ahc_fn or (lambda d, thr: clustering.ahc_average_device(d, thr)[0])
    n = emb_all.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n, rank, world)
    tm = _PhaseTimer(emb_all.device, timings is not None)
    tm.mark()
    if hi > lo:
        block = distance_fn(emb_all, lo, hi - lo)
    else:
        block = torch.empty((0, n), dtype=torch.float32, device=emb_all.device)
    tm.mark("affinity_rowblock")
    full = gather_row_blocks(block, n, 0, group)
    tm.mark("gather_rowblocks")
    labels = torch.empty((n,), dtype=torch.int32, device=emb_all.device)
    if rank == 0:
        labels.copy_(ahc_fn(full.contiguous(), 1 - cos_thr))
    del full
    tm.mark("ahc")
    if world > 1:
        dist.broadcast(labels, src=0, group=group)
    tm.mark("broadcast_labels")
    if timings is not None:
        timings.update(tm.read())
    return labels


def embed_windows_sharded(audio: torch.Tensor, win: int, hop: int, encoder, group=None,
                          n_total_samples: int | None = None, timings: dict | None = None):
    """Embeds this rank's window range and all-gathers the L2-normalised embeddings.

    audio: 1-D f32 tensor on this rank's device holding EITHER the full recording (n_total_samples None) OR only
    this rank's slice of it — samples [a0, a1) = audio_slice_for(*shard_range(n, rank, world), win, hop) of a
    recording of n_total_samples samples ("each rank receives only its audio slice plus a win-hop
    overlap").  Returns (all embeddings [N, 192] on every rank, (lo, hi) = this rank's window range)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sliced = n_total_samples is not None
    n = window_count(int(n_total_samples) if sliced else audio.numel(), win, hop)
    lo, hi = shard_range(n, rank, world)
    a0, a1 = audio_slice_for(lo, hi, win, hop)
    if sliced:
        if audio.numel() < a1 - a0:
            raise ValueError(f"rank {rank}: slice of {audio.numel()} samples is shorter than the {a1 - a0} "
                             f"its windows [{lo}, {hi}) cover")
        local_audio = audio
    else:
        local_audio = audio[a0:]
    tm = _PhaseTimer(audio.device, timings is not None)
    tm.mark()
    local = encoder.embed_device(local_audio, hop, hi - lo, win, l2_normalize=True)
    tm.mark("embed_shard")
    out = gather_embeddings(local, n, group)
    tm.mark("allgather_embeddings")
    if timings is not None:
        timings.update(tm.read())
    return out, (lo, hi)


def embed_windows_sharded_host(host_slice: torch.Tensor, win: int, hop: int, encoder, n_total_samples: int,
                               group=None, timings: dict | None = None, chunk_windows: int = 2048):
    """embed_windows_sharded for a rank whose audio slice is still in (page-locked) HOST memory: the slice is
    uploaded in chunks on a side stream while the previous chunk's windows are embedded, so only the first chunk's
    copy is exposed (the whole-slice upload is 10-17 ms of the 8 h corpus leg).

    host_slice: 1-D f32 CPU tensor with this rank's samples [a0, a1) = audio_slice_for(*shard_range(n, rank, world),
    win, hop) of a recording of n_total_samples samples; pinned memory makes the copies asynchronous.  Returns
    (all embeddings [N, 192] on every rank, (lo, hi))."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = window_count(int(n_total_samples), win, hop)
    lo, hi = shard_range(n, rank, world)
    a0, a1 = audio_slice_for(lo, hi, win, hop)
    if host_slice.dim() != 1 or host_slice.dtype != torch.float32 or host_slice.is_cuda:
        raise ValueError("embed_windows_sharded_host needs a 1-D float32 CPU tensor")
    if host_slice.numel() < a1 - a0:
        raise ValueError(f"rank {rank}: slice of {host_slice.numel()} samples is shorter than the {a1 - a0} "
                         f"its windows [{lo}, {hi}) cover")
    dev = encoder.device
    nw = hi - lo
    tm = _PhaseTimer(dev, timings is not None)
    tm.mark()
    local = torch.empty((nw, 192), dtype=torch.float32, device=dev)
    if nw > 0:
        buf = torch.empty(a1 - a0, dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(main)            # the buffer's allocation is ordered before the copies
        copied = 0
        step = max(1, int(chunk_windows))
        for w0 in range(0, nw, step):
            w1 = min(nw, w0 + step)
            end = (w1 - 1) * hop + win    # samples of the slice the windows [w0, w1) need
            with torch.cuda.stream(side):
                buf[copied:end].copy_(host_slice[copied:end], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            copied = end
            main.wait_event(ev)
            encoder.embed_device(buf[w0 * hop:end], hop, w1 - w0, win, l2_normalize=True, out=local[w0:w1])
        buf.record_stream(side)
    tm.mark("upload_and_embed_shard")
    out = gather_embeddings(local, n, group)
    tm.mark("allgather_embeddings")
    if timings is not None:
        timings.update(tm.read())
    return out, (lo, hi)


def diarize_windows(audio: torch.Tensor, sr: int, encoder, win_s: float = 1.5, hop_s: float = 0.75,
                    cos_thr: float = 0.68, group=None) -> list:
    """Window-level diarization of one recording on 1..G GPUs: embed -> all-gather -> row-block
    affinity -> AHC -> run-length encode to (start, end, "SPEAKER_xx") tuples
    (output format of diarization_baseline.py:259-261)."""
    import numpy as np
    from .anti_stick_diarize import _labels_to_segments, merge_adjacent
    from .diarization_baseline import segments_to_tuples
    win, hop = int(round(win_s * sr)), int(round(hop_s * sr))
    emb, _ = embed_windows_sharded(audio, win, hop, encoder, group)
    labels = cluster_sharded(emb, cos_thr, group).cpu().numpy()
    n = labels.shape[0]
    starts = np.arange(n) * hop
    segs = _labels_to_segments(starts, np.arange(n), labels, sr, audio.numel() / sr)
    return segments_to_tuples(merge_adjacent(segs, gap=0.05))
