"""SURVEY §8f rank 2 on the GPU: the callers and operators either side of the embedding kernels in
anti_stick_diarize.py — SCD (z-score + peak picking), speaker centroids, dense frame reassignment, run-length
label -> segment encoding, neighbour merging, variable-length segment batches.

Goldens (tests/golden/f2_ref.npz, windows_ref.npz) are outputs of the REFERENCE's own functions run with a
recording stand-in encoder (tests/golden/make_golden.py::make_f2_golden); the tests hand the product the very
embeddings the stand-in returned, so everything after the encoder is compared number for number."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_emb
from speech_diarization_b200 import anti_stick_diarize as asd, dense_ops

pytestmark = pytest.mark.gpu


def _segs(a):
    return [asd.Segment(float(s), float(e)) for s, e in a]


def _inject(monkeypatch, rows, normalize=False):
    """Make the dense passes 'embed' by returning the golden's recorded embeddings, in call order."""
    state = {"pos": 0, "calls": []}

    def fake(audio, offsets, n_samples, l2_normalize=False):
        n = len(offsets)
        e = rows[state["pos"]:state["pos"] + n].copy()
        state["pos"] += n
        state["calls"].append((np.asarray(offsets).copy(), int(n_samples)))
        if l2_normalize:
            e /= np.linalg.norm(e, axis=1, keepdims=True) + 1e-8          # anti_stick_diarize.py:430
        return torch.from_numpy(e).cuda()

    monkeypatch.setattr(asd, "_embed_rows_device", fake)
    return state


@pytest.mark.parametrize("tag,kw", [("scd", dict(win_ms=1000.0, hop_ms=200.0, thr=1.25, min_speech_ms=1000.0)),
                                    ("scd2", dict(win_ms=800.0, hop_ms=100.0, thr=0.8, min_speech_ms=500.0))])
def test_scd_split_segments_matches_reference(monkeypatch, tag, kw):
    g = golden("f2_ref.npz")
    st = _inject(monkeypatch, g[f"{tag}_embs"])
    y = np.zeros(int(g["ylen"]), np.float32)
    out = asd.scd_split_segments(y, int(g["sr"]), _segs(g["scd_in"]), **kw)
    np.testing.assert_array_equal(np.array([[s.start, s.end] for s in out]), g[f"{tag}_out"])
    assert st["pos"] == len(g[f"{tag}_embs"]) and len(st["calls"]) == 1       # ONE pass over all segments' windows
    if tag == "scd":
        # the windows of the four long-enough segments, in segment order, hop 3200 samples
        offs, n = st["calls"][0]
        assert n == 16000 and len(offs) == int(g["scd_calls"].sum())
        assert offs[0] == 0 and offs[1] - offs[0] == 3200


def test_scd_peaks_kernel_against_scipy():
    """sd_scd_peaks on random per-segment embeddings vs numpy z-score + scipy.signal.find_peaks, incl. flat tops."""
    from scipy.signal import find_peaks
    rng = np.random.default_rng(3)
    counts = [3, 4, 40, 2, 300, 17, 1000]
    emb = rng.standard_normal((sum(counts), 192)).astype(np.float32)
    emb[50:53] = emb[50]                      # identical neighbours: zero distances, a flat valley
    emb[400:900:7] *= 3.0
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    for thr in (0.5, 1.25):
        peak, z = dense_ops.scd_peaks_device(torch.from_numpy(emb).cuda(), torch.from_numpy(off).cuda(), thr)
        peak, z = peak.cpu().numpy(), z.cpu().numpy()
        for s, c in enumerate(counts):
            e = emb[off[s]:off[s + 1]]
            if c < 2:
                assert not peak[off[s]:off[s + 1]].any()
                continue
            sims = np.einsum("id,id->i", e[:-1], e[1:]) / (np.linalg.norm(e[:-1], axis=1) * np.linalg.norm(e[1:], axis=1) + 1e-8)
            d = 1 - sims
            zr = (d - d.mean()) / d.std() if np.std(d) > 1e-6 else d
            np.testing.assert_allclose(z[off[s]:off[s + 1] - 1], zr, rtol=0, atol=2e-4)
            ref, _ = find_peaks(z[off[s]:off[s + 1] - 1], height=thr)          # same z: index-exact comparison
            np.testing.assert_array_equal(np.flatnonzero(peak[off[s]:off[s + 1]]), ref)
    # flat tops: the middle (floor) of a plateau, never the borders
    zflat = np.array([0, 2, 2, 2, 0, 3, 3, 1, 5, 5], np.float32)
    ref, _ = find_peaks(zflat, height=1.0)
    assert ref.tolist() == [2, 5]


def test_speaker_centroids_matches_reference():
    g = golden("f2_ref.npz")
    segs = [asd.Segment(float(a), float(b), None if k == -2 else int(k)) for a, b, k in g["cent_segs"]]
    ids, cents = asd.speaker_centroids(segs, g["cent_embs"])
    np.testing.assert_array_equal(ids, g["cent_ids"])
    assert cents.dtype == np.float32 and cents.shape == g["cent_out"].shape
    np.testing.assert_allclose(cents, g["cent_out"], rtol=0, atol=2e-7)
    ids0, c0 = asd.speaker_centroids([asd.Segment(0, 1, -1), asd.Segment(1, 2)], g["cent_embs"][:2])
    assert ids0.shape == (0,) and c0.shape == (0, 192)                        # :345-346


def test_frame_reassign_matches_reference(monkeypatch):
    g = golden("f2_ref.npz")
    st = _inject(monkeypatch, g["fr_window_embs"])
    y = np.zeros(int(g["ylen"]), np.float32)
    segs = [asd.Segment(float(a), float(b), int(k)) for a, b, k in g["fr_segs"]]
    mask = _segs(g["fr_mask"])
    out = asd.frame_reassign(y, int(g["sr"]), mask, segs, g["fr_embs_in"], smooth_step=0.1, win=1.0, batch_size=128)
    np.testing.assert_array_equal(np.array([[s.start, s.end, s.spk] for s in out]), g["fr_out"])
    assert st["pos"] == len(g["fr_window_embs"]) and len(st["calls"]) == 1
    assert asd.frame_reassign(y, 16000, mask, [], g["fr_embs_in"]) == []       # :400-401
    assert asd.frame_reassign(y, 16000, [], segs, g["fr_embs_in"]) is segs     # no speech windows: input returned (:417-418)


def test_labels_to_segments_and_merge_match_reference():
    g = golden("windows_ref.npz")
    ws, vi = g["window_starts"], g["valid_indices"]
    segs = asd._labels_to_segments(ws, vi, g["window_labels"], 16000, int(g["ylen"]) / 16000)
    np.testing.assert_array_equal(np.array([[s.start, s.end, s.spk] for s in segs]), g["segs"])
    merged = asd.merge_adjacent(segs, 0.05)
    np.testing.assert_array_equal(np.array([[s.start, s.end, s.spk] for s in merged]), g["merged"])
    assert asd.merge_adjacent([]) == []
    assert asd._labels_to_segments(np.arange(0), np.arange(0), np.arange(0), 16000, 1.0) == []
    # None speakers, scores, sub- and super-gap neighbours (reference output in f2_ref.npz)
    f = golden("f2_ref.npz")
    ml = [asd.Segment(float(a), float(b), None if k == -2 else int(k), None if sc == -1 else float(sc)) for a, b, k, sc in f["merge_in"]]
    mm = asd.merge_adjacent(ml, gap=0.05)
    got = np.array([[s.start, s.end, -2 if s.spk is None else s.spk, -1.0 if s.score is None else s.score] for s in mm])
    np.testing.assert_array_equal(got, f["merge_out"])
    assert mm[1] is ml[2] and mm[4] is ml[7]                                   # untouched segments are the same objects


def test_label_runs_large_random_against_numpy():
    """35 990 windows (1 h at the reference's 0.1 s step): run-length + merge kernels vs the reference's numpy logic."""
    from oracle import cluster_oracle as co
    rng = np.random.default_rng(9)
    n = 35990
    ws = np.arange(n) * 1600
    labels_full = np.repeat(rng.integers(-1, 4, n // 9 + 1), 9)[:n]
    vi = np.flatnonzero(labels_full >= 0)
    wl = labels_full[vi]
    max_t = (n * 1600 + 16000) / 16000
    ref = co.labels_to_segments(ws, vi, wl, 16000, max_t)
    got = asd._labels_to_segments(ws, vi, wl, 16000, max_t)
    assert [(s.start, s.end, s.spk) for s in got] == [(s.start, s.end, s.spk) for s in ref]
    for gap in (0.0, 0.05, 0.95, 5.0):
        mr, mg = co.merge_adjacent(ref, gap), asd.merge_adjacent(got, gap)
        assert [(s.start, s.end, s.spk) for s in mg] == [(s.start, s.end, s.spk) for s in mr]


def test_embed_segments_batching_matches_reference(monkeypatch):
    g = golden("windows_ref.npz")
    calls = []

    def fake(batch):                                  # the golden's recording stand-in, on the device-built batch
        b = batch.cpu().numpy()
        calls.append(b.copy())
        return torch.from_numpy(np.tile(b.sum(axis=1, keepdims=True), (1, 192)).astype(np.float32)).cuda()

    monkeypatch.setattr(asd, "_encode_batch_device", fake)
    segs = [asd.Segment(a, b) for a, b in g["embed_segs"]]
    out = asd.embed_segments(g["embed_y"], 16000, segs, batch_size=2)
    np.testing.assert_array_equal(np.array([c.shape for c in calls]), g["embed_shapes"])
    np.testing.assert_array_equal(np.concatenate([c.sum(axis=1) for c in calls]), g["embed_sums"])
    np.testing.assert_array_equal(out, g["embed_out"])
    empty = asd.embed_segments(g["embed_y"], 16000, [])
    assert empty.shape == tuple(g["empty_out_shape"]) == (0, 192) and empty.dtype == np.float32


def test_embed_offsets_equals_strided_and_batched_paths(oracle_model):
    """sd_ecapa_embed_offsets: windows at arbitrary offsets give bit-identical embeddings to the strided path and
    to materialised batches (slot-invariant forward)."""
    from conftest import synth_wave
    from speech_diarization_b200 import speech_encode as se
    y = torch.from_numpy(synth_wave(1, 16000 * 12, 4)[0]).cuda()
    enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=32, max_samples=16000)
    try:
        strided = enc.embed_device(y, 1600, 100, 16000)
        offs = np.arange(100, dtype=np.int64) * 1600
        assert torch.equal(enc.embed_offsets_device(y, offs, 16000), strided)
        pick = np.array([99, 3, 3, 57, 0], dtype=np.int64)
        got = enc.embed_offsets_device(y, offs[pick], 16000)
        assert torch.equal(got, strided[torch.from_numpy(pick).cuda()])
        with pytest.raises(ValueError):
            enc.embed_offsets_device(y, np.array([y.numel() - 100]), 16000)
        assert enc.embed_offsets_device(y, np.zeros(0, np.int64), 16000).shape == (0, 192)
    finally:
        enc.close()
