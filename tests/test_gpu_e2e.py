"""End-to-end parity on BASELINE.json's configurations 1 and 5 (small enough for the CPU oracle):

config 1: a 60 s synthetic 2-speaker 16 kHz clip, random-init ECAPA -> 1.5 s / 0.75 s windows -> embeddings ->
          cosine affinity -> AHC (cos_thr 0.68) -> (start, end, speaker) tuples / RTTM.
config 5: the dense short-window reassignment pass of anti_stick_diarize.frame_reassign (1.0 s windows,
          0.5 s hop) against unit speaker centroids.

The GPU pipeline must give the same window labels (up to permutation) and therefore the same segments
as the oracle pipeline (CPU fp32 ECAPA + sklearn), per the north_star gates."""
import numpy as np
import pytest
import torch

from oracle import cluster_oracle as co
from oracle import ecapa_oracle as eo
from speech_diarization_b200 import anti_stick_diarize as asd
from speech_diarization_b200 import diarization_baseline as db
from speech_diarization_b200 import sharded, speech_encode as se, vad

pytestmark = pytest.mark.gpu
SR = 16000


def two_speaker_clip(seconds=60, seed=0):
    """Alternating 2-6 s turns of two harmonic 'voices' (different f0 and spectral envelope) with 0.3 s gaps."""
    rng = np.random.default_rng(seed)
    n = seconds * SR
    t = np.arange(n) / SR
    voices = []
    for k, (f0, tilt) in enumerate(((110.0, 1.0), (205.0, 2.2))):
        sig = sum(np.sin(2 * np.pi * f0 * h * t + 0.9 * h * (k + 1)) / h ** tilt for h in range(1, 30) if f0 * h < 7500)
        voices.append(0.12 * sig / np.abs(sig).max())
    y = np.zeros(n)
    spk_of = np.full(n, -1)
    pos, spk = 0, 0
    while pos < n:
        dur = int(rng.uniform(2.0, 6.0) * SR)
        y[pos:pos + dur] = voices[spk][pos:pos + dur]
        spk_of[pos:pos + dur] = spk
        pos += dur + int(0.3 * SR)
        spk ^= 1
    y += 0.003 * rng.standard_normal(n)
    return y.astype(np.float32), spk_of


def test_config1_60s_two_speaker_clip_matches_oracle_pipeline(oracle_model, encoder, tmp_path):
    y, _ = two_speaker_clip(60, seed=1)
    frames = vad.frame_audio(y, SR, 1500.0, 750.0)
    assert frames.shape == (79, 24000)                       # 60 s -> 79 windows
    # oracle pipeline (the reference's CPU path): encode_batch -> cluster_embeddings("agglo")
    with torch.inference_mode():
        emb_ref = eo.encode_batch(oracle_model, torch.from_numpy(np.ascontiguousarray(frames))).squeeze(1).numpy()
    lab_ref = co.cluster_embeddings(emb_ref, "agglo", 0.68)
    # product pipeline through the reference-facing callables
    se.register_ecapa_state_dict(oracle_model.state_dict())
    try:
        emb = se.ecapa_encode_batch(frames)
        from speech_diarization_b200.diar_diag import cluster_embeddings
        lab = cluster_embeddings(emb, method="agglo", cos_thr=0.68)
    finally:
        se.register_ecapa_state_dict(None)
    cos = np.sum(emb * emb_ref, 1) / (np.linalg.norm(emb, axis=1) * np.linalg.norm(emb_ref, axis=1))
    assert (1 - cos).max() < 1e-4
    assert co.same_partition(lab, lab_ref)
    assert 2 <= len(set(lab.tolist())) <= 79
    # segments / RTTM from window labels: identical to what the oracle labels give (after renaming)
    starts = np.arange(len(lab)) * 12000
    def to_tuples(labels):
        first = {}
        ren = np.array([first.setdefault(int(l), len(first)) for l in labels])     # rename by first appearance
        segs = asd.merge_adjacent(asd._labels_to_segments(starts, np.arange(len(ren)), ren, SR, len(y) / SR), 0.05)
        return db.segments_to_tuples(segs)
    tup, tup_ref = to_tuples(lab), to_tuples(lab_ref)
    assert tup == tup_ref
    assert all(a[0] <= b[0] for a, b in zip(tup, tup[1:]))                          # sorted by start
    p = tmp_path / "clip.rttm"
    db.write_rttm(tup, p)
    lines = p.read_text().splitlines()
    assert len(lines) == len(tup) and all(l.startswith("SPEAKER clip 1 ") and l.endswith(" <NA> <NA>") for l in lines)
    # the multi-GPU driver on a single rank gives the same tuples
    tup_dev = sharded.diarize_windows(torch.from_numpy(y).cuda(), SR, encoder)
    assert [(round(a, 6), round(b, 6)) for a, b, _ in tup_dev] == [(round(a, 6), round(b, 6)) for a, b, _ in tup]


def test_config5_dense_reassignment_pass_matches_oracle(oracle_model):
    y, spk_of = two_speaker_clip(40, seed=2)
    win, step = 1.0, 0.5
    speech = [asd.Segment(0.0, len(y) / SR)]                   # whole clip is "speech" for this test
    ws, vi = asd._get_speech_windows(y, SR, speech, int(win * SR), int(step * SR))
    snippets = np.stack([y[s:s + SR] for s in ws[vi]])
    with torch.inference_mode():
        emb_ref = eo.encode_batch(oracle_model, torch.from_numpy(snippets)).squeeze(1).numpy()
    # unit centroids of the two planted speakers from the oracle embeddings
    centre_spk = spk_of[(ws[vi] + SR // 2)]
    cents = []
    for k in (0, 1):
        c = emb_ref[centre_spk == k].mean(0)
        cents.append(c / (np.linalg.norm(c) + 1e-8))
    c_matrix = np.stack(cents).astype(np.float32)
    labels_ref = co.window_argmax(emb_ref.copy(), c_matrix)
    segs_ref = co.merge_adjacent(co.labels_to_segments(ws, vi, labels_ref, SR, len(y) / SR), 0.05)
    se.register_ecapa_state_dict(oracle_model.state_dict())
    try:
        segs = asd.reassign_windows(y, SR, speech, np.array([0, 1]), c_matrix, smooth_step=step, win=win)
    finally:
        se.register_ecapa_state_dict(None)
    assert [(s.start, s.end, s.spk) for s in segs] == [(s.start, s.end, s.spk) for s in segs_ref]
    assert len(segs) >= 4


def test_host_buffer_path_equals_device_path(oracle_model, monkeypatch):
    """ecapa_encode_batch from host memory (sd_ecapa_embed_host: chunked upload overlapped with fbank) gives
    bit-identical embeddings to the explicit upload + device path, for overlapping window views, contiguous
    [B, n] batches, few windows (single chunk) and non-contiguous input (falls back to the torch upload); and for
    batches whose quarter-chunks end on 128-row boundaries, where with SD_ECAPA_PIPE=1 the front of the trunk
    (block0 .. b1.tdnn2) runs per upload chunk (96 windows: odd m-block count per chunk, 128 windows: even)."""
    import numpy as np
    from speech_diarization_b200 import speech_encode, vad
    from conftest import synth_wave
    speech_encode.register_ecapa_state_dict(oracle_model.state_dict())
    try:
        y = synth_wave(1, 16000 * 40, 5)[0]
        frames = vad.frame_audio(y, 16000, 1000.0, 250.0)           # 157 overlapping windows (4 upload chunks)
        dense = np.ascontiguousarray(frames[:70])                    # contiguous [70, 16000]
        few = frames[:5]                                             # one chunk
        strided = dense[::2]                                         # not contiguous, not a hop view
        cases = {"view": frames, "dense": dense, "few": few, "strided": strided,
                 "piped96": frames[:96], "piped128": np.ascontiguousarray(frames[:128])}
        monkeypatch.setenv("SD_ECAPA_PIPE", "1")                     # off by default (measured slower)
        host = {k: speech_encode.ecapa_encode_batch(v) for k, v in cases.items()}
        for _ in range(3):   # the third and later runs of a shape replay the captured tail graph
            np.testing.assert_array_equal(speech_encode.ecapa_encode_batch(cases["piped128"]), host["piped128"])
        monkeypatch.setenv("SD_ECAPA_HOST_PATH", "0")
        dev = {k: speech_encode.ecapa_encode_batch(v) for k, v in cases.items()}
        for k in cases:
            assert host[k].shape == (len(cases[k]), 192) and host[k].dtype == np.float32
            np.testing.assert_array_equal(host[k], dev[k], err_msg=k)
        np.testing.assert_array_equal(host["view"][:70], host["dense"])
    finally:
        speech_encode.register_ecapa_state_dict(None)


def test_sharded_embedding_from_host_slice_equals_device_path(oracle_model):
    """embed_windows_sharded_host (chunked upload on a side stream, overlapped with the embedding of the previous
    chunk) returns exactly what embed_windows_sharded returns for the same samples already on the device."""
    from conftest import synth_wave
    y = synth_wave(1, 16000 * 60, 12)[0]
    enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=64, max_samples=24000)
    try:
        ref, rng_ref = sharded.embed_windows_sharded(torch.from_numpy(y).cuda(), 24000, 12000, enc)
        for pinned in (True, False):
            host = torch.from_numpy(y.copy())
            if pinned:
                host = host.pin_memory()
            t = {}
            got, rng = sharded.embed_windows_sharded_host(host, 24000, 12000, enc, n_total_samples=y.size, timings=t,
                                                           chunk_windows=17)
            assert rng == rng_ref and torch.equal(got, ref)
            assert "upload_and_embed_shard" in t
        with pytest.raises(ValueError):
            sharded.embed_windows_sharded_host(torch.from_numpy(y[:1000].copy()), 24000, 12000, enc, n_total_samples=y.size)
    finally:
        enc.close()
