"""GPU parity of the score post-processing and VAD-mask operators through the
reference-facing callables diar_diag.asnorm_scores / viterbi_hmm and vad.hysteresis_binarize /
morph_open_close / mask_to_segments: against the goldens produced by the reference's own functions and
against the oracle on seeded inputs.  Integer / boolean results must be identical; AS-norm is floating
point (f32 on both sides): |delta| <= 2e-4 on z-scores of magnitude up to ~50 (relative 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_emb

pytestmark = pytest.mark.gpu


def _po():
    from oracle import post_oracle
    return post_oracle


# ---------------------------------------------------------------------------------- Viterbi
@pytest.mark.parametrize("tag,alpha", [("as", 0.995), ("f64", 0.9), ("ties", 0.5), ("k2", 0.995), ("t1", 0.995)])
def test_viterbi_matches_reference_golden(tag, alpha):
    from speech_diarization_b200 import diar_diag
    g = golden("post_ref.npz")
    path = diar_diag.viterbi_hmm(g[f"vt_scores_{tag}"], alpha=alpha)
    assert path.dtype == np.int32
    np.testing.assert_array_equal(path, g[f"vt_path_{tag}"])


@pytest.mark.parametrize("T,K,dtype,alpha", [(4097, 8, np.float32, 0.995), (1500, 32, np.float64, 0.8),
                                              (129, 5, np.float32, 0.3), (2, 3, np.float32, 0.995),
                                              (20000, 4, np.float32, 0.999)])
def test_viterbi_matches_oracle(T, K, dtype, alpha):
    from speech_diarization_b200 import diar_diag
    rng = np.random.default_rng(T + K)
    scores = rng.standard_normal((T, K)).astype(dtype)
    if K == 5:
        scores = np.round(scores)            # many exact ties
    np.testing.assert_array_equal(diar_diag.viterbi_hmm(scores, alpha), _po().viterbi_hmm(scores, alpha))


def test_viterbi_errors_like_reference():
    from speech_diarization_b200 import _lib, diar_diag
    with pytest.raises(ZeroDivisionError):
        diar_diag.viterbi_hmm(np.zeros((5, 1), np.float32))          # (1 - alpha) / (K - 1)
    with pytest.raises(IndexError):
        diar_diag.viterbi_hmm(np.zeros((0, 3), np.float32))          # dp[0] = scores[0]
    with pytest.raises(_lib.SdError):
        diar_diag.viterbi_hmm(np.zeros((5, 33), np.float32))         # more than 32 states


# ---------------------------------------------------------------------------------- AS-norm
def test_asnorm_matches_reference_golden():
    from speech_diarization_b200 import diar_diag
    g = golden("post_ref.npz")
    X, cent = g["as_X"], g["as_cent"]
    got = diar_diag.asnorm_scores(X, cent, X, topk=min(200, len(X)))
    assert got.dtype == np.float32 and got.shape == (300, 5)
    assert np.abs(got - g["as_self"]).max() <= 2e-4, np.abs(got - g["as_self"]).max()
    got = diar_diag.asnorm_scores(X[:70], cent, g["as_cohort"], topk=200)
    assert np.abs(got - g["as_small"]).max() <= 2e-4
    # the decision the pipeline takes from the scores (diar_diag.py:393-396) is the reference's
    np.testing.assert_array_equal(np.argmax(diar_diag.asnorm_scores(X, cent, X, 200), axis=1),
                                  np.argmax(g["as_self"], axis=1))


@pytest.mark.parametrize("nq,nr,nc,topk", [(5000, 8, 5000, 200), (257, 3, 9000, 50), (64, 2, 130, 1),
                                            (600, 4, 55000, 200)])
def test_asnorm_matches_oracle(nq, nr, nc, topk):
    from speech_diarization_b200 import diar_diag
    X, lab = synth_emb(max(nq, nc), nr, 0.05, nq + nc)
    Q, C = X[:nq], X[:nc]
    R = np.stack([X[lab == k].mean(0) for k in range(nr)]).astype(np.float32)
    got = diar_diag.asnorm_scores(Q, R, C, topk)
    ref = _po().asnorm_scores(Q, R, C, topk)
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(got - ref).max() <= 2e-5 * scale + 2e-4, (np.abs(got - ref).max(), scale)


def test_asnorm_empty_queries():
    from speech_diarization_b200 import diar_diag
    X, _ = synth_emb(40, 2, 0.05, 1)
    assert diar_diag.asnorm_scores(X[:0], X[:2], X, 10).shape == (0, 2)


# ---------------------------------------------------------------------------------- VAD mask operators
@pytest.mark.parametrize("tag", ["f32", "f64", "short"])
def test_vad_mask_ops_match_reference_golden(tag):
    from speech_diarization_b200 import vad
    g = golden("post_ref.npz")
    p, hop = g[f"vad_{tag}_probs"], float(g[f"vad_{tag}_hop"])
    m0 = vad.hysteresis_binarize(p, on=0.6, off=0.4)
    assert m0.dtype == np.bool_
    np.testing.assert_array_equal(m0, g[f"vad_{tag}_hyst"])
    np.testing.assert_array_equal(vad.morph_open_close(m0, hop, open_ms=80.0, close_ms=40.0), g[f"vad_{tag}_morph"])
    np.testing.assert_array_equal(vad.morph_open_close(m0, hop, open_ms=50.0, close_ms=70.0), g[f"vad_{tag}_morph_b"])
    np.testing.assert_array_equal(vad.morph_open_close(m0, hop, open_ms=0.0, close_ms=100.0), g[f"vad_{tag}_morph_c"])
    segs = vad.mask_to_segments(g[f"vad_{tag}_morph"], hop)
    np.testing.assert_array_equal(np.array(segs, dtype=np.float64).reshape(-1, 2), g[f"vad_{tag}_segs"])
    raw = vad.mask_to_segments(m0, hop, min_speech_ms=60.0, min_gap_ms=45.0, speech_pad_ms=30.0)
    np.testing.assert_array_equal(np.array(raw, dtype=np.float64).reshape(-1, 2), g[f"vad_{tag}_segs_raw"])


def test_vad_mask_ops_edge_cases():
    from speech_diarization_b200 import vad
    g = golden("post_ref.npz")
    np.testing.assert_array_equal(vad.hysteresis_binarize(g["vad_f32_probs"], on=0.3, off=0.7), g["vad_toggle_hyst"])
    assert vad.mask_to_segments(np.zeros(100, bool), 10.0) == []
    assert vad.mask_to_segments(np.zeros(0, bool), 10.0) == []
    full = vad.mask_to_segments(np.ones(100, bool), 10.0)
    np.testing.assert_array_equal(np.array(full), g["vad_full_segs"])
    np.testing.assert_array_equal(vad.morph_open_close(np.ones(100, bool), 10.0), g["vad_full_morph"])
    assert vad.hysteresis_binarize(np.zeros(0, np.float32)).shape == (0,)


@pytest.mark.parametrize("n,seed", [(1, 0), (1023, 1), (1024, 2), (1025, 3), (250_000, 4)])
def test_vad_mask_ops_match_oracle_at_chunk_boundaries(n, seed):
    """The scans walk 1024 frames at a time: lengths around the chunk size, and 250k frames (~42 min at 10 ms)."""
    from speech_diarization_b200 import vad
    po = _po()
    rng = np.random.default_rng(seed)
    p = np.clip(np.convolve(rng.random(n + 8), np.ones(9) / 9, mode="valid")[:n] * 1.6 - 0.3, 0, 1).astype(np.float32)
    m0 = vad.hysteresis_binarize(p, 0.55, 0.45)
    np.testing.assert_array_equal(m0, po.hysteresis_binarize(p, 0.55, 0.45))
    for open_ms, close_ms in ((80.0, 40.0), (30.0, 130.0), (10.0, 0.0)):
        np.testing.assert_array_equal(vad.morph_open_close(m0, 10.0, open_ms, close_ms),
                                      po.morph_open_close(m0, 10.0, open_ms, close_ms))
    for args in ((250.0, 100.0, 80.0), (10.0, 0.0, 0.0), (400.0, 300.0, 200.0)):
        assert vad.mask_to_segments(m0, 10.0, *args) == po.mask_to_segments(m0, 10.0, *args)


# ---------------------------------------------------------------------------------- whitening
def test_whiten_matches_reference_golden():
    """f64 on both sides; the only f32 step (the column mean) is summed in a different order, which moves the
    unit-norm output by < 1e-6.  The rank-deficient case (N = 60 < D) amplifies noise directions by
    1/sqrt(1e-6) = 1000 — the f32 centring residue of X in the null space of C, ~1e-7 relative, becomes 1e-4 of the
    row — and is held to 1e-4 (measured 2.3e-5)."""
    from speech_diarization_b200 import diar_diag
    g = golden("post_ref.npz")
    got = diar_diag.whiten_l2(g["wh_X"])
    assert got.dtype == np.float64 and got.shape == g["wh_out"].shape
    assert np.abs(got - g["wh_out"]).max() <= 1e-6, np.abs(got - g["wh_out"]).max()
    got = diar_diag.whiten_l2(g["wh_X_small"])
    assert np.abs(got - g["wh_out_small"]).max() <= 1e-4, np.abs(got - g["wh_out_small"]).max()


@pytest.mark.parametrize("N,scale", [(5000, 12.0), (20000, 1.0), (193, 30.0)])
def test_whiten_matches_oracle(N, scale):
    from speech_diarization_b200 import postproc
    X, _ = synth_emb(N, 6, 0.4, N)
    X = (X * scale).astype(np.float32)
    out, sweeps = postproc.whiten_l2_device(torch.from_numpy(X).cuda(), return_sweeps=True)
    got = out.cpu().numpy()
    ref = _po().whiten_l2(X)
    assert 2 <= sweeps < 40, sweeps                       # the Jacobi iteration converged
    assert np.abs(got - ref).max() <= 2e-6, np.abs(got - ref).max()
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-8)
    # whitened data has identity covariance (up to the 1e-6 ridge) before the row normalisation: check through
    # the property that survives it — off-diagonal correlations vanish
    if N >= 5000:
        c = np.corrcoef(got.T)
        assert np.abs(c - np.eye(192)).max() < 0.2


# ---------------------------------------------------------------------------------- diar_diag pipeline
@pytest.mark.parametrize("tag,wh,asn,vbx,thr", [("full", 1, 1, 1, 0.1), ("plain", 0, 0, 0, 0.68),
                                                 ("wh_argmax", 1, 1, 0, 0.1), ("vbx_raw", 0, 0, 1, 0.68)])
def test_label_segments_matches_reference_pipeline(tag, wh, asn, vbx, thr):
    """main()'s post-embedding path (diar_diag.py:352-411) in one GPU-resident call: cluster labels and final
    speaker path equal the reference's up to the numbering of the clusters; merged segments have the same
    boundaries."""
    from oracle import cluster_oracle as co
    from speech_diarization_b200 import diar_diag
    po = _po()
    g = golden("post_ref.npz")
    segs = [tuple(x) for x in g["pipe_segs"]]
    merged, final, labels = diar_diag.label_segments(g["pipe_X"], segs, whiten=wh, asnorm=asn, cluster="agglo",
                                                     cos_thr=thr, use_vbx=vbx)
    assert co.same_partition(labels, g[f"pipe_{tag}_labels"])
    assert co.same_partition(final, g[f"pipe_{tag}_final"])
    ref_merged = po.merge_segments(segs, g[f"pipe_{tag}_final"], gap=0.1)
    assert [(a, b) for a, b, _ in merged] == [(a, b) for a, b, _ in ref_merged]
    assert co.same_partition(np.array([m[2] for m in merged]), np.array([m[2] for m in ref_merged]))


def test_label_segments_rejects_hdbscan():
    from speech_diarization_b200 import diar_diag
    with pytest.raises(NotImplementedError):
        diar_diag.label_segments(np.zeros((4, 192), np.float32), [(0, 1)] * 4, cluster="hdbscan")


def test_cluster_centers_and_dot_scores_match_numpy():
    from speech_diarization_b200 import postproc
    X, lab = synth_emb(3000, 5, 0.1, 9)
    x64 = torch.from_numpy(X.astype(np.float64)).cuda()
    cent = postproc.cluster_centers_device(x64, torch.from_numpy(lab.astype(np.int32)).cuda(), 5)
    ref = np.stack([X.astype(np.float64)[lab == k].mean(0) for k in range(5)])
    ref /= np.linalg.norm(ref, axis=1, keepdims=True) + 1e-9
    assert np.abs(cent.cpu().numpy() - ref.astype(np.float32)).max() <= 1e-7
    sc = postproc.dot_scores_device(torch.from_numpy(X).cuda(), cent)
    assert np.abs(sc.cpu().numpy() - X @ ref.astype(np.float32).T).max() <= 1e-5


def test_speaker_encoder_embed_matches_oracle(oracle_model):
    """SpeakerEncoder(...).embed(y, sr) (diar_diag.py:161-170): one utterance -> float32 [192]."""
    from oracle import ecapa_oracle
    from speech_diarization_b200 import diar_diag, speech_encode
    from conftest import synth_wave
    speech_encode.register_ecapa_state_dict(oracle_model.state_dict())
    try:
        enc = diar_diag.SpeakerEncoder("speechbrain-ecapa", device="cuda")
        y = synth_wave(1, 11000, 3)[0]
        e = enc.embed(y, 16000)
        assert e.dtype == np.float32 and e.shape == (192,)
        with torch.inference_mode():
            ref = ecapa_oracle.encode_batch(oracle_model, torch.from_numpy(y)[None]).squeeze().numpy()
        cos = float(np.dot(e, ref) / (np.linalg.norm(e) * np.linalg.norm(ref)))
        assert 1 - cos < 1e-4, 1 - cos
        with pytest.raises(NotImplementedError):
            enc.embed(y, 8000)
        with pytest.raises(ValueError):
            diar_diag.SpeakerEncoder("kaldi")
        np.testing.assert_array_equal(diar_diag.pad_with_context(np.arange(100), 10, 2.0, 5.0, 0.5), np.arange(15, 55))
    finally:
        speech_encode.register_ecapa_state_dict(None)
