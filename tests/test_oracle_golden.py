"""CPU tests: the oracle is pinned against golden vectors produced by the REFERENCE's own
functions (tests/golden/make_golden.py), and the host-side (numpy) logic of the product
package reproduces the reference's window / segment / RTTM conventions bit-exactly."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_emb
from oracle import cluster_oracle as co
from oracle import ecapa_oracle as eo
from oracle import fbank_oracle as fo


# ------------------------------------------------------------------ oracle vs reference goldens
@pytest.mark.parametrize("tag", ["short", "win15"])
def test_fbank_oracle_matches_reference_fbank_batch(tag):
    g = golden(f"fbank_ref_{tag}.npz")
    np.testing.assert_array_equal(fo.fbank_batch(g["wav"], mean_nor=True), g["cmn"])
    np.testing.assert_array_equal(fo.fbank_batch(g["wav"], mean_nor=False), g["raw"])
    assert g["cmn"].shape == (int(g["B"]), 1 + int(g["n"]) // 160, 80)


@pytest.mark.parametrize("tag", ["clean", "noisy", "tiny"])
def test_cluster_oracle_matches_reference_cluster_embeddings(tag):
    g = golden("cluster_ref.npz")
    X = g[f"{tag}_X"]
    np.testing.assert_array_equal(co.cluster_embeddings(X, "agglo", 0.68), g[f"{tag}_labels"])
    np.testing.assert_array_equal(co.cluster_embeddings(X, "agglo", 0.5), g[f"{tag}_labels_thr05"])
    if tag == "clean":
        assert co.same_partition(g["clean_labels"], g["clean_true"])


def test_cluster_oracle_rejects_unknown_method():
    with pytest.raises(ValueError):
        co.cluster_embeddings(np.zeros((4, 192), np.float32), "kmeans")


def test_window_helpers_match_reference():
    g = golden("windows_ref.npz")
    y = np.arange(5000, dtype=np.float32)
    np.testing.assert_array_equal(co.frame_audio(y, 16000, 30.0, 10.0), g["frames_30_10"])
    np.testing.assert_array_equal(co.frame_audio(y, 16000, 100.0, 37.5), g["frames_100_37"])
    mask = [co.Segment(a, b) for a, b in g["mask"]]
    ws, vi = co.get_speech_windows(int(g["ylen"]), 16000, mask, 16000, 1600)
    np.testing.assert_array_equal(ws, g["window_starts"])
    np.testing.assert_array_equal(vi, g["valid_indices"])
    segs = co.labels_to_segments(ws, vi, g["window_labels"], 16000, int(g["ylen"]) / 16000)
    np.testing.assert_array_equal(np.array([[s.start, s.end, s.spk] for s in segs]), g["segs"])
    merged = co.merge_adjacent(segs, 0.05)
    np.testing.assert_array_equal(np.array([[s.start, s.end, s.spk] for s in merged]), g["merged"])


# ------------------------------------------------------------------------- ECAPA oracle (unpinned)
def test_ecapa_oracle_topology():
    m = eo.ECAPA_TDNN()
    assert eo.count_params(m) == 20_767_552
    keys = set(m.state_dict())
    for k in ["blocks.0.conv.conv.weight", "blocks.1.tdnn1.norm.norm.running_var",
              "blocks.2.res2net_block.blocks.6.conv.conv.bias", "blocks.3.se_block.conv2.conv.weight",
              "mfa.conv.conv.weight", "asp.tdnn.conv.conv.weight", "asp.conv.conv.bias",
              "asp_bn.norm.running_mean", "fc.conv.weight"]:
        assert k in keys, k
    assert m.state_dict()["asp.tdnn.conv.conv.weight"].shape == (128, 9216, 1)
    with torch.inference_mode():
        out = m.eval()(eo.synth_features(2, 37, 0))
    assert out.shape == (2, 1, 192)


def test_speechbrain_fbank_oracle_shape_and_norm():
    w = torch.randn(3, 24000) * 0.1
    f = eo.fbank_speechbrain(w)
    assert f.shape == (3, 151, 80)
    assert float(f.mean(dim=1).abs().max()) < 1e-4           # sentence mean normalisation
    raw = eo.fbank_speechbrain(w, mean_norm=False)
    assert float((raw.amax(dim=(1, 2)) - raw.amin(dim=(1, 2))).max()) <= 80.0 + 1e-3   # top_db


# ----------------------------------------------------------- product host logic (numpy, no GPU)
def test_product_frame_audio_is_view_and_matches_reference():
    from speech_diarization_b200 import vad, diar_diag
    g = golden("windows_ref.npz")
    y = np.arange(5000, dtype=np.float32)
    fr = vad.frame_audio(y, 16000, 30.0, 10.0)
    np.testing.assert_array_equal(fr, g["frames_30_10"])
    assert fr.base is not None and fr.strides == (160 * 4, 4)      # no copy
    np.testing.assert_array_equal(vad.frame_audio(y, 16000, 100.0, 37.5), g["frames_100_37"])
    fr2, hop = diar_diag.frame_audio(y, 16000, 30.0, 10.0)
    np.testing.assert_array_equal(fr2, g["frames_30_10"])
    assert hop == 160
    short, _ = diar_diag.frame_audio(np.ones(100, np.float32), 16000)       # pads to one window
    assert short.shape == (1, 480) and short[0, 100:].sum() == 0
    with pytest.raises(ValueError):
        vad.frame_audio(np.ones(100, np.float32), 16000)


def test_product_overlap_span_detection():
    from speech_diarization_b200 import vad, speech_encode
    y = np.random.default_rng(0).standard_normal(100000).astype(np.float32)
    fr = vad.frame_audio(y, 16000, 1500.0, 750.0)
    span, hop = speech_encode._overlap_span(fr)
    assert hop == 12000 and span.shape[0] == (fr.shape[0] - 1) * 12000 + 24000
    np.testing.assert_array_equal(span, y[: span.shape[0]])
    assert speech_encode._overlap_span(np.ascontiguousarray(fr)) is None


def test_product_speech_windows_match_reference():
    """_get_speech_windows (host index arithmetic) against the reference's own output; the run-length / merge kernels
    that consume it are checked on the GPU (tests/test_gpu_dense.py)."""
    from speech_diarization_b200 import anti_stick_diarize as asd
    g = golden("windows_ref.npz")
    mask = [asd.Segment(a, b) for a, b in g["mask"]]
    ws, vi = asd._get_speech_windows(np.zeros(int(g["ylen"]), np.float32), 16000, mask, 16000, 1600)
    np.testing.assert_array_equal(ws, g["window_starts"])
    np.testing.assert_array_equal(vi, g["valid_indices"])
    # overlapping, reversed and out-of-range mask segments behave like the reference's slice assignments
    odd = [asd.Segment(1.0, 2.0), asd.Segment(1.5, 2.5), asd.Segment(3.0, 2.0), asd.Segment(5.9, 99.0)]
    n_frames = int(np.ceil(int(g["ylen"]) / 16000 / 0.01))
    sm = np.zeros(n_frames, bool)
    for s_ in odd:
        sm[int(s_.start / 0.01):int(s_.end / 0.01)] = True
    ws2, vi2 = asd._get_speech_windows(np.zeros(int(g["ylen"]), np.float32), 16000, odd, 16000, 1600)
    cf = np.clip(((ws2 + 8000) / 16000 / 0.01).astype(int), 0, n_frames - 1)
    np.testing.assert_array_equal(vi2, np.where(sm[cf])[0])


def test_product_rttm_format(tmp_path):
    from speech_diarization_b200 import diarization_baseline as db, anti_stick_diarize as asd
    segs = [asd.Segment(1.5, 3.25, 1), asd.Segment(0.0, 1.5, 0)]
    tup = db.segments_to_tuples(segs)
    assert tup == [(0.0, 1.5, "SPEAKER_00"), (1.5, 3.25, "SPEAKER_01")]
    p = tmp_path / "clip.rttm"
    db.write_rttm(tup, p)
    assert p.read_text().splitlines() == [
        "SPEAKER clip 1 0.000 1.500 <NA> <NA> SPEAKER_00 <NA> <NA>",
        "SPEAKER clip 1 1.500 1.750 <NA> <NA> SPEAKER_01 <NA> <NA>"]
    assert db.rttm_lines(tup, "clip") == co.rttm_lines(tup, "clip")


def test_product_cluster_embeddings_error_behaviour():
    from speech_diarization_b200 import diar_diag
    with pytest.raises(ValueError):
        diar_diag.cluster_embeddings(np.zeros((4, 192), np.float32), method="kmeans")
    with pytest.raises(NotImplementedError):
        diar_diag.cluster_embeddings(np.zeros((4, 192), np.float32))       # default "hdbscan": out of scope


def test_product_fails_loudly_without_gpu():
    from speech_diarization_b200 import _lib, speech_encode
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SdError):
        speech_encode.fbank_batch(np.zeros((1, 4000), np.float32))


def test_product_fcluster_distance_matches_scipy_with_inversions():
    """Host-side flat cut of a centroid linkage (speech_diarization_b200.diarization_baseline.fcluster_distance)
    against scipy.cluster.hierarchy.fcluster(criterion="distance") on dendrograms WITH inversions."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from speech_diarization_b200.diarization_baseline import fcluster_distance
    rng = np.random.default_rng(0)
    for n in (2, 5, 50, 400):
        X = rng.standard_normal((n, 8))
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Z = linkage(X, "centroid", "euclidean")
        for t in (0.0, 0.3, 0.7, 1.0, 1.2, 5.0):
            assert co.same_partition(fcluster(Z, t, "distance") - 1, fcluster_distance(Z, t)), (n, t)
    assert (np.diff(Z[:, 2]) < 0).any()


# ------------------------------------------------------------------ post-processing oracle
def test_post_oracle_matches_reference_asnorm_and_viterbi():
    from oracle import post_oracle as po
    g = golden("post_ref.npz")
    X, cent = g["as_X"], g["as_cent"]
    np.testing.assert_array_equal(po.asnorm_scores(X, cent, X, topk=min(200, len(X))), g["as_self"])
    np.testing.assert_array_equal(po.asnorm_scores(X[:70], cent, g["as_cohort"], topk=200), g["as_small"])
    for tag, alpha in (("as", 0.995), ("f64", 0.9), ("ties", 0.5), ("k2", 0.995), ("t1", 0.995)):
        path = po.viterbi_hmm(g[f"vt_scores_{tag}"], alpha=alpha)
        assert path.dtype == np.int32
        np.testing.assert_array_equal(path, g[f"vt_path_{tag}"])
    assert len(set(g["vt_path_ties"].tolist())) > 1 and len(set(g["vt_path_f64"].tolist())) > 1


@pytest.mark.parametrize("tag", ["f32", "f64", "short"])
def test_post_oracle_matches_reference_vad_mask_ops(tag):
    from oracle import post_oracle as po
    g = golden("post_ref.npz")
    p, hop = g[f"vad_{tag}_probs"], float(g[f"vad_{tag}_hop"])
    m0 = po.hysteresis_binarize(p, 0.6, 0.4)
    np.testing.assert_array_equal(m0, g[f"vad_{tag}_hyst"])
    np.testing.assert_array_equal(po.morph_open_close(m0, hop, 80.0, 40.0), g[f"vad_{tag}_morph"])
    np.testing.assert_array_equal(po.morph_open_close(m0, hop, 50.0, 70.0), g[f"vad_{tag}_morph_b"])
    np.testing.assert_array_equal(po.morph_open_close(m0, hop, 0.0, 100.0), g[f"vad_{tag}_morph_c"])
    segs = np.array(po.mask_to_segments(g[f"vad_{tag}_morph"], hop), dtype=np.float64).reshape(-1, 2)
    np.testing.assert_array_equal(segs, g[f"vad_{tag}_segs"])
    raw = np.array(po.mask_to_segments(m0, hop, 60.0, 45.0, 30.0), dtype=np.float64).reshape(-1, 2)
    np.testing.assert_array_equal(raw, g[f"vad_{tag}_segs_raw"])
    if tag == "f32":
        np.testing.assert_array_equal(po.hysteresis_binarize(p, 0.3, 0.7), g["vad_toggle_hyst"])
        assert len(segs) > 5 and m0.any() and not m0.all()
        assert po.mask_to_segments(np.zeros(100, bool), 10.0) == []
        np.testing.assert_array_equal(np.array(po.mask_to_segments(np.ones(100, bool), 10.0)), g["vad_full_segs"])


def test_post_oracle_matches_reference_whiten():
    from oracle import post_oracle as po
    g = golden("post_ref.npz")
    np.testing.assert_array_equal(po.whiten_l2(g["wh_X"]), g["wh_out"])
    np.testing.assert_array_equal(po.whiten_l2(g["wh_X_small"]), g["wh_out_small"])
    assert g["wh_out"].dtype == np.float64
    np.testing.assert_allclose(np.linalg.norm(g["wh_out"], axis=1), 1.0, atol=1e-8)


@pytest.mark.parametrize("tag,wh,asn,vbx,thr", [("full", 1, 1, 1, 0.1), ("plain", 0, 0, 0, 0.68),
                                                 ("wh_argmax", 1, 1, 0, 0.1), ("vbx_raw", 0, 0, 1, 0.68)])
def test_post_oracle_pipeline_matches_reference_functions(tag, wh, asn, vbx, thr):
    """oracle.label_segments (diar_diag.py:352-411) against the reference's own functions chained in main()'s order."""
    from oracle import post_oracle as po
    g = golden("post_ref.npz")
    segs = [tuple(x) for x in g["pipe_segs"]]
    merged, final, labels = po.label_segments(g["pipe_X"], segs, whiten=wh, asnorm=asn, cos_thr=thr, use_vbx=vbx)
    np.testing.assert_array_equal(labels, g[f"pipe_{tag}_labels"])
    np.testing.assert_array_equal(final, g[f"pipe_{tag}_final"])
    assert co.same_partition(labels, g["pipe_true"])
    assert merged[0][0] == segs[0][0] and merged[-1][1] == segs[-1][1] and len(merged) >= 20


# ------------------------------------------------------------------ ECAPA-TDNN trunk vs an installed third party
def _neutral_bn(model):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.fill_(1); m.bias.data.zero_(); m.running_mean.zero_(); m.running_var.fill_(1.0 - m.eps)
    return model


def test_ecapa_oracle_trunk_matches_transformers_ecapa_golden():
    """tests/golden/ecapa_hf_ref.npz = transformers' ECAPA_TimeDelayNet (speechbrain's ECAPA_TDNN without the
    BatchNorms) run with the oracle's seed-0 conv weights (make_ecapa_hf_golden.py).  The oracle with identity
    BatchNorms reproduces it bit for bit: its conv / Res2Net / SE / ASP / fc wiring is pinned to third-party code."""
    g = golden("ecapa_hf_ref.npz")
    model = _neutral_bn(eo.make_random_ecapa(0))
    with torch.inference_mode():
        out = model(torch.from_numpy(g["feats"])).squeeze(1).numpy()
    np.testing.assert_array_equal(out, g["emb"])
    assert out.shape == (3, 192) and np.abs(out).max() > 0.05
