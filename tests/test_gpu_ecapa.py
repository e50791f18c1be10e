"""ECAPA-TDNN trunk + end-to-end embedding parity against the CPU fp32 oracle on identical
synthetic audio and identical random-init weights.

Gate (BASELINE.json north_star): cosine similarity of embeddings >= 1 - 1e-4.
Layer-wise bound: relative L2 error <= 5e-3 (f16 operands / f16 activations, f32 accumulation;
measured 4e-4 after block0 growing to 1.4e-3 after the MFA layer)."""
import numpy as np
import pytest
import torch

from conftest import synth_wave
from oracle import ecapa_oracle as eo
from speech_diarization_b200 import _lib, speech_encode as se

pytestmark = pytest.mark.gpu
COS_TOL = 1e-4


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("B,T", [(3, 151), (5, 101), (2, 26), (1, 10), (7, 248),
                                 (3, 249), (2, 301), (2, 513), (1, 1001)])   # > 248 frames: chunked online-softmax pooling
def test_trunk_layerwise_and_embedding_parity(oracle_model, encoder, B, T):
    x = eo.synth_features(B, T, seed=B * 100 + T)
    taps = {}
    with torch.inference_mode():
        ref = oracle_model(x, taps).squeeze(1)
    got = encoder.forward_feats(x).cpu()
    for name in ("block0", "b1.out", "b2.out", "b3.out", "mfa"):
        assert _rel(encoder.debug_fetch(name, B, T), taps[name].transpose(1, 2)) < 5e-3, name
    assert _rel(encoder.debug_fetch("pooled", B, T), taps["pooled"]) < 5e-3
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert float((1 - cos).max()) < COS_TOL
    assert _rel(got, ref) < 5e-3


@pytest.mark.parametrize("B,T", [(300, 151), (700, 151), (5, 101), (3, 160), (2, 16), (9, 129), (4, 128), (2, 161), (3, 131),
                                 (3, 137), (149, 152), (597, 40)])
def test_fused_res2net_is_bit_identical_to_the_per_conv_chain(oracle_model, B, T, monkeypatch):
    """The fused Res2Net kernels keep the operand order and f16 rounding points of the per-convolution chain, so v
    after every block and the embeddings must be bit-identical, for the four-window pipeline (res2net_pipe_kernel,
    T + 2 dil <= 160) and for the first fused kernel (res2net_fused_kernel: SD_R2_PIPE=0, and T = 153 .. 160).
    B = 700 > 4 x 148 exercises a second round of window slots, B = 149 / 597 a last round with one active slot,
    T = 161 falls back to the chain, T = 129 .. 152 run the transposed second tile (T = 131: the mirrored right-halo
    frames straddle the two tiles)."""
    x = eo.synth_features(B, T, seed=7 * B + T)
    out = {}
    for tag, fused, pipe in (("pipe", "1", "1"), ("v1", "1", "0"), ("chain", "0", "0")):
        if tag == "v1" and B > 300:
            continue
        monkeypatch.setenv("SD_ECAPA_R2FUSED", fused)
        monkeypatch.setenv("SD_R2_PIPE", pipe)
        monkeypatch.setenv("SD_ECAPA_GRAPH", "0")
        enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=B, max_samples=(T - 1) * 160)
        try:
            emb = enc.forward_feats(x).cpu()
            out[tag] = (emb, enc.debug_fetch("b3.res2net", B, T).cpu(), enc.debug_fetch("b3.out", B, T).cpu())
        finally:
            enc.close()
    for tag in out:
        for a, b in zip(out[tag], out["chain"]):
            assert torch.equal(a, b), tag
    with torch.inference_mode():
        ref = oracle_model(x[:64]).squeeze(1)
    cos = torch.nn.functional.cosine_similarity(out["pipe"][0][:64], ref, dim=1)
    assert float((1 - cos).max()) < COS_TOL


@pytest.mark.parametrize("B,T", [(300, 151), (7, 113), (5, 120), (3, 248), (2, 301), (3, 1001), (4, 101)])
def test_time_statistics_from_the_gemm_writeout_match_the_separate_passes(oracle_model, B, T, monkeypatch):
    """The SE squeeze mean and the ASP mean/std are accumulated per group of gcd(128, Tp) rows inside the tdnn2 / MFA
    write-outs (EpiParams::colsum) instead of by extra passes over the activations.  Same f16 values, different
    f32 summation order and a different variance shift (BN shift instead of the first frame): the SE gate agrees
    to 1e-5, mean|std to 2e-4 relative, embeddings to 1 - cos < 1e-6.  Tp = 112 (T = 101) has 16-row groups,
    Tp = 128 one group per tile, Tp = 160 32-row groups, Tp = 256 / 1024 whole-tile groups."""
    x = eo.synth_features(B, T, seed=11 * B + T)
    out = {}
    for on in ("1", "0"):
        monkeypatch.setenv("SD_ECAPA_COLSUM", on)
        monkeypatch.setenv("SD_ECAPA_GRAPH", "0")
        enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=B, max_samples=(T - 1) * 160)
        try:
            emb = enc.forward_feats(x).cpu()
            out[on] = (emb, enc.debug_fetch("b3.se", B, T).cpu(), enc.debug_fetch("asp.stats", B, T).cpu())
        finally:
            enc.close()
    assert _rel(out["1"][1], out["0"][1]) < 1e-5          # SE gate
    assert _rel(out["1"][2], out["0"][2]) < 2e-4          # ASP mean | std
    cos = torch.nn.functional.cosine_similarity(out["1"][0], out["0"][0], dim=1)
    assert float((1 - cos).max()) < 1e-6
    with torch.inference_mode():
        ref = oracle_model(x).squeeze(1)
    cos = torch.nn.functional.cosine_similarity(out["1"][0], ref, dim=1)
    assert float((1 - cos).max()) < COS_TOL


@pytest.mark.parametrize("n", [24000, 16000, 4000, 8123, 48000, 160000])      # up to 10 s (pyannote chunk length)
def test_encode_batch_matches_oracle(oracle_model, encoder, n):
    w = synth_wave(6, n, n)
    w[2, n // 2:] = 0.0           # zero-padded member of a variable-length batch
    with torch.inference_mode():
        ref = eo.encode_batch(oracle_model, torch.from_numpy(w)).squeeze(1)
    out = encoder.encode_batch(torch.from_numpy(w))
    assert out.shape == (6, 1, 192) and out.dtype == torch.float32 and out.is_cuda
    cos = torch.nn.functional.cosine_similarity(out.squeeze(1).cpu(), ref, dim=1)
    assert float((1 - cos).max()) < COS_TOL


def test_reference_callables(oracle_model):
    """ecapa_encode_batch / using_ecapa_encoder / ECAPAEncoder keep the reference's contracts
    (speech_encode.py:64-78, ecapa_annote.py:6-22)."""
    from speech_diarization_b200 import ecapa_annote, vad
    se.register_ecapa_state_dict(oracle_model.state_dict())
    try:
        enc = se.using_ecapa_encoder()
        assert enc is se.using_ecapa_encoder()                      # lru_cache singleton
        y = synth_wave(1, 16000 * 6, 3)[0]
        frames = vad.frame_audio(y, 16000, 1500.0, 750.0)           # strided view: uploaded once
        e_view = se.ecapa_encode_batch(frames)
        e_copy = se.ecapa_encode_batch(np.ascontiguousarray(frames))
        assert e_view.shape == (frames.shape[0], 192) and e_view.dtype == np.float32
        np.testing.assert_array_equal(e_view, e_copy)
        with torch.inference_mode():
            ref = eo.encode_batch(oracle_model, torch.from_numpy(np.ascontiguousarray(frames))).squeeze(1)
        cos = torch.nn.functional.cosine_similarity(torch.from_numpy(e_view), ref, dim=1)
        assert float((1 - cos).max()) < COS_TOL
        # float64 input is cast like the reference's .float() (speech_encode.py:76)
        e64 = se.ecapa_encode_batch(np.ascontiguousarray(frames).astype(np.float64))
        np.testing.assert_array_equal(e64, e_copy)
        m = ecapa_annote.ECAPAEncoder(device="cuda")
        assert m.dimension == 192
        out = m(torch.from_numpy(np.ascontiguousarray(frames[:3])).cuda())
        assert out.shape == (3, 192) and out.is_cuda
        np.testing.assert_allclose(out.cpu().numpy(), e_copy[:3], rtol=0, atol=0)
    finally:
        se.register_ecapa_state_dict(None)


def test_no_weights_is_an_error(monkeypatch):
    monkeypatch.delenv("SD_ECAPA_CKPT", raising=False)
    se.register_ecapa_state_dict(None)
    with pytest.raises(_lib.SdError):
        se.using_ecapa_encoder()


def test_missing_tensor_is_reported(oracle_model):
    sd = dict(oracle_model.state_dict())
    del sd["mfa.conv.conv.weight"]
    with pytest.raises(_lib.SdError, match="mfa.conv.conv.weight"):
        se.EcapaEncoderB200(sd, device="cuda:0", max_batch=1, max_samples=4000)


def test_batch_properties_full_size(oracle_model, encoder):
    """BASELINE batch: determinism, batch-composition independence and chunking across the
    plan's capacity (512 windows through a 64-window plan)."""
    w = torch.from_numpy(synth_wave(16, 24000, 11)).cuda().repeat(32, 1)       # 512 windows
    e1 = encoder.embed_device(w, 24000, 512, 24000)
    e2 = encoder.embed_device(w, 24000, 512, 24000)
    assert torch.equal(e1, e2)                              # run-to-run: bit-identical
    # same window, different batch slot / batch size: bit-identical.  The SE and ASP time statistics are summed
    # inside the GEMM write-outs per group of gcd(128, Tp) rows — window-relative groups added in a fixed order —
    # so nothing in the forward depends on where a window's rows fall relative to the 128-row tiles
    for k in range(1, 32):
        assert torch.equal(e1[:16], e1[16 * k:16 * (k + 1)]), k
    single = encoder.embed_device(w[3:4].contiguous(), 24000, 1, 24000)
    assert torch.equal(single[0], e1[3])
    n = encoder.embed_device(w, 24000, 512, 24000, l2_normalize=True)
    assert torch.allclose(n.norm(dim=1), torch.ones(512, device="cuda"), atol=1e-5)
    assert float((n - e1 / (e1.norm(dim=1, keepdim=True) + 1e-8)).abs().max()) < 1e-6


@pytest.mark.parametrize("colsum", ["1", "0"])
@pytest.mark.parametrize("n", [24000, 16000, 39520, 4000])      # Tp = 160 / 112 / 256 / 48: 32-, 16-, 128-, 16-row groups
def test_slot_independence_is_bit_exact(oracle_model, monkeypatch, colsum, n):
    """embed(x)[i] does not depend on the batch position of window i or on the batch size, with the fused
    statistics (default) and with the separate passes."""
    monkeypatch.setenv("SD_ECAPA_COLSUM", colsum)
    enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=70, max_samples=n)
    try:
        w = torch.from_numpy(synth_wave(7, n, 11)).cuda().repeat(10, 1)
        e1 = enc.embed_device(w, n, 70, n)
        for k in range(1, 10):
            assert torch.equal(e1[:7], e1[7 * k:7 * (k + 1)]), k
        single = enc.embed_device(w[3:4].contiguous(), n, 1, n)
        assert torch.equal(single[0], e1[3])
        five = enc.embed_device(w[2:7].contiguous(), n, 5, n)
        assert torch.equal(five, e1[2:7])
    finally:
        enc.close()


def _scaled_bn_model(oracle_model, key, factor):
    sd = {k: v.clone() for k, v in oracle_model.state_dict().items()}
    sd[key] = sd[key] * factor
    m = eo.ECAPA_TDNN().eval()
    m.load_state_dict(sd)
    return m


def test_large_batchnorm_scale_stays_finite_and_within_the_gate(oracle_model):
    """A checkpoint with one BatchNorm scale 200x the usual: activations of O(1000) from there on.  They are
    stored as f16 (max 65504): everything must stay finite, within the cosine gate, and the overflow flag clear."""
    m = _scaled_bn_model(oracle_model, "blocks.1.tdnn1.norm.norm.weight", 200.0)
    w = synth_wave(6, 24000, 21)
    with torch.inference_mode():
        ref = eo.encode_batch(m, torch.from_numpy(w)).squeeze(1)
    enc = se.EcapaEncoderB200(m.state_dict(), device="cuda:0", max_batch=6, max_samples=24000)
    try:
        got = enc.encode_batch(torch.from_numpy(w)).squeeze(1).cpu()
        assert not enc.overflowed()
    finally:
        enc.close()
    assert torch.isfinite(got).all()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert float((1 - cos).max()) < COS_TOL


def test_f16_overflow_is_loud_not_silent(oracle_model):
    """A BatchNorm scale of 1e6 drives activations past the f16 range.  Stores saturate (no inf - inf = NaN inside
    the trunk), the overflow flag is raised, the device path delivers NaN embeddings and the host path raises."""
    m = _scaled_bn_model(oracle_model, "blocks.2.tdnn2.norm.norm.weight", 1e6)
    w = synth_wave(64, 24000, 22)
    enc = se.EcapaEncoderB200(m.state_dict(), device="cuda:0", max_batch=64, max_samples=24000)
    try:
        got = enc.encode_batch(torch.from_numpy(w)).squeeze(1).cpu()
        assert torch.isnan(got).all()
        assert enc.overflowed() and not enc.overflowed()        # reported once, then reset
        with pytest.raises(_lib.SdError, match="f16 range"):
            enc.embed_host(w, 24000, 64, 24000)
        # the next forward with sane inputs through a sane plan is unaffected: the flag is per forward
    finally:
        enc.close()
    enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=64, max_samples=24000)
    try:
        assert torch.isfinite(enc.encode_batch(torch.from_numpy(w))).all() and not enc.overflowed()
    finally:
        enc.close()


def test_pageable_and_pinned_host_paths_agree(oracle_model):
    """sd_ecapa_embed_host: page-locked memory is read by the copy engine directly, pageable memory goes through
    the pinned staging ring (host_stage.cuh); both give the embeddings of the device path bit for bit."""
    n_win, hop, win = 300, 12000, 24000
    y = synth_wave(1, (n_win - 1) * hop + win, 31)[0]
    enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=n_win, max_samples=win)
    try:
        dev = enc.embed_device(torch.from_numpy(y).cuda(), hop, n_win, win).cpu().numpy()
        pageable = enc.embed_host(y, hop, n_win, win)
        pinned_t = torch.from_numpy(y).pin_memory()
        pinned = enc.embed_host(pinned_t.numpy(), hop, n_win, win)
        batch = np.ascontiguousarray(np.lib.stride_tricks.as_strided(y, (n_win, win), (4 * hop, 4)))
        stacked = enc.embed_host(batch.reshape(-1), win, n_win, win)
    finally:
        enc.close()
    np.testing.assert_array_equal(pageable, dev)
    np.testing.assert_array_equal(pinned, dev)
    np.testing.assert_array_equal(stacked, dev)


def test_unsupported_and_bad_arguments(encoder):
    with pytest.raises(_lib.SdError):
        encoder.encode_batch(torch.zeros(2, 24000), wav_lens=torch.tensor([1.0, 0.5]))
    with pytest.raises(ValueError):
        encoder.embed_device(torch.zeros(100, device="cuda"), 100, 1, 100)
    assert encoder.embed_device(torch.zeros(0, device="cuda"), 1, 0, 24000).shape == (0, 192)


def test_trunk_matches_transformers_ecapa_golden():
    """CUDA trunk against the third-party golden (transformers' ECAPA_TimeDelayNet, tests/golden/ecapa_hf_ref.npz):
    the oracle's seed-0 conv weights with every BatchNorm an exact identity.  Gate: 1 - cos <= 1e-4."""
    from conftest import golden
    from oracle import ecapa_oracle
    from speech_diarization_b200 import speech_encode
    g = golden("ecapa_hf_ref.npz")
    model = ecapa_oracle.make_random_ecapa(0)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.fill_(1); m.bias.data.zero_(); m.running_mean.zero_(); m.running_var.fill_(1.0 - m.eps)
    enc = speech_encode.EcapaEncoderB200(model.state_dict(), device="cuda:0", max_batch=4, max_samples=16000)
    try:
        got = enc.forward_feats(torch.from_numpy(g["feats"])).cpu()
    finally:
        enc.close()
    ref = torch.from_numpy(g["emb"])
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert float((1 - cos).max()) <= 1e-4, float((1 - cos).max())
    assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
