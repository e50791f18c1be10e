"""fbank kernels through the reference-facing callable ``fbank_batch`` and the C ABI.

Tolerances (f32 path, stated per north_star "within the tolerance written in the test"):
  torchaudio variant : |d log-mel| <= 1e-3   (FFT factorisation and the f64-built mel table differ
                       from torch's pocketfft / f32-built table by ~1e-5 relative in power; measured 2e-4)
  speechbrain variant: |d dB|      <= 3e-3   (x 10/ln10 = 4.34 on the same relative error; measured 7e-4)
"""
import numpy as np
import pytest
import torch

from conftest import golden, synth_wave
from oracle import ecapa_oracle as eo
from oracle import fbank_oracle as fo
from speech_diarization_b200 import _lib, speech_encode as se

pytestmark = pytest.mark.gpu
TOL_LOG, TOL_DB = 1e-3, 3e-3


@pytest.fixture(params=[0, 1], ids=["fft", "tensorcore"], autouse=True)
def fbank_kernel(request):
    """Every test of this file runs on both frames kernels: the FFT on the FP32 pipe (default) and the folded
    split-f16 DFT on the tensor cores (sd_fbank_kernel(1))."""
    lib = _lib.load()
    before = lib.sd_fbank_kernel(-1)
    lib.sd_fbank_kernel(request.param)
    yield request.param
    lib.sd_fbank_kernel(before)


@pytest.mark.parametrize("tag", ["short", "win15"])
def test_fbank_batch_matches_reference_golden(tag):
    g = golden(f"fbank_ref_{tag}.npz")
    got = se.fbank_batch(g["wav"], mean_nor=True)
    assert got.shape == g["cmn"].shape and got.dtype == np.float32
    assert np.abs(got - g["cmn"]).max() < TOL_LOG
    assert np.abs(se.fbank_batch(g["wav"], mean_nor=False) - g["raw"]).max() < TOL_LOG


@pytest.mark.parametrize("n", [400, 401, 559, 560, 4000, 16000, 24000, 48123])
def test_fbank_torchaudio_variant_vs_oracle(n):
    w = synth_wave(3, n, n)
    ref = fo.fbank_batch(w)
    got = se.fbank_batch(w)
    assert got.shape == ref.shape == (3, 1 + n // 160, 80)
    assert np.abs(got - ref).max() < TOL_LOG


@pytest.mark.parametrize("n", [400, 4000, 16000, 24000])
@pytest.mark.parametrize("mean_norm", [True, False])
def test_fbank_speechbrain_variant_vs_oracle(n, mean_norm):
    w = synth_wave(4, n, n + 1)
    w[1, n // 3:] = 0.0                      # zero padding inside the batch: hits the top_db floor
    ref = eo.fbank_speechbrain(torch.from_numpy(w), mean_norm=mean_norm).numpy()
    got = se.fbank_batch_device(torch.from_numpy(w).cuda(), variant=1, mean_nor=mean_norm).cpu().numpy()
    assert np.abs(got - ref).max() < TOL_DB


def test_fbank_in_place_windows_equal_materialised_windows():
    """Windows addressed in place in one audio buffer (stride = hop) == frame_audio copies."""
    from speech_diarization_b200 import vad
    y = synth_wave(1, 100000, 5)[0]
    fr = vad.frame_audio(y, 16000, 1500.0, 750.0)
    a = se.fbank_batch(np.ascontiguousarray(fr))
    b = se.fbank_batch_device(torch.from_numpy(y).cuda(), variant=0, wav_stride=12000, n_windows=fr.shape[0],
                              n_samples=24000).cpu().numpy()
    np.testing.assert_array_equal(a, b)


def test_fbank_properties_at_full_batch():
    """BASELINE batch (512 x 1.5 s): CMN output has zero time-mean; silence gives the log floor."""
    w = torch.from_numpy(synth_wave(8, 24000, 9)).cuda().repeat(64, 1)
    w[5] = 0.0
    out = se.fbank_batch_device(w, variant=0, mean_nor=True)
    assert out.shape == (512, 151, 80)
    assert float(out.mean(dim=1).abs().max()) < 2e-4
    raw = se.fbank_batch_device(w, variant=0, mean_nor=False)
    assert torch.allclose(raw[5], torch.full_like(raw[5], float(np.log(np.float32(1e-6)))), atol=1e-6)
    assert torch.equal(out[0], out[8])       # identical windows -> identical features (determinism)


def test_fbank_error_behaviour():
    with pytest.raises(AssertionError):
        se.fbank_batch(np.zeros(4000, np.float32))            # ndim != 2, as the reference asserts
    with pytest.raises(_lib.SdError):
        se.fbank_batch(np.zeros((2, 100), np.float32))        # shorter than one frame
    assert se.fbank_batch(np.zeros((0, 4000), np.float32)).shape == (0, 26, 80)


@pytest.mark.parametrize("scale", [1e-5, 1e-3, 30.0, 3e4])
@pytest.mark.parametrize("variant", [0, 1])
def test_fbank_input_scale_invariance(variant, scale):
    """Quiet and loud inputs (the tensor-core kernel rescales every span by a power of two, so neither the f16 lo part
    nor the f16 range is a limit): log-mel of s*x == log-mel of x + log(s^2) wherever no floor is active."""
    if variant == 0 and scale < 1e-2:
        pytest.skip("log(x + 1e-6): a 1e-6 floor leaves nothing to compare at this scale")
    w = synth_wave(2, 8000, 21)
    a = se.fbank_batch_device(torch.from_numpy(w).cuda(), variant=variant, mean_nor=False).double().cpu().numpy()
    b = se.fbank_batch_device(torch.from_numpy(w * np.float32(scale)).cuda(), variant=variant, mean_nor=False).double().cpu().numpy()
    assert np.isfinite(b).all()
    if variant == 1:
        shift = 20.0 * np.log10(scale)
        ok = (a > a.max() - 70.0) & (b > b.max() - 70.0) & (a > -90.0) & (b > -90.0)   # away from the top_db / amin floors
        assert ok.sum() >= 100
        assert np.abs((b - a - shift)[ok]).max() < TOL_DB
    else:
        pa, pb = np.exp(a) - 1e-6, np.exp(b) - 1e-6                                   # undo log(x + 1e-6)
        ok = (pa > 1e-1) & (pb > 1e-1)
        assert ok.sum() >= 100
        assert np.abs(np.log(pb[ok] / pa[ok]) - 2 * np.log(scale)).max() < TOL_LOG


def test_fbank_kernels_agree_on_embeddings(oracle_model):
    """The two frames kernels feed the same ECAPA trunk: embeddings agree to 1 - cos < 1e-6."""
    lib = _lib.load()
    before = lib.sd_fbank_kernel(-1)
    y = torch.from_numpy(synth_wave(1, 16000 * 8, 33)[0]).cuda()
    out = []
    try:
        for which in (0, 1):
            lib.sd_fbank_kernel(which)
            enc = se.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=16, max_samples=24000)
            try:
                out.append(enc.embed_device(y, 12000, 9, 24000).double())
            finally:
                enc.close()
    finally:
        lib.sd_fbank_kernel(before)
    cos = torch.nn.functional.cosine_similarity(out[0], out[1], dim=1)
    assert float((1 - cos).max()) < 1e-5
