import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def synth_wave(B, n, seed):
    """Deterministic speech-like test audio: harmonic stacks + noise, f32 in [-1, 1]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    out = np.zeros((B, n), np.float32)
    for b in range(B):
        f0 = 100.0 + 17.0 * (b % 11)
        sig = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) / h for h in range(1, 15))
        out[b] = (0.08 * sig + 0.004 * rng.standard_normal(n)).astype(np.float32)
    return out


def synth_emb(N, K, sigma, seed, D=192):
    """K unit centroids + sigma * N(0, I) (SURVEY.md §8d clustering inputs)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((K, D))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    lab = rng.integers(0, K, N)
    return (c[lab] + sigma * rng.standard_normal((N, D))).astype(np.float32), lab


@pytest.fixture(scope="session")
def oracle_model():
    from oracle import ecapa_oracle
    return ecapa_oracle.make_random_ecapa(0)


@pytest.fixture(scope="session")
def encoder(oracle_model):
    from speech_diarization_b200 import speech_encode
    enc = speech_encode.EcapaEncoderB200(oracle_model.state_dict(), device="cuda:0", max_batch=64, max_samples=24000)
    yield enc
    enc.close()
