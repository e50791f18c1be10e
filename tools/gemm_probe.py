"""GPU bring-up probe for the tcgen05 GEMM kernel (run under gpurun).
Prints error statistics per configuration; exits non-zero on mismatch."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_diarization_b200 import _lib

lib = ctypes.CDLL(_lib.LIB_PATH)
lib.sd_debug_gemm_f16.restype = ctypes.c_int
lib.sd_debug_gemm_f16.argtypes = _lib.SIGNATURES["sd_debug_gemm_f16"][1]
lib.sd_last_error.restype = ctypes.c_char_p
dev = torch.device("cuda:0")
torch.manual_seed(0)
bad = 0

def run(M, N, K, taps=1, dil=1, n_tile=128, reps=1):
    global bad
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    B = (torch.randn(N, taps * K, device=dev) * 0.5).half()
    D = torch.full((M, N), float("nan"), device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        st = lib.sd_debug_gemm_f16(A.data_ptr(), M, K, B.data_ptr(), N, taps, dil, n_tile, D.data_ptr(), _lib.stream_ptr())
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print(f"M={M} N={N} K={K} taps={taps} dil={dil} n_tile={n_tile}: CUDA ERROR {e}")
        sys.exit(3)
    dt = (time.time() - t0) / reps
    if st != 0:
        print("status", st, lib.sd_last_error().decode()); bad += 1; return
    Af = A.float(); Bf = B.float()
    ref = torch.zeros(M, N, device=dev)
    for j in range(taps):
        off = (j - taps // 2) * dil
        As = torch.zeros_like(Af)
        lo, hi = max(0, -off), min(M, M - off)
        As[lo:hi] = Af[lo + off:hi + off]
        ref += As @ Bf[:, j * K:(j + 1) * K].T
    err = (D - ref).abs()
    nan = torch.isnan(D).sum().item()
    mx = err[~torch.isnan(err)].max().item() if nan < D.numel() else float("nan")
    ok = nan == 0 and mx < 2e-2 * (K * taps / 64) ** 0.5
    if not ok:
        bad += 1
        # structure of the error: which rows / cols are wrong
        wrong = (err > 1e-1) | torch.isnan(D)
        rows = wrong.any(1).nonzero().flatten()[:16].tolist()
        cols = wrong.any(0).nonzero().flatten()[:16].tolist()
        print("   wrong rows", rows, "cols", cols, "frac", wrong.float().mean().item())
        print("   D[0,:8]", D[0, :8].tolist(), "\n   R[0,:8]", ref[0, :8].tolist())
    tf = 2.0 * M * N * K * taps / dt / 1e12
    print(f"M={M} N={N} K={K} taps={taps} dil={dil} n_tile={n_tile}: max_err={mx:.3e} nan={nan} {'OK' if ok else 'FAIL'}  {dt*1e3:.3f} ms {tf:.1f} TF/s")

print("version", lib.sd_version(), torch.cuda.get_device_name(0))
run(128, 128, 64, n_tile=128)
run(128, 128, 128, n_tile=128)
run(256, 256, 256, n_tile=256)
run(300, 200, 192, n_tile=128)
run(300, 320, 128, n_tile=160)
run(1000, 128, 128, taps=3, dil=2, n_tile=128)
run(1000, 1024, 128, taps=5, dil=1, n_tile=256)
run(5000, 1024, 1024, n_tile=256)
run(81920, 3072, 3072, n_tile=256, reps=3)
run(81920, 1024, 1024, n_tile=256, reps=3)
run(81920, 128, 128, taps=3, dil=3, n_tile=128, reps=3)
sys.exit(1 if bad else 0)
