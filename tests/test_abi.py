"""The C-ABI library loads on a box without a GPU and exports exactly what include/sd_b200.h
declares; the ctypes signature table covers the same set."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT
from speech_diarization_b200 import _lib

HEADER = os.path.join(ROOT, "include", "sd_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src))


def test_header_matches_ctypes_table():
    assert header_symbols() == set(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sd_[a-z0-9_]+)$", out, flags=re.M))
    assert header_symbols() <= exported


def test_library_loads_without_gpu_and_reports_version():
    lib = _lib.load()
    assert lib.sd_version() >= 100
    assert lib.sd_status_string(0) == b"ok"
    assert lib.sd_fbank_num_frames(24000) == 151
    assert lib.sd_fbank_num_frames(16000) == 101
    # 5.428 GFLOP per 1.5 s window (ASP decomposed)
    assert abs(lib.sd_ecapa_flops_per_window(151) / 1e9 - 5.428) < 0.01
    assert lib.sd_affinity_workspace_bytes(1000, 192) >= 1000 * 384 * 2
    assert lib.sd_ahc_workspace_bytes(1000) >= 8 * 1000 * 1000


def test_no_cpu_fallback_in_product_path():
    """The product package must not import the oracle (or sklearn/scipy/torchaudio arithmetic)."""
    pkg = os.path.join(ROOT, "speech_diarization_b200")
    banned = re.compile(r"^\s*(?:import|from)\s+(oracle|sklearn|scipy|torchaudio|librosa)\b", re.M)
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            hit = banned.search(open(os.path.join(pkg, f)).read())
            assert hit is None, f"{f} imports {hit.group(1)}"
