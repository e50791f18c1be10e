"""ctypes loader for libsd_b200.so (the C-ABI declared in include/sd_b200.h).

There is deliberately no fallback: if the CUDA extension is missing or a call
fails, the caller gets an exception (BASELINE.json north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_long, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# SD_LIB_PATH: load another build of the same library (A/B timing of compile-time variants); never a fallback
LIB_PATH = os.environ.get("SD_LIB_PATH") or os.path.join(_HERE, "csrc", "libsd_b200.so")


class SdError(RuntimeError):
    """A libsd_b200 entry point returned a non-zero status."""


# name -> (restype, argtypes); must list every symbol include/sd_b200.h declares
# (tests/test_abi.py checks the two against each other).
SIGNATURES = {
    "sd_version": (c_int, []),
    "sd_status_string": (c_char_p, [c_int]),
    "sd_last_error": (c_char_p, []),
    "sd_launch_count": (c_long, []),
    "sd_fbank_num_frames": (c_int, [c_int]),
    "sd_fbank_f32": (c_int, [c_void_p, c_long, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_fbank_kernel": (c_int, [c_int]),
    "sd_ecapa_plan_create": (c_int, [POINTER(c_char_p), POINTER(c_void_p), POINTER(c_int64), c_int, c_int, c_int, POINTER(c_void_p)]),
    "sd_ecapa_plan_destroy": (c_int, [c_void_p]),
    "sd_ecapa_embed": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_ecapa_embed_offsets": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_ecapa_embed_host": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_ecapa_overflow": (c_int, [c_void_p, c_int, POINTER(c_int), c_void_p]),
    "sd_ecapa_forward_feats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_ecapa_debug_fetch": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int), c_void_p]),
    "sd_ecapa_profile": (c_int, [c_void_p, c_int]),
    "sd_ecapa_num_stages": (c_int, []),
    "sd_ecapa_profile_read": (c_int, [c_void_p, c_int, c_char_p, POINTER(c_float), POINTER(c_int)]),
    "sd_ecapa_flops_per_window": (c_double, [c_int]),
    "sd_l2norm_f32": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "sd_affinity_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sd_cosine_distance_rowblock": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sd_ahc_workspace_bytes": (c_size_t, [c_int]),
    "sd_ahc_average_f32": (c_int, [c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sd_ahc_read_stats": (c_int, [c_void_p, c_int, POINTER(c_int32), POINTER(c_int32)]),
    "sd_centroid_linkage_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sd_centroid_linkage_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sd_window_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sd_adjacent_cosine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sd_viterbi_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sd_viterbi_hmm": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "sd_asnorm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sd_asnorm_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sd_whiten_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sd_whiten_l2_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, POINTER(c_int32), c_void_p]),
    "sd_cluster_centers_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sd_dot_scores": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sd_hysteresis_u8": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_void_p, c_void_p]),
    "sd_morph_open_close_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sd_mask_segments_workspace_bytes": (c_size_t, [c_int]),
    "sd_mask_segments_i32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sd_scd_peaks": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "sd_speaker_centroids": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "sd_scatter_labels": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sd_label_runs": (c_int, [c_void_p, c_int, c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sd_merge_adjacent": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "sd_gather_pad_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sd_debug_gemm_f16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol. Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(speech_diarization_b200/csrc/build.sh). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.sd_status_string(status).decode()
        detail = lib.sd_last_error().decode()
        raise SdError(f"{what or 'libsd_b200'}: status {status} ({msg}) {detail}")


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream, so our launches are ordered with torch's."""
    import torch

    return torch.cuda.current_stream().cuda_stream
