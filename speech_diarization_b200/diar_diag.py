"""Drop-in for the hot-path pieces of /root/reference/diar_diag.py:
``cluster_embeddings`` (:213-229, "agglo" branch), the ``frame_audio`` duplicate (:48-56), and the score
post-processing of its pipeline: ``whiten_l2`` (:187-194), ``asnorm_scores`` (:196-208) and ``viterbi_hmm``
(:231-247).  Audio loading, plotting and the JSON / SRT / CSV writers are not built."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, postproc
from ._device import require_cuda
from .clustering import cluster_embeddings_device, to_cuda_embeddings


def frame_audio(y: np.ndarray, sr: int, win_ms: float = 30.0, hop_ms: float = 10.0):
    """diar_diag.py:48-56 — pads a too-short signal to one window; returns (frames, hop)."""
    win = int(round(win_ms / 1000.0 * sr))
    hop = int(round(hop_ms / 1000.0 * sr))
    if len(y) < win:
        y = np.pad(y, (0, win - len(y)))
    n = 1 + (len(y) - win) // hop
    idx = np.arange(win)[None, :] + hop * np.arange(n)[:, None]
    return y[idx], hop


def cluster_embeddings(embs: np.ndarray, method="hdbscan", cos_thr: float = 0.68):
    """diar_diag.py:213-229.  method="agglo": average-linkage AHC on 1 - cosine similarity, cut at
    1 - cos_thr, on the GPU (tensor-core affinity + RNN-parallel Lance-Williams).  Labels equal
    sklearn's up to a permutation (numbered by first appearance)."""
    if method == "agglo":
        x = to_cuda_embeddings(embs)
        if x.shape[1] % 64:
            raise _lib.SdError(f"embedding dimension {x.shape[1]} must be a multiple of 64 (ECAPA: 192)")
        return cluster_embeddings_device(x, cos_thr).cpu().numpy().astype(np.int64)
    if method == "hdbscan":
        raise NotImplementedError("method='hdbscan' is outside the B200 hot path (SURVEY.md §2 #9-10); use 'agglo'")
    raise ValueError("method 必须是 hdbscan 或 agglo")      # diar_diag.py:228


def whiten_l2(embs: np.ndarray) -> np.ndarray:
    """diar_diag.py:187-194 — centre, ZCA-whiten with (cov + 1e-6 I)^(-1/2), L2-normalise; float64 [N, D].
    Covariance, eigendecomposition (Jacobi) and the projection all run on the GPU in f64."""
    x = to_cuda_embeddings(embs)
    return postproc.whiten_l2_device(x).cpu().numpy()


def asnorm_scores(query_embs: np.ndarray, ref_centers: np.ndarray, cohort_embs: np.ndarray,
                  topk: int = 200) -> np.ndarray:
    """diar_diag.py:196-208 — adaptive symmetric normalisation of query x centre cosine scores against the
    top-k cohort similarities of each side.  Computed on the GPU in f32 (cohort similarities on the tensor
    cores, statistics in f64); the result has the inputs' result dtype."""
    res_dtype = np.result_type(np.asarray(query_embs).dtype, np.asarray(ref_centers).dtype,
                               np.asarray(cohort_embs).dtype, np.float32)
    q, r, c = (to_cuda_embeddings(x) for x in (query_embs, ref_centers, cohort_embs))
    if q.shape[1] % 64:
        raise _lib.SdError(f"embedding dimension {q.shape[1]} must be a multiple of 64 (ECAPA: 192)")
    return postproc.asnorm_device(q, r, c, topk).cpu().numpy().astype(res_dtype, copy=False)


def viterbi_hmm(scores: np.ndarray, alpha: float = 0.995) -> np.ndarray:
    """diar_diag.py:231-247 — most likely state path of a sticky HMM (stay probability alpha) over per-step
    scores [T, K]; int32 [T].  The float32 recursion of the reference is reproduced exactly on the GPU."""
    s = np.asarray(scores)
    T, K = s.shape
    if s.dtype == np.float16:
        s = s.astype(np.float32)
    elif s.dtype not in (np.float32, np.float64):
        s = s.astype(np.float64)
    d = torch.from_numpy(np.ascontiguousarray(s)).to(require_cuda())
    return postproc.viterbi_device(d, alpha).cpu().numpy().astype(np.int32)


class SpeakerEncoder:
    """diar_diag.py:127-178 — `SpeakerEncoder(backend, device).embed(y, sr) -> np.float32[192]`.
    Only the "speechbrain-ecapa" backend is the B200 path (the shared ECAPA-TDNN plan of
    speech_encode.using_ecapa_encoder); the ONNX back-ends are outside it."""

    def __init__(self, backend="speechbrain-ecapa", device="cuda", ali_model_id=None):
        self.backend = backend
        self.device = device
        self.sr = 16000
        if backend == "speechbrain-ecapa":
            from .speech_encode import using_ecapa_encoder
            self.model = using_ecapa_encoder(device)
            self.kind = "sb"
        elif backend in ("ali-eres2netv2", "ali-campp"):
            raise NotImplementedError(f"backend {backend!r} (ONNX Runtime) is outside the B200 hot path; use speechbrain-ecapa")
        else:
            raise ValueError("backend 必须是 speechbrain-ecapa / ali-eres2netv2 / ali-campp")     # diar_diag.py:159

    def embed(self, y: np.ndarray, sr: int) -> np.ndarray:
        if sr != self.sr:
            raise NotImplementedError("resampling (librosa) is outside the B200 hot path; pass 16 kHz audio")
        wav = torch.from_numpy(np.ascontiguousarray(y)).float().unsqueeze(0)          # :166
        with torch.inference_mode():
            e = self.model.encode_batch(wav).squeeze(0).squeeze(0).cpu().numpy()         # :167-168
        return e.astype(np.float32)


def pad_with_context(y: np.ndarray, sr: int, start: float, end: float, ctx: float = 0.2):
    """diar_diag.py:182-185."""
    s = max(0, int((start - ctx) * sr))
    e = min(len(y), int((end + ctx) * sr))
    return y[s:e]


def merge_segments(segs, labs, gap: float = 0.10):
    """The nested helper of main() (diar_diag.py:398-409): join consecutive same-speaker segments whose gap <= `gap`."""
    out = []
    cur_lab, s, e = labs[0], segs[0][0], segs[0][1]
    for (st, ed), lb in zip(segs[1:], labs[1:]):
        if lb == cur_lab and st - e <= gap:
            e = ed
        else:
            out.append([s, e, int(cur_lab)])
            cur_lab, s, e = lb, st, ed
    out.append([s, e, int(cur_lab)])
    return out


def label_segments(embs: np.ndarray, segs, whiten: int = 1, asnorm: int = 1, cluster: str = "agglo",
                   cos_thr: float = 0.68, use_vbx: int = 1, alpha: float = 0.995, min_gap_ms: float = 100.0):
    """The body of main() between the embedding loop and the export (diar_diag.py:352-411) as one call with every
    array kept on the GPU: whiten -> cluster -> unit-norm centres -> (AS-norm) scores -> (Viterbi | argmax) ->
    merge.  Returns (merged [[start, end, speaker_index]], final_labels int [N], cluster labels int [N]).
    Speaker indices follow this package's cluster numbering (order of first appearance), i.e. they equal the
    reference's up to a permutation."""
    if cluster != "agglo":
        raise NotImplementedError("cluster='hdbscan' is outside the B200 hot path (SURVEY.md §2 #9-10); use 'agglo'")
    x32 = to_cuda_embeddings(embs)
    if x32.shape[1] % 64:
        raise _lib.SdError(f"embedding dimension {x32.shape[1]} must be a multiple of 64 (ECAPA: 192)")
    x64 = postproc.whiten_l2_device(x32) if whiten else x32.to(torch.float64)           # :352
    xw = x64.to(torch.float32)
    labels = cluster_embeddings_device(xw, cos_thr)                                     # :371-374
    K = int(labels.max().item()) + 1
    centers = postproc.cluster_centers_device(x64, labels, K)                           # :377-383
    if asnorm:
        scores = postproc.asnorm_device(xw, centers, xw, min(200, xw.shape[0]))         # :389
    else:
        scores = postproc.dot_scores_device(xw, centers)                                 # :386
    if use_vbx:
        final = postproc.viterbi_device(scores, alpha)                                   # :393
    else:
        final = scores.argmax(dim=1)                                                     # :396 (index bookkeeping)
    final = final.cpu().numpy()
    merged = merge_segments(list(segs), final, gap=min_gap_ms / 1000.0)                  # :411
    return merged, final, labels.cpu().numpy()
