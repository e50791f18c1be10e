"""Drop-in for the OUTPUT FORMAT of /root/reference/diarization_baseline.py (:259-266):
``list[(start, end, speaker)]`` sorted by start and NIST RTTM text.  The pyannote pipeline
itself (segmentation model, HF-gated weights) is out of scope."""
from __future__ import annotations

from pathlib import Path


def segments_to_tuples(segments, label_fmt: str = "SPEAKER_{:02d}") -> list:
    """Segment(start, end, spk) -> (start, end, "SPEAKER_xx") tuples sorted by start, the shape
    diarize_audio returns (diarization_baseline.py:259-261); pyannote names speakers SPEAKER_00…"""
    out = [(float(s.start), float(s.end), label_fmt.format(int(s.spk)) if isinstance(s.spk, (int,)) or
            hasattr(s.spk, "__index__") else s.spk) for s in segments]
    return sorted(out, key=lambda t: (t[0], t[1]))


def rttm_lines(segments: list, uri: str) -> list:
    """One RTTM line per turn, as pyannote.core.Annotation.write_rttm formats it."""
    return [f"SPEAKER {uri} 1 {start:.3f} {end - start:.3f} <NA> <NA> {label} <NA> <NA>\n"
            for start, end, label in segments]


def write_rttm(segments: list, rttm_filepath: str | Path, uri: str | None = None) -> None:
    """diarization_baseline.py:263-265."""
    uri = uri or Path(rttm_filepath).stem
    with open(rttm_filepath, "w") as f:
        f.writelines(rttm_lines(segments, uri))


# ------------------------------------------------------------------ pyannote-style clustering
import numpy as np  # noqa: E402


def linkage_centroid(embeddings: np.ndarray) -> np.ndarray:
    """scipy.cluster.hierarchy.linkage(embeddings, "centroid", "euclidean") on the GPU (csrc/centroid.cu):
    the (N-1) x 4 linkage matrix, f64."""
    import torch
    from . import _lib
    from ._device import require_cuda, to_device_f32
    lib = _lib.load()
    x = to_device_f32(np.ascontiguousarray(embeddings, dtype=np.float32), require_cuda())
    N, D = x.shape
    Z = torch.empty((N - 1, 4), dtype=torch.float64, device=x.device)
    ws = torch.empty((lib.sd_centroid_linkage_workspace_bytes(N, D),), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sd_centroid_linkage_f64(x.data_ptr(), N, D, Z.data_ptr(), ws.data_ptr(), _lib.stream_ptr()),
                   "sd_centroid_linkage_f64")
    return Z.cpu().numpy()


def fcluster_distance(Z: np.ndarray, t: float) -> np.ndarray:
    """scipy.cluster.hierarchy.fcluster(Z, t, criterion="distance") - 1 as a partition (cluster ids 0..k-1 in
    order of first appearance): observations share a flat cluster iff their cophenetic distance is <= t.  Safe
    under the inversions centroid linkage produces: a merge node forms a cluster only if the MAXIMUM height in
    its whole subtree is <= t (scipy's ``get_max_dist_for_each_cluster`` + ``cluster_monocrit``)."""
    n = Z.shape[0] + 1
    left = Z[:, 0].astype(np.int64)
    right = Z[:, 1].astype(np.int64)
    md = Z[:, 2].copy()
    for i in range(n - 1):                     # children always precede their parent
        a, b = left[i], right[i]
        if a >= n and md[a - n] > md[i]:
            md[i] = md[a - n]
        if b >= n and md[b - n] > md[i]:
            md[i] = md[b - n]
    root = np.full(2 * n - 1, -1, dtype=np.int64)   # id of the subtree root that defines the node's flat cluster
    for i in range(n - 2, -1, -1):             # parents before children
        node = n + i
        if root[node] < 0 and md[i] <= t:
            root[node] = node
        if root[node] >= 0:
            root[left[i]] = root[node]
            root[right[i]] = root[node]
    leaf_root = root[:n].copy()
    single = leaf_root < 0
    leaf_root[single] = np.flatnonzero(single)  # singletons are their own cluster
    first = {}
    return np.array([first.setdefault(int(r), len(first)) for r in leaf_root], dtype=np.int64)


class AgglomerativeClustering:
    """Stand-in for ``pipeline.clustering`` of pyannote/speaker-diarization-3.1 as the reference configures it
    (diarization_baseline.py:180 sets ``.threshold``; ``min_speakers`` / ``max_speakers`` arrive as
    ``min_clusters`` / ``max_clusters``, :252-257).  Method "centroid", metric "cosine"; SURVEY.md App. B
    (recalled — pyannote is not installable here, parity is against oracle/cluster_oracle.pyannote_agglomerative)."""

    def __init__(self, threshold: float = 0.7045654963945799, min_cluster_size: int = 12):
        self.threshold = threshold
        self.min_cluster_size = min_cluster_size
        self.method = "centroid"
        self.metric = "cosine"

    @staticmethod
    def _large(clusters, min_cluster_size):
        uniq, counts = np.unique(clusters, return_counts=True)
        return uniq, counts, uniq[counts >= min_cluster_size]

    def cluster(self, embeddings: np.ndarray, min_clusters: int = 1, max_clusters=np.inf) -> np.ndarray:
        n = embeddings.shape[0]
        min_cluster_size = min(self.min_cluster_size, max(1, round(0.1 * n)))
        if n == 1:
            return np.zeros((1,), dtype=np.uint8)
        emb = embeddings / np.linalg.norm(embeddings, axis=-1, keepdims=True)
        Z = linkage_centroid(emb)
        clusters = fcluster_distance(Z, self.threshold)
        uniq, counts, large = self._large(clusters, min_cluster_size)
        if len(large) < min_clusters:
            num_clusters = min_clusters
        elif len(large) > max_clusters:
            num_clusters = max_clusters
        else:
            num_clusters = None
        if num_clusters is not None:
            Zi = Z.copy()
            Zi[:, 2] = np.arange(n - 1)
            best_iteration, best_num_large = n - 1, 1
            for iteration in np.argsort(np.abs(Z[:, 2] - self.threshold)):
                if Zi[iteration, 3] < min_cluster_size:
                    continue
                clusters = fcluster_distance(Zi, iteration)
                uniq, counts, large = self._large(clusters, min_cluster_size)
                if abs(len(large) - num_clusters) < abs(best_num_large - num_clusters):
                    best_iteration, best_num_large = iteration, len(large)
                if len(large) == num_clusters:
                    break
            if best_num_large != num_clusters:
                clusters = fcluster_distance(Zi, best_iteration)
                uniq, counts, large = self._large(clusters, min_cluster_size)
        if len(large) == 0:
            clusters[:] = 0
            return clusters
        small = uniq[counts < min_cluster_size]
        if len(small) == 0:
            return clusters
        large_c = np.vstack([emb[clusters == k].mean(axis=0) for k in large])
        small_c = np.vstack([emb[clusters == k].mean(axis=0) for k in small])
        ln = large_c / np.linalg.norm(large_c, axis=1, keepdims=True)
        sn = small_c / np.linalg.norm(small_c, axis=1, keepdims=True)
        cos_dist = 1.0 - ln @ sn.T                                   # cdist(..., metric="cosine")
        for small_k, large_k in enumerate(np.argmin(cos_dist, axis=0)):
            clusters[clusters == small[small_k]] = large[large_k]
        _, clusters = np.unique(clusters, return_inverse=True)
        return clusters
