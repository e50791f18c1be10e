"""Centroid-linkage clustering (what diarization_baseline.py's pyannote pipeline runs):
GPU linkage matrix against scipy's, and the pyannote-style flat clustering against the oracle restatement
(PARITY UNPINNED for the pyannote post-processing: recalled from SURVEY App. B, see oracle/cluster_oracle.py)."""
import numpy as np
import pytest

from conftest import synth_emb
from oracle import cluster_oracle as co
from speech_diarization_b200 import diarization_baseline as db

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,sigma", [(2, 1, 0.1), (3, 2, 0.1), (50, 3, 0.05), (500, 5, 0.05), (2000, 8, 0.03),
                                       (1500, 4, 0.3)])     # sigma 0.3: no cluster structure, many inversions
def test_linkage_matrix_matches_scipy(N, K, sigma):
    from scipy.cluster.hierarchy import linkage
    X, _ = synth_emb(N, K, sigma, N + K)
    X = (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32)
    ref = linkage(X, method="centroid", metric="euclidean")
    got = db.linkage_centroid(X)
    assert got.shape == ref.shape == (N - 1, 4)
    np.testing.assert_array_equal(got[:, [0, 1, 3]], ref[:, [0, 1, 3]])      # same merges in the same order
    np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=0, atol=1e-9)      # f64 centroid distances
    if N >= 500:
        assert (np.diff(ref[:, 2]) < 0).any()                                 # the data does produce inversions


@pytest.mark.parametrize("N,K,sigma,kw", [
    (300, 4, 0.05, {}),                                   # plain threshold cut + small-cluster reassignment
    (300, 4, 0.05, {"min_clusters": 6}),                  # too few large clusters -> dendrogram re-cut
    (300, 4, 0.05, {"max_clusters": 2}),                  # too many -> re-cut
    (1000, 6, 0.2, {"min_clusters": 2, "max_clusters": 6}),   # min/max_speakers as the reference passes them
    (40, 2, 0.05, {}),                                    # min_cluster_size shrinks to round(0.1 N)
    (5, 1, 0.05, {}),
])
def test_pyannote_style_clustering_matches_oracle(N, K, sigma, kw):
    X, _ = synth_emb(N, K, sigma, 3 * N + K)
    ref = co.pyannote_agglomerative(X.copy(), threshold=0.7045654963945799, min_cluster_size=12,
                                    min_clusters=kw.get("min_clusters", 1), max_clusters=kw.get("max_clusters", np.inf))
    got = db.AgglomerativeClustering().cluster(X.copy(), **kw)
    assert co.same_partition(got, ref), (len(set(ref.tolist())), len(set(got.tolist())))


def test_threshold_attribute_is_honoured():
    """diarization_baseline.py:180: ``diarizer.clustering.threshold = clustering_threshold``."""
    X, lab = synth_emb(400, 4, 0.05, 9)
    c = db.AgglomerativeClustering()
    c.threshold = 0.70                                     # the value the reference hard-codes (:248)
    got = c.cluster(X)
    assert co.same_partition(got, co.pyannote_agglomerative(X, threshold=0.70))
    c.threshold = 5.0                                      # everything merges
    assert len(set(c.cluster(X).tolist())) == 1
    assert db.AgglomerativeClustering().cluster(X[:1]).tolist() == [0]
