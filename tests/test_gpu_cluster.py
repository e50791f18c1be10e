"""Affinity + AHC + scoring kernels against the oracle (sklearn / scipy, the reference's own calls).

Gates (BASELINE.json north_star): affinity within 1e-5 absolute; cluster labels identical up to
permutation."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_emb
from oracle import cluster_oracle as co
from speech_diarization_b200 import clustering as cl
from speech_diarization_b200 import diar_diag, anti_stick_diarize as asd

pytestmark = pytest.mark.gpu
AFF_TOL = 1e-5


@pytest.mark.parametrize("N", [2, 127, 128, 129, 1000, 5000])
def test_cosine_distance_within_1e5(N):
    X, _ = synth_emb(N, 8, 0.02, N)
    X *= np.random.default_rng(N).uniform(0.1, 30.0, (N, 1)).astype(np.float32)     # un-normalised input
    ref = co.cosine_distance(X)
    got = cl.cosine_distance_device(torch.from_numpy(X).cuda()).cpu().numpy()
    assert got.shape == (N, N) and got.dtype == np.float32
    assert np.abs(got - ref).max() <= AFF_TOL
    assert np.array_equal(got, got.T)                 # exactly symmetric
    assert np.abs(np.diag(got)).max() <= AFF_TOL


def test_cosine_distance_zero_rows_and_random_embeddings():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((300, 192)).astype(np.float32)
    X[7] = 0.0                                        # sklearn.normalize keeps zero rows at zero -> D = 1
    ref = co.cosine_distance(X)
    got = cl.cosine_distance_device(torch.from_numpy(X).cuda()).cpu().numpy()
    assert np.abs(got - ref).max() <= AFF_TOL
    assert np.all(got[7] == 1.0)


def test_row_block_sharding_assembles_full_matrix():
    X, _ = synth_emb(1000, 8, 0.02, 1)
    xd = torch.from_numpy(X).cuda()
    full = cl.cosine_distance_device(xd)
    parts = [cl.cosine_distance_device(xd, r0, min(250, 1000 - r0)) for r0 in range(0, 1000, 250)]
    assert torch.equal(torch.cat(parts, 0), full)
    f32, f64 = cl.cosine_distance_device(xd, 100, 300, want_f64=True)
    assert torch.equal(f32.double(), f64)


def test_empty_row_block_is_not_an_error():
    """A rank beyond the last shard (n = 9 windows on 8 GPUs) asks for zero rows: an empty block, no SdError
    (the other ranks would otherwise wait for it in the gather forever)."""
    X, _ = synth_emb(9, 2, 0.02, 1)
    xd = torch.from_numpy(X).cuda()
    assert cl.cosine_distance_device(xd, 9, 0).shape == (0, 9)
    assert cl.cosine_distance_device(xd, 4, 0).shape == (0, 9)
    from speech_diarization_b200 import sharded
    assert [sharded.shard_range(9, r, 8) for r in (4, 5, 7)] == [(8, 9), (9, 9), (9, 9)]


def test_anti_stick_cosine_distance_head():
    X, _ = synth_emb(500, 4, 0.05, 2)
    ref = 1 - co.cosine_similarity(co.l2_normalize(X))          # anti_stick_diarize.py:176-177
    assert np.abs(asd.cosine_distance(X) - ref).max() <= AFF_TOL


@pytest.mark.parametrize("tag", ["clean", "noisy", "tiny"])
def test_cluster_embeddings_matches_reference_golden(tag):
    g = golden("cluster_ref.npz")
    got = diar_diag.cluster_embeddings(g[f"{tag}_X"], method="agglo", cos_thr=0.68)
    assert got.dtype == np.int64 and got.shape == g[f"{tag}_labels"].shape
    assert co.same_partition(got, g[f"{tag}_labels"])
    got05 = diar_diag.cluster_embeddings(g[f"{tag}_X"], method="agglo", cos_thr=0.5)
    assert co.same_partition(got05, g[f"{tag}_labels_thr05"])


@pytest.mark.parametrize("N,K,sigma,thr", [
    (2, 1, 0.02, 0.68), (2, 2, 0.02, 0.68), (50, 3, 0.02, 0.68), (1000, 8, 0.02, 0.68),
    (1000, 8, 0.05, 0.68),      # sits on the threshold: shatters into hundreds of clusters
    (1000, 8, 0.05, 0.3),       # same data, looser cut
    (3000, 16, 0.03, 0.68), (5000, 8, 0.02, 0.68),
])
def test_ahc_labels_identical_up_to_permutation(N, K, sigma, thr):
    X, _ = synth_emb(N, K, sigma, N + K)
    ref = co.cluster_embeddings(X, "agglo", thr)
    got = diar_diag.cluster_embeddings(X, "agglo", thr)
    assert co.same_partition(got, ref), (len(set(ref)), len(set(got)))
    # labels are numbered by first appearance
    first = [np.flatnonzero(got == l)[0] for l in range(got.max() + 1)]
    assert first == sorted(first)


def test_ahc_on_arbitrary_precomputed_matrix():
    """Not a cosine matrix: random symmetric distances (no cluster structure, long chains)."""
    rng = np.random.default_rng(5)
    A = rng.uniform(0.0, 1.0, (400, 400)).astype(np.float32)
    D = np.triu(A, 1)
    D = D + D.T
    for thr in (0.2, 0.45, 0.6):
        ref = co.ahc_average_precomputed(D, thr)
        got, ncl = cl.ahc_average_device(torch.from_numpy(D).cuda(), thr)
        assert co.same_partition(got.cpu().numpy(), ref)
        assert int(ncl.item()) == len(set(ref.tolist()))


def test_ahc_properties_at_n20k():
    """BASELINE size (N = 20 000, K = 8, sigma = 0.02): the oracle needs ~30 s here, so check
    size-independent properties: the planted partition is recovered, labels are a valid
    numbering, and clustering a permuted copy gives the permuted partition."""
    X, lab = synth_emb(20000, 8, 0.02, 0)
    xd = torch.from_numpy(X).cuda()
    got = cl.cluster_embeddings_device(xd, 0.68).cpu().numpy()
    assert co.same_partition(got, lab)
    assert sorted(set(got.tolist())) == list(range(8))
    perm = np.random.default_rng(1).permutation(20000)
    got_p = cl.cluster_embeddings_device(xd[torch.from_numpy(perm).cuda()].contiguous(), 0.68).cpu().numpy()
    assert co.same_partition(got_p, got[perm])


def test_ahc_repeated_runs_agree():
    """The round loop is a chain of grid-wide phases over shared counters: repeated runs must give the same labels, the
    same number of rounds and N - K merges (a counter reset one barrier too early once let a slow CTA leave the loop:
    right answer on most runs, wrong partition on some)."""
    for n, k in ((6000, 8), (20000, 8)):
        X, lab = synth_emb(n, k, 0.02, 11)
        dist = cl.cosine_distance_device(torch.from_numpy(X).cuda())
        cl.ahc_keep_stats(True)
        try:
            first, stats0 = None, None
            for rep in range(6):
                got, ncl = cl.ahc_average_device(dist, 1 - 0.68)
                stats = cl.ahc_last_stats()
                g = got.cpu().numpy()
                assert int(ncl.item()) == k and stats["merges"] == n - k
                if first is None:
                    first, stats0 = g, stats
                    assert co.same_partition(g, lab)
                else:
                    assert np.array_equal(g, first) and stats == stats0
        finally:
            cl.ahc_keep_stats(False)


@pytest.mark.parametrize("tag", ["clean", "edge"])
def test_ahc_n20k_labels_match_reference_golden(tag):
    """BASELINE config 4 at N = 20 000 against the labels of the reference's own cluster_embeddings (diar_diag.py:213-229;
    tests/golden/make_golden.py cluster20k): sigma = 0.02 -> 8 clusters, and sigma = 0.045 where the intra-cluster
    cosine sits near the 0.68 threshold -> 38 clusters.

    "clean" runs the whole device path (affinity kernel + AHC kernel).  "edge" feeds the AHC kernel the oracle's own
    distance matrix: that data set has near-tied sub-clusters just under the cut, and which of them pairs up first
    flips with perturbations of the distances far inside the 1e-5 affinity gate (measured: the device affinity is
    within 1.6e-6 of sklearn's everywhere, yet scipy itself run on it finds 39 clusters, its last merge moving from
    2.0e-5 below the threshold to 2.0e-5 above).  So at this tolerance only the clustering kernel can be held to the
    reference's labels there, and it is: same matrix in, same partition out."""
    g = golden("cluster_ref_20k.npz")
    N, K, seed = (int(v) for v in g[f"{tag}_params"])
    X, _ = synth_emb(N, K, float(g[f"{tag}_sigma"]), seed)
    assert abs(float(X.astype(np.float64).sum()) - float(g[f"{tag}_xsum"])) < 1e-6      # same inputs as the golden run
    ref = g[f"{tag}_labels"].astype(np.int64)
    if tag == "clean":
        got = diar_diag.cluster_embeddings(X, method="agglo", cos_thr=0.68)
    else:
        D = torch.from_numpy(co.cosine_distance(X)).cuda()
        assert float((cl.cosine_distance_device(torch.from_numpy(X).cuda()) - D).abs().max()) <= AFF_TOL
        got = cl.ahc_average_device(D, 1 - 0.68)[0].cpu().numpy()
    assert len(set(got.tolist())) == len(set(ref.tolist())) == (8 if tag == "clean" else 38)
    assert co.same_partition(got, ref)


def test_ahc_with_exactly_duplicated_rows():
    """Many exactly duplicated embeddings (silent windows): zero distances and exact f64 ties everywhere.  The
    reciprocal-nearest-neighbour rounds must still terminate with sklearn's partition."""
    X, _ = synth_emb(600, 5, 0.03, 17)
    X[100:300] = X[100]                  # 200 identical rows
    X[300:340] = X[301]
    X[500:] = X[[3, 7]].repeat(50, axis=0)
    for thr in (0.68, 0.9):
        ref = co.cluster_embeddings(X, "agglo", thr)
        got = diar_diag.cluster_embeddings(X, "agglo", thr)
        assert co.same_partition(got, ref), (thr, len(set(ref)), len(set(got)))
    Y = np.tile(X[:1], (257, 1))         # nothing but duplicates
    assert len(set(diar_diag.cluster_embeddings(Y, "agglo", 0.68).tolist())) == 1


def test_single_sample_raises_like_sklearn():
    with pytest.raises(ValueError):
        diar_diag.cluster_embeddings(np.ones((1, 192), np.float32), "agglo")


def test_window_argmax_and_adjacent_cosine():
    X, _ = synth_emb(3000, 5, 0.3, 9)
    C, _ = synth_emb(5, 5, 0.0, 10)
    C = co.l2_normalize(C).astype(np.float32)
    ref = co.window_argmax(X.copy(), C)
    xn = cl.l2_normalize_device(torch.from_numpy(X).cuda())
    best, score = cl.window_argmax_device(xn, torch.from_numpy(C).cuda())
    assert np.array_equal(best.cpu().numpy(), ref)
    assert np.abs(xn.cpu().numpy() - co.l2_normalize(X)).max() < 1e-6
    assert np.abs(asd.adjacent_cosine(X) - co.adjacent_cosine(X)).max() < 1e-5
    assert asd.adjacent_cosine(X[:1]).shape == (0,)
