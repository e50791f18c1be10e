"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): shard ranges, the embedding
all-gather with a padded tail shard, row-block assembly and the label broadcast.  The CUDA kernels
are replaced by numpy/sklearn stand-ins (the oracle), which is what these functions accept."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_emb
from speech_diarization_b200 import sharded


def test_shard_range_and_slices():
    for n in (0, 1, 7, 4799, 38399):
        for world in (1, 2, 4, 8):
            rs = [sharded.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) == (-(-n // world) if n else 0)
    assert sharded.window_count(57_600_000, 24000, 12000) == 4799      # 1 h  
    assert sharded.window_count(8 * 57_600_000, 24000, 12000) == 38399  # 8 h
    assert sharded.window_count(23999, 24000, 12000) == 0
    lo, hi = sharded.shard_range(4799, 1, 2)
    a0, a1 = sharded.audio_slice_for(lo, hi, 24000, 12000)
    assert a0 == lo * 12000 and a1 == (hi - 1) * 12000 + 24000 <= 57_600_000


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cluster_oracle as co
        X, lab = synth_emb(n, 4, 0.02, 3)
        lo, hi = sharded.shard_range(n, rank, world)
        local = torch.from_numpy(X[lo:hi])
        allx = sharded.gather_embeddings(local, n)
        assert torch.equal(allx, torch.from_numpy(X))

        def distance_fn(emb, row0, rows):
            assert rows > 0, "an empty shard must not call the distance kernel"
            return torch.from_numpy(co.cosine_distance(emb.numpy())[row0:row0 + rows].copy())

        def ahc_fn(d, thr):
            return torch.from_numpy(co.ahc_average_precomputed(d.numpy(), thr).astype(np.int32))

        labels = sharded.cluster_sharded(allx, 0.68, distance_fn=distance_fn, ahc_fn=ahc_fn)
        ok = co.same_partition(labels.numpy(), lab)
        block = distance_fn(allx, lo, hi - lo) if hi > lo else torch.empty((0, n), dtype=torch.float32)
        full = sharded.gather_row_blocks(block, n, 0)
        if rank == 0:
            ok = ok and np.array_equal(full.numpy(), co.cosine_distance(X))
        else:
            ok = ok and full is None
        q.put((rank, bool(ok), labels.numpy().tolist()))
    finally:
        dist.destroy_process_group()


# odd n: the last shard is shorter and gets padded; n = 5 on 4 ranks: shards of 2, 2, 1 and an EMPTY one (ranks
# beyond the last shard must still take part in the gather and the broadcast)
@pytest.mark.parametrize("n,world", [(101, 2), (64, 2), (5, 4)])
def test_gather_and_cluster(n, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(r[2] == res[0][2] for r in res)            # every rank holds the same labels
