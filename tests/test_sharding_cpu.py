"""-m "not gpu": the N>1 plumbing (tile maps, query shards, boundary-node all-gather) with the
gloo backend, world_size 2, on CPU."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_tiles_form_one_continuous_heightfield(pkg):
    a = pkg.terrain.mountain(120, h=0.1, seed=2, tile=(0, 0), world_tiles=(2, 1), shuffle=False)
    b = pkg.terrain.mountain(120, h=0.1, seed=2, tile=(1, 0), world_tiles=(2, 1), shuffle=False)
    assert a[:, 0].max() < 12.0 + 0.05 and b[:, 0].min() > 12.0 - 0.15
    # heights on both sides of the shared border x = 12 m agree to the terrain's local slope
    la = a[a[:, 0] > 11.85]
    lb = b[b[:, 0] < 12.05]
    from scipy.spatial import cKDTree
    d, j = cKDTree(lb[:, :2]).query(la[:, :2])
    near = d < 0.15
    assert near.sum() > 50
    assert np.abs(la[near, 2] - lb[j[near], 2]).max() < 0.25
    # a single-tile world is the unsharded map
    c = pkg.terrain.mountain(60, h=0.1, seed=2)
    d2 = pkg.terrain.mountain(60, h=0.1, seed=2, tile=(0, 0), world_tiles=(1, 1))
    np.testing.assert_array_equal(c, d2)


def test_query_shards_partition_the_batch(pkg):
    from trg_planner_b200 import sharding
    for n in (0, 1, 7, 1000, 10001):
        for w in (1, 2, 4, 8):
            sl = [sharding.query_shard(n, r, w) for r in range(w)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(sl[i].stop == sl[i + 1].start for i in range(w - 1))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    _pkg.load()
    from trg_planner_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 50 + 30 * rank                      # ragged sizes; rank-dependent
    x_lo, x_hi = 10.0 * rank, 10.0 * (rank + 1)
    pos = np.column_stack([rng.uniform(x_lo, x_hi, n), rng.uniform(0, 10, n), rng.normal(size=n)]).astype(np.float32)
    ids = (np.arange(n) + 20_000_000 * rank).astype(np.int32)   # beyond 2^24: must survive the float bit-cast
    sel = sharding.boundary_nodes(pos, x_lo, x_hi, 0.9, rank, world)
    got = sharding.allgather_boundary(dist, torch, pos[sel], ids[sel], torch.device("cpu"))
    q.put((rank, pos, ids, sel, got))
    dist.barrier()
    dist.destroy_process_group()


def test_boundary_allgather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        _, pos, ids, sel, got = res[rank]
        # rank 0 shares only its right border, rank 1 only its left one
        if rank == 0:
            assert (pos[sel, 0] > 10.0 - 0.9).all()
        else:
            assert (pos[sel, 0] < 10.0 + 0.9).all()
        for other in (0, 1):
            _, opos, oids, osel, _ = res[other]
            np.testing.assert_array_equal(got[other][0], opos[osel])
            np.testing.assert_array_equal(got[other][1], oids[osel])
    from trg_planner_b200 import sharding
    pairs = sharding.cross_tile_candidates(res[0][4][0][0], res[0][4][1][0], 0.9)
    for i, j in pairs:
        a, b = res[0][4][0][0][i], res[0][4][1][0][j]
        assert np.hypot(a[0] - b[0], a[1] - b[1]) <= 0.9 + 1e-6
