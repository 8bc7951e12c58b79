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
    dd = np.hypot(la[:, None, 0] - lb[None, :, 0], la[:, None, 1] - lb[None, :, 1])
    j, d = dd.argmin(1), dd.min(1)
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


def _stitch_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    trg = _pkg.load()
    from trg_planner_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = trg.MOUNTAIN
    side = 140
    pts = trg.terrain.mountain(side, h=0.1, seed=2, tile=(rank, 0), world_tiles=(world, 1))
    x_lo, x_hi = rank * side * 0.1, (rank + 1) * side * 0.1
    oracle = _pkg.load_oracle().oracle
    o = oracle(P)
    o.seed(42 + rank)
    o.set_global_map(pts)
    assert o.init_graph((0.5 * (x_lo + x_hi), 7.0, 0.0)) == 0
    g = o.export()

    def edge_eval(strip_pts, p1, p2):      # CPU tests: the oracle stands in for the K4 kernels
        e = oracle(P)
        e.set_global_map(strip_pts)
        r = e.edge_eval(p1, p2)
        return r["stage"], r["weight"], r["dist"]

    edges, st = sharding.stitch_tiles(dist, torch, torch.device("cpu"), rank, world, pts, g.pos, g.ids, x_lo, x_hi,
                                      P.expand_dist, P.robot_size, edge_eval)
    m = sharding.merge_graphs(dist, torch, torch.device("cpu"), rank, world, g, edges)
    merged = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in m.items()}
    q.put((rank, edges, st, g.pos, g.ids, pts, merged, (g.row_ptr, g.col, g.weight, g.dist)))
    dist.barrier()
    dist.destroy_process_group()


def test_stitch_tiles_gloo_world2(pkg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_stitch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=300)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    e0, e1 = res[0][1], res[1][1]
    np.testing.assert_array_equal(e0, e1)              # every rank ends with the same stitched edge list
    assert res[0][2]["stitched_edges_total"] == len(e0) > 20
    assert res[0][2]["strip_points"] > 1000 and res[0][2]["candidate_pairs"] >= len(e0)
    # every stitched edge joins a node of tile 0 to a node of tile 1 closer than expand_dist, and
    # re-evaluating it on the merged strips gives the same risk
    P = pkg.MOUNTAIN
    pos0 = {int(i): p for i, p in zip(res[0][4], res[0][3])}
    pos1 = {int(i): p for i, p in zip(res[1][4], res[1][3])}
    merged = np.concatenate([res[0][5][res[0][5][:, 0] > 14.0 - 2.0], res[1][5][res[1][5][:, 0] < 14.0 + 2.0]])
    o = pkg.oracle(P)
    o.set_global_map(merged)
    p1 = np.array([pos0[int(e[1])] for e in e0], np.float32)
    p2 = np.array([pos1[int(e[3])] for e in e0], np.float32)
    assert (e0[:, 0] == 0).all() and (e0[:, 2] == 1).all()
    d = np.hypot(p1[:, 0] - p2[:, 0], p1[:, 1] - p2[:, 1])
    # (a node may sit up to robot_size beyond the last points of its own tile)
    assert (d < P.expand_dist + 1e-6).all() and (p1[:, 0] < 14.0 + P.robot_size + 0.05).all() and (p2[:, 0] > 14.0 - P.robot_size - 0.05).all()
    ev = o.edge_eval(p1, p2)
    assert (ev["stage"] == 0).all()
    np.testing.assert_allclose(ev["weight"], e0[:, 4], rtol=1e-6)
    np.testing.assert_allclose(ev["dist"], e0[:, 5], rtol=1e-6)


def test_stitch_tiles_gloo_world3_middle_tile_has_two_borders(pkg):
    """N >= 3: the middle rank shares a border with both neighbours (what N = 4 / 8 runs look like)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 3
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_stitch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=400)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    e = res[0][1]
    for r in range(1, world):
        np.testing.assert_array_equal(e, res[r][1])
    assert res[1][2]["stitched_edges_total"] == len(e)
    # edges only between x-adjacent tiles, and both borders are stitched
    assert (e[:, 2] == e[:, 0] + 1).all()
    for left in (0, 1):
        sel = e[e[:, 0] == left]
        assert len(sel) > 10
        border = 14.0 * (left + 1)
        pl = {int(i): p for i, p in zip(res[left][4], res[left][3])}
        pr = {int(i): p for i, p in zip(res[left + 1][4], res[left + 1][3])}
        a = np.array([pl[int(v[1])] for v in sel], np.float32)
        b = np.array([pr[int(v[3])] for v in sel], np.float32)
        assert (np.hypot(a[:, 0] - b[:, 0], a[:, 1] - b[:, 1]) < pkg.MOUNTAIN.expand_dist + 1e-6).all()
        assert (np.abs(a[:, 0] - border) < 1.0).all() and (np.abs(b[:, 0] - border) < 1.0).all()


def test_merged_graph_gloo_world2_paths_cross_the_border(pkg):
    """ONE graph across the ranks: both ranks assemble the same global CSR (local edges in local order, then
    the stitched ones, both directions), and a shortest path from tile 0 to tile 1 exists on it."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 150)
    procs = [ctx.Process(target=_stitch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=300)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m0, m1 = res[0][6], res[1][6]
    for k in ("pos", "state", "row_ptr", "col", "weight", "dist", "node_off"):
        np.testing.assert_array_equal(m0[k], m1[k], err_msg=k)
    n0, n1 = len(res[0][4]), len(res[1][4])
    st = res[0][1]
    assert m0["n_nodes"] == n0 + n1 and list(m0["node_off"]) == [0, n0, n0 + n1]
    assert m0["n_edges"] == len(res[0][7][1]) + len(res[1][7][1]) + 2 * len(st)
    row, col, w, d = m0["row_ptr"], m0["col"], m0["weight"], m0["dist"]
    assert row[0] == 0 and row[-1] == len(col) and (np.diff(row) >= 0).all() and col.min() >= 0 and col.max() < m0["n_nodes"]
    # tile-local adjacency survives as the head of every row (ids shifted by the owner's offset)
    for rank, off in ((0, 0), (1, n0)):
        lrow, lcol, lw, ld = res[rank][7]
        for v in (0, len(lrow) // 3, len(lrow) - 2):
            a, b = lrow[v], lrow[v + 1]
            np.testing.assert_array_equal(col[row[v + off]: row[v + off] + (b - a)], lcol[a:b] + off)
            np.testing.assert_array_equal(w[row[v + off]: row[v + off] + (b - a)], lw[a:b])
    # every stitched edge is there in both directions with its risk and length
    for e in st[:50]:
        ga, gb = int(e[1]), n0 + int(e[3])
        ia = np.nonzero(col[row[ga]:row[ga + 1]] == gb)[0]
        ib = np.nonzero(col[row[gb]:row[gb + 1]] == ga)[0]
        assert len(ia) == 1 and len(ib) == 1
        assert w[row[ga] + ia[0]] == np.float32(e[4]) and d[row[gb] + ib[0]] == np.float32(e[5])
    # Dijkstra over the merged CSR from the root of tile 0 reaches the far side of tile 1
    import heapq
    cost = (np.float32(pkg.MOUNTAIN.safety_factor) * w + 1) * d
    dist_ = np.full(m0["n_nodes"], np.inf)
    src = 0
    dist_[src] = 0.0
    pq = [(0.0, src)]
    while pq:
        du, u = heapq.heappop(pq)
        if du > dist_[u]:
            continue
        for k in range(row[u], row[u + 1]):
            v = col[k]
            if m0["state"][v] == -1:
                continue
            nd = du + cost[k]
            if nd < dist_[v]:
                dist_[v] = nd
                heapq.heappush(pq, (nd, v))
    far = np.nonzero(m0["pos"][:, 0] > 24.0)[0]
    assert len(far) > 10 and np.isfinite(dist_[far]).mean() > 0.9


def _qlist_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    trg = _pkg.load()
    from trg_planner_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = trg.MOUNTAIN
    pts = trg.terrain.stairs(160, h=0.1, seed=5)
    rng = np.random.default_rng(3)
    qxy = rng.uniform(pts[:, :2].min(0) - 0.5, pts[:, :2].max(0) + 0.5, size=(3000, 2)).astype(np.float32)
    oracle = _pkg.load_oracle().oracle

    def k2_k3(part, idx):   # CPU tests: the oracle stands in for the K2 / K3 kernels
        o = oracle(P)
        o.set_global_map(part)
        coll = o.is_collision(qxy[idx], P.collision_threshold).astype(np.float32)
        cnt = o.range_count(qxy[idx], P.robot_size).astype(np.int32).view(np.float32)
        z, _, tie = o.nearest_z(qxy[idx])
        return np.column_stack([coll, cnt, z, tie.astype(np.float32)])

    # K3 looks at the nearest point wherever it is: on this dense map it is closer than 0.3 m from every query
    rows, st = sharding.sharded_query_list(dist, torch, torch.device("cpu"), rank, world, pts, qxy, max(P.robot_size, 0.3) + 0.05, k2_k3)
    q.put((rank, rows, st, qxy, pts))
    dist.barrier()
    dist.destroy_process_group()


def test_query_list_morton_partition_gloo_world4_equals_unsharded(pkg):
    """SURVEY 8(e), pure kernels on a fixed query list: Morton partition of the queries, every rank holds only the map
    points within the halo of its run, all-gather of the outputs — and the answers are those of the whole map."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 4
    port = 29450 + (os.getpid() % 100)
    procs = [ctx.Process(target=_qlist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=400)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows, qxy, pts = res[0][1], res[0][3], res[0][4]
    for r in range(1, world):
        np.testing.assert_array_equal(rows, res[r][1])
    assert sum(res[r][2]["mine"] for r in range(world)) == len(qxy)
    # a quarter of a Z-curve is a quadrant: every rank holds about a quarter of the map plus the halo
    assert all(res[r][2]["map_points_mine"] < 0.4 * len(pts) for r in range(world))
    P = pkg.MOUNTAIN
    o = pkg.oracle(P)
    o.set_global_map(pts)
    np.testing.assert_array_equal(rows[:, 0] > 0, o.is_collision(qxy, P.collision_threshold).astype(bool))
    np.testing.assert_array_equal(np.ascontiguousarray(rows[:, 1]).view(np.int32), o.range_count(qxy, P.robot_size))
    z, _, tie = o.nearest_z(qxy)
    inside = (qxy[:, 0] > pts[:, 0].min()) & (qxy[:, 0] < pts[:, 0].max()) & (qxy[:, 1] > pts[:, 1].min()) & (qxy[:, 1] < pts[:, 1].max())
    ok = inside & (tie == 0)
    np.testing.assert_array_equal(rows[ok, 2], z[ok])
    # the Morton order itself: a permutation, and spatially coherent (consecutive queries are close)
    from trg_planner_b200 import sharding
    order = sharding.morton_order(qxy, qxy.min(0), qxy.max(0))
    assert sorted(order.tolist()) == list(range(len(qxy)))
    step = np.hypot(*(qxy[order][1:] - qxy[order][:-1]).T)
    assert np.median(step) < 0.1 * np.hypot(*(qxy.max(0) - qxy.min(0)))
