"""-m gpu: the multi-GPU path end to end on ONE device: two ranks (gloo between them, both on cuda:0) each build
their tile's TRG with the CUDA engine, stitch across the border with the K4 kernels, merge into one global search
graph on the device and answer a shared query batch — the same code `bench.py --gpus N` runs over NCCL."""
import heapq
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(pkg, built):
    from trg_planner_b200 import kernels
    if kernels.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return kernels


def _merged_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    trg = _pkg.load()
    from trg_planner_b200 import kernels as K, sharding
    torch.cuda.set_device(0)
    K.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = trg.MOUNTAIN
    side = 300
    pts = trg.terrain.mountain(side, h=0.1, seed=2, tile=(rank, 0), world_tiles=(world, 1))
    bb = trg.terrain.bbox(pts)
    t = trg.product(P)
    t.seed(42)
    t.set_global_map(pts)
    assert t.init_graph((0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)) == 0
    dev = torch.device("cpu")       # gloo collectives on the host (two ranks cannot share one GPU under NCCL) ...
    mg = sharding.build_merged_graph(dist, torch, dev, rank, world, t, pts, bb, P, K, compute_device=torch.device("cuda", 0))  # ... merge on the GPU
    # starts in tile 0, goals in tile 1 (and a few the other way round): every path has to cross x = 30 m
    rng = np.random.default_rng(5)
    n_q = 48
    a = np.column_stack([rng.uniform(2, 26, n_q), rng.uniform(2, 28, n_q)])
    b = np.column_stack([rng.uniform(34, 58, n_q), rng.uniform(2, 28, n_q)])
    swap = rng.uniform(size=n_q) < 0.25
    s = np.where(swap[:, None], b, a)
    g = np.where(swap[:, None], a, b)
    queries = np.column_stack([s, g, np.zeros(n_q)]).astype(np.float32)
    res = sharding.plan_sharded(dist, torch, dev, rank, world, mg, queries, P, K)
    out = dict(rank=rank, res={k: (np.asarray(v) if not isinstance(v, int) else v) for k, v in res.items()},
               node_off=np.asarray(mg["node_off"]), stats=mg["stats"], n_nodes=mg["n_nodes"], n_edges=mg["n_edges"])
    if rank == 0:
        out.update({k: mg[k].cpu().numpy() for k in ("pos", "state", "row_ptr", "col", "weight", "dist")})
        out["queries"] = queries
    q.put(out)
    dist.barrier()
    mg["graph"].close()
    dist.destroy_process_group()


def test_merged_graph_on_gpu_paths_cross_the_tile_border(pkg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 150)
    procs = [ctx.Process(target=_merged_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r = q.get(timeout=600)
        got[r["rank"]] = r
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = got[0], got[1]
    # both ranks hold the same merged graph and the same gathered answers
    assert r0["n_nodes"] == r1["n_nodes"] and r0["n_edges"] == r1["n_edges"]
    for k in ("found", "cost", "path_length", "offsets", "ids"):
        np.testing.assert_array_equal(r0["res"][k], r1["res"][k], err_msg=k)
    res, P = r0["res"], pkg.MOUNTAIN
    n_q = len(r0["queries"])
    assert res["found"].all()
    assert int(res["cross_tile_paths"]) == n_q          # every path uses nodes of both tiles
    assert r0["stats"]["stitched_edges_total"] > 50
    row, col, w, d, state = r0["row_ptr"], r0["col"], r0["weight"], r0["dist"], r0["state"]
    sf = np.float32(P.safety_factor)
    ecost = ((sf * w).astype(np.float32) + np.float32(1)).astype(np.float32) * d
    n0 = int(r0["node_off"][1])
    crossings = 0
    for i in range(n_q):
        ids = res["ids"][res["offsets"][i]:res["offsets"][i + 1]]
        c = np.float32(0)
        for a, b in zip(ids[:-1], ids[1:]):
            k = np.nonzero(col[row[a]:row[a + 1]] == b)[0]
            assert len(k) == 1, "consecutive path nodes must be joined by an edge of the merged graph"
            c = np.float32(c + ecost[row[a] + k[0]])
            crossings += int((a < n0) != (b < n0))
        assert abs(float(c) - float(res["cost"][i])) <= 1e-5 * max(float(c), 1e-9)
    assert crossings >= n_q
    # optimality on the merged graph: plain Dijkstra (float64) from a few starts
    for i in range(0, n_q, 8):
        ids = res["ids"][res["offsets"][i]:res["offsets"][i + 1]]
        src, dst = int(ids[0]), int(ids[-1])
        dist_ = np.full(len(state), np.inf)
        dist_[src] = 0.0
        pq = [(0.0, src)]
        while pq:
            du, u = heapq.heappop(pq)
            if u == dst:
                break
            if du > dist_[u]:
                continue
            for k in range(row[u], row[u + 1]):
                v = col[k]
                if state[v] == -1:
                    continue
                nd = du + float(ecost[k])
                if nd < dist_[v]:
                    dist_[v] = nd
                    heapq.heappush(pq, (nd, v))
        assert abs(dist_[dst] - float(res["cost"][i])) <= 1e-4 * dist_[dst], (i, dist_[dst], res["cost"][i])


def _qlist_gpu_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import _pkg
    trg = _pkg.load()
    from trg_planner_b200 import kernels as K, sharding
    torch.cuda.set_device(0)
    K.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = trg.MOUNTAIN
    pts = trg.terrain.mountain(300, h=0.1, seed=2)
    rng = np.random.default_rng(4)
    n = 20000
    qxy = rng.uniform(pts[:, :2].min(0) + 0.5, pts[:, :2].max(0) - 0.5, size=(n, 2)).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n)
    d = rng.uniform(0.1, P.expand_dist, n)
    p2 = np.column_stack([qxy + np.stack([d * np.cos(ang), d * np.sin(ang)], 1), np.zeros(n)]).astype(np.float32)   # (x, y, z)
    p1 = np.column_stack([qxy, np.zeros(n, np.float32)]).astype(np.float32)

    def k2_k3(part, idx):
        dm = K.DeviceMap(part, P.robot_size)
        coll = dm.collision(qxy[idx], P.robot_size, P.height_threshold, P.collision_threshold).astype(np.float32)
        z, _, tie = dm.nearest_z(qxy[idx])
        dm.close()
        return np.column_stack([coll, z, tie.astype(np.float32)])

    def k4(part, idx):
        dm = K.DeviceMap(part, P.robot_size)
        r = dm.edge_eval(p1[idx], p2[idx], P.robot_size, P.height_threshold, P.collision_threshold)
        dm.close()
        return np.column_stack([r["stage"].astype(np.float32), r["weight"], r["dist"], r["npts"].astype(np.int32).view(np.float32)])

    cpu = torch.device("cpu")
    a, sa = sharding.sharded_query_list(dist, torch, cpu, rank, world, pts, qxy, P.robot_size + 0.3, k2_k3)
    # an edge reads its segment samples (robot_size around points up to expand_dist away) and the ellipse around its centre
    halo = P.expand_dist + float(np.hypot(0.5 * P.expand_dist, P.robot_size)) + 0.05
    b, sb = sharding.sharded_query_list(dist, torch, cpu, rank, world, pts, qxy, halo, k4)
    q.put(dict(rank=rank, a=a, b=b, sa=sa, sb=sb, pts=pts, qxy=qxy, p1=p1, p2=p2))
    dist.barrier()
    dist.destroy_process_group()


def test_query_list_sharded_over_ranks_equals_one_map(pkg, K):
    """SURVEY 8(e), config #5 style: K2 / K3 / K4 over a fixed query list, Morton-partitioned over two ranks that each
    index only the part of the map their queries reach — bit-identical to the same kernels on the whole map."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29750 + (os.getpid() % 150)
    procs = [ctx.Process(target=_qlist_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r = q.get(timeout=600)
        got[r["rank"]] = r
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = got[0], got[1]
    np.testing.assert_array_equal(r0["a"], r1["a"])
    np.testing.assert_array_equal(r0["b"], r1["b"])
    assert r0["sa"]["map_points_mine"] < 0.75 * len(r0["pts"]) and r1["sa"]["map_points_mine"] < 0.75 * len(r0["pts"])
    P = pkg.MOUNTAIN
    dm = K.DeviceMap(r0["pts"], P.robot_size)
    coll = dm.collision(r0["qxy"], P.robot_size, P.height_threshold, P.collision_threshold)
    z, _, tie = dm.nearest_z(r0["qxy"])
    ev = dm.edge_eval(r0["p1"], r0["p2"], P.robot_size, P.height_threshold, P.collision_threshold)
    np.testing.assert_array_equal(r0["a"][:, 0] > 0, coll.astype(bool))
    np.testing.assert_array_equal(r0["a"][:, 1], z)
    np.testing.assert_array_equal(r0["a"][:, 2] > 0, tie.astype(bool))
    assert (ev["stage"] == 0).mean() > 0.3            # a good share of real edges, not only rejections
    np.testing.assert_array_equal(r0["b"][:, 0].astype(np.uint8), ev["stage"])
    np.testing.assert_array_equal(r0["b"][:, 1], ev["weight"])
    np.testing.assert_array_equal(r0["b"][:, 2], ev["dist"])
    np.testing.assert_array_equal(np.ascontiguousarray(r0["b"][:, 3]).view(np.int32), ev["npts"])
