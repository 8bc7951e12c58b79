"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF: oracle/_ref/libtrg_ref.so is the
reference's unmodified src/graph/trg.cpp + src/kdtree/kdtree.c, compiled where they lie by
oracle/Makefile against the stand-in headers of oracle/shim/ (Eigen / PCL / OpenCV / yaml-cpp are
absent from the image; only Eigen's JacobiSVD and float reductions are restated there).

The reference ships no golden vectors and no tests, so these fixtures are outputs of the reference
run here on seeded synthetic maps (graph, CSR, draw count, kernel-level answers, paths). The three
fields the reference cannot report (wireEdge's return stage, the number of PCA points, the float64
twin of the weight) come from the restated oracle, which `PinnedOracle.edge_eval` first requires
to agree bit for bit with the reference on every created edge. Run in the build container:
    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

trg = _pkg.load()
OUT = Path(__file__).resolve().parent


def case(name, P, pts, start, seed, n_q=24, updates=None):
    o = _pkg.load_oracle().oracle(P, kind="ref")
    o.seed(seed)
    o.set_global_map(pts)
    assert o.init_graph(start) == 0
    g = o.export()
    rng = np.random.default_rng(123)
    lo, hi = pts[:, :2].min(0) - 0.5, pts[:, :2].max(0) + 0.5
    q = rng.uniform(lo, hi, size=(4000, 2)).astype(np.float32)
    coll = o.is_collision(q, P.collision_threshold)
    cnt = o.range_count(q, P.robot_size)
    z, idx, tie = o.nearest_z(q)
    a = q[:1500]
    ang = rng.uniform(0, 2 * np.pi, a.shape[0])
    d = rng.uniform(0.1, 1.4 * P.expand_dist, a.shape[0])
    b = (a + np.stack([d * np.cos(ang), d * np.sin(ang)], 1)).astype(np.float32)
    za, _, _ = o.nearest_z(a)
    zb, _, _ = o.nearest_z(b)
    p1 = np.column_stack([a, za]).astype(np.float32)
    p2 = np.column_stack([b, zb]).astype(np.float32)
    ev = o.edge_eval(p1, p2)
    qs = trg.terrain.query_pairs(((float(lo[0]), float(hi[0])), (float(lo[1]), float(hi[1]))), n_q, seed=7)
    plans = [o.plan(r[:2], r[2:5]) for r in qs]
    path_ids = np.concatenate([p["ids"] for p in plans]) if plans else np.zeros(0, np.int32)
    path_off = np.cumsum([0] + [len(p["ids"]) for p in plans]).astype(np.int64)
    np.savez_compressed(
        OUT / f"{name}.npz", pts=pts, params=np.asarray([P.expand_dist, P.robot_size, P.sample_num, P.height_threshold, P.collision_threshold, P.update_collision_threshold, P.safety_factor, P.goal_tolerance], np.float64), seed=seed, start=np.asarray(start, np.float32), rng_draws=o.stat("rng_draws"),
        iter_ids=g.iter_ids, pos=g.pos, state=g.state, row_ptr=g.row_ptr, col=g.col, weight=g.weight, dist=g.dist,
        q=q, coll=coll, cnt=cnt, z=z, idx=idx, tie=tie, p1=p1, p2=p2, ev_stage=ev["stage"], ev_weight=ev["weight"],
        ev_weight64=ev["weight64"], ev_dist=ev["dist"], ev_npts=ev["npts"], queries=qs, path_ids=path_ids,
        path_off=path_off, path_found=np.array([p["found"] for p in plans]),
        path_len=np.array([p["path_length"] for p in plans], np.float32),
        path_risk=np.array([p["avg_risk"] for p in plans], np.float32),
        goal_known=np.array([p["goal_known"] for p in plans]))
    print(name, "nodes", g.n_nodes, "edges", g.n_edges, "draws", o.stat("rng_draws"))


def update_scans(pts, n_steps=3):
    """The scan sequence of the incremental-update cases (same recipe as tests/test_trg_gpu.py)."""
    rng = np.random.default_rng(9)
    out = []
    for step in range(n_steps):
        cx, cy = 4.0 + 3.0 * step, 6.0
        m = (np.abs(pts[:, 0] - cx) < 3.5) & (np.abs(pts[:, 1] - cy) < 3.5)
        scan = pts[m].copy()
        scan[:, 2] += rng.normal(0, 0.01, size=scan.shape[0]).astype(np.float32)
        if step == 1:   # an obstacle appears in the scan: a 1 m block 0.8 m high
            blk = (np.abs(scan[:, 0] - cx - 1.5) < 0.5) & (np.abs(scan[:, 1] - cy) < 0.5)
            scan[blk, 2] += np.where(rng.uniform(size=int(blk.sum())) < 0.5, 0.8, 0.0).astype(np.float32)
        out.append((cx, cy, scan))
    return out


def update_case(name, P, pts, start, seed):
    """setLocalMap + updateGraph (trg.cpp:195-231, 456-489): the graph after every scan."""
    o = _pkg.load_oracle().oracle(P, kind="ref")
    o.seed(seed)
    o.set_global_map(pts)
    assert o.init_graph(start) == 0
    rec = dict(pts=pts, params=np.asarray([P.expand_dist, P.robot_size, P.sample_num, P.height_threshold, P.collision_threshold, P.update_collision_threshold, P.safety_factor, P.goal_tolerance], np.float64),
               seed=seed, start=np.asarray(start, np.float32))
    for i, (cx, cy, scan) in enumerate(update_scans(pts)):
        o.set_local_map(cx, cy, scan)
        rec[f"local_before_{i}"] = o.export("local").iter_ids
        o.update_graph()
        g = o.export()
        for k in ("iter_ids", "pos", "state", "row_ptr", "col", "weight", "dist"):
            rec[f"{k}_{i}"] = getattr(g, k)
        rec[f"local_after_{i}"] = o.export("local").iter_ids
        rec[f"draws_{i}"] = o.stat("rng_draws")
    fr = np.random.default_rng(5).uniform(1, 11, size=(300, 2)).astype(np.float32)
    rec["frontier_q"] = fr
    rec["frontier"] = o.is_frontier(fr)
    np.savez_compressed(OUT / f"{name}.npz", **rec)
    print(name, "nodes after last scan", len(rec["iter_ids_2"]), "draws", rec["draws_2"])


if __name__ == "__main__":
    update_case("update_120", trg.MOUNTAIN, trg.terrain.mountain(120, h=0.1, seed=4), (4.0, 6.0, 0.0), 3)
    case("mountain_120", trg.MOUNTAIN, trg.terrain.mountain(120, h=0.1, seed=2), (6.0, 6.0, 0.0), 42)
    case("indoor_70", trg.INDOOR, trg.terrain.indoor(70, h=0.2, seed=1), (3.27, 4.12, 0.0), 42)
    case("stairs_100", trg.MOUNTAIN, trg.terrain.stairs(100, h=0.1, seed=5, riser=0.10), (5.0, 5.0, 0.0), 11)
