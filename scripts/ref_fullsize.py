#!/usr/bin/env python
"""Full-size CPU run of the REFERENCE ITSELF (oracle/_ref/libtrg_ref.so = unmodified trg.cpp + kdtree.c)
on a benchmark configuration, at the reference's own timer sites ("Graph expansion time"
src/planner/trg_planner.cpp:197-199, "Path planning time" :262-270), 1 thread.

Writes  profiles/_big/<tag>_ref.npz   the reference's graph / CSR / paths (git-ignored: too large for
                                      history, but it travels to the GPU box, where
                                      `bench.py --verify` compares the CUDA build with it)
        profiles/<round>_<tag>_reference_cpu.json   timings + SHA-256 digests of every array

  python scripts/ref_fullsize.py --config c2          # 3163 x 3163 = 10 M points, 1 000 queries
"""
import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402

CONFIGS = {  # tag: (side, map seed, n queries, query seed)  — SURVEY.md §8d
    "c2": (3163, 2, 1000, 7),
    "c2s": (1000, 2, 100, 7),     # the 1 M-point sample bench.py's default CPU leg uses
    "c3t": (3536, 3, 1000, 8),    # one 12.5 M-point tile of C3's generator
    "c3": (7072, 3, 200, 8),      # the whole 50 M-point C3 map (the first 200 of its 10 000 queries); written --slim
}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


C4 = dict(side=4473, map_seed=4, scan_seed=9, half=22.36, scans=7)   # = scripts/bench_extra.py run_c4 (2 warm-up + 5 timed scans)
C4_KEYS = ("iter_ids", "pos", "state", "row_ptr", "col", "dist")


def run_c4(a):
    """Config #4 at full size: the reference builds the TRG of the 20 M-point map, then takes the 200 k-point scans of
    bench.py --workload c4 one by one (setLocalMap + updateGraph, trg.cpp:195-231, 456-489). The global graph after
    every scan is recorded as SHA-256 digests; the last one also keeps its node positions and edge risks as data."""
    sys.path.insert(0, str(ROOT / "scripts"))
    from bench_extra import _scans
    trg = _pkg.load()
    P = trg.MOUNTAIN
    side = C4["side"]
    pts = trg.terrain.mountain(side, h=0.1, seed=C4["map_seed"])
    ext = side * 0.1
    o = _pkg.load_oracle().oracle(P, kind=a.kind)
    o.seed(42)
    t0 = time.perf_counter()
    o.set_global_map(pts)
    t1 = time.perf_counter()
    assert o.init_graph((ext / 2, ext / 2, 0.0)) == 0
    t2 = time.perf_counter()
    rng = np.random.default_rng(C4["scan_seed"])
    per_scan = []
    g = o.export()
    rec_scans = [{"scan": -1, "nodes": g.n_nodes, "edges": g.n_edges, "rng_draws": o.stat("rng_draws"),
                  "digests": {k: digest(getattr(g, k)) for k in C4_KEYS}}]
    for k, (cx, cy, scan) in enumerate(_scans(pts, ext, C4["scans"], C4["half"], rng, ext / 2 - 50.0)):
        s0 = time.perf_counter()
        o.set_local_map(cx, cy, scan)
        s1 = time.perf_counter()
        o.update_graph()
        s2 = time.perf_counter()
        g = o.export()
        ln, le = o.counts("local")
        per_scan.append((s1 - s0, s2 - s1))
        rec_scans.append({"scan": k, "scan_points": int(scan.shape[0]), "set_local_map_s": s1 - s0, "update_graph_s": s2 - s1,
                          "nodes": g.n_nodes, "edges": g.n_edges, "local_nodes": ln, "local_edges": le,
                          "rng_draws": o.stat("rng_draws"), "digests": {k: digest(getattr(g, k)) for k in C4_KEYS}})
        print(json.dumps(rec_scans[-1]), flush=True)
    big = ROOT / "profiles" / "_big"
    big.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(big / "c4_ref.npz", slim=1, side=side, scans=json.dumps(rec_scans), pos=g.pos, weight=g.weight)
    L = np.array(per_scan)
    T = L[2:] if len(L) > 2 else L
    rec = {"what": "the reference's own unmodified trg.cpp + kdtree.c (oracle/_ref/libtrg_ref.so) on config #4 at full size"
                   if a.kind == "ref" else f"oracle kind {a.kind}",
           "config": "c4", "side": side, "points": int(pts.shape[0]), "mt19937_seed": 42, "cores_used": 1,
           "host_cores": os.cpu_count(), "set_global_map_s": round(t1 - t0, 3), "init_graph_s": round(t2 - t1, 3),
           "per_scan_ms": {"set_local_map": float(1e3 * T[:, 0].mean()), "update_graph": float(1e3 * T[:, 1].mean()),
                           "note": "mean of scans 2..6 (the five the GPU arm times)"},
           "scans": rec_scans}
    out = ROOT / "profiles" / f"{a.round}_c4_reference_cpu.json"
    out.write_text(json.dumps(rec, indent=1))
    print(json.dumps({k: v for k, v in rec.items() if k != "scans"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["c4"])
    ap.add_argument("--round", default="r02")
    ap.add_argument("--kind", default="ref", choices=["ref", "refkd", "port"])
    ap.add_argument("--slim", action="store_true",
                    help="keep only what cannot be compared by digest (edge risks, queries, paths) + the digests of the rest")
    a = ap.parse_args()
    if a.config == "c4":
        return run_c4(a)
    side, map_seed, nq, q_seed = CONFIGS[a.config]
    trg = _pkg.load()
    P = trg.MOUNTAIN
    pts = trg.terrain.mountain(side, h=0.1, seed=map_seed, tile=(0, 0), world_tiles=(1, 1))
    bb = trg.terrain.bbox(pts)
    start = (0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)
    queries = trg.terrain.query_pairs(bb, nq, seed=q_seed)
    o = _pkg.load_oracle().oracle(P, kind=a.kind)
    o.seed(42)
    t0 = time.perf_counter()
    o.set_global_map(pts)
    t1 = time.perf_counter()
    assert o.init_graph(start) == 0
    t2 = time.perf_counter()
    g = o.export()
    plans = []
    t3 = time.perf_counter()
    for row in queries:
        plans.append(o.plan(row[:2], row[2:5], max_pts=1 << 14))
    t4 = time.perf_counter()
    n = int(pts.shape[0])
    path_ids = np.concatenate([p["ids"] for p in plans]) if plans else np.zeros(0, np.int32)
    path_off = np.cumsum([0] + [len(p["ids"]) for p in plans]).astype(np.int64)
    arrays = dict(iter_ids=g.iter_ids, pos=g.pos, state=g.state, row_ptr=g.row_ptr, col=g.col, weight=g.weight, dist=g.dist,
                  queries=queries, path_ids=path_ids, path_off=path_off,
                  path_found=np.array([p["found"] for p in plans]),
                  path_len=np.array([p["path_length"] for p in plans], np.float32),
                  path_risk=np.array([p["avg_risk"] for p in plans], np.float32),
                  direct=np.array([p["direct_dist"] for p in plans], np.float32),
                  goal_known=np.array([p["goal_known"] for p in plans]))
    big = ROOT / "profiles" / "_big"
    big.mkdir(parents=True, exist_ok=True)
    digests = {k: digest(v) for k, v in arrays.items()}
    if a.slim or a.config == "c3":
        # a 50 M-point graph is 0.4 GB: the arrays that must match bit for bit travel as SHA-256 digests, only the
        # edge risks (compared within 1e-5), the paths and the node positions (so that a nearest-map-point tie, the one
        # documented deviation, can be counted row by row) travel as data
        keep = {k: arrays[k] for k in ("pos", "weight", "queries", "path_ids", "path_off", "path_found", "path_len", "path_risk", "direct", "goal_known")}
        np.savez_compressed(big / f"{a.config}_ref.npz", seed=42, start=np.asarray(start, np.float32), side=side, map_seed=map_seed,
                            rng_draws=o.stat("rng_draws"), n_nodes=g.n_nodes, n_edges=g.n_edges, slim=1,
                            digest_names=np.array(sorted(digests)), digest_values=np.array([digests[k] for k in sorted(digests)]), **keep)
    else:
        np.savez(big / f"{a.config}_ref.npz", seed=42, start=np.asarray(start, np.float32), side=side, map_seed=map_seed,
                 rng_draws=o.stat("rng_draws"), **arrays)
    rec = {
        "what": "the reference's own unmodified trg.cpp + kdtree.c (oracle/_ref/libtrg_ref.so) on the full configuration"
                if a.kind == "ref" else f"oracle kind {a.kind}",
        "config": a.config, "side": side, "points": n, "map_seed": map_seed, "mt19937_seed": 42, "queries": nq,
        "cores_used": 1, "host_cores": os.cpu_count(), "host": os.uname().nodename,
        "set_global_map_s": round(t1 - t0, 3), "init_graph_s": round(t2 - t1, 3), "build_s": round(t2 - t0, 3),
        "plan_s": round(t4 - t3, 3),
        "points_per_sec": n / (t2 - t0), "nodes_per_sec": g.n_nodes / (t2 - t0), "paths_per_sec": nq / (t4 - t3),
        "nodes": g.n_nodes, "edges": g.n_edges, "rng_draws": o.stat("rng_draws"),
        "paths_found": int(sum(p["found"] for p in plans)),
        "mean_path_nodes": float(np.mean([len(p["ids"]) for p in plans if p["found"]] or [0])),
        "digests": digests,
    }
    out = ROOT / "profiles" / f"{a.round}_{a.config}_reference_cpu.json"
    out.write_text(json.dumps(rec, indent=1))
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
