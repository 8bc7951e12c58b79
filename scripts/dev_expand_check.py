#!/usr/bin/env python
"""Device-resident BFS (K9) vs the CPU oracle on small maps + timing of a larger build. GPU box only."""
import sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg
trg = _pkg.load()
F = _pkg.load_oracle()
KEYS = ("pops", "device_builds", "device_steps", "device_steps_active", "device_rounds", "device_redo_pops", "device_interrupts",
        "device_polls", "window_tests", "us_device_bfs", "us_device_edges", "us_materialize", "z_ties", "node_ties", "rng_draws",
        "device_expand_unavailable", "device_expand_capacity")

def compare(a, b):
    out = {}
    for k in ("iter_ids", "ids", "pos", "state", "row_ptr", "col", "dist"):
        x, y = getattr(a, k), getattr(b, k)
        out[k] = bool(x.shape == y.shape and np.array_equal(x, y))
    if a.weight.shape == b.weight.shape:
        rel = np.abs(a.weight - b.weight) / np.maximum(np.abs(b.weight), 1e-12)
        rel[(a.weight == 0) & (b.weight == 0)] = 0
        out["weight_bad"] = int((rel > 1e-5).sum())
    else:
        out["weight_bad"] = -1
    out["nodes"] = (a.n_nodes, b.n_nodes); out["edges"] = (a.n_edges, b.n_edges)
    return out

def case(name, P, pts, start, seed, tuning=None, repeat=1):
    t, o = trg.product(P), F.oracle(P)
    for k, v in (tuning or {}).items():
        t.set_tuning(k, v)
    t.set_global_map(pts); o.set_global_map(pts)
    for r in range(repeat):
        t.seed(seed + r); o.seed(seed + r)
        w0 = time.perf_counter(); rc = t.init_graph(start); w1 = time.perf_counter()
        assert rc == 0 and o.init_graph(start) == 0
        res = compare(t.export(), o.export())
        res["draws"] = (t.stat("rng_draws"), o.stat("rng_draws"))
        ok = all(v for k, v in res.items() if isinstance(v, bool)) and res["draws"][0] == res["draws"][1]
        print(name, tuning, "rep", r, "OK" if ok else "MISMATCH", f"{1e3*(w1-w0):.1f} ms", res, {k: t.stat(k) for k in KEYS if t.stat(k)}, flush=True)
    return t

if __name__ == "__main__":
    import os
    sm = trg.terrain.mountain(300, h=0.1, seed=2) if not os.environ.get("ONLY_BIG") else None
    if sm is not None:
        case("mountain300", trg.MOUNTAIN, sm, (15.0, 15.0, 0.0), 9)
        case("mountain300", trg.MOUNTAIN, sm, (15.0, 15.0, 0.0), 9, dict(device_expand=0))
        case("mountain300", trg.MOUNTAIN, sm, (15.0, 15.0, 0.0), 5, dict(expand_max_pops=32), repeat=3)
        case("mountain300", trg.MOUNTAIN, sm, (15.0, 15.0, 0.0), 5, dict(expand_window_words=4, expand_steps=1, expand_max_pops=4096))
        st = trg.terrain.stairs(200, h=0.1, seed=5, riser=0.10)
        case("stairs200", trg.MOUNTAIN, st, (10.0, 10.0, 0.0), 11)
    if len(sys.argv) > 1:
        side = int(sys.argv[1])
        big = trg.terrain.mountain(side, h=0.1, seed=2)
        bb = trg.terrain.bbox(big)
        start = (0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)
        t = trg.product(trg.MOUNTAIN)
        t.set_global_map(big)
        for r in range(3):
            t.seed(42)
            w0 = time.perf_counter(); t.init_graph(start); w1 = time.perf_counter()
            print("big", side, f"{1e3*(w1-w0):.1f} ms", t.counts(), {k: t.stat(k) for k in KEYS if t.stat(k)}, flush=True)
        from trg_planner_b200 import kernels as K
        K.prof_enable(True); K.prof_reset()
        t.seed(42)
        w0 = time.perf_counter(); t.init_graph(start); w1 = time.perf_counter()
        pr = K.prof_collect(); K.prof_enable(False)
        print("profiled build", f"{1e3*(w1-w0):.1f} ms")
        for k, v in sorted(pr.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"  {k:22s} launches {v['launches']:6d}  total {v['ms']:9.3f} ms  avg {1e3*v['ms']/max(v['launches'],1):8.1f} us")
