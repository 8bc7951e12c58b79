"""BASELINE config #1 (1 M-point indoor map, config/indoor.yaml: step-3 neighbour wiring on) at full size on the CPU:
the reference itself (oracle/_ref/libtrg_ref.so) and the restated oracle (port), compared digest for digest —
graph, CSR, edge risks, draw count and the findPath node sequence. Writes profiles/r02_c1_reference_cpu.json."""
import json, sys, time, hashlib
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import _pkg
trg = _pkg.load()
P = trg.INDOOR
pts = trg.terrain.indoor(1000, h=0.2, seed=1)
start = (3.27, 4.12, 0.0); goal = (190.0, 185.0, 0.0)
F = _pkg.load_oracle()
def dg(a): return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]
out = {"config": "c1 (BASELINE configs[0]): synthetic indoor map, config/indoor.yaml parameters", "points": int(pts.shape[0]), "mt19937_seed": 42}
snaps = {}
for kind in ("ref", "port"):
    o = F.oracle(P, kind=kind); o.seed(42)
    w0 = time.perf_counter(); o.set_global_map(pts); w1 = time.perf_counter(); assert o.init_graph(start) == 0; w2 = time.perf_counter()
    g = o.export(); r = o.plan(start[:2], goal); w3 = time.perf_counter()
    snaps[kind] = (g, r)
    out[kind] = {"set_global_map_s": round(w1 - w0, 3), "init_graph_s": round(w2 - w1, 3), "plan_s": round(w3 - w2, 4), "nodes": g.n_nodes, "edges": g.n_edges,
                 "rng_draws": o.stat("rng_draws"), "path_nodes": len(r["ids"]), "path_found": bool(r["found"]),
                 "digests": {k: dg(getattr(g, k)) for k in ("iter_ids", "pos", "state", "row_ptr", "col", "weight", "dist")} | {"path_ids": dg(r["ids"])}}
out["restatement_equals_reference"] = out["ref"]["digests"] == out["port"]["digests"] and out["ref"]["rng_draws"] == out["port"]["rng_draws"]
print(json.dumps(out, indent=1))
json.dump(out, open(ROOT / 'profiles' / 'r02_c1_reference_cpu.json', 'w'), indent=1)
