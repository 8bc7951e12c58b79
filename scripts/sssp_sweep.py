"""Dev tool: time the K7 query batch of the C2 workload for several threshold steps (TRGB_SSSP_DELTA)."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg
trg = _pkg.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 3163
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
pts = trg.terrain.mountain(side, h=0.1, seed=2)
bb = trg.terrain.bbox(pts)
t = trg.product(trg.MOUNTAIN); t.seed(42); t.set_global_map(pts)
t.init_graph((0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0))
q = trg.terrain.query_pairs(bb, nq, seed=7)
for delta in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1.5"]):
    os.environ["TRGB_SSSP_DELTA"] = delta
    best = 1e9
    for rep in range(4):
        r0 = t.stat("sssp_relaxed_edges")
        w0 = time.perf_counter(); r = t.plan_batch(q); w1 = time.perf_counter()
        best = min(best, w1 - w0)
    print(f"delta={delta} plan_batch {1e3*best:.2f} ms  relaxed/query {(t.stat('sssp_relaxed_edges')-r0)/nq:.0f} found {int(r['found'].sum())}", flush=True)
