import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pts = trg.terrain.mountain(side, h=0.1, seed=2)
bb = trg.terrain.bbox(pts)
t = trg.product(trg.MOUNTAIN); t.seed(42); t.set_global_map(pts)
t.init_graph((0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0))
g = t.export()
order = {int(i): k for k, i in enumerate(g.ids)}
xy = np.ascontiguousarray(g.pos[[order[int(i)] for i in g.iter_ids], :2])
for rep in range(3):
    w0 = time.perf_counter(); lo, hi, par, ax = K.kdtree_build(xy); w1 = time.perf_counter()
    print(f"kdtree_build n={len(xy)}: {1e3*(w1-w0):.2f} ms")
depth = np.zeros(len(xy), np.int32)
for i in range(1, len(xy)):
    depth[i] = depth[par[i]] + 1
print("max depth", depth.max(), "mean depth", depth.mean())
K.prof_enable(True); K.prof_reset(); K.kdtree_build(xy); print(K.prof_collect()); K.prof_enable(False)
