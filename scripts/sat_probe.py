"""Dev tool: K2 / K4 alone at saturating batch sizes on the resident C2 map (the `kernels_saturated` block of bench.py)."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg, bench
trg = _pkg.load()
from trg_planner_b200 import kernels as K
side = int(sys.argv[1]) if len(sys.argv) > 1 else 3163
pts = trg.terrain.mountain(side, h=0.1, seed=2)
bb = trg.terrain.bbox(pts)
P = trg.MOUNTAIN
t = trg.product(P); t.seed(42); t.set_global_map(pts)
rho = len(pts) / ((bb[0][1] - bb[0][0]) * (bb[1][1] - bb[1][0]))
peak, _ = bench.measured_peak_gbs()
r = bench.saturated_kernels(trg, K, torch, t, P, bb, rho, 0.67 * P.robot_size, peak)
for k, v in r.items():
    print(k, json.dumps(v))
