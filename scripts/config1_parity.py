"""BASELINE config #1 at full size: TRG build + findPath on a synthetic 1 M-point indoor map with
config/indoor.yaml parameters — the case the CPU reference can run. Product (B200) vs CPU oracle
(restated trg.cpp + verbatim reference kdtree.c): graphs compared bit for bit, both timed."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import _pkg
trg = _pkg.load()
P = trg.INDOOR
pts = trg.terrain.indoor(1000, h=0.2, seed=1)
start = (3.27, 4.12, 0.0)
goal = (190.0, 185.0, 0.0)
res = {"points": int(pts.shape[0])}
t = trg.product(P); t.seed(42)
w0 = time.time(); t.set_global_map(pts); w1 = time.time(); t.init_graph(start); w2 = time.time()
t2 = trg.product(P); t2.seed(42)   # second build: warm CUDA context
w0 = time.time(); t2.set_global_map(pts); w1 = time.time(); t2.init_graph(start); w2 = time.time()
a = t2.export()
rp = t2.plan(start[:2], goal)
res["b200"] = dict(map_s=round(w1 - w0, 4), init_s=round(w2 - w1, 4), nodes=a.n_nodes, edges=a.n_edges,
                   points_per_s=round(pts.shape[0] / (w2 - w0)), nodes_per_s=round(a.n_nodes / (w2 - w0)),
                   stalls=t2.stat("stalls"), flush_launches=t2.stat("flush_launches"), batches=t2.stat("batches"),
                   path_nodes=len(rp["ids"]), path_found=rp["found"])
refkd = (Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "liboracle_refkd.so").exists()
o = _pkg.load_oracle().oracle(P, ref_kdtree=refkd); o.seed(42)
w0 = time.time(); o.set_global_map(pts); w1 = time.time(); o.init_graph(start); w2 = time.time()
b = o.export()
w3 = time.time(); ro = o.plan(start[:2], goal); w4 = time.time()
res["cpu_oracle"] = dict(map_s=round(w1 - w0, 3), init_s=round(w2 - w1, 3), nodes=b.n_nodes, edges=b.n_edges,
                         points_per_s=round(pts.shape[0] / (w2 - w0)), nodes_per_s=round(b.n_nodes / (w2 - w0)),
                         plan_s=round(w4 - w3, 4), kdtree="verbatim reference kdtree.c" if refkd else "port")
same = {k: bool(np.array_equal(getattr(a, k), getattr(b, k))) for k in ("iter_ids", "pos", "state", "row_ptr", "col", "dist")}
rel = np.abs(a.weight - b.weight) / np.maximum(np.abs(b.weight), 1e-12) if a.n_edges == b.n_edges else np.array([np.inf])
rel[(a.weight == 0) & (b.weight == 0)] = 0
res["parity"] = dict(bit_exact=same, rng_draws_equal=t2.stat("rng_draws") == o.stat("rng_draws"),
                     weight_rel_gt_1e5=int((rel > 1e-5).sum()), weight_max_rel=float(rel.max()),
                     path_ids_equal=bool(np.array_equal(rp["ids"], ro["ids"])), path_len=(rp["path_length"], ro["path_length"]))
res["speedup_build"] = round((res["cpu_oracle"]["map_s"] + res["cpu_oracle"]["init_s"]) / (res["b200"]["map_s"] + res["b200"]["init_s"]), 1)
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/config1.json", "w"), indent=1)
