"""Probe: time one TRG build (+ path batch) at a given size and print scheduler / kernel stats."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--h", type=float, default=0.1)
ap.add_argument("--kind", default="mountain")
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--window", type=int, default=0)
ap.add_argument("--cell", type=float, default=0)
ap.add_argument("--tcell", type=float, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--reuse", action="store_true")
ap.add_argument("--noq", action="store_true")
ap.add_argument("--lib", default=None, help="host library variant (e.g. built with -DTRG_FINE_TIMERS)")
a = ap.parse_args()
if a.lib:
    from trg_planner_b200 import binding as _b
    _b.PRODUCT_LIB = Path(a.lib).resolve()
P = trg.MOUNTAIN if a.kind == "mountain" else trg.INDOOR
t0 = time.time()
pts = trg.terrain.mountain(a.n, h=a.h, seed=2) if a.kind == "mountain" else trg.terrain.indoor(a.n, h=a.h, seed=1)
print(f"gen {pts.shape[0]} pts {time.time()-t0:.1f}s", flush=True)
ext = a.n * a.h
start = (ext / 2, ext / 2, 0.0) if a.kind == "mountain" else (3.27, 4.12, 0.0)
t = None
for rep in range(a.reps):
    if t is None or not a.reuse:
        t = trg.product(P)
    prev = {k: t.stat(k) for k in ("us_sample", "us_eval", "us_commit", "us_draws", "us_clean", "us_wait", "us_reset", "us_root", "us_expand_end", "us_prepare", "us_serial_feed", "batches")}
    if a.chunk: t.set_tuning("chunk_nodes", a.chunk)
    if a.window: t.set_tuning("window", a.window)
    if a.cell: t.set_tuning("map_cell_scale", a.cell)
    if a.tcell: t.set_tuning("table_cell_scale", a.tcell)
    t.seed(42)
    K.prof_reset(); K.prof_enable(True)
    w0 = time.time(); t.set_global_map(pts); w1 = time.time()
    t.init_graph(start); w2 = time.time()
    nn, ne = t.counts()
    q = trg.terrain.query_pairs(trg.terrain.bbox(pts), a.queries, seed=7)
    r = t.plan_batch(q) if not a.noq else dict(found=np.zeros(1)); w3 = time.time()
    prof = K.prof_collect(); K.prof_enable(False)
    stats = {k: t.stat(k) for k in ("pops", "rng_draws", "window_launches", "eval_launches", "flush_launches", "stalls",
                                    "window_tests", "edge_evals", "nearest_map", "batches", "node_ties", "z_ties",
                                    "us_sample", "us_eval", "us_commit", "us_draws", "us_clean", "cyc_nearest", "cyc_wire", "cyc_newnode", "cyc_nn_a", "cyc_nn_b", "cyc_nn_c", "cyc_pre", "cyc_alloc", "cyc_umap", "cyc_index", "cyc_table", "us_w_draws", "us_w_prep", "us_w_gpu")}
    print(json.dumps(dict(rep=rep, map_s=round(w1 - w0, 4), init_s=round(w2 - w1, 4), plan_s=round(w3 - w2, 4),
                          snap_s=round(t.seconds("plan_snap"), 4), nodes=nn, edges=ne, found=int(r["found"].sum()),
                          pts_per_s=round(pts.shape[0] / (w2 - w0)), nodes_per_s=round(nn / (w2 - w0)),
                          paths_per_s=round(a.queries / (w3 - w2)), stats=stats)), flush=True)
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:18s} launches={v['launches']:6d} ms={v['ms']:10.3f} avg_us={1e3*v['ms']/max(1,v['launches']):9.2f}")
    print("   delta", {k: t.stat(k) - v for k, v in prev.items()})
    if not a.reuse:
        t.close()
