"""Saturated timing of the sampling-window kernel: staged (shared memory) vs plain (per-thread global loads)."""
import ctypes as C, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K
P = trg.MOUNTAIN
side = 2000
pts = trg.terrain.mountain(side, h=0.1, seed=2)
ext = side * 0.1
rng = np.random.default_rng(4)
n_nodes, W = 100_000, 128
nodes = rng.uniform(1, ext - 1, size=(n_nodes, 2)).astype(np.float32)
key = np.floor(nodes[:, 1] / 2.0).astype(np.int64) * 100000 + np.floor(nodes[:, 0] / 2.0).astype(np.int64)
nodes = nodes[np.argsort(key, kind="stable")]
ang = rng.uniform(0, 2 * np.pi, 1_000_000)
draws = (P.expand_dist * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(np.float32)
first = rng.integers(0, 1_000_000 - W, n_nodes).astype(np.int32)
d_nodes, d_draws, d_first = (torch.from_numpy(a).cuda() for a in (nodes, draws, first))
L = K.lib()
L.trgb_sample_window_launch2.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
res = {}
for use_staging in (1, 0):
    dm = K.DeviceMap(pts, 0.5 * P.robot_size)
    dm.set_option("use_staging", use_staging)
    mask = torch.zeros(n_nodes * 2, dtype=torch.int64, device="cuda")
    for rep in range(4):
        if rep == 1:
            K.prof_reset(); K.prof_enable(True)
        mask.zero_(); torch.cuda.synchronize()
        L.trgb_sample_window_launch2(dm.h, C.c_void_p(d_nodes.data_ptr()), C.c_void_p(d_first.data_ptr()), C.c_void_p(d_draws.data_ptr()),
                                     n_nodes, W, C.c_float(P.expand_dist * 1.0001), P.robot_size, P.height_threshold,
                                     P.collision_threshold, C.c_void_p(mask.data_ptr()))
        dm.sync()
    pr = K.prof_collect()["k_sample_window"]; K.prof_enable(False)
    ups = pr["units"] / pr["ms"] * 1e3
    res["staged" if use_staging else "plain"] = dict(avg_ms=round(pr["ms"] / pr["launches"], 3), tests_per_s=round(ups),
                                                     alg_gbs=round(ups * 461.6 / 1e9, 1), frac=round(ups * 461.6 / 1e9 / 6551.7, 4),
                                                     checksum=int(mask.sum().item()))
print(json.dumps(res))
