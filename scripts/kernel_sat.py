"""Saturated-batch timing of the query kernels alone (config #5 style): K2 collision and K4 edge
evaluation on a resident map, CUDA events on the library's stream. Also the ncu target for the
per-kernel captures (few launches, short)."""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=2000)
ap.add_argument("--nq", type=int, default=4_000_000)
ap.add_argument("--ne", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--radius", type=float, default=0.3)
ap.add_argument("--warp", action="store_true")
ap.add_argument("--cell", type=float, default=0.5, help="map cell as a multiple of robot_size")
ap.add_argument("--sorted", action="store_true", help="sort queries by cell (spatially coherent threads)")
ap.add_argument("--lib", default=None, help="time a build variant of the kernel library (scripts/build_variants.sh)")
a = ap.parse_args()
if a.lib:
    K.KERNEL_LIB = Path(a.lib).resolve()
import torch
P = trg.MOUNTAIN
pts = trg.terrain.mountain(a.side, h=0.1, seed=2)
ext = a.side * 0.1
dm = K.DeviceMap(pts, a.cell * P.robot_size)
if a.warp:
    dm.set_option("force_warp_path", 1)
rng = np.random.default_rng(10)
q = rng.uniform(1.0, ext - 1.0, size=(a.nq, 2)).astype(np.float32)
if a.sorted:
    key = (np.floor(q[:, 1] / 0.6).astype(np.int64) * 100000 + np.floor(q[:, 0] / 0.6).astype(np.int64))
    q = q[np.argsort(key, kind="stable")]
ang = rng.uniform(0, 2 * np.pi, a.ne)
p1 = np.column_stack([q[:a.ne], np.zeros(a.ne, np.float32)]).astype(np.float32)
p2 = (q[:a.ne] + P.expand_dist * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(np.float32)
dq, dp1, dp2 = torch.from_numpy(q).cuda(), torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
out8 = torch.empty(a.nq, dtype=torch.uint8, device="cuda")
st8 = torch.empty(a.ne, dtype=torch.uint8, device="cuda")
w = torch.zeros(a.ne, dtype=torch.float32, device="cuda")
dd = torch.zeros(a.ne, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for rep in range(a.reps + 1):
    if rep == 1:
        K.prof_reset(); K.prof_enable(True)
    dm.collision_launch(dq.data_ptr(), a.nq, a.radius, P.height_threshold, P.collision_threshold, out8.data_ptr())
    dm.edge_eval_launch(dp1.data_ptr(), dp2.data_ptr(), a.ne, P.robot_size, P.height_threshold, P.collision_threshold,
                        st8.data_ptr(), w.data_ptr(), dd.data_ptr())
    dm.sync()
pr = K.prof_collect()
rho = pts.shape[0] / (ext * ext)
for k, v in pr.items():
    ups = v["units"] / v["ms"] * 1e3
    # algorithmic bytes per unit: K2 16 k(r) + 9; K4 split into its two kernels:
    # collide 16 m k(r) + 21 with m = 4 segment samples, PCA 16 k(a) + 29 with a^2 = (e/2)^2 + r^2 = 0.18
    if "edge_collide" in k:
        per = 16 * 4 * np.pi * 0.09 * rho + 21
    elif "edge_pca" in k:
        per = 16 * np.pi * 0.18 * rho + 29
    elif "edge_eval" in k:
        per = 16 * (4 * np.pi * 0.09 * rho + np.pi * 0.18 * rho) + 41
    else:
        per = 16 * np.pi * a.radius ** 2 * rho + 9
    print(json.dumps(dict(kernel=k, launches=v["launches"], avg_ms=round(v["ms"] / v["launches"], 4), units_per_s=round(ups),
                          alg_gbs=round(ups * per / 1e9, 1), frac_hbm=round(ups * per / 1e9 / 6551.7, 4))))
import hashlib
sig = hashlib.sha1(out8.cpu().numpy().tobytes() + st8.cpu().numpy().tobytes() + w.cpu().numpy().tobytes()).hexdigest()[:16]
print("collision rate", float(out8.float().mean()), "edge ok", float((st8 == 0).float().mean()), "sha1", sig, "lib", a.lib)
