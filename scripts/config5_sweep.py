"""BASELINE config #5 — stress: 200 M-point multi-level synthetic terrain (stairs + overhang slabs),
K2 (cylinder collision test) and K4 (edge evaluation) roofline sweep vs search radius.

The cloud is generated on the GPU (torch, seeded) so that 200 M points do not have to cross PCIe:
stairs along x (0.10 m risers, 0.3 m treads, up for 20 m then down), slabs 2 m above ~10 % of the
4 m x 4 m tiles, every point jittered by +-0.2 h. Queries: uniform over the map, sorted by tile.
"""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=14143)
ap.add_argument("--nq", type=int, default=10_000_000)
ap.add_argument("--ne", type=int, default=2_000_000)
ap.add_argument("--radii", default="0.15,0.3,0.6,1.2,2.4")
a = ap.parse_args()
h = 0.1
g = torch.Generator(device="cuda"); g.manual_seed(5)
side = a.side
ext = side * h
t0 = time.time()
ix = torch.arange(side, device="cuda", dtype=torch.float32)
x = (ix[:, None] * h).expand(side, side).reshape(-1) + (torch.rand(side * side, device="cuda", generator=g) - 0.5) * 0.4 * h
y = (ix[None, :] * h).expand(side, side).reshape(-1) + (torch.rand(side * side, device="cuda", generator=g) - 0.5) * 0.4 * h
u = torch.remainder(x, 40.0)
up = torch.where(u < 20.0, u, 40.0 - u)
z = torch.floor(up / 0.3) * 0.10 + torch.randn(side * side, device="cuda", generator=g) * 0.005
tiles = int(np.ceil(ext / 4.0)) + 1
sel = torch.rand((tiles, tiles), device="cuda", generator=g) < 0.10
slab = sel[(x / 4.0).long().clamp(0, tiles - 1), (y / 4.0).long().clamp(0, tiles - 1)]
xs, ys, zs = x[slab] + 0.013, y[slab] - 0.017, z[slab] + 2.0
pts = torch.stack([torch.cat([x, xs]), torch.cat([y, ys]), torch.cat([z, zs])], 1).contiguous()
del x, y, z, u, up, slab, xs, ys, zs
n = pts.shape[0]
torch.cuda.synchronize()
print(f"generated {n} points on the GPU in {time.time()-t0:.1f}s ({pts.numel()*4/1e9:.2f} GB)", flush=True)
P = trg.MOUNTAIN
t0 = time.time()
dm = K.DeviceMap(None, 0.67 * P.robot_size, dev_ptr=pts.data_ptr(), n=n, stride=3)
dm.sync()
t_idx = time.time() - t0
mi = dm.info()
print(f"map index: {t_idx*1e3:.1f} ms  ({n/t_idx/1e9:.2f} G points/s, {32*n/t_idx/1e9:.0f} GB/s algorithmic = {32*n/t_idx/1e9/6551.7:.3f} of HBM peak); "
      f"grid {mi.grid_w}x{mi.grid_h}, {mi.device_bytes/1e9:.2f} GB", flush=True)
del pts
rng = np.random.default_rng(10)
q = rng.uniform(2.0, ext - 2.0, size=(a.nq, 2)).astype(np.float32)
key = np.floor(q[:, 1] / 0.6).astype(np.int64) * 1_000_000 + np.floor(q[:, 0] / 0.6).astype(np.int64)
q = q[np.argsort(key, kind="stable")]
dq = torch.from_numpy(q).cuda()
out8 = torch.empty(a.nq, dtype=torch.uint8, device="cuda")
ang = rng.uniform(0, 2 * np.pi, a.ne)
sub = np.sort(rng.choice(a.nq, a.ne, replace=False))
p1 = np.column_stack([q[sub], np.zeros(a.ne, np.float32)]).astype(np.float32)
p2 = (q[sub] + P.expand_dist * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(np.float32)
dp1, dp2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
st8 = torch.empty(a.ne, dtype=torch.uint8, device="cuda")
w = torch.empty(a.ne, dtype=torch.float32, device="cuda")
dd = torch.empty(a.ne, dtype=torch.float32, device="cuda")
rho = n / (ext * ext)
rows = []
for r in [float(v) for v in a.radii.split(",")]:
    nq = a.nq if r <= 0.6 else a.nq // 8          # large radii: fewer queries, same order of work
    ne = a.ne if r <= 0.6 else a.ne // 8
    for rep in range(3):
        if rep == 1:
            K.prof_reset(); K.prof_enable(True)
        dm.collision_launch(dq.data_ptr(), nq, r, P.height_threshold, P.collision_threshold, out8.data_ptr())
        dm.edge_eval_launch(dp1.data_ptr(), dp2.data_ptr(), ne, r, P.height_threshold, P.collision_threshold,
                            st8.data_ptr(), w.data_ptr(), dd.data_ptr())
        dm.sync()
    pr = K.prof_collect(); K.prof_enable(False)
    k_r = np.pi * r * r * rho
    row = {"radius": r, "pts_in_cylinder": round(k_r, 1), "collision_rate": round(float(out8[:nq].float().mean()), 4)}
    for name, v in pr.items():
        ups = v["units"] / v["ms"] * 1e3
        if "collision" in name:
            per = 16 * k_r + 9
            row["K2"] = dict(kernel=name, queries_per_s=round(ups), alg_gbs=round(ups * per / 1e9, 1), frac_hbm=round(ups * per / 1e9 / 6551.7, 3))
        else:
            row.setdefault("K4", {})[name] = dict(edges_per_s=round(ups), ms=round(v["ms"] / v["launches"], 3))
    if "K4" in row:
        ms = sum(v["ms"] for v in row["K4"].values())
        e = P.expand_dist; m = int(np.ceil(e / (0.5 * r))); c = 0.5 * e
        aa = np.sqrt(c * c + r * r) if c >= r else r
        per = 16 * (m * k_r + np.pi * aa * aa * rho) + 41
        row["K4_total"] = dict(edges_per_s=round(ne / ms * 1e3), alg_gbs=round(ne / ms * 1e3 * per / 1e9, 1),
                               frac_hbm=round(ne / ms * 1e3 * per / 1e9 / 6551.7, 3))
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(dict(points=n, index_ms=t_idx * 1e3, rows=rows), open("gpurun_out/config5.json", "w"), indent=1)
