"""K1 (map index build) steady-state timing on a device-resident cloud, shuffled vs raster order."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K
for side, label in ((3163, "10M"), (7072, "50M"), (14143, "200M")):
    n = side * side
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ix = torch.arange(side, device="cuda", dtype=torch.float32)
    x = (ix[:, None] * 0.1).expand(side, side).reshape(-1) + (torch.rand(n, device="cuda", generator=g) - 0.5) * 0.04
    y = (ix[None, :] * 0.1).expand(side, side).reshape(-1) + (torch.rand(n, device="cuda", generator=g) - 0.5) * 0.04
    z = torch.sin(x * 0.1) + torch.cos(y * 0.13)
    pts = torch.stack([x, y, z], 1).contiguous()
    del x, y, z
    for order in ("raster", "shuffled"):
        if order == "shuffled":
            pts = pts[torch.randperm(n, device="cuda", generator=g)].contiguous()
        torch.cuda.synchronize()
        for rep in range(3):
            if rep == 1:
                K.prof_reset(); K.prof_enable(True)
            t0 = time.perf_counter()
            dm = K.DeviceMap(None, 0.2, dev_ptr=pts.data_ptr(), n=n, stride=3); dm.sync()
            dt = time.perf_counter() - t0
            dm.close()
        pr = K.prof_collect(); K.prof_enable(False)
        kms = {k: round(v["ms"] / v["launches"], 3) for k, v in pr.items()}
        tot = sum(kms.values())
        print(json.dumps(dict(points=label, order=order, wall_ms=round(dt * 1e3, 2), kernels_ms=kms, kernel_total_ms=round(tot, 3),
                              gpts_per_s=round(n / tot / 1e6, 2), alg_gbs=round(32 * n / tot / 1e6, 1), frac_hbm=round(32 * n / tot / 1e6 / 6551.7, 3))), flush=True)
    del pts
    torch.cuda.empty_cache()
