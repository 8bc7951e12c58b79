import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg
trg = _pkg.load()
side = 3163
pts = trg.terrain.mountain(side, h=0.1, seed=2)
bb = trg.terrain.bbox(pts)
d = torch.from_numpy(pts).cuda()
t = trg.product(trg.MOUNTAIN); t.seed(42)
start = (0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)
for rep in range(4):
    t.seed(42)
    t.set_global_map_dev(d.data_ptr(), len(pts), 3)
    t.init_graph(start)
    w0 = time.perf_counter()
    g = t.export()
    w1 = time.perf_counter()
    torch.cuda.synchronize(); w2 = time.perf_counter()
    px = d[:, 0]
    m = px > float(bb[0][1] - 2.0)
    torch.cuda.synchronize(); w3 = time.perf_counter()
    pr = d[m][:, :3]
    torch.cuda.synchronize(); w4 = time.perf_counter()
    out = torch.cat([pr, torch.ones((pr.shape[0], 1), dtype=pr.dtype, device=pr.device)], 1)
    torch.cuda.synchronize(); w5 = time.perf_counter()
    print(f"export {1e3*(w1-w0):.1f} sync {1e3*(w2-w1):.2f} mask {1e3*(w3-w2):.2f} index {1e3*(w4-w3):.2f} cat {1e3*(w5-w4):.2f} rows {pr.shape[0]}")
