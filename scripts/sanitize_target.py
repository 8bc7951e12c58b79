import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import _pkg
trg = _pkg.load()
from trg_planner_b200 import kernels as K
for P, pts, start in ((trg.MOUNTAIN, trg.terrain.mountain(90, h=0.1, seed=2), (4.5, 4.5, 0.0)),
                      (trg.INDOOR, trg.terrain.indoor(50, h=0.2, seed=1), (3.27, 4.12, 0.0))):
    t = trg.product(P); t.seed(1); t.set_global_map(pts); t.init_graph(start)
    q = trg.terrain.query_pairs(trg.terrain.bbox(pts), 40, seed=3)
    r = t.plan_batch(q)
    scan = pts[(np.abs(pts[:, 0] - start[0]) < 2) & (np.abs(pts[:, 1] - start[1]) < 2)]
    t.set_local_map(start[0], start[1], scan); t.update_graph()
    dm = K.DeviceMap(pts, 0.2)
    rng = np.random.default_rng(0)
    qq = rng.uniform(-1, 10, size=(5000, 2)).astype(np.float32)
    dm.collision(qq, 0.3, 0.16, 0.1); dm.collision(qq, 1.2, 0.16, 0.1); dm.nearest_z(qq); dm.range_count(qq, 0.6)
    dm.set_option("force_warp_path", 1); dm.collision(qq, 0.3, 0.16, 0.1)
    K.voxel_filter(pts, 0.25)
    print("ok", t.counts(), int(r["found"].sum()))
