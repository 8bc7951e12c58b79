import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import _pkg; trg=_pkg.load()
from trg_planner_b200 import kernels as K
pts = trg.terrain.mountain(3163, h=0.1, seed=2)
for i in range(8):
    t0=time.time(); dm = K.DeviceMap(pts, 0.15); t1=time.time(); dm.close(); t2=time.time()
    print(f"create {t1-t0:.4f}s destroy {t2-t1:.4f}s", flush=True)
import torch
d = torch.from_numpy(pts).cuda(); torch.cuda.synchronize()
for i in range(6):
    t0=time.time(); dm = K.DeviceMap(None, 0.15, dev_ptr=d.data_ptr(), n=pts.shape[0], stride=3); t1=time.time(); dm.close(); t2=time.time()
    print(f"dev create {t1-t0:.4f}s destroy {t2-t1:.4f}s", flush=True)
