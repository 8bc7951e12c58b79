import sys, time, json, os
sys.path.insert(0, "/root/repo")
import numpy as np
import _pkg; trg=_pkg.load()
from trg_planner_b200 import kernels as K
P=trg.MOUNTAIN
side=int(sys.argv[1]) if len(sys.argv)>1 else 3163
nq=int(sys.argv[2]) if len(sys.argv)>2 else 1000
pts=trg.terrain.mountain(side,h=0.1,seed=2); ext=side*0.1
t=trg.product(P); t.seed(42); t.set_global_map(pts); t.init_graph((ext/2,ext/2,0.0))
q=trg.terrain.query_pairs(trg.terrain.bbox(pts), nq, seed=7)
for rep in range(3):
    K.prof_reset(); K.prof_enable(True)
    w=time.time(); r=t.plan_batch(q, max_total_nodes=nq*4096); dt=time.time()-w
    pr=K.prof_collect(); K.prof_enable(False)
    print(json.dumps(dict(delta=os.environ.get("TRGB_SSSP_DELTA","2"), rep=rep, plan_s=round(dt,4), snap_s=round(t.seconds("plan_snap"),4), csr_s=round(t.seconds("plan_csr"),4), tree_s=round(t.seconds("plan_tree"),4), grid_s=round(t.seconds("plan_grid"),4), sssp_ms=round(pr["k_sssp"]["ms"],2), found=int(r["found"].sum()), cost_sum=float(r["cost"].sum()))), flush=True)
