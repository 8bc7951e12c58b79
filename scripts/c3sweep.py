import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
import _pkg; trg=_pkg.load()
P=trg.MOUNTAIN
side=int(sys.argv[1]) if len(sys.argv)>1 else 7072
pts=trg.terrain.mountain(side,h=0.1,seed=3); ext=side*0.1
for chunk,look in ((2048,256),(2048,1024),(4096,1024),(4096,2048),(8192,2048),(1024,512)):
    t=trg.product(P); t.seed(42); t.set_tuning("chunk_nodes",chunk); t.set_tuning("lookahead",look)
    t.set_global_map(pts); w=time.time(); t.init_graph((ext/2,ext/2,0.0)); dt=time.time()-w
    print(json.dumps(dict(chunk=chunk,lookahead=look,init_s=round(dt,3),**{k:t.stat(k) for k in ("us_commit","us_wait","us_sample","us_eval","us_clean","batches")})),flush=True)
    t.close()
