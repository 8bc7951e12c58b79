import time, torch, numpy as np
n = 10_000_000
pts = np.random.default_rng(0).uniform(0, 316, size=(n, 3)).astype(np.float32)
d = torch.from_numpy(pts).cuda()
torch.cuda.synchronize()
for rep in range(4):
    t0 = time.perf_counter()
    px = d[:, 0]
    m = px < 2.0
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pl = d[m]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    pl = pl[:, :3]
    out = torch.cat([pl, torch.zeros((pl.shape[0], 1), dtype=pl.dtype, device=pl.device)], 1)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    idx = torch.nonzero(m).squeeze(1)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    pl2 = d.index_select(0, idx)
    torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"mask {1e3*(t1-t0):.2f} index {1e3*(t2-t1):.2f} cat {1e3*(t3-t2):.2f} nonzero {1e3*(t4-t3):.2f} index_select {1e3*(t5-t4):.2f} rows {pl.shape[0]}")
