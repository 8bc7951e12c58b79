"""Multi-GPU sharding of the TRG hot path: one process per GPU, torch.distributed
for the plumbing (NCCL on the GPU box, gloo in the CPU tests).

The graph build does not shard bit-exactly (one BFS, one RNG stream), so large maps are cut into
tiles laid out along x, one per rank: every rank builds the TRG of its own tile with its own root
(weak scaling, no collective on the data path). The one real exchange step is the all-gather of
*boundary nodes* — nodes within `band` of a shared tile border — which is what a neighbour needs
to stitch cross-tile edges. Path-query batches are split contiguously by rank.
"""
from __future__ import annotations

import numpy as np


def query_shard(n_queries: int, rank: int, world: int) -> slice:
    """Contiguous shard of a query batch for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def boundary_nodes(pos: np.ndarray, x_lo: float, x_hi: float, band: float, rank: int, world: int) -> np.ndarray:
    """Indices of the nodes within `band` of a border this tile shares with a neighbour."""
    x = pos[:, 0]
    sel = np.zeros(pos.shape[0], bool)
    if rank > 0:
        sel |= x < x_lo + band
    if rank < world - 1:
        sel |= x > x_hi - band
    return np.nonzero(sel)[0]


def allgather_boundary(dist, torch, pos_sel: np.ndarray, ids_sel: np.ndarray, device) -> list[tuple[np.ndarray, np.ndarray]]:
    """All-gather (positions, local ids) of every rank's boundary nodes. Variable sizes are padded
    to the maximum count (two collectives: counts, then payload). Returns one (pos, ids) per rank."""
    world = dist.get_world_size()
    cnt = torch.tensor([pos_sel.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    mx = max(1, max(counts))
    pad = torch.zeros((mx, 4), dtype=torch.float32, device=device)
    if pos_sel.shape[0]:
        pad[: pos_sel.shape[0], :3] = torch.from_numpy(np.ascontiguousarray(pos_sel, np.float32)).to(device)
        # ids < 2^24 are exact in float32; larger ids travel bit-cast
        pad[: pos_sel.shape[0], 3] = torch.from_numpy(np.ascontiguousarray(ids_sel, np.int32)).to(device).view(torch.float32)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    res = []
    for r in range(world):
        a = out[r][: counts[r]].cpu()
        res.append((a[:, :3].numpy().copy(), a[:, 3].contiguous().view(torch.int32).numpy().copy()))
    return res


def cross_tile_candidates(mine_pos: np.ndarray, other_pos: np.ndarray, max_dist: float):
    """Pairs (i, j) of boundary nodes from two neighbouring tiles closer than `max_dist` in 2-D:
    the candidate stitching edges (to be validated with the batched edge kernel)."""
    if mine_pos.shape[0] == 0 or other_pos.shape[0] == 0:
        return np.zeros((0, 2), np.int64)
    from scipy.spatial import cKDTree
    t = cKDTree(other_pos[:, :2])
    pairs = []
    for i, nb in enumerate(t.query_ball_point(mine_pos[:, :2], max_dist)):
        for j in nb:
            pairs.append((i, j))
    return np.asarray(pairs, np.int64).reshape(-1, 2)
