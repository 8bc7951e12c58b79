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


def allgather_rows(dist, torch, rows: np.ndarray, device) -> list[np.ndarray]:
    """All-gather one float32 matrix per rank (same column count, ragged row counts): two
    collectives (counts, then the payload padded to the longest). Returns the matrix of every rank."""
    world = dist.get_world_size()
    if hasattr(rows, "is_cuda"):
        rows_t = rows.to(device=device, dtype=torch.float32).contiguous()
    else:
        rows_t = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).to(device)
    ncol = rows_t.shape[1]
    cnt = torch.tensor([rows_t.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = [int(c) for c in torch.cat(cnts).cpu()]
    mx = max(1, max(counts))
    pad = torch.zeros((mx, ncol), dtype=torch.float32, device=device)
    if rows_t.shape[0]:
        pad[: rows_t.shape[0]] = rows_t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [out[r][: counts[r]].cpu().numpy().copy() for r in range(world)]


def stitch_tiles(dist, torch, device, rank: int, world: int, tile_pts: np.ndarray, node_pos: np.ndarray,
                 node_ids: np.ndarray, x_lo: float, x_hi: float, expand_dist: float, robot_size: float,
                 edge_eval, strip: float = 2.0):
    """Wire the TRGs of neighbouring tiles together (tiles laid out along x).

    Exchange step (NCCL / gloo all-gather, three collectives' worth of payload):
      1. boundary nodes: nodes within expand_dist + robot_size of a shared border (pos, id);
      2. boundary strips: the map points within `strip` of a shared border, so that the rank that
         evaluates a cross-tile edge sees the terrain on BOTH sides of the border;
      3. the stitched edges, so both owners of a border know them.
    Rank r owns the border between tiles r and r+1: it builds a map of the two strips, pairs its
    right-band nodes with the neighbour's left-band nodes closer than expand_dist (the radius of
    TRG::expandGraph's neighbour wiring, trg.cpp:429-444) and validates every pair with the same edge
    evaluation as inside a tile (`edge_eval(strip_pts, p1_xyz, p2_xyz) -> stage, weight, dist`;
    the batched K4 kernel on the GPU, the oracle in the CPU tests).
    Returns (edges, stats): edges = float64 array of rows (rank_a, id_a, rank_b, id_b, weight, dist)
    for EVERY border (identical on all ranks).
    """
    import time
    tm = {}
    t0 = time.perf_counter()

    def lap(name):
        nonlocal t0
        t1 = time.perf_counter()
        tm[name] = tm.get(name, 0.0) + 1e3 * (t1 - t0)
        t0 = t1

    band = expand_dist + robot_size
    x = node_pos[:, 0]
    left = np.nonzero(x < x_lo + band)[0] if rank > 0 else np.zeros(0, np.int64)
    right = np.nonzero(x > x_hi - band)[0] if rank < world - 1 else np.zeros(0, np.int64)
    sel = np.concatenate([left, right])
    side = np.concatenate([np.zeros(left.size), np.ones(right.size)]).astype(np.float32)
    ids_f = np.ascontiguousarray(node_ids[sel], np.int32).view(np.float32)   # ids travel bit-cast
    nodes_rows = np.column_stack([node_pos[sel], ids_f, side]).astype(np.float32) if sel.size else np.zeros((0, 5), np.float32)
    if hasattr(tile_pts, "is_cuda"):
        # the tile's cloud as a torch tensor (resident in HBM on the GPU box): select the strips there
        px = tile_pts[:, 0]
        parts = []
        if rank > 0:
            pl = tile_pts[px < x_lo + strip][:, :3]
            parts.append(torch.cat([pl, torch.zeros((pl.shape[0], 1), dtype=pl.dtype, device=pl.device)], 1))
        if rank < world - 1:
            pr = tile_pts[px > x_hi - strip][:, :3]
            parts.append(torch.cat([pr, torch.ones((pr.shape[0], 1), dtype=pr.dtype, device=pr.device)], 1))
        strip_rows = torch.cat(parts).to(torch.float32) if parts else torch.zeros((0, 4), dtype=torch.float32, device=device)
    else:
        # one pass over the 10 M-point cloud (combined mask, one index gather), then split the ~1 %
        # that survives; row order inside each strip stays the cloud's order
        px = tile_pts[:, 0]
        lo_cut = x_lo + strip if rank > 0 else -np.inf
        hi_cut = x_hi - strip if rank < world - 1 else np.inf
        if lo_cut > hi_cut:   # tile narrower than two strips: a point may belong to both
            pl, pr = tile_pts[px < lo_cut][:, :3], tile_pts[px > hi_cut][:, :3]
        else:
            sub = tile_pts[np.flatnonzero((px < lo_cut) | (px > hi_cut))][:, :3]
            is_left = sub[:, 0] < lo_cut
            pl, pr = sub[is_left], sub[~is_left]
        strip_rows = np.concatenate([np.column_stack([pl, np.zeros(len(pl), np.float32)]),
                                     np.column_stack([pr, np.ones(len(pr), np.float32)])]).astype(np.float32)
    lap("select_ms")
    all_nodes = allgather_rows(dist, torch, nodes_rows, device)
    lap("gather_nodes_ms")   # first collective of the step: includes waiting for the slowest rank's build
    all_strips = allgather_rows(dist, torch, strip_rows, device)
    lap("gather_strips_ms")
    mine = np.zeros((0, 6), np.float64)
    n_pairs = 0
    if rank < world - 1:
        a = all_nodes[rank][all_nodes[rank][:, 4] == 1.0]          # my right band
        b = all_nodes[rank + 1][all_nodes[rank + 1][:, 4] == 0.0]  # neighbour's left band
        pairs = cross_tile_candidates(a[:, :3], b[:, :3], expand_dist)
        n_pairs = len(pairs)
        lap("pairs_ms")
        if n_pairs:
            sp = np.concatenate([all_strips[rank][all_strips[rank][:, 3] == 1.0][:, :3],
                                 all_strips[rank + 1][all_strips[rank + 1][:, 3] == 0.0][:, :3]]).astype(np.float32)
            stage, w, d = edge_eval(np.ascontiguousarray(sp), np.ascontiguousarray(a[pairs[:, 0], :3]),
                                    np.ascontiguousarray(b[pairs[:, 1], :3]))
            lap("edge_eval_ms")
            ok = np.nonzero(np.asarray(stage) == 0)[0]
            ida = np.ascontiguousarray(a[:, 3]).view(np.int32)[pairs[ok, 0]]
            idb = np.ascontiguousarray(b[:, 3]).view(np.int32)[pairs[ok, 1]]
            mine = np.column_stack([np.full(ok.size, rank), ida, np.full(ok.size, rank + 1), idb,
                                    np.asarray(w)[ok], np.asarray(d)[ok]]).astype(np.float64)
    # share the stitched edges (float32 payload: ids < 2^24 exact; larger ids would need the bit-cast trick)
    got = allgather_rows(dist, torch, mine.astype(np.float32), device)
    lap("gather_edges_ms")
    edges = np.concatenate(got) if got else mine
    stats = dict(boundary_nodes=int(sum(len(m) for m in all_nodes)), strip_points=int(sum(len(m) for m in all_strips)),
                 candidate_pairs=int(n_pairs), stitched_edges_total=int(len(edges)),
                 bytes=int(sum(m.nbytes for m in all_nodes) + sum(m.nbytes for m in all_strips) + sum(m.nbytes for m in got)),
                 **{k: round(v, 2) for k, v in tm.items()})
    return edges, stats
