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
    the candidate stitching edges (to be validated with the batched edge kernel). Cell hash on a
    max_dist grid: a node only meets the nodes of the 3 x 3 cells around it."""
    if mine_pos.shape[0] == 0 or other_pos.shape[0] == 0:
        return np.zeros((0, 2), np.int64)
    a, b = np.asarray(mine_pos[:, :2], np.float64), np.asarray(other_pos[:, :2], np.float64)
    org = np.minimum(a.min(0), b.min(0)) - max_dist
    ca, cb = np.floor((a - org) / max_dist).astype(np.int64), np.floor((b - org) / max_dist).astype(np.int64)
    W = int(max(ca[:, 0].max(), cb[:, 0].max())) + 3
    kb = cb[:, 1] * W + cb[:, 0]
    order = np.argsort(kb, kind="stable")
    kb_sorted = kb[order]
    out_i, out_j = [], []
    idx_a = np.arange(a.shape[0])
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            key = (ca[:, 1] + dy) * W + (ca[:, 0] + dx)
            lo, hi = np.searchsorted(kb_sorted, key, "left"), np.searchsorted(kb_sorted, key, "right")
            cnt = hi - lo
            if not cnt.any():
                continue
            ii = np.repeat(idx_a, cnt)
            off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            jj = order[np.repeat(lo, cnt) + off]
            d = a[ii] - b[jj]
            keep = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) <= max_dist * max_dist
            out_i.append(ii[keep]); out_j.append(jj[keep])
    if not out_i:
        return np.zeros((0, 2), np.int64)
    pairs = np.stack([np.concatenate(out_i), np.concatenate(out_j)], 1).astype(np.int64)
    return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]


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
    lap("select_nodes_ms")
    if hasattr(tile_pts, "is_cuda"):
        # the tile's cloud as a torch tensor (resident in HBM on the GPU box): select the strips there
        px = tile_pts[:, 0]
        parts = []
        if rank > 0:
            ml = px < float(x_lo + strip)
            pl = tile_pts[ml][:, :3]
            parts.append(torch.cat([pl, torch.zeros((pl.shape[0], 1), dtype=pl.dtype, device=pl.device)], 1))
        if rank < world - 1:
            mr = px > float(x_hi - strip)
            pr = tile_pts[mr][:, :3]
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
    lap("select_strips_ms")
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
    # share the stitched edges (float32 payload; the two id columns travel bit-cast)
    pay = mine.astype(np.float32)
    if len(pay):
        pay[:, 1] = mine[:, 1].astype(np.int32).view(np.float32)
        pay[:, 3] = mine[:, 3].astype(np.int32).view(np.float32)
    got = allgather_rows(dist, torch, pay, device)
    lap("gather_edges_ms")
    edges = (np.concatenate(got) if got else pay).astype(np.float64)
    if len(edges):
        cat = np.concatenate(got).astype(np.float32)
        edges[:, 1] = np.ascontiguousarray(cat[:, 1]).view(np.int32)
        edges[:, 3] = np.ascontiguousarray(cat[:, 3]).view(np.int32)
    stats = dict(boundary_nodes=int(sum(len(m) for m in all_nodes)), strip_points=int(sum(len(m) for m in all_strips)),
                 candidate_pairs=int(n_pairs), stitched_edges_total=int(len(edges)),
                 bytes=int(sum(m.nbytes for m in all_nodes) + sum(m.nbytes for m in all_strips) + sum(m.nbytes for m in got)),
                 **{k: round(v, 2) for k, v in tm.items()})
    return edges, stats


# ---------------------------------------------------------------------------------------------------
# ONE graph across the GPUs: merge the per-tile TRGs + the stitched edges, answer a query batch on it
# ---------------------------------------------------------------------------------------------------
def allgather_concat(dist, torch, t, device):
    """All-gather of one tensor per rank with ragged first dimension; returns (concatenation, row counts)."""
    world = dist.get_world_size()
    t = t.to(device).contiguous()
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = [int(c) for c in torch.cat(cnts).cpu()]
    mx = max(1, max(counts))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    if t.shape[0]:
        pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: counts[r]] for r in range(world)]), counts


def merge_graphs(dist, torch, device, rank, world, local, stitched, compute_device=None):
    """Every rank's tile graph (`local`: GraphSnapshot, CSR by local id) + the stitched cross-tile edges
    (rows rank_a, id_a, rank_b, id_b, weight, dist) -> ONE global CSR, identical on every rank.
    Global id = node offset of the owner rank + local id; a node's edge list keeps its local order, the
    stitched edges follow. Exchange: all-gathers of positions / states and of the edge lists."""
    # the tile graphs travel as they are (CSR, local ids, 12 bytes per edge), one ragged all-gather per array
    # (packing them into one or two buffers was measured slower: the host-side packing and the slicing on the
    # device cost more than the four extra collectives; merge 12 -> 22 ms at N = 2, 16 -> 36 ms at N = 4);
    # global ids are put on after the gather
    T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt))
    pos, counts = allgather_concat(dist, torch, T(local.pos, np.float32), device)
    state, _ = allgather_concat(dist, torch, T(local.state, np.int32), device)
    deg, _ = allgather_concat(dist, torch, T(np.diff(local.row_ptr), np.int32), device)
    col_all, ecounts = allgather_concat(dist, torch, T(local.col, np.int32), device)
    w_all, _ = allgather_concat(dist, torch, T(local.weight, np.float32), device)
    d_all, _ = allgather_concat(dist, torch, T(local.dist, np.float32), device)
    node_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if compute_device is not None:   # collectives on `device` (gloo: the host), the merge itself where the graph will live
        pos, state, deg, col_all, w_all, d_all = (x.to(compute_device) for x in (pos, state, deg, col_all, w_all, d_all))
    dev = pos.device
    owner_off = torch.repeat_interleave(torch.from_numpy(node_off[:-1]).to(dev), torch.tensor(ecounts, dtype=torch.int64, device=dev))
    dst = col_all.to(torch.int64) + owner_off
    src = torch.repeat_interleave(torch.arange(int(node_off[-1]), dtype=torch.int64, device=dev), deg.to(torch.int64))
    coo_all = torch.stack([src, dst], 1)
    wd_all = torch.stack([w_all, d_all], 1)
    nbytes = int(pos.numel() * 4 + state.numel() * 4 + deg.numel() * 4 + col_all.numel() * 4 + wd_all.numel() * 4)
    if len(stitched):
        st = np.asarray(stitched, np.float64)
        ga = node_off[st[:, 0].astype(np.int64)] + st[:, 1].astype(np.int64)
        gb = node_off[st[:, 2].astype(np.int64)] + st[:, 3].astype(np.int64)
        s_coo = torch.from_numpy(np.concatenate([np.stack([ga, gb], 1), np.stack([gb, ga], 1)])).to(dev)
        s_wd = torch.from_numpy(np.concatenate([st[:, 4:6], st[:, 4:6]]).astype(np.float32)).to(dev)
        coo_all = torch.cat([coo_all, s_coo])
        wd_all = torch.cat([wd_all, s_wd])
    n = int(node_off[-1])
    order = torch.sort(coo_all[:, 0], stable=True).indices   # rows by source id; local edges first, stitched after
    col = coo_all[order, 1].to(torch.int32)
    wd_sorted = wd_all[order]
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=coo_all.device)
    row_ptr[1:] = torch.cumsum(torch.bincount(coo_all[:, 0], minlength=n), 0)
    return dict(n_nodes=n, n_edges=int(col.shape[0]), node_off=node_off, pos=pos, state=state, row_ptr=row_ptr, col=col,
                weight=wd_sorted[:, 0].contiguous(), dist=wd_sorted[:, 1].contiguous(), bytes=nbytes, edge_counts=ecounts)


def build_merged_graph(dist, torch, device, rank, world, trg_handle, cloud, bb, P, K, compute_device=None):
    """Stitch the tiles (stitch_tiles), merge them (merge_graphs) and upload the merged CSR (K7) on this rank."""
    import time
    t0 = time.perf_counter()
    g = trg_handle.export()
    t1 = time.perf_counter()

    def edge_eval(strip_pts, p1, p2):
        dm = K.DeviceMap(strip_pts, 0.67 * P.robot_size)
        r = dm.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
        dm.close()
        return r["stage"], r["weight"], r["dist"]

    stitched, st = stitch_tiles(dist, torch, device, rank, world, cloud, g.pos, g.ids, bb[0][0], bb[0][1], P.expand_dist,
                                P.robot_size, edge_eval)
    t2 = time.perf_counter()
    m = merge_graphs(dist, torch, device, rank, world, g, stitched, compute_device)
    t3 = time.perf_counter()
    if m["pos"].is_cuda:   # merged on the device: the search graph is built from those tensors where they are
        graph = K.DeviceGraph.from_device(m["row_ptr"], m["col"], m["weight"], m["dist"], m["pos"], m["state"])
    else:
        graph = K.DeviceGraph(m["row_ptr"].numpy(), m["col"].numpy(), m["weight"].numpy(), m["dist"].numpy(),
                              m["pos"].numpy(), m["state"].numpy())
    grid = K.DeviceNodeGrid(m["pos"], P.robot_size)   # nearest-node snapping on the merged graph
    t4 = time.perf_counter()
    stats = dict(st)
    stats["bytes"] = int(st["bytes"] + m["bytes"])
    stats.update(export_ms=round(1e3 * (t1 - t0), 2), stitch_ms=round(1e3 * (t2 - t1), 2), merge_ms=round(1e3 * (t3 - t2), 2),
                 upload_ms=round(1e3 * (t4 - t3), 2))
    m.update(graph=graph, grid=grid, stats=stats)
    return m


def plan_sharded(dist, torch, device, rank, world, mg, queries, P, K):
    """One global query batch on the merged graph: rank r answers the contiguous share query_shard(n, r, world)
    (start / goal = nearest node of the merged graph), the (found, cost, length, node sequence) records are
    all-gathered. Returns the full batch on every rank."""
    sl = query_shard(len(queries), rank, world)
    q = np.ascontiguousarray(queries[sl], np.float32)
    s_ids = mg["grid"].nearest(q[:, 0:2])
    g_ids = mg["grid"].nearest(q[:, 2:4])
    r = mg["graph"].sssp(s_ids, g_ids, P.safety_factor)
    lens = np.diff(r["offsets"]).astype(np.int64)
    rec = torch.from_numpy(np.stack([r["found"].astype(np.float32), r["cost"], r["path_length"], lens.astype(np.float32)], 1))
    rec_all, _ = allgather_concat(dist, torch, rec, device)
    ids_all, _ = allgather_concat(dist, torch, torch.from_numpy(r["ids"].astype(np.int32)), device)
    rec_np, ids_np = rec_all.cpu().numpy(), ids_all.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(rec_np[:, 3].astype(np.int64))])
    owner = np.searchsorted(mg["node_off"], ids_np, side="right") - 1   # tile of every path node
    cross = 0
    for i in range(len(rec_np)):
        seg = owner[offs[i]:offs[i + 1]]
        cross += int(seg.size > 0 and seg.min() != seg.max())
    return dict(found=rec_np[:, 0] > 0, cost=rec_np[:, 1], path_length=rec_np[:, 2], offsets=offs, ids=ids_np,
                cross_tile_paths=cross, gather_bytes=int(rec_all.numel() * 4 + ids_all.numel() * 4),
                d2h_bytes=int(rec_np.nbytes + ids_np.nbytes))


# ---------------------------------------------------------------------------------------------------
# Pure query kernels (K2 / K3 / K4) over a fixed query list: Morton partition of the queries, each rank
# holds only the map points its queries can reach
# ---------------------------------------------------------------------------------------------------
def morton_order(xy: np.ndarray, lo, hi) -> np.ndarray:
    """Indices that sort 2-D points along a Morton (Z-order) curve over the box [lo, hi] (16 bits per axis);
    ties keep the input order. Identical on every rank for identical input."""
    xy = np.asarray(xy, np.float64)
    span = np.maximum(np.asarray(hi, np.float64) - np.asarray(lo, np.float64), 1e-9)
    q = np.clip((xy - np.asarray(lo, np.float64)) / span * 65535.0, 0, 65535).astype(np.uint32)

    def spread(v):
        v = v & np.uint32(0xFFFF)
        v = (v | (v << np.uint32(8))) & np.uint32(0x00FF00FF)
        v = (v | (v << np.uint32(4))) & np.uint32(0x0F0F0F0F)
        v = (v | (v << np.uint32(2))) & np.uint32(0x33333333)
        v = (v | (v << np.uint32(1))) & np.uint32(0x55555555)
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint32(1))
    return np.argsort(code, kind="stable")


def sharded_query_list(dist, torch, device, rank: int, world: int, cloud: np.ndarray, anchors: np.ndarray, halo: float,
                       evaluate):
    """One batch of independent map queries split over the ranks, no communication while they run.

    anchors (n, 2): the point that locates query i (the query point for K2 / K3, the first end point for K4).
    The queries are ordered along a Morton curve and cut into `world` contiguous runs; rank r keeps only the map
    points within `halo` of the bounding box of its run (halo >= the largest radius a query of the run reads:
    robot_size for K2, the search ring for K3, edge length + ellipse semi-axis for K4), so every query sees exactly
    the points it would see on the whole map and the answers are bit-identical to the unsharded ones.
    evaluate(map_points (m, 3) float32, indices of my queries) -> float32 array (len(indices), k); integers
    travel bit-cast. The rows are all-gathered (one ragged collective) and put back in the caller's order.
    Returns (rows (n, k) float32 in query order, stats)."""
    anchors = np.ascontiguousarray(anchors, np.float32)
    n = anchors.shape[0]
    lo, hi = anchors.min(0), anchors.max(0)
    order = morton_order(anchors, lo, hi)
    sl = query_shard(n, rank, world)
    mine = order[sl]
    if mine.size:
        # the map cells my queries can reach: coarse tiles (>= the halo wide) that hold a query of mine, grown by one
        # tile. (A bounding box would not do: a run of a Z-curve that spills a few queries into the next quadrant
        # has a box twice its area.)
        ext = np.maximum(hi - lo, 1e-6)
        T = float(max(halo, float(ext.max()) / 256.0))
        org = lo - T
        W, H = int(np.floor((ext[0] + 2 * T) / T)) + 2, int(np.floor((ext[1] + 2 * T) / T)) + 2
        tq = np.floor((anchors[mine] - org) / T).astype(np.int64)
        grid = np.zeros((H, W), bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                grid[np.clip(tq[:, 1] + dy, 0, H - 1), np.clip(tq[:, 0] + dx, 0, W - 1)] = True
        if hasattr(cloud, "is_cuda"):
            g = torch.from_numpy(grid).to(cloud.device)
            tx = torch.floor((cloud[:, 0] - float(org[0])) / T).long()
            ty = torch.floor((cloud[:, 1] - float(org[1])) / T).long()
            inside = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
            m = inside & g[ty.clamp(0, H - 1), tx.clamp(0, W - 1)]
            part = cloud[m][:, :3].contiguous()
        else:
            tx = np.floor((cloud[:, 0] - org[0]) / T).astype(np.int64)
            ty = np.floor((cloud[:, 1] - org[1]) / T).astype(np.int64)
            inside = (tx >= 0) & (tx < W) & (ty >= 0) & (ty < H)
            m = inside & grid[np.clip(ty, 0, H - 1), np.clip(tx, 0, W - 1)]
            part = np.ascontiguousarray(cloud[m][:, :3], np.float32)
        rows = np.ascontiguousarray(evaluate(part, mine), np.float32)
        n_part = int(part.shape[0])
    else:
        rows, n_part = None, 0
    k = rows.shape[1] if rows is not None else 0
    kk = torch.tensor([k], dtype=torch.int64, device=device)
    dist.all_reduce(kk, op=dist.ReduceOp.MAX)   # ranks with an empty share learn the row width
    k = int(kk.item())
    if rows is None:
        rows = np.zeros((0, k), np.float32)
    got = allgather_rows(dist, torch, rows, device)
    out = np.empty((n, k), np.float32)
    for r in range(world):
        out[order[query_shard(n, r, world)]] = got[r]
    return out, dict(queries=n, mine=int(mine.size), map_points_mine=n_part, map_points_all=int(cloud.shape[0]),
                     gathered_bytes=int(sum(g.nbytes for g in got)))
