#!/usr/bin/env python
"""bench.py — TRG build + risk-aware path batch on synthetic terrain (BASELINE.json configs[1]).

One "step" = one full pass of the hot path over one map: map-index build (TRG::setGlobalMap),
graph construction (TRG::initGraph) and a 1k-query planSafePath batch, all through the
reference-facing C facade (include/trg_b200.h).

  value   points/s of the build with the cloud already resident in HBM (trg_set_global_map_dev);
          the per-kernel timings behind `roofline` / `kernels` come from one more step of the same
          workload run with the library's event profiler switched on, right after the timed region
  e2e     the same through host buffers: pinned-host cloud -> H2D inside the timed region,
          paths / costs copied back to host
  N > 1   launched by torchrun: one process per GPU, each rank builds the TRG of its own
          10 M-point tile of one continuous heightfield and answers its shard of the queries
          (weak scaling, no data-path collective); boundary nodes are all-gathered over NCCL
          for stitching, and that exchange step counts as part of the build (build time at
          N > 1 = tile build + stitching). Times are max over ranks.

`--impl reference` times the CPU oracle (restated trg.cpp + the reference's own kdtree.c when
oracle/_ref was built) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED_RNG = 42          # mt19937 seed of TRG::gen_ (the reference seeds from random_device)
SEED_MAP = 2
SEED_QUERIES = 7
CPU_SAMPLE_SIDE = 1000  # bounded CPU sample: 1000 x 1000 lattice = 1 M points of the same generator


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 200):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tile_map(trg, side, rank, world):
    """Rank's tile of one continuous world heightfield (tiles laid out along x)."""
    return trg.terrain.mountain(side, h=0.1, seed=SEED_MAP, tile=(rank, 0), world_tiles=(world, 1))


# algorithmic bytes per work unit (DESIGN.md §5 / SURVEY.md §8d); rho = map points per m^2
def unit_bytes(kernel: str, P, rho: float, cell: float) -> float:
    r = P.robot_size
    k_r = np.pi * r * r * rho
    if kernel in ("k_sample_window", "k_collision"):
        return 16.0 * k_r + 8 + 1
    if kernel == "k_range_count":
        return 16.0 * k_r + 8 + 4
    if kernel == "k_nearest_z":
        return 16.0 * (9.0 * cell * cell * rho) + 12
    # K4 = segment collision samples (k_edge_collide) + ellipse PCA (k_edge_pca); the two halves of
    # SURVEY.md's 16*(m*k(r) + k_e) + 32 + 9 figure, each reading the two endpoints (20 B)
    e = P.expand_dist
    m = int(np.ceil(e / (0.5 * r)))
    c = 0.5 * e
    a = np.sqrt(c * c + r * r) if c >= r else r
    if kernel == "k_edge_collide":
        return 16.0 * m * k_r + 20 + 1
    if kernel == "k_edge_pca":
        return 16.0 * np.pi * a * a * rho + 20 + 9
    if kernel in ("k_edge_eval", "k_edge_eval_warp"):
        return 16.0 * (m * k_r + np.pi * a * a * rho) + 32 + 9
    if kernel in ("k_nodes_nearest",):
        return 0.0
    if kernel in ("k_bbox", "k_count"):
        return 12.0
    if kernel in ("k_scatter", "k_sort_cell"):
        return 32.0
    if kernel == "k_sssp":
        return 0.0  # reported per relaxed edge elsewhere
    return 0.0


def run_product(a):
    import torch
    import _pkg
    trg = _pkg.load()
    from trg_planner_b200 import kernels as K

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if K.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local)
    K.set_device(local)
    dist = None
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; the driver wants
        # exactly one JSON line there, so stdout points at stderr until the result is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    P = trg.MOUNTAIN
    pts = tile_map(trg, a.side, rank, world)
    n = int(pts.shape[0])
    bb = trg.terrain.bbox(pts)
    start = (0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)
    queries = trg.terrain.query_pairs(bb, a.queries, seed=SEED_QUERIES + rank)
    d_pts = torch.from_numpy(pts).cuda()
    h_pin = torch.from_numpy(pts).pin_memory()
    h_np = h_pin.numpy()
    t = trg.product(P)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    last_stitch = {}
    resident_leg = [True]

    def exchange_boundary():
        """N > 1: stitch the per-tile TRGs. NCCL all-gathers of boundary nodes, boundary strips of
        the map and the stitched edges; cross-tile candidate edges are validated by the K4 kernels
        on a map of the two strips (trg-planner_b200/sharding.py). Returns bytes gathered."""
        if dist is None:
            return 0
        from trg_planner_b200 import sharding
        tx = time.perf_counter()
        g = t.export(edges=False)
        last_stitch["export_ms"] = round(1e3 * (time.perf_counter() - tx), 2)

        def edge_eval(strip_pts, p1, p2):
            dm = K.DeviceMap(strip_pts, 0.67 * P.robot_size)
            r = dm.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
            dm.close()
            return r["stage"], r["weight"], r["dist"]

        cloud = d_pts if resident_leg[0] else pts   # strips are cut where the step's input lives
        _, st = sharding.stitch_tiles(dist, torch, torch.device("cuda", local), rank, world, cloud, g.pos, g.ids,
                                      bb[0][0], bb[0][1], P.expand_dist, P.robot_size, edge_eval)
        last_stitch.update(st)
        return st["bytes"]

    STAT_KEYS = ("us_sample", "us_eval", "us_commit", "us_wait", "us_clean", "us_draws", "pops", "window_launches",
                 "eval_launches", "window_tests", "edge_evals", "us_device_bfs", "us_device_edges", "us_materialize",
                 "device_steps", "device_rounds", "device_redo_pops", "device_interrupts", "device_builds")

    def one_step(resident: bool):
        resident_leg[0] = resident
        t.seed(SEED_RNG)
        s0 = {k: t.stat(k) for k in STAT_KEYS}
        w0 = time.perf_counter()
        if resident:
            t.set_global_map_dev(d_pts.data_ptr(), n, 3)
        else:
            t.set_global_map(h_np)
        rc = t.init_graph(start)
        assert rc == 0
        w1 = time.perf_counter()
        xb = exchange_boundary()
        w2 = time.perf_counter()
        r = t.plan_batch(queries)
        w3 = time.perf_counter()
        nn, ne = t.counts()
        host = {k: t.stat(k) - s0[k] for k in STAT_KEYS}
        host.update(map_ms=round(1e3 * t.seconds("set_global_map"), 2), init_ms=round(1e3 * t.seconds("init_graph"), 2),
                    snap_ms=round(1e3 * t.seconds("plan_snap"), 2), plan_ms=round(1e3 * t.seconds("plan_batch"), 2))
        return dict(host=host, build_s=w1 - w0, exch_s=w2 - w1, query_s=w3 - w2, step_s=w3 - w0, nodes=nn, edges=ne,
                    found=int(r["found"].sum()), d2h=int(r["ids"].nbytes + 4 * 4 * a.queries + 2 * a.queries +
                                                         8 * (a.queries + 1)), xbytes=xb)

    def timed(resident: bool, warmup: int, steps: int):
        # nvidia-smi is started BEFORE the warm-up: its start-up (NVML init) holds driver locks for
        # a second or so and stalled whichever CUDA call of the first timed step ran into it
        sampler = ClockSampler(local, a.clock_ms).start() if (rank == 0 and a.clock_ms > 0) else None
        for i in range(warmup):
            one_step(resident)
        sync_all()
        if sampler:
            sampler.rows.clear()   # keep only samples taken during the timed region
        l0 = K.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rows = [one_step(resident) for _ in range(steps)]
        e1.record()
        sync_all()
        total_ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = K.launch_count() - l0
        for r in rows:
            r["buildx_s"] = r["build_s"] + r["exch_s"]   # N > 1: a tile's TRG is finished when it is stitched
        agg = {k: float(np.mean([r[k] for r in rows])) for k in ("build_s", "buildx_s", "exch_s", "query_s", "step_s")}
        agg.update(total_ms=total_ms, rows=rows, clocks=clocks, launches=launches)
        return agg

    def profiled_steps(n_steps: int):
        """Per-kernel CUDA-event timings (the library's profiler: two events around every launch, on
        the launching stream) of `n_steps` further steps of the same resident workload, run right
        after the timed region. They are NOT taken inside it: with ~28 k events per step the
        profiler's bookkeeping stalled the first timed map build by 0.3 - 1.1 s on some boxes (at
        any N), which is a property of the measurement, not of the path. One profiled step is run
        and discarded first (event pool, first use)."""
        K.prof_enable(True)
        one_step(True)
        K.prof_reset()
        for _ in range(n_steps):
            one_step(True)
        sync_all()
        profd = K.prof_collect()
        K.prof_enable(False)
        return profd

    val = timed(True, a.warmup, a.steps)
    val["prof"] = profiled_steps(1)
    val["prof_steps"] = 1
    e2e = timed(False, a.warmup, a.steps)

    # max over ranks for times, sum for units
    def reduce(x, op):
        if dist is None:
            return x
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=op)
        return float(tt.item())

    RMAX = dist.ReduceOp.MAX if dist is not None else None
    RSUM = dist.ReduceOp.SUM if dist is not None else None
    tot_pts = reduce(float(n), RSUM)
    tot_nodes = reduce(float(val["rows"][-1]["nodes"]), RSUM)
    tot_q = float(a.queries * world)
    v_build = reduce(val["buildx_s"], RMAX)
    v_query = reduce(val["query_s"], RMAX)
    v_step = reduce(val["total_ms"] / a.steps, RMAX)
    e_build = reduce(e2e["buildx_s"], RMAX)
    e_query = reduce(e2e["query_s"], RMAX)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel inside the timed region -------------------------------
    rho = n / ((bb[0][1] - bb[0][0]) * (bb[1][1] - bb[1][0]))
    cell = 0.67 * P.robot_size
    peak, peak_src = measured_peak_gbs()
    prof = {k: v for k, v in val["prof"].items() if v["launches"] > 0}
    dom = max((k for k in prof if unit_bytes(k, P, rho, cell) > 0), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ub = unit_bytes(dom, P, rho, cell)
    avg_ms = d["ms"] / d["launches"]
    bytes_per_launch = ub * d["units"] / d["launches"]
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(dom)
        except Exception:
            traffic = None
    psteps = val["prof_steps"]   # steps the kernel profile covers (run right after the timed region)
    kern = {k: dict(launches=int(v["launches"] / psteps), ms_per_step=round(v["ms"] / psteps, 3),
                    units_per_step=int(v["units"] / psteps),
                    achieved_gbs=round(unit_bytes(k, P, rho, cell) * v["units"] / max(v["ms"], 1e-9) / 1e6, 1))
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- saturated-batch kernel numbers (config #5 style, isolated) ----------------------------
    sat = saturated_kernels(trg, K, torch, t, P, bb, rho, cell, peak) if not a.no_sat else None

    # ---- CPU baseline: oracle on a bounded sample, rank 0, N = 1 only ---------------------------
    cpu = None
    if world == 1 and not a.no_cpu:
        cpu = cpu_sample(trg, a.cpu_side, queries_n=100)

    line = {
        "metric": "trg_build_points_per_sec", "value": tot_pts / v_build, "unit": "points/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": v_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"C2 synthetic mountain heightfield {a.side}x{a.side} (h=0.1 m, {n} points per GPU), "
                               f"config/mountain.yaml params, full TRG build + {a.queries} path queries per GPU",
                   "points_per_gpu": n, "queries_per_gpu": a.queries, "mt19937_seed": SEED_RNG,
                   "l2": "inputs larger than L2 (160 MB float4 cloud per build); every step rebuilds from scratch",
                   "parallelism": f"tiles x{world}" if world > 1 else "single GPU"},
        "nodes_per_sec": tot_nodes / v_build, "paths_per_sec": tot_q / v_query,
        "build_ms": 1e3 * v_build, "query_ms": 1e3 * v_query,
        "graph": {"nodes": int(tot_nodes), "edges_rank0": val["rows"][-1]["edges"], "paths_found_rank0": val["rows"][-1]["found"]},
        "e2e": {"value": tot_pts / e_build, "unit": "points/s",
                "h2d_bytes_per_step": int(n * 12 + queries.nbytes), "d2h_bytes_per_step": e2e["rows"][-1]["d2h"],
                "build_ms": 1e3 * e_build, "query_ms": 1e3 * e_query, "paths_per_sec": tot_q / e_query,
                "nodes_per_sec": tot_nodes / e_build},
        "gpu_launches": int(val["launches"]),
        "host_breakdown_per_step": {"value_leg": [r["host"] for r in val["rows"]], "e2e_leg": [r["host"] for r in e2e["rows"]]},
        "clocks": val["clocks"],
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "bytes_per_unit": ub, "units_per_launch": d["units"] / d["launches"],
                     "avg_launch_ms": avg_ms, "share_of_step": d["ms"] / psteps / v_step,
                     "profiled_steps": psteps,
                     "profiled": "same workload, extra step(s) right after the timed region (see bench.py: profiled_steps)"},
        "kernels": kern,
        "kernels_saturated": sat,
        "cpu_baseline": cpu,
    }
    if world > 1:
        line["exchange"] = {"allgather_bytes_per_step": val["rows"][-1]["xbytes"], "ms": 1e3 * val["exch_s"], **last_stitch}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def saturated_kernels(trg, K, torch, t, P, bb, rho, cell, peak):
    """K2 / K4 alone at saturating batch sizes on the resident map (device pointers, CUDA events on
    the library's own stream via the built-in profiler)."""
    import ctypes as C
    rng = np.random.default_rng(10)
    nq = 4_000_000
    q = np.stack([rng.uniform(bb[0][0], bb[0][1], nq), rng.uniform(bb[1][0], bb[1][1], nq)], 1).astype(np.float32)
    key = np.floor((q[:, 1] - bb[1][0]) / 0.6).astype(np.int64) * 1_000_000 + np.floor((q[:, 0] - bb[0][0]) / 0.6).astype(np.int64)
    q = q[np.argsort(key, kind="stable")]
    ne = 1_000_000
    ang = rng.uniform(0, 2 * np.pi, ne)
    p1 = np.column_stack([q[:ne], np.zeros(ne, np.float32)]).astype(np.float32)
    p2 = (q[:ne] + P.expand_dist * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(np.float32)
    dq, dp1, dp2 = torch.from_numpy(q).cuda(), torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
    out8 = torch.empty(nq, dtype=torch.uint8, device="cuda")
    st8 = torch.empty(ne, dtype=torch.uint8, device="cuda")
    w = torch.empty(ne, dtype=torch.float32, device="cuda")
    dd = torch.empty(ne, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    L = K.lib()
    tl = C.CDLL(str(ROOT / "trg-planner_b200" / "lib" / "libtrg_b200.so"), mode=C.RTLD_GLOBAL)
    tl.trg_device_map.restype = C.c_void_p
    tl.trg_device_map.argtypes = [C.c_void_p, C.c_char_p]
    m = C.c_void_p(tl.trg_device_map(t.h, b"global"))
    prm = K.EdgeParams(P.robot_size, P.height_threshold, P.collision_threshold, 0)
    res = {}
    for rep in range(4):
        if rep == 1:
            K.prof_reset(); K.prof_enable(True)
        L.trgb_collision_launch(m, C.c_void_p(dq.data_ptr()), nq, P.robot_size, P.height_threshold,
                                P.collision_threshold, C.c_void_p(out8.data_ptr()))
        L.trgb_edge_eval_launch(m, C.c_void_p(dp1.data_ptr()), C.c_void_p(dp2.data_ptr()), ne, C.byref(prm),
                                C.c_void_p(st8.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(dd.data_ptr()), None)
        L.trgb_map_sync(m)
    pr = K.prof_collect(); K.prof_enable(False)
    for k in ("k_collision", "k_edge_collide", "k_edge_pca"):
        if k not in pr:
            continue
        v = pr[k]
        gbs = unit_bytes(k, P, rho, cell) * v["units"] / v["ms"] / 1e6
        res[k] = {"units_per_launch": int(v["units"] / v["launches"]), "avg_launch_ms": round(v["ms"] / v["launches"], 4),
                  "units_per_s": v["units"] / v["ms"] * 1e3, "achieved_gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}
    if "k_edge_collide" in pr and "k_edge_pca" in pr:   # K4 as a whole
        ms = pr["k_edge_collide"]["ms"] + pr["k_edge_pca"]["ms"]
        units = pr["k_edge_pca"]["units"]
        gbs = unit_bytes("k_edge_eval", P, rho, cell) * units / ms / 1e6
        res["edge_eval_total"] = {"units_per_s": units / ms * 1e3, "achieved_gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}
    res["note"] = "queries uniform over the map, sorted by 0.6 m tile (spatially coherent threads)"
    return res


def cpu_sample(trg, side, queries_n):
    """Oracle (restated trg.cpp; verbatim reference kdtree.c when oracle/_ref exists) on a bounded
    sample: one `side` x `side` tile of the same generator, same parameters, same seed."""
    P = trg.MOUNTAIN
    pts = trg.terrain.mountain(side, h=0.1, seed=SEED_MAP, tile=(0, 0), world_tiles=(1, 1))
    refkd = (ROOT / "oracle" / "_ref" / "liboracle_refkd.so").exists()
    import _pkg
    o = _pkg.load_oracle().oracle(P, ref_kdtree=refkd)   # the checker, timed as the CPU baseline
    o.seed(SEED_RNG)
    bb = trg.terrain.bbox(pts)
    w0 = time.perf_counter()
    o.set_global_map(pts)
    o.init_graph((0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0))
    w1 = time.perf_counter()
    q = trg.terrain.query_pairs(bb, queries_n, seed=SEED_QUERIES)
    for row in q:
        o.plan(row[:2], row[2:5])
    w2 = time.perf_counter()
    nn, ne = o.counts()
    return {"value": pts.shape[0] / (w1 - w0), "unit": "points/s", "cores": 1, "kind": "port",
            "sample": f"{side}x{side} tile ({pts.shape[0]} points) of the same generator, full build + {queries_n} queries; "
                      f"kd-tree = {'verbatim reference kdtree.c' if refkd else 'restated port'}; single thread "
                      f"(the reference build is single-threaded, serialised by TRG::mtx.graph); host has {os.cpu_count()} cores",
            "build_s": w1 - w0, "nodes_per_sec": nn / (w1 - w0), "paths_per_sec": queries_n / (w2 - w1),
            "nodes": nn, "edges": ne}


def run_reference(a):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    import _pkg
    trg = _pkg.load()
    rows = []
    for i in range(a.warmup + a.steps):
        r = cpu_sample(trg, a.cpu_side, queries_n=50)
        if i >= a.warmup:
            rows.append(r)
    v = float(np.mean([r["value"] for r in rows]))
    step_ms = 1e3 * float(np.mean([r["build_s"] + 50 / r["paths_per_sec"] for r in rows]))
    cb = dict(rows[-1]); cb["value"] = v
    line = {"impl": "reference", "metric": "trg_build_points_per_sec", "value": v, "unit": "points/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2 synthetic mountain heightfield {a.side}x{a.side} (h=0.1 m), config/mountain.yaml params, "
                                   f"full TRG build + path queries",
                       "sample": f"each step = the reference's CPU path (oracle: restated trg.cpp + kdtree.c) on a bounded sample: "
                                 f"one {a.cpu_side}x{a.cpu_side} tile ({a.cpu_side * a.cpu_side} points) of the same generator + 50 queries, 1 thread",
                       "mt19937_seed": SEED_RNG},
            "nodes_per_sec": float(np.mean([r["nodes_per_sec"] for r in rows])),
            "paths_per_sec": float(np.mean([r["paths_per_sec"] for r in rows])),
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=3163, help="lattice side per GPU (3163 -> 10 M points)")
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--cpu-side", type=int, default=CPU_SAMPLE_SIDE)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sat", action="store_true")
    ap.add_argument("--clock-ms", type=int, default=200, help="nvidia-smi sampling period during the timed region (0 = off)")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = max(a.warmup, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_product(a)


if __name__ == "__main__":
    main()
