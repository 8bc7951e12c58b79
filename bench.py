#!/usr/bin/env python
"""bench.py — TRG build + risk-aware path batch on synthetic terrain.

Default workload (--workload c2 = BASELINE.json configs[1]): one "step" = one full pass of the hot path over
one 10 M-point map: map-index build (TRG::setGlobalMap), graph construction (TRG::initGraph) and a 1k-query
planSafePath batch, all through the reference-facing C facade (include/trg_b200.h).

  value   points/s of the build with the cloud already resident in HBM (trg_set_global_map_dev); the
          per-kernel timings behind `roofline` / `kernels` come from one more step of the same workload run
          with the library's event profiler switched on, right after the timed region
  e2e     the same through host buffers: pinned-host cloud -> H2D inside the timed region, paths / costs
          copied back to host
  N > 1   torchrun, one process per GPU. c2: every rank builds the TRG of its own 10 M-point tile of one
          continuous heightfield (weak scaling); c3: ONE 50 M-point map cut into N tiles (strong scaling).
          Either way the tiles are stitched into ONE graph: boundary nodes and map strips are all-gathered
          (NCCL), cross-tile edges validated by K4, every rank receives the merged CSR, answers its contiguous
          share of the query batch on it (paths cross tile borders) and the results are all-gathered.
          Times are max over ranks.

Other workloads: --workload c3 | c4 (incremental updates: 200 k-point scans against a 20 M-point map) |
c5 (200 M-point multi-level terrain, K2 / K4 roofline sweep over the radius). --verify compares the CUDA
build with the full-size run of the reference itself (profiles/_big/<tag>_ref.npz, scripts/ref_fullsize.py).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libtrg_ref.so = the unmodified
trg.cpp + kdtree.c; the restated oracle when that library is absent) on a bounded sample of the workload.
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
CPU_SAMPLE_SIDE = 1000  # bounded CPU sample: 1000 x 1000 lattice = 1 M points of the same generator
WORKLOADS = {           # tag: (lattice side, map seed, queries, query seed)      SURVEY.md 8d
    "c2": (3163, 2, 1000, 7),
    "c3": (7072, 3, 10000, 8),
    "c4": (4473, 4, 0, 9),
    "c5": (14143, 5, 0, 10),
}


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


# algorithmic bytes per work unit (DESIGN.md §5 / SURVEY.md §8d); rho = map points per m^2
def unit_bytes(kernel: str, P, rho: float, cell: float) -> float:
    r = P.robot_size
    k_r = np.pi * r * r * rho
    e = P.expand_dist
    m = int(np.ceil(e / (0.5 * r)))
    c = 0.5 * e
    a = np.sqrt(c * c + r * r) if c >= r else r
    table = {
        "k_sample_window": 16.0 * k_r + 8 + 1, "k_collision": 16.0 * k_r + 8 + 1, "k_exp_window": 16.0 * k_r + 8 + 1,
        "k_range_count": 16.0 * k_r + 8 + 4,
        "k_nearest_z": 16.0 * (9.0 * cell * cell * rho) + 12,
        # K4 = segment collision samples (k_edge_collide) + ellipse PCA (k_edge_pca); the two halves of
        # SURVEY.md's 16*(m*k(r) + k_e) + 32 + 9 figure, each reading the two endpoints (20 B)
        "k_edge_collide": 16.0 * m * k_r + 20 + 1,
        "k_edge_pca": 16.0 * np.pi * a * a * rho + 20 + 9,
        "k_edge_eval": 16.0 * (m * k_r + np.pi * a * a * rho) + 32 + 9,
        "k_edge_eval_warp": 16.0 * (m * k_r + np.pi * a * a * rho) + 32 + 9,
        "k_bbox": 12.0, "k_count": 12.0, "k_scatter": 32.0, "k_sort_cell": 32.0,
        # device BFS bookkeeping kernels (latency-bound; bytes of one sample / pop record read + written)
        "k_exp_commit": 64.0, "k_exp_emit": 52.0, "k_exp_deps": 8.0 + 8.0 * 4, "k_exp_tables": 2.0 * 32 + 224,
        # K7: 20 B per relaxed edge (12 B edge + 4 B label read + 4 B label write); units = relaxed edges
        "k_sssp": 20.0,
    }
    return float(table.get(kernel, 0.0))


def workload_map(trg, tag, side, rank, world, strong):
    """This rank's cloud: c2 weak = its own side x side tile of a continuous world laid out along x;
    strong (c3) = its x-slab of ONE side x side map."""
    seed = WORKLOADS[tag][1]
    if world == 1:
        return trg.terrain.mountain(side, h=0.1, seed=seed, tile=(0, 0), world_tiles=(1, 1))
    if not strong:
        return trg.terrain.mountain(side, h=0.1, seed=seed, tile=(rank, 0), world_tiles=(world, 1))
    return trg.terrain.mountain_slab(side, h=0.1, seed=seed, slab=rank, n_slabs=world)


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

    tag = a.workload
    side, _, n_queries, q_seed = WORKLOADS[tag]
    if a.side:
        side = a.side
    if a.queries is not None:
        n_queries = a.queries
    strong = tag == "c3"
    P = trg.MOUNTAIN
    pts = workload_map(trg, tag, side, rank, world, strong)
    n = int(pts.shape[0])
    bb = trg.terrain.bbox(pts)
    start = (0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0)
    # queries: N = 1 -> the workload's batch over the map; N > 1 -> ONE global batch over the whole world,
    # the same on every rank, answered in contiguous shares on the merged graph
    if world == 1:
        queries = trg.terrain.query_pairs(bb, n_queries, seed=q_seed)
    else:
        lo = torch.tensor([bb[0][0], bb[1][0]], device="cuda", dtype=torch.float64)
        hi = torch.tensor([bb[0][1], bb[1][1]], device="cuda", dtype=torch.float64)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        world_bb = ((float(lo[0]), float(hi[0])), (float(lo[1]), float(hi[1])))
        queries = trg.terrain.query_pairs(world_bb, n_queries * (1 if strong else world), seed=q_seed)
    d_pts = torch.from_numpy(pts).cuda()
    h_pin = torch.from_numpy(pts).pin_memory()
    h_np = h_pin.numpy()
    t = trg.product(P)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    STAT_KEYS = ("us_draws", "pops", "window_tests", "edge_evals", "us_device_bfs", "us_device_finalize", "us_device_edges", "us_materialize",
                 "device_steps", "device_rounds", "device_redo_pops", "device_interrupts", "device_builds",
                 "us_prep_csr", "us_prep_tree", "us_prep_grid", "us_tree_split", "us_tree_device", "us_tree_adopt", "sssp_relaxed_edges", "sssp_queries",
                 "us_commit", "us_clean", "us_wait")
    merged_stats = {}

    def one_step(resident: bool):
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
        xb = 0
        if dist is None:
            w2 = w1
            r = t.plan_batch(queries)
            found = int(r["found"].sum())
            d2h = int(r["ids"].nbytes + 4 * 4 * len(queries) + 2 * len(queries) + 8 * (len(queries) + 1))
        else:
            from trg_planner_b200 import sharding
            # boundary strips are cut on the device: from the resident input, or (host-buffer leg) from the cloud
            # as the product's map index holds it in HBM after the upload - no second trip over PCIe
            cloud = d_pts if resident else K.map_points_view(device_map_of(t))
            dev = torch.device("cuda", local)
            mg = sharding.build_merged_graph(dist, torch, dev, rank, world, t, cloud, bb, P, K)
            w2 = time.perf_counter()
            res = sharding.plan_sharded(dist, torch, dev, rank, world, mg, queries, P, K)
            found = int(res["found"].sum())
            d2h = int(res["d2h_bytes"])
            xb = int(mg["stats"]["bytes"] + res["gather_bytes"])
            merged_stats.clear()
            merged_stats.update(mg["stats"], cross_tile_paths=int(res["cross_tile_paths"]), merged_nodes=int(mg["n_nodes"]),
                                merged_edges=int(mg["n_edges"]), result_gather_bytes=int(res["gather_bytes"]))
            mg["graph"].close()
        w3 = time.perf_counter()
        nn, ne = t.counts()
        host = {k: t.stat(k) - s0[k] for k in STAT_KEYS}
        host.update(map_ms=round(1e3 * t.seconds("set_global_map"), 2), init_ms=round(1e3 * t.seconds("init_graph"), 2))
        if dist is None:
            host.update(prep_ms=round(1e3 * t.seconds("plan_prep"), 2),
                        snap_ms=round(1e3 * (t.seconds("plan_snap") - t.seconds("plan_prep")), 2),
                        plan_ms=round(1e3 * t.seconds("plan_batch"), 2))
        return dict(host=host, build_s=w1 - w0, exch_s=w2 - w1, query_s=w3 - w2, step_s=w3 - w0, nodes=nn, edges=ne,
                    found=found, d2h=d2h, xbytes=xb, merged=dict(merged_stats))

    def device_map_of(handle):
        import ctypes as C
        tl = C.CDLL(str(ROOT / "trg-planner_b200" / "lib" / "libtrg_b200.so"), mode=C.RTLD_GLOBAL)
        tl.trg_device_map.restype = C.c_void_p
        tl.trg_device_map.argtypes = [C.c_void_p, C.c_char_p]
        return C.c_void_p(tl.trg_device_map(handle.h, b"global"))

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
            r["buildx_s"] = r["build_s"] + r["exch_s"]   # N > 1: the TRG is finished when the tiles are merged
        agg = {k: float(np.mean([r[k] for r in rows])) for k in ("build_s", "buildx_s", "exch_s", "query_s", "step_s")}
        agg.update(total_ms=total_ms, rows=rows, clocks=clocks, launches=launches)
        return agg

    def profiled_steps(n_steps: int):
        """Per-kernel CUDA-event timings (the library's profiler: two events around every launch, on
        the launching stream) of `n_steps` further steps of the same resident workload, run right
        after the timed region. They are NOT taken inside it: the profiler replaces the captured CUDA
        graphs of the device BFS by individual launches and adds ~10 k events per step, which is a
        property of the measurement, not of the path. One profiled step is run and discarded first."""
        K.prof_enable(True)
        one_step(True)
        K.prof_reset()
        r0 = t.stat("sssp_relaxed_edges")
        rows = [one_step(True) for _ in range(n_steps)]
        sync_all()
        profd = K.prof_collect()
        K.prof_enable(False)
        if "k_sssp" in profd:
            profd["k_sssp"]["units"] = float(t.stat("sssp_relaxed_edges") - r0)
        # the per-sample kernels of the device BFS are launched for the step's capacity (a power of two of pops);
        # their work units are the samples the steps actually held
        samples = float(sum(r["host"]["pops"] for r in rows)) * P.sample_num
        for kname in ("k_exp_commit", "k_exp_emit", "k_exp_deps"):
            if kname in profd and samples > 0:
                profd[kname]["units"] = samples
        profd["_wall"] = {"build_ms": 1e3 * float(np.mean([r["build_s"] for r in rows])),
                          "query_ms": 1e3 * float(np.mean([r["query_s"] for r in rows]))}
        return profd

    verify = None
    if a.verify and rank == 0 and world == 1:
        verify = verify_against_reference(trg, t, tag, one_step, P)

    val = timed(True, a.warmup, a.steps)
    val["prof"] = profiled_steps(1) if (not a.no_prof and world == 1) else {}
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
    tot_q = float(len(queries))
    v_build = reduce(val["buildx_s"], RMAX)
    v_tile = reduce(val["build_s"], RMAX)
    v_query = reduce(val["query_s"], RMAX)
    v_step = reduce(val["total_ms"] / a.steps, RMAX)
    e_build = reduce(e2e["buildx_s"], RMAX)
    e_query = reduce(e2e["query_s"], RMAX)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the profiled step among kernels with a byte model)
    rho = n / ((bb[0][1] - bb[0][0]) * (bb[1][1] - bb[1][0]))
    cell = 0.67 * P.robot_size
    peak, peak_src = measured_peak_gbs()
    prof_wall = val["prof"].pop("_wall", None) if isinstance(val["prof"], dict) else None
    prof = {k: v for k, v in val["prof"].items() if v["launches"] > 0}
    roof = None
    kern = {}
    if prof:
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
        psteps = val["prof_steps"]
        kern = {k: dict(launches=int(v["launches"] / psteps), ms_per_step=round(v["ms"] / psteps, 3),
                        units_per_step=int(v["units"] / psteps),
                        achieved_gbs=round(unit_bytes(k, P, rho, cell) * v["units"] / max(v["ms"], 1e-9) / 1e6, 1),
                        frac_of_peak=round(unit_bytes(k, P, rho, cell) * v["units"] / max(v["ms"], 1e-9) / 1e6 / peak, 4))
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_unit": ub, "units_per_launch": d["units"] / d["launches"],
                "avg_launch_ms": avg_ms, "share_of_profiled_step": d["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                "profiled_steps": psteps,
                "note": "dominant kernel by time of one profiled step (individual launches instead of the captured graphs); "
                        "k_sssp units = edges relaxed (20 B each); the device-BFS kernels are launch / latency bound "
                        "(sub-wave batches of one BFS frontier), see kernels_saturated for K2 / K4 at saturating sizes"}

    # ---- how busy the device is during a build: kernel time of the profiled step over its wall time ----
    gpu_busy = None
    if prof and prof_wall:
        side = ("k_exp_deps", "k_nearest_z")          # run on the side stream beside K4: not on the critical path
        query_k = ("k_sssp", "k_edge_cost", "k_sssp_order")
        build_ms = sum(v["ms"] for k, v in prof.items() if k not in query_k) / val["prof_steps"]
        crit_ms = sum(v["ms"] for k, v in prof.items() if k not in query_k and k not in side) / val["prof_steps"]
        gpu_busy = {"build_kernel_ms": round(build_ms, 2), "build_kernel_ms_main_stream": round(crit_ms, 2),
                    "build_wall_ms": round(prof_wall["build_ms"], 2), "frac": round(min(1.0, crit_ms / max(prof_wall["build_ms"], 1e-9)), 3),
                    "note": "profiled step (individual launches + events, slower than the timed steps): share of the build's wall "
                            "time in which a kernel of the main stream is executing; the timed build itself is one chain of "
                            "captured graphs, the host only polls"}

    # ---- saturated-batch kernel numbers (config #5 style, isolated) ----------------------------
    sat = saturated_kernels(trg, K, torch, t, P, bb, rho, cell, peak) if (not a.no_sat and world == 1) else None

    # ---- CPU baseline: the reference on a bounded sample, rank 0, N = 1 only ---------------------
    cpu = None
    if world == 1 and not a.no_cpu:
        cpu = cpu_sample(trg, a.cpu_side, queries_n=100)

    last = val["rows"][-1]
    hb = last["host"]
    line = {
        "metric": "trg_build_points_per_sec", "value": tot_pts / v_build, "unit": "points/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": v_step,
        "higher_is_better": True, "scaling": "strong" if (strong and world > 1) else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": (f"{tag.upper()} synthetic mountain heightfield, h=0.1 m, {n} points on this GPU "
                                f"({int(tot_pts)} in all), config/mountain.yaml params, full TRG build + {int(tot_q)} path queries"),
                   "points_per_gpu": n, "queries": int(tot_q), "mt19937_seed": SEED_RNG,
                   "l2": "inputs larger than L2 (160 MB float4 cloud per build); every step rebuilds from scratch",
                   "parallelism": (("one map in " if strong else "") + f"{world} tiles, one merged graph") if world > 1 else "single GPU"},
        "nodes_per_sec": tot_nodes / v_build, "paths_per_sec": tot_q / v_query,
        "build_ms": 1e3 * v_build, "query_ms": 1e3 * v_query,
        "graph": {"nodes": int(tot_nodes), "edges_rank0": last["edges"], "paths_found": last["found"]},
        "e2e": {"value": tot_pts / e_build, "unit": "points/s",
                "h2d_bytes_per_step": int(n * 12 + queries.nbytes), "d2h_bytes_per_step": e2e["rows"][-1]["d2h"],
                "build_ms": 1e3 * e_build, "query_ms": 1e3 * e_query, "paths_per_sec": tot_q / e_query,
                "nodes_per_sec": tot_nodes / e_build},
        "gpu_launches": int(val["launches"]),
        "gpu_busy": gpu_busy,
        "device_bfs": {"steps": hb.get("device_steps"), "reservation_sweeps": hb.get("device_rounds"),
                       "redone_pops": hb.get("device_redo_pops"), "host_handled_pops": hb.get("device_interrupts"),
                       "useful_over_window_tests": (hb["pops"] * (P.sample_num + 0.2)) / max(hb["window_tests"], 1)},
        "host_breakdown_per_step": {"value_leg": [r["host"] for r in val["rows"]], "e2e_leg": [r["host"] for r in e2e["rows"]]},
        "clocks": val["clocks"],
        "roofline": roof,
        "kernels": kern,
        "kernels_saturated": sat,
        "cpu_baseline": cpu,
    }
    if verify is not None:
        line["verify"] = verify
    if world > 1:
        line["exchange"] = {"leg": "value (resident cloud), last timed step of rank 0", "bytes_per_step": last["xbytes"],
                            "tile_build_ms_max": 1e3 * v_tile, "merge_ms_rank0": 1e3 * val["exch_s"], **val["rows"][-1]["merged"]}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def pos_rows_report(t, mine, theirs):
    """Node positions that differ from the reference's, row by row: x / y must still be bit-identical, and every
    differing z must sit on a query K3 itself flags as an exact tie (two map points at the same float distance:
    the reference returns whichever its kd-tree visits first, kdtree.c:303-350; K3 the lowest index)."""
    import ctypes as C
    from trg_planner_b200 import kernels as K   # (_pkg.load() ran in run_product)
    bad = np.nonzero((mine.view(np.uint32) != theirs.view(np.uint32)).any(axis=1))[0]
    xy_bad = int((mine[bad, :2].view(np.uint32) != theirs[bad, :2].view(np.uint32)).any(axis=1).sum())
    rep = {"rows": int(len(mine)), "rows_differing": int(len(bad)), "rows_with_xy_differing": xy_bad,
           "max_abs_z_diff": float(np.abs(mine[bad, 2] - theirs[bad, 2]).max()) if len(bad) else 0.0,
           "build_z_ties_stat": int(t.stat("z_ties"))}
    flagged = other = 0
    if len(bad):
        L = K.lib()
        tl = C.CDLL(str(ROOT / "trg-planner_b200" / "lib" / "libtrg_b200.so"), mode=C.RTLD_GLOBAL)
        tl.trg_device_map.restype = C.c_void_p
        tl.trg_device_map.argtypes = [C.c_void_p, C.c_char_p]
        m = C.c_void_p(tl.trg_device_map(t.h, b"global"))
        xy = np.ascontiguousarray(mine[bad, :2], np.float32)
        z, idx, tie = np.empty(len(bad), np.float32), np.empty(len(bad), np.int64), np.empty(len(bad), np.uint8)
        rc = L.trgb_nearest_z_batch(m, K._p(xy), len(bad), K._p(z), K._p(idx), K._p(tie))
        assert rc == 0, "trgb_nearest_z_batch"
        flagged = int((tie != 0).sum())
        rep["first_rows"] = [{"node_row": int(b), "xy": [float(v) for v in mine[b, :2]], "z_product": float(mine[b, 2]),
                              "z_reference": float(theirs[b, 2]), "k3_tie_flag": int(tie[j])} for j, b in enumerate(bad[:8])]
    rep["differing_rows_flagged_as_ties_by_k3"] = flagged
    rep["every_differing_row_is_a_flagged_z_tie"] = bool(xy_bad == 0 and flagged == len(bad))
    return rep


def verify_against_reference(trg, t, tag, one_step, P):
    """Full-size parity: the CUDA build of this workload against the reference's own run on the CPU
    (profiles/_big/<tag>_ref.npz, written by scripts/ref_fullsize.py; git-ignored, travels with gpurun)."""
    f = ROOT / "profiles" / "_big" / f"{tag}_ref.npz"
    if not f.exists():
        return {"skipped": f"{f.relative_to(ROOT)} not present (run scripts/ref_fullsize.py --config {tag})"}
    ref = np.load(f)
    one_step(True)
    g = t.export()
    slim = "slim" in ref.files   # big graphs: the arrays that must match bit for bit travel as SHA-256 digests
    if slim:
        import hashlib
        want = dict(zip([str(x) for x in ref["digest_names"]], [str(x) for x in ref["digest_values"]]))
        n_ref, e_ref = int(ref["n_nodes"]), int(ref["n_edges"])
    else:
        n_ref, e_ref = int(len(ref["iter_ids"])), int(len(ref["col"]))
    out = {"reference_file": str(f.relative_to(ROOT)), "nodes": [int(g.n_nodes), n_ref],
           "edges": [int(g.n_edges), e_ref], "rng_draws": [int(t.stat("rng_draws")), int(ref["rng_draws"])]}
    bit = {}
    for k in ("iter_ids", "pos", "state", "row_ptr", "col", "dist"):
        x = getattr(g, k)
        if slim:
            bit[k] = hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:32] == want[k]
        else:
            y = ref[k]
            bit[k] = bool(x.shape == y.shape and np.array_equal(x, y))
    out["bit_exact"] = bit
    if slim:
        out["bit_exact_by"] = "SHA-256 digests of the reference's arrays (scripts/ref_fullsize.py --slim)"
        # the CSR is the reference's once its digests match: the path checks below walk the product's copy
        csr = ({"row_ptr": g.row_ptr, "col": g.col, "dist": g.dist, "weight": ref["weight"]}
               if bit["row_ptr"] and bit["col"] and bit["dist"] else None)
    else:
        csr = {"row_ptr": ref["row_ptr"], "col": ref["col"], "dist": ref["dist"], "weight": ref["weight"]}
    z_tie_only = False
    if not bit["pos"] and "pos" in ref.files and ref["pos"].shape == g.pos.shape:
        out["pos_rows"] = pos_rows_report(t, g.pos, ref["pos"])
        z_tie_only = out["pos_rows"]["every_differing_row_is_a_flagged_z_tie"]
    if g.weight.shape == ref["weight"].shape:
        rel = np.abs(g.weight - ref["weight"]) / np.maximum(np.abs(ref["weight"]), 1e-12)
        rel[(g.weight == 0) & (ref["weight"] == 0)] = 0
        # `if (weight < 0.1) weight = 0` (trg.cpp:360-362) is a step: a risk within rounding of 0.1 can land on either side
        cross = (g.weight == 0) != (ref["weight"] == 0)
        out["edge_risk"] = {"tolerance": 1e-5, "beyond_tolerance": int((rel > 1e-5).sum()),
                            "max_rel": float(rel[~cross].max()) if (~cross).any() else 0.0,
                            "max_rel_note": "over the edges on the same side of the 0.1 step",
                            "fraction_beyond": float((rel > 1e-5).mean()) if rel.size else 0.0,
                            "threshold_crossings_at_0.1": int(cross.sum()),
                            "nonzero_side_of_the_crossings": [float(x) for x in np.maximum(g.weight[cross], ref["weight"][cross])[:16]]}
    # paths
    q = ref["queries"]
    r = t.plan_batch(q)
    same = cost_ok = tie = 0
    sf = np.float32(P.safety_factor)

    def cost_of(ids):
        c = np.float32(0)
        for a_, b_ in zip(ids[:-1], ids[1:]):
            e = csr["row_ptr"][a_] + np.nonzero(csr["col"][csr["row_ptr"][a_]:csr["row_ptr"][a_ + 1]] == b_)[0][0]
            c = np.float32(c + np.float32(np.float32(np.float32(sf * csr["weight"][e]) + np.float32(1)) * csr["dist"][e]))
        return float(c)
    if csr is None:
        out["pass"] = False
        return out
    found_equal = bool(np.array_equal(r["found"], ref["path_found"]))
    known_equal = bool(np.array_equal(r["goal_known"], ref["goal_known"]))
    ends_equal = 0
    worst = 0.0
    for i in range(len(q)):
        if not ref["path_found"][i]:
            continue
        mine = r["ids"][r["offsets"][i]:r["offsets"][i + 1]]
        theirs = ref["path_ids"][ref["path_off"][i]:ref["path_off"][i + 1]]
        ends_equal += int(len(mine) > 0 and mine[0] == theirs[0] and mine[-1] == theirs[-1])
        if np.array_equal(mine, theirs):
            same += 1
            cost_ok += 1
            continue
        cm, ct = cost_of(mine), cost_of(theirs)
        rel = abs(cm - ct) / max(ct, 1e-9)
        worst = max(worst, rel)
        if rel <= 1e-5:
            cost_ok += 1
            tie += 1
    nf = int(ref["path_found"].sum())
    out["paths"] = {"queries": int(len(q)), "found_flags_equal": found_equal, "goal_known_equal": known_equal,
                    "start_goal_nodes_equal": ends_equal, "found": nf, "identical_node_sequences": same,
                    "different_sequence_equal_cost_within_1e-5": tie, "cost_within_1e-5": cost_ok, "worst_rel_cost_diff": worst}
    # node z: bit-exact, or differing only where K3 flagged an exact float-distance tie between two map points
    # (DESIGN.md "Known deviations"; the reference's answer there depends on its kd-tree's visit order)
    out["pos_ok_by"] = "bit-exact" if bit["pos"] else ("bit-exact but for flagged nearest-map-point ties" if z_tie_only else "MISMATCH")
    bit_ok = all(v for k, v in bit.items() if k != "pos") and (bit["pos"] or z_tie_only)
    out["pass"] = bool(bit_ok and out["rng_draws"][0] == out["rng_draws"][1] and found_equal and known_equal
                       and ends_equal == nf and cost_ok == nf and out.get("edge_risk", {}).get("fraction_beyond", 1.0) <= 0.005)
    return out


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
    # sample slots per edge = what edges of length expand_dist need (trg.cpp:279-289 steps by robot_size / 2), plus
    # one for the edges whose float length lands just above it (the device build passes the same count,
    # expand.cu launch_step); the last slot's thread walks any further sample, so the setting changes the thread
    # layout, never the result
    prm = K.EdgeParams(P.robot_size, P.height_threshold, P.collision_threshold, int(np.ceil(P.expand_dist / (0.5 * P.robot_size))) + 1)
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


def cpu_sample(trg, side, queries_n, tag="c2"):
    """The reference's own CPU implementation (oracle/_ref/libtrg_ref.so: unmodified trg.cpp + kdtree.c; the
    restated oracle when that library is absent) on a bounded sample: one `side` x `side` tile of the same
    generator, same parameters, same seed. The full-size run of the same code is recorded once in
    profiles/r02_c2_reference_cpu.json (scripts/ref_fullsize.py)."""
    P = trg.MOUNTAIN
    pts = trg.terrain.mountain(side, h=0.1, seed=WORKLOADS[tag][1], tile=(0, 0), world_tiles=(1, 1))
    import _pkg
    F = _pkg.load_oracle()
    kind = "ref" if F.available("ref") else ("refkd" if F.available("refkd") else "port")
    o = F.oracle(P, kind=kind)   # the checker, timed as the CPU baseline
    o.seed(SEED_RNG)
    bb = trg.terrain.bbox(pts)
    w0 = time.perf_counter()
    o.set_global_map(pts)
    o.init_graph((0.5 * (bb[0][0] + bb[0][1]), 0.5 * (bb[1][0] + bb[1][1]), 0.0))
    w1 = time.perf_counter()
    q = trg.terrain.query_pairs(bb, queries_n, seed=WORKLOADS[tag][3])
    for row in q:
        o.plan(row[:2], row[2:5])
    w2 = time.perf_counter()
    nn, ne = o.counts()
    out = {"value": pts.shape[0] / (w1 - w0), "unit": "points/s", "cores": 1,
           "kind": "reference" if kind == "ref" else "port",
           "sample": f"{side}x{side} tile ({pts.shape[0]} points) of the same generator, full build + {queries_n} queries; "
                     + {"ref": "the reference's unmodified trg.cpp + kdtree.c (oracle/_ref/libtrg_ref.so)",
                        "refkd": "restated trg.cpp on the reference's kdtree.c", "port": "restated oracle"}[kind]
                     + f"; single thread (the reference build is single-threaded, serialised by TRG::mtx.graph); host has {os.cpu_count()} cores",
           "build_s": w1 - w0, "nodes_per_sec": nn / (w1 - w0), "paths_per_sec": queries_n / (w2 - w1),
           "nodes": nn, "edges": ne}
    full = ROOT / "profiles" / "r02_c2_reference_cpu.json"
    if full.exists():
        try:
            fr = json.loads(full.read_text())
            out["full_config_record"] = {"file": str(full.relative_to(ROOT)), "points": fr["points"], "build_s": fr["build_s"],
                                         "points_per_sec": fr["points_per_sec"], "paths_per_sec": fr["paths_per_sec"],
                                         "host": fr.get("host"), "note": "same code, the whole 10 M-point C2 map, run once in the build container"}
        except Exception:
            pass
    return out


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
    side = WORKLOADS["c2"][0]
    line = {"impl": "reference", "metric": "trg_build_points_per_sec", "value": v, "unit": "points/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2 synthetic mountain heightfield {side}x{side} (h=0.1 m), config/mountain.yaml params, "
                                   f"full TRG build + path queries",
                       "sample": f"each step = the reference's CPU path ({cb['kind']}) on a BOUNDED sample: one {a.cpu_side}x{a.cpu_side} "
                                 f"tile ({a.cpu_side * a.cpu_side} points, not the 10 M of the GPU arm) of the same generator + 50 queries, "
                                 f"1 thread; the CPU's points/s FALLS with map size (unbalanced kd-tree), so the per-point ratio is "
                                 f"conservative for the GPU: the full 10 M-point run of the same code is in cpu_baseline.full_config_record",
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--side", type=int, default=0, help="override the lattice side per GPU (c2: 3163 -> 10 M points)")
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--cpu-side", type=int, default=CPU_SAMPLE_SIDE)
    ap.add_argument("--verify", action="store_true", help="compare with the reference's full-size CPU run (profiles/_big)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sat", action="store_true")
    ap.add_argument("--no-prof", action="store_true")
    ap.add_argument("--clock-ms", type=int, default=200, help="nvidia-smi sampling period during the timed region (0 = off)")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = max(a.warmup, 1)
    if a.impl == "reference":
        run_reference(a)
    elif a.workload in ("c4", "c5"):
        sys.path.insert(0, str(ROOT / "scripts"))
        import bench_extra
        getattr(bench_extra, "run_" + a.workload)(a)
    else:
        run_product(a)


if __name__ == "__main__":
    main()
