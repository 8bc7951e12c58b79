"""ctypes binding shared by the product host library (prefix `trg_`, include/trg_b200.h) and
the CPU oracle (prefix `orc_`, oracle/trg_oracle.h): both export the same C facade of the
reference's `TRG` class (trg.h:50-98), so parity tests drive them through one class.

This is the reference-side stub a maintainer would write for a Python front end in place of
python/trg_planner/pybind/trg_planner_pybind.cpp:19-78.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from .params import CParams, TrgParams

ROOT = Path(__file__).resolve().parent.parent
PRODUCT_LIB = ROOT / "trg-planner_b200" / "lib" / "libtrg_b200.so"
KERNEL_LIB = ROOT / "trg-planner_b200" / "lib" / "libtrgb_kernels.so"

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GraphSnapshot:
    """Graph export: CSR in id order (row = id, columns in `edges_` order)."""

    def __init__(self, iter_ids, ids, pos, state, row_ptr, col, weight, dist):
        self.iter_ids, self.ids, self.pos, self.state = iter_ids, ids, pos, state
        self.row_ptr, self.col, self.weight, self.dist = row_ptr, col, weight, dist

    @property
    def n_nodes(self):
        return int(self.ids.shape[0])

    @property
    def n_edges(self):
        return int(self.col.shape[0])


class TrgFacade:
    """One TRG instance behind the C facade. prefix = 'trg' (product) or 'orc' (oracle)."""

    def __init__(self, lib_path: os.PathLike, prefix: str, params: TrgParams):
        if not Path(lib_path).exists():
            raise FileNotFoundError(
                f"{lib_path} is missing — run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.lib = C.CDLL(str(lib_path), mode=C.RTLD_GLOBAL)
        self.p = prefix
        self.params = params
        self._sig()
        cp = params.to_c()
        self.h = self._f("create")(C.byref(cp))
        if not self.h:
            raise RuntimeError(f"{prefix}_create failed: {self.last_error()}")

    def _f(self, name):
        return getattr(self.lib, f"{self.p}_{name}")

    def _sig(self):
        f = self._f
        f("create").restype = _vp
        f("create").argtypes = [C.POINTER(CParams)]
        f("destroy").argtypes = [_vp]
        f("seed").argtypes = [_vp, C.c_uint32]
        f("set_global_map").argtypes = [_vp, _vp, C.c_int64]
        f("set_local_map").argtypes = [_vp, C.c_float, C.c_float, _vp, C.c_int64]
        f("init_graph").argtypes = [_vp, C.c_int, C.c_float, C.c_float, C.c_float]
        f("update_graph").argtypes = [_vp]
        f("graph_counts").argtypes = [_vp, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        f("graph_export").argtypes = [_vp, C.c_char_p] + [_vp] * 8
        f("plan").argtypes = [_vp] + [C.c_float] * 5 + [_vp, _vp, C.c_int, C.POINTER(C.c_int)] + \
            [C.POINTER(C.c_float)] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        f("refine_path").argtypes = [_vp, _vp, C.c_int, _vp, C.POINTER(C.c_int)]
        f("is_collision_batch").argtypes = [_vp, C.c_char_p, _vp, C.c_int64, C.c_float, _vp]
        f("range_count_batch").argtypes = [_vp, C.c_char_p, _vp, C.c_int64, C.c_float, _vp]
        f("nearest_z_batch").argtypes = [_vp, C.c_char_p, _vp, C.c_int64, _vp, _vp, _vp]
        f("edge_eval_batch").argtypes = [_vp, C.c_char_p, _vp, _vp, C.c_int64] + [_vp] * 5
        f("is_frontier_batch").argtypes = [_vp, _vp, C.c_int64, _vp]
        f("last_seconds").restype = C.c_double
        f("last_seconds").argtypes = [_vp, C.c_char_p]
        f("stat").restype = C.c_int64
        f("stat").argtypes = [_vp, C.c_char_p]
        f("check_reached").argtypes = [_vp, C.c_float, C.c_float]
        f("check_replan").argtypes = [_vp, C.c_float, C.c_float, _vp, C.c_int]
        if self.p == "trg":
            f("last_error").restype = C.c_char_p
            f("last_error").argtypes = []
            f("plan_batch").argtypes = [_vp, _vp, C.c_int64] + [_vp] * 8 + [C.c_int64]
            f("save_graph").argtypes = [_vp, C.c_char_p]
            f("load_graph").argtypes = [_vp, C.c_char_p]
            f("set_tuning").argtypes = [_vp, C.c_char_p, C.c_double]
            f("set_global_map_dev").argtypes = [_vp, _vp, C.c_int64, C.c_int]

    def last_error(self) -> str:
        if self.p != "trg":
            return ""
        return (self._f("last_error")() or b"").decode()

    def _chk(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{self.p}_{what} failed rc={rc}: {self.last_error()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- TRG API mirror ----
    def seed(self, s: int):
        self._f("seed")(self.h, s)

    def set_global_map(self, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        self._chk(self._f("set_global_map")(self.h, _ptr(pts), pts.shape[0]), "set_global_map")

    def set_local_map(self, sx, sy, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        self._chk(self._f("set_local_map")(self.h, sx, sy, _ptr(pts), pts.shape[0]), "set_local_map")

    def init_graph(self, start, is_pre_map=True):
        return self._chk(self._f("init_graph")(self.h, int(is_pre_map), *[float(v) for v in start]),
                         "init_graph")

    def update_graph(self):
        self._chk(self._f("update_graph")(self.h), "update_graph")

    def counts(self, type_="global"):
        n, e = C.c_int64(), C.c_int64()
        self._chk(self._f("graph_counts")(self.h, type_.encode(), C.byref(n), C.byref(e)), "graph_counts")
        return n.value, e.value

    def export(self, type_="global", edges: bool = True) -> GraphSnapshot:
        n, e = self.counts(type_)
        iter_ids = np.empty(n, np.int32)
        ids = np.empty(n, np.int32)
        pos = np.empty((n, 3), np.float32)
        state = np.empty(n, np.int32)
        row_ptr = np.empty(n + 1, np.int64) if edges else None
        col = np.empty(e, np.int32) if edges else None
        w = np.empty(e, np.float32) if edges else None
        d = np.empty(e, np.float32) if edges else None
        self._chk(self._f("graph_export")(self.h, type_.encode(), *[_ptr(a) for a in
                                          (iter_ids, ids, pos, state, row_ptr, col, w, d)]), "graph_export")
        return GraphSnapshot(iter_ids, ids, pos, state, row_ptr, col, w, d)

    def plan(self, start2d, goal3d, max_pts=1 << 16):
        path = np.empty((max_pts, 3), np.float32)
        ids = np.empty(max_pts, np.int32)
        n = C.c_int()
        dd, pl, ar = C.c_float(), C.c_float(), C.c_float()
        known, nexp = C.c_int(), C.c_int64()
        rc = self._chk(self._f("plan")(self.h, float(start2d[0]), float(start2d[1]),
                                       float(goal3d[0]), float(goal3d[1]), float(goal3d[2]),
                                       _ptr(path), _ptr(ids), max_pts, C.byref(n), C.byref(dd),
                                       C.byref(pl), C.byref(ar), C.byref(known), C.byref(nexp)), "plan")
        k = min(n.value, max_pts)
        return dict(found=bool(rc), path=path[:k].copy(), ids=ids[:k].copy(), direct_dist=dd.value,
                    path_length=pl.value, avg_risk=ar.value, goal_known=bool(known.value),
                    n_expanded=nexp.value)

    def plan_batch(self, queries: np.ndarray, max_total_nodes: int | None = None):
        """queries: (n,5) float32 rows (sx, sy, gx, gy, gz). Product only (GPU batched SSSP)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        n = q.shape[0]
        cap = max_total_nodes or max(1 << 20, n * 4096)
        found = np.zeros(n, np.uint8)
        cost = np.zeros(n, np.float32)
        length = np.zeros(n, np.float32)
        risk = np.zeros(n, np.float32)
        direct = np.zeros(n, np.float32)
        known = np.zeros(n, np.uint8)
        offs = np.zeros(n + 1, np.int64)
        ids = np.empty(cap, np.int32)
        self._chk(self._f("plan_batch")(self.h, _ptr(q), n, _ptr(found), _ptr(cost), _ptr(length),
                                        _ptr(risk), _ptr(direct), _ptr(known), _ptr(offs), _ptr(ids),
                                        cap), "plan_batch")
        return dict(found=found.astype(bool), cost=cost, path_length=length, avg_risk=risk,
                    direct_dist=direct, goal_known=known.astype(bool), offsets=offs,
                    ids=ids[:int(offs[-1])].copy())

    def save_graph(self, path: str):
        self._chk(self._f("save_graph")(self.h, str(path).encode()), "save_graph")

    def load_graph(self, path: str):
        self._chk(self._f("load_graph")(self.h, str(path).encode()), "load_graph")

    def set_tuning(self, key: str, value: float):
        self._chk(self._f("set_tuning")(self.h, key.encode(), float(value)), "set_tuning")

    def set_global_map_dev(self, dev_ptr: int, n: int, stride: int = 3):
        """Map cloud already resident in HBM (bench `value` leg)."""
        self._chk(self._f("set_global_map_dev")(self.h, C.c_void_p(dev_ptr), n, stride), "set_global_map_dev")

    def check_reached(self, xy) -> bool:
        return bool(self._chk(self._f("check_reached")(self.h, float(xy[0]), float(xy[1])), "check_reached"))

    def check_replan(self, xy, path: np.ndarray) -> bool:
        p = np.ascontiguousarray(path, dtype=np.float32)
        return bool(self._chk(self._f("check_replan")(self.h, float(xy[0]), float(xy[1]), _ptr(p), p.shape[0]), "check_replan"))

    def load_prebuilt_map(self, pcd_path, is_voxelize=False, voxel_size=0.1):
        """TRGPlanner::loadPrebuiltMap (trg_planner.cpp:76-101). Product only. Returns (n_raw, n_map)."""
        self._f("load_prebuilt_map").argtypes = [_vp, C.c_char_p, C.c_int, C.c_float, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        a, b = C.c_int64(), C.c_int64()
        self._chk(self._f("load_prebuilt_map")(self.h, str(pcd_path).encode(), int(is_voxelize), float(voxel_size),
                                               C.byref(a), C.byref(b)), "load_prebuilt_map")
        return a.value, b.value

    def refine_path(self, path: np.ndarray):
        p = np.ascontiguousarray(path, dtype=np.float32)
        out = np.empty((2 * max(p.shape[0], 1), 3), np.float32)
        n = C.c_int()
        self._f("refine_path")(self.h, _ptr(p), p.shape[0], _ptr(out), C.byref(n))
        return out[:n.value].copy()

    # ---- pure kernels ----
    def is_collision(self, xy, threshold, type_="global"):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.empty(xy.shape[0], np.uint8)
        self._chk(self._f("is_collision_batch")(self.h, type_.encode(), _ptr(xy), xy.shape[0],
                                                float(threshold), _ptr(out)), "is_collision_batch")
        return out

    def range_count(self, xy, radius, type_="global"):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.empty(xy.shape[0], np.int32)
        self._chk(self._f("range_count_batch")(self.h, type_.encode(), _ptr(xy), xy.shape[0],
                                               float(radius), _ptr(out)), "range_count_batch")
        return out

    def nearest_z(self, xy, type_="global"):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        z = np.empty(xy.shape[0], np.float32)
        idx = np.empty(xy.shape[0], np.int64)
        tie = np.empty(xy.shape[0], np.uint8)
        self._chk(self._f("nearest_z_batch")(self.h, type_.encode(), _ptr(xy), xy.shape[0], _ptr(z),
                                             _ptr(idx), _ptr(tie)), "nearest_z_batch")
        return z, idx, tie

    def edge_eval(self, p1, p2, type_="global"):
        p1 = np.ascontiguousarray(p1, dtype=np.float32)
        p2 = np.ascontiguousarray(p2, dtype=np.float32)
        n = p1.shape[0]
        stage = np.empty(n, np.uint8)
        w = np.empty(n, np.float32)
        w64 = np.empty(n, np.float64)
        d = np.empty(n, np.float32)
        npts = np.empty(n, np.int32)
        self._chk(self._f("edge_eval_batch")(self.h, type_.encode(), _ptr(p1), _ptr(p2), n, _ptr(stage),
                                             _ptr(w), _ptr(w64), _ptr(d), _ptr(npts)), "edge_eval_batch")
        return dict(stage=stage, weight=w, weight64=w64, dist=d, npts=npts)

    def is_frontier(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.empty(xy.shape[0], np.uint8)
        self._chk(self._f("is_frontier_batch")(self.h, _ptr(xy), xy.shape[0], _ptr(out)), "is_frontier_batch")
        return out

    def seconds(self, what: str) -> float:
        return float(self._f("last_seconds")(self.h, what.encode()))

    def stat(self, what: str) -> int:
        return int(self._f("stat")(self.h, what.encode()))


def _product_lib():
    if not PRODUCT_LIB.exists():
        raise FileNotFoundError(f"{PRODUCT_LIB} is missing — run __graft_entry__.build()")
    L = C.CDLL(str(PRODUCT_LIB), mode=C.RTLD_GLOBAL)
    L.trg_last_error.restype = C.c_char_p
    L.trg_load_params_yaml.argtypes = [C.c_char_p, C.POINTER(CParams), C.POINTER(C.c_int), C.c_char_p, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.trg_load_pcd.argtypes = [C.c_char_p, _vp, C.c_int64, C.POINTER(C.c_int64)]
    L.trg_save_pcd.argtypes = [C.c_char_p, _vp, C.c_int64, C.c_int]
    L.trg_load_prebuilt_map.argtypes = [_vp, C.c_char_p, C.c_int, C.c_float, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    return L


def load_params_yaml(path) -> dict:
    """TRGPlanner::setParams (trg_planner.cpp:103-129) through the C facade."""
    L = _product_lib()
    cp = CParams()
    pre, vox, upd = C.c_int(), C.c_int(), C.c_int()
    vs = C.c_float()
    buf = C.create_string_buffer(1024)
    if L.trg_load_params_yaml(str(path).encode(), C.byref(cp), C.byref(pre), buf, 1024, C.byref(vox), C.byref(vs), C.byref(upd)) < 0:
        raise RuntimeError((L.trg_last_error() or b"").decode())
    trg = TrgParams(bool(cp.is_verbose), cp.expand_dist, cp.robot_size, cp.sample_num, cp.height_threshold,
                    cp.collision_threshold, cp.update_collision_threshold, cp.safety_factor, cp.goal_tolerance)
    return dict(trg=trg, is_prebuilt_map=bool(pre.value), prebuilt_map_path=buf.value.decode(),
                is_voxelize=bool(vox.value), voxel_size=vs.value, is_update=bool(upd.value))


def load_pcd(path) -> np.ndarray:
    L = _product_lib()
    n = C.c_int64()
    if L.trg_load_pcd(str(path).encode(), None, 0, C.byref(n)) < 0:
        raise RuntimeError((L.trg_last_error() or b"").decode())
    out = np.empty((n.value, 3), np.float32)
    if L.trg_load_pcd(str(path).encode(), _ptr(out), n.value, C.byref(n)) < 0:
        raise RuntimeError((L.trg_last_error() or b"").decode())
    return out


def save_pcd(path, xyz: np.ndarray, binary: bool = True):
    L = _product_lib()
    a = np.ascontiguousarray(xyz, np.float32)
    if L.trg_save_pcd(str(path).encode(), _ptr(a), a.shape[0], int(binary)) < 0:
        raise RuntimeError((L.trg_last_error() or b"").decode())



def product(params: TrgParams) -> TrgFacade:
    """The B200 path. Fails loudly when the CUDA libraries are missing."""
    return TrgFacade(PRODUCT_LIB, "trg", params)
