"""ctypes binding of the kernel-level C ABI (include/trgb_kernels.h, libtrgb_kernels.so).

Tier-1 entry points only (host buffers in, host buffers out). This is the stub a maintainer of
the reference would bind in place of the per-call kd-tree API
(cpp/trg_planner/core/trg_planner/include/kdtree/kdtree.h:30-115). There is no CPU fallback:
every call fails with TRGB_E_CUDA when no device is usable.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
KERNEL_LIB = ROOT / "trg-planner_b200" / "lib" / "libtrgb_kernels.so"

EDGE_OK, EDGE_SLOPE, EDGE_COLLISION, EDGE_EMPTY, EDGE_FEWPTS = range(5)
_vp = C.c_void_p


class MapInfo(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("grid_w", C.c_int32), ("grid_h", C.c_int32),
                ("origin_x", C.c_float), ("origin_y", C.c_float), ("cell_size", C.c_float),
                ("device_bytes", C.c_int64)]


class EdgeParams(C.Structure):
    _fields_ = [("robot_size", C.c_float), ("height_threshold", C.c_float),
                ("collision_threshold", C.c_float), ("max_edge_samples", C.c_int32)]


class GraphDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_edges", C.c_int64), ("row_ptr", _vp), ("col", _vp),
                ("weight", _vp), ("dist", _vp), ("pos_xyz", _vp), ("state", _vp)]


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double),
                ("units", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not KERNEL_LIB.exists():
            raise FileNotFoundError(f"{KERNEL_LIB} missing — run __graft_entry__.build()")
        L = C.CDLL(str(KERNEL_LIB), mode=C.RTLD_GLOBAL)
        L.trgb_last_error.restype = C.c_char_p
        L.trgb_map_create.argtypes = [C.POINTER(_vp), _vp, C.c_int64, C.c_int, C.c_float]
        L.trgb_map_create_dev.argtypes = [C.POINTER(_vp), _vp, C.c_int64, C.c_int, C.c_float]
        L.trgb_map_destroy.argtypes = [_vp]
        L.trgb_map_destroy.restype = None
        L.trgb_map_info.argtypes = [_vp, C.POINTER(MapInfo)]
        L.trgb_map_stream.argtypes = [_vp]
        L.trgb_map_stream.restype = _vp
        L.trgb_map_sync.argtypes = [_vp]
        L.trgb_map_set_option.argtypes = [_vp, C.c_char_p, C.c_int]
        L.trgb_collision_batch.argtypes = [_vp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float, _vp]
        L.trgb_range_count_batch.argtypes = [_vp, _vp, C.c_int64, C.c_float, _vp]
        L.trgb_nearest_z_batch.argtypes = [_vp, _vp, C.c_int64, _vp, _vp, _vp]
        L.trgb_edge_eval_batch.argtypes = [_vp, _vp, _vp, C.c_int64, C.POINTER(EdgeParams), _vp, _vp, _vp, _vp]
        L.trgb_collision_launch.argtypes = [_vp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float, _vp]
        L.trgb_range_count_launch.argtypes = [_vp, _vp, C.c_int64, C.c_float, _vp]
        L.trgb_nearest_z_launch.argtypes = [_vp, _vp, C.c_int64, _vp, _vp, _vp]
        L.trgb_edge_eval_launch.argtypes = [_vp, _vp, _vp, C.c_int64, C.POINTER(EdgeParams), _vp, _vp, _vp, _vp]
        L.trgb_sample_window_launch.argtypes = [_vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_float,
                                                C.c_float, C.c_float, _vp]
        L.trgb_graph_upload.argtypes = [C.POINTER(_vp), C.POINTER(GraphDesc)]
        L.trgb_graph_upload_device.argtypes = [C.POINTER(_vp), C.POINTER(GraphDesc), _vp]
        L.trgb_graph_destroy.argtypes = [_vp]
        L.trgb_graph_destroy.restype = None
        L.trgb_sssp_batch.argtypes = [_vp, _vp, _vp, C.c_int64, C.c_float] + [_vp] * 6 + [C.c_int64]
        L.trgb_voxel_filter.argtypes = [_vp, C.c_int64, C.c_int, C.c_float, _vp, C.POINTER(C.c_int64)]
        L.trgb_prof_enable.argtypes = [C.c_int]
        L.trgb_prof_collect.argtypes = [C.POINTER(ProfEntry), C.c_int]
        L.trgb_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed rc={rc}: {(lib().trgb_last_error() or b'').decode()}")


def device_count() -> int:
    return int(lib().trgb_device_count())


def set_device(i: int):
    _chk(lib().trgb_set_device(i), "trgb_set_device")


def launch_count() -> int:
    return int(lib().trgb_launch_count())


def prof_enable(on: bool):
    lib().trgb_prof_enable(int(on))


def prof_reset():
    lib().trgb_prof_reset()


def prof_collect() -> dict:
    buf = (ProfEntry * 64)()
    n = lib().trgb_prof_collect(buf, 64)
    return {buf[i].name.decode(): dict(launches=buf[i].launches, ms=buf[i].total_ms, units=buf[i].units)
            for i in range(min(n, 64))}


def voxel_filter(xyz: np.ndarray, leaf: float) -> np.ndarray:
    """K8: pcl::VoxelGrid centroid filter on the device (trg_planner.cpp:90-94)."""
    a = np.ascontiguousarray(xyz, np.float32)
    out = np.empty((a.shape[0], 3), np.float32)
    n = C.c_int64()
    rc = lib().trgb_voxel_filter(_p(a), a.shape[0], a.shape[1], float(leaf), _p(out), C.byref(n))
    if rc not in (0, -4):   # -4 = TRGB_E_STATE: leaf too small, cloud passed through (PCL does the same)
        _chk(rc, "trgb_voxel_filter")
    return out[: n.value].copy()


def kdtree_build(xy: np.ndarray):
    """Insertion-order 2-D kd-tree of the graph nodes grown on the device (trg.cpp:249, 528-530):
    children lo / hi, parent and split axis per node."""
    a = np.ascontiguousarray(xy, np.float32)
    n = a.shape[0]
    lo, hi, par = (np.empty(n, np.int32) for _ in range(3))
    ax = np.empty(n, np.uint8)
    L = lib()
    L.trgb_kdtree_build.argtypes = [_vp, C.c_int64, _vp, _vp, _vp, _vp]
    _chk(L.trgb_kdtree_build(_p(a), n, _p(lo), _p(hi), _p(par), _p(ax)), "trgb_kdtree_build")
    return lo, hi, par, ax


class DeviceMap:
    """K1: device cell index over one point cloud (replaces the kd_insert2 loop, trg.cpp:185-188)."""

    def __init__(self, pts: np.ndarray | None, cell: float, *, dev_ptr: int | None = None,
                 n: int | None = None, stride: int = 3):
        self.h = _vp()
        if dev_ptr is not None:
            _chk(lib().trgb_map_create_dev(C.byref(self.h), _vp(dev_ptr), n, stride, cell), "trgb_map_create_dev")
        else:
            pts = np.ascontiguousarray(pts, dtype=np.float32)
            _chk(lib().trgb_map_create(C.byref(self.h), _p(pts), pts.shape[0], pts.shape[1], cell),
                 "trgb_map_create")

    def close(self):
        if self.h:
            lib().trgb_map_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> MapInfo:
        mi = MapInfo()
        _chk(lib().trgb_map_info(self.h, C.byref(mi)), "trgb_map_info")
        return mi

    @property
    def stream(self) -> int:
        return int(lib().trgb_map_stream(self.h) or 0)

    def sync(self):
        _chk(lib().trgb_map_sync(self.h), "trgb_map_sync")

    def set_option(self, key: str, value: int):
        _chk(lib().trgb_map_set_option(self.h, key.encode(), int(value)), "trgb_map_set_option")

    def collision(self, xy, radius, height_thr, ratio_thr):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.empty(xy.shape[0], np.uint8)
        _chk(lib().trgb_collision_batch(self.h, _p(xy), xy.shape[0], radius, height_thr, ratio_thr, _p(out)),
             "trgb_collision_batch")
        return out

    def range_count(self, xy, radius):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        out = np.empty(xy.shape[0], np.int32)
        _chk(lib().trgb_range_count_batch(self.h, _p(xy), xy.shape[0], radius, _p(out)), "trgb_range_count_batch")
        return out

    def nearest_z(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        n = xy.shape[0]
        z, idx, tie = np.empty(n, np.float32), np.empty(n, np.int64), np.empty(n, np.uint8)
        _chk(lib().trgb_nearest_z_batch(self.h, _p(xy), n, _p(z), _p(idx), _p(tie)), "trgb_nearest_z_batch")
        return z, idx, tie

    def edge_eval(self, p1, p2, robot_size, height_thr, collision_thr):
        p1 = np.ascontiguousarray(p1, dtype=np.float32)
        p2 = np.ascontiguousarray(p2, dtype=np.float32)
        n = p1.shape[0]
        prm = EdgeParams(robot_size, height_thr, collision_thr)
        stage, w = np.empty(n, np.uint8), np.empty(n, np.float32)
        d, npts = np.empty(n, np.float32), np.empty(n, np.int32)
        _chk(lib().trgb_edge_eval_batch(self.h, _p(p1), _p(p2), n, C.byref(prm), _p(stage), _p(w), _p(d), _p(npts)),
             "trgb_edge_eval_batch")
        return dict(stage=stage, weight=w, dist=d, npts=npts)

    # tier 2 (device pointers as ints; asynchronous on self.stream)
    def collision_launch(self, d_xy: int, n: int, radius, height_thr, ratio_thr, d_out: int):
        _chk(lib().trgb_collision_launch(self.h, _vp(d_xy), n, radius, height_thr, ratio_thr, _vp(d_out)),
             "trgb_collision_launch")

    def edge_eval_launch(self, d_p1: int, d_p2: int, n: int, robot_size, height_thr, collision_thr,
                         d_stage: int, d_w: int, d_dist: int, d_npts: int = 0):
        prm = EdgeParams(robot_size, height_thr, collision_thr)
        _chk(lib().trgb_edge_eval_launch(self.h, _vp(d_p1), _vp(d_p2), n, C.byref(prm), _vp(d_stage), _vp(d_w),
                                         _vp(d_dist), _vp(d_npts) if d_npts else None), "trgb_edge_eval_launch")


class _DevArray:
    """Minimal __cuda_array_interface__ carrier: lets torch view device memory owned by the library."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}   # (torch refuses the read-only flag; callers only read)


def map_points_view(map_handle):
    """torch view (n, 4) float32 of the indexed cloud of a trgb_map (x, y, z, bit-cast original index): the points
    as the product holds them in HBM, e.g. to cut boundary strips without a second upload of the cloud."""
    import torch
    L = lib()
    L.trgb_map_points.argtypes = [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.trgb_map_sync.argtypes = [_vp]
    p, n = C.c_void_p(), C.c_int64()
    _chk(L.trgb_map_points(map_handle, C.byref(p), C.byref(n)), "trgb_map_points")
    _chk(L.trgb_map_sync(map_handle), "trgb_map_sync")
    return torch.as_tensor(_DevArray(p.value, (n.value, 4)), device="cuda")


class DeviceNodeGrid:
    """K5: device grid over graph nodes; batched exact nearest node (kd_nearest2 on the node tree, trg.cpp:615).
    `pos`: torch tensor (n, >= 2) on the device, or a numpy array."""

    def __init__(self, pos, robot_size: float):
        import torch
        L = lib()
        L.trgb_nodes_create.argtypes = [C.POINTER(_vp), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        L.trgb_nodes_destroy.argtypes = [_vp]
        L.trgb_nodes_destroy.restype = None
        L.trgb_nodes_append_launch.argtypes = [_vp, _vp, C.c_int64, _vp]
        L.trgb_nodes_nearest_launch.argtypes = [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]
        p = pos if hasattr(pos, "is_cuda") else torch.from_numpy(np.ascontiguousarray(pos, np.float32))
        self.xy = p[:, :2].to("cuda", torch.float32).contiguous()
        lo, hi = self.xy.min(0).values.cpu().numpy(), self.xy.max(0).values.cpu().numpy()
        self.h = _vp()
        _chk(L.trgb_nodes_create(C.byref(self.h), float(lo[0]) - 1.0, float(lo[1]) - 1.0, float(hi[0]) + 1.0, float(hi[1]) + 1.0,
                                 1.5 * float(robot_size)), "trgb_nodes_create")
        _chk(L.trgb_nodes_append_launch(self.h, _vp(self.xy.data_ptr()), self.xy.shape[0], None), "trgb_nodes_append_launch")
        torch.cuda.synchronize()

    def nearest(self, xy: np.ndarray) -> np.ndarray:
        import torch
        q = torch.from_numpy(np.ascontiguousarray(xy, np.float32)).cuda()
        n = q.shape[0]
        idx = torch.empty(n, dtype=torch.int32, device="cuda")
        d2 = torch.empty(n, dtype=torch.float32, device="cuda")
        tie = torch.empty(n, dtype=torch.uint8, device="cuda")
        _chk(lib().trgb_nodes_nearest_launch(self.h, _vp(q.data_ptr()), n, _vp(idx.data_ptr()), _vp(d2.data_ptr()),
                                             _vp(tie.data_ptr()), None), "trgb_nodes_nearest_launch")
        torch.cuda.synchronize()
        return idx.cpu().numpy()

    def close(self):
        if self.h:
            lib().trgb_nodes_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGraph:
    """K7: CSR graph resident in HBM + batched risk-aware shortest path (trg.cpp:618-688)."""

    def __init__(self, row_ptr, col, weight, dist, pos_xyz, state):
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        self.col = np.ascontiguousarray(col, np.int32)
        self.weight = np.ascontiguousarray(weight, np.float32)
        self.dist = np.ascontiguousarray(dist, np.float32)
        self.pos = np.ascontiguousarray(pos_xyz, np.float32)
        self.state = np.ascontiguousarray(state, np.int32)
        d = GraphDesc(self.state.shape[0], self.col.shape[0], _p(self.row_ptr), _p(self.col), _p(self.weight),
                      _p(self.dist), _p(self.pos), _p(self.state))
        self.h = _vp()
        _chk(lib().trgb_graph_upload(C.byref(self.h), C.byref(d)), "trgb_graph_upload")

    @classmethod
    def from_device(cls, row_ptr, col, weight, dist, pos_xyz, state):
        """The same graph from torch CUDA tensors (int64 row_ptr, int32 col / state, float32 rest): no PCIe trip."""
        import torch
        self = cls.__new__(cls)
        t = dict(row_ptr=row_ptr.to(torch.int64).contiguous(), col=col.to(torch.int32).contiguous(),
                 weight=weight.to(torch.float32).contiguous(), dist=dist.to(torch.float32).contiguous(),
                 pos=pos_xyz.to(torch.float32).contiguous(), state=state.to(torch.int32).contiguous())
        assert all(v.is_cuda for v in t.values())
        vp = lambda x: C.c_void_p(x.data_ptr())
        d = GraphDesc(int(t["state"].shape[0]), int(t["col"].shape[0]), vp(t["row_ptr"]), vp(t["col"]), vp(t["weight"]),
                      vp(t["dist"]), vp(t["pos"]), vp(t["state"]))
        self.h = _vp()
        stream = torch.cuda.current_stream().cuda_stream
        _chk(lib().trgb_graph_upload_device(C.byref(self.h), C.byref(d), C.c_void_p(stream)), "trgb_graph_upload_device")
        return self

    def close(self):
        if self.h:
            lib().trgb_graph_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sssp(self, starts, goals, safety_factor, capacity=None):
        s = np.ascontiguousarray(starts, np.int32)
        g = np.ascontiguousarray(goals, np.int32)
        n = s.shape[0]
        cap = capacity or max(1 << 20, 2048 * n)
        found, cost = np.zeros(n, np.uint8), np.zeros(n, np.float32)
        plen, risk = np.zeros(n, np.float32), np.zeros(n, np.float32)
        offs, ids = np.zeros(n + 1, np.int64), np.empty(cap, np.int32)
        rc = lib().trgb_sssp_batch(self.h, _p(s), _p(g), n, safety_factor, _p(found), _p(cost), _p(plen),
                                   _p(risk), _p(offs), _p(ids), cap)
        if rc == -3 and offs[n] > cap:  # TRGB_E_NOMEM: needed size returned
            return self.sssp(starts, goals, safety_factor, int(offs[n]))
        _chk(rc, "trgb_sssp_batch")
        return dict(found=found.astype(bool), cost=cost, path_length=plen, avg_risk=risk, offsets=offs,
                    ids=ids[:int(offs[n])].copy())
