"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement (numpy) of pcl::VoxelGrid<pcl::PointXYZ>::filter
as called by TRGPlanner::loadPrebuiltMap (src/planner/trg_planner.cpp:90-94).

PCL is a third-party dependency that is absent from /root/reference and from this image
(`find_package(PCL REQUIRED)`, cpp/trg_planner/CMakeLists.txt:22, version unpinned), so this follows
PCL's published algorithm (pcl/filters/impl/voxel_grid.hpp, PCL 1.10-1.12, recalled):
  * getMinMax3D over the finite points; min_b = floor(min_p * inverse_leaf_size), max_b likewise,
    div_b = max_b - min_b + 1; if div_b.x*div_b.y*div_b.z overflows int: warn and pass the input through;
  * idx = (floor(x*inv) - min_b.x) + (floor(y*inv) - min_b.y)*div_b.x + (floor(z*inv) - min_b.z)*div_b.x*div_b.y
    evaluated in float like PCL (`static_cast<int>(std::floor(x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]))`);
  * sort by idx; one centroid per occupied leaf (min_points_per_voxel_ = 0), output in ascending idx.
PARITY UNPINNED: the reference holds no golden vectors for this step, and PCL sums each leaf in float
in an order that std::sort (not stable) leaves unspecified; this restatement sums in float64.
"""
import numpy as np


def voxel_grid(xyz: np.ndarray, leaf: float) -> np.ndarray:
    p = np.ascontiguousarray(xyz, np.float32)
    finite = np.isfinite(p).all(1)
    q = p[finite]
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = q.min(0), q.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    if int(div[0]) * int(div[1]) * int(div[2]) > np.iinfo(np.int32).max:
        return p.copy()
    ijk = (np.floor(q * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    idx_s = idx[order]
    starts = np.concatenate([[0], np.nonzero(np.diff(idx_s))[0] + 1])
    sums = np.add.reduceat(q[order].astype(np.float64), starts, axis=0)
    counts = np.diff(np.concatenate([starts, [len(idx_s)]]))[:, None]
    return (sums / counts).astype(np.float32)
