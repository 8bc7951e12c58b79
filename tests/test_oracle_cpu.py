"""-m "not gpu": the CPU oracle against the committed golden fixtures (generated with the VERBATIM
reference kdtree.c linked in, tests/golden/make_golden.py), the restated kd-tree port against the
reference kd-tree, and the restated Eigen JacobiSVD against numpy."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
REFKD_LIB = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "liboracle_refkd.so"
CASES = ["mountain_120", "indoor_70", "stairs_100"]


def params_of(pkg, g):
    e, r, s, h, c, u, sf, gt = [float(v) for v in g["params"]]
    return pkg.TrgParams(False, e, r, int(s), h, c, u, sf, gt)


def oracles(pkg):
    out = [("port", False)]
    if REFKD_LIB.exists():
        out.append(("refkd", True))
    return out


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(pkg, built, case):
    g = np.load(GOLD / f"{case}.npz")
    P = params_of(pkg, g)
    for name, refkd in oracles(pkg):
        o = pkg.oracle(P, ref_kdtree=refkd)
        o.seed(int(g["seed"]))
        o.set_global_map(g["pts"])
        assert o.init_graph(tuple(g["start"])) == 0
        assert o.stat("rng_draws") == int(g["rng_draws"]), name
        e = o.export()
        for k in ("iter_ids", "pos", "state", "row_ptr", "col", "weight", "dist"):
            np.testing.assert_array_equal(getattr(e, k), g[k], err_msg=f"{name}:{k}")
        np.testing.assert_array_equal(o.is_collision(g["q"], P.collision_threshold), g["coll"])
        np.testing.assert_array_equal(o.range_count(g["q"], P.robot_size), g["cnt"])
        z, idx, tie = o.nearest_z(g["q"])
        np.testing.assert_array_equal(z, g["z"]); np.testing.assert_array_equal(idx, g["idx"])
        np.testing.assert_array_equal(tie, g["tie"])
        ev = o.edge_eval(g["p1"], g["p2"])
        np.testing.assert_array_equal(ev["stage"], g["ev_stage"])
        np.testing.assert_array_equal(ev["dist"], g["ev_dist"])
        np.testing.assert_array_equal(ev["npts"], g["ev_npts"])
        # the float PCA sums follow the kd-tree's result order: identical for port and reference tree
        np.testing.assert_array_equal(ev["weight"], g["ev_weight"])
        for i, row in enumerate(g["queries"]):
            r = o.plan(row[:2], row[2:5])
            assert r["found"] == bool(g["path_found"][i])
            assert r["goal_known"] == bool(g["goal_known"][i])
            np.testing.assert_array_equal(r["ids"], g["path_ids"][g["path_off"][i]:g["path_off"][i + 1]])
            if r["found"]:
                assert np.float32(r["path_length"]) == g["path_len"][i]
                assert np.float32(r["avg_risk"]) == g["path_risk"][i]


def test_oracle_update_path_matches_golden(pkg, built):
    """setLocalMap / setLocalGraph / updateGraph / isFrontier (trg.cpp:195-231, 456-489, 780-803):
    graph, local node set and draw count after every scan, port and reference kd-tree builds."""
    sys_path_added = str(GOLD) not in __import__("sys").path
    if sys_path_added:
        __import__("sys").path.insert(0, str(GOLD))
    from make_golden import update_scans
    g = np.load(GOLD / "update_120.npz")
    P = params_of(pkg, g)
    for name, refkd in oracles(pkg):
        o = pkg.oracle(P, ref_kdtree=refkd)
        o.seed(int(g["seed"]))
        o.set_global_map(g["pts"])
        assert o.init_graph(tuple(g["start"])) == 0
        for i, (cx, cy, scan) in enumerate(update_scans(g["pts"])):
            o.set_local_map(cx, cy, scan)
            np.testing.assert_array_equal(o.export("local").iter_ids, g[f"local_before_{i}"], err_msg=f"{name}: local set {i}")
            o.update_graph()
            assert o.stat("rng_draws") == int(g[f"draws_{i}"]), (name, i)
            e = o.export()
            for k in ("iter_ids", "pos", "state", "row_ptr", "col", "weight", "dist"):
                np.testing.assert_array_equal(getattr(e, k), g[f"{k}_{i}"], err_msg=f"{name}: scan {i}: {k}")
            np.testing.assert_array_equal(o.export("local").iter_ids, g[f"local_after_{i}"])
        np.testing.assert_array_equal(o.is_frontier(g["frontier_q"]), g["frontier"])
        assert len(g["iter_ids_2"]) != len(g["iter_ids_0"]) or not np.array_equal(g["state_2"], g["state_0"])  # the scans did change the graph


def test_kdtree_port_equals_reference_kdtree(pkg, built):
    """Restated kd-tree (oracle/kdtree_port.h) vs the reference's kdtree.c compiled where it lies:
    same result ORDER for range queries is what setGoal / wireEdge order depend on."""
    if not REFKD_LIB.exists():
        pytest.skip("oracle/_ref not built (reference tree absent on this box)")
    P = pkg.INDOOR
    pts = pkg.terrain.indoor(90, h=0.2, seed=3)
    a, b = pkg.oracle(P, False), pkg.oracle(P, True)
    for o in (a, b):
        o.seed(5); o.set_global_map(pts); assert o.init_graph((3.27, 4.12, 0.0)) == 0
    ea, eb = a.export(), b.export()
    for k in ("iter_ids", "pos", "state", "row_ptr", "col", "weight", "dist"):
        np.testing.assert_array_equal(getattr(ea, k), getattr(eb, k))
    # raster-ordered input (degenerate tree), duplicates in x and y
    raster = pkg.terrain.mountain(60, h=0.1, seed=9, shuffle=False)
    raster[::7, 0] = raster[0, 0]
    a2, b2 = pkg.oracle(P, False), pkg.oracle(P, True)
    rng = np.random.default_rng(1)
    q = rng.uniform(-0.5, 6.5, size=(3000, 2)).astype(np.float32)
    q[:200] = raster[:200, :2]   # queries exactly on points
    for o in (a2, b2):
        o.set_global_map(raster)
    for r in (0.05, 0.3, 1.1):
        np.testing.assert_array_equal(a2.range_count(q, r), b2.range_count(q, r))
    za, ia, ta = a2.nearest_z(q)
    zb, ib, tb = b2.nearest_z(q)
    np.testing.assert_array_equal(ia, ib)   # including tie resolution by tree visit order
    np.testing.assert_array_equal(ta, tb)


def _jacobi_weight_numpy(rows):
    """Independent float64 evaluation of trg.cpp:332-363 with numpy's symmetric eigensolver."""
    A = rows.astype(np.float64)
    C = np.cov(A.T, ddof=1)
    w, V = np.linalg.eigh(C)
    order = np.argsort(-w)
    V = V[:, order] / np.sqrt(3.0)
    weight = np.float32(0.8) * abs(V[2, 0]) + np.float64(np.float32(1 - np.float32(0.8))) * abs(V[2, 1])
    return 0.0 if weight < 0.1 else weight


def test_edge_weight_restatement_vs_numpy(pkg, built):
    """The restated Eigen arithmetic (mean / covariance / JacobiSVD / Frobenius normalisation):
    its float64 pipeline must agree with numpy.linalg.eigh to 1e-9; its float pipeline to 1e-5
    on well-conditioned patches."""
    P = pkg.MOUNTAIN
    pts = pkg.terrain.mountain(120, h=0.1, seed=2)
    o = pkg.oracle(P)
    o.set_global_map(pts)
    g = np.load(GOLD / "mountain_120.npz")
    ev = o.edge_eval(g["p1"], g["p2"])
    ok = np.nonzero(ev["stage"] == 0)[0][:300]
    assert len(ok) > 100
    checked = 0
    for i in ok:
        p1, p2 = g["p1"][i], g["p2"][i]
        d = np.float32(np.hypot(p1[0] - p2[0], p1[1] - p2[1]))
        dirv = (p2[:2] - p1[:2]) / np.float32(np.hypot(*(p2[:2] - p1[:2])))
        c = np.float32(0.5) * d
        b = np.float32(P.robot_size)
        a = np.sqrt(c * c + b * b) if c >= b else b
        ctr = p1[:2] + c * dirv
        q = pts[:, :2] - ctr
        inr = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) <= a * a
        px = dirv[0] * q[:, 0] - dirv[1] * q[:, 1]
        py = dirv[1] * q[:, 0] + dirv[0] * q[:, 1]
        keep = inr if a == b else inr & ((px * px) * (b * b) + (py * py) * (a * a) < a * a * b * b)
        if keep.sum() != ev["npts"][i]:
            continue   # float32-vs-float64 membership differs on a boundary point: not the point of this test
        rows = np.column_stack([px[keep], py[keep], pts[keep, 2]])
        w_np = _jacobi_weight_numpy(rows)
        w64 = ev["weight64"][i]
        if (w_np == 0) != (w64 == 0):
            continue
        assert abs(w64 - w_np) <= 1e-6 * max(1.0, abs(w_np)), (i, w64, w_np)
        checked += 1
    assert checked > 80


def test_refine_path_and_frontier(pkg, built):
    P = pkg.MOUNTAIN
    o = pkg.oracle(P)
    path = np.array([[0, 0, 0], [1, 0, 0.5], [2, 1, 1.0], [3, 3, 0.0]], np.float32)
    out = o.refine_path(path)
    assert out.shape == (6, 3)
    np.testing.assert_allclose(out[0], (path[0] + path[1]) / 2)          # 2 taps at the head
    np.testing.assert_allclose(out[1], (path[0] + 2 * path[1]) / 3, rtol=1e-6)
    np.testing.assert_array_equal(out[-1], path[-1])                        # last point kept
    assert o.refine_path(path[:1]).shape[0] == 0                            # single point -> empty
