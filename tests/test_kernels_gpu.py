"""-m gpu: kernel-level parity of the CUDA path (through the tier-1 C ABI of
include/trgb_kernels.h) against the CPU oracle on the same seeded inputs.
Bit-exact for sets / booleans / indices; edge weights within 1e-5 relative (north_star), with the
ill-conditioned classification of SURVEY.md A.4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(pkg, built):
    from trg_planner_b200 import kernels
    if kernels.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return kernels


def _queries(pts, n, seed, margin=0.5):
    rng = np.random.default_rng(seed)
    lo = pts[:, :2].min(0) - margin
    hi = pts[:, :2].max(0) + margin
    return rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)


@pytest.mark.parametrize("which,prm", [("mountain", "MOUNTAIN"), ("indoor", "INDOOR")])
def test_collision_and_count_parity(pkg, K, which, prm, small_mountain, small_indoor):
    pts = small_mountain if which == "mountain" else small_indoor
    P = getattr(pkg, prm)
    o = pkg.oracle(P)
    o.set_global_map(pts)
    dm = K.DeviceMap(pts, P.robot_size)
    q = _queries(pts, 200_000, 11)
    for r in (0.15, 0.3, 0.61):
        np.testing.assert_array_equal(dm.range_count(q[:50_000], r), o.range_count(q[:50_000], r))
    got = dm.collision(q, P.robot_size, P.height_threshold, P.collision_threshold)
    want = o.is_collision(q, P.collision_threshold)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    assert 0 < want.mean() < 1  # both outcomes exercised


def test_collision_cell_size_independent(pkg, K, small_mountain):
    P = pkg.MOUNTAIN
    q = _queries(small_mountain, 50_000, 12)
    ref = None
    for cell in (0.1, 0.3, 0.45, 1.0):
        dm = K.DeviceMap(small_mountain, cell)
        got = dm.collision(q, P.robot_size, P.height_threshold, P.collision_threshold)
        if ref is None:
            ref = got
        np.testing.assert_array_equal(got, ref)


def test_collision_large_radius_overflow_path(pkg, K, small_mountain):
    """radius 2.4 m on a 0.1 m lattice => ~1800 points per cylinder (shared buffer overflow path)."""
    import dataclasses
    P = dataclasses.replace(pkg.MOUNTAIN, robot_size=2.4)
    o = pkg.oracle(P)
    o.set_global_map(small_mountain)
    dm = K.DeviceMap(small_mountain, 0.3)
    q = _queries(small_mountain, 3_000, 13)
    np.testing.assert_array_equal(dm.collision(q, 2.4, P.height_threshold, P.collision_threshold),
                                  o.is_collision(q, P.collision_threshold))
    # force the tiny-buffer path too: 4-point cells, many points
    dm2 = K.DeviceMap(small_mountain, 5.0)
    np.testing.assert_array_equal(dm2.collision(q, 2.4, P.height_threshold, P.collision_threshold),
                                  o.is_collision(q, P.collision_threshold))


def test_nearest_z_parity(pkg, K, small_mountain, small_indoor):
    for pts, P in ((small_mountain, pkg.MOUNTAIN), (small_indoor, pkg.INDOOR)):
        o = pkg.oracle(P)
        o.set_global_map(pts)
        dm = K.DeviceMap(pts, P.robot_size)
        q = _queries(pts, 100_000, 14, margin=3.0)
        z, idx, tie = dm.nearest_z(q)
        oz, oidx, otie = o.nearest_z(q)
        np.testing.assert_array_equal(tie, otie)
        ok = tie == 0
        np.testing.assert_array_equal(idx[ok], oidx[ok])
        np.testing.assert_array_equal(z[ok], oz[ok])
        assert ok.mean() > 0.999


def test_nearest_z_tie_rule_on_a_duplicate_bearing_cloud(pkg, K, small_mountain):
    """K3 deviation, quantified (deviations): among map points at the IDENTICAL float distance the
    kernel returns the lowest index, the reference whichever its insertion-order kd-tree visits first
    (kdtree.c:303-362). The tie FLAG is exact, so an affected query is never silent. Jittered clouds (every
    generator of terrain.py, any voxel-filtered PCD) hold no such ties; a cloud with repeated (x, y) columns -
    two returns at different heights - ties on every query that lands on a repeated column."""
    P = pkg.MOUNTAIN
    rng = np.random.default_rng(21)
    base = small_mountain[:: 2].copy()
    for frac in (0.0, 0.05):
        n_dup = int(frac * len(base))
        dup = base[rng.choice(len(base), n_dup, replace=False)].copy()
        dup[:, 2] += rng.uniform(0.3, 1.5, n_dup).astype(np.float32)       # same (x, y), another height
        pts = np.concatenate([base, dup]).astype(np.float32)
        pts = pts[rng.permutation(len(pts))]
        o = pkg.oracle(P)
        o.set_global_map(pts)
        dm = K.DeviceMap(pts, P.robot_size)
        q = _queries(pts, 60_000, 15, margin=0.0)
        z, idx, tie = dm.nearest_z(q)
        oz, oidx, otie = o.nearest_z(q)
        np.testing.assert_array_equal(tie, otie)                 # the flag itself is exact
        differ = z != oz
        assert not (differ & (tie == 0)).any()                   # no disagreement without the flag
        tied = tie != 0
        # a flagged query returns the z of ONE of the tied points (here: of the repeated column)
        print(f"duplicate fraction {frac}: {tied.mean():.4f} of the queries tie, {differ.mean():.4f} return another z than the reference"
              f" ({(differ.sum() / max(1, tied.sum())):.2f} of the tied ones)")
        if frac == 0.0:
            assert tied.sum() == 0
        else:
            assert 0.5 * frac < tied.mean() < 2.5 * frac          # ~ the share of repeated columns
            assert differ.sum() <= tied.sum()


def test_map_index_with_crowded_cells(pkg, K, small_mountain):
    """K1 on a cloud whose cells hold thousands of points (raw, un-voxelised scans; duplicates): the per-cell ordering
    by original index switches to a heap sort beyond 48 points, the layout stays deterministic and every query
    agrees with the reference."""
    P = pkg.MOUNTAIN
    rng = np.random.default_rng(33)
    base = small_mountain[:: 3]
    blobs = []
    for cx, cy, k in ((5.03, 7.01, 6000), (12.4, 3.3, 900), (20.0, 20.0, 60)):
        b = np.column_stack([rng.normal(cx, 0.02, k), rng.normal(cy, 0.02, k), rng.normal(0.5, 0.3, k)])
        blobs.append(b.astype(np.float32))
    pts = np.concatenate([base] + blobs).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    o = pkg.oracle(P)
    o.set_global_map(pts)
    q = np.concatenate([_queries(pts, 4000, 3, margin=0.0),
                        np.column_stack([rng.normal(5.03, 0.3, 500), rng.normal(7.01, 0.3, 500)]).astype(np.float32)])
    outs = []
    for rep in range(2):
        dm = K.DeviceMap(pts, P.robot_size)
        z, idx, tie = dm.nearest_z(q)
        cnt = dm.range_count(q, P.robot_size)
        coll = dm.collision(q, P.robot_size, P.height_threshold, P.collision_threshold)
        outs.append((z, idx, cnt, coll))
        dm.close()
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)          # two builds, one layout
    oz, oidx, otie = o.nearest_z(q)
    ok = otie == 0
    np.testing.assert_array_equal(outs[0][1][ok], oidx[ok])
    np.testing.assert_array_equal(outs[0][0][ok], oz[ok])
    np.testing.assert_array_equal(outs[0][2], o.range_count(q, P.robot_size))
    np.testing.assert_array_equal(outs[0][3].astype(bool), o.is_collision(q, P.collision_threshold).astype(bool))


def _edge_pairs(pts, o, n, seed, e):
    rng = np.random.default_rng(seed)
    a = _queries(pts, n, seed, margin=-1.0)
    ang = rng.uniform(0, 2 * np.pi, n)
    d = rng.uniform(0.1, 1.45 * e, n)
    b = (a + np.stack([d * np.cos(ang), d * np.sin(ang)], 1)).astype(np.float32)
    za, _, _ = o.nearest_z(a)
    zb, _, _ = o.nearest_z(b)
    return np.column_stack([a, za]).astype(np.float32), np.column_stack([b, zb]).astype(np.float32)


@pytest.mark.parametrize("which,prm", [("mountain", "MOUNTAIN"), ("indoor", "INDOOR")])
def test_edge_eval_parity(pkg, K, which, prm, small_mountain, small_indoor):
    pts = small_mountain if which == "mountain" else small_indoor
    P = getattr(pkg, prm)
    o = pkg.oracle(P)
    o.set_global_map(pts)
    dm = K.DeviceMap(pts, P.robot_size)
    p1, p2 = _edge_pairs(pts, o, 60_000, 15, P.expand_dist)
    got = dm.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
    want = o.edge_eval(p1, p2)
    np.testing.assert_array_equal(got["stage"], want["stage"])
    np.testing.assert_array_equal(got["dist"], want["dist"])
    ok = want["stage"] == 0
    assert ok.sum() > 1000
    np.testing.assert_array_equal(got["npts"][ok], want["npts"][ok])
    gw, ww, w64 = got["weight"][ok], want["weight"][ok], want["weight64"][ok]
    tol = 1e-5  # north_star: edge risks within 1e-5 relative
    rel = np.abs(gw - ww) / np.maximum(np.abs(ww), 1e-12)
    rel[(gw == 0) & (ww == 0)] = 0
    bad = rel > tol
    # SURVEY A.4 protocol: a mismatch is a real failure only if the float oracle agrees with its
    # own float64 evaluation (i.e. the case is well-conditioned)
    rel_o = np.abs(ww - w64) / np.maximum(np.abs(w64), 1e-12)
    rel_o[(ww == 0) & (w64 == 0)] = 0
    ill = bad & (rel_o > tol)
    real = bad & ~ill
    # GPU (double accumulation) must match the float64 oracle everywhere it is defined
    rel64 = np.abs(gw - w64) / np.maximum(np.abs(w64), 1e-12)
    rel64[(gw == 0) & (w64 < 0.1)] = 0
    thr_cross = (gw == 0) != (w64 < 0.1)
    print(f"{which}: ok={ok.sum()} bad={bad.sum()} ill={ill.sum()} real={real.sum()} "
          f"max_rel64={rel64[~thr_cross].max():.3g} thr_cross={thr_cross.sum()}")
    # the kernel accumulates the covariance in double and runs the reference's float Jacobi on the
    # rounded result: well-conditioned edges must agree with BOTH oracle pipelines; ill-conditioned
    # ones (float vs float64 oracle already apart) are counted, not asserted
    assert real.sum() == 0
    assert ill.sum() <= 0.01 * ok.sum()
    well = (rel_o <= 0.1 * tol) & ~thr_cross
    assert rel64[well].max() <= tol
    assert thr_cross.sum() <= 3


def test_sssp_parity(pkg, K, small_mountain):
    P = pkg.MOUNTAIN
    o = pkg.oracle(P)
    o.seed(42)
    o.set_global_map(small_mountain)
    assert o.init_graph((15.0, 15.0, 0.0)) == 0
    g = o.export()
    assert np.array_equal(g.ids, np.arange(g.n_nodes))
    dg = K.DeviceGraph(g.row_ptr, g.col, g.weight, g.dist, g.pos, g.state)
    q = pkg.terrain.query_pairs(pkg.terrain.bbox(small_mountain), 300, seed=7)
    starts, goals, ocost, opaths = [], [], [], []
    sf = np.float32(P.safety_factor)
    for row in q:
        r = o.plan(row[:2], row[2:5])
        assert r["found"]
        ids = r["ids"]
        starts.append(ids[0]); goals.append(ids[-1]); opaths.append(ids)
        c = np.float32(0)
        for a, b in zip(ids[:-1], ids[1:]):
            e = g.row_ptr[a] + np.nonzero(g.col[g.row_ptr[a]:g.row_ptr[a + 1]] == b)[0][0]
            c = np.float32(c + np.float32(np.float32(np.float32(sf * g.weight[e]) + np.float32(1)) * g.dist[e]))
        ocost.append(c)
    res = dg.sssp(starts, goals, P.safety_factor)
    assert res["found"].all()
    ocost = np.array(ocost, np.float32)
    rel = np.abs(res["cost"] - ocost) / np.maximum(ocost, 1e-9)
    assert rel.max() <= 1e-5, rel.max()
    same = ties = 0
    for i, ids in enumerate(opaths):
        mine = res["ids"][res["offsets"][i]:res["offsets"][i + 1]]
        assert mine[0] == starts[i] and mine[-1] == goals[i]
        # path_length / avg_risk follow from the node sequence; compare when it is identical
        if len(mine) == len(ids) and np.array_equal(mine, ids):
            same += 1
            assert abs(res["path_length"][i] - r_len(g, ids)) <= 1e-5 * max(1.0, r_len(g, ids))
        else:
            # a different sequence must be a walk over reference edges whose cost TIES with the reference's
            c = np.float32(0)
            for a, b in zip(mine[:-1], mine[1:]):
                e = g.row_ptr[a] + np.nonzero(g.col[g.row_ptr[a]:g.row_ptr[a + 1]] == b)[0][0]
                c = np.float32(c + np.float32(np.float32(np.float32(sf * g.weight[e]) + np.float32(1)) * g.dist[e]))
            assert abs(float(c) - float(ocost[i])) <= 1e-6 * max(float(ocost[i]), 1e-9), (i, c, ocost[i])
            ties += 1
    print(f"identical node sequences: {same}/{len(opaths)}, equal-cost ties: {ties}")
    assert same + ties == len(opaths)


def r_len(g, ids):
    s = np.float32(0)
    for a, b in zip(ids[::-1][:-1], ids[::-1][1:]):  # goal -> start accumulation (trg.cpp:641-659)
        e = g.row_ptr[a] + np.nonzero(g.col[g.row_ptr[a]:g.row_ptr[a + 1]] == b)[0][0]
        s = np.float32(s + g.dist[e])
    return float(s)


def test_thread_path_equals_warp_path(pkg, K, small_mountain, small_indoor):
    """The thread-per-item fast path and the warp-per-item kernels must agree bit for bit
    (collision booleans, edge stage / dist / npts / weight)."""
    for pts, P in ((small_mountain, pkg.MOUNTAIN), (small_indoor, pkg.INDOOR)):
        o = pkg.oracle(P)
        o.set_global_map(pts)
        a, b = K.DeviceMap(pts, 0.5 * P.robot_size), K.DeviceMap(pts, 0.5 * P.robot_size)
        b.set_option("force_warp_path", 1)
        q = _queries(pts, 100_000, 21)
        ca = a.collision(q, P.robot_size, P.height_threshold, P.collision_threshold)
        cb = b.collision(q, P.robot_size, P.height_threshold, P.collision_threshold)
        np.testing.assert_array_equal(ca, cb)
        np.testing.assert_array_equal(ca, o.is_collision(q, P.collision_threshold))
        p1, p2 = _edge_pairs(pts, o, 30_000, 22, P.expand_dist)
        ea = a.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
        eb = b.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
        for k in ("stage", "dist", "npts"):
            np.testing.assert_array_equal(ea[k], eb[k])
        # double sums are accumulated in a different order (per thread vs across lanes): <= 1 ulp-ish
        np.testing.assert_allclose(ea["weight"], eb["weight"], rtol=2e-6, atol=0)


def test_thread_path_overflow_falls_back(pkg, K):
    """A clustered map: most cylinders hold far more points than the per-thread column sized from
    the mean density -> the in-kernel warp fallback must take over without changing results."""
    rng = np.random.default_rng(5)
    base = pkg.terrain.mountain(150, h=0.1, seed=3)
    cl = rng.normal(0, 0.12, size=(60_000, 2)) + rng.uniform(2, 13, size=(30, 1, 2)).repeat(2000, 1).reshape(-1, 2)
    zc = rng.normal(0, 0.1, size=cl.shape[0])
    pts = np.concatenate([base, np.column_stack([cl, zc]).astype(np.float32)])
    P = pkg.MOUNTAIN
    o = pkg.oracle(P)
    o.set_global_map(pts)
    dm = K.DeviceMap(pts, 0.5 * P.robot_size)
    q = np.concatenate([_queries(pts, 20_000, 23), (cl[:20_000] + 0.05).astype(np.float32)])
    np.testing.assert_array_equal(dm.collision(q, P.robot_size, P.height_threshold, P.collision_threshold),
                                  o.is_collision(q, P.collision_threshold))
    cnt = o.range_count(q, P.robot_size)
    assert cnt.max() > 500   # far above any per-thread column


def test_thread_path_dense_map_long_columns(pkg, K):
    """h = 0.085 m: ~39 points per cylinder, so most warps take the 64-value selection network,
    columns run close to their capacity (the per-group overflow check) and some overflow into the
    warp fallback; cell sizes on both sides of the register-held row count of the gather."""
    P = pkg.MOUNTAIN
    pts = pkg.terrain.mountain(170, h=0.085, seed=6)
    rng = np.random.default_rng(33)
    patch = rng.uniform(5.0, 8.0, size=(750, 2))       # +60 % density on 9 m^2: columns of 55 - 70 values
    pts = np.concatenate([pts, np.column_stack([patch, rng.normal(3.0, 0.05, 750)]).astype(np.float32)])
    o = pkg.oracle(P)
    o.set_global_map(pts)
    q = np.concatenate([_queries(pts, 50_000, 31), rng.uniform(4.5, 8.5, size=(10_000, 2)).astype(np.float32)])
    want = o.is_collision(q, P.collision_threshold)
    cnt = o.range_count(q, P.robot_size)
    assert 33 < np.median(cnt) < 48 and (cnt > 64).sum() > 100 and ((cnt > 56) & (cnt <= 64)).sum() > 100
    for cell in (0.4, 0.67, 1.0):
        dm = K.DeviceMap(pts, cell * P.robot_size)
        np.testing.assert_array_equal(dm.collision(q, P.robot_size, P.height_threshold, P.collision_threshold), want)
    p1, p2 = _edge_pairs(pts, o, 20_000, 32, P.expand_dist)
    a, b = K.DeviceMap(pts, 0.67 * P.robot_size), K.DeviceMap(pts, 0.67 * P.robot_size)
    b.set_option("force_warp_path", 1)
    ea = a.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
    eb = b.edge_eval(p1, p2, P.robot_size, P.height_threshold, P.collision_threshold)
    for k in ("stage", "dist", "npts"):
        np.testing.assert_array_equal(ea[k], eb[k])


def test_edge_cases_empty_tiny_and_far(pkg, K):
    """Empty batches, a one-point map, duplicate points (exact distance ties), queries far outside
    the map, and a map far from the origin (float resolution ~1e-4 m at 1 km)."""
    P = pkg.MOUNTAIN
    pts = pkg.terrain.mountain(60, h=0.1, seed=4)
    dm = K.DeviceMap(pts, 0.2)
    e2 = np.zeros((0, 2), np.float32)
    e3 = np.zeros((0, 3), np.float32)
    assert dm.collision(e2, 0.3, 0.16, 0.1).shape == (0,)
    assert dm.range_count(e2, 0.3).shape == (0,)
    assert dm.nearest_z(e2)[0].shape == (0,)
    assert dm.edge_eval(e3, e3, 0.3, 0.16, 0.1)["stage"].shape == (0,)
    # far outside: empty cylinder => collision (trg.cpp:749-752); nearest still defined
    far = np.array([[1e4, 1e4], [-500.0, 3.0], [3.0, 1e6]], np.float32)
    o = pkg.oracle(P); o.set_global_map(pts)
    np.testing.assert_array_equal(dm.collision(far, 0.3, 0.16, 0.1), [1, 1, 1])
    np.testing.assert_array_equal(dm.range_count(far, 0.3), [0, 0, 0])
    z, idx, tie = dm.nearest_z(far)
    oz, oidx, otie = o.nearest_z(far)
    np.testing.assert_array_equal(idx[tie == 0], oidx[tie == 0])
    # one-point map
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    d1 = K.DeviceMap(one, 0.2)
    q = np.array([[1.0, 2.0], [1.2, 2.0], [1.31, 2.0], [50.0, 50.0]], np.float32)
    np.testing.assert_array_equal(d1.range_count(q, 0.3), [1, 1, 0, 0])
    np.testing.assert_array_equal(d1.collision(q, 0.3, 0.16, 0.1), [0, 0, 1, 1])   # 1 point: ratio 0
    np.testing.assert_array_equal(d1.nearest_z(q)[0], [3.0, 3.0, 3.0, 3.0])
    # duplicates: every point twice -> nearest-z ties flagged, collision / counts still exact
    dup = np.concatenate([pts, pts + np.float32([0, 0, 0.5])])
    dd = K.DeviceMap(dup, 0.2)
    od = pkg.oracle(P); od.set_global_map(dup)
    qq = _queries(pts, 20_000, 31)
    np.testing.assert_array_equal(dd.range_count(qq, 0.3), od.range_count(qq, 0.3))
    np.testing.assert_array_equal(dd.collision(qq, 0.3, 0.16, 0.1), od.is_collision(qq, 0.1))
    assert dd.nearest_z(qq)[2].all()     # every nearest is an exact tie between the two copies
    # map 1 km from the origin
    shifted = pts + np.float32([1000.0, -2000.0, 50.0])
    ds = K.DeviceMap(shifted, 0.2)
    os_ = pkg.oracle(P); os_.set_global_map(shifted)
    qs = _queries(shifted, 50_000, 32)
    np.testing.assert_array_equal(ds.collision(qs, 0.3, 0.16, 0.1), os_.is_collision(qs, 0.1))
    np.testing.assert_array_equal(ds.range_count(qs, 0.3), os_.range_count(qs, 0.3))
    with pytest.raises(RuntimeError):
        K.DeviceMap(np.zeros((0, 3), np.float32), 0.2)
    with pytest.raises(RuntimeError):
        K.DeviceMap(np.array([[np.nan, 0, 0]], np.float32), 0.2)


def _kd_insert_all(xy):
    """kd_insert2 in a loop (kdtree.c:167-194): go left iff pos[dir] < node.pos[dir], dir alternating x, y."""
    n = len(xy)
    lo, hi, par = np.full(n, -1, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32)
    ax = np.zeros(n, np.uint8)
    x = xy.tolist()
    for i in range(1, n):
        at, a = 0, 0
        while True:
            low = x[i][a] < x[at][a]
            nxt = lo[at] if low else hi[at]
            if nxt < 0:
                if low:
                    lo[at] = i
                else:
                    hi[at] = i
                par[i], ax[i] = at, a ^ 1
                break
            at, a = nxt, a ^ 1
    return lo, hi, par, ax


def test_parallel_kdtree_build_equals_sequential_insertion(pkg, K):
    """The node kd-tree grown on the device (one level per round) is the tree of n sequential insertions:
    random points, a lattice full of equal coordinates (ties go right), sorted input (a degenerate list)."""
    rng = np.random.default_rng(5)
    cases = [rng.uniform(0, 50, size=(30000, 2)).astype(np.float32),
             (rng.integers(0, 40, size=(20000, 2)) * 0.25).astype(np.float32),
             np.stack([np.arange(3000, dtype=np.float32), np.zeros(3000, np.float32)], 1),
             np.zeros((1, 2), np.float32)]
    for xy in cases:
        want = _kd_insert_all(xy)
        got = K.kdtree_build(xy)
        for a, b, name in zip(got, want, ("lo", "hi", "parent", "axis")):
            np.testing.assert_array_equal(a, b, err_msg=name)
