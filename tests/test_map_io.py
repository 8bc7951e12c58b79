"""Map ingestion and configuration, the step before the hot path:
YAML parameters (TRGPlanner::setParams), PCD files, and the voxel-grid filter (device, -m gpu)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))

MOUNTAIN_YAML = """isVerbose: false
timer:
  graphRate: 5.0
  planningRate: 10.0
map:
  isPrebuiltMap: true
  prebuiltMapPath: "prebuilt_maps/sim_mountain_0.1.pcd"  # downloaded by shellscripts/download_maps.sh
  isVoxelize: false
  voxelSize: 0.1
trg:
  isPrebuiltTRG: false
  prebuiltTRGPath: "prebuilt_graphs/predefined_trg_mountain.pcd"
  isUpdate: false
  expandDist: 0.6
  robotSize: 0.3
  sampleNum: 7
  heightThreshold: 0.16
  collisionThreshold: 0.1
  updateCollisionThreshold: 0.5
  safetyFactor: 3.0
  goalTolerance: 0.8
"""


def same_params(a, b):
    """TrgParams equality at float32 (the C structs hold floats)."""
    from dataclasses import astuple
    return all(np.float32(x) == np.float32(y) for x, y in zip(astuple(a), astuple(b)))


def test_yaml_params_match_reference_config(pkg, built, tmp_path):
    f = tmp_path / "mountain.yaml"
    f.write_text(MOUNTAIN_YAML)
    cfg = pkg.load_params_yaml(f)
    assert same_params(cfg["trg"], pkg.MOUNTAIN)
    assert cfg["is_prebuilt_map"] and not cfg["is_voxelize"] and not cfg["is_update"]
    assert cfg["prebuilt_map_path"] == "prebuilt_maps/sim_mountain_0.1.pcd"
    assert cfg["voxel_size"] == pytest.approx(0.1)
    # defaults of setParams (trg_planner.cpp:107-128) for a config that sets nothing
    g = tmp_path / "empty.yaml"
    g.write_text("isVerbose: true\n")
    d = pkg.load_params_yaml(g)
    assert same_params(d["trg"], pkg.TrgParams(True, 0.6, 0.3, 20, 0.15, 0.2, 0.2, 1.0, 0.8))
    assert d["voxel_size"] == pytest.approx(0.1) and not d["is_prebuilt_map"]
    ref = Path("/root/reference/config")
    if ref.exists():   # build container only
        assert same_params(pkg.load_params_yaml(ref / "mountain.yaml")["trg"], pkg.MOUNTAIN)
        i = pkg.load_params_yaml(ref / "indoor.yaml")
        assert same_params(i["trg"], pkg.INDOOR) and i["is_voxelize"] and i["voxel_size"] == pytest.approx(0.2)
    with pytest.raises(RuntimeError):
        pkg.load_params_yaml(tmp_path / "missing.yaml")


def _lzf_literal(raw: bytes) -> bytes:
    """LZF stream made of literal runs only (valid input for any LZF decoder)."""
    out = bytearray()
    for i in range(0, len(raw), 32):
        chunk = raw[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def test_pcd_ascii_binary_compressed_roundtrip(pkg, built, tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(257, 3)).astype(np.float32) * 50
    for binary in (True, False):
        f = tmp_path / f"c_{int(binary)}.pcd"
        pkg.save_pcd(f, pts, binary=binary)
        back = pkg.load_pcd(f)
        if binary:
            np.testing.assert_array_equal(back, pts)
        else:
            np.testing.assert_allclose(back, pts, rtol=1e-7)
    # extra fields and a different field order (x y z intensity is what PCL writes for PointXYZI)
    inten = rng.uniform(size=257).astype(np.float32)
    rec = np.column_stack([inten, pts]).astype(np.float32)
    hdr = ("# .PCD v0.7\nVERSION 0.7\nFIELDS intensity x y z\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           "WIDTH 257\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 257\nDATA binary\n").encode()
    (tmp_path / "xyzi.pcd").write_bytes(hdr + rec.tobytes())
    np.testing.assert_array_equal(pkg.load_pcd(tmp_path / "xyzi.pcd"), pts)
    # binary_compressed: LZF payload, fields stored one after the other
    soa = np.concatenate([rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3]]).astype(np.float32).tobytes()
    comp = _lzf_literal(soa)
    hdr_c = hdr.replace(b"DATA binary", b"DATA binary_compressed")
    payload = np.array([len(comp), len(soa)], np.uint32).tobytes() + comp
    (tmp_path / "xyzi_c.pcd").write_bytes(hdr_c + payload)
    np.testing.assert_array_equal(pkg.load_pcd(tmp_path / "xyzi_c.pcd"), pts)
    with pytest.raises(RuntimeError, match="Failed to load"):
        pkg.load_pcd(tmp_path / "nope.pcd")


def test_voxel_grid_oracle_basics():
    from voxel_grid_oracle import voxel_grid
    pts = np.array([[0.05, 0.05, 0.0], [0.15, 0.05, 0.0], [0.25, 0.05, 0.0], [0.05, 0.25, 0.0], [-0.05, 0.05, 0.0]], np.float32)
    out = voxel_grid(pts, 0.2)
    # leaves along x: [-0.2,0) [0,0.2) [0.2,0.4) in row y [0,0.2), then row y [0.2,0.4)
    np.testing.assert_allclose(out, [[-0.05, 0.05, 0], [0.10, 0.05, 0], [0.25, 0.05, 0], [0.05, 0.25, 0]], atol=1e-7)


@pytest.mark.gpu
def test_voxel_filter_matches_oracle(pkg, built):
    from voxel_grid_oracle import voxel_grid
    from trg_planner_b200 import kernels as K
    for pts, leaf in ((pkg.terrain.indoor(200, h=0.1, seed=1), 0.2), (pkg.terrain.mountain(300, h=0.1, seed=2), 0.25),
                      (pkg.terrain.stairs(150, h=0.1, seed=5) - np.float32([7.3, 7.1, 0.4]), 0.33)):
        want = voxel_grid(pts, leaf)
        got = K.voxel_filter(pts, leaf)
        assert got.shape == want.shape and got.shape[0] < pts.shape[0]
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)   # float64 vs float64 sums, same order of leaves
    # leaf too small for the extent: passed through unfiltered, like PCL
    far = np.array([[0, 0, 0], [5000, 5000, 100]], np.float32)
    np.testing.assert_array_equal(K.voxel_filter(far, 0.001), far)


@pytest.mark.gpu
def test_load_prebuilt_map_end_to_end(pkg, built, tmp_path):
    """config yaml -> PCD -> voxel filter -> setGlobalMap -> initGraph, against the oracle fed with the
    oracle-filtered cloud (the path `run_trg_planner <config>` takes in the reference)."""
    from voxel_grid_oracle import voxel_grid
    raw = pkg.terrain.indoor(140, h=0.1, seed=1)
    pcd = tmp_path / "sim_indoor_0.1.pcd"
    pkg.save_pcd(pcd, raw, binary=True)
    cfg = tmp_path / "indoor.yaml"
    cfg.write_text(MOUNTAIN_YAML.replace("isVoxelize: false", "isVoxelize: true").replace("voxelSize: 0.1", "voxelSize: 0.2")
                   .replace("expandDist: 0.6", "expandDist: 0.4").replace("sampleNum: 7", "sampleNum: 15")
                   .replace("heightThreshold: 0.16", "heightThreshold: 0.15").replace("updateCollisionThreshold: 0.5", "updateCollisionThreshold: 0.1"))
    c = pkg.load_params_yaml(cfg)
    assert same_params(c["trg"], pkg.INDOOR) and c["is_voxelize"]
    t = pkg.product(c["trg"])
    n_raw, n_map = t.load_prebuilt_map(pcd, c["is_voxelize"], c["voxel_size"])
    filt = voxel_grid(raw, c["voxel_size"])
    assert n_raw == raw.shape[0] and n_map == filt.shape[0]
    t.seed(3)
    assert t.init_graph((3.27, 4.12, 0.0)) == 0
    # the device filter's centroids may differ from the oracle's in the last ulp (float64 sum order),
    # so the oracle is given the product's filtered cloud: the build on it must then be bit-exact
    from trg_planner_b200 import kernels as K
    same_cloud = K.voxel_filter(raw, c["voxel_size"])
    o = pkg.oracle(c["trg"]); o.seed(3); o.set_global_map(same_cloud)
    assert o.init_graph((3.27, 4.12, 0.0)) == 0
    a, b = t.export(), o.export()
    np.testing.assert_array_equal(a.pos, b.pos)
    np.testing.assert_array_equal(a.col, b.col)
    np.testing.assert_allclose(same_cloud, filt, rtol=1e-6, atol=1e-6)
