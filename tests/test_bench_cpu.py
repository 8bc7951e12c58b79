"""-m "not gpu": the reference arm of bench.py (CPU oracle on a bounded sample) prints ONE JSON line
with the keys the driver reads; the B200 arm must refuse to run without a CUDA device."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line(built):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-side", "120"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "trg_build_points_per_sec" and d["unit"] == "points/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] == 1 and cb["kind"] in ("port", "reference") and cb["sample"]
    assert "workload" in d["config"]


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # covered by the GPU tier
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_c4_verify_checker_accepts_an_equal_run_and_catches_a_different_one(pkg, built, tmp_path, monkeypatch):
    """`bench.py --workload c4 --verify` compares the global graph after the build and after every scan with the
    record of scripts/ref_fullsize.py --config c4. Here the checker itself is checked, on a small map: the record is
    written by the reference's own code (or the strongest oracle present), the 'product' is the restated port driven
    through the same calls — it must pass; the same port with another mt19937 seed must fail."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "scripts"))
    import bench_extra as B
    import ref_fullsize as R
    import _pkg
    F = _pkg.load_oracle()
    monkeypatch.setattr(R, "ROOT", tmp_path)
    monkeypatch.setattr(B, "ROOT", tmp_path)
    monkeypatch.setitem(R.C4, "side", 500)
    monkeypatch.setitem(R.C4, "scans", 2)
    monkeypatch.setitem(R.C4, "half", 30.0)   # scans centred at x = -25, -23 m: 5 / 7 m wide strips of the 50 m map

    class A:
        kind = "ref" if F.available("ref") else "port"
        round = "rtest"
    R.run_c4(A)
    rec = json.loads((tmp_path / "profiles" / "rtest_c4_reference_cpu.json").read_text())
    assert [s["scan"] for s in rec["scans"]] == [-1, 0, 1] and rec["scans"][1]["scan_points"] > 1000
    pts = pkg.terrain.mountain(500, h=0.1, seed=4)

    def drive(seed):
        t = F.oracle(pkg.MOUNTAIN, kind="port")
        t.seed(seed)
        t.set_global_map(pts)
        t.init_graph((25.0, 25.0, 0.0))
        v = B._C4Verify(t)
        v.check(-1)
        for k, (cx, cy, scan) in enumerate(B._scans(pts, 50.0, 2, 30.0, np.random.default_rng(9), -25.0)):
            t.set_local_map(cx, cy, scan)
            t.update_graph()
            v.check(k)
        return v.report()
    good = drive(42)
    assert good["pass"] and good["scans_compared"] == 3 and good["scans_pos_bit_exact"] == 3
    assert good["edge_risk_last_scan"]["beyond_tolerance"] == 0
    bad = drive(43)
    assert not bad["pass"]
