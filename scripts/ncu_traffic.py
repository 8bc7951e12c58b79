"""profiles/traffic.json from ncu --set full reports: mean dram__bytes_read.sum + dram__bytes_write.sum per launch
of every kernel (bench.py reads it for `roofline.traffic`; names are the library profiler's).
python scripts/ncu_traffic.py gpurun_out/prof_build.ncu-rep gpurun_out/prof_query.ncu-rep > profiles/traffic.json"""
import csv, io, json, subprocess, sys
from collections import defaultdict

ALIAS = {"k_edge_collide_tq": "k_edge_collide", "k_edge_pca_t": "k_edge_pca", "k_collision_tq": "k_collision",
         "k_sample_window_tq": "k_sample_window"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = defaultdict(list)
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in body:
        name = r[ik].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        b = float(r[ir].replace(",", "")) * UNIT.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * UNIT.get(units[iw], 1.0)
        acc[ALIAS.get(name, name)].append(b)
out = {k: round(sum(v) / len(v), 1) for k, v in sorted(acc.items())}
out["_note"] = "mean dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, scripts/ncu_capture_r02.sh); launches per kernel: " + \
    ", ".join(f"{k}={len(v)}" for k, v in sorted(acc.items()))
print(json.dumps(out, indent=1))
