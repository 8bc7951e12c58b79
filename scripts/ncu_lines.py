#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples / executed instructions, from an ncu report
captured with `--set full --import-source on` (kernels are compiled with -lineinfo).
  python scripts/ncu_lines.py gpurun_out/commit_full.ncu-rep [n_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["/usr/local/cuda/bin/ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
i_thr = hdr.index("Avg. Threads Executed")
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
lines = [(num(r[i_s]), num(r[i_e]), r[0], r[1].strip(), r[i_thr]) for r in rows if r and r[0].isdigit() and len(r) > i_e]
raw = subprocess.run(["/usr/local/cuda/bin/ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
if len(rr) >= 3:
    for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"):
        if k in rr[0]:
            print(k, rr[2][rr[0].index(k)], rr[1][rr[0].index(k)])
ts, te = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"samples {ts}  warp instructions {te}")
for l in sorted(lines, reverse=True)[:top]:
    print(f"{l[0]:6d} {100*l[0]/max(ts,1):5.1f}%  inst {l[1]:8d} {100*l[1]/max(te,1):5.1f}%  thr {l[4]:>5s}  L{l[2]:>4s}  {l[3][:110]}")
