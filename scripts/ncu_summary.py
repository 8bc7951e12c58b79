"""Turn an .ncu-rep (ncu --set full) into the compact per-launch CSV kept under profiles/:
python scripts/ncu_summary.py gpurun_out/prof_tq.ncu-rep "<command line that was profiled>" > profiles/rNN_ncu_full_*.csv"""
import csv, io, subprocess, sys

KEEP = """dram__bytes_read.sum dram__bytes_read.sum.pct_of_peak_sustained_elapsed dram__bytes_read.sum.per_second
dram__bytes_write.sum dram__bytes_write.sum.pct_of_peak_sustained_elapsed dram__bytes_write.sum.per_second
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct
l1tex__throughput.avg.pct_of_peak_sustained_active l1tex__throughput.avg.pct_of_peak_sustained_elapsed
launch__block_size launch__grid_size launch__occupancy_limit_barriers launch__occupancy_limit_blocks
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__occupancy_limit_warps
launch__registers_per_thread launch__registers_per_thread_allocated launch__shared_mem_per_block_dynamic
lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio""".split()

rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
cols = [h for h in KEEP if h in hdr] + sorted(stall)
idx = [hdr.index(c) for c in cols]
ik = hdr.index("Kernel Name")
out = csv.writer(sys.stdout, lineterminator="\n")
print(f"# {cmd}")
out.writerow(["ID", "Kernel Name"] + cols)
out.writerow(["", ""] + [units[i] for i in idx])
for r in body:
    out.writerow([r[0], r[ik].split("(")[0]] + [r[i] for i in idx])
