#!/bin/bash
# Run under gpurun. 1) plain bench (must exit 0), 2) launch list of the same command,
# 3) --set full capture of the dominant kernels (build) and of the saturated K2/K4 launches.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
echo "plain ok"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_sample_window|k_edge_eval|k_nearest_z' -s 600 -c 6 -f -o gpurun_out/prof_build $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full build rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_collision|k_edge_eval' -s 1650 -c 4 -f -o gpurun_out/prof_sat $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full sat rc=$?"
ls -la gpurun_out/
