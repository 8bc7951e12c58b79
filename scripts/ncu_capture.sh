#!/bin/bash
# Run under gpurun. 1) plain bench (must exit 0), 2) launch list of the same command (first 2600
# launches = the map build + one whole TRG build), 3) --set full on a few launches of each hot kernel.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
echo "plain ok"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_sample_window_tq|k_edge_collide_tq|k_edge_pca|k_nodes_nearest|k_nearest_z' -s 1000 -c 5 -f -o gpurun_out/prof_build $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full build rc=$?"
bash scripts/ncu_kernels.sh
ls -la gpurun_out/ | tail -12
