#!/bin/bash
# ncu --set full on the saturated thread-per-item kernels (2 launches each), after a plain run.
set -u
mkdir -p gpurun_out
CMD="python scripts/kernel_sat.py --side 1500 --nq 2000000 --ne 500000 --reps 1 --sorted --cell 0.67"
$CMD > gpurun_out/sat_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/sat_plain.log; exit 1; }
cat gpurun_out/sat_plain.log
ncu --set full --clock-control none --import-source on -k regex:'k_collision_tq|k_edge_collide_tq|k_edge_pca' -s 3 -c 3 -f -o gpurun_out/prof_tq $CMD > gpurun_out/ncu_tq.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tq.log
