#!/bin/bash
# Round-2 ncu evidence; run under gpurun (one GPU). The launch list covers the map index + one whole device-resident
# TRG build + its query batch (4 600 launches: ncu takes ~0.1 s per launch in this process).
# 1) the plain bench command must exit 0, 2) launch list of the
# same command (map build + one whole device-resident TRG build + the query batch), 3) --set full on a few
# launches of every hot kernel in the middle of a build, of the query / graph-preparation kernels, and of K2 / K4
# at saturation. Reports land in gpurun_out/; scripts/ncu_summary.py + scripts/ncu_traffic.py turn them into profiles/.
set -u
mkdir -p gpurun_out /tmp/ncu_r02
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-sat"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
echo "plain ok"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# individual launches instead of the captured graphs for the replayed captures
export TRGB_EXPAND_GRAPHS=0
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_exp_window|k_exp_tables|k_exp_emit|k_exp_deps|k_exp_commit|k_nearest_z|k_edge_collide_tq|k_edge_pca' \
  -s 2400 -c 24 -f -o /tmp/ncu_r02/prof_build $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full build rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_sssp$|k_kd_build|k_g_fill|k_g_keys|k_fin_csr|k_fin_edges|k_scatter|k_sort_cell' \
  -c 10 -f -o /tmp/ncu_r02/prof_query $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full query rc=$?"
unset TRGB_EXPAND_GRAPHS
SAT="python scripts/sat_probe.py"
$SAT > gpurun_out/sat_plain.log 2>&1 || { echo "sat plain failed"; tail -5 gpurun_out/sat_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_collision_tq|k_edge_collide_tq|k_edge_pca_t' -s 3 -c 3 -f \
  -o /tmp/ncu_r02/prof_sat $SAT > gpurun_out/ncu_sat.log 2>&1
echo "full saturated rc=$?"
# the reports stay on the box (gpurun_out/ is capped at 64 MiB): summaries, DRAM traffic and the source page of the two
# kernels with the largest share come back instead
python scripts/ncu_summary.py /tmp/ncu_r02/prof_build.ncu-rep "TRGB_EXPAND_GRAPHS=0 $CMD  (-s 2400 -c 24: two steps in the middle of the build)" > gpurun_out/ncu_full_build.csv
python scripts/ncu_summary.py /tmp/ncu_r02/prof_query.ncu-rep "TRGB_EXPAND_GRAPHS=0 $CMD  (query / graph preparation / map index kernels)" > gpurun_out/ncu_full_query.csv
python scripts/ncu_summary.py /tmp/ncu_r02/prof_sat.ncu-rep "$SAT  (K2 / K4 at saturation)" > gpurun_out/ncu_full_saturated.csv
python scripts/ncu_traffic.py /tmp/ncu_r02/prof_build.ncu-rep /tmp/ncu_r02/prof_query.ncu-rep > gpurun_out/traffic.json
ncu -i /tmp/ncu_r02/prof_query.ncu-rep --page source --csv -k regex:'k_sssp$' 2>/dev/null | head -400 > gpurun_out/ncu_source_k_sssp.csv
ncu -i /tmp/ncu_r02/prof_build.ncu-rep --page source --csv -k regex:'k_exp_commit' 2>/dev/null | head -600 > gpurun_out/ncu_source_k_exp_commit.csv
ls -la gpurun_out/ | tail -14
