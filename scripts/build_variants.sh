#!/bin/bash
# Build-time sweep of the thread-per-query kernel knobs (common.cuh: TQ_FLAT, TQ_UNROLL, TQ_MINBLK,
# TQ_CAP): one libtrgb_kernels_<tag>.so per variant under trg-planner_b200/lib/variants/ (git-ignored),
# timed on the GPU with `python scripts/kernel_sat.py --lib <path>`.
set -e
# (only queries.cu depends on the knobs; the other objects come from the regular `make`)
cd "$(dirname "$0")/../trg-planner_b200/csrc"
OUT=../lib/variants; mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off -I../../include -I."
build() {  # tag, defines...
  tag=$1; shift
  $NV "$@" -Xptxas -v -c queries.cu -o $OUT/queries_$tag.o 2> $OUT/$tag.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libtrgb_kernels_$tag.so core.o map_index.o $OUT/queries_$tag.o node_grid.o sssp.o voxel.o -lcudart
  rm -f $OUT/queries_$tag.o
  grep -A2 "k_collision_tq\|k_edge_collide_tq\|k_sample_window_tq" $OUT/$tag.ptxas.log | grep -E "Used" | tr '\n' ' '; echo " <- $tag"
}
for v in "$@"; do
  case $v in
    old)     build old     -DTQ_FLAT=0 -DTQ_CAP=64 -DTQ_MINBLK=0 ;;
    g1r5)    build g1r5    -DTQ_UNROLL=1 -DTQ_ROWS=5 -DTQ_CAP=48 -DTQ_MINBLK=8 ;;
    g2r5)    build g2r5    -DTQ_UNROLL=2 -DTQ_ROWS=5 -DTQ_CAP=48 -DTQ_MINBLK=8 ;;
    g1r4)    build g1r4    -DTQ_UNROLL=1 -DTQ_ROWS=4 -DTQ_CAP=48 -DTQ_MINBLK=8 ;;
    g2r4)    build g2r4    -DTQ_UNROLL=2 -DTQ_ROWS=4 -DTQ_CAP=48 -DTQ_MINBLK=8 ;;
    g1r5c64) build g1r5c64 -DTQ_UNROLL=1 -DTQ_ROWS=5 -DTQ_CAP=64 -DTQ_MINBLK=6 ;;
    g2r5c64) build g2r5c64 -DTQ_UNROLL=2 -DTQ_ROWS=5 -DTQ_CAP=64 -DTQ_MINBLK=6 ;;
    g3r5c64) build g3r5c64 -DTQ_UNROLL=3 -DTQ_ROWS=5 -DTQ_CAP=64 -DTQ_MINBLK=6 ;;
    w2)      build w2      -DTQ_LD256=1 -DTQ_UNROLL=2 ;;
    w1)      build w1      -DTQ_LD256=1 -DTQ_UNROLL=1 ;;
    n2)      build n2      -DTQ_LD256=0 -DTQ_UNROLL=2 ;;
    w2c48)   build w2c48   -DTQ_LD256=1 -DTQ_UNROLL=2 -DTQ_CAP=48 -DTQ_MINBLK=8 ;;
    g1r5c40) build g1r5c40 -DTQ_UNROLL=1 -DTQ_ROWS=5 -DTQ_CAP=40 -DTQ_MINBLK=10 ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
