#!/bin/bash
# ncu full captures of selected kernels: bash scripts/gpu_prof.sh <tag> <workload:kernel-regex> ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  W=${spec%%:*}; K=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_${W}_${K}_$TAG \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_${W}_${K}_$TAG.log 2>&1
  tail -1 $OUT/ncu_full_${W}_${K}_$TAG.log | cut -c1-200
done
