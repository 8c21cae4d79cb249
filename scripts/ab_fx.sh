#!/bin/bash
# A/B of environment settings on the f32 receiver's bench lines: bash scripts/ab_fx.sh <tag> "ENV=VAL" ...   ("X=0" = default)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for E in "$@"; do
  for W in cfg2 cfg3; do
  env $E timeout 300 python bench.py --workload $W --no-cpu-baseline --no-e2e --no-extra --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$E $W ms_per_step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'step_frac',r['whole_step_frac'])" | tee -a $OUT/ab_fx_$TAG.txt
  done
done
