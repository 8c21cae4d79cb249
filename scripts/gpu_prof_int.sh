#!/bin/bash
# ncu --set full of the integer kernel at the given downsamples:  bash scripts/gpu_prof_int.sh <tag> 2 3
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for D in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_demod -s 2 -c 1 -f -o $OUT/prof_int_d${D}_$TAG \
    python scripts/gpu_generic.py int $D > $OUT/prof_int_d${D}_$TAG.log 2>&1
  tail -2 $OUT/prof_int_d${D}_$TAG.log
done
python scripts/gpu_generic.py int "$@"
