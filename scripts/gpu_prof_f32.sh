#!/bin/bash
# ncu --set full of the fused FIR kernel at the given (taps,decimation) shapes:  bash scripts/gpu_prof_f32.sh <tag> 129,16 31,10
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
python scripts/gpu_generic.py f32 "$@"       # compiles (NVRTC) and caches the shapes, prints the un-profiled rates
for S in "$@"; do
  N=${S/,/_}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir -s 2 -c 1 -f -o $OUT/prof_f32_${N}_$TAG \
    python scripts/gpu_generic.py f32 $S > $OUT/prof_f32_${N}_$TAG.log 2>&1
  tail -2 $OUT/prof_f32_${N}_$TAG.log
done
