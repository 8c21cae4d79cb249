#!/bin/bash
# CTA-shape sweep of the fused FIR kernel: bash scripts/gpu_tune.sh
OUT=gpurun_out; mkdir -p $OUT
run() { W=$1; NT=$2; B=$3; echo -n "== $W NT=$NT B=$B: "; SDR_FIR_NT=$NT SDR_FIR_B=$B timeout 300 python bench.py --workload $W --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'])"; }
run cfg2 32 2; run cfg2 64 2; run cfg2 32 4; run cfg2 64 4
run cfg3 128 1; run cfg3 96 1; run cfg3 160 1; run cfg3 64 1; run cfg3 64 2; run cfg3 32 2
