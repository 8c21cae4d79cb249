#!/bin/bash
# Integer-path iteration: parity tests, then cfg1 kernel time for a sweep of tile sizes.  bash scripts/gpu_int.sh [passes...]
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_ring_gpu.py -m gpu -q -x 2>&1 | tail -15
for P in "${@:-3}"; do
  echo -n "== cfg1 SDR_INT_PASSES=$P: "
  SDR_INT_PASSES=$P timeout 300 python bench.py --workload cfg1 --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'])"
done
