#!/bin/bash
# cfg1 sweep of the integer batch kernel: bash scripts/sweep_int.sh "PASSES,PERSIST" ...
#   PASSES = tile size in 1024-window passes (SDR_INT_PASSES), PERSIST = resident CTAs per SM (SDR_INT_PERSIST, 0 = one CTA per tile)
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_ring_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in "$@"; do
  IFS=, read P C <<< "$c"
  echo -n "== passes=$P persist=$C: "
  SDR_INT_PASSES=$P SDR_INT_PERSIST=$C timeout 300 python bench.py --workload cfg1 --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('kernel_ms',r['kernel_ms'],'frac',r['frac'])"
done
