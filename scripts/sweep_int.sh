#!/bin/bash
# cfg1 sweep of the pipelined integer kernel: bash scripts/sweep_int.sh "P,STAGES,CTAS" ...   (STAGES=0: one CTA per tile)
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_ring_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in "$@"; do
  IFS=, read P S C W <<< "$c"; W=${W:-8}
  echo -n "== P=$P stages=$S ctas/SM=$C warps=$W: "
  SDR_INT_WARPS=$W SDR_INT_PASSES=$P SDR_INT_STAGES=$S SDR_INT_CTAS=$C timeout 300 python bench.py --workload cfg1 --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('kernel_ms',r['kernel_ms'],'frac',r['frac'])"
done
