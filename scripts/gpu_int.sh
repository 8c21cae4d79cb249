#!/bin/bash
# Integer-path iteration: parity tests, then the cfg1 bench line (kernel time + per-buffer latency) for each ENV combo.
#   bash scripts/gpu_int.sh "A=1 B=2" "A=3" ...
timeout 900 python -m pytest tests/test_demod_gpu.py tests/test_ring_gpu.py -m gpu -q -x 2>&1 | tail -5
for E in "${@:-X=0}"; do
  echo -n "== $E: "
  env $E timeout 300 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('kernel_ms',r['kernel_ms'],'frac',r['frac'],'e2e',d['e2e']['value'],'per_buffer',json.dumps(d.get('per_buffer')))"
done
