#!/bin/bash
# f32-receiver iteration: parity tests, then the cfg2 / cfg3 bench lines per environment setting.
#   bash scripts/gpu_fx.sh "A=1" "A=2 B=3" ...
timeout 900 python -m pytest tests/test_fmrx_gpu.py tests/test_chan_gpu.py -m gpu -q -x 2>&1 | tail -5
for E in "${@:-X=0}"; do
  for W in cfg2 cfg3; do
  echo -n "== $W $E: "
  env $E timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'e2e',(d.get('e2e') or {}).get('value'),'launches',d['gpu_launches'])"
  done
done
