#!/bin/bash
# A/B of integer-path builds on cfg1: [PASSES="2 3"] [ENVS="A=1 B=2"] bash scripts/ab_int.sh <lib> ...   (each also runs the parity tests first)
for lib in "$@"; do
  export SDR_B200_LIB=$lib
  ok=$(timeout 600 python -m pytest tests/test_demod_gpu.py -m gpu -q -x 2>&1 | tail -1)
  for P in ${PASSES:-3}; do
  for E in ${ENVS:-X=0}; do
  env $E SDR_INT_PASSES=$P timeout 300 python bench.py --workload cfg1 --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$(basename $lib) P=$P $E kernel_ms',r['kernel_ms'],'frac',r['frac'],'| $ok')"
  done
  done
done
