#!/bin/bash
# Channel-shard bench at N ranks for several NCCL CTA budgets: bash scripts/gpu_bcast_sweep.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== 2-GPU parity test"; timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q 2>&1 | tail -4
for C in default 2 4 8 16; do
  if [ $C = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$C; fi
  timeout 600 $TR bench.py --gpus $N --workload chan --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); m=d['multi_gpu']
print('NCCL_MAX_CTAS=$C', {k:m[k] for k in ('ms_per_step','ms_per_step_without_bcast','exposed_bcast_ms','bcast_ms_per_slab','bcast_gbs','kernel_ms_per_slab','input_msamples_per_s')})" | tee -a $OUT/bcast_sweep_${TAG}_n$N.txt
done
