#!/bin/bash
# N-GPU bench lines only (run under `gpurun --gpus N`): bash scripts/gpu_scale.sh <N> <tag>
N=${1:-8}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/gpus_${TAG}_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for W in cfg2 cfg1 chan; do
  echo "== bench $W N=$N"; timeout 900 $TR bench.py --gpus $N --workload $W --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${W}_${TAG}_n$N.json | cut -c1-400
done
echo "== reference arm N=$N"; timeout 600 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_${TAG}_n$N.json | cut -c1-300
