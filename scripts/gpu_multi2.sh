#!/bin/bash
# Round-2 multi-GPU session (run under `gpurun --gpus N`): bash scripts/gpu_multi2.sh <N> <tag>
# The driver-shaped default bench line at N ranks (carries `extra` and `multi_gpu`), the 2-rank parity test, NCCL's own log.
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/gpus_${TAG}_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== 2-GPU parity test"; timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_multirank_${TAG}_n$N.log
echo "== default bench N=$N"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/nccl_${TAG}_n$N.%p.log timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_default_${TAG}_n$N.json 2> $OUT/bench_default_${TAG}_n$N.err
tail -1 $OUT/bench_default_${TAG}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','multi_gpu')}, indent=1)); print(json.dumps(d['extra']['chan'], indent=1))"
tail -3 $OUT/bench_default_${TAG}_n$N.err
grep -h -i "broadcast\|NVLS\|Channel 00\|via" $OUT/nccl_${TAG}_n$N.*.log | head -12 > $OUT/nccl_${TAG}_n${N}_summary.txt; rm -f $OUT/nccl_${TAG}_n$N.*.log; cat $OUT/nccl_${TAG}_n${N}_summary.txt
echo "== chan bench N=$N"; timeout 900 $TR bench.py --gpus $N --workload chan --steps 10 2>&1 | tail -1 | tee $OUT/bench_chan_${TAG}_n$N.json | cut -c1-1500
