#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`): bash scripts/gpu_multi.sh <N> [tag]
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus_${TAG}_n$N.txt
echo "== 1-GPU sanity (tests that changed since the last full round)"
timeout 900 python -m pytest tests/test_fmrx_gpu.py tests/test_demod_gpu.py -m gpu -q 2>&1 | tail -3
echo "== 2-GPU parity test"; timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_multirank_${TAG}_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for W in cfg2 chan; do
  echo "== bench $W N=1"; timeout 600 python bench.py --workload $W --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${W}_${TAG}_n1.json | cut -c1-600
  echo "== bench $W N=$N"; timeout 900 $TR bench.py --gpus $N --workload $W --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${W}_${TAG}_n$N.json | cut -c1-600
done
echo "== reference arm N=$N"; timeout 600 $TR bench.py --gpus $N --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_${TAG}_n$N.json | cut -c1-400
echo "== streaming shell"; ls tests/golden/_ref/capture.bin && (time ./rtl-sdr-rs_b200/bin/simple_fm_b200 tests/golden/_ref/capture.bin | sha256sum) 2>&1 | tail -8 | tee $OUT/shell_${TAG}.log
