#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines, ncu launch list + one full capture.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep "Model name" >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
for W in cfg2 cfg3 cfg1; do
  echo "== bench $W"; timeout 600 python bench.py --workload $W 2>&1 | tail -1 | tee $OUT/bench_${W}_$TAG.json
done
echo "== ncu launch list (cfg2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg2_$TAG.csv \
    python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_run_$TAG.log 2>&1
echo "== ncu full capture of the fused kernel (cfg2, 2^28 samples)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_fast -s 3 -c 1 -f -o $OUT/prof_fir_cfg2_$TAG \
    python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --n-log2 28 > $OUT/ncu_full_run_$TAG.log 2>&1
tail -2 $OUT/ncu_full_run_$TAG.log
ls -la $OUT | tail -15
