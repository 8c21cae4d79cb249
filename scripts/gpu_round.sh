#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines, ncu launch lists + full captures.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag] [quick]
TAG=${1:-r02}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep "Model name" >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench default (the driver's command: headline cfg2 + extra)"; timeout 900 python bench.py 2> $OUT/bench_default_$TAG.err | tail -1 > $OUT/bench_default_$TAG.json; cut -c1-400 $OUT/bench_default_$TAG.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_reference_$TAG.json; cut -c1-300 $OUT/bench_reference_$TAG.json
for W in cfg3 cfg1 chan; do
  echo "== bench $W"; timeout 600 python bench.py --workload $W --no-extra 2>/dev/null | tail -1 > $OUT/bench_${W}_$TAG.json; cut -c1-300 $OUT/bench_${W}_$TAG.json
done
cp $OUT/bench_default_$TAG.json $OUT/bench_cfg2_$TAG.json
[ -n "$QUICK" ] && exit 0
echo "== ncu launch lists"
for W in cfg2 cfg3 cfg1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${W}_$TAG.csv \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/ncu_launch_run_${W}_$TAG.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_chan_$TAG.csv \
    python bench.py --workload chan --steps 1 --warmup 3 > $OUT/ncu_launch_run_chan_$TAG.log 2>&1
echo "== ncu full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_fast -s 3 -c 1 -f -o $OUT/prof_fir_cfg2_$TAG \
    python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/ncu_full_cfg2_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_fast -s 3 -c 1 -f -o $OUT/prof_fir_cfg3_$TAG \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/ncu_full_cfg3_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_demod_d -s 3 -c 1 -f -o $OUT/prof_int_cfg1_$TAG \
    python bench.py --workload cfg1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/ncu_full_cfg1_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chan_bank -s 3 -c 1 -f -o $OUT/prof_chan_$TAG \
    python bench.py --workload chan --steps 1 --warmup 3 > $OUT/ncu_full_chan_$TAG.log 2>&1
ls -la $OUT | tail -30
