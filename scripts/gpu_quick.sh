#!/bin/bash
# Quick GPU session: selected tests + the default bench line.  Usage: bash scripts/gpu_quick.sh <tag> "<pytest args>" [bench args]
TAG=${1:-q}
PYT=${2:-tests -m gpu}
BARGS=${3:-}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest $PYT"; timeout 1500 python -m pytest $PYT -q -x 2>&1 | tail -60 | tee $OUT/pytest_$TAG.log
echo "== bench $BARGS"; timeout 900 python bench.py $BARGS 2> $OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_$TAG.json | cut -c1-6000
tail -5 $OUT/bench_$TAG.err
