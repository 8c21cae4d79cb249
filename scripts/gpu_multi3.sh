#!/bin/bash
# Driver-shaped multi-GPU bench lines only (no tests): bash scripts/gpu_multi3.sh <N> <tag>
N=${1:-8}; TAG=${2:-r02z}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_default_${TAG}_n$N.json 2> $OUT/bench_default_${TAG}_n$N.err
tail -1 $OUT/bench_default_${TAG}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks')})); print(d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step']); print(json.dumps(d['multi_gpu'], indent=0)); print(d['e2e']['value'], d['e2e']['fraction_of_h2d_ceiling'])"
tail -2 $OUT/bench_default_${TAG}_n$N.err
timeout 600 $TR bench.py --gpus $N --workload chan --steps 20 2>/dev/null | tail -1 > $OUT/bench_chan_${TAG}_n$N.json
python -c "import sys,json; d=json.loads(open('$OUT/bench_chan_${TAG}_n$N.json').read()); print(d['ms_per_step'], d['value'], json.dumps(d['multi_gpu']))"
