#!/bin/bash
# A/B of channeliser builds: bash scripts/ab_chan.sh <tag> <lib> ...   ("default" = the in-tree library); then the channeliser
# parity tests on the in-tree library and one ncu capture of its bank kernel
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for lib in "$@"; do
  if [ "$lib" = default ]; then unset SDR_B200_LIB; else export SDR_B200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --workload chan --steps 5 2> $OUT/ab_err.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$(basename $lib) kernel_ms_per_slab',r['kernel_ms_per_slab'],'ms_per_step',d['ms_per_step'],r['kernel'])" | tee -a $OUT/ab_chan_$TAG.txt || tail -3 $OUT/ab_err.txt
done
unset SDR_B200_LIB
timeout 900 python -m pytest tests/test_chan_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_chan_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chan_bank -s 3 -c 1 -f -o $OUT/prof_chan_bank_$TAG \
    python bench.py --workload chan --steps 1 --warmup 3 > $OUT/ncu_full_chan_bank_$TAG.log 2>&1
python scripts/ncu_summary.py $OUT/prof_chan_bank_$TAG.ncu-rep | tee $OUT/ncu_chan_bank_$TAG.txt | grep -E "duration|issue_active|pipe_fma|warps_active|registers|stalled_(wait|short|long|math|not_sel|no_inst)|inst_executed.sum"
