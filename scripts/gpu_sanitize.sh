#!/bin/bash
# compute-sanitizer over the parity tests that fit (small sizes; the persistent ring and the 1 GiB runs are excluded).
#   bash scripts/gpu_sanitize.sh <tag>
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
SEL='not full_golden and not device_resident and not full_size and not ring'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 \
  python -m pytest tests/test_demod_gpu.py tests/test_fmrx_gpu.py tests/test_chan_gpu.py -m gpu -q -x -k "$SEL and not wide_downsamples and not every_tile" > $OUT/sanitizer_memcheck_$TAG.log 2>&1
echo "memcheck rc=$?"; tail -4 $OUT/sanitizer_memcheck_$TAG.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 \
  python -m pytest tests/test_demod_gpu.py tests/test_fmrx_gpu.py -m gpu -q -x -k "$SEL and (ragged or kat or streaming or rtc or carried)" > $OUT/sanitizer_racecheck_$TAG.log 2>&1
echo "racecheck rc=$?"; tail -4 $OUT/sanitizer_racecheck_$TAG.log
# round 2: the bank channeliser, the wide-downsample passes (a few of them) and the post-stages under memcheck as well
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 \
  python -m pytest tests/test_chan_gpu.py tests/test_demod_gpu.py -m gpu -q -x -k "bank_kernel or post_stages or (wide_downsamples and (15 or 16 or 21 or 30 or 32) and not 8-)" > $OUT/sanitizer_memcheck_b_$TAG.log 2>&1
echo "memcheck (round-2 kernels) rc=$?"; tail -4 $OUT/sanitizer_memcheck_b_$TAG.log
