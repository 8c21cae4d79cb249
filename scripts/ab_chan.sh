for c in dcb9003 dfdebbb 8bbefcc dcb9003 8bbefcc; do
SDR_B200_LIB=rtl-sdr-rs_b200/lib/ab/lib_$c.so python bench.py --workload chan --steps 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$c value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms_per_slab'],'frac',r['frac'], d['gpu_launches'], d['clocks'])"
done
python -m pytest tests/test_chan_gpu.py -m gpu -q -k device_resident 2>&1 | grep -E "assert|where" | head
