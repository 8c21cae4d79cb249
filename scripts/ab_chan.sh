# A/B harness: SDR_B200_LIB=<lib> selects an alternative build of the same ABI (see _ffi.py)
for lib in "$@"; do
SDR_B200_LIB=$lib python bench.py --workload chan --steps 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',r['kernel_ms_per_slab'],'frac',r['frac'], d['gpu_launches'], d['clocks'])"
done
