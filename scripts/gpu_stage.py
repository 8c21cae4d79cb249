#!/usr/bin/env python3
"""Pageable-input end-to-end rate of sdr_fmrx_process for one SDR_STAGE_THREADS setting (argv[1]); argv[2] = 0 disables the staging."""
import os, sys, time
os.environ["SDR_STAGE_THREADS"] = sys.argv[1]
if len(sys.argv) > 2:
    os.environ["SDR_STAGE_PAGEABLE"] = sys.argv[2]
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] * 1
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import numpy as np
import sdrpkg
S = sdrpkg.load()
from rtl_sdr_rs_b200 import _ffi as F  # noqa: E402
from sigutil import channel_taps, lowpass_taps
n = 1 << 27
iq = np.random.default_rng(1).integers(0, 256, 2 * n, dtype=np.uint8)
h = S.FmRx(channel_taps(127, 75), 75, lowpass_taps(63, 0.45), 1, 1)
out = np.empty(n // 75 + 64, np.float32)
def step():
    return F.check(F.lib().sdr_fmrx_process(h._h, F.ptr(iq), n, None, 0, None, 0, F.ptr(out), out.size))
for _ in range(2):
    step()
t0 = time.perf_counter()
for _ in range(6):
    step()
dt = (time.perf_counter() - t0) / 6
print(f"threads={sys.argv[1]} staging={'on' if len(sys.argv) < 3 else sys.argv[2]}: {n / dt / 1e6:.0f} Msamples/s = {2 * n / dt / 1e9:.1f} GB/s")
