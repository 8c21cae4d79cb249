#!/usr/bin/env python3
"""Throughput of the non-specialised shapes: integer Demod for several `downsample`, f32 receiver for (T, D) that have no
compiled k_fir_fast instance.  Device-resident 2^28 samples, CUDA-event kernel time.  python scripts/gpu_generic.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import sdrpkg
from sigutil import channel_taps
S = sdrpkg.load()
n = 1 << 28
d_in = S.DevBuffer(2 * n)
S.synth_fill_dev(d_in, 2 * n, 0xB2000001)
BUF = 262144
for D, fast, slow in ((6, 170_000, 32_000), (2, 96_000, 48_000), (4, 250_000, 48_000), (8, 125_000, 32_000), (10, 100_000, 32_000),
                      (12, 170_000, 32_000), (3, 334_000, 48_000), (5, 200_000, 32_000), (7, 143_000, 32_000), (9, 112_000, 32_000),
                      (11, 100_000, 32_000), (13, 80_000, 32_000)) + tuple((d, 160_000, 32_000) for d in range(14, 33)) + (
                      (40, 32_000, 32_000), (100, 48_000, 32_000)):
    if len(sys.argv) > 1 and sys.argv[1] == "f32":
        break
    if len(sys.argv) > 2 and str(D) not in sys.argv[2:]:
        continue
    cfg = S.DemodConfig(fast * D, fast, slow, D, 42)
    h = S.Demod(cfg)
    n_bufs = 2 * n // BUF
    cap = (h.out_len(BUF) + 2) * n_bufs + 64
    d_out = S.DevBuffer(2 * cap)
    best = 1e9
    for _ in range(5):
        h.demodulate_batch_dev(d_in, BUF, n_bufs, d_out, cap); h.sync()
        ms, _ = h.last_timing(); best = min(best, ms)
    print(f"int D={D:3d} {fast}->{slow}: {best:.3f} ms  {2.0 * n / best / 1e6:.0f} GB/s")
    d_out.free(); h.close()
if len(sys.argv) > 1 and sys.argv[1] == "int":
    sys.exit(0)
F32_SHAPES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] if len(sys.argv) > 2 and sys.argv[1] == "f32" else None
for T, D in F32_SHAPES or ((127, 75), (255, 100), (127, 50), (63, 25), (201, 64), (511, 100), (31, 10), (129, 16), (65, 32), (127, 48), (255, 96),
             (33, 8), (127, 40), (200, 3)):
    taps = channel_taps(T, D)
    h = S.FmRx(taps, D, None, 1, 1)
    _, na = h.out_lens(n)
    d_out = S.DevBuffer(4 * (na + n // D + 64))
    best = 1e9
    for _ in range(5):
        h.timing_totals(reset=True)
        h.process_dev(d_in, n, d_out, na + n // D + 64); h.sync()
        sums, calls = h.timing_totals(); best = min(best, sums[0] / max(calls, 1))
    print(f"f32 T={T:3d} D={D:3d} kind={h.kernel_kind()[0]}: {best:.3f} ms  {2.0 * n / best / 1e6:.0f} GB/s")
    d_out.free(); h.close()
