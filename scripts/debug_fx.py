"""GPU debug helper: where does low_pass differ from the oracle?  (not a test)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_ffi as O
import sdrpkg
from sigutil import channel_taps
S = sdrpkg.load()
rng = np.random.default_rng(1)
for T, D in ((255, 100), (127, 75), (6, 6)):
    taps = channel_taps(T, D)
    for first, n in ((0, 3000), (54, 2385), (1, 2400), (0, 100 * 700 + 13), (33, 75 * 900)):
        iq = rng.integers(0, 256, 2 * (first + n), dtype=np.uint8)
        g, o = S.FmRx(taps, D), O.FxChain(taps, D)
        if first:
            g.low_pass(iq[:2 * first]); o.process(iq[:2 * first])
        got = g.low_pass(iq[2 * first:]).astype(np.float64)
        want = o.process(iq[2 * first:])[0]
        err = np.abs(got - want).max(axis=1) if want.size else np.zeros(0)
        bad = np.nonzero(err > 1e-3)[0]
        print(f"T={T} D={D} first={first} n={n} outs={want.shape[0]} bad={bad.size} idx={bad[:12].tolist()} maxerr={err.max() if err.size else 0:.3e}")
