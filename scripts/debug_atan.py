import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import test_demod_gpu as T
import sdrpkg; S = sdrpkg.load()
class Rec:
    def __init__(s): s.d = S.Demod()
    def fast_atan2(s, y, x):
        got = s.d.fast_atan2(y, x); want = T.np_fast_atan2(y, x)
        bad = np.nonzero(got != want)[0]
        print("n", len(y), "mismatches", len(bad))
        for i in bad[:30]:
            print(int(y[i]), int(x[i]), "got", int(got[i]), "want", int(want[i]))
        raise SystemExit
class FS:
    def Demod(s): return Rec()
T.test_fast_atan2_and_polar_vs_oracle(FS())
