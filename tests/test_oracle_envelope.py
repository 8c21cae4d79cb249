"""CPU pin of the oracle on the i32-overflow envelope: for downsample >= 256 the boxcar sums reach
128*D >= 2^15, a*b.conj() leaves i32 (examples/simple_fm.rs:371,378) and fast_atan2 (:383-405) runs on wrapped
operands.  The reference's three KATs do not reach that region, so the C oracle is cross-checked here against an
INDEPENDENT numpy restatement of the same reference lines (int64 arithmetic with explicit 32-bit wraps), and against
its own Vec-per-stage `ref_like` form."""
import numpy as np
import pytest

import oracle_ffi as O
from sigutil import saturated_stream

I32 = 1 << 31


def wrap32(v):
    return ((np.asarray(v, np.int64) + I32) % (1 << 32)) - I32


def np_fast_atan2(y, x):
    """Demod::fast_atan2, :383-405 (release build: wrapping +,-; `as i32` of the i64 product; truncating /).  den == 0
    and INT_MIN / -1 would panic in the reference; the oracle and the library define them as 0 / INT_MIN."""
    y, x = np.asarray(y, np.int64), np.asarray(x, np.int64)
    yabs = np.where(y < 0, wrap32(-y), y)
    xpos = x >= 0
    t = np.where(xpos, wrap32(x - yabs), wrap32(x + yabs))
    num = wrap32(4096 * t)
    den = np.where(xpos, wrap32(x + yabs), wrap32(yabs - x))
    safe = np.where(den == 0, 1, den)
    quo = np.where(den == 0, 0, np.sign(num) * np.sign(safe) * (np.abs(num) // np.abs(safe)))
    quo = np.where((num == -I32) & (den == -1), -I32, quo)
    angle = wrap32(np.where(xpos, 4096, 12288) - quo)
    res = np.where(y < 0, wrap32(-angle), angle)
    return np.where((x == 0) & (y == 0), 0, res)


class NpDemod:
    """struct Demod (:232-427) restated with numpy; one instance carries the five state fields."""

    def __init__(self, D, fast, slow):
        self.D, self.fast, self.slow = D, fast, slow
        self.prev_index, self.lp_now = 0, np.zeros(2, np.int64)
        self.demod_pre = np.zeros(2, np.int64)
        self.now_lpr, self.prev_lpr_index = 0, 0

    def demodulate(self, buf):
        b = np.asarray(buf, np.uint8).copy().reshape(-1, 8)
        r = np.stack([b[:, 0], b[:, 1], 255 - b[:, 3], b[:, 2], 255 - b[:, 4], 255 - b[:, 5], b[:, 7], 255 - b[:, 6]], axis=1)
        v = r.reshape(-1, 2).astype(np.int64) - 127                                   # :258, :441-450
        # low_pass_complex :337-352
        n = v.shape[0]
        csum = np.cumsum(v, axis=0)
        first_end = self.D - self.prev_index                                          # samples that complete window 0
        ends = np.arange(first_end, n + 1, self.D)
        lp = np.empty((ends.size, 2), np.int64)
        if ends.size:
            at = csum[ends - 1]
            lp[0] = wrap32(at[0] + self.lp_now)
            lp[1:] = wrap32(at[1:] - at[:-1])
            self.lp_now = wrap32(csum[-1] - at[-1])
            self.prev_index = n - ends[-1]
        else:
            self.lp_now = wrap32(self.lp_now + csum[-1])
            self.prev_index += n
        assert lp.shape[0] > 1                                                        # :356
        # fm_demod :355-367
        prev = np.vstack([self.demod_pre[None, :], lp[:-1]])
        cre = wrap32(lp[:, 0] * prev[:, 0] + lp[:, 1] * prev[:, 1])
        cim = wrap32(lp[:, 1] * prev[:, 0] - lp[:, 0] * prev[:, 1])
        pcm = np_fast_atan2(cim, cre)
        pcm[0] = int(np.trunc(np.arctan2(float(cim[0]), float(cre[0])) / np.pi * 16384.0))   # :370-374
        dm = wrap32(pcm << 16) >> 16                                                  # `as i16`
        self.demod_pre = lp[-1].copy()
        # low_pass_real :408-426
        out, div = [], self.fast // self.slow
        for x in dm.tolist():
            self.now_lpr = int(wrap32(self.now_lpr + x))
            self.prev_lpr_index += self.slow
            if self.prev_lpr_index < self.fast:
                continue
            q = abs(self.now_lpr) // div
            out.append(int(wrap32((q if self.now_lpr >= 0 else -q) << 16) >> 16))
            self.prev_lpr_index -= self.fast
            self.now_lpr = 0
        return np.asarray(out, np.int16), lp, dm.astype(np.int16)


@pytest.mark.parametrize("D,fast,slow", [(6, 170_000, 32_000), (256, 170_000, 32_000), (300, 96_000, 48_000),
                                         (1000, 50_000, 32_000)])
def test_oracle_matches_numpy_restatement_on_saturated_streams(D, fast, slow):
    rng = np.random.default_rng(D)
    o, o2, r = O.Demod(O.DemodConfig(fast, fast, slow, D, 1)), O.Demod(O.DemodConfig(fast, fast, slow, D, 1)), NpDemod(D, fast, slow)
    wrapped = 0
    for ln in (262144, 8 * 4321, 8 * 2 * D, 262144):
        buf = saturated_stream(rng, ln)
        a, lp, dm = o.demodulate(buf, stages=True)
        ra, rlp, rdm = r.demodulate(buf)
        assert np.array_equal(lp, rlp) and np.array_equal(dm, rdm) and np.array_equal(a, ra), (D, ln)
        assert np.array_equal(o2.demodulate(buf, ref_like=True), a)
        st = o.state()
        assert st["prev_index"] == r.prev_index and st["now_lpr"] == r.now_lpr and st["prev_lpr_index"] == r.prev_lpr_index
        assert st["lp_now"] == tuple(int(x) for x in r.lp_now) and st["demod_pre"] == tuple(int(x) for x in r.demod_pre)
        lp64 = lp.astype(np.int64)
        wrapped += int(np.count_nonzero(np.abs(lp64[1:, 0] * lp64[:-1, 0] + lp64[1:, 1] * lp64[:-1, 1]) >= I32))
    if D >= 256:
        assert wrapped > 0      # the envelope is actually entered


def test_numpy_fast_atan2_equals_oracle_on_wrapped_operands():
    rng = np.random.default_rng(9)
    y = rng.integers(-I32, I32, 20000, dtype=np.int64)
    x = rng.integers(-I32, I32, 20000, dtype=np.int64)
    edge = np.array([0, 1, -1, I32 - 1, -I32, -I32 + 1, 1 << 24, -(1 << 24), 4096, -4096], np.int64)
    y = np.concatenate([y, np.repeat(edge, edge.size)])
    x = np.concatenate([x, np.tile(edge, edge.size)])
    got = np_fast_atan2(y, x)
    want = np.array([O.Demod.fast_atan2(int(a), int(b)) for a, b in zip(y, x)], np.int64)
    assert np.array_equal(got, want)
