"""GPU parity tests of the f32 tap'd-FIR receiver (sdr_fmrx_*) against the f64 oracle.

Bar (BASELINE.json north_star): within 1e-5 relative.  The oracle for this path is "parity
unpinned" by the reference (there is no tap'd FIR in it); where the two paths coincide (boxcar
taps) the f32 path is additionally checked bit-exactly against the PINNED integer path.
"""
import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg
from sigutil import (assert_angle_close, assert_close, assert_demod_propagated, channel_taps, disc_f64, fm_test_signal,
                     lowpass_taps, rel_err)

pytestmark = pytest.mark.gpu
GAIN = 16384.0 / np.pi


@pytest.fixture(scope="module")
def S():
    m = sdrpkg.load()
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return m


def ragged_cuts(n, rng, pieces):
    cuts = np.sort(rng.choice(np.arange(1, n), size=pieces - 1, replace=False))
    return [0, *cuts.tolist(), n]


SHAPES = [  # (T, D, kernel kind: 1 = pre-compiled k_fir_fast, 2 = k_fir_fast compiled by NVRTC for the shape,
           #  3 = k_fir_slide (output-owner kernel, NVRTC): more than 16 lags per sample, or a decimation up to 4 with 8+ lags; 0 = generic)
    (127, 75, 1), (255, 100, 1), (6, 6, 1),
    (63, 20, 2), (31, 7, 2), (1, 1, 2), (7, 3, 2), (5, 64, 2), (127, 50, 2), (201, 64, 2), (33, 125, 2), (129, 16, 2), (65, 32, 2),
    (200, 3, 3), (65, 4, 3), (31, 2, 3), (255, 8, 3), (16, 1, 3),
    (300, 301, 0), (1001, 250, 0),   # decim > 256 / too many taps for either unrolled kernel: generic kernel
]


@pytest.fixture()
def no_rtc(monkeypatch):
    monkeypatch.setenv("SDR_FIR_RTC", "0")


@pytest.mark.parametrize("T,D,kind", SHAPES)
def test_low_pass_streaming_vs_oracle(S, T, D, kind):
    rng = np.random.default_rng(T * 1000 + D)
    taps = channel_taps(T, D) if T > 1 else np.ones(1, np.float32)
    n = max(40 * D, 3 * T) + 12345
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    g = S.FmRx(taps, D)
    assert g.kernel_kind()[0] == kind, g.kernel_kind()
    assert g.last_timing()[2] == int(kind != 0)
    o = O.FxChain(taps, D)
    # ragged calls: odd lengths exercise both 4-byte phases, tiny calls exercise the carry
    cuts = [0, 1, 2, 3 + D // 2, 3 + D // 2 + 1] + [c for c in ragged_cuts(n, rng, 7)[1:-1] if c > 3 + D // 2 + 1] + [n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        want, _, _ = o.process(iq[2 * lo:2 * hi])
        got = g.low_pass(iq[2 * lo:2 * hi])
        assert got.shape == want.shape, (lo, hi)
        assert_close(got, want, what=f"low_pass T={T} D={D} [{lo},{hi})")


@pytest.mark.parametrize("T,D", [(63, 20), (31, 7), (5, 64), (127, 50)])
def test_generic_kernel_matches_oracle_when_rtc_is_off(S, no_rtc, T, D):
    """SDR_FIR_RTC=0: shapes without a pre-compiled instance run k_fir_generic (the path a box without libnvrtc takes)."""
    rng = np.random.default_rng(T + D)
    taps = channel_taps(T, D)
    n = 50 * D + 4321
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    g = S.FmRx(taps, D)
    assert g.kernel_kind()[0] == 0
    o = O.FxChain(taps, D)
    for lo, hi in ((0, n // 3), (n // 3, n // 3 + 1), (n // 3 + 1, n)):
        want, _, _ = o.process(iq[2 * lo:2 * hi])
        assert_close(g.low_pass(iq[2 * lo:2 * hi]), want, what=f"generic T={T} D={D}")


@pytest.mark.parametrize("T,D", [(127, 75), (255, 100)])
def test_rtc_instance_is_bit_identical_to_the_precompiled_one(S, monkeypatch, T, D):
    """SDR_FIR_RTC=force re-compiles a pre-compiled shape with NVRTC (same CTA shape, same source): every output bit
    of y and of the discriminator must agree, for ragged calls that exercise every load phase."""
    rng = np.random.default_rng(T * 7 + D)
    taps = channel_taps(T, D)
    n = 300 * D + 777
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    a = S.FmRx(taps, D)
    monkeypatch.setenv("SDR_FIR_RTC", "force")
    b = S.FmRx(taps, D)
    assert a.kernel_kind()[0] == 1 and b.kernel_kind()[0] == 2, (a.kernel_kind(), b.kernel_kind())
    cuts = [0, 1, 3, 10, 10 + D, 11 + 3 * D, n // 2 + 1, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ya, da, _ = a.process(iq[2 * lo:2 * hi])
        yb, db, _ = b.process(iq[2 * lo:2 * hi])
        assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32)), (lo, hi)
        assert np.array_equal(da.view(np.uint32), db.view(np.uint32)), (lo, hi)


@pytest.mark.parametrize("T,D", [(127, 75), (255, 100), (63, 20), (129, 16), (65, 4), (200, 3)])
def test_streaming_is_bitwise_chunk_invariant(S, T, D):
    """Fixed-order partial sums: how the stream is cut into calls must not change a single bit."""
    rng = np.random.default_rng(T)
    taps = channel_taps(T, D)
    n = 300 * D + 77
    iq = fm_test_signal(n, fs=2.4e6, seed=T)
    one = S.FmRx(taps, D).process(iq)
    g = S.FmRx(taps, D)
    parts = [g.process(iq[2 * lo:2 * hi]) for lo, hi in zip(*(lambda c: (c[:-1], c[1:]))(ragged_cuts(n, rng, 9)))]
    for idx, name in ((0, "y"), (1, "demod"), (2, "audio")):
        cat = np.concatenate([p[idx] for p in parts])
        assert np.array_equal(cat, one[idx]), name


def test_boxcar_taps_equal_the_pinned_integer_path(S):
    """T=D=6, h=1: low_pass == Demod::low_pass_complex (examples/simple_fm.rs:337-352) on centred samples."""
    rng = np.random.default_rng(66)
    iq = rng.integers(0, 256, 2 * 60_001, dtype=np.uint8)
    g = S.FmRx(np.ones(6, np.float32), 6)
    d = S.Demod()
    for lo, hi in ((0, 1000), (1000, 1003), (1003, 60_001)):
        y = g.low_pass(iq[2 * lo:2 * hi])
        lp = d.low_pass_complex(d.buf_to_complex(iq[2 * lo:2 * hi]))
        assert np.array_equal(y, lp.astype(np.float32))


@pytest.mark.parametrize("T,D,fs", [(127, 75, 2.4e6), (255, 100, 20e6), (63, 20, 1.0e6), (65, 4, 1.0e6), (200, 3, 1.0e6)])
def test_fused_chain_on_fm_parity_signal(S, T, D, fs):
    """SURVEY §8d parity set: FM test signal, 75 kHz deviation, 1 kHz tone, N(0, 8^2) noise."""
    n = 1 << 20
    iq = fm_test_signal(n, fs=fs, f_dev=75e3 if fs > 2e6 else 30e3)
    taps = channel_taps(T, D)
    fs_out = fs / D
    up, down = (1, 1) if abs(fs_out - 32e3) < 1 else ((4, 25) if abs(fs_out - 200e3) < 1 else (16, 25))
    taps2 = lowpass_taps(32 * up - 1 if up > 1 else 63, 0.45 / max(up, down), gain=up)
    g, o = S.FmRx(taps, D, taps2, up, down), O.FxChain(taps, D, taps2, up, down)
    y, d, a = g.process(iq)
    yo, do, ao = o.process(iq)
    assert y.shape[0] == n // D and a.shape == ao.shape
    assert_close(y, yo, what="y")
    # the discriminator stage itself: against an f64 discriminator of the GPU's own y
    assert_angle_close(d, disc_f64(y, GAIN), GAIN * np.pi, what="demod stage")
    # end to end: the first-order bound that a 1e-5 error in y allows
    assert_demod_propagated(d, yo, do, GAIN, what="demod end-to-end")
    # the resampler stage itself: against an f64 polyphase FIR of the GPU's own discriminator output
    import ctypes as C
    oo = O.FxChain(taps, D, taps2, up, down)
    d64 = np.ascontiguousarray(d, np.float64)
    buf = np.empty(ao.size + 4, np.float64)
    na = O.lib().orc_fx_resample(C.byref(oo.s), O._p(d64, C.c_double), d64.size, O._p(buf, C.c_double))
    assert_close(a, buf[:na], what="resample stage")
    print(f"\n[T={T} D={D}] rel_err y={rel_err(y, yo):.2e}  demod={rel_err(d, do):.2e}  audio={rel_err(a, ao):.2e}")
    # y never needs to leave the chip: same audio without asking for y / demod
    g2 = S.FmRx(taps, D, taps2, up, down)
    _, _, a2 = g2.process(iq, want_y=False, want_demod=False)
    assert np.array_equal(a, a2)


def test_stage_entry_points_fm_demod_and_resample(S):
    rng = np.random.default_rng(8)
    taps = channel_taps(63, 20)
    for up, down, t2 in ((1, 1, 63), (4, 25, 127), (1, 5, 31), (3, 2, 64)):
        taps2 = lowpass_taps(t2, 0.45 / max(up, down), gain=up)
        g, o = S.FmRx(taps, 20, taps2, up, down), O.FxChain(taps, 20, taps2, up, down)
        y = (rng.standard_normal((5000, 2)) * 50).astype(np.float32)
        import ctypes as C
        got_d, got_a, want_d, want_a = [], [], [], []
        for lo, hi in ((0, 1), (1, 14), (14, 2000), (2000, 5000)):
            seg = y[lo:hi]
            gd = g.fm_demod(seg)
            got_d.append(gd)
            got_a.append(g.resample(gd))
            seg64 = np.ascontiguousarray(seg, np.float64)
            od = np.empty(hi - lo, np.float64)
            O.lib().orc_fx_fm_demod(C.byref(o.s), O._p(seg64, C.c_double), hi - lo, O._p(od, C.c_double))
            oa = np.empty((hi - lo) * up // down + 3, np.float64)
            na = O.lib().orc_fx_resample(C.byref(o.s), O._p(od, C.c_double), hi - lo, O._p(oa, C.c_double))
            want_d.append(od), want_a.append(oa[:na])
        assert_angle_close(np.concatenate(got_d), np.concatenate(want_d), GAIN * np.pi, what=f"fm_demod {up}/{down}")
        ga, wa = np.concatenate(got_a), np.concatenate(want_a)
        assert ga.shape == wa.shape == (-(-5000 * up // down),)
        # the resampler is linear: feed it the oracle's own demod values to isolate it
        g2 = S.FmRx(taps, 20, taps2, up, down)
        assert_close(g2.resample(np.concatenate(want_d).astype(np.float32)), wa, rtol=2e-5, what=f"resample {up}/{down}")


def test_errors_are_codes_not_truncation(S):
    g = S.FmRx(channel_taps(127, 75), 75)
    iq = np.zeros(2 * 7500, np.uint8)
    import ctypes as C
    from rtl_sdr_rs_b200 import _ffi as F
    out = np.empty((10, 2), np.float32)
    rc = F.lib().sdr_fmrx_low_pass(g._h, F.ptr(iq), 7500, F.ptr(out), 10)
    assert rc == -3                                     # SDR_E_CAP
    assert g.out_lens(7500) == (100, 100)              # and the failed call consumed nothing
    with pytest.raises(S.SdrError):
        S.FmRx(np.ones(4, np.float32), 0)
    with pytest.raises(S.SdrError):
        S.FmRx(np.ones(4, np.float32), 2, np.ones(3, np.float32), 0, 1)


def test_device_resident_full_size_properties(S):
    """2^28 complex samples resident in HBM (config 2 shape): chunk-invariance and prefix parity."""
    T, D = 127, 75
    n = 1 << 28
    seed = 0xB2000001
    taps = channel_taps(T, D)
    taps2 = lowpass_taps(63, 0.45)
    d_iq = S.DevBuffer(2 * n)
    S.synth_fill_dev(d_iq, 2 * n, seed)
    g = S.FmRx(taps, D, taps2, 1, 1)
    ny, na = g.out_lens(n)
    assert ny == na == n // D
    d_a, d_d = S.DevBuffer(4 * na), S.DevBuffer(4 * ny)
    assert g.process_dev(d_iq, n, d_a, na, d_demod=d_d) == na
    g.sync()
    ms, launches, spec = g.last_timing()
    assert spec == 1 and ms[0] > 0 and launches == 2      # fused convert+FIR+demod (carry and history folded in) + audio FIR
    audio, demod = d_a.download(np.float32, na), d_d.download(np.float32, ny)
    # (1) prefix parity against the oracle
    k = 300 * D
    yo, do, ao = O.FxChain(taps, D, taps2, 1, 1).process(O.synth_fill(2 * k, seed))
    assert_demod_propagated(demod[: k // D], yo, do, GAIN, what="prefix demod")
    # (2) the same stream in three ragged device-resident calls is bit-identical
    g2 = S.FmRx(taps, D, taps2, 1, 1)
    cuts = [0, 8 * 12_345_67, 8 * 20_000_001, n]     # 16-byte aligned offsets, arbitrary phase mod 75
    got = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        m = g2.process_dev(d_iq, hi - lo, d_a, na, iq_offset=2 * lo)
        g2.sync()
        got.append(d_a.download(np.float32, m))
    assert np.array_equal(np.concatenate(got), audio)
    # (3) linearity of the whole FIR stage in the taps: y(2h) == 2*y(h) exactly in f32
    g3, g4 = S.FmRx(taps, D), S.FmRx((2 * taps).astype(np.float32), D)
    d_y = S.DevBuffer(8 * 4096)
    small = 4096 * D
    g3.process_dev(d_iq, small, d_d, ny, d_y=d_y); g3.sync()
    y1 = d_y.download(np.float32, 2 * 4096)
    g4.process_dev(d_iq, small, d_d, ny, d_y=d_y); g4.sync()
    assert np.array_equal(d_y.download(np.float32, 2 * 4096), 2 * y1)
    for b in (d_iq, d_a, d_d, d_y):
        b.free()
