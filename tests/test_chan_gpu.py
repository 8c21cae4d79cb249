"""GPU parity tests of the wideband channeliser (sdr_chan_*) against the f64 oracle's DIRECT definition
(mix by a 32-bit-phase NCO, then FIR/decimate, then discriminate).  Extension path: parity unpinned by the
reference; bar = 1e-5 relative.  The kernel folds the NCO into per-channel complex taps, so this also
checks that algebra."""
import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg
from sigutil import assert_angle_close, assert_close, assert_demod_propagated, channel_taps, disc_f64, fm_test_signal

pytestmark = pytest.mark.gpu
GAIN = 16384.0 / np.pi


@pytest.fixture(scope="module")
def S():
    m = sdrpkg.load()
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return m


def freq_words(offsets_hz, fs):
    return (np.round(np.asarray(offsets_hz, np.float64) / fs * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)


@pytest.mark.parametrize("C,T,D", [(64, 255, 100), (3, 31, 7), (130, 63, 20), (64, 127, 75), (8, 16, 4)])
def test_channeliser_vs_direct_definition(S, C, T, D):
    rng = np.random.default_rng(C * 100 + D)
    fs = 20e6
    taps = channel_taps(T, D)
    fw = freq_words((np.arange(C) - (C - 1) / 2) * (fs / max(C, 2)) * 0.9, fs)
    n = D * 150 + 37
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    ch = S.Channeliser(taps, D, fw)
    y, d = ch.process(iq)
    yo, do = O.channelise(iq, taps, D, fw)
    assert y.shape == yo.shape and d.shape == do.shape == (C, n // D)
    for c in range(C):
        assert_close(y[c], yo[c], what=f"y ch{c}")
        assert_angle_close(d[c], disc_f64(y[c], GAIN), GAIN * np.pi, what=f"demod stage ch{c}")
        assert_demod_propagated(d[c], yo[c], do[c], GAIN, what=f"demod ch{c}")


def test_streaming_chunks_are_bitwise_identical(S):
    rng = np.random.default_rng(4)
    C, T, D = 64, 255, 100
    taps = channel_taps(T, D)
    fw = freq_words((np.arange(C) - 31.5) * 200e3, 20e6)
    n = D * 400 + 61
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    y1, d1 = S.Channeliser(taps, D, fw).process(iq)
    ch = S.Channeliser(taps, D, fw)
    parts = [ch.process(iq[2 * lo:2 * hi]) for lo, hi in ((0, 1), (1, 6777), (6777, 6800), (6800, n))]
    y2 = np.concatenate([p[0] for p in parts], axis=1)
    d2 = np.concatenate([p[1] for p in parts], axis=1)
    assert np.array_equal(y1, y2) and np.array_equal(d1, d2)


def test_zero_offset_channel_equals_single_channel_receiver(S):
    """fw = 0: the channeliser's channel is the plain FIR of the f32 receiver (same taps, same input)."""
    T, D = 255, 100
    taps = channel_taps(T, D)
    iq = fm_test_signal(D * 500, fs=20e6, f_dev=40e3)
    y, d = S.Channeliser(taps, D, np.zeros(2, np.uint32)).process(iq)
    y1, d1, _ = S.FmRx(taps, D).process(iq)
    assert_close(y[0], y1, what="fw=0 channel vs FmRx")
    assert np.array_equal(y[0], y[1])


def test_fm_channels_are_recovered(S):
    """Two FM carriers at +-1 MHz in a 20 Msps stream come out of their channels with the 1 kHz tone."""
    fs, D, T = 20e6, 100, 255
    n = D * 4000
    a = fm_test_signal(n, fs, seed=1, f_c=1.0e6, f_dev=50e3, amp=50, noise=2).astype(np.float64) - 127.5
    b = fm_test_signal(n, fs, seed=2, f_c=-1.0e6, f_dev=50e3, f_mod=2e3, amp=50, noise=2).astype(np.float64) - 127.5
    iq = np.clip(np.rint(a + b + 127.5), 0, 255).astype(np.uint8)
    taps = channel_taps(T, D)
    _, d = S.Channeliser(taps, D, freq_words([1.0e6, -1.0e6, 3.0e6], fs)).process(iq, want_y=False)
    spec = np.abs(np.fft.rfft(d[:, 200:] * np.hanning(d.shape[1] - 200), axis=1))
    f = np.fft.rfftfreq(d.shape[1] - 200, D / fs)
    assert abs(f[np.argmax(spec[0][1:]) + 1] - 1e3) < 150
    assert abs(f[np.argmax(spec[1][1:]) + 1] - 2e3) < 150
    assert spec[2].max() < 0.2 * spec[0].max()      # empty channel: no tone


def test_device_resident_and_errors(S):
    C, T, D = 64, 255, 100
    taps = channel_taps(T, D)
    fw = freq_words((np.arange(C) - 31.5) * 200e3, 20e6)
    n = 1 << 22
    d_iq = S.DevBuffer(2 * n)
    S.synth_fill_dev(d_iq, 2 * n, 7)
    ch = S.Channeliser(taps, D, fw)
    cap = n // D + 1
    d_d = S.DevBuffer(4 * C * cap)
    m = ch.process_dev(d_iq, n, d_d, cap)
    ch.sync()
    ms, launches = ch.last_timing()
    assert m == n // D and ms > 0 and launches == 2              # cfg4 plan: one bank launch (discriminator fused) + carry
    d = d_d.download(np.float32, C * cap).reshape(C, cap)[:, :m]
    _, d_host = S.Channeliser(taps, D, fw).process(O.synth_fill(2 * D * 300, 7), want_y=False)
    assert np.array_equal(d[:, :300], d_host)
    with pytest.raises(S.SdrError) as e:
        ch.process_dev(d_iq, n, d_d, 10)
    assert e.value.code == -3
    with pytest.raises(S.SdrError):
        S.Channeliser(taps, 0, fw)
    d_iq.free(), d_d.free()


# ---- two-stage polyphase bank (k_chan_bank): uniformly spaced channels --------------------------------------------------
BANK_CASES = [  # (name, C, T, D, offsets(C, fs), K)
    ("cfg4", 64, 255, 100, lambda C, fs: (np.arange(C) - 31.5) * 200e3, 100),
    ("cfg5-interleaved-rank3of8", 64, 255, 100, lambda C, fs: ((3 + 8 * np.arange(C)) - 255.5) * (fs / 512), 64),
    ("100-of-128", 100, 63, 20, lambda C, fs: (np.arange(C) - 50) * (fs / 128), 128),
    ("K16-aliasing-odd-D", 64, 31, 7, lambda C, fs: np.arange(C) * (fs / 16), 16),
    ("prime-K53", 64, 255, 50, lambda C, fs: (np.arange(C) - 20) * (fs / 53), 53),
    ("K96-K2=8", 70, 127, 75, lambda C, fs: (np.arange(C) - 10) * (fs / 96), 96),
]


@pytest.mark.parametrize("name,C,T,D,offs,K", BANK_CASES)
def test_bank_kernel_vs_direct_definition(S, name, C, T, D, offs, K):
    """Uniform grids take k_chan_bank; outputs against the oracle's direct NCO-mix definition, ragged streaming calls."""
    fs = 20e6
    taps = channel_taps(T, D)
    fw = freq_words(offs(C, fs), fs)
    n = D * 300 + 37
    iq = np.random.default_rng(K).integers(0, 256, 2 * n, dtype=np.uint8)
    ch = S.Channeliser(taps, D, fw)
    kind, info = ch.kernel_kind()
    assert kind == 2 and info[0] == K and info[1] * info[2] == K, (kind, info)
    cuts = [0, 1, D - 1, D + 3, 127 * D + 5, 128 * D, n]        # first tile boundary, sub-decimation calls, carry
    parts = [ch.process(iq[2 * lo:2 * hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
    y = np.concatenate([p[0] for p in parts], axis=1)
    d = np.concatenate([p[1] for p in parts], axis=1)
    yo, do = O.channelise(iq, taps, D, fw)
    assert y.shape == yo.shape and d.shape == do.shape == (C, n // D)
    for c in range(C):
        assert_close(y[c], yo[c], what=f"{name} y ch{c}")
        assert_angle_close(d[c], disc_f64(y[c], GAIN), GAIN * np.pi, what=f"{name} demod stage ch{c}")
        assert_demod_propagated(d[c], yo[c], do[c], GAIN, what=f"{name} demod ch{c}")
    from sigutil import rel_err
    assert rel_err(y, yo) <= 1e-5
    # one call == ragged calls, bit for bit; and the discriminator does not depend on whether y is asked for
    ch1 = S.Channeliser(taps, D, fw)
    y1, d1 = ch1.process(iq)
    assert np.array_equal(y1, y) and np.array_equal(d1, d)
    ch2 = S.Channeliser(taps, D, fw)
    _, d2 = ch2.process(iq, want_y=False)
    assert np.array_equal(d2, d)


def test_bank_and_direct_kernels_agree(S, monkeypatch):
    fs, C, T, D = 20e6, 64, 255, 100
    taps = channel_taps(T, D)
    fw = freq_words((np.arange(C) - 31.5) * 200e3, fs)
    iq = fm_test_signal(D * 2000, fs=fs, f_c=500e3)
    a = S.Channeliser(taps, D, fw)
    monkeypatch.setenv("SDR_CHAN_BANK", "0")
    b = S.Channeliser(taps, D, fw)
    assert a.kernel_kind()[0] == 2 and b.kernel_kind()[0] == 1
    ya, da = a.process(iq)
    yb, db = b.process(iq)
    from sigutil import rel_err
    assert rel_err(ya, yb) < 3e-6, rel_err(ya, yb)
    strong = np.hypot(yb[..., 0], yb[..., 1]) > 1e-2 * np.abs(yb).max()
    strong[:, 1:] &= strong[:, :-1]
    dd = (da - db + GAIN * np.pi) % (2 * GAIN * np.pi) - GAIN * np.pi
    assert np.abs(dd[strong]).max() < 1e-4 * GAIN * np.pi


def test_bank_device_resident_slab(S):
    """The bench shape: 2^22 samples resident, demod only (y never leaves the chip), two back-to-back slabs."""
    fs, C, T, D = 20e6, 64, 255, 100
    taps = channel_taps(T, D)
    fw = freq_words((np.arange(C) - 31.5) * 200e3, fs)
    n = 1 << 22
    d_in = S.DevBuffer(2 * n)
    S.synth_fill_dev(d_in, 2 * n, 0xB2000001)
    cap = n // D + 1
    d_d = S.DevBuffer(4 * C * cap)
    ch = S.Channeliser(taps, D, fw)
    m1 = ch.process_dev(d_in, n // 2, d_d, cap)
    ch.sync()
    first = d_d.download(np.float32, C * cap).reshape(C, cap)[:, :m1].copy()
    from rtl_sdr_rs_b200 import _ffi as F
    m2 = F.check(F.lib().sdr_chan_process_dev(ch._h, d_in.at(n), n // 2, None, d_d.ptr, cap))
    ch.sync()
    second = d_d.download(np.float32, C * cap).reshape(C, cap)[:, :m2].copy()
    assert m1 + m2 == n // D
    k = 200 * D
    yo, do = O.channelise(O.synth_fill(2 * k, 0xB2000001), taps, D, fw)
    for c in (0, 17, 63):
        assert_demod_propagated(first[c, :200], yo[c], do[c], GAIN, what=f"resident ch{c}")
    # the same stream in one host call
    host = S.Channeliser(taps, D, fw).process(O.synth_fill(2 * n, 0xB2000001), want_y=False)[1]
    assert np.array_equal(np.concatenate([first, second], axis=1), host)
    ms, launches = ch.last_timing()
    assert launches == 2 and ms > 0          # one bank launch + the carry update


@pytest.mark.parametrize("plan", ["cfg4", "cfg5-interleaved", "non-uniform"])
def test_channeliser_against_a_numpy_expectation_computed_here(S, plan):
    """Not through oracle/: the direct definition restated with numpy/scipy in f64 — mix by the 32-bit-phase NCO, FIR, decimate,
    discriminate — for a handful of channels of each plan (bank kernel for the two uniform plans, direct form for the third)."""
    from scipy.signal import lfilter
    from sigutil import rel_err
    fs, C, T, D = 20e6, 64, 255, 100
    taps = channel_taps(T, D)
    if plan == "cfg4":
        fw = freq_words((np.arange(C) - 31.5) * 200e3, fs)
    elif plan == "cfg5-interleaved":
        fw = freq_words(((5 + 8 * np.arange(C)) - 255.5) * (fs / 512), fs)
    else:
        fw = freq_words(np.sort(np.random.default_rng(9).uniform(-9e6, 9e6, C)), fs)
    n = D * 500 + 11
    iq = fm_test_signal(n, fs=fs, f_c=1.3e6, f_dev=60e3)
    ch = S.Channeliser(taps, D, fw)
    assert ch.kernel_kind()[0] == (1 if plan == "non-uniform" else 2)
    y, d = ch.process(iq)
    x = (iq[0::2].astype(np.float64) - 127.0) + 1j * (iq[1::2].astype(np.float64) - 127.0)
    nn = np.arange(n, dtype=np.uint64)
    worst = 0.0
    for c in (0, 1, 17, 40, 63):
        theta = 2.0 * np.pi * ((np.uint64(fw[c]) * nn) % np.uint64(1 << 32)).astype(np.float64) / 2.0 ** 32
        yc = lfilter(taps.astype(np.float64), [1.0], x * np.exp(-1j * theta))[D - 1::D]
        want = np.stack([yc.real, yc.imag], axis=1)
        assert_close(y[c], want, what=f"{plan} ch{c} vs numpy")
        worst = max(worst, rel_err(y[c], want))
        z = yc * np.conj(np.concatenate([[0.0], yc[:-1]]))
        dc = GAIN * np.arctan2(z.imag, z.real)
        dc[z == 0] = 0.0
        assert_demod_propagated(d[c], want, dc, GAIN, what=f"{plan} demod ch{c} vs numpy")
    assert worst <= 1e-5, worst
