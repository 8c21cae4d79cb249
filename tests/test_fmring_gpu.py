"""GPU tests of the f32 receiver's persistent ring (sdr_fmrx_ring_*): successive USB-sized buffers stream through ONE
resident kernel; the audio is BIT-IDENTICAL to one sdr_fmrx_process call per buffer (same FIR tile code, same ascending-tap
fma chain in the audio stage), and close hands the stream back to the handle."""
import threading
import time

import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg
from sigutil import channel_taps, fm_test_signal, lowpass_taps

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
BUF = O.DEFAULT_BUF_LENGTH


@pytest.fixture(scope="module")
def S():
    m = sdrpkg.load()
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return m


def make(S, T, D, T2, up, down):
    taps = channel_taps(T, D)
    taps2 = lowpass_taps(T2, 0.45 / max(up, down), gain=up) if T2 else None
    return (lambda: S.FmRx(taps, D, taps2, up, down))


SHAPES = [  # (T, D, T2, up, down, buf_len, kernel kind)
    (127, 75, 63, 1, 1, BUF, 1),          # cfg2
    (255, 100, 127, 4, 25, BUF, 1),       # cfg3
    (63, 20, 95, 16, 25, 65536, 2),       # run-time-compiled shape, rational resampler
    (127, 75, 0, 1, 1, BUF, 1),           # no resample stage: the discriminator output is the audio
    (31, 7, 31, 1, 5, 16 * 1000 + 16, 2), # odd decimation, buffer length not a multiple of anything but 16
]


@pytest.mark.parametrize("T,D,T2,up,down,buf_len,kind", SHAPES)
def test_ring_is_bit_identical_to_one_call_per_buffer(S, T, D, T2, up, down, buf_len, kind):
    new = make(S, T, D, T2, up, down)
    n_bufs = 13
    iq = fm_test_signal(n_bufs * buf_len // 2 + 5000, fs=2.4e6, seed=T)
    ref, rx = new(), new()
    assert rx.kernel_kind()[0] == kind
    # open the ring on a stream that is already under way (odd sample count: non-trivial phase and history)
    pre = iq[:2 * 2501]
    assert np.array_equal(ref.process(pre)[2], rx.process(pre)[2])
    body = iq[2 * 2501:2 * 2501 + n_bufs * buf_len]
    want = [ref.process(body[i * buf_len:(i + 1) * buf_len], want_y=False, want_demod=False)[2] for i in range(n_bufs)]
    launches0 = S.kernel_launch_count()
    ring = S.FmRing(rx, buf_len, n_slots=4)
    got = []
    for i in range(n_bufs):
        ring.submit(body[i * buf_len:(i + 1) * buf_len])
        if i >= 2:
            got.append(ring.collect())            # keep three buffers in flight
    got += [ring.collect() for _ in range(2)]
    with pytest.raises(S.SdrError) as e:
        rx.process(pre)                           # the ring owns the handle
    assert e.value.code == -7
    ring.close()
    assert S.kernel_launch_count() - launches0 == 1          # ONE kernel for the whole stream
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, (i, g.shape, w.shape)
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), i
    # the handle continues the same stream with ordinary calls
    tail = iq[2 * 2501 + n_bufs * buf_len:]
    for a, b in zip(ref.process(tail), rx.process(tail)):
        assert np.array_equal(a, b)


def test_ring_producer_consumer_threads_and_latency(S):
    new = make(S, 127, 75, 63, 1, 1)
    n_bufs = 200
    data = np.random.default_rng(5).integers(0, 256, 8 * BUF, dtype=np.uint8)
    ref = new()
    want = [ref.process(data[(i % 8) * BUF:(i % 8 + 1) * BUF], want_y=False, want_demod=False)[2] for i in range(n_bufs)]
    rx = new()
    ring = S.FmRing(rx, BUF, n_slots=8)
    got = []

    def producer():
        for i in range(n_bufs):
            ring.submit(data[(i % 8) * BUF:(i % 8 + 1) * BUF])    # blocks while the ring is full
    t = threading.Thread(target=producer)
    t0 = time.perf_counter()
    t.start()
    while len(got) < n_bufs:
        try:
            got.append(ring.collect())
        except S.SdrError as e:
            if e.code != -7:
                raise
            time.sleep(0.0002)
    dt = time.perf_counter() - t0
    t.join()
    ring.close()
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), i
    print(f"\nf32 ring: {dt / n_bufs * 1e6:.1f} us per 262144-byte buffer ({n_bufs * BUF / 2 / dt / 1e6:.0f} Msamples/s) driven from Python")


def test_ring_validates(S):
    rx = make(S, 127, 75, 63, 1, 1)()
    for bad in (12, 24, 16 * 3):                  # not a multiple of 16 / shorter than the filter history
        with pytest.raises(S.SdrError) as e:
            S.FmRing(rx, bad)
        assert e.value.code == -2
    with pytest.raises(S.SdrError):
        S.FmRing(rx, BUF, n_slots=1)
    import os
    os.environ["SDR_FORCE_GENERIC"] = "1"
    try:
        gen = make(S, 127, 75, 63, 1, 1)()
    finally:
        del os.environ["SDR_FORCE_GENERIC"]
    with pytest.raises(S.SdrError) as e:
        S.FmRing(gen, BUF)                        # the generic kernel has no ring
    assert e.value.code == -7
    ring = S.FmRing(rx, BUF, n_slots=2)
    with pytest.raises(S.SdrError):
        ring.collect()                            # nothing outstanding
    rx.close()                                    # freeing the handle retires its ring
    ring._h = None
    assert make(S, 127, 75, 63, 1, 1)().process(np.zeros(2 * 7500, np.uint8))[2].size == 100
