"""GPU test of the persistent-kernel ring (sdr_demod_ring_*): successive USB-sized buffers stream through ONE
resident kernel and the audio is bit-identical to one Demod::demodulate call per buffer
(examples/simple_fm.rs:108-128,145-160)."""
import hashlib
import json
import threading

import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]
BUF = O.DEFAULT_BUF_LENGTH


@pytest.fixture(scope="module")
def S():
    m = sdrpkg.load()
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return m


def test_ring_streams_capture_head_bit_exact(S, golden_dir):
    head = np.fromfile(golden_dir / "capture_head.bin", np.uint8)
    want = np.fromfile(golden_dir / "capture_head_audio.s16le", "<i2")
    d = S.Demod()
    launches0 = S.kernel_launch_count()
    ring = S.Ring(d, BUF, n_slots=3)
    got = []
    for c in range(4):
        ring.submit(head[c * BUF:(c + 1) * BUF])
        if c >= 1:
            got.append(ring.collect())          # keep two buffers in flight
    got.append(ring.collect())
    ring.close()
    assert S.kernel_launch_count() - launches0 == 1          # ONE kernel for the whole stream
    assert np.array_equal(np.concatenate(got), want)
    o = O.Demod()
    for c in range(4):
        o.demodulate(head[c * BUF:(c + 1) * BUF])
    assert d.state() == o.state()                            # the ring hands the carried state back
    # and the handle continues the same stream with ordinary calls
    extra = np.random.default_rng(1).integers(0, 256, 8 * 999, dtype=np.uint8)
    assert np.array_equal(d.demodulate(extra), o.demodulate(extra))


@pytest.mark.parametrize("D,fast,slow,buf_len,n_bufs,slots", [(6, 170_000, 32_000, 262144, 40, 8), (15, 160_000, 32_000, 4096, 300, 5),
                                                              (7, 100_003, 31_999, 40 * 8, 500, 2), (6, 170_000, 32_000, 8 * 3, 64, 64)])
def test_ring_producer_consumer_threads(S, D, fast, slow, buf_len, n_bufs, slots):
    """Reader thread -> ring -> processor thread, like the reference's two threads; ragged configs and a 1-tile buffer."""
    rng = np.random.default_rng(D)
    data = rng.integers(0, 256, n_bufs * buf_len, dtype=np.uint8)
    d = S.Demod(S.DemodConfig(fast, fast, slow, D, 1))
    o = O.Demod(O.DemodConfig(fast, fast, slow, D, 1))
    d.demodulate(data[:buf_len]), o.demodulate(data[:buf_len])      # open the ring on a non-trivial state
    want = [o.demodulate(data[i * buf_len:(i + 1) * buf_len]) for i in range(n_bufs)]
    ring = S.Ring(d, buf_len, n_slots=slots)
    got = []

    def producer():
        for i in range(n_bufs):
            ring.submit(data[i * buf_len:(i + 1) * buf_len])    # blocks while the ring is full
    t = threading.Thread(target=producer)
    t.start()
    for _ in range(n_bufs):
        got.append(_collect(ring, S))
    t.join()
    ring.close()
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), i
    assert d.state() == o.state()


def _collect(ring, S):
    import time
    while True:
        try:
            return ring.collect()
        except S.SdrError as e:
            if e.code != -7:        # SDR_E_STATE: nothing outstanding yet
                raise
            time.sleep(0.0005)


def test_ring_full_capture_golden(S, golden_dir):
    full = golden_dir / "_ref" / "capture.bin"
    if not full.exists():
        pytest.skip("full capture.bin copy not present")
    pins = json.loads((golden_dir / "capture_pins.json").read_text())
    cap = np.fromfile(full, np.uint8)
    d = S.Demod()
    ring = S.Ring(d, BUF, n_slots=8)
    got = []
    for c in range(pins["n_calls"]):
        ring.submit(cap[c * BUF:(c + 1) * BUF])
        if c >= 7:
            got.append(ring.collect())
    for _ in range(7):
        got.append(ring.collect())
    ring.close()
    au = np.concatenate(got)
    assert hashlib.sha256(au.astype("<i2").tobytes()).hexdigest() == pins["audio_sha256"]


def test_ring_owns_the_handle_and_validates(S):
    d = S.Demod()
    with pytest.raises(S.SdrError) as e:
        S.Ring(d, 12)                      # len % 8 != 0
    assert e.value.code == -2
    ring = S.Ring(d, 4096, n_slots=2)
    with pytest.raises(S.SdrError) as e:
        d.demodulate(np.zeros(4096, np.uint8))
    assert e.value.code == -7
    with pytest.raises(S.SdrError):
        ring.collect()                     # nothing outstanding
    ring.close()
    assert d.demodulate(np.zeros(4096, np.uint8)).size > 0
