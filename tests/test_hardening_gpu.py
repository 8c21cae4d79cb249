"""GPU regression tests for handle-lifetime hazards (none of them arithmetic):

  * a persistent ring keeps one kernel resident; other handles on the same GPU must keep working (allocate, launch,
    free) and nothing may wait for the whole device — run in a child process under a timeout so that a regression
    shows up as a failed test, not as a hung box;
  * the per-kernel dynamic shared-memory limit is process-global: a later handle with a smaller tile must not lower it;
  * the shared-memory-tap channeliser must not store the padding rows of its last 64-channel group;
  * commit followed at once by close must still process the committed buffer.
"""
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg
from sigutil import assert_close, channel_taps

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
ROOT = Path(__file__).resolve().parent.parent
BUF = O.DEFAULT_BUF_LENGTH


@pytest.fixture(scope="module")
def S():
    m = sdrpkg.load()
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return m


def _child(code: str, timeout: int = 150):
    """Run `code` in a fresh interpreter; a hang is a failure of the test, not of the session."""
    prelude = f"import sys; sys.path[:0] = [{str(ROOT)!r}, {str(ROOT / 'tests')!r}]\n"
    try:
        r = subprocess.run([sys.executable, "-c", prelude + textwrap.dedent(code)], capture_output=True, text=True,
                           timeout=timeout)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"child did not finish within {timeout} s (device-wide wait behind a resident ring kernel?)\n"
                    f"stdout: {(e.stdout or b'')[-2000:]}\nstderr: {(e.stderr or b'')[-2000:]}")
    assert r.returncode == 0, f"child failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r.stdout


def test_other_handles_work_while_a_ring_is_resident():
    out = _child("""
        import numpy as np
        import oracle_ffi as O, sdrpkg
        from sigutil import channel_taps
        S = sdrpkg.load()
        BUF = O.DEFAULT_BUF_LENGTH
        rng = np.random.default_rng(7)
        data = rng.integers(0, 256, 6 * BUF, dtype=np.uint8)
        d1, o1 = S.Demod(), O.Demod()
        def step(msg):
            print(msg, file=sys.stderr, flush=True)      # a timeout report shows how far the child got
        ring = S.Ring(d1, BUF, n_slots=4)
        ring.submit(data[:BUF])
        step("ring open, one buffer committed")
        # a second integer Demod: new (allocates), several calls (the first grows its buffers), free
        d2, o2 = S.Demod(S.DemodConfig(160000, 160000, 32000, 5, 1)), O.Demod(O.DemodConfig(160000, 160000, 32000, 5, 1))
        for c in range(3):
            assert np.array_equal(d2.demodulate(data[c * BUF:(c + 1) * BUF]), o2.demodulate(data[c * BUF:(c + 1) * BUF]))
        step("second Demod: small calls done")
        big = np.tile(data, 8)                       # 12 MiB: the chunked path, fresh device buffers
        assert np.array_equal(d2.demodulate(big), o2.demodulate(big))
        d2.close()
        step("second Demod: chunked call and free done")
        # an f32 receiver and a device buffer come and go as well
        taps = channel_taps(127, 75)
        rx = S.FmRx(taps, 75)
        y, dm, _ = rx.process(data[:75 * 2 * 2000])
        yo, do, _ = O.FxChain(taps, 75).process(data[:75 * 2 * 2000])
        assert np.allclose(y, yo, rtol=1e-4, atol=1e-2)
        rx.close()
        step("FmRx done")
        b = S.DevBuffer(1 << 20); b.free()
        h = S.HostBuffer(1 << 20); h.free()
        step("buffers done")
        # a device-wide wait is refused, not entered
        try:
            S.lib().sdr_device_sync(0)
            rc = S.lib().sdr_device_sync(0)
        except Exception as e:
            rc = None
        assert rc == -7, rc
        # the ring is still alive and exact
        got = [ring.collect()]
        for c in range(1, 4):
            ring.submit(data[c * BUF:(c + 1) * BUF])
            got.append(ring.collect())
        # commit immediately followed by close: the committed buffer must still be processed and its state handed back
        ring.submit(data[4 * BUF:5 * BUF])
        ring.close()
        step("ring closed")
        want = [o1.demodulate(data[c * BUF:(c + 1) * BUF]) for c in range(5)]
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        assert d1.state() == o1.state(), (d1.state(), o1.state())
        assert np.array_equal(d1.demodulate(data[5 * BUF:]), o1.demodulate(data[5 * BUF:]))
        assert S.lib().sdr_device_sync(0) == 0       # allowed again, and the parked frees have been carried out
        print("OK")
    """)
    assert "OK" in out


def test_freeing_a_demod_closes_its_ring_and_two_rings_coexist():
    out = _child("""
        import numpy as np
        import oracle_ffi as O, sdrpkg
        S = sdrpkg.load()
        BUF = O.DEFAULT_BUF_LENGTH
        data = np.random.default_rng(3).integers(0, 256, 4 * BUF, dtype=np.uint8)
        da, db = S.Demod(), S.Demod()
        oa, ob = O.Demod(), O.Demod()
        ra, rb = S.Ring(da, BUF, n_slots=2), S.Ring(db, BUF, n_slots=3)      # two resident kernels on one GPU
        for c in range(4):
            ra.submit(data[c * BUF:(c + 1) * BUF]); rb.submit(data[(3 - c) * BUF:(4 - c) * BUF])
            assert np.array_equal(ra.collect(), oa.demodulate(data[c * BUF:(c + 1) * BUF]))
            assert np.array_equal(rb.collect(), ob.demodulate(data[(3 - c) * BUF:(4 - c) * BUF]))
        rb.close()
        assert db.state() == ob.state()
        # free the Demod while its ring is still open (what Python's GC order can do): must retire the kernel, not hang
        ra._h = None                                 # the wrapper forgets the ring; the C handle is still open
        da.close()
        d = S.Demod()
        assert np.array_equal(d.demodulate(data[:BUF]), O.Demod().demodulate(data[:BUF]))
        assert S.lib().sdr_device_sync(0) == 0
        print("OK")
    """)
    assert "OK" in out


def test_dynamic_shared_memory_limit_is_only_ever_raised(S):
    """A handle with a large tile, then one with a small tile of the SAME kernel: the first must still launch."""
    rng = np.random.default_rng(11)
    data = rng.integers(0, 256, 8 * 4096, dtype=np.uint8)
    # k_demod_fused<0>: downsample 1 (large staged tile) then downsample 20 (small)
    big, obig = S.Demod(S.DemodConfig(170000, 170000, 32000, 1, 1)), O.Demod(O.DemodConfig(170000, 170000, 32000, 1, 1))
    assert np.array_equal(big.demodulate(data), obig.demodulate(data))
    small, osmall = S.Demod(S.DemodConfig(170000, 170000, 32000, 20, 1)), O.Demod(O.DemodConfig(170000, 170000, 32000, 20, 1))
    assert np.array_equal(small.demodulate(data), osmall.demodulate(data))
    assert np.array_equal(big.demodulate(data), obig.demodulate(data))
    # k_fir_generic (SDR_FORCE_GENERIC is read at handle creation): long taps, then short taps
    import os
    os.environ["SDR_FORCE_GENERIC"] = "1"
    try:
        iq = rng.integers(0, 256, 2 * 40000, dtype=np.uint8)
        t_long, t_short = channel_taps(2001, 40), channel_taps(9, 4)
        a = S.FmRx(t_long, 40)
        ya = a.process(iq)[0]
        b = S.FmRx(t_short, 4)
        b.process(iq)
        a.reset()
        assert np.array_equal(a.process(iq)[0], ya)
        assert_close(ya, O.FxChain(t_long, 40).process(iq)[0], what="generic FIR after a smaller handle")
    finally:
        del os.environ["SDR_FORCE_GENERIC"]


def test_smem_tap_channeliser_does_not_store_padding_rows(S, monkeypatch):
    """n_channels % 64 != 0 through the shared-memory-tap kernel into a caller buffer of exactly [C][cap]."""
    monkeypatch.setenv("SDR_CHAN_SMEM_TAPS", "1")
    C_, T, D, n = 5, 33, 8, 8 * 3000
    taps = channel_taps(T, D)
    fw = (np.arange(C_, dtype=np.uint64) * 0x12345679 % (1 << 32)).astype(np.uint32)
    iq = np.random.default_rng(5).integers(0, 256, 2 * n, dtype=np.uint8)
    cap = n // D
    ch = S.Channeliser(taps, D, fw)
    d_in = S.DevBuffer(2 * n).upload(iq)
    guard = 64 * cap * 8                                        # room for the 59 rows a regression would write
    d_y = S.DevBuffer(C_ * cap * 8 + guard)
    d_d = S.DevBuffer(C_ * cap * 4)
    canary = np.full(guard // 4, 0x7FC0DEAD, np.uint32)
    d_y.upload(canary, offset=C_ * cap * 8)
    m = ch.process_dev(d_in, n, d_d, cap, d_y=d_y)
    ch.sync()
    assert m == cap
    assert np.array_equal(d_y.download(np.uint32, guard // 4, offset=C_ * cap * 8), canary), "rows >= n_channels were written"
    y = d_y.download(np.float32, C_ * cap * 2).reshape(C_, cap, 2)
    yo, _ = O.channelise(iq, taps, D, fw)
    assert_close(y, yo, what="smem-tap channeliser, 5 channels")


def test_pageable_pinned_and_registered_inputs_give_the_same_audio(S):
    """The same 24 MiB stream handed over as an ordinary (pageable) array — staged through pinned pieces by the copy threads —,
    from sdr_host_alloc memory, and from a caller-owned array page-locked with sdr_host_register: identical audio from the
    integer Demod and the f32 receiver, and the first buffers agree with the oracle."""
    rng = np.random.default_rng(77)
    n_bufs = 96
    data = rng.integers(0, 256, n_bufs * BUF, dtype=np.uint8)
    pinned = S.HostBuffer(data.size)
    pinned.array[:] = data
    owned = data.copy()

    def run_int(buf):
        h = S.Demod()
        try:
            return h.demodulate_batch(buf, BUF)
        finally:
            h.close()

    taps = channel_taps(63, 20)
    def run_f32(buf):
        h = S.FmRx(taps, 20, None, 1, 1)
        try:
            return h.process(buf, want_y=False, want_demod=False)[2]
        finally:
            h.close()

    a_page, f_page = run_int(data), run_f32(data)
    a_pin, f_pin = run_int(pinned.array), run_f32(pinned.array)
    S.host_register(owned)
    S.host_register(owned)            # twice is fine
    try:
        a_reg, f_reg = run_int(owned), run_f32(owned)
    finally:
        S.host_unregister(owned)
        S.host_unregister(owned)      # unknown pointer is fine
    assert np.array_equal(a_page, a_pin) and np.array_equal(a_page, a_reg)
    assert np.array_equal(f_page, f_pin) and np.array_equal(f_page, f_reg)
    o = O.Demod()
    want = np.concatenate([o.demodulate(data[i * BUF:(i + 1) * BUF]) for i in range(3)])
    assert np.array_equal(a_page[:want.size], want)
    pinned.free()
