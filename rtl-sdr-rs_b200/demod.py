"""`Demod` — host-side mirror of the reference's struct Demod (examples/simple_fm.rs:232-427).

Every method forwards to the CUDA library through the C ABI; names, argument meaning and error
behaviour follow the reference (its panics become SdrError with SDR_E_LEN).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class Demod:
    def __init__(self, config: F.DemodConfig | None = None, device: int = 0):
        """Demod::new(config), examples/simple_fm.rs:243-252."""
        if config is None:
            _, config = F.optimal_settings()
        self.config = config
        self.device = device
        h = C.c_void_p()
        F.check(F.lib().sdr_demod_new(C.byref(config), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            F.lib().sdr_demod_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- whole pipeline ------------------------------------------------------------------
    def out_len(self, n_bytes: int) -> int:
        return F.check(F.lib().sdr_demod_out_len(self._h, n_bytes))

    def demodulate(self, buf: np.ndarray) -> np.ndarray:
        """demodulate(&mut self, Vec<u8>) -> Vec<i16>, :256-269 (one fused kernel launch)."""
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        out = np.empty(b.size // 2 + 8, np.int16)
        n = F.check(F.lib().sdr_demod_demodulate(self._h, F.ptr(b), b.size, F.ptr(out), out.size))
        return out[:n].copy()

    def demodulate_batch(self, buf: np.ndarray, buf_len: int, with_lens: bool = False):
        """n consecutive demodulate() calls of buf_len bytes in one pipelined submission."""
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        assert buf_len > 0 and b.size % buf_len == 0
        n_bufs = b.size // buf_len
        out = np.empty(b.size // 2 + 8, np.int16)
        lens = np.zeros(n_bufs, np.uint32)
        n = F.check(F.lib().sdr_demod_demodulate_batch(self._h, F.ptr(b), buf_len, n_bufs, F.ptr(out), out.size,
                                                       lens.ctypes.data_as(C.POINTER(C.c_uint32))))
        return (out[:n].copy(), lens) if with_lens else out[:n].copy()

    def demodulate_batch_dev(self, d_in: F.DevBuffer, buf_len: int, n_bufs: int, d_out: F.DevBuffer, out_cap: int) -> int:
        return F.check(F.lib().sdr_demod_demodulate_batch_dev(self._h, d_in.ptr, buf_len, n_bufs, d_out.ptr, out_cap))

    def sync(self):
        F.check(F.lib().sdr_demod_sync(self._h))

    def last_timing(self):
        ms, n = C.c_float(0), C.c_uint32(0)
        F.check(F.lib().sdr_demod_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- stages (names as in the reference) ------------------------------------------------
    def rotate_90(self, buf: np.ndarray) -> np.ndarray:
        """Demod::rotate_90, scalar branch :276-299."""
        b = np.ascontiguousarray(buf, dtype=np.uint8).copy()
        F.check(F.lib().sdr_rotate_90(self._h, F.ptr(b), b.size))
        return b

    def buf_to_complex(self, buf: np.ndarray) -> np.ndarray:
        """`v as i16 - 127` (:258) then buf_to_complex (:441-450): u8 -> Complex<i32> pairs."""
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        out = np.empty((b.size // 2, 2), np.int32)
        n = F.check(F.lib().sdr_buf_to_complex(self._h, F.ptr(b), b.size, F.ptr(out), out.shape[0]))
        return out[:n]

    def low_pass_complex(self, pairs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty((x.shape[0] // max(self.config.downsample, 1) + 2, 2), np.int32)
        n = F.check(F.lib().sdr_low_pass_complex(self._h, F.ptr(x), x.shape[0], F.ptr(out), out.shape[0]))
        return out[:n].copy()

    def fm_demod(self, pairs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty(max(x.shape[0], 1), np.int16)
        n = F.check(F.lib().sdr_fm_demod(self._h, F.ptr(x), x.shape[0], F.ptr(out), out.size))
        return out[:n].copy()

    def low_pass_real(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty(x.size + 2, np.int16)
        n = F.check(F.lib().sdr_low_pass_real(self._h, F.ptr(x), x.size, F.ptr(out), out.size))
        return out[:n].copy()

    def fast_atan2(self, y, x) -> np.ndarray:
        """Demod::fast_atan2(y, x), :383-405, element-wise."""
        y = np.ascontiguousarray(np.atleast_1d(y), dtype=np.int32)
        x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.int32)
        assert y.shape == x.shape
        out = np.empty(y.size, np.int32)
        F.check(F.lib().sdr_fast_atan2(self._h, F.ptr(y), F.ptr(x), y.size, F.ptr(out)))
        return out

    def _polar(self, a, b, fast: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 2)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, 2)
        assert a.shape == b.shape
        out = np.empty(a.shape[0], np.int32)
        F.check(F.lib().sdr_polar_discriminant(self._h, F.ptr(a), F.ptr(b), a.shape[0], fast, F.ptr(out)))
        return out

    def polar_discriminant(self, a, b) -> np.ndarray:
        """:370-374 (f64 atan2)."""
        return self._polar(a, b, 0)

    def polar_discriminant_fast(self, a, b) -> np.ndarray:
        """:377-380 (integer fast_atan2)."""
        return self._polar(a, b, 1)

    # ---- carried state -----------------------------------------------------------------------
    def state(self) -> dict:
        s = F.DemodState()
        F.check(F.lib().sdr_demod_get_state(self._h, C.byref(s)))
        return dict(prev_index=int(s.prev_index), now_lpr=s.now_lpr, prev_lpr_index=s.prev_lpr_index,
                    lp_now=(s.lp_now_re, s.lp_now_im), demod_pre=(s.demod_pre_re, s.demod_pre_im))

    def set_state(self, prev_index=0, now_lpr=0, prev_lpr_index=0, lp_now=(0, 0), demod_pre=(0, 0)):
        s = F.DemodState(prev_index, now_lpr, prev_lpr_index, lp_now[0], lp_now[1], demod_pre[0], demod_pre[1])
        F.check(F.lib().sdr_demod_set_state(self._h, C.byref(s)))

    def span_begin(self):
        F.check(F.lib().sdr_demod_span_begin(self._h))

    def span_end(self) -> float:
        ms = C.c_float(0)
        F.check(F.lib().sdr_demod_span_end(self._h, C.byref(ms)))
        return ms.value


class Ring:
    """Persistent-kernel ring over a Demod (sdr_demod_ring_*): buffers stream through one resident kernel."""

    def __init__(self, demod: Demod, buf_len: int, n_slots: int = 8):
        self.demod, self.buf_len = demod, buf_len
        h = C.c_void_p()
        F.check(F.lib().sdr_demod_ring_open(demod._h, buf_len, n_slots, C.byref(h)))
        self._h = h
        self._cap = buf_len // 2 + 8

    def submit(self, buf: np.ndarray):
        """acquire the next pinned slot, copy `buf` into it, commit (H2D + doorbell, no kernel launch)."""
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        assert b.size == self.buf_len
        p = C.c_void_p()
        F.check(F.lib().sdr_ring_acquire(self._h, C.byref(p)))
        C.memmove(p, b.ctypes.data, b.size)
        F.check(F.lib().sdr_ring_commit(self._h))

    def collect(self) -> np.ndarray:
        out = np.empty(self._cap, np.int16)
        n = F.check(F.lib().sdr_ring_collect(self._h, F.ptr(out), out.size))
        return out[:n].copy()

    def close(self):
        if getattr(self, "_h", None):
            F.check(F.lib().sdr_ring_close(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AudioPost:
    """Optional post-stages after low_pass_real (sdr_post_*): output_scale, squelch, de-emphasis, DC block.  All off by default."""

    def __init__(self, output_scale: int = 0, squelch_level: int = 0, deemph_a: int = 0, dc_block: bool = False, device: int = 0):
        cfg = F.PostConfig(output_scale, squelch_level, deemph_a, int(dc_block))
        h = C.c_void_p()
        F.check(F.lib().sdr_post_new(C.byref(cfg), device, C.byref(h)))
        self._h = h

    @staticmethod
    def deemph_a(rate: int, tau_us: float = 75.0) -> int:
        return int(F.lib().sdr_post_deemph_a(rate, tau_us))

    def process(self, audio: np.ndarray, raw: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(audio, np.int16).copy()
        r = np.ascontiguousarray(raw, np.uint8) if raw is not None else None
        F.check(F.lib().sdr_post_process(self._h, F.ptr(a), a.size, F.ptr(r) if r is not None else None, r.size if r is not None else 0))
        return a

    def close(self):
        if getattr(self, "_h", None):
            F.lib().sdr_post_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
