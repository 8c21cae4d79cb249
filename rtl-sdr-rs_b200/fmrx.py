"""`FmRx` — host-side mirror of the f32 tap'd-FIR receiver (sdr_fmrx_*, include/sdr_b200.h §2).

Stage names follow BASELINE.json's north_star: low_pass / fm_demod / resample.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class FmRx:
    def __init__(self, taps, decim: int, taps2=None, up: int = 1, down: int = 1, gain: float = 0.0, device: int = 0):
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.taps2 = np.ascontiguousarray(taps2 if taps2 is not None else [], np.float32)
        self.decim, self.up, self.down, self.device = int(decim), int(up), int(down), device
        cfg = F.FmrxConfig(self.taps.size, decim, self.taps2.size, up, down, gain)
        h = C.c_void_p()
        F.check(F.lib().sdr_fmrx_new(C.byref(cfg), F.ptr(self.taps), F.ptr(self.taps2) if self.taps2.size else None,
                                     device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            F.lib().sdr_fmrx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        F.check(F.lib().sdr_fmrx_reset(self._h))

    def out_lens(self, n_samples: int):
        ny, na = C.c_size_t(0), C.c_size_t(0)
        F.check(F.lib().sdr_fmrx_out_lens(self._h, n_samples, C.byref(ny), C.byref(na)))
        return ny.value, na.value

    def process(self, iq_u8: np.ndarray, want_y: bool = True, want_demod: bool = True):
        """u8 IQ -> (y complex f32 pairs, discriminator f32, audio f32); streaming state carried."""
        b = np.ascontiguousarray(iq_u8, np.uint8)
        n = b.size // 2
        ny, na = self.out_lens(n)
        y = np.empty((ny, 2), np.float32) if want_y else None
        d = np.empty(ny, np.float32) if want_demod else None
        a = np.empty(na, np.float32)
        r = F.check(F.lib().sdr_fmrx_process(self._h, F.ptr(b), n,
                                             F.ptr(y) if want_y else None, ny,
                                             F.ptr(d) if want_demod else None, ny, F.ptr(a), na))
        assert r == na
        return y, d, a

    def process_dev(self, d_iq: F.DevBuffer, n_samples: int, d_audio: F.DevBuffer, audio_cap: int,
                    d_y: F.DevBuffer | None = None, d_demod: F.DevBuffer | None = None, iq_offset: int = 0) -> int:
        return F.check(F.lib().sdr_fmrx_process_dev(self._h, d_iq.at(iq_offset), n_samples,
                                                    d_y.ptr if d_y else None, d_demod.ptr if d_demod else None,
                                                    d_audio.ptr, audio_cap))

    def low_pass(self, iq_u8: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(iq_u8, np.uint8)
        n = b.size // 2
        out = np.empty((n // self.decim + 2, 2), np.float32)
        r = F.check(F.lib().sdr_fmrx_low_pass(self._h, F.ptr(b), n, F.ptr(out), out.shape[0]))
        return out[:r].copy()

    def fm_demod(self, y_pairs: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y_pairs, np.float32).reshape(-1, 2)
        out = np.empty(max(y.shape[0], 1), np.float32)
        r = F.check(F.lib().sdr_fmrx_fm_demod(self._h, F.ptr(y), y.shape[0], F.ptr(out), out.size))
        return out[:r].copy()

    def resample(self, d: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(d, np.float32)
        out = np.empty(x.size * self.up // self.down + 2, np.float32)
        r = F.check(F.lib().sdr_fmrx_resample(self._h, F.ptr(x), x.size, F.ptr(out), out.size))
        return out[:r].copy()

    def sync(self):
        F.check(F.lib().sdr_fmrx_sync(self._h))

    def last_timing(self):
        ms = (C.c_float * 3)()
        n, spec = C.c_uint32(0), C.c_int(0)
        F.check(F.lib().sdr_fmrx_last_timing(self._h, C.byref(ms), C.byref(n), C.byref(spec)))
        return list(ms), n.value, spec.value

    def span_begin(self):
        F.check(F.lib().sdr_fmrx_span_begin(self._h))

    def span_end(self) -> float:
        ms = C.c_float(0)
        F.check(F.lib().sdr_fmrx_span_end(self._h, C.byref(ms)))
        return ms.value

    def seek(self, global_sample_index: int):
        F.check(F.lib().sdr_fmrx_seek(self._h, global_sample_index))

    def kernel_kind(self):
        """(kind, note): 0 generic, 1 pre-compiled k_fir_fast, 2 run-time-compiled k_fir_fast (sdr_fmrx_kernel_kind)."""
        note = C.c_char_p()
        kind = F.check(F.lib().sdr_fmrx_kernel_kind(self._h, C.byref(note)))
        return kind, (note.value or b"").decode()

    def timing_totals(self, reset: bool = False):
        sums = (C.c_double * 3)()
        n = C.c_uint64(0)
        F.check(F.lib().sdr_fmrx_timing_totals(self._h, C.byref(sums), C.byref(n), int(reset)))
        return list(sums), n.value


class FmRing:
    """Persistent-kernel ring over an FmRx (sdr_fmrx_ring_*): buffers stream through one resident kernel."""

    def __init__(self, rx: FmRx, buf_len: int, n_slots: int = 8):
        self.rx, self.buf_len = rx, buf_len
        h = C.c_void_p()
        F.check(F.lib().sdr_fmrx_ring_open(rx._h, buf_len, n_slots, C.byref(h)))
        self._h = h
        self._cap = buf_len // 2 // max(rx.decim, 1) * max(rx.up, 1) // max(rx.down, 1) + 16

    def submit(self, buf: np.ndarray):
        """acquire the next pinned slot, copy `buf` into it, commit (H2D + doorbell, no kernel launch)."""
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        assert b.size == self.buf_len
        p = C.c_void_p()
        F.check(F.lib().sdr_fmrx_ring_acquire(self._h, C.byref(p)))
        C.memmove(p, b.ctypes.data, b.size)
        F.check(F.lib().sdr_fmrx_ring_commit(self._h))

    def collect(self) -> np.ndarray:
        out = np.empty(self._cap, np.float32)
        n = F.check(F.lib().sdr_fmrx_ring_collect(self._h, F.ptr(out), out.size))
        return out[:n].copy()

    def close(self):
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            F.check(F.lib().sdr_fmrx_ring_close(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
