"""`Channeliser` / `Comm` — host-side mirror of sdr_chan_* and sdr_comm_* (include/sdr_b200.h §3)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class Channeliser:
    def __init__(self, taps, decim: int, freq_words, gain: float = 0.0, device: int = 0):
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.freq_words = np.ascontiguousarray(freq_words, np.uint32)
        self.decim, self.device = int(decim), device
        cfg = F.ChanConfig(self.freq_words.size, self.taps.size, decim, gain)
        h = C.c_void_p()
        F.check(F.lib().sdr_chan_new(C.byref(cfg), F.ptr(self.taps), F.ptr(self.freq_words), device, C.byref(h)))
        self._h = h

    @property
    def n_channels(self) -> int:
        return int(self.freq_words.size)

    def close(self):
        if getattr(self, "_h", None):
            F.lib().sdr_chan_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        F.check(F.lib().sdr_chan_reset(self._h))

    def process(self, iq_u8: np.ndarray, want_y: bool = True):
        b = np.ascontiguousarray(iq_u8, np.uint8)
        n = b.size // 2
        cap = n // self.decim + 2
        y = np.empty((self.n_channels, cap, 2), np.float32) if want_y else None
        d = np.empty((self.n_channels, cap), np.float32)
        m = F.check(F.lib().sdr_chan_process(self._h, F.ptr(b), n, F.ptr(y) if want_y else None, F.ptr(d), cap))
        return (y[:, :m].copy() if want_y else None), d[:, :m].copy()

    def process_dev(self, d_iq: F.DevBuffer, n_samples: int, d_demod: F.DevBuffer, cap_per_channel: int,
                    d_y: F.DevBuffer | None = None) -> int:
        return F.check(F.lib().sdr_chan_process_dev(self._h, d_iq.ptr, n_samples, d_y.ptr if d_y else None,
                                                    d_demod.ptr, cap_per_channel))

    def kernel_kind(self):
        """(kind, (K, K1, K2, groups)): 0 shared-memory taps, 1 direct form with uniform taps, 2 two-stage polyphase bank."""
        info = (C.c_uint32 * 4)()
        kind = F.check(F.lib().sdr_chan_kernel_kind(self._h, C.byref(info)))
        return kind, tuple(info)

    def sync(self):
        F.check(F.lib().sdr_chan_sync(self._h))

    def last_timing(self):
        ms, n = C.c_float(0), C.c_uint32(0)
        F.check(F.lib().sdr_chan_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def bank_plan(taps, decim: int, freq_words, want_tables: bool = True):
    """sdr_chan_bank_plan: (K, K1, K2, groups) and the per-group coefficient tables [groups][3840][2] the bank kernel would
    receive, or None when the channels are not a uniform bank.  Pure host arithmetic."""
    taps = np.ascontiguousarray(taps, np.float32)
    fw = np.ascontiguousarray(freq_words, np.uint32)
    cfg = F.ChanConfig(fw.size, taps.size, decim, 0.0)
    info = (C.c_uint32 * 4)()
    g = F.check(F.lib().sdr_chan_bank_plan(C.byref(cfg), F.ptr(taps), F.ptr(fw), C.byref(info), None, 0))
    if g == 0:
        return None
    tabs = None
    if want_tables:
        tabs = np.zeros((g, 3840, 2), np.float32)
        F.check(F.lib().sdr_chan_bank_plan(C.byref(cfg), F.ptr(taps), F.ptr(fw), C.byref(info), F.ptr(tabs), tabs.size))
    return tuple(info), tabs


class Comm:
    """NCCL communicator used only for the raw-slab broadcast."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * F.NCCL_ID_BYTES)()
        F.check(F.lib().sdr_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, device: int, rank: int, world: int, uid: bytes):
        assert len(uid) == F.NCCL_ID_BYTES
        h = C.c_void_p()
        buf = (C.c_uint8 * F.NCCL_ID_BYTES).from_buffer_copy(uid)
        F.check(F.lib().sdr_comm_init(device, rank, world, buf, C.byref(h)))
        self._h = h

    def bcast_u8(self, d_buf: F.DevBuffer, nbytes: int, root: int = 0, offset: int = 0):
        F.check(F.lib().sdr_comm_bcast_u8(self._h, d_buf.at(offset), nbytes, root))

    def chan_wait(self, ch: Channeliser):
        F.check(F.lib().sdr_comm_chan_wait(self._h, ch._h))

    def wait_chan(self, ch: Channeliser):
        F.check(F.lib().sdr_comm_wait_chan(self._h, ch._h))

    def mark_chan(self, ch: Channeliser, slot: int):
        F.check(F.lib().sdr_comm_mark_chan(self._h, ch._h, slot))

    def wait_mark(self, slot: int):
        F.check(F.lib().sdr_comm_wait_mark(self._h, slot))

    def sync(self):
        F.check(F.lib().sdr_comm_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            F.lib().sdr_comm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
