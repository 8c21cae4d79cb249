"""`Source` — the read_sync-shaped buffer source (RtlSdr::read_sync, src/lib.rs:153-155)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class Source:
    def __init__(self, handle):
        self._h = handle
        self._cb_keepalive = None

    @classmethod
    def open_file(cls, path: str, loop: bool = False) -> "Source":
        h = C.c_void_p()
        F.check(F.lib().sdr_source_open_file(str(path).encode(), int(loop), C.byref(h)))
        return cls(h)

    @classmethod
    def open_synth(cls, seed: int, total_bytes: int = 0) -> "Source":
        h = C.c_void_p()
        F.check(F.lib().sdr_source_open_synth(seed, total_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def open_rtl_tcp(cls, host: str, port: int = 1234) -> "Source":
        """rtl_tcp client (wire format of examples/rtl_tcp.rs:633-697)."""
        h = C.c_void_p()
        F.check(F.lib().sdr_source_open_rtl_tcp(host.encode(), port, C.byref(h)))
        return cls(h)

    def rtl_tcp_info(self):
        t, g = C.c_uint32(0), C.c_uint32(0)
        F.check(F.lib().sdr_source_rtl_tcp_info(self._h, C.byref(t), C.byref(g)))
        return t.value, g.value

    def rtl_tcp_command(self, cmd: int, param: int):
        F.check(F.lib().sdr_source_rtl_tcp_command(self._h, cmd, param & 0xFFFFFFFF))

    def read_sync(self, buf: np.ndarray) -> int:
        """read_sync(&self, buf: &mut [u8]) -> Result<usize>: fills the caller's buffer, returns bytes read."""
        assert buf.dtype == np.uint8 and buf.flags.c_contiguous
        return F.check(F.lib().sdr_source_read_sync(self._h, F.ptr(buf), buf.size))

    def read_async(self, callback, buf_num: int = 15, buf_len: int = 16 * 16384):
        """Blocks; callback(np.ndarray[u8]) runs for every full buffer until the source ends or cancel()."""
        def tramp(p, n, _ctx):
            callback(np.ctypeslib.as_array(p, shape=(n,)))
        cb = F.READ_ASYNC_CB(tramp)
        self._cb_keepalive = cb
        F.check(F.lib().sdr_source_read_async(self._h, cb, None, buf_num, buf_len))

    def cancel(self):
        F.check(F.lib().sdr_source_cancel_async(self._h))

    def close(self):
        if self._h:
            F.lib().sdr_source_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
