"""ctypes declarations for include/sdr_b200.h (one entry per exported symbol)."""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libsdr_b200.so"

SDR_OK, SDR_E_ARG, SDR_E_LEN, SDR_E_CAP, SDR_E_CUDA, SDR_E_NCCL, SDR_E_IO, SDR_E_STATE = 0, -1, -2, -3, -4, -5, -6, -7
_CODE_NAMES = {-1: "SDR_E_ARG", -2: "SDR_E_LEN", -3: "SDR_E_CAP", -4: "SDR_E_CUDA", -5: "SDR_E_NCCL",
               -6: "SDR_E_IO", -7: "SDR_E_STATE"}
NCCL_ID_BYTES = 128


class SdrError(RuntimeError):
    """Raised for any negative return of the C ABI (the Rust wrapper maps these to
    RtlsdrError::RtlsdrErr(String), src/error.rs:43)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{_CODE_NAMES.get(code, code)}: {msg}")
        self.code = code


class DemodConfig(C.Structure):
    """sdr_demod_config == DemodConfig, examples/simple_fm.rs:179-185."""
    _fields_ = [(n, C.c_uint32) for n in ("rate_in", "rate_out", "rate_resample", "downsample", "output_scale")]


class RadioConfig(C.Structure):
    """sdr_radio_config == RadioConfig, examples/simple_fm.rs:173-176."""
    _fields_ = [("capture_freq", C.c_uint32), ("capture_rate", C.c_uint32)]


class DemodState(C.Structure):
    """sdr_demod_state == the carried fields of struct Demod, examples/simple_fm.rs:234-238."""
    _fields_ = [("prev_index", C.c_uint64), ("now_lpr", C.c_int32), ("prev_lpr_index", C.c_int32),
                ("lp_now_re", C.c_int32), ("lp_now_im", C.c_int32),
                ("demod_pre_re", C.c_int32), ("demod_pre_im", C.c_int32)]


class FmrxConfig(C.Structure):
    _fields_ = [("n_taps", C.c_uint32), ("decim", C.c_uint32), ("n_taps2", C.c_uint32),
                ("up", C.c_uint32), ("down", C.c_uint32), ("gain", C.c_float)]


class PostConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("output_scale", "squelch_level", "deemph_a", "dc_block")]


class ChanConfig(C.Structure):
    _fields_ = [("n_channels", C.c_uint32), ("n_taps", C.c_uint32), ("decim", C.c_uint32), ("gain", C.c_float)]


READ_ASYNC_CB = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p)

_vp, _sz, _i, _l = C.c_void_p, C.c_size_t, C.c_int, C.c_long
_u8p, _i16p, _i32p, _u32p, _f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_int32, C.c_uint32, C.c_float))

# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/sdr_b200.h.
SIGNATURES = {
    "sdr_last_error": (C.c_char_p, []),
    "sdr_abi_version": (_i, []),
    "sdr_device_count": (_i, []),
    "sdr_device_info": (_i, [_i, C.c_char_p, _sz, C.POINTER(_i), C.POINTER(C.c_uint64)]),
    "sdr_kernel_launch_count": (C.c_uint64, []),
    "sdr_dev_alloc": (_vp, [_i, _sz]),
    "sdr_dev_free": (None, [_i, _vp]),
    "sdr_host_alloc": (_vp, [_sz]),
    "sdr_host_free": (None, [_vp]),
    "sdr_host_register": (_i, [_vp, _sz]),
    "sdr_host_unregister": (_i, [_vp]),
    "sdr_memcpy_h2d": (_i, [_i, _vp, _vp, _sz]),
    "sdr_memcpy_d2h": (_i, [_i, _vp, _vp, _sz]),
    "sdr_dev_memset": (_i, [_i, _vp, _i, _sz]),
    "sdr_synth_fill_dev": (_i, [_i, _vp, _sz, C.c_uint64, C.c_uint64]),
    "sdr_device_sync": (_i, [_i]),
    "sdr_optimal_settings": (_i, [C.c_uint32] * 4 + [C.POINTER(RadioConfig), C.POINTER(DemodConfig)]),
    "sdr_demod_new": (_i, [C.POINTER(DemodConfig), _i, C.POINTER(_vp)]),
    "sdr_demod_free": (None, [_vp]),
    "sdr_demod_get_state": (_i, [_vp, C.POINTER(DemodState)]),
    "sdr_demod_set_state": (_i, [_vp, C.POINTER(DemodState)]),
    "sdr_demod_out_len": (_l, [_vp, _sz]),
    "sdr_demod_demodulate": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_demod_demodulate_batch": (_l, [_vp, _vp, _sz, _sz, _vp, _sz, _u32p]),
    "sdr_demod_demodulate_batch_dev": (_l, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "sdr_demod_sync": (_i, [_vp]),
    "sdr_demod_last_timing": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "sdr_demod_span_begin": (_i, [_vp]),
    "sdr_demod_span_end": (_i, [_vp, C.POINTER(C.c_float)]),
    "sdr_demod_ring_open": (_i, [_vp, _sz, C.c_uint32, C.POINTER(_vp)]),
    "sdr_ring_acquire": (_i, [_vp, C.POINTER(_vp)]),
    "sdr_ring_commit": (_i, [_vp]),
    "sdr_ring_collect": (_l, [_vp, _vp, _sz]),
    "sdr_ring_close": (_i, [_vp]),
    "sdr_rotate_90": (_l, [_vp, _vp, _sz]),
    "sdr_buf_to_complex": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_low_pass_complex": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fm_demod": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_low_pass_real": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fast_atan2": (_l, [_vp, _vp, _vp, _sz, _vp]),
    "sdr_polar_discriminant": (_l, [_vp, _vp, _vp, _sz, _i, _vp]),
    "sdr_post_deemph_a": (C.c_uint32, [C.c_uint32, C.c_double]),
    "sdr_post_new": (_i, [C.POINTER(PostConfig), _i, C.POINTER(_vp)]),
    "sdr_post_free": (None, [_vp]),
    "sdr_post_process": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fmrx_new": (_i, [C.POINTER(FmrxConfig), _vp, _vp, _i, C.POINTER(_vp)]),
    "sdr_fmrx_free": (None, [_vp]),
    "sdr_fmrx_reset": (_i, [_vp]),
    "sdr_fmrx_out_lens": (_i, [_vp, _sz, C.POINTER(_sz), C.POINTER(_sz)]),
    "sdr_fmrx_process": (_l, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "sdr_fmrx_process_dev": (_l, [_vp, _vp, _sz, _vp, _vp, _vp, _sz]),
    "sdr_fmrx_low_pass": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fmrx_fm_demod": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fmrx_resample": (_l, [_vp, _vp, _sz, _vp, _sz]),
    "sdr_fmrx_sync": (_i, [_vp]),
    "sdr_fmrx_last_timing": (_i, [_vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_uint32), C.POINTER(_i)]),
    "sdr_fmrx_timing_totals": (_i, [_vp, C.POINTER(C.c_double * 3), C.POINTER(C.c_uint64), _i]),
    "sdr_fmrx_kernel_kind": (_i, [_vp, C.POINTER(C.c_char_p)]),
    "sdr_rtc_selftest": (_l, [C.c_uint32, C.c_uint32, C.POINTER(_i * 4)]),
    "sdr_rtc_pick_shape": (_i, [C.c_uint32, C.c_uint32, C.POINTER(_i * 4)]),
    "sdr_rtc_compile_ring": (_l, [C.c_uint32, C.c_uint32]),
    "sdr_rtc_compile_slide": (_l, [C.c_uint32, C.c_uint32, C.POINTER(_i)]),
    "sdr_fmrx_ring_open": (_i, [_vp, _sz, C.c_uint32, C.POINTER(_vp)]),
    "sdr_fmrx_ring_acquire": (_i, [_vp, C.POINTER(_vp)]),
    "sdr_fmrx_ring_commit": (_i, [_vp]),
    "sdr_fmrx_ring_collect": (_l, [_vp, _vp, _sz]),
    "sdr_fmrx_ring_close": (_i, [_vp]),
    "sdr_fmrx_span_begin": (_i, [_vp]),
    "sdr_fmrx_span_end": (_i, [_vp, C.POINTER(C.c_float)]),
    "sdr_fmrx_seek": (_i, [_vp, C.c_uint64]),
    "sdr_fmrx_plan": (_i, [C.POINTER(FmrxConfig), C.c_uint64, _sz, C.POINTER(C.c_uint64), C.POINTER(_sz),
                           C.POINTER(C.c_uint64), C.POINTER(_sz)]),
    "sdr_demod_plan": (_i, [C.POINTER(DemodConfig), C.POINTER(DemodState), _sz, _sz, C.POINTER(_sz), C.POINTER(_sz),
                            C.POINTER(DemodState)]),
    "sdr_shard_range": (_i, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sdr_chan_new": (_i, [C.POINTER(ChanConfig), _vp, _vp, _i, C.POINTER(_vp)]),
    "sdr_chan_free": (None, [_vp]),
    "sdr_chan_reset": (_i, [_vp]),
    "sdr_chan_process": (_l, [_vp, _vp, _sz, _vp, _vp, _sz]),
    "sdr_chan_process_dev": (_l, [_vp, _vp, _sz, _vp, _vp, _sz]),
    "sdr_chan_sync": (_i, [_vp]),
    "sdr_chan_last_timing": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "sdr_chan_kernel_kind": (_i, [_vp, C.POINTER(C.c_uint32 * 4)]),
    "sdr_chan_bank_plan": (_l, [C.POINTER(ChanConfig), _vp, _vp, C.POINTER(C.c_uint32 * 4), _vp, _sz]),
    "sdr_comm_unique_id": (_i, [_vp]),
    "sdr_comm_init": (_i, [_i, _i, _i, _vp, C.POINTER(_vp)]),
    "sdr_comm_bcast_u8": (_i, [_vp, _vp, _sz, _i]),
    "sdr_comm_chan_wait": (_i, [_vp, _vp]),
    "sdr_comm_wait_chan": (_i, [_vp, _vp]),
    "sdr_comm_mark_chan": (_i, [_vp, _vp, C.c_uint32]),
    "sdr_comm_wait_mark": (_i, [_vp, C.c_uint32]),
    "sdr_comm_sync": (_i, [_vp]),
    "sdr_comm_free": (None, [_vp]),
    "sdr_source_open_file": (_i, [C.c_char_p, _i, C.POINTER(_vp)]),
    "sdr_source_open_synth": (_i, [C.c_uint64, C.c_uint64, C.POINTER(_vp)]),
    "sdr_source_open_rtl_tcp": (_i, [C.c_char_p, C.c_uint16, C.POINTER(_vp)]),
    "sdr_source_rtl_tcp_info": (_i, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "sdr_source_rtl_tcp_command": (_i, [_vp, C.c_uint8, C.c_uint32]),
    "sdr_source_read_sync": (_l, [_vp, _vp, _sz]),
    "sdr_source_read_async": (_i, [_vp, READ_ASYNC_CB, _vp, C.c_uint32, C.c_uint32]),
    "sdr_source_cancel_async": (_i, [_vp]),
    "sdr_source_close": (None, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libsdr_b200.so; fail loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = pathlib.Path(os.environ.get("SDR_B200_LIB", LIB_PATH))   # override: A/B builds of the same ABI
    if not path.exists():
        raise SdrError(SDR_E_STATE, f"{path} is missing: build it with `make -C {PKG_DIR}` "
                                    "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if L.sdr_abi_version() != 1:
        raise SdrError(SDR_E_STATE, "ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise SdrError(int(rc), (lib().sdr_last_error() or b"").decode(errors="replace"))
    return int(rc)


def ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def device_count() -> int:
    return int(lib().sdr_device_count())


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(128)
    sm, mem = C.c_int(0), C.c_uint64(0)
    check(lib().sdr_device_info(device, name, 128, C.byref(sm), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "mem_bytes": mem.value}


def kernel_launch_count() -> int:
    return int(lib().sdr_kernel_launch_count())


def optimal_settings(freq: int = 94_900_000, rate: int = 170_000, sample_rate: int = 170_000,
                     rate_resample: int = 32_000):
    """optimal_settings(freq, rate), examples/simple_fm.rs:189-214 (defaults = the example's constants :25-27)."""
    r, c = RadioConfig(), DemodConfig()
    check(lib().sdr_optimal_settings(freq, rate, sample_rate, rate_resample, C.byref(r), C.byref(c)))
    return r, c


class DevBuffer:
    """Device memory from sdr_dev_alloc (16-B aligned, head/tail room)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, int(nbytes)
        p = lib().sdr_dev_alloc(device, self.nbytes)
        if not p:
            raise SdrError(SDR_E_CUDA, (lib().sdr_last_error() or b"").decode())
        self.ptr = C.c_void_p(p)

    def upload(self, a: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(a)
        assert offset + a.nbytes <= self.nbytes
        check(lib().sdr_memcpy_h2d(self.device, C.c_void_p(self.ptr.value + offset), ptr(a), a.nbytes))
        return self

    def download(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        out = np.empty(count, dtype)
        assert offset + out.nbytes <= self.nbytes
        check(lib().sdr_memcpy_d2h(self.device, ptr(out), C.c_void_p(self.ptr.value + offset), out.nbytes))
        return out

    def at(self, offset: int) -> C.c_void_p:
        return C.c_void_p(self.ptr.value + offset)

    def free(self):
        if self.ptr:
            lib().sdr_dev_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_register(arr: np.ndarray) -> None:
    """Page-lock a numpy array the caller owns (sdr_host_register): later calls DMA straight from it."""
    check(lib().sdr_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes))


def host_unregister(arr: np.ndarray) -> None:
    check(lib().sdr_host_unregister(C.c_void_p(arr.ctypes.data)))


class HostBuffer:
    """Pinned host memory from sdr_host_alloc, exposed as a numpy array."""

    def __init__(self, nbytes: int, dtype=np.uint8):
        self.nbytes = int(nbytes)
        p = lib().sdr_host_alloc(self.nbytes)
        if not p:
            raise SdrError(SDR_E_CUDA, (lib().sdr_last_error() or b"").decode())
        self.ptr = C.c_void_p(p)
        raw = (C.c_uint8 * self.nbytes).from_address(p)
        self.array = np.frombuffer(raw, dtype=np.uint8).view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().sdr_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synth_fill_dev(buf: DevBuffer, nbytes: int, seed: int, byte_offset: int = 0, dst_offset: int = 0):
    check(lib().sdr_synth_fill_dev(buf.device, buf.at(dst_offset), nbytes, seed, byte_offset))


def fmrx_plan(cfg: FmrxConfig, n_in0: int, n_samples: int):
    """(y0, n_y, a0, n_audio) of a call at global sample n_in0 — pure host arithmetic."""
    y0, a0, ny, na = C.c_uint64(0), C.c_uint64(0), C.c_size_t(0), C.c_size_t(0)
    check(lib().sdr_fmrx_plan(C.byref(cfg), n_in0, n_samples, C.byref(y0), C.byref(ny), C.byref(a0), C.byref(na)))
    return y0.value, ny.value, a0.value, na.value


def demod_plan(cfg: DemodConfig, buf_len: int, n_bufs: int, state: DemodState | None = None):
    """(n_lowpassed, n_audio, DemodState-after[index fields]) — pure host arithmetic."""
    nl, na, after = C.c_size_t(0), C.c_size_t(0), DemodState()
    check(lib().sdr_demod_plan(C.byref(cfg), C.byref(state) if state is not None else None, buf_len, n_bufs,
                               C.byref(nl), C.byref(na), C.byref(after)))
    return nl.value, na.value, after


def shard_range(total: int, world: int, rank: int, align: int = 1):
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    check(lib().sdr_shard_range(total, world, rank, align, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value
