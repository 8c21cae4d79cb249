"""ctypes binding for the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATH = ORACLE_DIR / "_build" / "liboracle.so"


class DemodConfig(C.Structure):
    """DemodConfig, examples/simple_fm.rs:179-185."""

    _fields_ = [(n, C.c_uint32) for n in ("rate_in", "rate_out", "rate_resample", "downsample", "output_scale")]


class RadioConfig(C.Structure):
    _fields_ = [("capture_freq", C.c_uint32), ("capture_rate", C.c_uint32)]


class DemodState(C.Structure):
    """struct Demod, examples/simple_fm.rs:232-239 (oracle layout)."""

    _fields_ = [
        ("config", DemodConfig),
        ("prev_index", C.c_uint64),
        ("now_lpr", C.c_int32),
        ("prev_lpr_index", C.c_int32),
        ("lp_now_re", C.c_int32),
        ("lp_now_im", C.c_int32),
        ("demod_pre_re", C.c_int32),
        ("demod_pre_im", C.c_int32),
    ]


class Post(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("output_scale", "squelch_level", "deemph_a", "dc_block")] + \
               [("deemph_avg", C.c_int32), ("dc_avg", C.c_int32)]


class Fx(C.Structure):
    _fields_ = [
        ("n_taps", C.c_uint32), ("decim", C.c_uint32),
        ("up", C.c_uint32), ("down", C.c_uint32), ("n_taps2", C.c_uint32),
        ("gain", C.c_double),
        ("n_in", C.c_uint64), ("n_y", C.c_uint64), ("n_a", C.c_uint64),
        ("taps", C.c_void_p), ("taps2", C.c_void_p),
        ("hist_re", C.c_void_p), ("hist_im", C.c_void_p),
        ("prev_re", C.c_double), ("prev_im", C.c_double),
        ("dhist", C.c_void_p), ("dhist_len", C.c_size_t),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src_m = max((ORACLE_DIR / f).stat().st_mtime for f in ("sdr_oracle.c", "sdr_oracle.h", "Makefile"))
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    u8p, i16p, i32p, f64p, f32p, u32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_int32, C.c_double, C.c_float, C.c_uint32))
    sz = C.c_size_t
    L.orc_optimal_settings.argtypes = [C.c_uint32] * 4 + [C.POINTER(RadioConfig), C.POINTER(DemodConfig)]
    L.orc_optimal_settings.restype = None
    L.orc_demod_init.argtypes = [C.POINTER(DemodState), C.POINTER(DemodConfig)]
    L.orc_rotate_90.argtypes = [u8p, sz]
    L.orc_centre.argtypes = [u8p, sz, i16p]
    L.orc_buf_to_complex.argtypes = [i16p, sz, i32p]
    L.orc_buf_to_complex.restype = sz
    L.orc_low_pass_complex.argtypes = [C.POINTER(DemodState), i32p, sz, i32p]
    L.orc_low_pass_complex.restype = sz
    L.orc_fast_atan2.argtypes = [C.c_int32, C.c_int32]
    L.orc_fast_atan2.restype = C.c_int32
    for f in (L.orc_polar_discriminant, L.orc_polar_discriminant_fast):
        f.argtypes = [C.c_int32] * 4
        f.restype = C.c_int32
    L.orc_fm_demod.argtypes = [C.POINTER(DemodState), i32p, sz, i16p]
    L.orc_fm_demod.restype = C.c_long
    L.orc_low_pass_real.argtypes = [C.POINTER(DemodState), i16p, sz, i16p]
    L.orc_low_pass_real.restype = sz
    L.orc_demodulate.argtypes = [C.POINTER(DemodState), u8p, sz, i16p, i32p, i16p, C.POINTER(sz)]
    L.orc_demodulate.restype = C.c_long
    L.orc_demodulate_ref_like.argtypes = [C.POINTER(DemodState), u8p, sz, i16p]
    L.orc_demodulate_ref_like.restype = C.c_long
    L.orc_demodulate_fused.argtypes = [C.POINTER(DemodState), u8p, sz, i16p]
    L.orc_demodulate_fused.restype = C.c_long
    L.orc_demodulate_many_mt.argtypes = [C.POINTER(DemodConfig), u8p, sz, sz, i16p, sz, C.c_int]
    L.orc_demodulate_many_mt.restype = C.c_long
    L.orc_demodulate_many_mt2.argtypes = [C.POINTER(DemodConfig), u8p, sz, sz, i16p, sz, C.c_int, C.c_int]
    L.orc_demodulate_many_mt2.restype = C.c_long
    L.orc_post_init.argtypes = [C.POINTER(Post)] + [C.c_uint32] * 4
    L.orc_post_process.argtypes = [C.POINTER(Post), i16p, sz, u8p, sz]
    L.orc_fx_init.argtypes = [C.POINTER(Fx), f32p, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
    L.orc_fx_init.restype = C.c_int
    L.orc_fx_free.argtypes = [C.POINTER(Fx)]
    L.orc_fx_low_pass.argtypes = [C.POINTER(Fx), u8p, sz, f64p]
    L.orc_fx_low_pass.restype = sz
    L.orc_fx_fm_demod.argtypes = [C.POINTER(Fx), f64p, sz, f64p]
    L.orc_fx_fm_demod.restype = sz
    L.orc_fx_resample.argtypes = [C.POINTER(Fx), f64p, sz, f64p]
    L.orc_fx_resample.restype = sz
    L.orc_fx_process.argtypes = [C.POINTER(Fx), u8p, sz, f64p, f64p, f64p, C.POINTER(sz)]
    L.orc_fx_process.restype = sz
    L.orc_fx_channelise.argtypes = [u8p, sz, f32p, C.c_uint32, C.c_uint32, u32p, C.c_uint32, C.c_double, f64p, f64p]
    L.orc_fx_channelise.restype = sz
    L.orc_fx_process_f32_mt.argtypes = [u8p, sz, f32p, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, f32p, sz, C.c_int]
    L.orc_fx_process_f32_mt.restype = sz
    L.orc_synth_fill.argtypes = [u8p, sz, C.c_uint64, C.c_uint64]
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


# ---------------------------------------------------------------------------------------
# numpy-level wrappers named after the reference's functions (examples/simple_fm.rs)

# the example's constants, examples/simple_fm.rs:25-27
FREQUENCY, SAMPLE_RATE, RATE_RESAMPLE = 94_900_000, 170_000, 32_000
DEFAULT_BUF_LENGTH = 16 * 16384  # src/lib.rs:25


def optimal_settings(freq=FREQUENCY, rate=SAMPLE_RATE, sample_rate_const=SAMPLE_RATE, rate_resample=RATE_RESAMPLE):
    r, c = RadioConfig(), DemodConfig()
    lib().orc_optimal_settings(freq, rate, sample_rate_const, rate_resample, C.byref(r), C.byref(c))
    return r, c


class Demod:
    """Oracle Demod: same method names as the reference's struct (examples/simple_fm.rs:242-427)."""

    def __init__(self, config: DemodConfig | None = None):
        if config is None:
            _, config = optimal_settings()
        self.st = DemodState()
        lib().orc_demod_init(C.byref(self.st), C.byref(config))

    @staticmethod
    def rotate_90(buf: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(buf, dtype=np.uint8).copy()
        lib().orc_rotate_90(_p(b, C.c_uint8), b.size)
        return b

    @staticmethod
    def fast_atan2(y: int, x: int) -> int:
        return lib().orc_fast_atan2(int(y), int(x))

    def low_pass_complex(self, pairs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty((x.shape[0] + 1, 2), np.int32)
        n = lib().orc_low_pass_complex(C.byref(self.st), _p(x, C.c_int32), x.shape[0], _p(out, C.c_int32))
        return out[:n].copy()

    def fm_demod(self, pairs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty(max(x.shape[0], 1), np.int16)
        n = lib().orc_fm_demod(C.byref(self.st), _p(x, C.c_int32), x.shape[0], _p(out, C.c_int16))
        if n < 0:
            raise ValueError("fm_demod needs more than one sample (examples/simple_fm.rs:356)")
        return out[:n].copy()

    def low_pass_real(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty(x.size + 1, np.int16)
        n = lib().orc_low_pass_real(C.byref(self.st), _p(x, C.c_int16), x.size, _p(out, C.c_int16))
        return out[:n].copy()

    def demodulate(self, buf: np.ndarray, stages: bool = False, ref_like: bool = False, fused: bool = False):
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        out = np.empty(b.size // 2 + 1, np.int16)
        if ref_like or fused:
            fn = lib().orc_demodulate_fused if fused else lib().orc_demodulate_ref_like
            n = fn(C.byref(self.st), _p(b, C.c_uint8), b.size, _p(out, C.c_int16))
            if n < 0:
                raise ValueError("demodulate: reference would panic on this input")
            return out[:n].copy()
        lp = np.empty((b.size // 2 + 1, 2), np.int32)
        dm = np.empty(b.size // 2 + 1, np.int16)
        nlp = C.c_size_t(0)
        n = lib().orc_demodulate(C.byref(self.st), _p(b, C.c_uint8), b.size, _p(out, C.c_int16),
                                 _p(lp, C.c_int32), _p(dm, C.c_int16), C.byref(nlp))
        if n < 0:
            raise ValueError("demodulate: reference would panic on this input")
        if stages:
            return out[:n].copy(), lp[: nlp.value].copy(), dm[: nlp.value].copy()
        return out[:n].copy()

    def set_state(self, prev_index=0, now_lpr=0, prev_lpr_index=0, lp_now=(0, 0), demod_pre=(0, 0)):
        s = self.st
        s.prev_index, s.now_lpr, s.prev_lpr_index = prev_index, now_lpr, prev_lpr_index
        s.lp_now_re, s.lp_now_im = lp_now
        s.demod_pre_re, s.demod_pre_im = demod_pre

    def state(self) -> dict:
        s = self.st
        return dict(prev_index=int(s.prev_index), now_lpr=s.now_lpr, prev_lpr_index=s.prev_lpr_index,
                    lp_now=(s.lp_now_re, s.lp_now_im), demod_pre=(s.demod_pre_re, s.demod_pre_im))


class AudioPost:
    """Oracle of the optional post-stages (orc_post_*)."""

    def __init__(self, output_scale=0, squelch_level=0, deemph_a=0, dc_block=False):
        self.s = Post()
        lib().orc_post_init(C.byref(self.s), output_scale, squelch_level, deemph_a, int(dc_block))

    def process(self, audio: np.ndarray, raw: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(audio, np.int16).copy()
        r = np.ascontiguousarray(raw, np.uint8) if raw is not None else None
        lib().orc_post_process(C.byref(self.s), _p(a, C.c_int16), a.size, _p(r, C.c_uint8) if r is not None else None,
                               r.size if r is not None else 0)
        return a


def buf_to_complex(i16: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(i16, dtype=np.int16)
    out = np.empty((x.size // 2, 2), np.int32)
    lib().orc_buf_to_complex(_p(x, C.c_int16), x.size, _p(out, C.c_int32))
    return out


class FxChain:
    """f64 oracle of the tap'd FIR -> discriminator -> rational resampler chain."""

    def __init__(self, taps, decim, taps2=None, up=1, down=1, gain=16384.0 / np.pi):
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.taps2 = np.ascontiguousarray(taps2 if taps2 is not None else [], np.float32)
        self.s = Fx()
        t2 = _p(self.taps2, C.c_float) if self.taps2.size else None
        rc = lib().orc_fx_init(C.byref(self.s), _p(self.taps, C.c_float), self.taps.size, decim, t2,
                               self.taps2.size, up, down, float(gain))
        if rc != 0:
            raise ValueError("bad FxChain parameters")
        self.decim = decim

    def __del__(self):
        try:
            lib().orc_fx_free(C.byref(self.s))
        except Exception:
            pass

    def process(self, iq_u8: np.ndarray):
        b = np.ascontiguousarray(iq_u8, np.uint8)
        n = b.size // 2
        cap = n // self.decim + 2
        y = np.empty((cap, 2), np.float64)
        d = np.empty(cap, np.float64)
        acap = cap * int(self.s.up) // int(self.s.down) + 2
        a = np.empty(acap, np.float64)
        ny = C.c_size_t(0)
        na = lib().orc_fx_process(C.byref(self.s), _p(b, C.c_uint8), n, _p(y, C.c_double), _p(d, C.c_double),
                                  _p(a, C.c_double), C.byref(ny))
        return y[: ny.value].copy(), d[: ny.value].copy(), a[:na].copy()


def channelise(iq_u8, taps, decim, freq_words, gain=16384.0 / np.pi):
    b = np.ascontiguousarray(iq_u8, np.uint8)
    t = np.ascontiguousarray(taps, np.float32)
    fw = np.ascontiguousarray(freq_words, np.uint32)
    n = b.size // 2
    M = n // decim
    y = np.empty((fw.size, M, 2), np.float64)
    d = np.empty((fw.size, M), np.float64)
    lib().orc_fx_channelise(_p(b, C.c_uint8), n, _p(t, C.c_float), t.size, decim, _p(fw, C.c_uint32), fw.size,
                            float(gain), _p(y, C.c_double), _p(d, C.c_double))
    return y, d


def synth_fill(n_bytes: int, seed: int, byte_offset: int = 0) -> np.ndarray:
    out = np.empty(n_bytes, np.uint8)
    lib().orc_synth_fill(_p(out, C.c_uint8), n_bytes, seed, byte_offset)
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
