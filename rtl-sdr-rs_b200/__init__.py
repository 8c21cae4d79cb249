"""rtl-sdr-rs_b200 — B200-native IQ-sample DSP hot path (host-side Python binding over the C ABI).

The product is `lib/libsdr_b200.so` (hand-written CUDA for sm_100a, `csrc/`) behind the C ABI of
`include/sdr_b200.h`.  This package is a thin ctypes mirror of that ABI whose classes keep the
names of the reference's operator interface (`Demod.demodulate / rotate_90 / low_pass_complex /
fm_demod / low_pass_real / fast_atan2`, examples/simple_fm.rs:242-427) so that parity tests read
like the reference's own tests.  It contains no arithmetic: there is NO CPU fallback — if the
shared library is missing or no CUDA device is visible, compute calls raise `SdrError`.

The directory name contains a hyphen, so import it through `sdrpkg.load()` at the repo root
(registers the package as `rtl_sdr_rs_b200`).
"""
from ._ffi import (  # noqa: F401
    LIB_PATH,
    SdrError,
    DemodConfig,
    DemodState,
    RadioConfig,
    FmrxConfig,
    ChanConfig,
    PostConfig,
    lib,
    device_count,
    device_info,
    kernel_launch_count,
    optimal_settings,
    DevBuffer,
    HostBuffer,
    host_register,
    host_unregister,
    synth_fill_dev,
    fmrx_plan,
    demod_plan,
    shard_range,
)
from .demod import Demod, Ring, AudioPost  # noqa: F401
from .fmrx import FmRx, FmRing  # noqa: F401
from .chan import Channeliser, Comm, bank_plan  # noqa: F401
from .source import Source  # noqa: F401

DEFAULT_BUF_LENGTH = 16 * 16384  # src/lib.rs:25
