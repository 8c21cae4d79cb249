"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/sdr_b200.h declares, the ctypes table covers exactly that set, and compute entry points fail
LOUDLY (SDR_E_CUDA) instead of falling back to anything when no device is present."""
import ctypes as C
import os
import re
from pathlib import Path

import numpy as np
import pytest

import sdrpkg

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "sdr_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    text = re.sub(r"typedef\s+void\s*\(\*\w+\)\([^;]*\);", "", text)
    return sorted(set(re.findall(r"\b(sdr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def S():
    return sdrpkg.load()


def test_library_loads_and_exports_every_declared_symbol(S):
    names = declared_functions()
    assert len(names) >= 55
    L = C.CDLL(str(S.LIB_PATH))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/sdr_b200.h but not exported: {missing}"
    assert S.lib().sdr_abi_version() == 1


def test_ctypes_table_matches_header(S):
    from rtl_sdr_rs_b200 import _ffi
    assert sorted(_ffi.SIGNATURES) == declared_functions()


def test_header_cites_the_reference_interface_it_replaces():
    text = HEADER.read_text()
    for cite in ("src/lib.rs:153", ":256-269", "examples/simple_fm.rs:179", ":337-352", ":355-367",
                 ":408-426", ":383-405", ":276-299", "src/error.rs"):
        assert cite in text, cite


def test_optimal_settings_is_host_arithmetic(S):
    r, c = S.optimal_settings()            # examples/simple_fm.rs:189-214
    assert (c.downsample, c.output_scale, r.capture_rate, r.capture_freq) == (6, 42, 1_020_000, 95_155_000)
    r, c = S.optimal_settings(100_000_000, 2_400_000, 2_400_000, 48_000)
    assert (c.downsample, c.output_scale) == (1, 256)


def test_no_cpu_fallback_without_a_device(S):
    if S.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(S.SdrError) as e:
        S.Demod()
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    with pytest.raises(S.SdrError):
        S.DevBuffer(1024)


def test_package_has_no_oracle_dependency():
    """The product package must never import or link the oracle (parity would be void)."""
    pkg = ROOT / "rtl-sdr-rs_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")) + list(pkg.glob("host/*")):
        if f.is_file():
            t = f.read_text(errors="ignore")
            assert "oracle_ffi" not in t and "liboracle" not in t and "sdr_oracle" not in t, f


def test_source_read_sync_file_and_synth(S, tmp_path, golden_dir):
    import oracle_ffi as O
    head = golden_dir / "capture_head.bin"
    src = S.Source.open_file(head)
    buf = np.empty(262144, np.uint8)
    total, first = 0, None
    while True:
        n = src.read_sync(buf)
        if first is None:
            first = buf[:16].copy()
        total += n
        if n < buf.size:           # short read == end of data (examples/simple_fm.rs:121-125)
            break
    assert total == head.stat().st_size and first.tolist() == np.fromfile(head, np.uint8, 16).tolist()
    src.close()
    syn = S.Source.open_synth(0xB2000001, total_bytes=1000)
    b = np.empty(600, np.uint8)
    assert syn.read_sync(b) == 600 and np.array_equal(b, O.synth_fill(600, 0xB2000001))
    assert syn.read_sync(b) == 400 and np.array_equal(b[:400], O.synth_fill(400, 0xB2000001, 600))
    assert syn.read_sync(b) == 0
    with pytest.raises(S.SdrError):
        S.Source.open_file(tmp_path / "does-not-exist.bin")


def test_source_read_async_delivers_every_full_buffer_in_order(S):
    import oracle_ffi as O
    n_bufs, blen = 37, 4096
    syn = S.Source.open_synth(42, total_bytes=n_bufs * blen + 100)   # trailing short read is dropped
    got = []
    syn.read_async(lambda a: got.append(a.copy()), buf_num=4, buf_len=blen)
    assert len(got) == n_bufs
    assert np.array_equal(np.concatenate(got), O.synth_fill(n_bufs * blen, 42))
    # cancel from inside the callback stops delivery
    syn2 = S.Source.open_synth(43)
    seen = []

    def cb(a):
        seen.append(a[0])
        if len(seen) == 5:
            syn2.cancel()
    syn2.read_async(cb, buf_num=3, buf_len=1024)
    assert 5 <= len(seen) <= 8


def test_source_rtl_tcp_wire_format(S):
    """A local server speaking examples/rtl_tcp.rs's protocol: 12-byte RTL0 greeting, raw IQ, 5-byte BE commands."""
    import socket
    import struct
    import threading
    import oracle_ffi as O
    payload = O.synth_fill(3 * 4096 + 100, 77).tobytes()
    got_cmds = []
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]

    def serve():
        c, _ = srv.accept()
        c.sendall(b"RTL0" + struct.pack(">II", 5, 29))            # R820T, 29 gain steps (send_handshake :691-697)
        for _ in range(2):
            got_cmds.append(struct.unpack(">BI", c.recv(5, socket.MSG_WAITALL)))
        c.sendall(payload)
        c.close()
    t = threading.Thread(target=serve, daemon=True)
    t.start()
    src = S.Source.open_rtl_tcp("127.0.0.1", port)
    assert src.rtl_tcp_info() == (5, 29)
    src.rtl_tcp_command(0x01, 94_900_000 + 255_000)               # set frequency (:658)
    src.rtl_tcp_command(0x02, 1_020_000)                          # set sample rate (:659)
    buf = np.empty(4096, np.uint8)
    chunks = []
    while True:
        n = src.read_sync(buf)
        chunks.append(buf[:n].copy())
        if n < buf.size:
            break
    t.join(5)
    assert b"".join(c.tobytes() for c in chunks) == payload
    assert got_cmds == [(0x01, 95_155_000), (0x02, 1_020_000)]
    src.close()
    srv.close()
    with pytest.raises(S.SdrError):
        S.Source.open_rtl_tcp("127.0.0.1", 1)                     # nothing listens there


def test_rtc_compiles_a_new_shape_without_a_gpu(tmp_path, monkeypatch):
    """NVRTC specialisation of k_fir_fast (csrc/rtc.cpp): compile-only self-test, no device needed.  A fresh cache
    directory forces a real compile; the second call must be served from the on-disk cache."""
    import ctypes as C
    import subprocess
    import sys
    code = (
        "import ctypes as C, sys; sys.path.insert(0, %r)\n"
        "import sdrpkg; sdrpkg.load()\n"
        "from rtl_sdr_rs_b200 import _ffi as F\n"
        "L = F.lib(); sh = (C.c_int * 4)()\n"
        "n1 = L.sdr_rtc_selftest(31, 10, C.byref(sh)); c1 = sh[3]\n"
        "n2 = L.sdr_rtc_selftest(31, 10, C.byref(sh)); c2 = sh[3]\n"
        "bad = L.sdr_rtc_selftest(200, 3, C.byref(sh))\n"
        "print(n1, c1, n2, c2, bad, sh[0] if n1 > 0 else L.sdr_last_error().decode())\n" % str(ROOT))
    env = dict(os.environ, SDR_RTC_CACHE=str(tmp_path / "cubins"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    n1, c1, n2, c2, bad = (int(v) for v in r.stdout.split()[:5])
    assert n1 > 10_000 and n2 == n1, r.stdout          # a real cubin, identical bytes the second time
    assert c1 >= 1 and c2 == 0, r.stdout               # compiled once, then cached
    assert bad == -1                                   # SDR_E_ARG: 67 lags per sample is outside the kernel's range
    assert len(list((tmp_path / "cubins").glob("*.cubin"))) == c1


def test_rtc_shape_picker_only_picks_shapes_the_kernel_accepts(S):
    """Every (taps, decimation) the picker accepts must satisfy k_fir_fast's compile-time requirements (FastGeom's
    static_asserts, re-stated on the host by sdr_rtc_pick_shape): a violation would only show up as an NVRTC error when
    somebody asks for that shape."""
    from rtl_sdr_rs_b200 import _ffi as F
    L = F.lib()
    sh = (C.c_int * 4)()
    picked = outside = 0
    for D in range(1, 300):
        for T in sorted({1, 2, D - 1, D, D + 1, 2 * D - 1, 2 * D + 1, 3 * D, 5 * D + 3, 16 * D, 16 * D + 1, 127, 255, 1023, 4096, 4097}):
            if T < 1:
                continue
            rc = L.sdr_rtc_pick_shape(T, D, C.byref(sh))
            assert rc >= 0, (T, D, list(sh), L.sdr_last_error().decode())
            if rc:
                picked += 1
                B, NT, WB, PAD = list(sh)
                assert (B * D) % (WB // 2) == 0 and B * ((T + D - 1) // D) <= 16 and D <= 256
            else:
                outside += 1
    assert picked > 1500 and outside > 300, (picked, outside)


def test_rtc_compiles_the_output_owner_kernel_without_a_gpu():
    """k_fir_slide (csrc/fir_fast.cuh): the shapes the block-owner kernel cannot take compile through NVRTC with no device;
    shapes outside BOTH unrolled kernels are refused (sdr_fmrx_new() then runs k_fir_generic)."""
    import ctypes as C
    sdrpkg.load()
    from rtl_sdr_rs_b200 import _ffi as F
    L = F.lib()
    sh = (C.c_int * 2)()
    n = L.sdr_rtc_compile_slide(31, 2, sh)
    assert n > 10_000, L.sdr_last_error()
    assert sh[0] == 8 and sh[1] == 128          # 8 outputs per thread, 128 threads
    n = L.sdr_rtc_compile_slide(200, 3, sh)
    assert n > 10_000 and sh[0] == 3            # 768 / 200 outputs per thread
    assert L.sdr_rtc_compile_slide(1001, 250, sh) == -1
    assert L.sdr_rtc_compile_slide(300, 301, sh) == -1
