"""Shared test/bench helpers: tap design, the seeded synthetic FM parity signal, tolerances."""
from __future__ import annotations

import numpy as np


def lowpass_taps(n_taps: int, cutoff_cyc_per_sample: float, gain: float = 1.0) -> np.ndarray:
    """Hamming-windowed sinc designed in f64, rounded to f32 (the taps are DATA shared bit-for-bit by
    the oracle and the CUDA path)."""
    k = np.arange(n_taps, dtype=np.float64) - (n_taps - 1) / 2.0
    h = 2.0 * cutoff_cyc_per_sample * np.sinc(2.0 * cutoff_cyc_per_sample * k)
    h *= np.hamming(n_taps)
    h *= gain / h.sum()
    return h.astype(np.float32)


def channel_taps(n_taps: int, decim: int) -> np.ndarray:
    """The configs' channel filter: cutoff 0.4 * fs_out (SURVEY §8d cfg 2)."""
    return lowpass_taps(n_taps, 0.4 / decim)


def fm_test_signal(n: int, fs: float, seed: int = 0xB2000001, f_dev: float = 75e3, f_mod: float = 1e3,
                   f_c: float = 0.0, amp: float = 100.0, noise: float = 8.0) -> np.ndarray:
    """u8 interleaved IQ: 127.5 + amp*exp(j*phi[n]) + N(0, noise^2), phi' = 2*pi*(f_c + f_dev*sin(2*pi*f_mod*t))/fs."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    inst = f_c + f_dev * np.sin(2 * np.pi * f_mod * t)
    phi = 2 * np.pi * np.cumsum(inst) / fs
    i = 127.5 + amp * np.cos(phi) + rng.normal(0, noise, n)
    q = 127.5 + amp * np.sin(phi) + rng.normal(0, noise, n)
    out = np.empty(2 * n, np.uint8)
    out[0::2] = np.clip(np.rint(i), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(q), 0, 255).astype(np.uint8)
    return out


def saturated_stream(rng, n_bytes):
    """Bytes whose rotate_90 + (-127) image sits on the rails for long runs, so that boxcar sums reach +-128*D and the
    products of two of them leave i32 (D >= 256), mixed with noise runs and sign flips."""
    assert n_bytes % 8 == 0
    out = np.empty(n_bytes, np.uint8)
    pos = 0
    while pos < n_bytes:
        ln = min(n_bytes - pos, 8 * int(rng.integers(20, 900)))
        kind = rng.integers(0, 4)
        if kind == 0:
            out[pos:pos + ln] = rng.integers(0, 256, ln, dtype=np.uint8)
        else:
            # rotated group = [b0, b1, 255-b3, b2, 255-b4, 255-b5, b7, 255-b6]; pick rails (re_hi, im_hi) for the run
            re_hi, im_hi = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            R, I = (255 if re_hi else 0), (255 if im_hi else 0)
            grp = np.array([R, I, R, 255 - I, 255 - R, 255 - I, 255 - R, I], np.uint8)   # rotates to (R, I) x 4
            out[pos:pos + ln] = np.tile(grp, ln // 8)
            if kind == 3:   # sprinkle noise on the rails
                idx = rng.integers(0, ln, ln // 16)
                out[pos + idx] = rng.integers(0, 256, idx.size, dtype=np.uint8)
        pos += ln
    return out


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    """Norm-wise relative error max|got-ref| / max|ref|."""
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(got, np.float64) - ref)) / max(np.max(np.abs(ref)), 1e-300))


def assert_close(got, ref, rtol=1e-5, what=""):
    """The f32-path bar of BASELINE.json's north_star: within 1e-5 relative.  Element-wise
    |got-ref| <= rtol*|ref| + rtol*rms(ref) (the rms term covers elements that cancel to ~0)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    rms = float(np.sqrt(np.mean(ref * ref)))
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + rtol * rms
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{ref.size} outside 1e-5 rel; worst err {err.max():.3e} (rms {rms:.3e})"


def assert_angle_close(got, ref, full_scale: float, rtol=1e-5, what=""):
    """Discriminator outputs live on a circle of circumference 2*full_scale (full_scale == gain*pi)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    d = got - ref
    d = (d + full_scale) % (2 * full_scale) - full_scale
    tol = rtol * np.abs(ref) + rtol * full_scale
    bad = np.abs(d) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{ref.size} outside 1e-5 rel; worst {np.abs(d).max():.3e} of {full_scale}"


def disc_f64(y_pairs, gain, prev=(0.0, 0.0)):
    """f64 polar discriminator of a complex stream given as (n,2) pairs (same definition as the oracle)."""
    y = np.asarray(y_pairs, np.float64)
    z = y[:, 0] + 1j * y[:, 1]
    zp = np.concatenate([[complex(*prev)], z[:-1]])
    c = z * np.conj(zp)
    out = gain * np.arctan2(c.imag, c.real)
    out[(c.real == 0) & (c.imag == 0)] = 0.0
    return out


def assert_demod_propagated(d_got, y_ref_pairs, d_ref, gain, rtol=1e-5, what=""):
    """End-to-end discriminator check.  If y is within rtol (the f32 FIR bar), the angle between two
    successive samples can move by at most dy[m]/|y[m]| + dy[m-1]/|y[m-1]| rad (first order); allow that
    plus rtol of full scale.  Differences are taken on the circle."""
    y = np.asarray(y_ref_pairs, np.float64)
    mag = np.hypot(y[:, 0], y[:, 1])
    rms = float(np.sqrt(np.mean(mag * mag))) if mag.size else 0.0
    dy = rtol * mag + rtol * rms
    rel = dy / np.maximum(mag, 1e-300)
    relp = np.concatenate([[np.inf], rel[:-1]])
    full = gain * np.pi
    tol = gain * np.minimum(rel + relp, np.pi) + rtol * full
    d = np.asarray(d_got, np.float64) - np.asarray(d_ref, np.float64)
    d = (d + full) % (2 * full) - full
    bad = np.abs(d) > tol
    bad[0] = False if mag.size and not np.isfinite(relp[0]) else bad[0]
    assert not bad.any(), f"{what}: {int(bad.sum())}/{d.size} beyond the propagated 1e-5 bound; worst {np.abs(d).max():.3e}"
