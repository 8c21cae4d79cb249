"""CPU check of the two-stage polyphase bank's host side (no GPU): sdr_chan_bank_plan() returns the plan and the very
coefficient blobs k_chan_bank would receive; a numpy model that consumes those blobs exactly the way the kernel does
(csrc/chan_bank.cuh: stage-1 sub-filters per residue r1, stage-2 combination per channel, phi-shifted discriminator) is
compared with the oracle's DIRECT definition (mix by the 32-bit NCO, FIR, decimate, discriminate).  This pins the grid
detection, the K = K1*K2 split and every table formula before the kernel ever runs."""
import numpy as np
import pytest

import oracle_ffi as O
import sdrpkg
from sigutil import assert_close, channel_taps, rel_err

GAIN = 16384.0 / np.pi
CH = 64


def words(offsets_hz, fs):
    return (np.round(np.asarray(offsets_hz, np.float64) / fs * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)


def bank_model(iq, taps, D, fw, plan, tabs):
    """S_c[m], y_c[m], d_c[m] from the blobs, in f64, with the kernel's loop structure."""
    K, K1, K2, groups = plan
    T, C = taps.size, fw.size
    x = (iq[0::2].astype(np.float64) - 127.0) + 1j * (iq[1::2].astype(np.float64) - 127.0)
    n = x.size
    M = n // D
    PAD = T + 1100
    xp = np.concatenate([np.zeros(PAD, complex), x])          # x[n < 0] = 127 -> centred 0
    y = np.zeros((C, M), complex)
    d = np.zeros((C, M))
    for g in range(groups):
        v = tabs[g, :, 0].astype(np.float64) + 1j * tabs[g, :, 1].astype(np.float64)
        S = np.zeros((CH, M), complex)
        NJ = -(-T // K1)
        Tp = K1 * NJ                                          # zero-padded taps
        eoff = (Tp * K2 + 1) & ~1
        gidx, eidx = 0, eoff
        for r1 in range(K1):
            A = np.zeros((K2, M), complex)
            for j in range(NJ):
                k = r1 + j * K1
                xs = xp[PAD + (np.arange(M) + 1) * D - 1 - k]  # x[n_m - k]
                for b in range(K2):
                    A[b] += v[gidx + b] * xs
                gidx += K2
            for c in range(CH):
                # E is stored per channel pair: entry (re c, re c+1), then entry (im c, im c+1)
                pair = tabs[g, eidx + (c & ~1):eidx + (c & ~1) + 2].astype(np.float64)
                S[c] += (pair[0, c & 1] + 1j * pair[1, c & 1]) * A[c % K2]
            eidx += CH
        phi = tabs[g, eoff + K1 * CH:eoff + K1 * CH + CH, 0].astype(np.float64)
        nm = (np.arange(M, dtype=np.uint64) + 1) * D - 1
        for c in range(CH):
            ch = g * CH + c
            if ch >= C:
                break
            ph = ((np.uint64(fw[ch]) * nm) % np.uint64(1 << 32)).astype(np.float64) / 2.0 ** 32
            y[ch] = S[c] * np.exp(-2j * np.pi * ph)
            z = S[c] * np.conj(np.concatenate([[0.0], S[c][:-1]]))
            t = np.arctan2(z.imag, z.real) - phi[c]
            t = np.where(t > np.pi, t - 2 * np.pi, np.where(t <= -np.pi, t + 2 * np.pi, t))
            d[ch] = np.where(z == 0, 0.0, GAIN * t)
    return y, d


CASES = [  # (name, C, T, D, offsets builder, expected K)
    ("cfg4: (c - 31.5) * 200 kHz at 20 Msps", 64, 255, 100, lambda C, fs: (np.arange(C) - 31.5) * 200e3, 100),
    ("cfg5 interleaved: rank 3 of 8 owns c = 3 (mod 8) of 512", 64, 255, 100,
     lambda C, fs: ((3 + 8 * np.arange(C)) - 255.5) * (fs / 512), 64),
    ("two groups: 100 channels on a 128-bin grid", 100, 63, 20, lambda C, fs: (np.arange(C) - 50) * (fs / 128), 128),
    ("short filter, odd decimation, K = 16 < C = 64 (bins alias)", 64, 31, 7, lambda C, fs: np.arange(C) * (fs / 16), 16),
    ("prime K = 53: stage 1 is the plain filter, K1 = 53", 64, 255, 50, lambda C, fs: (np.arange(C) - 20) * (fs / 53), 53),
]


@pytest.mark.parametrize("name,C,T,D,offs,K", CASES)
def test_bank_tables_reproduce_the_direct_definition(name, C, T, D, offs, K):
    S = sdrpkg.load()
    fs = 20e6
    taps = channel_taps(T, D)
    fw = words(offs(C, fs), fs)
    got = S.bank_plan(taps, D, fw)
    assert got is not None, name
    plan, tabs = got
    assert plan[0] == K and plan[1] * plan[2] == K and plan[3] == -(-C // CH), plan
    n = D * 40 + 7
    iq = np.random.default_rng(C + T).integers(0, 256, 2 * n, dtype=np.uint8)
    y, d = bank_model(iq, taps, D, fw, plan, tabs)
    yo, do = O.channelise(iq, taps, D, fw)
    yoc = yo[..., 0] + 1j * yo[..., 1]
    # f32 tables + the grid approximation of the rounded NCO words: far inside the 1e-5 bar
    assert rel_err(np.stack([y.real, y.imag], -1), yo) < 2e-6, rel_err(np.stack([y.real, y.imag], -1), yo)
    for c in range(C):
        assert_close(np.stack([y[c].real, y[c].imag], -1), yo[c], what=f"{name} ch {c}")
    # discriminator: compare on the circle where |y| is not tiny
    strong = (np.abs(yoc) > 1e-3 * np.abs(yoc).max())
    strong[:, 1:] &= strong[:, :-1]
    dd = (d - do + GAIN * np.pi) % (2 * GAIN * np.pi) - GAIN * np.pi
    assert np.abs(dd[strong]).max() < 1e-4 * GAIN * np.pi, np.abs(dd[strong]).max()


def test_non_uniform_or_unprofitable_plans_are_rejected():
    S = sdrpkg.load()
    fs, T, D = 20e6, 255, 100
    taps = channel_taps(T, D)
    assert S.bank_plan(taps, D, words((np.arange(64) - 31.5) * (fs / 64) * 0.9, fs)) is None      # 9 bins per 10 channels
    assert S.bank_plan(taps, D, words(np.sort(np.random.default_rng(1).uniform(-9e6, 9e6, 64)), fs)) is None
    assert S.bank_plan(taps, D, words((np.arange(4) - 1.5) * 200e3, fs)) is None                    # too few channels
    assert S.bank_plan(taps, D, words((31.5 - np.arange(64)) * 200e3, fs)) is None                  # descending grid
    fw = words((np.arange(64) - 31.5) * 200e3, fs).copy()
    fw[17] += 40                                                                                  # one channel off the grid
    assert S.bank_plan(taps, D, fw) is None
    # contiguous 64 of 512 bins: the tables would not fit the parameter blob -> direct form
    assert S.bank_plan(taps, D, words((np.arange(192, 256) - 255.5) * (fs / 512), fs)) is None
