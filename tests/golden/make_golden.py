#!/usr/bin/env python3
"""Generate the committed golden fixtures.  Run in the DEV container only (it reads
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes:
  tests/golden/kat_simple_fm.json   the three known-answer vectors of the reference's own
                                    test module (examples/simple_fm.rs:466-555), extracted
                                    textually from the Rust source (data, not code).
  tests/golden/capture_pins.json    sha256 pins of capture.bin and of the oracle's streams for
                                    it (75 calls x 262144 B, examples/simple_fm.rs:65-84), first/last
                                    samples, per-call audio lengths and Demod state after call 0.
  tests/golden/capture_head.bin     first 4 buffers (1 MiB) of capture.bin: the committed fixture.
  tests/golden/capture_head_audio.s16le   oracle audio for those 4 calls.
  tests/golden/_ref/capture.bin     full fixture copy (git-ignored; travels to the GPU box).
  tests/golden/_ref/capture_audio.s16le   full oracle audio (git-ignored).
The oracle is only trusted for the capture pins AFTER it reproduces the three KATs
(asserted below, and again in tests/test_oracle.py).
"""
import hashlib
import json
import re
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_ffi as O  # noqa: E402

REF = Path("/root/reference")
SRC = REF / "examples" / "simple_fm.rs"
BUF = O.DEFAULT_BUF_LENGTH


def vec_literals(text: str):
    """All `vec![ ... ]` integer literals inside the #[cfg(test)] module, in source order."""
    test_mod = text[text.index("#[cfg(test)]"):]
    out = []
    for m in re.finditer(r"let\s+(\w+)\s*=\s*vec!\[([^\]]*)\]", test_mod):
        nums = [int(t) for t in re.findall(r"-?\d+", m.group(2))]
        line = text[: text.index("#[cfg(test)]") + m.start()].count("\n") + 1
        out.append((m.group(1), line, nums))
    return out


def main():
    text = SRC.read_text()
    lits = vec_literals(text)
    byname = {}
    for name, line, nums in lits:
        byname.setdefault(name, []).append((line, nums))
    kat = {
        "source": "examples/simple_fm.rs #[cfg(test)] mod tests (:461-556), reference @ 8c32c118 v0.3.1",
        "provenance": "rtl_fm -f 92.5M -M fm -s 170k -A fast -r 32k -l 0 (comment at :467-468)",
        "test_lowpass": {
            "lines": "466-511",
            "buf_signed": byname["buf_signed"][0][1],
            "lowpass_expected": byname["lowpass"][0][1],
        },
        "test_demod": {
            "lines": "514-538",
            "lowpass": byname["lowpass"][1][1],
            "demod_expected": byname["demod_expected"][0][1],
        },
        "test_lowpass_real": {
            "lines": "541-555",
            "demodulated": byname["demodulated"][0][1],
            "result": byname["result"][-1][1],
        },
    }
    assert len(kat["test_lowpass"]["buf_signed"]) == 512
    assert len(kat["test_lowpass"]["lowpass_expected"]) == 84
    assert len(kat["test_demod"]["demod_expected"]) == 42
    assert kat["test_lowpass_real"]["result"] == [2588, 4030, -1212, -3430, 2585, 2110, -6110]
    (HERE / "kat_simple_fm.json").write_text(json.dumps(kat, indent=1) + "\n")

    # --- the oracle must pass the KATs before it is allowed to define the capture golden
    d = O.Demod()
    lp = d.low_pass_complex(O.buf_to_complex(np.array(kat["test_lowpass"]["buf_signed"], np.int16)))
    assert lp.reshape(-1).tolist() == kat["test_lowpass"]["lowpass_expected"]
    d = O.Demod()
    dm = d.fm_demod(np.array(kat["test_demod"]["lowpass"], np.int32).reshape(-1, 2))
    assert dm.tolist() == kat["test_demod"]["demod_expected"]
    d = O.Demod()
    au = d.low_pass_real(np.array(kat["test_lowpass_real"]["demodulated"], np.int16))
    assert au.tolist() == kat["test_lowpass_real"]["result"]

    cap = np.fromfile(REF / "capture.bin", dtype=np.uint8)
    assert cap.size % BUF == 0
    n_calls = cap.size // BUF
    d = O.Demod()
    audio, lps, dms, lens, lp_lens = [], [], [], [], []
    state0 = None
    for c in range(n_calls):
        a, lp, dm = d.demodulate(cap[c * BUF:(c + 1) * BUF], stages=True)
        audio.append(a), lps.append(lp), dms.append(dm), lens.append(int(a.size)), lp_lens.append(int(lp.shape[0]))
        if c == 0:
            state0 = d.state()
    audio, lps, dms = np.concatenate(audio), np.concatenate(lps), np.concatenate(dms)
    # reference-structured variant must agree bit for bit
    d2 = O.Demod()
    audio2 = np.concatenate([d2.demodulate(cap[c * BUF:(c + 1) * BUF], ref_like=True) for c in range(n_calls)])
    assert np.array_equal(audio, audio2)
    # chunking is part of the contract: one big call differs
    one = O.Demod().demodulate(cap)
    pins = {
        "capture_sha256": hashlib.sha256(cap.tobytes()).hexdigest(),
        "capture_bytes": int(cap.size),
        "n_calls": int(n_calls),
        "buf_len": BUF,
        "lowpassed_sha256": hashlib.sha256(lps.astype("<i4").tobytes()).hexdigest(),
        "lowpassed_count": int(lps.shape[0]),
        "demod_sha256": hashlib.sha256(dms.astype("<i2").tobytes()).hexdigest(),
        "audio_sha256": hashlib.sha256(audio.astype("<i2").tobytes()).hexdigest(),
        "audio_count": int(audio.size),
        "audio_first8": audio[:8].tolist(),
        "audio_last8": audio[-8:].tolist(),
        "audio_min_max": [int(audio.min()), int(audio.max())],
        "audio_lens_per_call": lens,
        "lowpassed_first4": lps[:4].tolist(),
        "demod_first8": dms[:8].tolist(),
        "demod_call1_first": int(dms[lp_lens[0]]),
        "lowpassed_lens_per_call": lp_lens,
        "state_after_call0": state0,
        "single_call_audio_diffs": int(np.count_nonzero(one[: audio.size] != audio)) if one.size == audio.size else -1,
        "head_calls": 4,
        "head_audio_sha256": hashlib.sha256(audio[: sum(lens[:4])].astype("<i2").tobytes()).hexdigest(),
    }
    (HERE / "capture_pins.json").write_text(json.dumps(pins, indent=1) + "\n")
    cap[: 4 * BUF].tofile(HERE / "capture_head.bin")
    audio[: sum(lens[:4])].astype("<i2").tofile(HERE / "capture_head_audio.s16le")
    ref = HERE / "_ref"
    ref.mkdir(exist_ok=True)
    shutil.copyfile(REF / "capture.bin", ref / "capture.bin")
    audio.astype("<i2").tofile(ref / "capture_audio.s16le")
    print(json.dumps({k: pins[k] for k in ("capture_sha256", "lowpassed_sha256", "demod_sha256", "audio_sha256",
                                            "audio_count", "single_call_audio_diffs", "state_after_call0")}, indent=1))


if __name__ == "__main__":
    main()
