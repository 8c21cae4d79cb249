"""world_size-2 gloo test (CPU) of the N>1 host logic: per-rank time slices cut by the C ABI's closed-form
planner tile one stream exactly (FIR outputs bit-identical, audio identical after the halo)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import sdrpkg

ROOT = Path(__file__).resolve().parent.parent


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_time_slices_tile_the_stream(tmp_path):
    out = tmp_path / "result.json"
    env = dict(os.environ, MULTIRANK_OUT=str(out), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "tests" / "multirank_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["world"] == 2 and len(res["slices"]) == 2


def test_planner_matches_sequential_bookkeeping():
    S = sdrpkg.load()
    rng = np.random.default_rng(0)
    for T, D, up, down in ((127, 75, 1, 1), (255, 100, 4, 25), (31, 7, 3, 2)):
        cfg = S.FmrxConfig(T, D, 16, up, down, 0.0)
        pos = 0
        ny_tot = na_tot = 0
        for n in rng.integers(0, 5000, 40):
            y0, ny, a0, na = S.fmrx_plan(cfg, pos, int(n))
            assert y0 == ny_tot and a0 == na_tot
            ny_tot += ny; na_tot += na; pos += int(n)
        assert ny_tot == pos // D and na_tot == -(-(pos // D) * up // down)
    import oracle_ffi as O
    _, ocfg = O.optimal_settings()
    _, cfg = S.optimal_settings()
    o, st = O.Demod(ocfg), S.DemodState()
    for nb, bl in ((1, 262144), (3, 4096), (2, 40), (5, 8 * 333)):
        buf = rng.integers(0, 256, nb * bl, dtype=np.uint8)
        n_audio = sum(o.demodulate(buf[i * bl:(i + 1) * bl]).size for i in range(nb))
        nl, na, st = S.demod_plan(cfg, bl, nb, st)
        assert na == n_audio
        assert (st.prev_index, st.prev_lpr_index) == (o.state()["prev_index"], o.state()["prev_lpr_index"])
    lo, hi = zip(*(S.shard_range(1003, 4, r, 8) for r in range(4)))
    assert lo[0] == 0 and hi[-1] == 1003 and list(hi[:-1]) == list(lo[1:]) and all(x % 8 == 0 for x in lo)


def test_demod_plan_matches_the_oracle_for_random_configs_and_states():
    """sdr_demod_plan (closed-form counts and index state, pure host arithmetic) against the oracle's sequential loops:
    random downsample / rate ratio / starting state / call geometry."""
    import oracle_ffi as O
    S = sdrpkg.load()
    rng = np.random.default_rng(20261017)
    for _ in range(60):
        D = int(rng.integers(1, 40))
        fast = int(rng.integers(1000, 400_000))
        slow = int(rng.integers(1, fast + 1))
        ocfg = O.DemodConfig(fast * D, fast, slow, D, 42)
        cfg = S.DemodConfig(fast * D, fast, slow, D, 42)
        o = O.Demod(ocfg)
        p0, q0 = int(rng.integers(0, D)), int(rng.integers(0, fast))
        o.set_state(prev_index=p0, prev_lpr_index=q0)
        st = S.DemodState(p0, 0, q0, 0, 0, 0, 0)
        min_len = ((2 * D * 2 + 7) // 8 + 1) * 8
        bl, nb = min_len + 8 * int(rng.integers(0, 600)), int(rng.integers(1, 6))
        buf = rng.integers(0, 256, nb * bl, dtype=np.uint8)
        n_audio = sum(o.demodulate(buf[i * bl:(i + 1) * bl]).size for i in range(nb))
        nl, na, after = S.demod_plan(cfg, bl, nb, st)
        assert na == n_audio, (D, fast, slow, p0, q0, bl, nb)
        assert nl == (p0 + nb * (bl // 2)) // D
        assert (after.prev_index, after.prev_lpr_index) == (o.state()["prev_index"], o.state()["prev_lpr_index"])
