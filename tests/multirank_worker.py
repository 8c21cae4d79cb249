"""Worker of tests/test_multirank_cpu.py: one rank of a world_size-2 gloo job on CPU.

Exercises the N>1 HOST logic without a GPU: (a) the time-sliced stream (bench.py --gpus N headline) — every
rank cuts its slice with the C ABI's closed-form planner, the ORACLE stands in for the kernels (same arithmetic by the
parity tests), and the slices must tile the single-stream result exactly; (b) north_star's channel partition — the raw
slab exists on rank 0 only, ONE broadcast delivers it, rank r channelises the interleaved channel set c = r (mod world),
for which sdr_chan_bank_plan() must find the rank's own uniform grid, and the gathered channels equal a single-rank run."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_ffi as O  # noqa: E402
import sdrpkg  # noqa: E402
from sigutil import channel_taps, lowpass_taps  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = sdrpkg.load()
    T, D, up, down = 63, 20, 4, 5
    taps, taps2 = channel_taps(T, D), lowpass_taps(31, 0.45 / down, gain=up)
    cfg = S.FmrxConfig(T, D, taps2.size, up, down, 0.0)
    total = 20 * 5 * 997 + 13                       # not a multiple of anything convenient
    seed = 0xB2000001
    # slices are cut on multiples of lcm(D*down, 8) so that FIR and resampler phases restart cleanly
    align = int(np.lcm(D * down, 8))
    lo, hi = S.shard_range(total, world, rank, align)
    y0, ny, a0, na = S.fmrx_plan(cfg, lo, hi - lo)
    # every rank primes its chain with the samples just before its slice (halo), then runs its slice
    halo = min(lo, align * ((T + D + align - 1) // align + 8))
    iq = O.synth_fill(2 * (hi - lo + halo), seed, 2 * (lo - halo))
    chain = O.FxChain(taps, D, taps2, up, down)
    if halo:
        chain.process(iq[: 2 * halo])
    y, d, a = chain.process(iq[2 * halo:])
    assert y.shape[0] == ny and a.shape[0] == na, (rank, y.shape, ny, a.shape, na)
    # gather the slices on rank 0 and compare with the single-stream oracle
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(rank=rank, lo=lo, hi=hi, y0=y0, a0=a0, y=y, d=d, a=a))
    # integer Demod: per-rank buffer ranges of a batch tile the audio the same way (closed-form state)
    _, dcfg = S.optimal_settings()
    n_bufs, buf_len = 7, 4096
    b_lo, b_hi = S.shard_range(n_bufs, world, rank, 1)
    st = S.DemodState()
    if b_lo:
        _, _, st = S.demod_plan(dcfg, buf_len, b_lo)
    nl, na_i, _ = S.demod_plan(dcfg, buf_len, b_hi - b_lo, st)
    t = torch.tensor([nl, na_i], dtype=torch.int64)
    dist.all_reduce(t)
    # ---- north_star's partition: channels across ranks, ONE broadcast of the raw u8 slab, nothing else ---------------
    # 16 channels per rank on the interleaved grid (rank r owns c = r mod world: its own channels sit on a uniform grid, so
    # the library plans the polyphase bank for them); the slab is generated on rank 0 only and broadcast (gloo here, one
    # ncclBroadcast on the GPUs); every rank channelises its share (oracle standing in for the kernel).
    Cc, Tc, Dc, n_c = 16, 63, 20, 20 * 300 + 7
    cc_tot = Cc * world
    ctaps = channel_taps(Tc, Dc)

    def plan(r):
        offs = ((r + world * np.arange(Cc)) - (cc_tot - 1) / 2.0) / cc_tot
        return (np.round(offs * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)
    slab = torch.from_numpy(O.synth_fill(2 * n_c, seed + 7).copy()) if rank == 0 else torch.zeros(2 * n_c, dtype=torch.uint8)
    dist.broadcast(slab, src=0)
    raw = slab.numpy()
    bank = S.bank_plan(ctaps, Dc, plan(rank), want_tables=False)
    yc, dc = O.channelise(raw, ctaps, Dc, plan(rank))
    cgath = [None] * world
    dist.all_gather_object(cgath, dict(rank=rank, fw=plan(rank), y=yc, d=dc, K=None if bank is None else int(bank[0][0])))
    ok = True
    if rank == 0:
        # every channel of the full plan exactly once, each rank's set on a K = 16 grid, results equal to one rank doing all
        cgath.sort(key=lambda g: g["rank"])
        fw_full = (np.round(((np.arange(cc_tot) - (cc_tot - 1) / 2.0) / cc_tot) * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)
        got_fw = np.concatenate([g["fw"] for g in cgath])
        ok &= sorted(got_fw.tolist()) == sorted(fw_full.tolist())
        ok &= all(g["K"] == Cc for g in cgath)
        y_all, d_all = O.channelise(O.synth_fill(2 * n_c, seed + 7), ctaps, Dc, fw_full)
        for g in cgath:
            idx = g["rank"] + world * np.arange(Cc)
            ok &= bool(np.array_equal(g["y"], y_all[idx]) and np.array_equal(g["d"], d_all[idx]))
        whole = O.synth_fill(2 * total, seed)
        yw, dw, aw = O.FxChain(taps, D, taps2, up, down).process(whole)
        gathered.sort(key=lambda g: g["rank"])
        assert gathered[0]["lo"] == 0 and gathered[-1]["hi"] == total
        for g0, g1 in zip(gathered[:-1], gathered[1:]):
            assert g0["hi"] == g1["lo"]
        ycat = np.concatenate([g["y"] for g in gathered]); acat = np.concatenate([g["a"] for g in gathered])
        assert ycat.shape == yw.shape and acat.shape == aw.shape
        # y is an exact function of its window: slices agree to the last bit; d/a after the halo transient too
        ok &= bool(np.array_equal(ycat, yw))
        ok &= bool(np.allclose(acat, aw, rtol=0, atol=1e-9))
        assert [g["y0"] for g in gathered] == [g["lo"] // D for g in gathered]
        nl_w, na_w, _ = S.demod_plan(dcfg, buf_len, n_bufs)
        ok &= (int(t[0]) == nl_w and int(t[1]) == na_w)
        Path(os.environ["MULTIRANK_OUT"]).write_text(json.dumps({"ok": ok, "world": world, "slices": [(g["lo"], g["hi"]) for g in gathered]}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
