"""Worker of tests/test_multirank_gpu.py: one rank (= one GPU) of a torchrun job.

(1) time-sliced single-channel receiver: rank r owns slice r of ONE stream (sdr_fmrx_seek + halo warm-up);
    the concatenated per-rank outputs must equal a single-GPU run of the whole stream BIT FOR BIT.
(2) channel-sharded channeliser: rank 0 generates the raw u8 slab, ONE ncclBroadcast (sdr_comm_bcast_u8)
    delivers it to every rank, rank r channelises channels [8r, 8r+8) with the direct-form kernel; the gathered
    result must equal a single-GPU run of all channels bit for bit (the direct form treats channels independently).
(3) the bench's plan: 64 channels per rank on the INTERLEAVED grid (rank r owns c = r mod world), which every rank runs
    through the two-stage polyphase bank.  The bank's coefficient tables belong to a handle's own channel set, so the
    reference is a single-GPU handle with the SAME set: bit for bit; and against the direct form: within 3e-6."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import sdrpkg  # noqa: E402
from sigutil import channel_taps, lowpass_taps  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    S = sdrpkg.load()
    seed = 0xB2000001
    ok = True

    # ---- (1) time slices of one stream ------------------------------------------------------------
    T, D, up, down = 127, 75, 1, 1
    taps, taps2 = channel_taps(T, D), lowpass_taps(63, 0.45)
    cfg = S.FmrxConfig(T, D, taps2.size, up, down, 0.0)
    total = 75 * 8 * 40_000 + 8 * 123
    align = int(np.lcm(D * down, 8))
    lo, hi = S.shard_range(total, world, rank, align)
    halo = min(lo, align * 12)                     # 96 FIR outputs: covers the 127-tap FIR transient and the 62
                                                   # discriminator values the 63-tap audio FIR looks back on
    d_iq = S.DevBuffer(2 * (hi - lo + halo), local)
    S.synth_fill_dev(d_iq, 2 * (hi - lo + halo), seed, byte_offset=2 * (lo - halo))
    rx = S.FmRx(taps, D, taps2, up, down, device=local)
    rx.seek(lo - halo)
    _, ny, _, na = S.fmrx_plan(cfg, lo, hi - lo)
    d_a, d_d = S.DevBuffer(4 * (na + 64), local), S.DevBuffer(4 * (ny + 64), local)
    if halo:
        rx.process_dev(d_iq, halo, d_a, na + 64, d_demod=d_d)
    n_a = rx.process_dev(d_iq, hi - lo, d_a, na + 64, d_demod=d_d, iq_offset=2 * halo)
    rx.sync()
    assert n_a == na
    mine = dict(rank=rank, lo=lo, hi=hi, audio=d_a.download(np.float32, na), demod=d_d.download(np.float32, ny))
    parts = [None] * world
    dist.all_gather_object(parts, mine)

    # ---- (2) channel shards fed by one NCCL broadcast ----------------------------------------------
    C_per, Tc, Dc, n_c = 8, 63, 20, 20 * 4096
    c_tot = C_per * world
    fw_all = (np.round(((np.arange(c_tot) - (c_tot - 1) / 2) / c_tot) * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)
    ctaps = channel_taps(Tc, Dc)
    slab = S.DevBuffer(2 * n_c, local)
    if rank == 0:
        S.synth_fill_dev(slab, 2 * n_c, seed + 7)
    uid = [S.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = S.Comm(local, rank, world, uid[0])
    os.environ["SDR_CHAN_BANK"] = "0"      # (2) is the direct form on every handle; read when a handle is created
    ch = S.Channeliser(ctaps, Dc, fw_all[rank * C_per:(rank + 1) * C_per], device=local)
    assert ch.kernel_kind()[0] == 1
    comm.bcast_u8(slab, 2 * n_c, 0)
    comm.chan_wait(ch)
    cap = n_c // Dc
    d_cd = S.DevBuffer(4 * C_per * cap, local)
    m = ch.process_dev(slab, n_c, d_cd, cap)
    ch.sync(); comm.sync()
    cparts = [None] * world
    dist.all_gather_object(cparts, dict(rank=rank, d=d_cd.download(np.float32, C_per * cap).reshape(C_per, cap)[:, :m]))

    # ---- (3) the interleaved 64-per-rank plan through the polyphase bank ----------------------------------
    del os.environ["SDR_CHAN_BANK"]
    Cb, Tb, Db = 64, 255, 100
    cb_tot = Cb * world
    n_b = Db * 800                                      # (the slab of (2) holds 81920 samples)
    btaps = channel_taps(Tb, Db)

    def plan(r):
        offs = ((r + world * np.arange(Cb)) - (cb_tot - 1) / 2.0) / cb_tot
        return (np.round(offs * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)
    chb = S.Channeliser(btaps, Db, plan(rank), device=local)
    assert chb.kernel_kind()[0] == 2 and chb.kernel_kind()[1][0] == 64, chb.kernel_kind()
    capb = n_b // Db
    d_bd = S.DevBuffer(4 * Cb * capb, local)
    mb = chb.process_dev(slab, n_b, d_bd, capb)          # the broadcast slab of (2), first n_b samples
    chb.sync()
    bparts = [None] * world
    dist.all_gather_object(bparts, dict(rank=rank, d=d_bd.download(np.float32, Cb * capb).reshape(Cb, capb)[:, :mb]))

    if rank == 0:
        whole = S.DevBuffer(2 * total, 0)
        S.synth_fill_dev(whole, 2 * total, seed)
        rx1 = S.FmRx(taps, D, taps2, up, down, device=0)
        ny1, na1 = rx1.out_lens(total)
        a1, d1 = S.DevBuffer(4 * na1, 0), S.DevBuffer(4 * ny1, 0)
        rx1.process_dev(whole, total, a1, na1, d_demod=d1); rx1.sync()
        parts.sort(key=lambda p: p["rank"])
        checks = {}
        checks["slices_demod"] = bool(np.array_equal(np.concatenate([p["demod"] for p in parts]), d1.download(np.float32, ny1)))
        checks["slices_audio"] = bool(np.array_equal(np.concatenate([p["audio"] for p in parts]), a1.download(np.float32, na1)))
        os.environ["SDR_CHAN_BANK"] = "0"      # the reference of (2) is the direct form as well
        ch1 = S.Channeliser(ctaps, Dc, fw_all, device=0)
        del os.environ["SDR_CHAN_BANK"]
        assert ch1.kernel_kind()[0] == 1
        d_all = S.DevBuffer(4 * c_tot * cap, 0)
        m1 = ch1.process_dev(slab, n_c, d_all, cap); ch1.sync()
        ref = d_all.download(np.float32, c_tot * cap).reshape(c_tot, cap)[:, :m1]
        cparts.sort(key=lambda p: p["rank"])
        checks["direct_shards"] = bool(np.array_equal(np.concatenate([p["d"] for p in cparts], axis=0), ref))
        bparts.sort(key=lambda p: p["rank"])
        gain = 16384.0
        for r in range(world):
            same = S.Channeliser(btaps, Db, plan(r), device=0)
            d_same = S.DevBuffer(4 * Cb * capb, 0)
            ms = same.process_dev(slab, n_b, d_same, capb); same.sync()
            checks[f"bank_rank{r}_same_handle"] = bool(np.array_equal(d_same.download(np.float32, Cb * capb).reshape(Cb, capb)[:, :ms], bparts[r]["d"]))
            os.environ["SDR_CHAN_BANK"] = "0"
            direct = S.Channeliser(btaps, Db, plan(r), device=0)
            del os.environ["SDR_CHAN_BANK"]
            d_dir = S.DevBuffer(4 * Cb * capb, 0)
            direct.process_dev(slab, n_b, d_dir, capb); direct.sync()
            dd = d_dir.download(np.float32, Cb * capb).reshape(Cb, capb)[:, :ms] - bparts[r]["d"]
            dd = (dd + gain) % (2 * gain) - gain              # on the circle (full scale = gain * pi / pi)
            # uniform random bytes: every channel carries noise of comparable power, the discriminator is well conditioned
            # almost everywhere; the few near-zero crossings are excluded by the median
            checks[f"bank_rank{r}_vs_direct"] = bool(np.median(np.abs(dd)) < 3e-6 * gain * np.pi)
            for b in (d_same, d_dir):
                b.free()
        ok = all(checks.values())
        if not ok:
            print("FAILED CHECKS:", {k: v for k, v in checks.items() if not v}, file=sys.stderr, flush=True)
        Path(os.environ["MULTIRANK_OUT"]).write_text(json.dumps({"ok": bool(ok), "world": world}))
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
