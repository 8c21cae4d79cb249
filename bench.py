#!/usr/bin/env python3
"""bench.py — headline benchmark of the IQ-sample DSP hot path (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg1|chan] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic IQ that is ALREADY RESIDENT in HBM.

  cfg2 (default; BASELINE.json configs[1]): single-channel 2.4 Msps-equivalent synthetic IQ, 127-tap
        FIR decimate-by-75, fused convert+FIR+polar-discriminator kernel, 63-tap audio FIR at 32 kHz;
        2^30 complex samples (2 GiB of u8) per GPU per step.
  cfg3 (configs[2]): 20 Msps-equivalent, 255-tap FIR /100, FM demod, 200k->32k (4/25) resampler.
  cfg1 (configs[0] semantics at scale): the reference-exact integer Demod (boxcar-6, fast_atan2,
        170k->32k) over 8192 x 262144-byte buffers per step.
  chan (configs[3]/[4]): the 64-channel-per-GPU wideband channeliser.

The ONE JSON line rank 0 prints carries the headline workload in `value` / `roofline` / `e2e` / `cpu_baseline`, and
— so that every BASELINE.json config is measured by the same driver-run command —
  `extra`     : cfg1 {ms, GB/s, frac, e2e pinned + pageable, per_buffer, CPU legs incl. the single-thread ref_like}, the
                other f32 config {ms, frac}, the channeliser {ms per slab, roofline}  (N = 1; measured after the timed region);
  `multi_gpu` : (N > 1, under torchrun) north_star's partition — 64 channels per rank, the raw u8 slabs broadcast from
                rank 0 with one ncclBroadcast each: {channels_total, input_msamples_per_s, channel_msamples_per_s,
                bcast_gbs, exposed_bcast_ms, ...} and the box's bare N-rank H2D ceiling next to the e2e figure.
The headline workload at N > 1: every rank owns one time slice of the same synthetic stream (weak scaling, no data-path
collective).  `--impl reference` times the CPU restatement of the same workload (oracle port; the reference itself is
Rust and cannot be built in this image) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SEED = 0xB2000001
METRIC = "IQ Msamples/s through fused convert+FIR+demod kernel; achieved HBM GB/s vs B200 peak"


def workload_spec(name: str) -> dict:
    if name == "cfg2":
        return dict(name="cfg2", T=127, D=75, T2=63, up=1, down=1, fs=2.4e6, n=1 << 30, dtype="f32",
                    desc="single-channel 2.4 Msps-equivalent synthetic IQ, 127-tap FIR decimate-by-75, fused "
                         "convert+FIR+FM-demod kernel, 63-tap audio FIR @32 kHz (BASELINE.json configs[1])")
    if name == "cfg3":
        return dict(name="cfg3", T=255, D=100, T2=127, up=4, down=25, fs=20e6, n=1 << 30, dtype="f32",
                    desc="single-channel 20 Msps-equivalent synthetic IQ, 255-tap FIR decimate-by-100, fused "
                         "convert+FIR+FM-demod kernel, 200k->32k (4/25) 127-tap resampler (BASELINE.json configs[2])")
    if name == "cfg1":
        return dict(name="cfg1", buf_len=262144, n_bufs=8192, n=8192 * 131072, dtype="i32",
                    desc="reference-exact integer Demod (rotate_90, -127, boxcar-6, fast_atan2, 170k->32k), "
                         "8192 x 262144-byte buffers per step (BASELINE.json configs[0] semantics at HBM scale)")
    if name == "chan":
        return dict(name="chan", T=255, D=100, C=64, fs=20e6, n=1 << 28, slab=1 << 25, dtype="f32",
                    desc="wideband channeliser, 64 channels per GPU (per-channel 32-bit NCO folded into 255-tap complex "
                         "taps, decimate-by-100, FM demod), 20 Msps-equivalent synthetic IQ; N>1: 64*N channels, raw u8 "
                         "slabs (64 MiB) broadcast from rank 0 with one ncclBroadcast each (BASELINE.json configs[3]/[4])")
    raise SystemExit(f"unknown workload {name}")


def config_of(w: dict, info: dict | None = None) -> dict:
    """The `config` object of a JSON line: identical keys for the GPU arm and the reference arm."""
    c = {"workload": w["desc"], "samples_per_gpu_per_step": w["n"], "input_bytes_per_gpu": 2 * w["n"],
         "l2_policy": "input (2 GiB) is larger than L2 (126 MB); no flush needed",
         "sharding": "each rank owns one time slice of the same seeded stream; no data-path collective"}
    if info is not None:
        c["device"], c["sm_count"] = info["name"], info["sm_count"]
    return c


def taps_for(w: dict):
    from sigutil import channel_taps, lowpass_taps
    taps = channel_taps(w["T"], w["D"])
    taps2 = lowpass_taps(w["T2"], 0.45 / max(w["up"], w["down"]), gain=w["up"])
    return taps, taps2


def alg_bytes_per_sample(w: dict) -> float:
    """Algorithmic HBM bytes of the DOMINANT kernel per complex input sample."""
    if w["name"] == "cfg1":
        return 2.0 + 2.0 * (32000 / 170000) / 6        # u8 IQ in, i16 audio out
    return 2.0 + 4.0 / w["D"]                          # u8 IQ in, f32 discriminator out


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def n_samples(self) -> int:
        try:
            return sum(1 for line in Path(self.path).read_text().splitlines() if line.count(",") >= 8)
        except Exception:
            return 0

    def wait_first(self, timeout: float = 5.0) -> None:
        """Block until nvidia-smi has printed its first sample: its start-up (NVML initialisation over every GPU of the
        box) takes driver locks and stalls kernel launches for milliseconds — it must be over BEFORE the timed region."""
        t0 = time.perf_counter()
        while self.proc is not None and self.n_samples() < 1 and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class Ctx:
    """Rank / device / torch.distributed plumbing shared by every leg."""

    def __init__(self):
        self.rank, self.local_rank, self.world = dist_env()
        self.dist = None
        if self.world > 1:
            # NCCL's broadcast kernel shares the SMs with the FMA-bound channeliser (multi_gpu leg): 16 CTAs move a 64 MiB
            # slab at 320 GB/s (8 ranks) to 520 GB/s (2 ranks), about the channeliser's own pace, where the default 32 CTAs
            # cost it more than they gain (ms per 2^28-sample step at N = 2: 32 CTAs 1.87, 16: 1.76, 8: 2.34; at N = 8
            # exposed broadcast time 0.92 ms with 32 CTAs, 0.23 ms with 16).  NCCL reads the variable once per process, at its
            # first initialisation — so it is set here, before torch creates its communicator.
            os.environ.setdefault("NCCL_MAX_CTAS", "16")
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl")
            self.dist, self.torch = dist, torch
        self.device = self.local_rank
        import sdrpkg
        self.S = sdrpkg.load()
        if self.S.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        self.info = self.S.device_info(self.device)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier(device_ids=[self.local_rank])

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, val: int) -> int:
        if self.dist is None:
            return int(val)
        t = self.torch.tensor([val], dtype=self.torch.int64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t[0])

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------
def measure_stream(cx: Ctx, w: dict, d_in, steps: int, warmup: int, sample_clocks: bool = False) -> dict:
    """cfg1 / cfg2 / cfg3 with the input resident in HBM: W warm-up steps, then exactly K steps between a barrier +
    stream sync on both sides, device time from CUDA events on the handle's stream, max over ranks."""
    S, n = cx.S, w["n"]
    if w["name"] == "cfg1":
        h = S.Demod(device=cx.device)
        out_cap = (h.out_len(w["buf_len"]) + 1) * w["n_bufs"] + 64
        d_out = S.DevBuffer(2 * out_cap, cx.device)

        def step():
            return h.demodulate_batch_dev(d_in, w["buf_len"], w["n_bufs"], d_out, out_cap)
    else:
        taps, taps2 = taps_for(w)
        h = S.FmRx(taps, w["D"], taps2, w["up"], w["down"], device=cx.device)
        h.seek(n * cx.rank)
        _, na = h.out_lens(n)
        out_cap = na + 64
        d_out = S.DevBuffer(4 * out_cap, cx.device)

        def step():
            return h.process_dev(d_in, n, d_out, out_cap)

    def barrier():
        cx.barrier()
        h.sync()

    # the clock sampler starts BEFORE the warm-up and is given time to print its first sample: a nvidia-smi process that
    # starts inside a timed region of a few milliseconds stalls the launches of that region (measured on the 8-GPU box:
    # 0.60 instead of 0.35 ms per step).  It then samples every 100 ms through warm-up, timed region and the load tail below.
    sampler = ClockSampler(cx.device) if (sample_clocks and cx.rank == 0) else None
    if sampler:
        sampler.start()
        sampler.wait_first()
    for _ in range(warmup):
        step()
    barrier()
    if w["name"] != "cfg1":
        h.timing_totals(reset=True)
    launches0 = S.kernel_launch_count()
    barrier()
    h.span_begin()
    for _ in range(steps):
        step()
    total_ms = h.span_end()          # records the closing event and waits for it
    barrier()
    launches = S.kernel_launch_count() - launches0
    if w["name"] != "cfg1":
        sums, calls = h.timing_totals()
    if sample_clocks:
        # load tail (untimed, every rank, same step): the timed region lasts a few milliseconds, the sampler's period is
        # 100 ms — keep the identical load running until it has taken three more samples (at most 0.6 s)
        have = cx.sum_over_ranks(sampler.n_samples() if sampler else 0)
        t_end = time.perf_counter() + 0.6
        while True:
            for _ in range(20):
                step()
            h.sync()
            now = cx.sum_over_ranks(sampler.n_samples() if sampler else 0)
            if cx.sum_over_ranks(int(now >= have + 3 or time.perf_counter() > t_end)) > 0:
                break
    clocks = sampler.stop() if sampler else None
    if w["name"] == "cfg1":
        kern_ms = total_ms / steps
    else:
        kern_ms = sums[0] / max(calls, 1)
    total_ms, kern_ms = cx.max_over_ranks(total_ms, kern_ms)
    launches = cx.sum_over_ranks(launches)
    h.close()
    d_out.free()
    return {"total_ms": total_ms, "ms_per_step": total_ms / steps, "kern_ms": kern_ms, "launches": launches,
            "clocks": clocks, "steps": steps}


def roofline_of(w: dict, m: dict) -> dict:
    peak, peak_src = measured_peak()
    bps = alg_bytes_per_sample(w)
    achieved = bps * w["n"] / (m["kern_ms"] * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / f"traffic_{w['name']}.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic,
            "traffic_source": f"profiles/traffic_{w['name']}.json (one `ncu --set full` capture of this kernel at this size; "
                              "not re-measured in this run)" if traffic else None,
            "peak_source": peak_src,
            "kernel": "k_demod_direct<6>" if w["name"] == "cfg1" else "k_fir_fast (fused convert+FIR+demod)",
            "kernel_ms": round(m["kern_ms"], 4), "alg_bytes_per_sample": round(bps, 4),
            "kernel_share_of_step": round(m["kern_ms"] / m["ms_per_step"], 4),
            "whole_step_gbs": round((bps * w["n"]) / (m["ms_per_step"] * 1e-3) / 1e9, 1),
            "whole_step_frac": round((bps * w["n"]) / (m["ms_per_step"] * 1e-3) / 1e9 / peak, 4),
            "note": "peak is the pool's COPY benchmark (half of its bytes are writes); this kernel's bytes are "
                    ">= 97 % reads, so frac can exceed 1 — against the 7.7 TB/s HBM3e nominal it is "
                    f"{achieved / 7700.0:.3f}"}


def measure_e2e(cx: Ctx, w: dict, steps: int, pageable: bool = False, registered: bool = False) -> dict:
    """The same metric through the public C-ABI call with HOST buffers: H2D of the input and D2H of the audio inside the
    timed region, every step.  pageable=False: pinned host memory (sdr_host_alloc); True: an ordinary heap array, which is
    what a drop-in caller's Vec<u8> is (examples/simple_fm.rs:80,153); registered=True: the same heap array, page-locked
    once with sdr_host_register before the timed region (what a caller with a long-lived buffer would do)."""
    S = cx.S
    from rtl_sdr_rs_b200 import _ffi as F
    n_e = 1 << 27
    hb = S.HostBuffer(2 * n_e)
    src = S.Source.open_synth(SEED + 1 + cx.rank)
    assert src.read_sync(hb.array) == 2 * n_e
    src.close()
    in_arr = hb.array
    if pageable:
        in_arr = np.empty(2 * n_e, np.uint8)
        in_arr[:] = hb.array
        hb.free()
        if registered:
            S.host_register(in_arr)
    in_ptr = F.ptr(in_arr)
    e_steps = max(3, min(steps, 10))
    if w["name"] == "cfg1":
        he = S.Demod(device=cx.device)
        n_out_cap = n_e // 6 + 64
        out_h = S.HostBuffer(2 * n_out_cap, np.int16) if not pageable else None
        out_arr = out_h.array if out_h else np.empty(n_out_cap, np.int16)

        def e_step():
            return F.check(F.lib().sdr_demod_demodulate_batch(he._h, in_ptr, w["buf_len"], 2 * n_e // w["buf_len"],
                                                              F.ptr(out_arr), out_arr.size, None))
    else:
        taps, taps2 = taps_for(w)
        he = S.FmRx(taps, w["D"], taps2, w["up"], w["down"], device=cx.device)
        n_out_cap = n_e // w["D"] * w["up"] // w["down"] + 64
        out_h = S.HostBuffer(4 * n_out_cap, np.float32) if not pageable else None
        out_arr = out_h.array if out_h else np.empty(n_out_cap, np.float32)

        def e_step():
            return F.check(F.lib().sdr_fmrx_process(he._h, in_ptr, n_e, None, 0, None, 0, F.ptr(out_arr), out_arr.size))
    n_out_e = 0
    for _ in range(2):
        n_out_e = e_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        n_out_e = e_step()          # synchronous: returns when the host output buffer is filled
    e_s = time.perf_counter() - t0
    (e_s,) = cx.max_over_ranks(e_s)
    he.close()
    if not pageable:
        hb.free()
    elif registered:
        S.host_unregister(in_arr)
    if out_h:
        out_h.free()
    return {"value": round(cx.world * n_e * e_steps / e_s / 1e6, 2), "unit": "Msamples/s",
            "h2d_bytes_per_step": 2 * n_e, "d2h_bytes_per_step": int(n_out_e) * (2 if w["name"] == "cfg1" else 4),
            "steps": e_steps, "samples_per_step": n_e,
            "api": "sdr_demod_demodulate_batch" if w["name"] == "cfg1" else "sdr_fmrx_process",
            "host_memory": ("caller-owned heap array page-locked once with sdr_host_register" if registered else
                            "pageable (ordinary heap array, like the caller's Vec<u8>)") if pageable else "pinned (sdr_host_alloc)",
            "note": "host input -> chunked H2D overlapped with the kernels -> D2H of the audio, per step"}


def measure_h2d_ceiling(cx: Ctx) -> dict:
    """Bare host->device copy rate of this box with all N ranks copying at once (pinned memory, 256 MiB per copy): the
    ceiling any end-to-end figure with host input can reach.  No kernels involved."""
    S = cx.S
    nbytes = 1 << 28
    hb = S.HostBuffer(nbytes)
    hb.array[::4096] = 1                      # touch every page
    db = S.DevBuffer(nbytes, cx.device)
    db.upload(hb.array)                       # warm-up
    cx.barrier()
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        db.upload(hb.array)                   # cudaMemcpy from pinned memory: one DMA, returns when done
    dt = time.perf_counter() - t0
    (dt,) = cx.max_over_ranks(dt)
    hb.free()
    db.free()
    gbs = cx.world * nbytes * reps / dt / 1e9
    return {"aggregate_gbs": round(gbs, 1), "msamples_per_s": round(gbs * 1e3 / 2, 1), "ranks": cx.world,
            "how": f"{reps} x cudaMemcpy(256 MiB, pinned -> device) per rank, all ranks at once, max over ranks"}


def per_buffer_bench(S, device: int, buf_len: int, n_bufs: int = 400) -> dict:
    """Latency/throughput of the drop-in call pattern: one `buf_len`-byte host buffer per call."""
    src = S.Source.open_synth(SEED + 77)
    bufs = np.empty((8, buf_len), np.uint8)
    for b in bufs:
        assert src.read_sync(b) == buf_len
    src.close()
    d = S.Demod(device=device)
    for i in range(20):
        d.demodulate(bufs[i % 8])
    lat = []
    t0 = time.perf_counter()
    for i in range(n_bufs):
        t1 = time.perf_counter()
        d.demodulate(bufs[i % 8])
        lat.append(time.perf_counter() - t1)
    sync_s = time.perf_counter() - t0
    lat.sort()
    out = {"buf_len": buf_len, "calls": n_bufs,
           "sync_call": {"api": "sdr_demod_demodulate", "us_per_call_median": round(lat[len(lat) // 2] * 1e6, 1),
                         "us_per_call_p99": round(lat[int(len(lat) * 0.99)] * 1e6, 1),
                         "msamples_per_s": round(n_bufs * (buf_len // 2) / sync_s / 1e6, 1)}}
    d.close()
    # ring: producer keeps up to n_slots - 1 buffers in flight, consumer collects in order
    d = S.Demod(device=device)
    slots = 8
    ring = S.Ring(d, buf_len, slots)
    for i in range(slots - 1):
        ring.submit(bufs[i % 8])
    t0 = time.perf_counter()
    for i in range(n_bufs):
        ring.collect()
        ring.submit(bufs[i % 8])
    ring_s = time.perf_counter() - t0
    for i in range(slots - 1):
        ring.collect()
    ring.close()
    d.close()
    out["ring"] = {"api": "sdr_ring_acquire/commit/collect", "slots": slots,
                   "us_per_buffer": round(ring_s / n_bufs * 1e6, 1),
                   "msamples_per_s": round(n_bufs * (buf_len // 2) / ring_s / 1e6, 1)}
    return out


def per_buffer_fx(S, device: int, w: dict, buf_len: int = 262144, n_bufs: int = 400) -> dict:
    """The f32 receiver on the reference's call pattern (one 262144-byte host buffer at a time): one synchronous
    sdr_fmrx_process per buffer, and the persistent ring (sdr_fmrx_ring_*: no launch per buffer)."""
    taps, taps2 = taps_for(w)
    src = S.Source.open_synth(SEED + 78)
    bufs = np.empty((8, buf_len), np.uint8)
    for b in bufs:
        assert src.read_sync(b) == buf_len
    src.close()
    rx = S.FmRx(taps, w["D"], taps2, w["up"], w["down"], device=device)
    for i in range(20):
        rx.process(bufs[i % 8], want_y=False, want_demod=False)
    lat = []
    t0 = time.perf_counter()
    for i in range(n_bufs):
        t1 = time.perf_counter()
        rx.process(bufs[i % 8], want_y=False, want_demod=False)
        lat.append(time.perf_counter() - t1)
    sync_s = time.perf_counter() - t0
    lat.sort()
    out = {"buf_len": buf_len, "calls": n_bufs,
           "sync_call": {"api": "sdr_fmrx_process", "us_per_call_median": round(lat[len(lat) // 2] * 1e6, 1),
                         "us_per_call_p99": round(lat[int(len(lat) * 0.99)] * 1e6, 1),
                         "msamples_per_s": round(n_bufs * (buf_len // 2) / sync_s / 1e6, 1)}}
    slots = 8
    ring = S.FmRing(rx, buf_len, slots)
    for i in range(slots - 1):
        ring.submit(bufs[i % 8])
    t0 = time.perf_counter()
    for i in range(n_bufs):
        ring.collect()
        ring.submit(bufs[i % 8])
    ring_s = time.perf_counter() - t0
    for i in range(slots - 1):
        ring.collect()
    ring.close()
    rx.close()
    out["ring"] = {"api": "sdr_fmrx_ring_acquire/commit/collect", "slots": slots,
                   "us_per_buffer": round(ring_s / n_bufs * 1e6, 1),
                   "msamples_per_s": round(n_bufs * (buf_len // 2) / ring_s / 1e6, 1)}
    return out


def chan_plan(w: dict, world: int, rank: int):
    """Channel plan of BASELINE.json configs[3]/[4].  One GPU: 64 channels at f_c = (c - 31.5) * 200 kHz
    (cfg4).  N GPUs: 64*N channels at f_c = (c - (64N-1)/2) * fs/(64N) (cfg5 is N = 8: (c - 255.5) * fs/512), 64 per
    rank, INTERLEAVED — rank r owns c = r (mod N), so that its own 64 channels sit on a uniform fs/64 grid and share the
    polyphase bank's first stage; which 64 a rank owns changes nothing else (outputs stay rank-local)."""
    from sigutil import channel_taps
    C = w["C"]
    c_tot = C * world
    taps = channel_taps(w["T"], w["D"])
    if world == 1:
        offs = (np.arange(C) - 31.5) * 200e3
    else:
        offs = ((rank + world * np.arange(C)) - (c_tot - 1) / 2.0) * (w["fs"] / c_tot)
    fw = (np.round(offs / w["fs"] * 2.0 ** 32).astype(np.int64) % (1 << 32)).astype(np.uint32)
    return taps, fw, c_tot


def chan_roofline(w: dict, info: dict, kern_ms: float, sm_mhz: float | None, kernel: str = "") -> dict:
    C, T, D, slab = w["C"], w["T"], w["D"], w["slab"]
    fma_per_sample = 4.0 * C * T / D                # direct form: complex tap x complex sample = 4 FMAs
    mhz = sm_mhz or 1965.0
    fma_peak = info["sm_count"] * 128 * mhz * 1e6 / 1e12          # TFMA/s at the observed clock
    ach = fma_per_sample * slab / (kern_ms * 1e-3) / 1e12
    peak, _ = measured_peak()
    hbm_bytes = (2.0 + C * 4.0 / D) * slab                        # u8 IQ in, f32 discriminator out per channel
    return {"kernel": kernel, "kernel_ms_per_slab": round(kern_ms, 4), "slab_samples": slab,
            "direct_form_tfma_per_s": round(ach, 2), "fp32_fma_peak_tfma_per_s": round(fma_peak, 2),
            "direct_form_fma_frac": round(ach / fma_peak, 4),
            "hbm_gbs": round(hbm_bytes / (kern_ms * 1e-3) / 1e9, 1), "hbm_frac": round(hbm_bytes / (kern_ms * 1e-3) / 1e9 / peak, 4),
            "note": "direct_form_* counts the 4*C*T/D FMAs per input sample of the direct definition (a value above 1 means "
                    "the kernel does less arithmetic than the direct form); hbm_* counts 2 B in + 4*C/D B out per sample"}


def measure_chan(cx: Ctx, w: dict, d_in, steps: int, warmup: int, shard: bool) -> dict:
    """The channeliser over n = 2^28 samples per step in 2^25-sample slabs.  shard=False: this GPU alone, input resident.
    shard=True (N > 1): rank r owns channels [64r, 64r+64) of 64*N; rank 0 broadcasts every raw slab with one
    ncclBroadcast(u8) on its own stream; a slab's broadcast waits only for the channelising of the slab that occupied the
    same memory one step earlier (per-slab marks), so it runs under the previous slab's kernels."""
    S = cx.S
    from rtl_sdr_rs_b200 import _ffi as F
    C, D, n, slab = w["C"], w["D"], w["n"], w["slab"]
    world = cx.world if shard else 1
    rank = cx.rank if shard else 0
    taps, fw, c_tot = chan_plan(w, world, rank)
    ch = S.Channeliser(taps, D, fw, device=cx.device)
    kind, kinfo = ch.kernel_kind()
    cap = slab // D + 1
    d_dem = S.DevBuffer(4 * C * cap, cx.device)
    comm = None
    if shard and world > 1:
        uid = [S.Comm.unique_id() if cx.rank == 0 else None]
        cx.dist.broadcast_object_list(uid, src=0)
        comm = S.Comm(cx.device, cx.rank, world, uid[0])
    n_slabs = n // slab

    def step(bcast=True):
        for s in range(n_slabs):
            if comm is not None and bcast:
                comm.wait_mark(s)                       # the slab memory is free once last step's slab s was consumed
                comm.bcast_u8(d_in, 2 * slab, 0, offset=2 * slab * s)
                comm.chan_wait(ch)
            F.check(F.lib().sdr_chan_process_dev(ch._h, d_in.at(2 * slab * s), slab, None, d_dem.ptr, cap))
            if comm is not None and bcast:
                comm.mark_chan(ch, s)

    def sync():
        ch.sync()
        if comm is not None:
            comm.sync()

    def timed(k, bcast=True):
        cx.barrier()
        sync()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            step(bcast)
        sync()
        (dt,) = cx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        return dt / k

    for _ in range(warmup):
        step()
    sync()
    launches0 = S.kernel_launch_count()
    ms_step = timed(steps)
    launches = cx.sum_over_ranks(S.kernel_launch_count() - launches0)
    (kern_ms,) = cx.max_over_ranks(ch.last_timing()[0])   # device time of the last slab's channeliser kernels
    out = {"channels_total": c_tot, "channels_per_gpu": C, "n_gpus": world, "samples_per_step": n, "slab_samples": slab,
           "ms_per_step": round(ms_step, 3), "input_msamples_per_s": round(n / (ms_step * 1e-3) / 1e6, 1),
           "channel_msamples_per_s": round(c_tot * n / (ms_step * 1e-3) / 1e6, 1), "gpu_launches": launches,
           "kernel_ms_per_slab": round(kern_ms, 4),
           "kernel": {0: "k_chan_fir (shared-memory taps)", 1: "k_chan_fir_u (direct form)",
                      2: f"k_chan_bank (two-stage polyphase bank, K = {kinfo[0]} = {kinfo[1]} x {kinfo[2]})"}[kind],
           "channel_plan": "f_c = (c - 31.5) * 200 kHz" if world == 1 else
                           f"f_c = (c - {(c_tot - 1) / 2}) * fs/{c_tot}, rank r owns c = r (mod {world})",
           "timing": "host wall clock around K steps bracketed by barrier + stream syncs (multi-stream pipeline), max over ranks"}
    if comm is not None:
        # the same steps without the broadcast (every rank channelises what is already in its buffer): the difference
        # is what the broadcast costs after overlap
        ms_nobc = timed(steps, bcast=False)
        # the broadcast alone, back to back
        cx.barrier()
        sync()
        t0 = time.perf_counter()
        for _ in range(2):
            for s in range(n_slabs):
                comm.bcast_u8(d_in, 2 * slab, 0, offset=2 * slab * s)
        comm.sync()
        (bc_ms,) = cx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        bc_ms /= 2 * n_slabs
        out.update({"ms_per_step_without_bcast": round(ms_nobc, 3), "exposed_bcast_ms": round(ms_step - ms_nobc, 3),
                    "bcast_ms_per_slab": round(bc_ms, 4), "bcast_gbs": round(2 * slab / (bc_ms * 1e-3) / 1e9, 1),
                    "collective": "one ncclBroadcast(ncclUint8) per 64 MiB slab from rank 0, own stream; no other collective",
                    "bcast_fraction_of_step_if_serial": round(bc_ms * n_slabs / ms_step, 3)})
        comm.close()
    ch.close()
    d_dem.free()
    return out


# ---------------------------------------------------------------------------------------------------
# CPU arm (cpu_baseline leg and --impl reference): the oracle port timed on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_run(w: dict, sample_samples: int, reps: int, threads: int, fused: bool = False):
    """Returns (seconds per rep list, samples per rep).  This is the ONLY place bench.py executes oracle/."""
    import ctypes as C
    import oracle_ffi as O
    L = O.lib()
    if w["name"] == "cfg1":
        buf_len = w["buf_len"]
        n_bufs = max(threads, sample_samples * 2 // buf_len)
        data = O.synth_fill(n_bufs * buf_len, SEED)
        _, cfg = O.optimal_settings()
        out = np.empty(n_bufs * 4200, np.int16)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = L.orc_demodulate_many_mt2(C.byref(cfg), O._p(data, C.c_uint8), buf_len, n_bufs, O._p(out, C.c_int16),
                                          out.size, threads, int(fused))
            times.append(time.perf_counter() - t0)
            assert r > 0
        return times, n_bufs * buf_len // 2
    taps, taps2 = taps_for(w)
    iq = O.synth_fill(2 * sample_samples, SEED)
    audio = np.empty(sample_samples // w["D"] * w["up"] // w["down"] + 16, np.float32)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        na = L.orc_fx_process_f32_mt(O._p(iq, C.c_uint8), sample_samples, O._p(taps, C.c_float), taps.size, w["D"],
                                     O._p(taps2, C.c_float), taps2.size, w["up"], w["down"], C.c_float(16384.0 / np.pi),
                                     O._p(audio, C.c_float), audio.size, threads)
        times.append(time.perf_counter() - t0)
        assert na > 0
    return times, sample_samples


def cpu_leg(w: dict, threads: int, fused: bool, sample: int, budget_s: float) -> dict:
    t, n = cpu_run(w, sample, 1, threads, fused)          # warm-up / calibration
    reps = int(max(2, min(200, budget_s / max(t[0], 1e-3))))
    times, n = cpu_run(w, sample, reps, threads, fused)
    kind = ("single-pass fused restatement (orc_demodulate_fused)" if fused else
            "ref-like integer Demod: one fresh vector per stage like examples/simple_fm.rs:256-269 (orc_demodulate_ref_like)") \
        if w["name"] == "cfg1" else "f32 FIR+atan2f+resampler port"
    return {"value": round(n / statistics.median(times) / 1e6, 2), "best": round(n / min(times) / 1e6, 2),
            "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": f"{reps} passes over {n} complex samples (seeded synthetic, same taps/config), oracle port: {kind}, "
                      f"{threads} thread{'s' if threads != 1 else ''}"
                      + ("; one independent Demod per thread" if w["name"] == "cfg1" and threads > 1 else "")}


def cpu_baseline(w: dict, budget_s: float = 8.0) -> dict:
    import oracle_ffi as O
    return cpu_leg(w, O.max_threads(), False, 1 << 26, budget_s)


def cpu_legs_cfg1(w: dict) -> dict:
    """(1) ref_like on ONE thread — the reference's `process` is one thread (examples/simple_fm.rs:60,135);
    (2) the fused single-pass restatement on one thread and on all threads; plus ref_like on all threads."""
    import oracle_ffi as O
    nt = O.max_threads()
    return {"ref_like_1_thread": cpu_leg(w, 1, False, 1 << 24, 2.5),
            "fused_1_thread": cpu_leg(w, 1, True, 1 << 24, 2.5),
            "ref_like_all_threads": cpu_leg(w, nt, False, 1 << 26, 3.0),
            "fused_all_threads": cpu_leg(w, nt, True, 1 << 26, 3.0)}


def run_reference_arm(args, w):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import oracle_ffi as O
    threads = O.max_threads()
    sample = 1 << 26
    times, n = cpu_run(w, sample, args.warmup + args.steps, threads)
    times = times[args.warmup:]
    total = sum(times)
    value = n * len(times) / total / 1e6
    info = None
    try:                                    # same `config` object as the GPU arm (the box's device is part of it)
        import sdrpkg
        S = sdrpkg.load()
        if S.device_count() > 0:
            info = S.device_info(0)
    except Exception:
        info = None
    cfg = config_of(w, info)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": cfg,
        "note": "the reference (Rust) cannot be built in this image: this arm is the oracle's CPU restatement of the same "
                f"workload on all host threads; each step is a bounded sample of {n} complex samples of it",
        "cpu_baseline": {"value": round(value, 2), "unit": "Msamples/s", "cores": threads, "kind": "port",
                         "sample": f"each step = {n} complex samples of the workload"},
        "e2e": {"value": round(value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------
def run_chan(args, w):
    """--workload chan: the channeliser as the headline line (1 GPU: alone; N > 1: the channel shard)."""
    cx = Ctx()
    S = cx.S
    d_in = S.DevBuffer(2 * w["n"], cx.device)
    if cx.rank == 0 or cx.world == 1:
        S.synth_fill_dev(d_in, 2 * w["n"], SEED)
    sampler = ClockSampler(cx.device) if cx.rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()          # nvidia-smi's start-up must not overlap the timed region
    m = measure_chan(cx, w, d_in, args.steps, max(args.warmup, 3), shard=cx.world > 1)
    if sampler:
        time.sleep(0.15)
    clocks = sampler.stop() if sampler else None
    if cx.rank == 0:
        line = {"metric": "channel-Msamples/s (input Msamples/s x channels) through the channeliser", "workload": "chan",
                "value": m["channel_msamples_per_s"], "unit": "channel-Msamples/s",
                "input_msamples_per_s": m["input_msamples_per_s"], "n_gpus": cx.world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "channels_total": m["channels_total"], "samples_per_step": w["n"],
                           "slab_samples": w["slab"], "timing": m["timing"], "device": cx.info["name"]},
                "roofline": chan_roofline(w, cx.info, m["kernel_ms_per_slab"], (clocks or {}).get("sm_mhz"), m["kernel"]),
                "multi_gpu": m if cx.world > 1 else None, "clocks": clocks, "gpu_launches": m["gpu_launches"]}
        print(json.dumps(line))
    cx.close()
    return 0


def guarded(errors: dict, name: str, fn, *a, **kw):
    """Run one of the legs that FOLLOW the headline's timed region; a failure there is recorded in the line
    (`extra_errors`) and must not cost the headline number."""
    try:
        return fn(*a, **kw)
    except Exception as e:   # noqa: BLE001 — reported, not swallowed
        errors[name] = f"{type(e).__name__}: {e}"[:400]
        print(f"bench.py: leg '{name}' failed: {errors[name]}", file=sys.stderr)
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg1", "chan"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (profiling runs)")
    ap.add_argument("--n-log2", type=int, default=0, help="override samples per GPU per step (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = workload_spec(args.workload)
    if args.n_log2:
        w["n"] = 1 << args.n_log2
        if w["name"] == "cfg1":
            w["n_bufs"] = w["n"] // 131072
    if args.impl == "reference":
        if w["name"] == "chan":
            raise SystemExit("--impl reference is defined for cfg1/cfg2/cfg3")
        return run_reference_arm(args, w)
    if w["name"] == "chan":
        return run_chan(args, w)

    cx = Ctx()
    S = cx.S
    n = w["n"]
    # each rank owns the time slice [rank*n, (rank+1)*n) of ONE synthetic stream
    d_in = S.DevBuffer(2 * n, cx.device)
    S.synth_fill_dev(d_in, 2 * n, SEED, byte_offset=2 * n * cx.rank)

    m = measure_stream(cx, w, d_in, args.steps, args.warmup, sample_clocks=True)
    value = cx.world * n / (m["ms_per_step"] * 1e-3) / 1e6               # whole-job Msamples/s
    alone, errors = None, {}
    if w["name"] != "cfg1":
        # In the timed region the audio kernel of step k runs UNDER the fused kernel of step k+1 (own stream), so the fused
        # kernel's event-bracketed time includes that contention.  The same kernel with the audio stage serialised behind it
        # (a handle created with SDR_FMRX_AUDIO_STREAM=serial), measured after the timed region:
        os.environ["SDR_FMRX_AUDIO_STREAM"] = "serial"
        try:
            alone = guarded(errors, "kernel_alone", measure_stream, cx, w, d_in, max(5, min(args.steps, 10)), 3)
        finally:
            del os.environ["SDR_FMRX_AUDIO_STREAM"]
    e2e = None if args.no_e2e else guarded(errors, "e2e", measure_e2e, cx, w, args.steps)

    # ---- everything below runs AFTER the headline's timed region --------------------------------------------------
    extra, multi = None, None
    full_size = not args.n_log2
    if not args.no_extra and full_size:
        extra = {}
        x_steps = max(5, min(args.steps, 10))
        h2d = guarded(errors, "h2d_ceiling", measure_h2d_ceiling, cx)
        if e2e is not None:
            e2e_page = guarded(errors, "e2e.pageable", measure_e2e, cx, w, args.steps, pageable=True)
            if e2e_page:
                e2e["pageable"] = {k: e2e_page[k] for k in ("value", "unit", "host_memory")}
                e2e["pageable"]["fraction_of_pinned"] = round(e2e_page["value"] / e2e["value"], 3)
            e2e_reg = guarded(errors, "e2e.registered", measure_e2e, cx, w, args.steps, pageable=True, registered=True)
            if e2e_reg:
                e2e["registered"] = {k: e2e_reg[k] for k in ("value", "unit", "host_memory")}
                e2e["registered"]["fraction_of_pinned"] = round(e2e_reg["value"] / e2e["value"], 3)
            if h2d:
                e2e["h2d_ceiling"] = h2d
                e2e["fraction_of_h2d_ceiling"] = round(e2e["value"] / h2d["msamples_per_s"], 3)
        for other in ("cfg1", "cfg3", "cfg2"):
            if other == w["name"]:
                continue
            wo = workload_spec(other)
            mo = guarded(errors, f"extra.{other}", measure_stream, cx, wo, d_in, x_steps, 3)
            if mo is None:
                continue
            ro = roofline_of(wo, mo)
            eo = {"workload": wo["desc"], "ms_per_step": round(mo["ms_per_step"], 4),
                  "msamples_per_s": round(cx.world * wo["n"] / (mo["ms_per_step"] * 1e-3) / 1e6, 1), "steps": x_steps,
                  "gpu_launches": mo["launches"],
                  "roofline": {k: ro[k] for k in ("achieved", "peak", "unit", "frac", "kernel", "kernel_ms", "alg_bytes_per_sample",
                                                  "kernel_share_of_step", "whole_step_frac", "traffic")}}
            if other == "cfg1" and not args.no_e2e:
                e1 = guarded(errors, "extra.cfg1.e2e", measure_e2e, cx, wo, x_steps)
                e1p = guarded(errors, "extra.cfg1.e2e.pageable", measure_e2e, cx, wo, x_steps, pageable=True)
                if e1:
                    eo["e2e"] = {"value": e1["value"], "unit": "Msamples/s", "api": e1["api"], "host_memory": e1["host_memory"],
                                 "h2d_bytes_per_step": e1["h2d_bytes_per_step"], "d2h_bytes_per_step": e1["d2h_bytes_per_step"],
                                 "pageable_value": e1p["value"] if e1p else None,
                                 "fraction_of_h2d_ceiling": round(e1["value"] / h2d["msamples_per_s"], 3) if h2d else None}
                if cx.rank == 0:
                    eo["per_buffer"] = guarded(errors, "extra.cfg1.per_buffer", per_buffer_bench, S, cx.device, wo["buf_len"])
            if other == "cfg1" and not args.no_cpu_baseline and cx.rank == 0:
                eo["cpu_baseline"] = guarded(errors, "extra.cfg1.cpu_baseline", cpu_legs_cfg1, wo)
            extra[other] = eo
        if not args.no_e2e and cx.rank == 0 and w["name"] != "cfg1":
            extra["per_buffer_f32"] = guarded(errors, "extra.per_buffer_f32", per_buffer_fx, S, cx.device, w)
        wc = workload_spec("chan")
        mc = guarded(errors, "extra.chan", measure_chan, cx, wc, d_in, max(3, x_steps // 2), 3, shard=False)
        if mc:
            extra["chan"] = {"workload": wc["desc"], "ms_per_step": mc["ms_per_step"], "input_msamples_per_s": mc["input_msamples_per_s"],
                             "channel_msamples_per_s": mc["channel_msamples_per_s"], "gpu_launches": mc["gpu_launches"],
                             "channel_plan": mc["channel_plan"],
                             "roofline": chan_roofline(wc, cx.info, mc["kernel_ms_per_slab"], (m["clocks"] or {}).get("sm_mhz"), mc["kernel"])}
        if cx.world > 1:
            # north_star's multi-GPU design: the channel shard with the NCCL slab broadcast.  Ranks other than 0 hold
            # whatever their time slice left in d_in; every slab is overwritten by the broadcast before it is read.
            multi = guarded(errors, "multi_gpu", measure_chan, cx, wc, d_in, 10, 5, shard=True)
            if multi and mc:
                multi["single_gpu_ms_per_step"] = mc["ms_per_step"]
                multi["efficiency_vs_one_gpu_same_run"] = round(mc["ms_per_step"] / multi["ms_per_step"], 4)

    if w["name"] == "cfg1" and not args.no_e2e and cx.rank == 0 and (args.no_extra or not full_size):
        extra = extra or {}
        extra["per_buffer"] = guarded(errors, "extra.per_buffer", per_buffer_bench, S, cx.device, w["buf_len"])

    if cx.rank != 0:
        cx.close()
        return 0

    cfg = config_of(w, cx.info)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "Msamples/s", "n_gpus": cx.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(m["ms_per_step"], 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic", "config": cfg,
        "roofline": roofline_of(w, m), "clocks": m["clocks"], "e2e": e2e, "gpu_launches": int(m["launches"]),
    }
    if alone is not None:
        ra = roofline_of(w, alone)
        line["roofline"]["kernel_alone"] = {
            "kernel_ms": ra["kernel_ms"], "achieved": ra["achieved"], "frac": ra["frac"], "step_ms_serialised": round(alone["ms_per_step"], 4),
            "note": "same kernel, audio stage serialised behind it instead of overlapped with the next step (measured in this run, "
                    "after the timed region): the kernel itself streams at this rate; overlapping costs it time but shortens the step"}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = guarded(errors, "cpu_baseline", cpu_baseline, w)
        if w["name"] == "cfg1":
            line["cpu_baseline_legs"] = guarded(errors, "cpu_baseline_legs", cpu_legs_cfg1, w)
    if extra:
        line["extra"] = extra
    if multi:
        line["multi_gpu"] = multi
    if errors:
        line["extra_errors"] = errors
    print(json.dumps(line))
    cx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
