#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: FIR Msamples/s (cf32, 256 taps) on 1/2/4/8 B200.

One "step" = one pass of the /comms/fir_filter hot path over one batch of synthetic
tone+noise input (per GPU: 2^28 complex-float32 samples = 2 GiB in + 2 GiB out, far larger
than the 126 MB L2, so nothing is cache-resident between steps).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's CPU path on the host cores

Multi-GPU (N>1): ONE long stream of N*2^28 samples is split into contiguous segments, one per
rank; before every pass rank r sends the last K-1 samples of its segment to rank r+1
(NCCL P2P over NVLink) -- the only exchange the path has.  Weak scaling.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

def _baseline_metric() -> str:
    """BASELINE.json's metric string, verbatim (the file travels with the repo)."""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "FIR Msamples/s (cf32, 256 taps) at 1/2/4/8 B200; % HBM roofline; vs host CPU"


METRIC = _baseline_metric()
UNIT = "Msamples/s"

WORKLOADS = {
    # name: (dtype name, taps config, decim, interp, log2 samples per GPU, bytes/input sample (algorithmic))
    "headline": ("complex_float32", "headline", 1, 1, 28, 16.0),
    "c1": ("complex_float32", "c1", 1, 1, 28, 16.0),
    "c1_real": ("complex_float32", "c1_real", 1, 1, 28, 16.0),
    "c2": ("complex_int16", "c2", 1, 1, 28, 8.0),
    "c3": ("complex_float32", "c3", 2, 3, 28, 20.0),
    "c5": ("complex_float32", "c5", 1, 1, 26, 16.0),
    # short-tap streams: the HBM-bound side of the path (north_star's ">= 70 % of HBM roofline
    # for short-tap FIR and resampling")
    "short": ("complex_float32", "short", 1, 1, 28, 16.0),
    "short_cx": ("complex_float32", "short_cx", 1, 1, 28, 16.0),
    "resamp_short": ("complex_float32", "resamp_short", 2, 3, 28, 20.0),
    "real64": ("float32", "real64", 1, 1, 29, 8.0),
    "real64_i16": ("int16", "real64", 1, 1, 29, 4.0),
    # the C3 resampler on the bit-exact fixed-point type (4 B in + 6 B out per input sample)
    "c3_i16": ("complex_int16", "c3_i16", 2, 3, 28, 10.0),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def workload_config(wl_name: str) -> dict:
    """The `config` object of a FIR workload: the same for the GPU arm and the reference arm."""
    from pothoscomms_b200 import workloads as wl
    dt_name, taps_name, M, L, log2n, _ = WORKLOADS[wl_name]
    taps, tt = wl.config_taps(taps_name)
    return {"workload": f"{wl_name}: /comms/fir_filter {dt_name} {len(taps)} {tt} taps decim={M} interp={L}, "
                        f"2^{log2n} samples per GPU tone+noise, one stream split in contiguous segments with K-1 halo",
            "samples_per_gpu": 1 << log2n, "l2_policy": "inputs larger than L2 (2 GiB in + 2 GiB out per pass)"}


def workload_traffic(wl_name: str, samples: int):
    """DRAM bytes one launch moves (ncu dram__bytes_read + write per sample, profiles/traffic.json, x the samples of
    a launch), or None when no capture of that workload's kernel is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(wl_name)
        return t["bytes_per_sample"] * samples if t else None
    except Exception:
        return None


def cpu_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_sample(wl_name: str, threads: int, log2_per_thread: int = 21, repeats: int = 1):
    """Times the FIR oracle port (kind "port": the reference FIR block needs PothosCore and
    cannot be compiled) on a bounded sample of the workload with `threads` host threads."""
    import numpy as np

    import oracle
    from pothoscomms_b200 import workloads as wl
    dt_name, taps_name, M, L, _, _ = WORKLOADS[wl_name]
    code = oracle.DTYPE_CODES[dt_name]
    taps, tt = wl.config_taps(taps_name)
    n = threads << log2_per_thread
    x = wl.tone_noise_numpy(code, min(n, 1 << 22), seed=0xC0FFEE01)
    if x.shape[0] < n:
        x = np.tile(x, (n // x.shape[0], 1))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, cons, _ = oracle.fir(code, tt == "COMPLEX", taps, M, L, x, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return cons / best / 1e6, cons, best


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores
    (the oracle port of filter/FIRFilter.cpp:278-302 -- the block itself cannot be compiled
    without PothosCore), all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = cpu_threads()
    # W warm-up steps, then exactly K timed steps; a step is a bounded sample of the workload (2^19 samples per
    # host thread: ~0.2 s at the measured 50-60 Msamples/s) so that any K the driver picks ends within minutes
    log2_step = 19
    for _ in range(args.warmup):
        run_cpu_sample(args.workload, threads, log2_per_thread=log2_step)
    total_s, total_n = 0.0, 0
    steps = max(1, args.steps)
    for _ in range(steps):
        _, cons, dt = run_cpu_sample(args.workload, threads, log2_per_thread=log2_step)
        total_s += dt
        total_n += cons
    value = total_n / total_s / 1e6
    dt_name = WORKLOADS[args.workload][0]
    config = workload_config(args.workload)
    config["sample"] = f"{threads} host threads x 2^{log2_step} samples per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": total_s / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if "float" in dt_name else dt_name, "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{threads} threads x 2^{log2_step} samples per step, {steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def bench_fft(args):
    """Auxiliary workload (BASELINE config 4): /comms/fft 4096-point forward then inverse over
    2^28 samples.  One step = two passes (2 launches); value counts samples per pass."""
    import numpy as np
    import torch

    import oracle
    from pothoscomms_b200 import Fft
    from pothoscomms_b200 import workloads as wl
    from pothoscomms_b200.handles import dtype_code
    args.warmup = max(args.warmup, 3)
    dt_name = "complex_int16" if args.workload == "c4_i16" else "complex_float32"
    code = dtype_code(dt_name)
    n, log2n = 4096, args.log2_samples or 28
    total = 1 << log2n
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    x = wl.tone_noise_torch(code, total, 0xC0FFEE04, dev)
    X = torch.empty_like(x)
    y = torch.empty_like(x)
    fwd, inv = Fft(code, n, False), Fft(code, n, True)
    for _ in range(args.warmup):
        fwd.run(x, out=X); inv.run(X, out=y)
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(0))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fwd.run(x, out=X); inv.run(X, out=y)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    value = 2 * total / (ms * 1e-3) / 1e6
    esz = x.element_size() * 2
    peaks, peak_kind = measured_peaks()
    achieved = 2 * esz * total * 2 / (ms * 1e-3) / 1e9    # (read + write) x two passes
    e2e = None
    if not args.no_e2e:
        h_in, h_out = x.cpu().pin_memory(), torch.empty(x.shape, dtype=x.dtype).pin_memory()
        fwd.run_host(h_in.numpy(), out=h_out.numpy())
        t0 = time.perf_counter()
        for _ in range(3):
            fwd.run_host(h_in.numpy(), out=h_out.numpy())
        e2e_s = (time.perf_counter() - t0) / 3
        e2e = {"value": total / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel() * h_in.element_size()),
               "d2h_bytes_per_step": int(h_out.numel() * h_out.element_size()), "note": "forward pass only"}
    cpu = None
    if not args.no_cpu:
        threads = cpu_threads()
        nb = threads * 512
        xs = x[: nb * n].cpu().numpy()
        t0 = time.perf_counter()
        if oracle.have_ref():
            oracle.ref_fft(code, n, False, xs, threads=threads)
            kind = "reference"
        else:
            oracle.fft(code, n, False, xs)
            kind, threads = "port", 1
        secs = time.perf_counter() - t0
        cpu = {"value": nb * n / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{nb} transforms of 4096 (kiss_fft from the reference sources), {secs:.2f} s"}
    line = {
        "metric": f"FFT Msamples/s per pass ({dt_name}, 4096-point, forward+inverse)", "value": value, "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if code == 1 else "i16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: /comms/fft {dt_name} 4096-point batched forward then inverse over 2^{log2n} samples",
                   "l2_policy": "inputs larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": workload_traffic(args.workload, total), "peak_kind": peak_kind,
                     "kernel": "fft"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 2 * args.steps, "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


def bench_bank(args):
    """BASELINE config 5: 1024-channel filter bank, one 1024-tap complex band-pass per channel,
    2^20 samples per channel, channels sharded over the ranks (no exchange: SURVEY.md 8e).
    Strong scaling: the bank is fixed, every rank takes 1024/N channels.  One step = one
    b200c_fir_bank_run over the rank's channels (one launch over (channel, block))."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from pothoscomms_b200 import FirFilterBank, sharding
    from pothoscomms_b200 import workloads as wl
    from pothoscomms_b200.handles import dtype_code
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    code = dtype_code("complex_float32")
    nchan_total, ntaps = args.channels or 1024, 1024
    log2n = args.log2_samples or 20
    c0, c1 = sharding.channel_range(nchan_total, world, rank)
    nch = c1 - c0
    n_new = 1 << log2n
    bank = FirFilterBank(code, "COMPLEX", nch, device=local_rank)
    for c in range(c0, c1):
        bank.set_taps(c - c0, wl.bank_taps(c, nchan_total, ntaps))
    K = bank.info()[1]
    x = torch.empty((nch, K - 1 + n_new, 2), dtype=torch.float32, device=dev)
    base = [wl.tone_noise_torch(code, K - 1 + n_new, 0xC0FFEE05 + i, dev) for i in range(8)]
    for c in range(nch):
        x[c] = base[(c0 + c) % 8]
    del base
    out = torch.empty((nch, n_new, 2), dtype=torch.float32, device=dev)
    for _ in range(args.warmup):
        cons, prod = bank.run(x, out)
    torch.cuda.synchronize()
    assert (cons, prod) == (n_new, n_new), (cons, prod)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        bank.run(x, out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = tmax.item()
    total = nchan_total * n_new
    value = total / (ms * 1e-3) / 1e6
    # spot check: channel 0 of this rank against a single-stream filter (same library path the tests pin to the oracle)
    from pothoscomms_b200 import FirFilter
    f1 = FirFilter(code, "COMPLEX", device=local_rank)
    f1.set_taps(wl.bank_taps(c0, nchan_total, ntaps))
    y1, _, _ = f1.run(x[0].contiguous())
    assert torch.equal(y1[:4096], out[0, :4096]), "bank channel differs from its single-stream filter"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks, peak_kind = measured_peaks()
    achieved = 16.0 * nch * n_new / (ms * 1e-3) / 1e9      # this rank's kernel: its channels' bytes over the step time
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c5_bank: {nchan_total}-channel filter bank, {ntaps}-tap complex_float32 band-pass per channel, "
                               f"2^{log2n} samples per channel, channels sharded over {world} GPU(s), no exchange",
                   "channels_per_gpu": nch, "l2_policy": "inputs larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind, "kernel": "fir_os64_kernel",
                     "note": "fused overlap-save, one launch over (channel, block); 16 B per sample"},
        "cpu_baseline": None, "e2e": None, "gpu_launches": args.steps, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS) + ["c4", "c4_i16", "c5_bank"])
    ap.add_argument("--ntaps", type=int, default=None, help="tap-count sweep: cf32 L=M=1 with this many complex taps")
    ap.add_argument("--channels", type=int, default=None, help="c5_bank: total channels (default 1024)")
    ap.add_argument("--log2-samples", type=int, default=None, help="override samples per GPU (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.workload in ("c4", "c4_i16"):
        return bench_fft(args)
    if args.workload == "c5_bank":
        return bench_bank(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    from pothoscomms_b200 import FirFilter
    from pothoscomms_b200 import workloads as wl
    from pothoscomms_b200.handles import dtype_code

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.ntaps:   # tap-count sweeps: cf32 stream, `ntaps` complex band-pass taps
        w0 = WORKLOADS[args.workload]   # the named workload's stream type, `ntaps` complex band-pass taps
        dt0 = w0[0] if w0[0] in ("complex_float32", "complex_int16") else "complex_float32"
        WORKLOADS[args.workload] = (dt0, f"sweep{args.ntaps}", 1, 1, 28, 16.0 if dt0 == "complex_float32" else 8.0)
    dt_name, taps_name, M, L, log2n, bytes_per_sample = WORKLOADS[args.workload]
    if args.log2_samples:
        log2n = args.log2_samples
    code = dtype_code(dt_name)
    taps, tt = wl.config_taps(taps_name)
    fir = FirFilter(code, tt, device=local_rank)
    fir.set_taps(taps)
    fir.set_rates(M, L)
    K = fir.K
    n_seg = (1 << log2n) // M * M           # new samples per rank and step (multiple of M: SURVEY 8e)
    nc = 2 if code & 1 else 1

    # [K-1 halo | n_seg samples]; rank r's segment is samples [r*n_seg, (r+1)*n_seg) of one stream
    buf = torch.empty((K - 1 + n_seg, nc), dtype=wl.tone_noise_torch(code, 1, 0, dev).dtype, device=dev)
    buf[K - 1:] = wl.tone_noise_torch(code, n_seg, 0xC0FFEE01 + rank, dev)
    buf[: K - 1] = 0   # rank 0: the stream's first K-1 samples are history only (FIRFilter.cpp:281)
    out_cap = n_seg // M * L
    out = torch.empty((out_cap, nc), dtype=buf.dtype, device=dev)
    torch.cuda.synchronize()

    from pothoscomms_b200 import sharding

    # Optional at N > 1 (B200C_BENCH_OVERLAP=1): start the halo P2P first, run the launch over
    # everything that does not read the halo (blocks q >= q0) while it is in flight, then a second,
    # tiny launch for the q0 halo-dependent blocks.  Measured on 2 GPUs it does not pay (headline 513
    # vs 522 Gsamples/s: the persistent FIR grid and the NCCL kernel contend for SMs), so the default
    # stays exchange-then-one-launch.
    q0, in0, out0 = sharding.split_at_halo(K, M, L, align=16)
    overlap = world > 1 and n_seg // M > 4 * q0 and bool(os.environ.get("B200C_BENCH_OVERLAP"))
    launches_per_step = 2 if overlap else 1

    def fir_pass():
        if not overlap:
            sharding.exchange_halo(buf, K, rank, world)
            _, cons, prod = fir.run(buf, out=out, out_capacity=out_cap)
            return cons, prod
        works = sharding.start_halo_exchange(buf, K, rank, world)
        _, c1, p1 = fir.run(buf[in0:], out=out[out0:], out_capacity=out_cap - out0)
        sharding.finish_halo_exchange(works)
        _, c0, p0 = fir.run(buf[: in0 + K - 1], out=out[:out0], out_capacity=out0)
        return c0 + c1, p0 + p1

    def step():
        return fir_pass()

    for _ in range(args.warmup):
        cons, prod = step()
    torch.cuda.synchronize()
    assert cons == n_seg and prod == out_cap, (cons, prod, n_seg, out_cap)

    sampler = ClockSampler(physical_gpu_index(local_rank))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        if overlap:
            kev[i][0].record()
            fir_pass()
            kev[i][1].record()
        else:
            sharding.exchange_halo(buf, K, rank, world)
            kev[i][0].record()
            fir.run(buf, out=out, out_capacity=out_cap)
            kev[i][1].record()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if world > 1:
        t = torch.tensor([elapsed_ms, kernel_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = t.tolist()
    ms_per_step = elapsed_ms / args.steps
    value = world * n_seg / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H timed) ----
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty((K - 1 + n_seg, nc), dtype=buf.dtype).pin_memory()
        h_out = torch.empty((out_cap, nc), dtype=buf.dtype).pin_memory()
        h_in.copy_(buf)
        x_np, y_np = h_in.numpy(), h_out.numpy()
        e2e_steps = max(3, min(args.steps, 5))
        fir.run_host(x_np, out=y_np, out_capacity=out_cap)   # warm-up (allocates the staging slots)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fir.run_host(x_np, out=y_np, out_capacity=out_cap)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = t.item()
        # the host path must agree with the resident path
        assert torch.equal(h_out[:4096], out[:4096].cpu()) if world == 1 else True
        e2e = {"value": world * n_seg / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel() * h_in.element_size()),
               "d2h_bytes_per_step": int(h_out.numel() * h_out.element_size()), "steps": e2e_steps}
        del h_in, h_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    algo_bytes = bytes_per_sample * n_seg
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = workload_traffic(args.workload, n_seg)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                "kernel": fir.kernel, "kernel_ms": kernel_ms,
                "note": ("spectral resampler (one 1024-point forward + one 1536-point inverse transform per block): one pass over "
                         "HBM, 8 B in + 12 B out per input sample whatever the tap count"
                         if fir.kernel == "fir_os32x_kernel" else
                         "fused overlap-save (fast convolution): one pass over HBM, 16 B per sample whatever the tap count"
                         if fir.kernel.startswith("fir_os") else
                         "bit-exact int16 as byte-limb Toeplitz GEMMs on the int8 tensor cores (tcgen05 kind::i8, accumulators "
                         "in tensor memory); the HBM figure is reported next to it, the kernel is shared-memory/tensor bound"
                         if fir.kernel.startswith("fir_umma") else
                         "bit-exact int16 as byte-limb Toeplitz GEMMs on mma.sync m16n8k32 (int8 tensor cores)"
                         if fir.kernel.startswith("fir_imma") else
                         "direct form: FMA/IMAD-issue bound once taps x MACs/tap exceed ~11 flop/B; see DESIGN.md")}
    flops = {"c1": 512, "c1_real": 256, "headline": 2048, "c3": 510, "c5": 8192}.get(args.workload)
    if flops and not fir.kernel.startswith("fir_os"):
        roofline["fp32_tflops"] = flops * n_seg / (kernel_ms * 1e-3) / 1e12

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = cpu_threads()
        # ~10-20 core-seconds of CPU work: 2^22 samples per thread, best of two passes
        v, cons_cpu, secs = run_cpu_sample(args.workload, threads, log2_per_thread=22, repeats=2)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{threads} threads x 2^22 samples of the same workload ({secs:.1f} s wall per pass, best of 2)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if code in (0, 1) else dt_name, "data": "synthetic",
        "config": workload_config(args.workload),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
