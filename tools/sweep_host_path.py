#!/usr/bin/env python
"""Sweep of the host-buffer entry point b200c_fir_run_host (the e2e number of bench.py): staging chunk size x
pipeline depth (streams / staging slots), headline workload, pinned host buffers, next to the box's plain
cudaMemcpy rates.  One subprocess per point (the knobs are read once per process).
  python tools/sweep_host_path.py > gpurun_out/r02_sweep_host_path.jsonl"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import json, os, sys, time
sys.path.insert(0, %r)
import torch
from pothoscomms_b200 import FirFilter
from pothoscomms_b200 import workloads as wl
taps, tt = wl.config_taps("headline")
f = FirFilter(1, tt); f.set_taps(taps)
n = 1 << 27
x = wl.tone_noise_torch(1, 255 + n, 1, torch.device("cuda", 0))
h_in = torch.empty((255 + n, 2), dtype=torch.float32).pin_memory(); h_in.copy_(x)
h_out = torch.empty((n, 2), dtype=torch.float32).pin_memory()
xi, yo = h_in.numpy(), h_out.numpy()
f.run_host(xi, out=yo, out_capacity=n)
best = 1e9
for _ in range(4):
    t0 = time.perf_counter(); f.run_host(xi, out=yo, out_capacity=n); best = min(best, time.perf_counter() - t0)
print(json.dumps({"chunk_mib": int(os.environ.get("B200C_HOST_CHUNK_MIB", "32")), "slots": int(os.environ.get("B200C_HOST_SLOTS", "3")),
                  "msamples_per_s": n / best / 1e6, "gbs_each_way": 8.0 * n / best / 1e9}))
""" % ROOT


def memcpy_peaks():
    import torch
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h2.copy_(d2, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        out[name + "_gbs"] = 3 * n / (time.perf_counter() - t0) / 1e9
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    out["bidirectional_gbs_each_way"] = 3 * n / (time.perf_counter() - t0) / 1e9
    return out


def main():
    print(json.dumps({"probe": "cudaMemcpy pinned, 1 GiB", **memcpy_peaks()}), flush=True)
    for chunk in (8, 16, 32, 64, 128):
        for slots in (2, 3, 4, 6):
            env = dict(os.environ, B200C_HOST_CHUNK_MIB=str(chunk), B200C_HOST_SLOTS=str(slots))
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else json.dumps({"chunk_mib": chunk, "slots": slots, "error": r.stderr[-300:]}), flush=True)


if __name__ == "__main__":
    main()
