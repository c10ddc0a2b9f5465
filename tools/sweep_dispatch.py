#!/usr/bin/env python
"""Where does each kernel family earn its place?  For a grid of (type, decimation, interpolation, taps) this times the
kernel the dispatcher picks (b200c_fir_kernel) against the direct-form kernel (B200C_FIR_ALGO=direct) and, for int16,
against the mma.sync kernel (B200C_FIR_ALGO=imma), on 2^26 input samples resident in HBM.  One JSON line per point:
a family that never beats the alternatives in the region it is dispatched for has no reason to stay in the library.
  python tools/sweep_dispatch.py > gpurun_out/r02_sweep_dispatch.jsonl"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pothoscomms_b200 import FirFilter  # noqa: E402
from pothoscomms_b200 import workloads as wl  # noqa: E402

CODES = {"complex_float32": 1, "float32": 0, "complex_int16": 7, "int16": 6}


@contextlib.contextmanager
def algo(name):
    old = os.environ.get("B200C_FIR_ALGO")
    if name is None:
        os.environ.pop("B200C_FIR_ALGO", None)
    else:
        os.environ["B200C_FIR_ALGO"] = name
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("B200C_FIR_ALGO", None)
        else:
            os.environ["B200C_FIR_ALGO"] = old


def measure(code, tt, taps, M, L, x, out, reps=5):
    f = FirFilter(code, tt)
    f.set_taps(taps)
    f.set_rates(M, L)
    cap = out.shape[0]
    for _ in range(2):
        f.run(x, out=out, out_capacity=cap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _, c, p = f.run(x, out=out, out_capacity=cap)
    e1.record()
    torch.cuda.synchronize()
    return f.kernel, c / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9


def main():
    dev = torch.device("cuda", 0)
    n = 1 << 26
    grid = []
    for (M, L) in ((2, 1), (1, 2), (1, 3), (3, 1), (4, 1), (1, 4), (2, 3), (3, 2), (4, 3), (3, 4), (1, 8), (1, 16), (5, 4)):
        for per_phase in (8, 24, 64, 128):
            grid.append(("complex_float32", M, L, per_phase * L, "REAL"))
    for (M, L) in ((1, 1), (2, 1), (1, 2), (2, 3)):
        for ntaps in (16, 64, 255):
            grid.append(("float32", M, L, ntaps, "REAL"))
    for ntaps in (16, 64, 128, 256, 400, 700, 1000, 2000):
        grid.append(("complex_int16", 1, 1, ntaps, "COMPLEX"))
        grid.append(("int16", 1, 1, ntaps, "REAL"))
    for (M, L) in ((2, 1), (1, 2), (2, 3), (3, 2), (4, 3), (5, 1)):
        grid.append(("complex_int16", M, L, 64 * L, "REAL"))
    bufs = {}
    for dt_name, M, L, ntaps, tt in grid:
        code = CODES[dt_name]
        if code not in bufs:
            bufs[code] = wl.tone_noise_torch(code, n + 4096, 7, dev)
        t = wl.sinc_lowpass(ntaps, 0.4 / max(M, L)) * L
        taps = t.astype(np.complex128) * np.exp(2j * np.pi * 0.03 * np.arange(ntaps)) if tt == "COMPLEX" else t
        if "int16" in dt_name:
            taps = taps * (0.45 / np.abs(taps).max())        # two-digit Q16 taps (|h| < 0.498)
        K = -(-ntaps // L)
        x = bufs[code][: K - 1 + n]
        out = torch.empty((n // M * L, x.shape[1]), dtype=x.dtype, device=dev)
        rec = {"dtype": dt_name, "decim": M, "interp": L, "ntaps": ntaps, "taps_type": tt}
        variants = [("auto", None), ("direct", "direct")] + ([("imma", "imma")] if "int16" in dt_name and M == 1 and L == 1 else [])
        for label, a in variants:
            try:
                with algo(a):
                    k, g = measure(code, tt, taps, M, L, x, out)
                rec[label] = {"kernel": k, "gsamples_per_s": round(g, 2)}
            except Exception as e:   # noqa: BLE001
                rec[label] = {"error": str(e)[:120]}
        best = max((v["gsamples_per_s"], k) for k, v in rec.items() if isinstance(v, dict) and "gsamples_per_s" in v)
        rec["best"] = best[1]
        print(json.dumps(rec), flush=True)
        del out


if __name__ == "__main__":
    main()
