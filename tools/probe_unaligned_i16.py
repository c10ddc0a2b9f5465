"""Throughput of the int16 tensor-core FIR when the input window starts at an element that is not 16-byte aligned
(what the block layer hands the kernel after the first work() call: the read pointer has advanced by n - (K - 1))."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pothoscomms_b200 import FirFilter  # noqa: E402

rng = np.random.default_rng(1)
for name, code, nc, taps in (("complex_int16 128 complex taps", 7, 2, (rng.standard_normal(128) + 1j * rng.standard_normal(128)) * 0.02),
                             ("int16 64 real taps", 6, 1, rng.standard_normal(64) * 0.03)):
    f = FirFilter(code, "COMPLEX" if np.iscomplexobj(taps) else "REAL")
    f.set_taps(taps)
    n = 1 << 27
    big = torch.randint(-30000, 30000, (n + 64, nc), dtype=torch.int16, device="cuda")
    out = torch.empty((n, nc), dtype=torch.int16, device="cuda")
    for off in (0, 1, 3):
        x = big[off: off + n]
        for _ in range(3):
            f.run(x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _, c, p = f.run(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name}: input offset {off} elements ({f.kernel}): {c / ms / 1e3:.0f} Msamples/s", flush=True)
