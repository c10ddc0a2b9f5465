#!/usr/bin/env python
"""Small invocations of every kernel family for
`compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_small.py` (the launch-shape knobs are
read once per process, so their other settings run in child processes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pothoscomms_b200 import Fft, FirFilter, handles  # noqa: E402

rng = np.random.default_rng(0)


def fir(dt, tt, ntaps, M, L, n, algo=None):
    if algo:
        os.environ["B200C_FIR_ALGO"] = algo
    else:
        os.environ.pop("B200C_FIR_ALGO", None)
    f = FirFilter(dt, tt)
    taps = rng.standard_normal(ntaps) * 0.2 / np.sqrt(ntaps)
    if tt == "COMPLEX":
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.2 / np.sqrt(ntaps)
    f.set_taps(taps)
    f.set_rates(M, L)
    code = handles.dtype_code(dt)
    nc = handles.ncomp(code)
    ts = handles.torch_scalar(code)
    x = (torch.randn((n, nc), device="cuda") * 1000).to(ts)
    y, c, p = f.run(x)
    torch.cuda.synchronize()
    print(dt, tt, ntaps, M, L, n, f.kernel, c, p)


for algo in ("umma32", "umma", "imma", "direct"):
    fir("complex_int16", "COMPLEX", 128, 1, 1, 20000, algo)
    fir("int16", "REAL", 64, 1, 1, 9001, algo)
fir("complex_int16", "COMPLEX", 255, 2, 3, 30001)
fir("int16", "REAL", 100, 3, 2, 30001)
fir("complex_int16", "REAL", 40, 1, 4, 5000)
fir("complex_float32", "COMPLEX", 256, 1, 1, 30001)
fir("complex_float32", "COMPLEX", 1024, 1, 1, 30001)
fir("complex_float32", "REAL", 255, 2, 3, 30001)
fir("complex_float32", "REAL", 255, 4, 3, 30001)
fir("float32", "REAL", 64, 1, 1, 30001)
fir("float32", "REAL", 600, 1, 1, 30001)
fir("float32", "REAL", 101, 2, 1, 30001)
# the same kernels in their one-warp-CTA forms, and the grouped polyphase kernel the spectral resampler replaced
for var, val in (("B200C_OS32_CFG", "112"), ("B200C_OS32_CFG", "1012"), ("B200C_OS32R_CFG", "112"), ("B200C_OSX_MINB", "12"), ("B200C_OSX", "0")):
    print("spawn", var, val)   # these knobs are read once per process: one child process per setting
    import subprocess
    child = ("import os, sys; sys.path.insert(0, %r); os.environ[%r] = %r; import numpy as np, torch; from pothoscomms_b200 import FirFilter; "
             "f = FirFilter('complex_float32', 'REAL'); f.set_taps(np.hanning(255) / 40); f.set_rates(2, 3); "
             "y, c, p = f.run(torch.randn((30001, 2), device='cuda')); torch.cuda.synchronize(); print(f.kernel, c, p); "
             "g = FirFilter('complex_float32', 'COMPLEX'); g.set_taps(np.hanning(256) / 40); y, c, p = g.run(torch.randn((30001, 2), device='cuda')); "
             "h = FirFilter('float32', 'REAL'); h.set_taps(np.hanning(64) / 10); y, c, p = h.run(torch.randn((30001, 1), device='cuda')); "
             "torch.cuda.synchronize(); print(g.kernel, h.kernel)") % (ROOT, var, val)
    subprocess.run([sys.executable, "-c", child], check=True)
t = (torch.randn((4096, 2), device="cuda")).float()
print(handles.table_source("complex_float32", t, 12345, 123, 100003).shape, handles.table_source("complex_float32", t, 5, 1, 7).shape)
x = (torch.randn((4 * 4096, 2), device="cuda")).float()
Fft("complex_float32", 4096, False).run(x)
xi = (torch.randn((10001, 2), device="cuda") * 1000).to(torch.int16)
handles.scale("complex_int16", 0.5, xi)
handles.rotate("complex_int16", 0.5, xi)
print(handles.probe("complex_int16", "RMS", xi), handles.probe("complex_int16", "MEAN", xi[1:]))
torch.cuda.synchronize()
print("done")
