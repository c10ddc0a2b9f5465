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


for algo in ("umma32t", "umma32", "umma", "imma", "direct"):
    fir("complex_int16", "COMPLEX", 128, 1, 1, 20000, algo)
    fir("int16", "REAL", 64, 1, 1, 9001, algo)
# several tiles per persistent CTA: the accumulator stages, plane stages and landing-ring slots of the tcgen05 kernels are reused
fir("complex_int16", "COMPLEX", 128, 1, 1, 4 * 148 * 3072 + 11, "umma32t")
fir("complex_int16", "REAL", 40, 1, 1, 3 * 148 * 3072 + 5, "umma32t")
fir("int16", "REAL", 64, 1, 1, 3 * 148 * 4096 + 7, "umma32")
fir("int16", "REAL", 64, 1, 1, 3 * 148 * 6144 + 7, "umma32t")      # real data on 64-output windows
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
fir("complex_float32", "COMPLEX", 256, 1, 1, 700)       # a launch smaller than one wave of warps (warp-major spread)
fir("complex_float32", "REAL", 5, 1, 1, 30001)           # <= 8 taps: the direct kernel
fir("complex_int8", "COMPLEX", 21, 1, 2, 2000)           # the round-1 R = 7 / R = 5 mismatch case
fir("complex_float32", "REAL", 64, 3, 1, 30001)          # grouped polyphase kernel
fir("complex_float32", "REAL", 128, 1, 8, 30001)         # general kernel, wide interpolation
# the grouped polyphase kernel the spectral resampler replaced (knob read once per process: a child process)
import subprocess  # noqa: E402
child = ("import os, sys; sys.path.insert(0, %r); os.environ['B200C_OSX'] = '0'; import numpy as np, torch; from pothoscomms_b200 import FirFilter; "
         "f = FirFilter('complex_float32', 'REAL'); f.set_taps(np.hanning(255) / 40); f.set_rates(2, 3); "
         "y, c, p = f.run(torch.randn((30001, 2), device='cuda')); torch.cuda.synchronize(); print(f.kernel, c, p)") % ROOT
subprocess.run([sys.executable, "-c", child], check=True)
# filter bank: one launch over (channel, block) with the 4096-point and the 1024-point kernels
from pothoscomms_b200 import FirFilterBank  # noqa: E402
for ntaps in (1024, 100):
    bank = FirFilterBank("complex_float32", "COMPLEX", 5)
    for c in range(5):
        bank.set_taps(c, (rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)) / ntaps)
    xb = torch.randn((5, ntaps - 1 + 20000, 2), device="cuda")
    ob = torch.empty((5, 20000, 2), device="cuda")
    print("bank", ntaps, bank.run(xb, ob))
# the halo copy kernel (same device: the peer-load path without a second GPU) and the int16 4096-point FFT (lazy wrap)
import ctypes  # noqa: E402
from pothoscomms_b200 import _abi  # noqa: E402
src, dst = torch.arange(2040, device="cuda", dtype=torch.uint8), torch.zeros(2040, device="cuda", dtype=torch.uint8)
_abi.check(_abi.lib().b200c_halo_exchange(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), 2040, 0, None))
torch.cuda.synchronize()
assert torch.equal(src, dst)
xq = (torch.randn((3 * 4096, 2), device="cuda") * 8000).to(torch.int16)
Fft("complex_int16", 4096, False).run(xq)
Fft("complex_int16", 4096, True).run(xq)
Fft("complex_int16", 1000, False).run(xq)
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
