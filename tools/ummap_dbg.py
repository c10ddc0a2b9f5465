#!/usr/bin/env python
"""Role timers of fir_ummap_kernel over a grid of rates / tap counts (run with B200C_UMMA_DBG=1):
prints what the library reports per launch, prefixed by the configuration, to see how the cycles
per MMA depend on N (columns), NB (k-blocks) and M (residue planes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pothoscomms_b200 import FirFilter  # noqa: E402


def main():
    n = 1 << 26
    for dtype, nc in (("complex_int16", 2), ("int16", 1)):
        x = torch.randint(-20000, 20000, (n, nc), dtype=torch.int16, device="cuda")
        for L, M, ntaps in ((1, 2, 64), (2, 1, 128), (2, 3, 128), (3, 2, 255), (3, 2, 96), (3, 2, 510), (4, 3, 256), (4, 1, 128), (3, 4, 255)):
            taps = np.hanning(ntaps) * 0.4 / max(1.0, ntaps / (8.0 * L))
            f = FirFilter(dtype, "REAL")
            f.set_taps(taps)
            f.set_rates(M, L)
            sys.stderr.write(f"{dtype} L={L} M={M} ntaps={ntaps} kernel={f.kernel}\n")
            sys.stderr.flush()
            f.run(x)
            torch.cuda.synchronize()
            f.run(x)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
