"""fir_umma32t_kernel against fir_umma32_kernel over tiles-per-CTA counts, each case in its own process under a
timeout (a deadlock must not take the whole GPU call with it).  Usage: python tools/probe_u32t_tiles.py [case]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case(tpc: int, ntaps: int) -> None:
    import numpy as np
    import torch
    from pothoscomms_b200 import FirFilter
    rng = np.random.default_rng(tpc)
    taps = (rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)) * 0.3 / np.sqrt(ntaps)
    n = tpc * 148 * 3072 + 100
    x = torch.randint(-32768, 32767, (ntaps - 1 + n, 2), dtype=torch.int16, device="cuda")
    outs = {}
    for algo in ("umma32", "umma32t"):
        os.environ["B200C_FIR_ALGO"] = algo
        f = FirFilter(7, "COMPLEX")     # complex_int16 (b200comms.h dtype codes)
        f.set_taps(taps)
        out = torch.zeros((n, 2), dtype=torch.int16, device="cuda")
        y, c, p = f.run(x, out=out, out_capacity=n)
        torch.cuda.synchronize()
        outs[algo] = (f.kernel, out.clone(), p)
    same = bool(torch.equal(outs["umma32"][1], outs["umma32t"][1]))
    print(f"tiles/CTA {tpc} K {ntaps}: {outs['umma32t'][0]} produced {outs['umma32t'][2]} equal={same}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        case(int(sys.argv[1]), int(sys.argv[2]))
        sys.exit(0)
    for tpc, k in ((1, 128), (2, 128), (3, 128), (4, 128), (9, 128), (17, 128), (64, 128), (3, 32), (9, 200)):
        try:
            r = subprocess.run([sys.executable, __file__, str(tpc), str(k)], timeout=60, capture_output=True, text=True,
                               env=dict(os.environ, B200C_UMMA_DBG="1"))
            tail = [l for l in (r.stdout + r.stderr).splitlines() if "tiles/CTA" in l or "watchdog" in l or "Error" in l or "error" in l]
            print(f"[{tpc},{k}] rc={r.returncode}", *tail[:24], sep="\n  ", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{tpc},{k}] TIMEOUT (hang)", flush=True)
