#!/usr/bin/env python
"""Stream-rate check of the HBM-resident neighbours (b200c_scale / b200c_rotate / b200c_probe / b200c_table_source):
one JSON line per (op, dtype) with Msamples/s and the fraction of the measured HBM copy peak.
Inputs (2^28 elements) are far larger than L2; CUDA-event timing over 10 launches after 3 warm-ups."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pothoscomms_b200 import handles  # noqa: E402
from pothoscomms_b200.handles import dtype_code, ncomp, torch_scalar  # noqa: E402


def main():
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    n = 1 << 28
    for op, dt in (("scale", "complex_float32"), ("scale", "complex_int16"), ("scale", "int16"), ("rotate", "complex_float32"),
                   ("rotate", "complex_int16"), ("probe_rms", "complex_float32"), ("probe_rms", "complex_int16"),
                   ("source", "complex_float32"), ("source", "complex_int16"), ("source_big_table", "complex_float32")):
        code = dtype_code(dt)
        nc, ts = ncomp(code), torch_scalar(code)
        x = (torch.randn((n, nc), device="cuda") * 1000).to(ts)
        out = torch.empty_like(x)
        esz = x.element_size() * nc
        # a source walks a 4096-entry table (staged in shared memory) or a 2^20-entry one (gathered through L2)
        table = x[: (1 << 20) if op == "source_big_table" else 4096].clone()

        def run():
            if op.startswith("source"):
                handles.table_source(code, table, 12345, 123, n, out=out)
                return
            if op == "scale":
                handles.scale(code, 0.37, x, out=out)
            elif op == "rotate":
                handles.rotate(code, 0.7, x, out=out)
            else:
                handles.probe(code, "RMS", x)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        if op.startswith("probe"):      # synchronous call returning a host value: wall clock is the honest figure
            import time
            t0 = time.perf_counter()
            for _ in range(10):
                run()
            ms = (time.perf_counter() - t0) / 10 * 1e3
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
        bytes_per = esz * (1 if op.startswith("probe") or op.startswith("source") else 2)
        gbs = n * bytes_per / (ms * 1e-3) / 1e9
        print(json.dumps({"op": op, "dtype": dt, "elements": n, "ms": ms, "Msamples_per_s": n / (ms * 1e-3) / 1e6,
                          "algorithmic_bytes_per_element": bytes_per, "GBps": gbs, "hbm_peak": peak, "frac": gbs / peak}))
        del x, out


if __name__ == "__main__":
    main()
