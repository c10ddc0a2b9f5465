"""Generates tests/golden/*.npz.  Run from the repo root in the BUILD container, where
/root/reference exists:   python tests/golden/make_golden.py

fft_*.npz : inputs + outputs of the REFERENCE's own kiss_fft sources (oracle/_ref, compiled
            from /root/reference/fft by oracle/Makefile) -- reference-pinned vectors.
fir_*.npz : inputs + outputs of the FIR oracle restatement (the reference FIR block cannot be
            compiled without PothosCore) -- regression vectors, NOT reference-pinned.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    oracle.build()
    assert oracle.have_ref(), "needs oracle/_ref (reference checkout)"
    rng = np.random.default_rng(20261017)
    for dt, name, sizes in ((oracle.CF32, "cf32", (4, 60, 1024, 4096)), (oracle.CI16, "ci16", (4, 60, 1024, 4096)),
                            (oracle.CF64, "cf64", (64, 1000))):
        for n in sizes:
            for inv in (0, 1):
                if dt == oracle.CI16:
                    x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
                else:
                    x = rng.standard_normal((n, 2)).astype(oracle.scalar_np(dt))
                y = oracle.ref_fft(dt, n, bool(inv), x)
                np.savez_compressed(os.path.join(OUT, f"fft_{name}_{n}_{'inv' if inv else 'fwd'}.npz"),
                                    dtype=dt, n=n, inverse=inv, x=x, y=y)
    # FIR regression vectors (oracle restatement)
    cases = [
        ("cf32_cc_64", oracle.CF32, True, 64, 1, 1),
        ("cf32_cr_255_l3m2", oracle.CF32, False, 255, 2, 3),
        ("ci16_cc_128", oracle.CI16, True, 128, 1, 1),
        ("ci16_cc_101_l2m3", oracle.CI16, True, 101, 3, 2),
        ("f32_rr_33_m2", oracle.F32, False, 33, 2, 1),
        ("i16_rr_17_l3", oracle.I16, False, 17, 1, 3),
    ]
    for name, dt, tcx, ntaps, M, L in cases:
        taps = rng.standard_normal(ntaps) * 0.1
        if tcx:
            taps = taps + 1j * rng.standard_normal(ntaps) * 0.1
        nc = 2 if dt & 1 else 1
        if dt in (oracle.I16, oracle.CI16):
            x = rng.integers(-20000, 20000, size=(2000, nc), dtype=np.int16)
        else:
            x = rng.standard_normal((2000, nc)).astype(np.float32)
        y, cons, prod = oracle.fir(dt, tcx, taps, M, L, x)
        tr = np.asarray(taps, dtype=np.complex128 if tcx else np.float64)
        np.savez_compressed(os.path.join(OUT, f"fir_{name}.npz"), dtype=dt, taps_complex=int(tcx), taps=tr, M=M, L=L,
                            x=x, y=y, consumed=cons, produced=prod)
    print("wrote", len([f for f in os.listdir(OUT) if f.endswith('.npz')]), "fixtures")


if __name__ == "__main__":
    main()
