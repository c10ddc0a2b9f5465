"""Generates tests/golden/*.npz.  Run from the repo root in the BUILD container, where
/root/reference exists:   python tests/golden/make_golden.py

fft_*.npz : inputs + outputs of the REFERENCE's own kiss_fft sources (oracle/_ref, compiled
            from /root/reference/fft by oracle/Makefile) -- reference-pinned vectors.
fir_*.npz : inputs + outputs of the REFERENCE's own /comms/fir_filter block (oracle/_ref/libfirref.so =
            /root/reference/filter/FIRFilter.cpp compiled unmodified, oracle/Makefile ref_fir) -- reference-pinned
            up to the one recalled external header (oracle/ref_include/Pothos/Util/QFormat.hpp).
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
    # FIR vectors from the reference block itself (one work() call, plus a burst flushed via a frame-end label)
    assert oracle.have_ref_fir(), "needs oracle/_ref/libfirref.so (reference checkout)"
    cases = [
        ("cf32_cc_64", oracle.CF32, True, 64, 1, 1),
        ("cf32_cr_255_l3m2", oracle.CF32, False, 255, 2, 3),
        ("ci16_cc_128", oracle.CI16, True, 128, 1, 1),
        ("ci16_cc_101_l2m3", oracle.CI16, True, 101, 3, 2),
        ("f32_rr_33_m2", oracle.F32, False, 33, 2, 1),
        ("i16_rr_17_l3", oracle.I16, False, 17, 1, 3),
        ("ci16_cr_48_m3", oracle.CI16, False, 48, 3, 1),
        ("ci8_cc_21_l2", oracle.CI8, True, 21, 1, 2),
        ("i32_rr_40", oracle.I32, False, 40, 1, 1),
        ("cf64_cc_77_l3m2", oracle.CF64, True, 77, 2, 3),
    ]
    for name, dt, tcx, ntaps, M, L in cases:
        taps = rng.standard_normal(ntaps) * 0.1
        if tcx:
            taps = taps + 1j * rng.standard_normal(ntaps) * 0.1
        nc = 2 if dt & 1 else 1
        sc = oracle.scalar_np(dt)
        if np.issubdtype(sc, np.integer):
            lim = min(np.iinfo(sc).max, 20000 if sc == np.int16 else 2 ** 31 - 1)
            x = rng.integers(-lim, lim, size=(2000, nc)).astype(sc)
        else:
            x = rng.standard_normal((2000, nc)).astype(sc)
        y, cons, prod = oracle.ref_fir(dt, tcx, taps, M, L, x)
        yb, cb, pb, _ = oracle.ref_fir_stream(dt, tcx, taps, M, L, x[:500], frame_end=True)
        tr = np.asarray(taps, dtype=np.complex128 if tcx else np.float64)
        np.savez_compressed(os.path.join(OUT, f"fir_{name}.npz"), dtype=dt, taps_complex=int(tcx), taps=tr, M=M, L=L,
                            x=x, y=y, consumed=cons, produced=prod, burst_y=yb, burst_consumed=cb, burst_produced=pb,
                            source="reference:filter/FIRFilter.cpp")
    print("wrote", len([f for f in os.listdir(OUT) if f.endswith('.npz')]), "fixtures")


if __name__ == "__main__":
    main()
