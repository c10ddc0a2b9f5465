"""Generates tests/golden/source_*.npz: regression vectors of the stream-source oracle (oracle/source_oracle.cpp).
Run from the repo root:   python tests/golden/make_source_golden.py

source_waveform.npz : tables, steps and 3000-element streams of WaveformSource::updateTable() / work() for a set of
                      (type, wave, frequency, resolution, amplitude, offset) cases -- NOT reference-pinned (the block
                      needs PothosCore); closed forms are checked separately in tests/test_oracle_source.py.
source_noise.npz    : pools and streams of a NoiseSource seeded with a fixed seed -- libstdc++ <random> output, so these
                      vectors are tied to the toolchain of this image (GCC's libstdc++).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

WAVE_CASES = [
    ("cf32_sine", "CF32", "SINE", 30e3, 1e6, 0.0, 1000.0, 0.0),
    ("ci16_sine", "CI16", "SINE", 30e3, 1e6, 0.0, 1000.0, 0.0),
    ("cf64_ramp", "CF64", "RAMP", 1e3, 1e6, 0.0, 2 - 1j, 0.5j),
    ("i8_square", "I8", "SQUARE", -45e3, 1e6, 0.0, 100.0, -3.0),
    ("f32_sine_res", "F32", "SINE", 1e3, 1e6, 10.0, 1.0, 0.25),
    ("ci64_const", "CI64", "CONST", 0.0, 1.0, 0.0, 7 + 2j, 1.0),
]
NOISE_CASES = [
    ("cf32_normal", "CF32", "NORMAL", 0.0, 1.0, 1.0, 0.0),
    ("ci16_uniform", "CI16", "UNIFORM", 0.5, 2.0, 1000.0, 1 - 1j),
    ("f64_laplace", "F64", "LAPLACE", 0.0, 0.9, 1.0, 0.0),
    ("ci8_poisson", "CI8", "POISSON", 4.0, 1.0, 3.0, 0.0),
]
NOISE_SEED, NOISE_WORK = 0xB200C0DE, [1000, 4096 + 17, 300]


def main():
    oracle.build()
    wave, noise = {}, {}
    for name, dt, kind, freq, rate, res, ampl, off in WAVE_CASES:
        code = getattr(oracle, dt)
        table, step = oracle.waveform_table(code, kind, freq, rate, res=res, ampl=ampl, offset=off)
        wave[name + "_table_head"] = table[:64]
        wave[name + "_entries_step"] = np.array([table.shape[0], step], dtype=np.uint64)
        wave[name + "_stream"] = oracle.table_walk(code, table, 12345, step, 3000)
    for name, dt, kind, mean, b, ampl, off in NOISE_CASES:
        code = getattr(oracle, dt)
        stream, pool = oracle.noise_stream(code, kind, mean, b, NOISE_SEED, NOISE_WORK, refill_before=[0, 0, 1], ampl=ampl, offset=off)
        noise[name + "_pool"] = pool[:256]
        noise[name + "_stream"] = np.concatenate([stream[:256], stream[-300:]])   # first work() and the one after the refill
    np.savez_compressed(os.path.join(OUT, "source_waveform.npz"), **wave)
    np.savez_compressed(os.path.join(OUT, "source_noise.npz"), **noise)
    print("wrote", len(wave), "+", len(noise), "arrays")


if __name__ == "__main__":
    main()
