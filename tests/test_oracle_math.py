"""Oracle pins for the HBM-resident neighbours: /comms/scale,
/comms/rotate, /comms/signal_probe.  The reference's own tests for these blocks are re-run against
the oracle: math/TestScale.cpp:14-70 and math/TestRotate.cpp:14-71 (every type, their factors and
phases, result within 1 of Type(input * factor)) -- the only in-tree tests that touch the Q-format
helpers.  They pin the shift count (factor 1.0 and phase 0 must be the identity) but, all their
cases being exact, not the rounding direction (oracle/qformat.h)."""
import numpy as np
import pytest

NUM_POINTS = 13   # math/TestScale.cpp:12, math/TestRotate.cpp:12

TYPES = ["F64", "F32", "I64", "I32", "I16", "I8"]


def _trunc_cast(v, sc):
    """C++ Type(double): truncation toward zero for integer types"""
    return np.trunc(v).astype(sc) if np.issubdtype(sc, np.integer) else v.astype(sc)


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("i", range(5))
def test_scale_reference_cases(oracle, dt, i):
    """math/TestScale.cpp:59-70: factor = i/2 - 1, input 10*i"""
    code = getattr(oracle, dt)
    sc = oracle.scalar_np(code)
    factor = i / 2.0 - 1.0
    x = (10 * np.arange(NUM_POINTS)).astype(sc)
    y = oracle.scale(code, factor, x)
    expected = _trunc_cast(x.astype(np.float64) * factor, sc)
    assert np.all(np.abs(y.astype(np.float64) - expected.astype(np.float64)) <= 1)   # POTHOS_TEST_CLOSE(..., 1)
    if factor == 1.0:
        assert np.array_equal(y, x)          # pins the shift count: in * 2^n >> n == in


@pytest.mark.parametrize("dt", TYPES)
@pytest.mark.parametrize("i", range(4))
def test_rotate_reference_cases(oracle, dt, i):
    """math/TestRotate.cpp:59-71: phase = i*pi/2, input (10 f, -20 f)"""
    code = getattr(oracle, "C" + dt)
    sc = oracle.scalar_np(code)
    phase = i * np.pi / 2
    f = np.arange(NUM_POINTS)
    x = np.stack([10 * f, -20 * f], axis=1).astype(sc)
    y = oracle.rotate(code, phase, x)
    z = (x[:, 0].astype(np.float64) + 1j * x[:, 1].astype(np.float64)) * np.exp(1j * phase)
    expected = np.stack([_trunc_cast(z.real, sc), _trunc_cast(z.imag, sc)], axis=1)
    assert np.all(np.abs(y.astype(np.float64) - expected.astype(np.float64)) <= 1)
    if i == 0:
        assert np.array_equal(y, x)


def test_scale_complex_and_wrapping(oracle):
    """complex data: the real factor scales both parts (Scale.cpp:149); integer products wrap in the Q type"""
    rng = np.random.default_rng(3)
    x = rng.integers(-32768, 32767, size=(1000, 2)).astype(np.int16)
    y = oracle.scale(oracle.CI16, 0.37, x)
    fq = int(np.trunc(0.37 * 65536))
    ref = ((x.astype(np.int64) * fq).astype(np.int32) >> 16).astype(np.int16)      # int32 Q, >> 16
    assert np.array_equal(y, ref)
    y = oracle.scale(oracle.I16, 3.5, x[:, 0])                                     # 3.5 * 2^16 * 30000 wraps int32
    ref = ((x[:, 0].astype(np.int64) * int(3.5 * 65536)).astype(np.int32) >> 16).astype(np.int16)
    assert np.array_equal(y, ref)


def test_rotate_int16_matches_bigint_model(oracle):
    rng = np.random.default_rng(4)
    x = rng.integers(-32768, 32767, size=(1000, 2)).astype(np.int16)
    phase = 0.7
    pr, pi = int(np.trunc(np.cos(phase) * 65536)), int(np.trunc(np.sin(phase) * 65536))
    a, b = x[:, 0].astype(np.int64), x[:, 1].astype(np.int64)
    re = ((pr * a - pi * b).astype(np.int32) >> 16).astype(np.int16)
    im = ((pr * b + pi * a).astype(np.int32) >> 16).astype(np.int16)
    assert np.array_equal(oracle.rotate(oracle.CI16, phase, x), np.stack([re, im], axis=1))


@pytest.mark.parametrize("dt", ["F32", "CF32", "I16", "CI16", "CF64", "I8"])
def test_probe_modes(oracle, dt):
    """utility/SignalProbe.cpp:140-160"""
    code = getattr(oracle, dt)
    sc = oracle.scalar_np(code)
    nc = 2 if code & 1 else 1
    rng = np.random.default_rng(code)
    x = (rng.standard_normal((777, nc)) * 50).astype(sc)
    xc = x[:, 0].astype(np.float64) + (1j * x[:, 1].astype(np.float64) if nc == 2 else 0)
    assert oracle.probe(code, "VALUE", x) == complex(xc[-1])
    assert abs(oracle.probe(code, "RMS", x).real - np.sqrt(np.mean(np.abs(xc) ** 2))) < 1e-9 * max(1.0, np.abs(xc).max())
    assert abs(oracle.probe(code, "MEAN", x) - np.mean(xc)) < 1e-9 * max(1.0, np.abs(xc).max())
