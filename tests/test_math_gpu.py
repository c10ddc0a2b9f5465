"""GPU parity of the HBM-resident neighbours (/comms/scale, /comms/rotate, /comms/signal_probe)
through the C-ABI against the CPU oracle: bit-exact for every integer type (Q-format shift and
wrap), bit-exact for float too (one rounding per operation, no contraction), probe sums within
1e-12 (double accumulation, different order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = ["F32", "CF32", "F64", "CF64", "I8", "CI8", "I16", "CI16", "I32", "CI32", "I64", "CI64"]


def _rand(oracle, code, n, rng):
    sc = oracle.scalar_np(code)
    nc = 2 if code & 1 else 1
    if np.issubdtype(sc, np.integer):
        info = np.iinfo(sc)
        return rng.integers(info.min, info.max, size=(n, nc), endpoint=True).astype(sc)
    return (rng.standard_normal((n, nc)) * 100).astype(sc)


@pytest.mark.parametrize("dt", ALL)
@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_scale_matches_oracle(oracle, cuda_device, dt, n):
    import torch
    from pothoscomms_b200 import handles
    code = getattr(oracle, dt)
    rng = np.random.default_rng(code * 131 + n)
    x = _rand(oracle, code, n, rng)
    for factor in (-1.0, -0.5, 0.0, 0.5, 1.0, 0.37, 3.5, -117.25):        # math/TestScale.cpp:61 and beyond
        y = handles.scale(code, factor, torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(y, oracle.scale(code, factor, x)), (dt, n, factor)


@pytest.mark.parametrize("dt", [t for t in ALL if t.startswith("C")])
@pytest.mark.parametrize("n", [1, 5, 4096, 100003])
def test_rotate_matches_oracle(oracle, cuda_device, dt, n):
    import torch
    from pothoscomms_b200 import handles
    code = getattr(oracle, dt)
    rng = np.random.default_rng(code * 17 + n)
    x = _rand(oracle, code, n, rng)
    for phase in (0.0, np.pi / 2, np.pi, 3 * np.pi / 2, 0.7, -2.1):       # math/TestRotate.cpp:61 and beyond
        y = handles.rotate(code, phase, torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(y, oracle.rotate(code, phase, x)), (dt, n, phase)


def test_unaligned_and_in_place(oracle, cuda_device):
    import torch
    from pothoscomms_b200 import handles
    rng = np.random.default_rng(5)
    x = _rand(oracle, oracle.CI16, 10001, rng)
    d = torch.zeros((10004, 2), dtype=torch.int16, device="cuda")
    for off in (1, 2, 3):
        d[off:off + 10001] = torch.from_numpy(x).cuda()
        view = d[off:off + 10001]
        out = torch.empty((10001 + off, 2), dtype=torch.int16, device="cuda")[off:]
        assert np.array_equal(handles.scale(oracle.CI16, 0.61, view, out=out).cpu().numpy(), oracle.scale(oracle.CI16, 0.61, x))
        assert np.array_equal(handles.rotate(oracle.CI16, 1.1, view, out=out).cpu().numpy(), oracle.rotate(oracle.CI16, 1.1, x))
    d0 = torch.from_numpy(x).cuda()
    handles.scale(oracle.CI16, -0.5, d0, out=d0)                            # in place
    assert np.array_equal(d0.cpu().numpy(), oracle.scale(oracle.CI16, -0.5, x))
    from pothoscomms_b200 import InvalidArgumentError
    with pytest.raises(InvalidArgumentError):                               # rotateFactory: complex types only
        handles.rotate(oracle.I16, 0.5, torch.zeros((4, 1), dtype=torch.int16, device="cuda"))


@pytest.mark.parametrize("dt", ALL)
def test_probe_matches_oracle(oracle, cuda_device, dt):
    import torch
    from pothoscomms_b200 import handles
    code = getattr(oracle, dt)
    rng = np.random.default_rng(code)
    for n in (1, 33, 100003):
        x = _rand(oracle, code, n, rng)
        if code >> 1 >= 4:
            x = (x >> 8).astype(x.dtype)        # keep 32/64-bit squares comparable in double
        d = torch.from_numpy(x).cuda()
        for mode in ("VALUE", "RMS", "MEAN"):
            got, ref = handles.probe(code, mode, d), oracle.probe(code, mode, x)
            assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (dt, n, mode, got, ref)


def test_fir_chain_stays_in_hbm(oracle, cuda_device):
    """source -> scale -> fir_filter -> probe, the topology of filter/TestFIRFilter.cpp:44-78 with
    every block's arithmetic on the device: the RMS the probe reports must be the oracle chain's."""
    import torch
    from pothoscomms_b200 import FirFilter, handles
    from pothoscomms_b200 import workloads as wl
    code = oracle.CI16
    x = wl.tone_noise_numpy(code, 1 << 16, seed=0xC0FFEE02)
    taps, tt = wl.config_taps("c2")
    d = torch.from_numpy(x).cuda()
    d = handles.scale(code, 0.5, d)
    f = FirFilter(code, tt)
    f.set_taps(taps)
    y, cons, prod = f.run(d)
    rms = handles.probe(code, "RMS", y).real
    y_ref, _, _ = oracle.fir(code, True, taps, 1, 1, oracle.scale(code, 0.5, x))
    assert np.array_equal(y.cpu().numpy(), y_ref)
    assert abs(rms - oracle.probe(code, "RMS", y_ref).real) < 1e-9 * rms
    assert rms > 0.1 * 500.0 * 0.5                                          # the reference's own check (:78), scaled
