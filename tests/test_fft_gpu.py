"""GPU parity: /comms/fft CUDA path (through the C-ABI) vs the oracle and the reference's own
compiled kiss_fft (oracle/_ref).  complex int16 is bit-exact (Q15, per-stage 1/radix scaling,
sround); cf32 within 1e-5 of output RMS (north_star), cf64 ~1e-14."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
IN4 = np.array([0.4 + 0.6j, -0.7 + 0.6j, -0.2 + 0.8j, 0.9 + 0.2j])
OUT4 = np.array([0.4 + 2.2j, 1.0 + 1.4j, 0.0 + 0.6j, 0.2 - 1.8j])


def _gpu_fft(code, n, inverse, x):
    import torch
    from pothoscomms_b200 import Fft
    f = Fft(code, n, inverse)
    y = f.run(torch.from_numpy(np.ascontiguousarray(x)).cuda())
    torch.cuda.synchronize()
    return y.cpu().numpy()


def _rel_rms(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return np.sqrt(np.mean((a - b) ** 2) / max(np.mean(b ** 2), 1e-300))


def test_golden_float_n4(oracle, cuda_device):
    """fft/TestFFT.cpp:11-82: numpy golden within 0.01; inverse returns N*input (unnormalised)."""
    y = _gpu_fft(oracle.CF32, 4, False, oracle.to_raw(IN4.astype(np.complex64), oracle.CF32)).view(np.complex64).ravel()
    assert np.all(np.abs(y.real - OUT4.real) < 0.01) and np.all(np.abs(y.imag - OUT4.imag) < 0.01)
    back = _gpu_fft(oracle.CF32, 4, True, oracle.to_raw(OUT4.astype(np.complex64), oracle.CF32)).view(np.complex64).ravel()
    assert np.all(np.abs(back - 4 * IN4) < 0.01)


def test_golden_short_n4(oracle, cuda_device):
    """fft/TestFFT.cpp:84-158: forward == golden/N exactly, inverse(golden) == input exactly."""
    xin = oracle.to_raw(np.round(IN4 * 1000), oracle.CI16)
    res = oracle.to_raw(np.round(OUT4 * 1000), oracle.CI16)
    assert np.array_equal(_gpu_fft(oracle.CI16, 4, False, xin), res // 4)
    assert np.array_equal(_gpu_fft(oracle.CI16, 4, True, res), xin)


SIZES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 25, 27, 30, 32, 49, 60, 64, 97, 100, 121, 125, 128, 242, 255, 256,
         343, 512, 1000, 1001, 1024, 2048, 3072, 4095, 4096, 5005, 8192, 16384, 30030]


@pytest.mark.parametrize("inverse", [False, True])
def test_int16_bit_exact_all_radices(oracle, cuda_device, inverse):
    rng = np.random.default_rng(17)
    for n in SIZES:
        batch = 5 if n <= 4096 else 2
        x = rng.integers(-32768, 32768, size=(batch * n, 2), dtype=np.int16)
        ref = oracle.ref_fft(oracle.CI16, n, inverse, x) if oracle.have_ref() else oracle.fft(oracle.CI16, n, inverse, x)
        y = _gpu_fft(oracle.CI16, n, inverse, x)
        assert np.array_equal(y, ref), f"n={n}"


@pytest.mark.parametrize("inverse", [False, True])
def test_float_within_tolerance_all_radices(oracle, cuda_device, inverse):
    rng = np.random.default_rng(19)
    for n in SIZES:
        batch = 5 if n <= 4096 else 2
        x = rng.standard_normal((batch * n, 2)).astype(np.float32)
        ref = oracle.ref_fft(oracle.CF32, n, inverse, x) if oracle.have_ref() else oracle.fft(oracle.CF32, n, inverse, x)
        y = _gpu_fft(oracle.CF32, n, inverse, x)
        assert _rel_rms(y, ref) < 1e-5, f"cf32 n={n}: {_rel_rms(y, ref):.3e}"
        assert np.max(np.abs(y - ref)) < 1e-4 * np.sqrt(np.mean(ref.astype(np.float64) ** 2)) + 1e-6, f"cf32 n={n}"
        xd = x.astype(np.float64)
        refd = oracle.fft(oracle.CF64, n, inverse, xd)
        yd = _gpu_fft(oracle.CF64, n, inverse, xd)
        assert _rel_rms(yd, refd) < 1e-13, f"cf64 n={n}: {_rel_rms(yd, refd):.3e}"


def test_batches_and_ragged_tail(oracle, cuda_device):
    """floor(len/numBins) transforms; a trailing partial frame is left untouched (the block's reserve, fft/FFT.cpp:50)."""
    import torch
    from pothoscomms_b200 import Fft
    rng = np.random.default_rng(23)
    for n, batch in ((64, 1000), (1024, 257), (4096, 33)):
        x = rng.standard_normal((batch * n + n // 2, 2)).astype(np.float32)
        f = Fft(oracle.CF32, n, False)
        out = torch.full((batch * n + n // 2, 2), 7.0, dtype=torch.float32, device="cuda")
        f.run(torch.from_numpy(x).cuda(), out=out)
        y = out.cpu().numpy()
        ref = oracle.fft(oracle.CF32, n, False, x)
        assert _rel_rms(y[: batch * n], ref) < 1e-5
        assert np.all(y[batch * n:] == 7.0)


def test_committed_reference_fixtures(oracle, cuda_device):
    """tests/golden/fft_*.npz: outputs of the reference's own sources (make_golden.py)."""
    n_checked = 0
    for fn in sorted(os.listdir(GOLDEN)):
        if not (fn.startswith("fft_") and fn.endswith(".npz")):
            continue
        g = np.load(os.path.join(GOLDEN, fn))
        dt, n, inv = int(g["dtype"]), int(g["n"]), bool(g["inverse"])
        y = _gpu_fft(dt, n, inv, g["x"])
        if dt == oracle.CI16:
            assert np.array_equal(y, g["y"]), fn
        else:
            assert _rel_rms(y, g["y"]) < (1e-5 if dt == oracle.CF32 else 1e-13), fn
        n_checked += 1
    assert n_checked >= 20


def test_forward_inverse_roundtrip_full_size(oracle, cuda_device):
    """BASELINE config 4: 4096-point fwd then inv over 2^26 samples: ifft(fft(x)) = N*x (float),
    and sampled transforms against the oracle."""
    import torch
    from pothoscomms_b200 import Fft
    from pothoscomms_b200 import workloads as wl
    n, total = 4096, 1 << 26
    x = wl.tone_noise_torch(oracle.CF32, total, 0xC0FFEE04, cuda_device)
    fwd, inv = Fft(oracle.CF32, n, False), Fft(oracle.CF32, n, True)
    X = fwd.run(x)
    back = inv.run(X)
    err = torch.sqrt(torch.mean((back / n - x).double() ** 2)).item()
    rms = torch.sqrt(torch.mean(x.double() ** 2)).item()
    assert err < 2e-6 * rms
    for b in (0, 1234, total // n - 1):
        seg = x[b * n: (b + 1) * n].cpu().numpy()
        ref = oracle.ref_fft(oracle.CF32, n, False, seg) if oracle.have_ref() else oracle.fft(oracle.CF32, n, False, seg)
        assert _rel_rms(X[b * n: (b + 1) * n].cpu().numpy(), ref) < 1e-5


def test_int16_full_size_sampled(oracle, cuda_device):
    import torch
    from pothoscomms_b200 import Fft
    from pothoscomms_b200 import workloads as wl
    n, total = 4096, 1 << 24
    x = wl.tone_noise_torch(oracle.CI16, total, 0xC0FFEE14, cuda_device)
    for inverse in (False, True):
        y = Fft(oracle.CI16, n, inverse).run(x)
        for b in (0, 777, total // n - 1):
            seg = x[b * n: (b + 1) * n].cpu().numpy()
            ref = oracle.ref_fft(oracle.CI16, n, inverse, seg) if oracle.have_ref() else oracle.fft(oracle.CI16, n, inverse, seg)
            assert np.array_equal(y[b * n: (b + 1) * n].cpu().numpy(), ref)


def test_host_buffer_entry_point(oracle, cuda_device):
    rng = np.random.default_rng(29)
    from pothoscomms_b200 import Fft
    x = rng.standard_normal((3000 * 4096, 2)).astype(np.float32)   # > one 32 MiB chunk
    y = Fft(oracle.CF32, 4096, False).run_host(x)
    for b in (0, 1500, 2999):
        ref = oracle.fft(oracle.CF32, 4096, False, x[b * 4096: (b + 1) * 4096])
        assert _rel_rms(y[b * 4096: (b + 1) * 4096], ref) < 1e-5


def test_factory_errors(oracle, cuda_device):
    from pothoscomms_b200 import Fft, InvalidArgumentError
    for bad in ("float32", "complex_int32", "int16", "complex_int8"):
        with pytest.raises(InvalidArgumentError, match="unsupported type"):   # fft/FFT.cpp:92
            Fft(bad, 1024, False)
    with pytest.raises(InvalidArgumentError):
        Fft("complex_float32", 0, False)


def test_large_transform_global_path(oracle, cuda_device):
    """numBins too large for shared memory: working buffer lives in the output slab."""
    rng = np.random.default_rng(31)
    for n in (1 << 16, 3 * 5 * 7 * 11 * 13 * 4, 2 * 13001):   # 13001 is prime: generic butterfly + global scratch
        x = rng.standard_normal((2 * n, 2)).astype(np.float32)
        ref = oracle.fft(oracle.CF32, n, False, x)
        y = _gpu_fft(oracle.CF32, n, False, x)
        assert _rel_rms(y, ref) < 1e-5, n
        xi = rng.integers(-32768, 32768, size=x.shape, dtype=np.int16)
        assert np.array_equal(_gpu_fft(oracle.CI16, n, True, xi), oracle.fft(oracle.CI16, n, True, xi)), n
