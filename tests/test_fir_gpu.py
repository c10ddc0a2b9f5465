"""GPU parity: /comms/fir_filter CUDA path (through the C-ABI) vs the CPU oracle.

Bit-exact for every integer type (the reference's Q-format shift and wrap); within 1e-5 of
the output RMS for float32/complex float32 (BASELINE.json north_star), tighter for float64.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FLOAT_TOL = 1e-5   # of output RMS, float32 paths (north_star)
DOUBLE_TOL = 1e-13


def _rand_input(oracle, code, n, rng, full_scale=False):
    nc = 2 if code & 1 else 1
    sc = oracle.scalar_np(code)
    if np.issubdtype(sc, np.integer):
        info = np.iinfo(sc)
        lim = info.max if full_scale else info.max // 2
        lim = min(lim, 2 ** 31 - 1)
        return rng.integers(-lim - 1, lim + 1, size=(n, nc)).astype(sc)
    return rng.standard_normal((n, nc)).astype(sc)


def _run_gpu(code, taps_type, taps, M, L, x, zero_tail=False, out_capacity=None):
    import torch
    from pothoscomms_b200 import FirFilter
    f = FirFilter(code, taps_type)
    f.set_taps(taps)
    f.set_rates(M, L)
    d = torch.from_numpy(x).cuda()
    y, cons, prod = f.run(d, zero_tail=zero_tail, out_capacity=out_capacity)
    torch.cuda.synchronize()
    return y.cpu().numpy(), cons, prod, f


def _compare(oracle, code, y_gpu, y_ref, what="", rms_hint=None):
    """rms_hint: the stream's expected output RMS, used as the denominator when the compared
    window is too short (< 256 values) for its own RMS to be a meaningful statistic."""
    assert y_gpu.shape == y_ref.shape, what
    sc = oracle.scalar_np(code)
    if np.issubdtype(sc, np.integer):
        assert np.array_equal(y_gpu, y_ref), f"{what}: integer path must be bit-exact"
    else:
        ref = y_ref.astype(np.float64)
        rms = np.sqrt(np.mean(ref ** 2)) if ref.size else 1.0
        if rms_hint is not None and ref.size < 256:
            rms = max(rms, rms_hint)
        err = np.sqrt(np.mean((y_gpu.astype(np.float64) - ref) ** 2)) if ref.size else 0.0
        tol = FLOAT_TOL if sc == np.float32 else DOUBLE_TOL
        assert err <= tol * max(rms, 1e-30), f"{what}: rel rms err {err / max(rms, 1e-30):.3e} > {tol}"
        mx = np.max(np.abs(y_gpu.astype(np.float64) - ref)) if ref.size else 0.0
        assert mx <= 50 * tol * max(rms, 1e-30), f"{what}: max err {mx}"


ALL_TYPES = ["F32", "CF32", "F64", "CF64", "I8", "CI8", "I16", "CI16", "I32", "CI32", "I64", "CI64"]


@pytest.mark.parametrize("dt", ALL_TYPES)
@pytest.mark.parametrize("taps_type", ["REAL", "COMPLEX"])
def test_all_18_factory_rows(oracle, cuda_device, dt, taps_type):
    """Every row of FIRFilterFactory (filter/FIRFilter.cpp:373-382), streaming L=M=1."""
    code = getattr(oracle, dt)
    if taps_type == "COMPLEX" and not (code & 1):
        from pothoscomms_b200 import FirFilter, InvalidArgumentError
        with pytest.raises(InvalidArgumentError):   # filter/FIRFilter.cpp:383
            FirFilter(code, "COMPLEX")
        return
    rng = np.random.default_rng(code * 2 + (taps_type == "COMPLEX"))
    ntaps = 37
    taps = rng.standard_normal(ntaps) * 0.2
    if taps_type == "COMPLEX":
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.2
    x = _rand_input(oracle, code, 5000, rng)
    y_ref, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps, 1, 1, x)
    y, cons, prod, _ = _run_gpu(code, taps_type, taps, 1, 1, x)
    assert (cons, prod) == (c_ref, p_ref)
    _compare(oracle, code, y, y_ref, f"{dt}/{taps_type}")


RATES = [(1, 1), (2, 1), (3, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3), (5, 4), (4, 6), (7, 3), (16, 1), (1, 16)]


@pytest.mark.parametrize("M,L", RATES)
@pytest.mark.parametrize("dt,taps_type", [("CF32", "COMPLEX"), ("CF32", "REAL"), ("F32", "REAL"), ("CI16", "COMPLEX"),
                                          ("CI16", "REAL"), ("I16", "REAL")])
def test_polyphase_rates(oracle, cuda_device, dt, taps_type, M, L):
    """decim x interp grid of filter/TestFIRFilter.cpp:68-70 (1..3 x 1..3) and beyond; 101 taps as :40."""
    code = getattr(oracle, dt)
    rng = np.random.default_rng(1000 + M * 31 + L)
    ntaps = 101
    taps = rng.standard_normal(ntaps) * 0.1
    if taps_type == "COMPLEX":
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.1
    x = _rand_input(oracle, code, 4096 + 200, rng)
    y_ref, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps, M, L, x)
    y, cons, prod, _ = _run_gpu(code, taps_type, taps, M, L, x)
    assert (cons, prod) == (c_ref, p_ref)
    _compare(oracle, code, y, y_ref, f"{dt}/{taps_type} M={M} L={L}")


@pytest.mark.parametrize("ntaps", [1, 2, 3, 8, 9, 10, 63, 64, 65, 128, 255, 256, 257, 1024])
def test_tap_counts(oracle, cuda_device, ntaps):
    rng = np.random.default_rng(ntaps)
    taps = (rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)) / np.sqrt(ntaps)
    for code, tt in ((oracle.CF32, "COMPLEX"), (oracle.CI16, "COMPLEX")):
        x = _rand_input(oracle, code, ntaps + 3000, rng)
        y_ref, c_ref, p_ref = oracle.fir(code, True, taps, 1, 1, x)
        y, cons, prod, _ = _run_gpu(code, tt, taps, 1, 1, x)
        assert (cons, prod) == (c_ref, p_ref)
        _compare(oracle, code, y, y_ref, f"ntaps={ntaps} code={code}")


def test_baseline_configs_small(oracle, cuda_device):
    """The BASELINE.json configs at sizes the oracle finishes in seconds (same taps, tone+noise)."""
    from pothoscomms_b200 import workloads as wl
    cases = [("c1", oracle.CF32, 1, 1, 1 << 18), ("c1_real", oracle.CF32, 1, 1, 1 << 18),
             ("headline", oracle.CF32, 1, 1, 1 << 17), ("c2", oracle.CI16, 1, 1, 1 << 18),
             ("c3", oracle.CF32, 2, 3, 1 << 18), ("c5", oracle.CF32, 1, 1, 1 << 15)]
    for name, code, M, L, n in cases:
        taps, tt = wl.config_taps(name)
        x = wl.tone_noise_numpy(code, n, seed=0xC0FFEE00 + len(name))
        y_ref, c_ref, p_ref = oracle.fir(code, tt == "COMPLEX", taps, M, L, x, threads=8)
        y, cons, prod, _ = _run_gpu(code, tt, taps, M, L, x)
        assert (cons, prod) == (c_ref, p_ref), name
        _compare(oracle, code, y, y_ref, name)


def test_int16_wrapping_accumulator(oracle, cuda_device):
    """Deliberately overflowing int32 accumulator must wrap exactly like the reference's QType."""
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal(128) + 1j * rng.standard_normal(128)) * 1.9
    x = _rand_input(oracle, oracle.CI16, 20000, rng, full_scale=True)
    y_ref, _, _ = oracle.fir(oracle.CI16, True, taps, 1, 1, x)
    y, _, _, _ = _run_gpu(oracle.CI16, "COMPLEX", taps, 1, 1, x)
    assert np.array_equal(y, y_ref)


def test_default_taps_passthrough(oracle, cuda_device):
    import torch
    from pothoscomms_b200 import FirFilter
    for dt in ("complex_float32", "complex_int16", "float32", "int16", "complex_float64", "complex_int8"):
        f = FirFilter(dt, "REAL")
        assert f.info() == (1, 1, 1, 1)
        code = f.dtype
        x = _rand_input(oracle, code, 1000, np.random.default_rng(3))
        y, cons, prod = f.run(torch.from_numpy(x).cuda())
        assert cons == prod == 1000
        assert np.array_equal(y.cpu().numpy(), x)


def test_setter_errors_and_state(oracle, cuda_device):
    from pothoscomms_b200 import FirFilter, InvalidArgumentError
    f = FirFilter("complex_float32", "COMPLEX")
    with pytest.raises(InvalidArgumentError, match="taps cannot be empty"):       # FIRFilter.cpp:140
        f.set_taps([])
    with pytest.raises(InvalidArgumentError, match="decimation cannot be 0"):     # FIRFilter.cpp:153
        f.set_rates(0, 1)
    with pytest.raises(InvalidArgumentError, match="interpolation cannot be 0"):  # FIRFilter.cpp:165
        f.set_rates(1, 0)
    # a failed setter leaves the previous configuration intact
    assert f.info() == (1, 1, 1, 1)
    f.set_taps(np.ones(101, dtype=complex))
    f.set_rates(2, 3)
    assert f.info() == (34, 2 + 34 - 1, 2, 3)    # K = ceil(101/3), _inputRequire = M + K - 1


def test_output_capacity_limits_N(oracle, cuda_device):
    rng = np.random.default_rng(5)
    taps = rng.standard_normal(33)
    x = _rand_input(oracle, oracle.CF32, 3000, rng)
    for cap in (0, 1, 2, 3, 10, 100, 1001):
        y_ref, c_ref, p_ref = oracle.fir(oracle.CF32, False, taps, 2, 3, x, out_capacity=cap)
        y, cons, prod, _ = _run_gpu(oracle.CF32, "REAL", taps, 2, 3, x, out_capacity=cap)
        assert (cons, prod) == (c_ref, p_ref), cap
        _compare(oracle, oracle.CF32, y, y_ref, f"cap={cap}")


def test_insufficient_input(oracle, cuda_device):
    y, cons, prod, f = _run_gpu(oracle.CF32, "REAL", np.ones(64), 1, 1, np.ones((63, 2), dtype=np.float32))
    assert (cons, prod) == (0, 0) and f.input_require == 64


@pytest.mark.parametrize("M,L", [(1, 1), (2, 3), (3, 2)])
def test_burst_zero_tail(oracle, cuda_device, M, L):
    """Burst flush (filter/FIRFilter.cpp:265-272): K-1 virtual zeros, never materialised on the device."""
    rng = np.random.default_rng(8)
    taps = rng.standard_normal(101) + 1j * rng.standard_normal(101)
    for code in (oracle.CF32, oracle.CI16):
        x = _rand_input(oracle, code, 1024, rng)
        y_ref, c_ref, p_ref = oracle.fir(code, True, taps * 0.1, M, L, x, zero_tail=True)
        y, cons, prod, _ = _run_gpu(code, "COMPLEX", taps * 0.1, M, L, x, zero_tail=True)
        assert (cons, prod) == (c_ref, p_ref)
        _compare(oracle, code, y, y_ref, f"zero tail code={code}")
    if (M, L) == (1, 1):
        assert prod == 1024   # filter/TestFIRDesigner.cpp:183


def test_streaming_in_pieces_equals_one_shot(oracle, cuda_device):
    """work() is stateless in the history: feeding consecutive windows that each start K-1
    elements before the new data reproduces the single-call result (FIRFilter.cpp:304-307).
    Integer streams: bit for bit.  Float streams on the fused overlap-save path: the transform
    block partition depends on where a call starts, so pieces agree to rounding (2e-6 of RMS,
    well inside the 1e-5 parity bar), and bit for bit on the direct kernel."""
    import torch
    from pothoscomms_b200 import FirFilter
    rng = np.random.default_rng(21)
    taps = rng.standard_normal(255)
    for code in (oracle.CF32, oracle.CI16):
        x = _rand_input(oracle, code, 50000, rng)
        f = FirFilter(code, "REAL")
        f.set_taps(taps * (1.0 if code == oracle.CF32 else 0.01))
        f.set_rates(2, 3)
        d = torch.from_numpy(x).cuda()
        whole, cons_all, _ = f.run(d)
        pos, outs = 0, []
        for piece in (1000, 37, 4096, 2, 10000, 12345, 50000):
            avail = min(piece, x.shape[0] - pos)
            y, cons, prod = f.run(d[pos: pos + avail].contiguous())
            outs.append(y.clone())
            pos += cons           # K-1 (+ remainder) elements stay in the "input buffer"
        got = torch.cat(outs)
        assert pos == cons_all and got.shape[0] == whole.shape[0]
        if code == oracle.CI16 or not f.kernel.startswith("fir_os"):
            assert torch.equal(got, whole)
        else:
            a, b = got.double().cpu().numpy(), whole.double().cpu().numpy()
            assert np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2)) < 2e-6


def test_host_buffer_entry_point(oracle, cuda_device):
    """b200c_fir_run_host: chunked H2D -> kernel -> D2H gives the same stream."""
    rng = np.random.default_rng(13)
    from pothoscomms_b200 import FirFilter
    for code, tt, M, L, n in ((oracle.CF32, "COMPLEX", 1, 1, 9_000_000), (oracle.CI16, "COMPLEX", 3, 2, 3_000_000)):
        taps = (rng.standard_normal(64) + 1j * rng.standard_normal(64)) * 0.05
        x = _rand_input(oracle, code, n, rng)
        f = FirFilter(code, tt)
        f.set_taps(taps)
        f.set_rates(M, L)
        y_host, cons, prod = f.run_host(x)
        import torch
        y_dev, c2, p2 = f.run(torch.from_numpy(x).cuda())
        assert (cons, prod) == (c2, p2)
        assert np.array_equal(y_host, y_dev.cpu().numpy())
        # spot-check against the oracle on a window
        w0 = 1_234_567 // M * M
        seg = x[w0: w0 + 20000]
        y_ref, _, p_ref = oracle.fir(code, True, taps, M, L, seg)
        got = y_host[w0 // M * L: w0 // M * L + p_ref]
        _compare(oracle, code, got, y_ref, "host path window")


def test_fir_regression_fixtures(oracle, cuda_device):
    for fn in sorted(os.listdir(GOLDEN)):
        if not (fn.startswith("fir_") and fn.endswith(".npz")):
            continue
        g = np.load(os.path.join(GOLDEN, fn))
        code, tcx = int(g["dtype"]), bool(g["taps_complex"])
        y, cons, prod, _ = _run_gpu(code, "COMPLEX" if tcx else "REAL", g["taps"], int(g["M"]), int(g["L"]), g["x"])
        assert cons == int(g["consumed"]) and prod == int(g["produced"]), fn
        _compare(oracle, code, y, g["y"], f"{fn} [{_.kernel}]")


def test_large_rates_use_generic_kernel(oracle, cuda_device):
    rng = np.random.default_rng(99)
    taps = rng.standard_normal(4000) * 0.05
    for code in (oracle.CF32, oracle.I16):
        x = _rand_input(oracle, code, 60000, rng)
        for M, L in ((1000, 1), (1, 300), (250, 7)):
            y_ref, c_ref, p_ref = oracle.fir(code, False, taps, M, L, x)
            y, cons, prod, _ = _run_gpu(code, "REAL", taps, M, L, x)
            assert (cons, prod) == (c_ref, p_ref)
            _compare(oracle, code, y, y_ref, f"M={M} L={L}")


def test_full_size_linearity_and_sampled_windows(oracle, cuda_device):
    """2^26-sample stream (too long for the oracle end to end): sampled windows vs the oracle,
    plus linearity FIR(a*x1 + x2) == a*FIR(x1) + FIR(x2) within tolerance."""
    import torch
    from pothoscomms_b200 import FirFilter
    from pothoscomms_b200 import workloads as wl
    n = 1 << 26
    taps, tt = wl.config_taps("headline")
    f = FirFilter(oracle.CF32, tt)
    f.set_taps(taps)
    x1 = wl.tone_noise_torch(oracle.CF32, n, 0xC0FFEE01, cuda_device)
    y1, cons, prod = f.run(x1)
    assert cons == n - 255 and prod == n - 255
    for w0 in (0, 12_345_678, n - 40_000):
        seg = x1[w0: w0 + 40_000].cpu().numpy()
        y_ref, _, p_ref = oracle.fir(oracle.CF32, True, taps, 1, 1, seg, threads=8)
        _compare(oracle, oracle.CF32, y1[w0: w0 + p_ref].cpu().numpy(), y_ref, f"window {w0}")
    x2 = wl.tone_noise_torch(oracle.CF32, n, 0xC0FFEE02, cuda_device)
    y2, _, _ = f.run(x2)
    y12, _, _ = f.run(x1 * 0.5 + x2)
    lin = y1 * 0.5 + y2
    err = torch.sqrt(torch.mean((y12 - lin).double() ** 2)).item()
    rms = torch.sqrt(torch.mean(lin.double() ** 2)).item()
    assert err <= 2e-6 * rms + 1e-7


def _with_algo(algo):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = os.environ.get("B200C_FIR_ALGO")
        if algo is None:
            os.environ.pop("B200C_FIR_ALGO", None)
        else:
            os.environ["B200C_FIR_ALGO"] = algo
        try:
            yield
        finally:
            if old is None:
                os.environ.pop("B200C_FIR_ALGO", None)
            else:
                os.environ["B200C_FIR_ALGO"] = old
    return cm()


@pytest.mark.parametrize("ntaps", [2, 3, 24, 64, 255, 256, 1024, 2048, 2049])
@pytest.mark.parametrize("taps_type", ["REAL", "COMPLEX"])
def test_overlap_save_path_matches_oracle_and_direct(oracle, cuda_device, ntaps, taps_type):
    """The fused FFT (overlap-save) kernel that long-tap cf32 streams take must agree with the
    oracle within the float tolerance, for ragged lengths, tiny inputs and the burst zero tail,
    and with the direct FFMA kernel it replaces."""
    rng = np.random.default_rng(ntaps * 3 + (taps_type == "COMPLEX"))
    taps = rng.standard_normal(ntaps) / np.sqrt(ntaps)
    if taps_type == "COMPLEX":
        taps = taps + 1j * rng.standard_normal(ntaps) / np.sqrt(ntaps)
    hop = 4096 - (ntaps - 1)
    rms_hint = float(np.sqrt(np.sum(np.abs(taps) ** 2)))   # unit-variance components in => this per component out
    for n_new, zero_tail in ((1, False), (hop - 1, False), (hop, False), (hop + 1, False), (3 * hop + 17, False),
                             (50000, False), (1000, True), (1, True)):
        x = _rand_input(oracle, oracle.CF32, ntaps - 1 + n_new, rng)
        y_ref, c_ref, p_ref = oracle.fir(oracle.CF32, taps_type == "COMPLEX", taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("fft"):
            y_os, cons, prod, _ = _run_gpu(oracle.CF32, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, oracle.CF32, y_os, y_ref, f"overlap-save K={ntaps} n={n_new} zt={zero_tail}", rms_hint)
        with _with_algo("direct"):
            y_d, cons_d, prod_d, _ = _run_gpu(oracle.CF32, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
        assert (cons_d, prod_d) == (c_ref, p_ref)
        _compare(oracle, oracle.CF32, y_d, y_ref, f"direct K={ntaps} n={n_new}", rms_hint)


def test_overlap_save_limits_fall_back_to_direct(oracle, cuda_device):
    """More than 2049 taps (L = M = 1) and integer types never take the FFT path."""
    rng = np.random.default_rng(4)
    with _with_algo("fft"):
        taps = rng.standard_normal(2050) / 45.0
        x = _rand_input(oracle, oracle.CF32, 2049 + 5000, rng)
        y_ref, c_ref, p_ref = oracle.fir(oracle.CF32, False, taps, 1, 1, x)
        y, cons, prod, _ = _run_gpu(oracle.CF32, "REAL", taps, 1, 1, x)
        assert (cons, prod) == (c_ref, p_ref)
        _compare(oracle, oracle.CF32, y, y_ref, "2050 taps")
        # int16 stays bit-exact (never FFT)
        xi = _rand_input(oracle, oracle.CI16, 5000, rng)
        t2 = rng.standard_normal(64) * 0.1
        yi_ref, _, _ = oracle.fir(oracle.CI16, False, t2, 1, 1, xi)
        yi, _, _, _ = _run_gpu(oracle.CI16, "REAL", t2, 1, 1, xi)
        assert np.array_equal(yi, yi_ref)


@pytest.mark.parametrize("M,L", [(1, 1), (2, 1), (1, 2), (1, 3), (2, 2), (2, 3), (2, 5), (1, 16), (1, 4), (3, 1), (3, 2),
                                 (4, 3), (3, 4), (4, 4), (2, 4)])
@pytest.mark.parametrize("dt,taps_type", [("CF32", "COMPLEX"), ("CF32", "REAL"), ("F32", "REAL")])
def test_overlap_save_polyphase_and_real_data(oracle, cuda_device, dt, taps_type, M, L):
    """The generalised fused kernel (decimation <= 2, any interpolation, complex or real float32
    data) against the oracle and against the direct kernel: ragged lengths around the transform
    hop, tiny inputs, the burst zero tail, tap counts that leave the last phases one tap short."""
    code = getattr(oracle, dt)
    if (dt, M, L) == ("CF32", 1, 1):
        pytest.skip("covered by test_overlap_save_path_matches_oracle_and_direct")
    if dt == "F32" and M > 2:
        pytest.skip("real float32 data: the fused kernel serves decimation <= 2")
    # complex float32 with 2 <= max(L, M) <= 4 (pure M <= 2 decimators aside): the multi-warp resampler kernel
    osp = dt == "CF32" and L <= 4 and M <= 4 and max(L, M) >= 2 and not (L == 1 and M <= 2)
    rng = np.random.default_rng(7000 + 97 * M + L + (taps_type == "COMPLEX"))
    for ntaps in (L * 13 + 1, 255):
        taps = rng.standard_normal(ntaps) / np.sqrt(ntaps / L)
        if taps_type == "COMPLEX":
            taps = taps + 1j * rng.standard_normal(ntaps) / np.sqrt(ntaps / L)
        K = -(-ntaps // L)
        rms_hint = float(np.sqrt(np.sum(np.abs(taps) ** 2) / L))
        for n_new, zero_tail in ((M, False), (M * 700 + 1, False), (M * 2049, False), (M * 5000 + M - 1, False), (997, True), (1, True)):
            x = _rand_input(oracle, code, K - 1 + n_new, rng)
            y_ref, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps, M, L, x, zero_tail=zero_tail)
            with _with_algo("fft"):
                y_os, cons, prod, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail)
                expect = ("fir_os32x_kernel",) if (dt, M, L) == ("CF32", 2, 3) else ("fir_ospg_kernel", "fir_osp_kernel") if osp else \
                    ("fir_os32r_kernel",) if (dt, M, L) == ("F32", 1, 1) else ("fir_os32g_kernel",)
                assert f.kernel in expect, f.kernel
            assert (cons, prod) == (c_ref, p_ref), (ntaps, n_new, zero_tail)
            _compare(oracle, code, y_os, y_ref, f"os32g {dt}/{taps_type} M={M} L={L} K={ntaps} n={n_new} zt={zero_tail}", rms_hint)
            with _with_algo("direct"):
                y_d, cons_d, prod_d, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail)
                assert f.kernel == "fir_tile_kernel", f.kernel
            assert (cons_d, prod_d) == (c_ref, p_ref)
            _compare(oracle, code, y_d, y_ref, f"direct {dt}/{taps_type} M={M} L={L} K={ntaps} n={n_new}", rms_hint)


def test_overlap_save_strong_attenuation_stays_in_tolerance(oracle, cuda_device):
    """Tone in the pass band + noise through a 60 dB stop-band filter: the FFT path's error is
    measured against the OUTPUT rms (north_star tolerance), the case where fast convolution is
    least comfortable."""
    from pothoscomms_b200 import workloads as wl
    taps, tt = wl.config_taps("headline")
    x = wl.tone_noise_numpy(oracle.CF32, 1 << 17, seed=99)
    y_ref, _, _ = oracle.fir(oracle.CF32, True, taps, 1, 1, x, threads=8)
    with _with_algo("fft"):
        y, _, _, _ = _run_gpu(oracle.CF32, tt, taps, 1, 1, x)
    _compare(oracle, oracle.CF32, y, y_ref, "headline via overlap-save")


@pytest.mark.parametrize("dt,taps_type,ntaps", [("CF32", "COMPLEX", 64), ("CF32", "COMPLEX", 1024), ("CF32", "REAL", 300),
                                                ("CI16", "COMPLEX", 33), ("F32", "REAL", 40)])
def test_filter_bank_matches_per_channel_filters(oracle, cuda_device, dt, taps_type, ntaps):
    """b200c_fir_bank_run: every channel's output equals what its own FIRFilter instance gives
    (the oracle per channel): one launch over (channel, block) for complex float32, the
    channels' kernels back to back otherwise; padded channel strides; zero tail."""
    import torch
    from pothoscomms_b200 import FirFilterBank
    code = getattr(oracle, dt)
    rng = np.random.default_rng(ntaps + code)
    nchan, n_in = 7, ntaps - 1 + 9001
    nc = 2 if code & 1 else 1
    bank = FirFilterBank(code, taps_type, nchan)
    taps = []
    for c in range(nchan):
        h = rng.standard_normal(ntaps) / np.sqrt(ntaps)
        if taps_type == "COMPLEX":
            h = h + 1j * rng.standard_normal(ntaps) / np.sqrt(ntaps)
        taps.append(h)
        bank.set_taps(c, h)
    assert bank.info()[:2] == (nchan, ntaps)
    x = np.stack([_rand_input(oracle, code, n_in, rng) for _ in range(nchan)])
    d = torch.from_numpy(x).cuda()
    for zero_tail in (False, True):
        cap = n_in + 5
        out = torch.zeros((nchan, cap, nc), dtype=d.dtype, device=d.device)
        cons, prod = bank.run(d, out, zero_tail=zero_tail)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for c in range(nchan):
            y_ref, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps[c], 1, 1, x[c], zero_tail=zero_tail)
            assert (cons, prod) == (c_ref, p_ref)
            _compare(oracle, code, got[c, :prod], y_ref, f"bank {dt} chan {c} zt={zero_tail}")
            assert not got[c, prod:].any(), "wrote past the channel's produced count"


# ---------------------------------------------------------------- int8 tensor-core path ---
@pytest.mark.parametrize("ntaps", [2, 12, 25, 26, 57, 121, 128, 129, 255, 1000, 2048])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_imma_path_is_bit_exact(oracle, cuda_device, dt, taps_type, ntaps):
    """int16 streams (L = M = 1) take the byte-limb Toeplitz GEMM kernel (fir_imma.cu); it must
    reproduce the reference's wrapping int32 accumulation and >>16 bit for bit, for ragged
    lengths, tiles that end mid-row, the burst zero tail and full-scale inputs, and agree with
    the direct IMAD kernel it replaces."""
    code = getattr(oracle, dt)
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(ntaps * 7 + code)
    taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    if cx:
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    for n_new, zero_tail in ((1, False), (127, False), (128, False), (4095, False), (4096, False), (4097, False),
                             (3 * 4096 + 1001, False), (70001, False), (1000, True), (1, True)):
        x = _rand_input(oracle, code, ntaps - 1 + n_new, rng, full_scale=True)
        y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("imma"):
            y, cons, prod, f = _run_gpu(code, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
            assert f.kernel == "fir_imma_kernel"
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, code, y, y_ref, f"imma K={ntaps} n={n_new} zt={zero_tail}")
    with _with_algo("direct"):
        y_d, _, _, f = _run_gpu(code, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
        assert f.kernel != "fir_imma_kernel"
    _compare(oracle, code, y_d, y_ref, "direct")


@pytest.mark.parametrize("scale,limbs", [(0.4, 2), (0.6, 3), (100.0, 3), (200.0, 4), (30000.0, 4)])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_imma_tap_limb_counts_and_wrapping(oracle, cuda_device, dt, taps_type, scale, limbs):
    """Q16 taps of any magnitude: 2, 3 or 4 balanced byte digits per tap; the int32 accumulator
    wraps exactly like the reference's (filter/FIRFilter.cpp:381) however large the taps are."""
    code = getattr(oracle, dt)
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(int(scale * 10) + code)
    ntaps = 77
    taps = rng.uniform(-scale, scale, ntaps)
    taps[5] = scale * 0.999                                  # make sure the top digit is in use
    taps[6] = -scale * 0.999
    if cx:
        taps = taps + 1j * rng.uniform(-scale, scale, ntaps)
    x = _rand_input(oracle, code, 9000, rng, full_scale=True)
    y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, 1, 1, x)
    y, cons, prod, f = _run_gpu(code, taps_type, taps, 1, 1, x)
    # tcgen05 paths take 2-limb taps (the operand-swapped kernel at this tap count)
    assert f.kernel == ("fir_umma32t_kernel" if limbs == 2 else "fir_imma_kernel")
    assert (cons, prod) == (c_ref, p_ref)
    _compare(oracle, code, y, y_ref, f"scale={scale} ({limbs} limbs)")


@pytest.mark.parametrize("kernel", ["fir_imma_kernel", "fir_umma_kernel", "fir_umma32_kernel", "fir_umma32t_kernel"])
@pytest.mark.parametrize("dt", ["CI16", "I16"])
def test_imma_unaligned_device_pointers(oracle, cuda_device, dt, kernel):
    """A ring-buffer window starts at any element: the kernel's 16-byte loads and 8-byte stores
    need their guarded fallbacks."""
    import torch
    from pothoscomms_b200 import FirFilter
    code = getattr(oracle, dt)
    rng = np.random.default_rng(99)
    taps = rng.standard_normal(64) * 0.05
    with _with_algo({"fir_imma_kernel": "imma", "fir_umma_kernel": "umma", "fir_umma32_kernel": "umma32",
                     "fir_umma32t_kernel": "umma32t"}[kernel]):
        f = FirFilter(code, "REAL")
        f.set_taps(taps)
    assert f.kernel == kernel
    x = _rand_input(oracle, code, 40000, rng, full_scale=True)
    y_ref, _, _ = oracle.fir(code, False, taps, 1, 1, x)
    xd = torch.from_numpy(x).cuda()
    # base pointers shifted by `off` elements (the swapped tcgen05 kernels bulk-copy from the aligned address below and
    # shift the bytes back in the stagers: every byte offset of a 16-byte unit for real data, every word offset for complex)
    for off in ((1, 2, 3, 4, 5, 6, 7, 9) if dt == "I16" else (1, 2, 3, 5)):
        shifted = torch.empty(x.shape[0] + off, x.shape[1], dtype=xd.dtype, device="cuda")
        shifted[off:] = xd
        out_big = torch.zeros(y_ref.shape[0] + off + 1, x.shape[1], dtype=xd.dtype, device="cuda")
        y, cons, prod = f.run(shifted[off:], out=out_big[off:])
        torch.cuda.synchronize()
        assert np.array_equal(y.cpu().numpy()[:prod], y_ref), off


# ------------------------------------------------------- tcgen05 / tensor-memory path ---
@pytest.mark.parametrize("ntaps", [2, 17, 18, 49, 50, 113, 128, 129, 255, 700])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_umma_path_is_bit_exact(oracle, cuda_device, dt, taps_type, ntaps):
    """The tcgen05.mma kind::i8 kernel (fir_umma.cu: Hankel A operand aliased onto the byte planes
    by the shared-memory descriptor, accumulators in tensor memory) against the oracle: ragged
    lengths, tiles ending mid-row, the zero tail, full-scale inputs; tap counts either side of
    the k-block boundaries (K + 15 = 32 j)."""
    code = getattr(oracle, dt)
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(ntaps * 11 + code)
    taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    if cx:
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    for n_new, zero_tail in ((1, False), (15, False), (16, False), (2047, False), (2048, False), (2049, False),
                             (5 * 2048 + 1001, False), (300001, False), (1000, True), (1, True)):
        x = _rand_input(oracle, code, ntaps - 1 + n_new, rng, full_scale=True)
        y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("umma"):
            y, cons, prod, f = _run_gpu(code, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
            assert f.kernel == "fir_umma_kernel"
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, code, y, y_ref, f"umma K={ntaps} n={n_new} zt={zero_tail}")


@pytest.mark.parametrize("ntaps", [2, 33, 34, 65, 66, 128, 129, 255])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_umma32_path_is_bit_exact(oracle, cuda_device, dt, taps_type, ntaps):
    """The second tcgen05 formulation (fir_umma32.cu): 32-byte-swizzled byte planes read through a
    SWIZZLE_32B descriptor advanced one row per k-block, re/im planes accumulating into shared
    tensor-memory regions through signed tap-digit matrices.  Tap counts either side of the k-block
    boundaries (K + 31 = 32 j), tiles of 4096 outputs ending mid-row, zero tail, full-scale input."""
    code = getattr(oracle, dt)
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(ntaps * 13 + code)
    taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    if cx:
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    for n_new, zero_tail in ((1, False), (31, False), (32, False), (4095, False), (4096, False), (4097, False),
                             (5 * 4096 + 1001, False), (700001, False), (1000, True), (1, True)):
        x = _rand_input(oracle, code, ntaps - 1 + n_new, rng, full_scale=True)
        y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("umma32"):
            y, cons, prod, f = _run_gpu(code, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
            assert f.kernel == "fir_umma32_kernel"
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, code, y, y_ref, f"umma32 K={ntaps} n={n_new} zt={zero_tail}")


@pytest.mark.parametrize("ntaps", [2, 33, 34, 65, 66, 128, 129, 193, 194, 225, 321, 449])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_umma32t_path_is_bit_exact(oracle, cuda_device, dt, taps_type, ntaps):
    """The operand-swapped formulation (fir_umma32t_kernel): the tap-digit tiles are the A operand and sit in
    tensor memory, the swizzled data planes are the B operand (96 windows per tile), the epilogue reads both digits
    of an output with the 16-lane tensor-memory load shape.  Complex data: windows of 32 outputs, 32-byte-swizzled
    rows; real data: windows of 64 outputs, 64-byte-swizzled rows read 32 bytes at a time.  Tap counts either side
    of the k-block boundaries up to the tensor-memory limit (complex: 8 k-blocks = 225 taps; real: 16 = 449), tiles of 3072 /
    6144 outputs ending mid-window, odd output counts (real: the last 32-bit word half used), zero tail, full-scale
    input."""
    code = getattr(oracle, dt)
    if dt == "CI16" and ntaps > 225:
        pytest.skip("complex data: 32 outputs + K - 1 positions must fit 8 k-blocks")
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(ntaps * 17 + 3 + code)
    taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    if cx:
        taps = taps + 1j * rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    tile = 3072 if dt == "CI16" else 6144
    for n_new, zero_tail in ((1, False), (31, False), (32, False), (33, False), (63, False), (64, False), (65, False),
                             (tile - 1, False), (tile, False), (tile + 1, False), (5 * tile + 1001, False), (148 * tile + 17, False),
                             (700001, False), (1000, True), (1, True)):
        x = _rand_input(oracle, code, ntaps - 1 + n_new, rng, full_scale=True)
        y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("umma32t"):
            y, cons, prod, f = _run_gpu(code, taps_type, taps, 1, 1, x, zero_tail=zero_tail)
            assert f.kernel == "fir_umma32t_kernel"
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, code, y, y_ref, f"umma32t {dt} K={ntaps} n={n_new} zt={zero_tail}")


@pytest.mark.parametrize("algo,kernel", [("umma32", "fir_umma32_kernel"), ("umma32t", "fir_umma32t_kernel")])
def test_umma32_kernels_over_many_tiles_per_cta(oracle, cuda_device, algo, kernel):
    """Twenty-odd tiles per persistent CTA: both accumulator stages and every slot of the bulk-copy landing ring are
    reused several times (the short cases above give a CTA two tiles at most, which never waits on a released stage).
    Oracle windows at the start, inside late tiles and at the ragged end; the whole output against the other variant."""
    import torch
    from pothoscomms_b200 import FirFilter
    code = oracle.CI16
    rng = np.random.default_rng(2024)
    ntaps = 128
    taps = (rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)) * 0.3 / np.sqrt(ntaps)
    n = 21 * 148 * 3072 + 777
    x = torch.randint(-32768, 32767, (ntaps - 1 + n, 2), dtype=torch.int16, device=cuda_device)
    outs = {}
    for a in ("umma32", "umma32t"):
        with _with_algo(a):
            f = FirFilter(code, "COMPLEX")
            f.set_taps(taps)
        y, cons, prod = f.run(x)
        torch.cuda.synchronize()
        assert (cons, prod) == (n, n)
        outs[a] = (f.kernel, y[:prod].clone())
    assert outs[algo][0] == kernel
    y = outs[algo][1]
    wlen = 20_000
    for w0 in (0, 3 * 148 * 3072 - 5000, 10 * 148 * 4096 + 123, n // 2, n - wlen):
        seg = x[w0: w0 + ntaps - 1 + wlen].cpu().numpy()
        y_ref, _, p_ref = oracle.fir(code, True, taps, 1, 1, seg)
        _compare(oracle, code, y[w0: w0 + p_ref].cpu().numpy(), y_ref, f"{algo} window at {w0}")
    assert torch.equal(outs["umma32"][1], outs["umma32t"][1])


def test_umma32t_falls_back_beyond_its_tensor_memory_budget(oracle, cuda_device):
    """More than 8 k-blocks of tap tiles do not fit beside the accumulators: the original formulation runs."""
    from pothoscomms_b200 import FirFilter
    taps = np.random.default_rng(5).standard_normal(300) * 0.01
    with _with_algo("umma32t"):
        f = FirFilter(oracle.CI16, "REAL")
        f.set_taps(taps)
    assert f.kernel == "fir_umma32_kernel"
    with _with_algo("umma32t"):
        f = FirFilter(oracle.I16, "REAL")
        f.set_taps(np.random.default_rng(6).standard_normal(460) * 0.01)   # real data: 64 + 459 positions are seventeen k-blocks
    assert f.kernel == "fir_umma32_kernel"


@pytest.mark.parametrize("M,L", [(2, 1), (1, 2), (3, 1), (1, 3), (2, 3), (3, 2), (4, 3), (3, 4), (4, 4), (2, 2), (1, 4), (4, 1)])
@pytest.mark.parametrize("dt,taps_type", [("CI16", "COMPLEX"), ("CI16", "REAL"), ("I16", "REAL")])
def test_ummap_polyphase_is_bit_exact(oracle, cuda_device, dt, taps_type, M, L):
    """int16 resampling (interpolation, decimation <= 4) on tcgen05 (fir_ummap.cu): one Toeplitz GEMM per
    input residue, all residues and components accumulating into shared tensor-memory regions.  Against
    the oracle's loop nest (filter/FIRFilter.cpp:286-302) for tap counts that leave the last phases one
    tap short, ragged lengths around the 2048-block tile, tiny inputs, the burst zero tail."""
    code = getattr(oracle, dt)
    cx = taps_type == "COMPLEX"
    rng = np.random.default_rng(9000 + 97 * M + L + cx)
    for ntaps in (L * 13 + 1, 255):
        taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps / L)
        if cx:
            taps = taps + 1j * rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps / L)
        K = -(-ntaps // L)
        for n_new, zero_tail in ((M, False), (M * 2047, False), (M * 2048, False), (M * 2049 + M - 1, False), (M * 9000 + 1, False),
                                 (997, True), (1, True)):
            x = _rand_input(oracle, code, K - 1 + n_new, rng, full_scale=True)
            y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, M, L, x, zero_tail=zero_tail)
            y, cons, prod, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail)
            assert f.kernel == "fir_ummap_kernel", f.kernel
            assert (cons, prod) == (c_ref, p_ref), (ntaps, n_new, zero_tail)
            _compare(oracle, code, y, y_ref, f"ummap {dt}/{taps_type} M={M} L={L} K={ntaps} n={n_new} zt={zero_tail}")


@pytest.mark.parametrize("ntaps", [2, 5, 16, 64, 65, 255, 448])
def test_real_float32_overlap_save_kernel(oracle, cuda_device, ntaps):
    """float32 streams (REAL taps, L = M = 1) take fir_os32r_kernel: two stream blocks per complex
    transform (z = x_A + i x_B).  Ragged lengths around the pair hop, odd block counts (a lone block A),
    tiny inputs, the zero tail, unaligned base pointers; against the oracle and the direct kernel."""
    import torch
    from pothoscomms_b200 import FirFilter
    rng = np.random.default_rng(ntaps * 5 + 1)
    taps = rng.standard_normal(ntaps) / np.sqrt(ntaps)
    hop = 1024 - (ntaps - 1)
    rms_hint = float(np.sqrt(np.sum(taps ** 2)))
    for n_new, zero_tail in ((1, False), (hop, False), (hop + 1, False), (2 * hop, False), (2 * hop + 1, False), (7 * hop + 13, False),
                             (200003, False), (1000, True), (1, True)):
        x = _rand_input(oracle, oracle.F32, ntaps - 1 + n_new, rng)
        y_ref, c_ref, p_ref = oracle.fir(oracle.F32, False, taps, 1, 1, x, zero_tail=zero_tail)
        with _with_algo("fft" if ntaps < 9 else None):   # up to 8 taps the automatic choice is the direct kernel
            y, cons, prod, f = _run_gpu(oracle.F32, "REAL", taps, 1, 1, x, zero_tail=zero_tail)
        assert f.kernel == "fir_os32r_kernel", f.kernel
        assert (cons, prod) == (c_ref, p_ref), (n_new, zero_tail)
        _compare(oracle, oracle.F32, y, y_ref, f"os32r K={ntaps} n={n_new} zt={zero_tail}", rms_hint)
    with _with_algo("direct"):
        y_d, _, _, f = _run_gpu(oracle.F32, "REAL", taps, 1, 1, x, zero_tail=zero_tail)
        assert f.kernel == "fir_tile_kernel"
    _compare(oracle, oracle.F32, y_d, y_ref, "direct", rms_hint)
    # base pointers shifted by 1..3 floats: the bulk copy re-aligns, guarded loads cover the edges
    x = _rand_input(oracle, oracle.F32, 30000, rng)
    y_ref, _, _ = oracle.fir(oracle.F32, False, taps, 1, 1, x)
    f = FirFilter(oracle.F32, "REAL")
    f.set_taps(taps)
    xd = torch.from_numpy(x).cuda()
    for off in (1, 2, 3):
        shifted = torch.empty((x.shape[0] + off, 1), dtype=xd.dtype, device="cuda")
        shifted[off:] = xd
        y, _, prod = f.run(shifted[off:])
        torch.cuda.synchronize()
        _compare(oracle, oracle.F32, y.cpu().numpy()[:prod], y_ref, f"os32r unaligned {off}", rms_hint)


def _with_env(name, value):
    import contextlib
    import os

    @contextlib.contextmanager
    def cm():
        old = os.environ.get(name)
        os.environ[name] = value
        try:
            yield
        finally:
            if old is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = old
    return cm()


@pytest.mark.parametrize("ntaps", [2, 7, 40, 48, 100, 255, 301, 700, 1200])
@pytest.mark.parametrize("taps_type", ["REAL", "COMPLEX"])
def test_spectral_resampler_interp3_decim2(oracle, cuda_device, ntaps, taps_type):
    """complex float32, interpolation 3 / decimation 2 (BASELINE config C3) takes fir_os32x_kernel: one 1024-point
    forward and one 1536-point inverse transform per block (spectrum replicated, multiplied, folded).  Ragged
    lengths around the block hop, tiny inputs, the burst zero tail, unaligned base pointers, an output capacity
    that ends mid-block; against the oracle and against the grouped polyphase kernel it replaces."""
    import torch
    from pothoscomms_b200 import FirFilter
    code, M, L = oracle.CF32, 2, 3
    rng = np.random.default_rng(31 * ntaps + (taps_type == "COMPLEX"))
    taps = rng.standard_normal(ntaps) / np.sqrt(ntaps / L)
    if taps_type == "COMPLEX":
        taps = taps + 1j * rng.standard_normal(ntaps) / np.sqrt(ntaps / L)
    K = -(-ntaps // L)
    m0 = (-(-(ntaps - 2) // 2) + 2) // 3 * 3
    hop_in = (1536 - m0) // 3 * 2
    rms_hint = float(np.sqrt(np.sum(np.abs(taps) ** 2) / L))
    for n_new, zero_tail in ((M, False), (hop_in - 2, False), (hop_in, False), (hop_in + 2, False), (3 * hop_in + 2, False),
                             (M * 2049, False), (M * 25000 + 1, False), (997, True), (1, True)):
        x = _rand_input(oracle, code, K - 1 + n_new, rng)
        y_ref, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps, M, L, x, zero_tail=zero_tail)
        with _with_algo("fft"):
            y, cons, prod, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail)
        assert f.kernel == "fir_os32x_kernel", f.kernel
        assert (cons, prod) == (c_ref, p_ref), (ntaps, n_new, zero_tail)
        _compare(oracle, code, y, y_ref, f"os32x {taps_type} K={ntaps} n={n_new} zt={zero_tail}", rms_hint)
    # the kernel it replaced, on the same stream (the grouped polyphase kernel is dispatched from 40 taps per phase)
    if K >= 40:
        with _with_env("B200C_OSX", "0"):
            y_g, _, _, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail)
            assert f.kernel in ("fir_ospg_kernel", "fir_osp_kernel"), f.kernel
        _compare(oracle, code, y_g, y_ref, "ospg", rms_hint)
    # unaligned base pointers (the bulk copy re-aligns by one element) and a capacity that ends inside a block
    x = _rand_input(oracle, code, 30001, rng)
    y_ref, _, _ = oracle.fir(code, taps_type == "COMPLEX", taps, M, L, x)
    with _with_algo("fft"):
        f = FirFilter(code, taps_type)
        f.set_taps(taps)
        f.set_rates(M, L)
    xd = torch.from_numpy(x).cuda()
    for off in (1, 2, 3):
        shifted = torch.empty((x.shape[0] + off, 2), dtype=xd.dtype, device="cuda")
        shifted[off:] = xd
        y, _, prod = f.run(shifted[off:])
        torch.cuda.synchronize()
        _compare(oracle, code, y.cpu().numpy()[:prod], y_ref, f"os32x unaligned {off}", rms_hint)
    cap = 3 * 1000 + 3
    y_ref_c, c_ref, p_ref = oracle.fir(code, taps_type == "COMPLEX", taps, M, L, x, out_capacity=cap)
    y, cons, prod = f.run(xd, out_capacity=cap)
    torch.cuda.synchronize()
    assert (cons, prod) == (c_ref, p_ref)
    _compare(oracle, code, y.cpu().numpy()[:prod], y_ref_c, "os32x capacity", rms_hint)


@pytest.mark.parametrize("dt", ["CF32", "F32"])
def test_short_float_filters_keep_time_domain_semantics(oracle, cuda_device, dt):
    """Filters of up to 8 taps stay on the direct kernel (automatic choice): what the reference's time-domain
    nest guarantees per output then holds here too -- a pure delay or an integer gain passes samples bit for bit,
    and a non-finite input sample contaminates exactly the K outputs whose window holds it
    (filter/FIRFilter.cpp:295-299).  Longer filters take fast convolution, where a NaN spreads over its transform
    block; documented in INTEGRATION.md, checked here as the contract it is."""
    code = getattr(oracle, dt)
    rng = np.random.default_rng(3)
    x = _rand_input(oracle, code, 5000, rng)
    for taps in ([0.0, 1.0], [2.0], [0.0, 0.0, 0.0, -4.0], [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]):
        y_ref, c_ref, p_ref = oracle.fir(code, False, taps, 1, 1, x)
        y, cons, prod, f = _run_gpu(code, "REAL", taps, 1, 1, x)
        assert f.kernel in ("fir_tile_kernel", "fir_generic_kernel"), f.kernel
        assert (cons, prod) == (c_ref, p_ref)
        assert np.array_equal(y.view(np.uint32), y_ref.view(np.uint32)), taps    # exact filters: bit for bit
    taps = rng.standard_normal(8)
    K = 8
    xn = x.copy()
    xn[1000, 0] = np.nan
    xn[3000, -1] = np.inf
    y, cons, prod, f = _run_gpu(code, "REAL", taps, 1, 1, xn)
    bad = ~np.isfinite(y).all(axis=1)
    expect = np.zeros(prod, dtype=bool)
    for pos in (1000, 3000):
        expect[max(pos - (K - 1), 0): pos + 1] = True        # outputs n with n <= pos <= n + K - 1 (history offset K-1)
    # exactly the K outputs of the reference, plus at most R - 1 <= 8 neighbours per sample: the tile kernel pads the tap
    # list to its register block with zero taps, and 0 * NaN is NaN (INTEGRATION.md "non-finite samples")
    near = np.zeros(prod, dtype=bool)
    for pos in (1000, 3000):
        near[max(pos - (K - 1) - 8, 0): pos + 1 + 8] = True
    assert bad[expect].all() and not bad[~near].any() and bad.sum() <= expect.sum() + 16, \
        (np.flatnonzero(bad).tolist(), int(bad.sum()), int(expect.sum()))
    # the fused path (forced, or chosen from 9 taps up) spreads the same sample over its transform block(s)
    with _with_algo("fft"):
        y2, _, _, f2 = _run_gpu(code, "REAL", taps, 1, 1, xn)
    assert f2.kernel.startswith("fir_os")
    bad2 = ~np.isfinite(y2).all(axis=1)
    assert bad2[expect].all() and bad2.sum() > expect.sum() and bad2.sum() <= 4 * 2048


@pytest.mark.parametrize("dt", ["I8", "CI8", "I32", "CI32", "I64", "CI64", "F64", "CF64"])
def test_untuned_types_at_tap_counts_that_prefer_wider_register_blocks(oracle, cuda_device, dt):
    """The direct kernel has R = 7 / 9 register blocks for the float32 and int16 families only; every other type must
    plan R = 5.  Round 1 planned R = 7 for int8 at 11 taps per phase while launching R = 5 (a tile covered 5/7 of its
    blocks): the reference-generated fixture fir_ci8_cc_21_l2 caught it.  Tap counts whose padding favours 7 or 9."""
    code = getattr(oracle, dt)
    tcx = bool(code & 1)
    rng = np.random.default_rng(11)
    x = _rand_input(oracle, code, 6000, rng)
    for M, L in ((1, 1), (1, 2), (2, 1), (3, 2)):
        for per_phase in (7, 9, 11, 14, 18, 27):
            ntaps = per_phase * L - (1 if L > 1 else 0)
            taps = rng.standard_normal(ntaps) * 0.2 + (1j * rng.standard_normal(ntaps) * 0.2 if tcx else 0)
            y_ref, c_ref, p_ref = oracle.fir(code, tcx, taps, M, L, x)
            y, cons, prod, f = _run_gpu(code, "COMPLEX" if tcx else "REAL", taps, M, L, x)
            assert (cons, prod) == (c_ref, p_ref)
            _compare(oracle, code, y, y_ref, f"{dt} M={M} L={L} ntaps={ntaps} [{f.kernel}]")


@pytest.mark.parametrize("cfg,code_name,M,L,log2n", [("c2", "CI16", 1, 1, 28), ("c3", "CF32", 2, 3, 30), ("c5", "CF32", 1, 1, 26)])
def test_baseline_configs_at_their_stated_sizes_sampled_windows(oracle, cuda_device, cfg, code_name, M, L, log2n):
    """BASELINE.json's C2 (complex int16, 128 taps, 2^28 samples), C3 (L = 3 / M = 2 resampler, ONE 2^30-sample stream on one
    GPU: 8 GiB in, 12 GiB out) and a C5 channel at the full stated sizes: the oracle cannot run these end to end, so windows
    are sampled -- the start (history-only prefix), block boundaries deep inside, the ragged end -- and compared with the
    oracle run on the same input slice: bit-exact for int16, 1e-5 of RMS for float."""
    import torch
    from pothoscomms_b200 import FirFilter
    from pothoscomms_b200 import workloads as wl
    code = getattr(oracle, code_name)
    taps, tt = wl.config_taps(cfg)
    f = FirFilter(code, tt)
    f.set_taps(taps)
    f.set_rates(M, L)
    K = f.K
    n = (1 << log2n) // M * M
    x = wl.tone_noise_torch(code, K - 1 + n, 0xC0FFEE00 + log2n, cuda_device)
    y, cons, prod = f.run(x)
    torch.cuda.synchronize()
    assert (cons, prod) == (n, n // M * L)
    wlen = 30_000 // M * M
    for w0 in (0, (n // 3) // M * M, (n // 2 + 7777) // M * M, n - wlen):       # input offsets, multiples of M
        seg = x[w0: w0 + K - 1 + wlen].cpu().numpy()
        y_ref, c_ref, p_ref = oracle.fir(code, tt == "COMPLEX", taps, M, L, seg, threads=8)
        o0 = w0 // M * L
        _compare(oracle, code, y[o0: o0 + p_ref].cpu().numpy(), y_ref, f"{cfg} 2^{log2n} window at {w0} [{f.kernel}]")
    del x, y
    torch.cuda.empty_cache()


def test_back_to_back_launches_chain_through_device_buffers(oracle, cuda_device):
    """Programmatic dependent launch (fir_os32_kernel's persistent form): a launch may start its prologue before its
    predecessor in the stream has finished, and must not touch stream data before `griddepcontrol.wait`.  Two filters
    chained through a device buffer (B reads what A just wrote), forty small buffers back to back with no host
    synchronisation in between, A's output buffer reused every round: every round's result against the oracle."""
    import torch
    from pothoscomms_b200 import FirFilter
    code = oracle.CF32
    rng = np.random.default_rng(77)
    ta = rng.standard_normal(64) / 8 + 1j * rng.standard_normal(64) / 8
    tb = rng.standard_normal(256) / 16 + 1j * rng.standard_normal(256) / 16
    fa, fb = FirFilter(code, "COMPLEX"), FirFilter(code, "COMPLEX")
    fa.set_taps(ta)
    fb.set_taps(tb)
    assert fa.kernel == "fir_os32_kernel" and fb.kernel == "fir_os32_kernel"
    n = 150_000
    rounds = 40
    xs = [torch.from_numpy(_rand_input(oracle, code, n, rng)).cuda() for _ in range(4)]
    mid = torch.empty((n, 2), dtype=torch.float32, device=cuda_device)
    outs = [torch.empty((n, 2), dtype=torch.float32, device=cuda_device) for _ in range(rounds)]
    prods = []
    torch.cuda.synchronize()
    for r in range(rounds):
        _, _, pa = fa.run(xs[r % 4], out=mid)
        _, _, pb = fb.run(mid[:pa], out=outs[r])
        prods.append((pa, pb))
    torch.cuda.synchronize()
    refs = {}
    for k in range(4):
        ya, _, pa = oracle.fir(code, True, ta, 1, 1, xs[k].cpu().numpy())
        yb, _, pb = oracle.fir(code, True, tb, 1, 1, ya)
        refs[k] = (yb, pa, pb)
    for r in range(rounds):
        yb, pa, pb = refs[r % 4]
        assert prods[r] == (pa, pb)
        _compare(oracle, code, outs[r][:pb].cpu().numpy(), yb, f"round {r}")


def test_back_to_back_launches_resampler_then_filter(oracle, cuda_device):
    """The same chain with the spectral resampler (fir_os32x_kernel, also a programmatic dependent) feeding the filter."""
    import torch
    from pothoscomms_b200 import FirFilter
    from pothoscomms_b200 import workloads as wl
    code = oracle.CF32
    rng = np.random.default_rng(78)
    ta, tta = wl.config_taps("c3")
    tb = rng.standard_normal(256) / 16 + 1j * rng.standard_normal(256) / 16
    fa, fb = FirFilter(code, tta), FirFilter(code, "COMPLEX")
    fa.set_taps(ta)
    fa.set_rates(2, 3)
    fb.set_taps(tb)
    assert fa.kernel == "fir_os32x_kernel" and fb.kernel == "fir_os32_kernel"
    n = 100_000
    rounds = 24
    xs = [torch.from_numpy(_rand_input(oracle, code, n, rng)).cuda() for _ in range(3)]
    mid = torch.empty((n * 3 // 2 + 8, 2), dtype=torch.float32, device=cuda_device)
    outs = [torch.empty((n * 3 // 2 + 8, 2), dtype=torch.float32, device=cuda_device) for _ in range(rounds)]
    prods = []
    torch.cuda.synchronize()
    for r in range(rounds):
        _, _, pa = fa.run(xs[r % 3], out=mid)
        _, _, pb = fb.run(mid[:pa], out=outs[r])
        prods.append((pa, pb))
    torch.cuda.synchronize()
    refs = {}
    for k in range(3):
        ya, _, pa = oracle.fir(code, tta == "COMPLEX", ta, 2, 3, xs[k].cpu().numpy())
        yb, _, pb = oracle.fir(code, True, tb, 1, 1, ya)
        refs[k] = (yb, pa, pb)
    for r in range(rounds):
        yb, pa, pb = refs[r % 3]
        assert prods[r] == (pa, pb)
        _compare(oracle, code, outs[r][:pb].cpu().numpy(), yb, f"round {r}")


def test_umma32t_real_data_over_many_tiles_per_cta(oracle, cuda_device):
    """The operand-swapped kernel on REAL int16 (64-output windows, 64-byte-swizzled planes): eleven tiles per persistent
    CTA, oracle windows at the start, deep inside and at the ragged (odd) end, the whole output against the original kernel."""
    import torch
    from pothoscomms_b200 import FirFilter
    code = oracle.I16
    rng = np.random.default_rng(2025)
    ntaps = 64
    taps = rng.standard_normal(ntaps) * 0.3 / np.sqrt(ntaps)
    n = 11 * 148 * 6144 + 333
    x = torch.randint(-32768, 32767, (ntaps - 1 + n, 1), dtype=torch.int16, device=cuda_device)
    outs = {}
    for a in ("umma32", "umma32t"):
        with _with_algo(a):
            f = FirFilter(code, "REAL")
            f.set_taps(taps)
        y, cons, prod = f.run(x)
        torch.cuda.synchronize()
        assert (cons, prod) == (n, n)
        outs[a] = (f.kernel, y[:prod].clone())
    assert outs["umma32"][0] == "fir_umma32_kernel" and outs["umma32t"][0] == "fir_umma32t_kernel"
    y = outs["umma32t"][1]
    wlen = 30_000
    for w0 in (0, 3 * 148 * 6144 - 7000, 7 * 148 * 6144 + 123, n // 2, n - wlen):
        seg = x[w0: w0 + ntaps - 1 + wlen].cpu().numpy()
        y_ref, _, p_ref = oracle.fir(code, False, taps, 1, 1, seg)
        _compare(oracle, code, y[w0: w0 + p_ref].cpu().numpy(), y_ref, f"window at {w0}")
    assert torch.equal(outs["umma32"][1], outs["umma32t"][1])


def test_gpu_equals_oracle_fuzz(oracle, cuda_device):
    """Property-based sweep through the C ABI: random row of the 18-row type table, rates, tap count and scale (two- to
    four-digit Q16 taps), window length, zero tail and output capacity -- whatever kernel the dispatch picks must reproduce
    the oracle: counts exactly, integers bit for bit, floats within the stated tolerance."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    rows = [("F32", "REAL"), ("CF32", "REAL"), ("CF32", "COMPLEX"), ("F64", "REAL"), ("CF64", "REAL"), ("CF64", "COMPLEX"),
            ("I8", "REAL"), ("CI8", "REAL"), ("CI8", "COMPLEX"), ("I16", "REAL"), ("CI16", "REAL"), ("CI16", "COMPLEX"),
            ("I32", "REAL"), ("CI32", "REAL"), ("CI32", "COMPLEX"), ("I64", "REAL"), ("CI64", "REAL"), ("CI64", "COMPLEX")]
    seen = set()

    @settings(max_examples=200, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(row=st.integers(0, len(rows) - 1), M=st.integers(1, 5), L=st.integers(1, 5), ntaps=st.integers(1, 300),
           n=st.integers(1, 30000), scale=st.sampled_from([0.02, 0.3, 0.7, 150.0]), zero_tail=st.booleans(),
           cap_frac=st.sampled_from([None, None, 1.0, 0.37]), seed=st.integers(0, 2 ** 31))
    def check(row, M, L, ntaps, n, scale, zero_tail, cap_frac, seed):
        dt, taps_type = rows[row]
        code = getattr(oracle, dt)
        cx = taps_type == "COMPLEX"
        rng = np.random.default_rng(seed)
        taps = rng.uniform(-scale, scale, ntaps) / np.sqrt(ntaps)
        if cx:
            taps = taps + 1j * rng.uniform(-scale, scale, ntaps) / np.sqrt(ntaps)
        x = _rand_input(oracle, code, n, rng, full_scale=True)
        cap = None if cap_frac is None else int(cap_frac * (n // M + 1) * L)
        y_ref, c_ref, p_ref = oracle.fir(code, cx, taps, M, L, x, zero_tail=zero_tail, out_capacity=cap)
        y, cons, prod, f = _run_gpu(code, taps_type, taps, M, L, x, zero_tail=zero_tail, out_capacity=cap)
        seen.add(f.kernel)
        what = f"{dt} {taps_type} M={M} L={L} ntaps={ntaps} n={n} scale={scale} zt={zero_tail} cap={cap} [{f.kernel}]"
        assert (cons, prod) == (c_ref, p_ref), what
        _compare(oracle, code, y[:prod], y_ref, what, rms_hint=float(np.sqrt(np.mean(np.abs(taps) ** 2) * ntaps)))

    check()
    assert len(seen) >= 5, seen        # the sweep crosses the dispatch table
