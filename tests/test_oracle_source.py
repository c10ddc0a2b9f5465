"""CPU checks of the stream-source oracle (oracle/source_oracle.cpp): the waveform tables against
closed forms written in numpy, the table size / step search on the cases the reference's doc and
test use (waveform/WaveformSource.cpp:184-214, filter/TestFIRFilter.cpp:19-23), the walk, and the noise
pool against an independent replay of libstdc++'s UNIFORM draws from numpy's MT19937 core -- which also
pins the (imaginary first) draw order this compiler gives std::complex<double>(dist(gen), dist(gen))."""
import numpy as np
import pytest


def test_sine_table_and_step_of_the_fir_test_tone(oracle):
    # TestFIRFilter: 30 kHz at 1 MHz, amplitude 1000 -> frac 0.03, 4096 entries (step 123 >= 16)
    for dt in ("CF64", "CF32"):
        code = getattr(oracle, dt)
        t, step = oracle.waveform_table(code, "SINE", 30e3, 1e6, ampl=1000.0)
        assert t.shape == (4096, 2) and step == 123
        ref = 1000.0 * np.exp(2j * np.pi * np.arange(4096) / 4096)
        got = t[:, 0].astype(np.float64) + 1j * t[:, 1].astype(np.float64)
        assert np.max(np.abs(got - ref)) < (1e-9 if dt == "CF64" else 1e-4)
    t, step = oracle.waveform_table(oracle.CI16, "SINE", 30e3, 1e6, ampl=1000.0)
    ref = 1000.0 * np.exp(2j * np.pi * np.arange(4096) / 4096)
    assert np.max(np.abs(t[:, 0] - np.trunc(ref.real))) <= 1                    # Type(double): truncation toward zero
    assert np.all(np.abs(t[:, 0].astype(np.float64)) <= np.abs(ref.real) + 1e-9)
    assert t[0, 0] == 1000 and t[1024, 1] == 1000 and abs(int(t[2048, 0]) + 1000) <= 1
    # real stream keeps the real part only
    tr, _ = oracle.waveform_table(oracle.F32, "SINE", 30e3, 1e6, ampl=2.0, offset=0.5)
    assert tr.shape == (4096, 1)
    assert np.allclose(tr[:, 0], 2.0 * np.cos(2 * np.pi * np.arange(4096) / 4096) + 0.5, atol=1e-6)


def test_const_ramp_square_tables(oracle):
    n = 4096
    i = np.arange(n)
    q = (i + 3 * n // 4) % n
    t, step = oracle.waveform_table(oracle.CF64, "CONST", 0.0, 1.0, ampl=2 - 1j, offset=0.25j)
    assert step == 0 and np.all(t[:, 0] == 2.0) and np.all(t[:, 1] == -0.75)
    t, _ = oracle.waveform_table(oracle.CF64, "RAMP", 30e3, 1e6)
    assert np.array_equal(t[:, 0], 2.0 * i / (n - 1) - 1.0) and np.array_equal(t[:, 1], 2.0 * q / (n - 1) - 1.0)
    t, _ = oracle.waveform_table(oracle.CI8, "SQUARE", 30e3, 1e6, ampl=100.0)
    assert np.array_equal(t[:, 0], np.where(i < n // 2, 0, 100)) and np.array_equal(t[:, 1], np.where(q < n // 2, 0, 100))
    # complex amplitude rotates: (0+1j) * (re + j im) = -im + j re
    t, _ = oracle.waveform_table(oracle.CI32, "SQUARE", 30e3, 1e6, ampl=7j)
    assert np.array_equal(t[:, 0], -7 * np.where(q < n // 2, 0, 1)) and np.array_equal(t[:, 1], 7 * np.where(i < n // 2, 0, 1))


def test_table_size_search_and_step(oracle):
    code = oracle.F32
    assert oracle.waveform_table(code, "SINE", 30e3, 1e6)[0].shape[0] == 4096
    # 1 kHz at 1 MHz: 4096 entries give a step of 4; the table doubles until the step reaches 16
    t, step = oracle.waveform_table(code, "SINE", 1e3, 1e6)
    assert (t.shape[0], step) == (16384, 16)
    # 1 Hz at 1 MHz: the search stops at the 2^20 limit with a step of 1
    t, step = oracle.waveform_table(code, "SINE", 1.0, 1e6)
    assert (t.shape[0], step) == (1 << 20, 1)
    # a frequency whose step rounds to zero even there is rejected (WaveformSource.cpp:209-212)
    with pytest.raises(ValueError):
        oracle.waveform_table(code, "SINE", 0.1, 1e6)
    # the resolution, when given, sizes the table instead of the frequency (:190)
    t, step = oracle.waveform_table(code, "SINE", 1e3, 1e6, res=10.0)
    assert (t.shape[0], step) == (1 << 20, 1049)
    # negative frequency: the step wraps as size_t and walks the table backwards
    t, step = oracle.waveform_table(oracle.CF64, "SINE", -30e3, 1e6)
    assert step == 2**64 - 123
    w = oracle.table_walk(oracle.CF64, t, 0, step, 10)
    assert np.array_equal(w, t[(-123 * np.arange(10)) % 4096])
    with pytest.raises(ValueError):
        oracle.waveform_table(code, "TRIANGLE", 30e3, 1e6)


@pytest.mark.parametrize("dt", ["F32", "CF32", "CF64", "I8", "CI16", "CI64"])
def test_table_walk(oracle, dt):
    code = getattr(oracle, dt)
    t, step = oracle.waveform_table(code, "RAMP", 30e3, 1e6, ampl=100.0)
    for index, st in ((0, step), (4095, 1), (2**63 + 5, 977), (17, 0)):
        w = oracle.table_walk(code, t, index, st, 9001)
        assert np.array_equal(w, t[((index + st * np.arange(9001, dtype=object)) % 4096).astype(np.int64)])


def _std_uniform_pairs(seed, count, lo, hi):
    """count draws of std::uniform_real_distribution<double>(lo, hi) on std::mt19937(seed): libstdc++'s
    generate_canonical<double, 53> takes two 32-bit outputs, low word first."""
    raw = np.frombuffer(np.random.RandomState(seed).bytes(8 * count), dtype="<u4").astype(np.float64)
    canon = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    canon = np.where(canon >= 1.0, np.nextafter(1.0, 0.0), canon)
    return canon * (hi - lo) + lo


def test_noise_uniform_pool_replayed_from_mt19937(oracle):
    seed, mean, b = 20240607, 0.25, 2.0
    _, pool = oracle.noise_stream(oracle.CF64, "UNIFORM", mean, b, seed, [1])
    draws = _std_uniform_pairs(seed, 2 * 4096, mean - b, mean + b)
    assert np.array_equal(pool[:, 1], draws[0::2]), "imaginary part = first draw of each pair (g++ argument order)"
    assert np.array_equal(pool[:, 0], draws[1::2])
    assert pool.min() >= mean - b and pool.max() < mean + b


@pytest.mark.parametrize("wave,mean,b", [("NORMAL", 0.5, 2.0), ("UNIFORM", -1.0, 0.5), ("LAPLACE", 0.0, 0.9), ("POISSON", 4.0, 1.0)])
def test_noise_pool_statistics_and_stream_structure(oracle, wave, mean, b):
    work = [100, 5000, 3 * 4096 + 5]
    s1, pool = oracle.noise_stream(oracle.CF64, wave, mean, b, 99, work)
    s2, _ = oracle.noise_stream(oracle.CF64, wave, mean, b, 99, work)
    s3, _ = oracle.noise_stream(oracle.CF64, wave, mean, b, 100, work)
    assert np.array_equal(s1, s2) and not np.array_equal(s1, s3)
    assert np.all(np.isfinite(pool))
    if wave == "NORMAL":
        assert abs(pool.mean() - mean) < 0.1 and abs(pool.std() - b) < 0.1
    elif wave == "UNIFORM":
        assert abs(pool.mean() - mean) < 0.05 and abs(pool.std() - b / np.sqrt(3)) < 0.02
    elif wave == "POISSON":
        assert np.array_equal(pool, np.round(pool)) and pool.min() >= 0 and abs(pool.mean() - mean) < 0.2
    # each work() is a contiguous run of the pool, entered at a new position (NoiseSource.cpp:108-113)
    pos, prev_end = 0, 0
    for n in work:
        chunk = s1[pos:pos + n]
        starts = np.flatnonzero(np.all(pool == chunk[0], axis=1))
        assert any(np.array_equal(chunk, pool[(st + np.arange(n)) % 4096]) for st in starts)
        pos += n
    # scaling and offset are applied when the pool is filled, per element type (setElem)
    si, pool_i = oracle.noise_stream(oracle.CI16, wave, mean, b, 99, work, ampl=100.0, offset=3 - 2j)
    assert np.array_equal(pool_i[:, 0], np.trunc(100.0 * pool[:, 0] + 3).astype(np.int16))
    assert np.array_equal(pool_i[:, 1], np.trunc(100.0 * pool[:, 1] - 2).astype(np.int16))
    # a setter call between two work() calls redraws the pool from the running generator
    s4, pool4 = oracle.noise_stream(oracle.CF64, wave, mean, b, 99, work, refill_before=[0, 1, 0])
    assert np.array_equal(s4[:100], s1[:100]) and not np.array_equal(pool4, pool)


def test_noise_unknown_wave(oracle):
    with pytest.raises(ValueError):
        oracle.noise_stream(oracle.F32, "PINK", 0.0, 1.0, 1, [4])


def test_source_oracle_regression_fixtures(oracle):
    """tests/golden/source_*.npz (made by tests/golden/make_source_golden.py): the oracle must keep producing them."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_source_golden", os.path.join(here, "make_source_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    wave = np.load(os.path.join(here, "source_waveform.npz"))
    for name, dt, kind, freq, rate, res, ampl, off in g.WAVE_CASES:
        code = getattr(oracle, dt)
        table, step = oracle.waveform_table(code, kind, freq, rate, res=res, ampl=ampl, offset=off)
        assert [table.shape[0], step] == [int(v) for v in wave[name + "_entries_step"]], name
        assert np.array_equal(table[:64], wave[name + "_table_head"]), name
        assert np.array_equal(oracle.table_walk(code, table, 12345, step, 3000), wave[name + "_stream"]), name
    noise = np.load(os.path.join(here, "source_noise.npz"))
    for name, dt, kind, mean, b, ampl, off in g.NOISE_CASES:
        code = getattr(oracle, dt)
        stream, pool = oracle.noise_stream(code, kind, mean, b, g.NOISE_SEED, g.NOISE_WORK, refill_before=[0, 0, 1], ampl=ampl, offset=off)
        assert np.array_equal(pool[:256], noise[name + "_pool"]), name
        assert np.array_equal(np.concatenate([stream[:256], stream[-300:]]), noise[name + "_stream"]), name
