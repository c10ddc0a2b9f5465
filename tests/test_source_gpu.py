"""GPU parity of the stream sources (/comms/waveform_source, /comms/noise_source): the device table
walk through the C-ABI (b200c_table_source) and the two blocks against oracle/source_oracle.cpp,
bit-exact for every element type -- the device only moves table entries, the tables are host work
as in the reference.  Ends with the reference's FIR test topology with its own source block in front
(filter/TestFIRFilter.cpp:19-53)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = ["F32", "CF32", "F64", "CF64", "I8", "CI8", "I16", "CI16", "I32", "CI32", "I64", "CI64"]
NAMES = {"F32": "float32", "CF32": "complex_float32", "F64": "float64", "CF64": "complex_float64", "I8": "int8",
         "CI8": "complex_int8", "I16": "int16", "CI16": "complex_int16", "I32": "int32", "CI32": "complex_int32",
         "I64": "int64", "CI64": "complex_int64"}


def _table(oracle, code, entries, rng):
    sc = oracle.scalar_np(code)
    nc = 2 if code & 1 else 1
    if np.issubdtype(sc, np.integer):
        info = np.iinfo(sc)
        return rng.integers(info.min, info.max, size=(entries, nc), endpoint=True).astype(sc)
    return rng.standard_normal((entries, nc)).astype(sc)


@pytest.mark.parametrize("dt", ALL)
def test_table_walk_matches_oracle(oracle, cuda_device, dt):
    import torch
    from pothoscomms_b200 import handles
    code = getattr(oracle, dt)
    rng = np.random.default_rng(code + 1)
    for entries in (1, 2, 4096, 1 << 17):            # 2^17 entries: beyond the shared-memory staging for every type
        t = _table(oracle, code, entries, rng)
        d_t = torch.from_numpy(t).cuda()
        for n in (1, 7, 4099, 100003):
            for index, step in ((0, 1), (4090, 123), (2**63 + 11, 2**64 - 123), (5, 0), (12345678901234, 977)):
                y = handles.table_source(code, d_t, index, step, n).cpu().numpy()
                assert np.array_equal(y, oracle.table_walk(code, t, index, step, n)), (dt, entries, n, index, step)


@pytest.mark.parametrize("dt", ["CI8", "I16", "CI16", "CF32", "CF64"])
def test_table_walk_unaligned_output(oracle, cuda_device, dt):
    """outputs that start between 16-byte boundaries (an output port mid-buffer), and complex elements aligned
    only to their scalar"""
    import torch
    from pothoscomms_b200 import handles
    code = getattr(oracle, dt)
    rng = np.random.default_rng(3)
    t = _table(oracle, code, 4096, rng)
    d_t = torch.from_numpy(t).cuda()
    nc = t.shape[1]
    n = 10001
    for off in (1, 2, 3):
        whole = torch.zeros((n + off, nc), dtype=d_t.dtype, device="cuda")
        handles.table_source(code, d_t, 77, 19, n, out=whole[off:])
        assert np.array_equal(whole[off:].cpu().numpy(), oracle.table_walk(code, t, 77, 19, n))
        assert not whole[:off].any()
    if nc == 2:   # scalar-aligned complex stream: start one scalar in
        flat = torch.zeros(2 * n + 1, dtype=d_t.dtype, device="cuda")
        view = flat[1:].view(n, 2)
        handles.table_source(code, d_t, 5, 3, n, out=view)
        assert np.array_equal(view.cpu().numpy(), oracle.table_walk(code, t, 5, 3, n)) and not flat[0].item()


def test_table_source_rejects_bad_tables(oracle, cuda_device):
    import torch
    from pothoscomms_b200 import InvalidArgumentError, handles
    t = torch.zeros((1000, 2), dtype=torch.float32, device="cuda")
    with pytest.raises(InvalidArgumentError, match="power of two"):
        handles.table_source(oracle.CF32, t, 0, 1, 16)


@pytest.mark.parametrize("dt", ["CF64", "CF32", "CI16", "F32", "I8", "CI64"])
def test_waveform_source_block(oracle, cuda_device, dt):
    """the registered calls of waveform/WaveformSource.cpp:79-90 and the stream of work() (:98-108) over
    several calls, with setters in between (the phase carries over, the table and step change)"""
    from pothoscomms_b200 import blocks
    code = getattr(oracle, dt)
    src = blocks.make("/comms/waveform_source", NAMES[dt], out_bytes=1 << 20)
    for c in ("setWaveform getWaveform setOffset getOffset setAmplitude getAmplitude setFrequency getFrequency "
              "setSampleRate getSampleRate setResolution getResolution").split():
        assert src.has_call(c), c
    assert (src.call("getWaveform"), src.call("getFrequency"), src.call("getSampleRate"), src.call("getResolution")) == ("CONST", 0.0, 1.0, 0.0)
    assert (src.call("getAmplitude"), src.call("getOffset")) == (1 + 0j, 0j)
    ampl = 100.0 if "I" in dt else 1.0
    src.call("setAmplitude", ampl)                      # TestFIRFilter.cpp:20-23
    src.call("setWaveform", "SINE")
    src.call("setFrequency", 30e3)
    src.call("setSampleRate", 1e6)
    src.activate()
    table, step = oracle.waveform_table(code, "SINE", 30e3, 1e6, ampl=ampl)
    y = src.run_source(nwork=3, elems=5000)
    assert np.array_equal(y, oracle.table_walk(code, table, 0, step, 15000))
    index = 15000 * step
    for wave, freq, res, off in (("RAMP", 1e3, 0.0, 0.0), ("SQUARE", -45e3, 0.0, 1 - 2j), ("SINE", 1e3, 10.0, 0.0), ("CONST", 0.0, 0.0, 0.5)):
        src.call("setWaveform", wave)
        src.call("setFrequency", freq)
        src.call("setResolution", res)
        src.call("setOffset", off)
        table, step = oracle.waveform_table(code, wave, freq, 1e6, res=res, ampl=ampl, offset=off)
        y = src.run_source(nwork=2, elems=4097)
        assert np.array_equal(y, oracle.table_walk(code, table, index, step, 2 * 4097)), (dt, wave)
        index += 2 * 4097 * step
    y = src.run_source(nwork=1)                        # the whole output buffer in one work()
    assert y.shape[0] == (1 << 20) // y.itemsize // y.shape[1]
    assert np.array_equal(y, oracle.table_walk(code, table, index, step, y.shape[0]))


def test_waveform_source_errors(oracle, cuda_device):
    from pothoscomms_b200 import blocks
    src = blocks.make("/comms/waveform_source", "complex_float32")
    src.call("setWaveform", "TRIANGLE")                # before activate() a setter only stores (:186)
    with pytest.raises(blocks.InvalidArgumentException, match="unknown waveform"):
        src.activate()
    src.call("setWaveform", "SINE")
    src.call("setSampleRate", 1e6)
    with pytest.raises(blocks.InvalidArgumentException, match="step size not achievable"):
        src.call("setFrequency", 0.1)                  # :209-212


@pytest.mark.parametrize("wave,mean,b", [("NORMAL", 0.0, 1.0), ("UNIFORM", 0.5, 2.0), ("LAPLACE", 0.0, 0.9), ("POISSON", 4.0, 1.0)])
@pytest.mark.parametrize("dt", ["CF32", "CI16", "F64"])
def test_noise_source_block_replays_the_oracle(oracle, cuda_device, monkeypatch, wave, mean, b, dt):
    """waveform/NoiseSource.cpp with a fixed seed: pool drawn on activate(), one random entry point per work()
    (:108), a setter redraws the pool from the running generator (:188-226)"""
    from pothoscomms_b200 import blocks
    code = getattr(oracle, dt)
    seed = 0xB200 + code
    monkeypatch.setenv("B200C_NOISE_SEED", str(seed))
    ampl = 50.0 if "I" in dt else 1.0
    src = blocks.make("/comms/noise_source", NAMES[dt], out_bytes=1 << 20)
    for c in "setWaveform getWaveform setOffset getOffset setAmplitude getAmplitude setMean getMean setB getB".split():
        assert src.has_call(c), c
    assert (src.call("getWaveform"), src.call("getMean"), src.call("getB")) == ("NORMAL", 0.0, 1.0)   # :76-83
    src.call("setWaveform", wave)
    src.call("setMean", mean)
    src.call("setB", b)
    src.call("setAmplitude", ampl)
    src.call("setOffset", 1 - 1j)
    src.activate()
    y1 = src.run_source(nwork=2, elems=3000)
    src.call("setB", b)                                 # any setter redraws the pool
    y2 = src.run_source(nwork=1, elems=3 * 4096 + 5)
    ref, _ = oracle.noise_stream(code, wave, mean, b, seed, [3000, 3000, 3 * 4096 + 5], refill_before=[0, 0, 1], ampl=ampl, offset=1 - 1j)
    assert np.array_equal(np.concatenate([y1, y2]), ref)


def test_noise_source_default_seed_is_random_and_rejects_unknown_wave(oracle, cuda_device, monkeypatch):
    from pothoscomms_b200 import blocks
    monkeypatch.delenv("B200C_NOISE_SEED", raising=False)
    outs = []
    for _ in range(2):
        src = blocks.make("/blocks/noise_source", "complex_float32", out_bytes=1 << 16)
        src.activate()
        y = src.run_source(nwork=1)
        assert y.shape == (8192, 2) and abs(float(y.std()) - 1.0) < 0.05 and abs(float(y.mean())) < 0.05   # NORMAL(0, 1) by default
        assert np.array_equal(y[:4096], y[4096:])       # the pool repeats every 4096 elements (fast mode)
        outs.append(y)
    assert not np.array_equal(outs[0], outs[1])        # std::random_device seeds each block (:84)
    with pytest.raises(blocks.InvalidArgumentException, match="unknown waveform"):
        src.call("setWaveform", "PINK")


@pytest.mark.parametrize("dtype", ["complex_float64", "complex_int16"])
def test_fir_test_topology_with_its_own_source(oracle, cuda_device, dtype):
    """filter/TestFIRFilter.cpp:19-82 with every block of the chain on the device: /comms/waveform_source (SINE,
    30 kHz at 1 MHz, amplitude 1000) -> 4096 elements -> /comms/fir_filter (101 complex taps) -> /comms/signal_probe
    in RMS mode; assertion rms > 0.1 * amplitude, plus parity of every stage with the oracle."""
    from pothoscomms_b200 import blocks
    from pothoscomms_b200 import workloads as wl
    code = oracle.DTYPE_CODES[dtype]
    amplitude, rate, freq = 1000.0, 1e6, 30e3
    src = blocks.make("/comms/waveform_source", dtype)
    src.call("setAmplitude", amplitude)
    src.call("setWaveform", "SINE")
    src.call("setFrequency", freq)
    src.call("setSampleRate", rate)
    src.activate()
    x = src.run_source(nwork=1, elems=4096)            # /blocks/finite_release: setTotalElements(4096)
    table, step = oracle.waveform_table(code, "SINE", freq, rate, ampl=amplitude)
    assert np.array_equal(x, oracle.table_walk(code, table, 0, step, 4096))
    for decim, interp in ((1, 1), (2, 3), (3, 2)):
        out_rate = rate * interp / decim
        taps = wl.complex_bandpass(101, freq / out_rate, 0.1 * rate / out_rate)
        fir = blocks.make("/comms/fir_filter", dtype, "COMPLEX")
        fir.call("setDecimation", decim)
        fir.call("setInterpolation", interp)
        fir.call("setTaps", taps)
        fir.activate()
        y = fir.push_through(x)
        y_ref, _, _ = oracle.fir(code, True, taps, decim, interp, x)
        assert np.array_equal(y, y_ref) if "int" in dtype else np.allclose(y, y_ref, rtol=0, atol=1e-9 * amplitude)
        probe = blocks.make("/comms/signal_probe", dtype)
        probe.call("setMode", "RMS")
        probe.call("setWindow", 1024)
        probe.activate()
        probe.feed(y[:1024])
        probe.run()
        rms = probe.call("value").real
        assert abs(rms - oracle.probe(code, "RMS", y[:1024]).real) <= 1e-9 * rms
        assert rms > 0.1 * amplitude                    # POTHOS_TEST_TRUE(rms > (0.1*amplitude)), :78
