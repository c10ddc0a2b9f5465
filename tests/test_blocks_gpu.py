"""The reference's own block tests, re-run against the B200 block layer.

Each test follows the named reference test: same registry path, same calls, same signal,
same assertion -- plus a stricter one: the block's output must equal the CPU oracle fed the
identical stream (bit-exact for int16, 1e-5 of RMS for float).  A single-block harness
(pothoscomms_b200/blocks/Harness.cpp) stands in for Pothos::Topology + feeder/collector.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sinc_complex_bandpass(ntaps, lo, hi):
    """Stand-in for FIRDesigner SINC/COMPLEX_BAND_PASS (Spuce is not available): Hann-windowed."""
    from pothoscomms_b200 import workloads as wl
    return wl.complex_bandpass(ntaps, (lo + hi) / 2, (hi - lo) / 2)


def rel_rms(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.sqrt(np.mean((a - b) ** 2) / max(np.mean(b ** 2), 1e-300))


@pytest.mark.parametrize("dtype", ["complex_float64", "complex_int16", "complex_float32"])
def test_fir_filter(oracle, cuda_device, dtype):
    """filter/TestFIRFilter.cpp:59-82 -- 30 kHz complex sine (amp 1000, fs 1 MHz), 4096 samples,
    101-tap complex band-pass delivered via setTaps while waitTaps is armed, decim x interp in
    1..3 x 1..3; assertion rms > 0.1*amplitude (:78)."""
    from pothoscomms_b200 import blocks
    code = oracle.DTYPE_CODES[dtype]
    amplitude, rate, freq = 1000.0, 1e6, 30e3
    n = 4096
    t = np.arange(n)
    wave = amplitude * np.exp(2j * np.pi * freq / rate * t)
    x = oracle.to_raw(np.rint(wave) if "int" in dtype else wave, code)
    for decim in (1, 2, 3):
        for interp in (1, 2, 3):
            out_rate = rate * interp / decim
            taps = sinc_complex_bandpass(101, (freq - 0.1 * rate) / out_rate, (freq + 0.1 * rate) / out_rate)   # designer runs at out_rate (:33-40)
            f = blocks.make("/comms/fir_filter", dtype, "COMPLEX")
            f.call("setDecimation", decim)
            f.call("setInterpolation", interp)
            f.call("setWaitTaps", True)
            f.activate()
            assert f.input_domain == "b200c_hbm"       # samples live in HBM between work() calls
            f.feed(x)
            f.run()
            assert f.collect().shape[0] == 0           # waitTaps: nothing flows before taps arrive (:209)
            f.call("setTaps", taps)                    # designer.tapsChanged -> filter.setTaps (:48)
            f.run()
            y = f.collect()
            y_ref, cons, prod = oracle.fir(code, True, taps, decim, interp, x)
            assert y.shape[0] == prod and f.total_consumed == cons
            if "int" in dtype:
                assert np.array_equal(y, y_ref)
            else:
                assert rel_rms(y, y_ref) < (1e-5 if dtype == "complex_float32" else 1e-13)
            yc = y[:, 0].astype(np.float64) + 1j * y[:, 1].astype(np.float64)
            rms = np.sqrt(np.mean(np.abs(yc) ** 2))    # /comms/signal_probe RMS mode
            assert rms > 0.1 * amplitude                # POTHOS_TEST_TRUE(rms > (0.1*amplitude)), :78


def test_registered_calls_and_getters(oracle, cuda_device):
    """The 12 registered calls of filter/FIRFilter.cpp:113-124, their defaults and errors."""
    from pothoscomms_b200 import blocks
    f = blocks.make("/blocks/fir_filter", "complex_float32", "REAL")   # legacy alias, :388-389
    for name in ("setTaps", "getTaps", "setDecimation", "getDecimation", "setInterpolation", "getInterpolation",
                 "setWaitTaps", "getWaitTaps", "setFrameStartId", "getFrameStartId", "setFrameEndId", "getFrameEndId"):
        assert f.has_call(name), name
    assert list(f.call("getTaps")) == [1.0]            # ctor default, :125
    assert f.call("getDecimation") == 1 and f.call("getInterpolation") == 1
    assert f.call("getWaitTaps") is False
    assert f.call("getFrameStartId") == "" and f.call("getFrameEndId") == ""
    with pytest.raises(blocks.InvalidArgumentException, match="taps cannot be empty"):
        f.call("setTaps", np.zeros(0))
    with pytest.raises(blocks.InvalidArgumentException, match="decimation cannot be 0"):
        f.call("setDecimation", 0)
    with pytest.raises(blocks.InvalidArgumentException, match="interpolation cannot be 0"):
        f.call("setInterpolation", 0)
    f.call("setTaps", np.arange(1, 8, dtype=float))
    f.call("setDecimation", 3)
    f.call("setInterpolation", 2)
    f.call("setFrameStartId", "START")
    f.call("setFrameEndId", "END")
    f.call("setWaitTaps", True)
    assert list(f.call("getTaps")) == [1, 2, 3, 4, 5, 6, 7]
    assert (f.call("getDecimation"), f.call("getInterpolation")) == (3, 2)
    assert f.call("getFrameStartId") == "START" and f.call("getFrameEndId") == "END" and f.call("getWaitTaps") is True
    c = blocks.make("/comms/fir_filter", "complex_int16", "COMPLEX")
    c.call("setTaps", np.array([0.5 + 0.25j, -0.125j]))
    assert np.array_equal(c.call("getTaps"), np.array([0.5 + 0.25j, -0.125j]))


def test_streaming_reserve_and_history(oracle, cuda_device):
    """Dribbling input: work() asks for _inputRequire = M + K - 1 via setReserve (:251-255),
    keeps K-1 elements as history in the circular buffer (:304-307) and the concatenated
    output equals the one-shot oracle."""
    from pothoscomms_b200 import blocks
    rng = np.random.default_rng(1)
    taps = rng.standard_normal(64)
    x = rng.standard_normal((20000, 2)).astype(np.float32)
    f = blocks.make("/comms/fir_filter", "complex_float32", "REAL")
    f.call("setTaps", taps)
    f.call("setDecimation", 2)
    f.activate()
    f.feed(x[:10])
    f.run()
    assert f.collect().shape[0] == 0 and f.reserve == 2 + 64 - 1
    outs, pos = [], 10
    for piece in (55, 1, 1000, 3, 7777, 20000):
        f.feed(x[pos: pos + piece])
        pos = min(pos + piece, x.shape[0])
        f.run()
        outs.append(f.collect())
    y = np.concatenate(outs)
    y_ref, cons, prod = oracle.fir(oracle.CF32, False, taps, 2, 1, x)
    assert f.total_consumed == cons and y.shape[0] == prod
    assert rel_rms(y, y_ref) < 1e-5


def test_burst_mode_impulse_response_through_fft(oracle, cuda_device):
    """filter/TestFIRDesigner.cpp:126-235 -- a 1024-element burst (START label carrying the length)
    holding an impulse of height 1024 at its LAST index goes through the FIR in burst mode, which
    must flush exactly fftSize outputs (:183), then through /comms/fft; the power spectrum is
    checked with the reference's -30 dB pass / -80 dB stop masks (:97-124) for LOW_PASS taps."""
    from pothoscomms_b200 import blocks
    fft_size, ntaps, rate = 1024, 101, 1e6
    lower = 1.5e5
    # Kaiser-windowed sinc low-pass (Spuce is unavailable; any taps meeting the mask will do)
    t = np.arange(ntaps) - (ntaps - 1) / 2
    taps = 2 * (lower / rate) * np.sinc(2 * (lower / rate) * t) * np.kaiser(ntaps, 12.0)
    taps /= taps.sum()
    for dtype in ("complex_float64", "complex_float32"):
        code = oracle.DTYPE_CODES[dtype]
        impulse = np.zeros(fft_size, dtype=complex)
        impulse[-1] = fft_size
        x = oracle.to_raw(impulse, code)
        fir = blocks.make("/comms/fir_filter", dtype, "COMPLEX")
        fir.call("setDecimation", 1)
        fir.call("setInterpolation", 1)
        fir.call("setWaitTaps", True)
        fir.call("setFrameStartId", "START")
        fir.activate()
        fir.post_label("START", 0, data=fft_size)      # vector_source setStartId("START"), :150
        fir.feed(x)
        fir.call("setTaps", taps.astype(complex))
        fir.run()
        y = fir.collect()
        assert y.shape[0] == fft_size                    # POTHOS_TEST_EQUAL(buff.elements(), fftSize), :183
        y_ref, _, _ = oracle.fir(code, True, taps.astype(complex), 1, 1, x, zero_tail=True)
        assert rel_rms(y, y_ref) < (1e-5 if dtype == "complex_float32" else 1e-13)

        fft = blocks.make("/comms/fft", dtype, fft_size, False)
        fft.activate()
        fft.feed(y)
        fft.run()
        bins = fft.collect()
        assert bins.shape[0] == fft_size
        spec = bins[:, 0].astype(np.float64) + 1j * bins[:, 1].astype(np.float64)
        # fftPowerSpectrum of the reference (:60-95): normalise by fftSize, reorder to [-fs/2, fs/2)
        power = 20 * np.log10(np.maximum(np.abs(np.fft.fftshift(spec)) / fft_size, 1e-12))

        def level(freq):
            return power[int(fft_size * ((freq + rate / 2) / rate))]
        assert level(0.0) > -30.0                                       # PASS @ middle of pass band
        assert level(-(lower + rate / 2) / 2) < -80.0                   # STOP @ middle of lower stop band
        assert level(+(lower + rate / 2) / 2) < -80.0                   # STOP @ middle of upper stop band


def test_frame_end_label_and_back_to_back_bursts(oracle, cuda_device):
    from pothoscomms_b200 import blocks
    rng = np.random.default_rng(6)
    taps = rng.standard_normal(33) + 1j * rng.standard_normal(33)
    f = blocks.make("/comms/fir_filter", "complex_int16", "COMPLEX")
    f.call("setTaps", taps * 0.1)
    f.call("setFrameEndId", "EOB")
    f.activate()
    bursts = [rng.integers(-20000, 20000, size=(n, 2), dtype=np.int16) for n in (500, 64, 1000)]
    pos, got = 0, []
    for b in bursts:
        f.post_label("EOB", pos + b.shape[0] - 1)       # end label on the burst's last element (:226-229)
        f.feed(b)
        f.run()
        got.append(f.collect())
        pos += b.shape[0]
    for b, y in zip(bursts, got):
        y_ref, cons, prod = oracle.fir(oracle.CI16, True, taps * 0.1, 1, 1, b, zero_tail=True)
        assert prod == b.shape[0] and np.array_equal(y, y_ref)


def test_propagate_labels_rescales_index_and_rxrate(oracle, cuda_device):
    """propagateLabels (:311-323): index*L/M (Label::toAdjusted) and "rxRate" payload * L / M."""
    from pothoscomms_b200 import blocks
    f = blocks.make("/comms/fir_filter", "float32", "REAL")
    f.call("setDecimation", 2)
    f.call("setInterpolation", 3)
    f.activate()
    f.post_label("rxRate", 10, data=1e6)
    f.post_label("marker", 100, data=7)
    f.feed(np.ones((1000, 1), dtype=np.float32))
    f.run()
    labels = {l["id"]: l for l in f.out_labels()}
    assert labels["rxRate"]["index"] == 10 * 3 // 2 and labels["rxRate"]["data"] == pytest.approx(1.5e6)
    assert labels["marker"]["index"] == 100 * 3 // 2 and labels["marker"]["data"] == 7


def test_fft_float(oracle, cuda_device):
    """fft/TestFFT.cpp:11-82 through the block."""
    from pothoscomms_b200 import blocks
    inp = np.array([0.4 + 0.6j, -0.7 + 0.6j, -0.2 + 0.8j, 0.9 + 0.2j])
    res = np.array([0.4 + 2.2j, 1.0 + 1.4j, 0.0 + 0.6j, 0.2 - 1.8j])
    fft = blocks.make("/comms/fft", "complex_float32", 4, False)
    fft.activate()
    y = fft.push_through(oracle.to_raw(inp.astype(np.complex64), oracle.CF32))
    assert y.shape[0] == 4
    assert np.all(np.abs(y[:, 0] - res.real) < 0.01) and np.all(np.abs(y[:, 1] - res.imag) < 0.01)
    ifft = blocks.make("/comms/fft", "complex_float32", 4, True)
    ifft.activate()
    y = ifft.push_through(oracle.to_raw(res.astype(np.complex64), oracle.CF32))
    assert np.all(np.abs(y[:, 0] - inp.real * 4) < 0.01) and np.all(np.abs(y[:, 1] - inp.imag * 4) < 0.01)


def test_fft_short(oracle, cuda_device):
    """fft/TestFFT.cpp:84-158 through the block: forward == golden/N, inverse(golden) == input, exactly."""
    from pothoscomms_b200 import blocks
    inp = oracle.to_raw(np.round(np.array([0.4 + 0.6j, -0.7 + 0.6j, -0.2 + 0.8j, 0.9 + 0.2j]) * 1000), oracle.CI16)
    res = oracle.to_raw(np.round(np.array([0.4 + 2.2j, 1.0 + 1.4j, 0.0 + 0.6j, 0.2 - 1.8j]) * 1000), oracle.CI16)
    fft = blocks.make("/comms/fft", "complex_int16", 4, False)
    fft.activate()
    assert np.array_equal(fft.push_through(inp), res // 4)
    ifft = blocks.make("/comms/fft", "complex_int16", 4, True)
    ifft.activate()
    assert np.array_equal(ifft.push_through(res), inp)


def test_fft_block_batches_and_reserve(oracle, cuda_device):
    """The block's reserve is numBins (fft/FFT.cpp:50); a partial frame waits; many frames go in one work()."""
    from pothoscomms_b200 import blocks
    rng = np.random.default_rng(2)
    n = 1024
    x = rng.standard_normal((10 * n + 100, 2)).astype(np.float32)
    fft = blocks.make("/comms/fft", "complex_float32", n, False)
    fft.activate()
    assert fft.reserve == n
    fft.feed(x[:100])
    fft.run()
    assert fft.collect().shape[0] == 0
    fft.feed(x[100:])
    fft.run()
    y = fft.collect()
    assert y.shape[0] == 10 * n and fft.work_calls <= 3          # batched, not one launch per transform
    ref = oracle.ref_fft(oracle.CF32, n, False, x[: 10 * n]) if oracle.have_ref() else oracle.fft(oracle.CF32, n, False, x[: 10 * n])
    assert rel_rms(y, ref) < 1e-5
    assert fft.has_call("setInverse")
    fft.call("setInverse", True)                                  # north_star's setInverse (an addition)
    assert fft.call("getInverse") is True


def test_fir_then_fft_chain_stays_consistent(oracle, cuda_device):
    """fir -> fft as in TestFIRDesigner's topology, streaming mode, int16 bit-exact end to end."""
    from pothoscomms_b200 import blocks
    from pothoscomms_b200 import workloads as wl
    taps, tt = wl.config_taps("c2")
    x = wl.tone_noise_numpy(oracle.CI16, 127 + 16 * 4096, seed=5)
    fir = blocks.make("/comms/fir_filter", "complex_int16", tt)
    fir.call("setTaps", taps)
    fir.activate()
    y = fir.push_through(x)
    y_ref, _, prod = oracle.fir(oracle.CI16, True, taps, 1, 1, x)
    assert prod == 16 * 4096 and np.array_equal(y, y_ref)
    fft = blocks.make("/comms/fft", "complex_int16", 4096, False)
    fft.activate()
    Y = fft.push_through(y)
    assert np.array_equal(Y, oracle.fft(oracle.CI16, 4096, False, y_ref))


def _designer_matrix():
    from _designer_cases import reference_matrix
    return list(reference_matrix())


@pytest.mark.parametrize("filter_type,band_type", _designer_matrix())
def test_fir_designer(oracle, cuda_device, filter_type, band_type):
    """filter/TestFIRDesigner.cpp:126-275, the whole topology: /comms/fir_designer emits
    "tapsChanged" into the filter's setTaps slot (:176) while the filter waits for taps, the
    START-labelled impulse burst goes FIR -> FFT in HBM, and the power spectrum must meet the
    reference's pass/stop mask for every (filter type, band type) pair it tests (:255-273)."""
    from pothoscomms_b200 import blocks
    from _designer_cases import configure, mask_points
    fft_size, rate, lower, upper = 1024, 1e6, 1.5e5, 3.0e5
    dtype = "complex_float64"                            # :141
    code = oracle.DTYPE_CODES[dtype]
    impulse = np.zeros(fft_size, dtype=complex)
    impulse[-1] = fft_size                               # :145-146
    fir = blocks.make("/comms/fir_filter", dtype, "COMPLEX")
    fir.call("setDecimation", 1)
    fir.call("setInterpolation", 1)
    fir.call("setWaitTaps", True)
    fir.call("setFrameStartId", "START")
    designer = blocks.make("/comms/fir_designer")
    configure(designer, filter_type, band_type, rate, lower, upper, 101)
    designer.connect("tapsChanged", fir, "setTaps")      # topology.connect(designer, "tapsChanged", filter, "setTaps")
    fir.activate()
    fir.post_label("START", 0, data=fft_size)
    fir.feed(oracle.to_raw(impulse, code))
    fir.run()
    assert fir.collect().shape[0] == 0                   # still waiting for taps
    designer.activate()                                  # emits the taps -> filter.setTaps (real taps convert to complex)
    taps, count = designer.last_signal()
    assert count == 1
    assert np.allclose(fir.call("getTaps"), taps.astype(complex))
    fir.run()
    y = fir.collect()
    assert y.shape[0] == fft_size                        # :183
    fft = blocks.make("/comms/fft", dtype, fft_size, False)
    fft.activate()
    fft.feed(y)
    fft.run()
    bins = fft.collect()
    spec = bins[:, 0] + 1j * bins[:, 1]
    power = 20 * np.log10(np.maximum(np.abs(np.fft.fftshift(spec)) / fft_size, 1e-15))
    for is_pass, freq in mask_points(band_type, rate, lower, upper):
        level = power[int(fft_size * ((freq + rate / 2) / rate))]
        assert (level > -30.0) if is_pass else (level < -80.0), (filter_type, band_type, freq, level)


@pytest.mark.parametrize("dtype", ["float64", "float32", "int64", "int32", "int16", "int8"])
def test_scale(oracle, cuda_device, dtype):
    """math/TestScale.cpp:14-70 -- 13 points 10*i through /comms/scale for factors -1 ... 1; each output
    within 1 of Type(input * factor) (:52), and equal to the oracle."""
    from pothoscomms_b200 import blocks
    code = oracle.DTYPE_CODES[dtype]
    sc = oracle.scalar_np(code)
    x = (10 * np.arange(13)).astype(sc).reshape(-1, 1)
    for i in range(5):
        factor = i / 2.0 - 1.0
        b = blocks.make("/comms/scale", dtype)
        assert all(b.has_call(c) for c in ("setFactor", "getFactor", "setLabelId", "getLabelId"))   # Scale.cpp:32-35
        b.call("setFactor", factor)
        assert b.call("getFactor") == factor
        b.activate()
        y = b.push_through(x)
        assert y.shape == x.shape
        expected = np.trunc(x.astype(np.float64) * factor) if "int" in dtype else x.astype(np.float64) * factor
        assert np.all(np.abs(y.astype(np.float64) - expected) <= 1)
        assert np.array_equal(y, oracle.scale(code, factor, x))


def test_scale_label_changes_factor_mid_stream(oracle, cuda_device):
    """math/Scale.cpp:86-108: a label with the configured id carries a new factor; it takes effect exactly
    at the label's element."""
    from pothoscomms_b200 import blocks
    code = oracle.CI16
    rng = np.random.default_rng(8)
    x = rng.integers(-20000, 20000, size=(3000, 2)).astype(np.int16)
    b = blocks.make("/comms/scale", "complex_int16")
    b.call("setFactor", 0.5)
    b.call("setLabelId", "gain")
    assert b.call("getLabelId") == "gain"
    b.activate()
    b.post_label("gain", 1000, data=0.25)
    b.post_label("other", 1500, data=9.0)             # a foreign id is ignored
    y = b.push_through(x)
    ref = np.concatenate([oracle.scale(code, 0.5, x[:1000]), oracle.scale(code, 0.25, x[1000:])])
    assert np.array_equal(y, ref)
    assert b.call("getFactor") == 0.25


@pytest.mark.parametrize("dtype", ["float64", "float32", "int64", "int32", "int16", "int8"])
def test_rotate(oracle, cuda_device, dtype):
    """math/TestRotate.cpp:14-71 -- (10 f, -20 f) through /comms/rotate at phases 0, pi/2, pi, 3pi/2"""
    from pothoscomms_b200 import blocks
    cdtype = "complex_" + dtype
    code = oracle.DTYPE_CODES[cdtype]
    sc = oracle.scalar_np(code)
    f = np.arange(13)
    x = np.stack([10 * f, -20 * f], axis=1).astype(sc)
    for i in range(4):
        phase = i * np.pi / 2
        b = blocks.make("/comms/rotate", cdtype)
        assert all(b.has_call(c) for c in ("setPhase", "getPhase", "setLabelId", "getLabelId"))    # Rotate.cpp:63-66
        b.call("setPhase", phase)
        b.activate()
        y = b.push_through(x)
        z = (x[:, 0].astype(np.float64) + 1j * x[:, 1].astype(np.float64)) * np.exp(1j * phase)
        exp = np.stack([z.real, z.imag], axis=1)
        if "int" in dtype:
            exp = np.trunc(exp)
        assert np.all(np.abs(y.astype(np.float64) - exp) <= 1)
        assert np.array_equal(y, oracle.rotate(code, phase, x))


def test_signal_probe_rms_at_the_end_of_the_fir_topology(oracle, cuda_device):
    """filter/TestFIRFilter.cpp:44-78: waveform -> /comms/fir_filter -> /comms/signal_probe (RMS mode), checked
    through the probe's "valueChanged" signal and value() call; here every block's samples stay in HBM."""
    from pothoscomms_b200 import blocks
    dtype = "complex_float32"
    code = oracle.DTYPE_CODES[dtype]
    amplitude, rate, freq, n = 1000.0, 1e6, 30e3, 4096
    wave = amplitude * np.exp(2j * np.pi * freq / rate * np.arange(n))
    x = oracle.to_raw(wave, code)
    taps = sinc_complex_bandpass(101, (freq - 0.1 * rate) / rate, (freq + 0.1 * rate) / rate)
    fir = blocks.make("/comms/fir_filter", dtype, "COMPLEX")
    fir.call("setTaps", taps)
    fir.activate()
    y = fir.push_through(x)
    probe = blocks.make("/comms/signal_probe", dtype)
    assert all(probe.has_call(c) for c in "value setMode getMode setWindow getWindow setRate getRate probeValue".split())
    assert probe.has_signal("valueChanged") and probe.has_signal("valueTriggered")
    assert (probe.call("getMode"), probe.call("getWindow"), probe.call("getRate")) == ("VALUE", 1024, 0.0)   # SignalProbe.cpp:63-66
    probe.call("setMode", "RMS")
    probe.call("setWindow", 1024)
    assert probe.reserve == 1024                       # setWindow sets the input reserve (:99)
    probe.activate()
    assert probe.input_domain == "b200c_hbm"
    probe.feed(y[:3 * 1024 + 100])
    probe.run()
    value, count = probe.last_signal_value("valueChanged")
    assert count == 3 and probe.total_consumed == 3 * 1024      # one emission per full window; the rest waits for the reserve
    ref = oracle.probe(code, "RMS", y[2 * 1024: 3 * 1024]).real
    assert abs(value.real - ref) <= 1e-9 * ref and value.imag == 0.0
    assert probe.call("value") == value
    assert value.real > 0.1 * amplitude                 # POTHOS_TEST_TRUE(rms > 0.1*waveAmplitude), :78
    for mode in ("MEAN", "VALUE"):
        p2 = blocks.make("/blocks/stream_probe", dtype)
        p2.call("setMode", mode)
        p2.call("setWindow", 512)
        p2.activate()
        p2.feed(y[:512])
        p2.run()
        got, c = p2.last_signal_value()
        assert c == 1 and abs(got - oracle.probe(code, mode, y[:512])) <= 1e-6 * max(1.0, abs(got))


@pytest.mark.parametrize("dtype,M,L", [("complex_int16", 1, 1), ("complex_int16", 2, 3), ("int16", 3, 1), ("complex_float32", 1, 1)])
def test_history_window_across_the_ring_seam(oracle, cuda_device, dtype, M, L):
    """The device replacement of BufferManager::make("circular") (filter/FIRFilter.cpp:196-199): the K-1 history
    elements left behind by consume() (:304-307) plus the new data must stay ONE contiguous window when it runs
    across the end of the ring (base + bytes) -- the second VMM mapping.  The stream is several ring lengths long
    and fed in pieces that do not divide the ring, so many work() calls see a window that straddles the seam;
    int16 streams must still equal the oracle's one-shot output bit for bit."""
    from pothoscomms_b200 import blocks
    code = oracle.DTYPE_CODES[dtype]
    rng = np.random.default_rng(42)
    nc = 2 if code & 1 else 1
    tcx = bool(code & 1)
    taps = (rng.standard_normal(128) * 0.05 + (1j * rng.standard_normal(128) * 0.05 if tcx else 0))
    f = blocks.make("/comms/fir_filter", dtype, "COMPLEX" if tcx else "REAL", in_bytes=1 << 21, out_bytes=1 << 23)
    f.call("setTaps", taps)
    f.call("setDecimation", M)
    f.call("setInterpolation", L)
    f.activate()
    esz = nc * (2 if "int16" in dtype else 4)
    ring_elems = f.ring_bytes // esz
    n = 3 * ring_elems + 12345
    if "int" in dtype:
        x = rng.integers(-30000, 30000, size=(n, nc), dtype=np.int16)
    else:
        x = rng.standard_normal((n, nc)).astype(np.float32)
    piece = ring_elems * 5 // 13 + 7          # does not divide the ring: the write and read windows both wrap
    outs, pos = [], 0
    while pos < n:
        got = f.feed(x[pos: pos + piece])
        pos += got
        f.run()
        outs.append(f.collect())
        assert got > 0
    y = np.concatenate(outs)
    y_ref, cons, prod = oracle.fir(code, tcx, taps, M, L, x)
    assert f.seam_windows >= 2, "no work() call saw a window across base + bytes: the test did not exercise the seam"
    assert f.total_consumed == cons and y.shape[0] == prod
    if "int" in dtype:
        assert np.array_equal(y, y_ref)
    else:
        assert rel_rms(y, y_ref) < 1e-5


def test_stream_bench_runs_without_host_copies(oracle, cuda_device):
    """blocks/Harness.cpp streamBench (bench.py --workload headline_blocks): work() buffers of a fixed size between
    device neighbours; every round is consumed completely and the ring wraps."""
    from pothoscomms_b200 import blocks
    from pothoscomms_b200 import workloads as wl
    taps, tt = wl.config_taps("headline")
    chunk = 1 << 17
    f = blocks.make("/comms/fir_filter", "complex_float32", tt, in_bytes=4 * chunk * 8, out_bytes=2 * chunk * 8)
    f.call("setTaps", taps)
    f.activate()
    pat = wl.tone_noise_numpy(oracle.CF32, 1 << 16, seed=3)
    secs = f.stream_bench(pat, chunk, 50)
    assert secs > 0
    assert f.total_consumed >= 52 * chunk - 255 and f.work_calls >= 53
    assert f.seam_windows > 0


@pytest.mark.parametrize("dtype,M,L", [("complex_int16", 1, 1), ("complex_float32", 2, 3), ("int16", 1, 2)])
def test_host_neighbours_go_through_the_bridge_blocks(oracle, cuda_device, dtype, M, L):
    """The reference's test topology has HOST neighbours (feeder_source -> fir_filter -> collector_sink,
    filter/TestFIRFilter.cpp:49-51).  The device block refuses to hand its HBM buffers to a host-domain peer
    (PortDomainError, checked inside the chain runner); wired through /b200c/host_to_hbm and /b200c/hbm_to_host the
    same topology produces the oracle's output: one copy per buffer at each edge, HBM in between."""
    from pothoscomms_b200 import blocks
    code = oracle.DTYPE_CODES[dtype]
    tcx = bool(code & 1)
    rng = np.random.default_rng(17)
    nc = 2 if code & 1 else 1
    taps = rng.standard_normal(101) * 0.05 + (1j * rng.standard_normal(101) * 0.05 if tcx else 0)
    n = 50000
    x = rng.integers(-30000, 30000, size=(n, nc), dtype=np.int16) if "int" in dtype else rng.standard_normal((n, nc)).astype(np.float32)
    f = blocks.make("/comms/fir_filter", dtype, "COMPLEX" if tcx else "REAL")
    f.call("setTaps", taps)
    f.call("setDecimation", M)
    f.call("setInterpolation", L)
    y, bridge_calls = f.run_host_chain(x, chunk=7000, out_capacity=(n // M + 1) * L)
    y_ref, cons, prod = oracle.fir(code, tcx, taps, M, L, x)
    assert y.shape[0] == prod and bridge_calls >= 2 * (n // 7000)
    if "int" in dtype:
        assert np.array_equal(y, y_ref)
    else:
        assert rel_rms(y, y_ref) < 1e-5


def test_fft_block_between_host_neighbours(oracle, cuda_device):
    """fft/TestFFT.cpp:33-46 (feeder -> fft -> collector) wired through the bridge blocks."""
    from pothoscomms_b200 import blocks
    rng = np.random.default_rng(5)
    x = rng.integers(-20000, 20000, size=(16 * 1024, 2), dtype=np.int16)
    f = blocks.make("/comms/fft", "complex_int16", 1024, False)
    y, _ = f.run_host_chain(x, chunk=3000, out_capacity=16 * 1024)
    ref = oracle.ref_fft(oracle.CI16, 1024, False, x) if oracle.have_ref() else oracle.fft(oracle.CI16, 1024, False, x)
    assert y.shape[0] == 16 * 1024 and np.array_equal(y, ref)
