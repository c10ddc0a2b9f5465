"""/comms/fir_designer and /comms/window_designer (host blocks).

The reference takes the tap maths from the external Spuce library and pins it only through a
pass/stop mask (filter/TestFIRDesigner.cpp:97-124, 237-275): the same matrix of filter and band
types is run here against the designer block's "tapsChanged" payload, with the response taken
by numpy instead of the FIR -> FFT chain (tests/test_blocks_gpu.py runs that chain on the GPU).
Windows and the Parks-McClellan routine are also compared with scipy's.
"""
import numpy as np
import pytest

from _designer_cases import configure, mask_points, reference_matrix


def power_bins(taps, fft_size=1024):
    """what the reference's impulse -> FIR -> FFT -> fftPowerSpectrum chain yields: |H|^2 in dB, [-fs/2, fs/2)"""
    H = np.fft.fft(np.asarray(taps, dtype=complex), fft_size)
    return 20 * np.log10(np.maximum(np.abs(np.fft.fftshift(H)), 1e-15))


@pytest.mark.parametrize("ft,bt", list(reference_matrix()))
def test_fir_designer_mask(ft, bt):
    from pothoscomms_b200 import blocks
    d = blocks.make("/comms/fir_designer")
    configure(d, ft, bt)
    assert d.last_signal()[1] == 0                       # nothing is emitted before activation (:389)
    d.activate()
    taps, count = d.last_signal()
    assert count == 1 and taps.shape[0] == 101
    assert np.iscomplexobj(taps) == ("COMPLEX" in bt)    # complex bands emit vector<complex<double>> (:470-476)
    power = power_bins(taps)
    for is_pass, freq in mask_points(bt, 1e6, 1.5e5, 3.0e5):
        level = power[int(1024 * ((freq + 5e5) / 1e6))]
        assert (level > -30.0) if is_pass else (level < -80.0), (ft, bt, freq, level)


def test_fir_designer_call_surface_and_defaults():
    from pothoscomms_b200 import blocks
    for path in ("/comms/fir_designer", "/blocks/fir_designer", "/comms/window_designer"):   # FIRDesigner.cpp:479-483, WindowDesigner.cpp:134
        assert blocks.registry_has(path)
    d = blocks.make("/blocks/fir_designer")
    for name in ("setBandType bandType setFilterType filterType setWindowType windowType setWindowArgs windowArgs setSampleRate "
                 "sampleRate setFrequencies setFrequencyLower frequencyLower setFrequencyUpper frequencyUpper setBandwidthTrans "
                 "bandwidthTrans setNumTaps numTaps setAlpha alpha setStopDB stopDB setPassDB passDB setGain gain").split():
        assert d.has_call(name), name                    # filter/FIRDesigner.cpp:143-168
    assert d.has_signal("tapsChanged")
    # constructor defaults, filter/FIRDesigner.cpp:128-141
    assert (d.call("filterType"), d.call("bandType"), d.call("windowType")) == ("GAUSSIAN", "LOW_PASS", "hann")
    assert (d.call("gain"), d.call("sampleRate"), d.call("frequencyLower"), d.call("frequencyUpper")) == (1.0, 1.0, 0.1, 0.2)
    assert (d.call("bandwidthTrans"), d.call("alpha"), d.call("stopDB"), d.call("passDB"), d.call("numTaps")) == (0.1, 0.5, 60.0, 0.1, 51)
    d.call("setFrequencies", [0.05, 0.3])
    assert (d.call("frequencyLower"), d.call("frequencyUpper")) == (0.05, 0.3)
    d.call("setWindowArgs", [7.5])
    assert list(d.call("windowArgs")) == [7.5]
    # a band name passed as filter type is the legacy usage: SINC of that band (:195-214)
    d.call("setFilterType", "HIGH_PASS")
    assert (d.call("filterType"), d.call("bandType")) == ("SINC", "HIGH_PASS")


def test_fir_designer_reemits_on_every_setter_and_applies_gain_and_window():
    from pothoscomms_b200 import blocks
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "SINC")
    d.call("setWindowType", "rectangular")
    d.activate()
    base, n0 = d.last_signal()
    d.call("setGain", 2.5)
    scaled, n1 = d.last_signal()
    assert n1 == n0 + 1 and np.allclose(scaled, 2.5 * base)
    d.call("setWindowType", "hamming")
    win, n2 = d.last_signal()
    assert n2 == n1 + 1 and np.allclose(win, scaled * np.hamming(51))
    d.call("setNumTaps", 33)
    assert d.last_signal()[0].shape[0] == 33


@pytest.mark.parametrize("setter,value,message", [
    ("setNumTaps", 0, "num taps must be positive"), ("setSampleRate", 0.0, "sample rate must be positive"),
    ("setFrequencyLower", 0.0, "lower frequency must be positive"), ("setFrequencyLower", 0.6, "lower frequency above Nyquist range")])
def test_fir_designer_parameter_errors(setter, value, message):
    """filter/FIRDesigner.cpp:396-402: thrown from the offending setter once the block is active"""
    from pothoscomms_b200 import blocks
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "SINC")
    d.activate()
    with pytest.raises(blocks.PothosException, match=message):
        d.call(setter, value)


def test_fir_designer_band_errors():
    from pothoscomms_b200 import blocks
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "SINC")
    d.call("setNumTaps", 50)
    d.activate()
    with pytest.raises(blocks.PothosException, match="odd number of taps"):          # :410
        d.call("setBandType", "BAND_PASS")
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "SINC")
    d.call("setFrequencies", [0.3, 0.2])
    d.activate()
    with pytest.raises(blocks.PothosException, match="upper frequency <= lower frequency"):   # :415
        d.call("setBandType", "BAND_STOP")
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "MAXFLAT")
    d.activate()
    with pytest.raises(blocks.PothosException, match="MAXFLAT"):                     # :419-422
        d.call("setBandType", "BAND_STOP")
    d = blocks.make("/comms/fir_designer")
    d.activate()
    with pytest.raises(blocks.InvalidArgumentException, match="Problem with creating taps"):   # :453-457
        d.call("setFilterType", "NO_SUCH_TYPE")
    d = blocks.make("/comms/fir_designer")
    d.call("setBandwidthTrans", 0.0)
    d.activate()
    with pytest.raises(blocks.PothosException, match="Transition Bandwidth"):        # :424
        d.call("setFilterType", "REMEZ")


def test_window_designer_against_scipy():
    from scipy.signal import windows
    from pothoscomms_b200 import blocks
    w = blocks.make("/comms/window_designer")
    assert all(w.has_call(c) for c in "setWindowType windowType setWindowArgs windowArgs setNumTaps numTaps".split())
    assert (w.call("windowType"), w.call("numTaps")) == ("hann", 51)                # window/WindowDesigner.cpp:61-62
    w.activate()
    for n in (51, 64, 7):
        w.call("setNumTaps", n)
        cases = {"rectangular": (None, np.ones(n)), "hamming": (None, windows.hamming(n)), "flattop": (None, windows.flattop(n)),
                 "kaiser": (6.5, windows.kaiser(n, 6.5)), "chebyshev": (70.0, windows.chebwin(n, 70.0)),
                 # open-ended variants: no tap is zeroed (the n+2 point textbook window without its end points)
                 "hann": (None, windows.hann(n + 2)[1:-1]), "blackman": (None, windows.blackman(n + 2)[1:-1]),
                 "bartlett": (None, windows.bartlett(n + 2)[1:-1])}
        for name, (arg, ref) in cases.items():
            w.call("setWindowArgs", [] if arg is None else [arg])
            w.call("setWindowType", name)
            got, _ = w.last_signal()
            assert got.shape[0] == n and np.allclose(got, ref, atol=1e-12), (name, n)
    with pytest.raises(blocks.PothosException, match="num taps must be positive"):   # :125
        w.call("setNumTaps", 0)
    w.call("setNumTaps", 5)                             # like the reference, a rejected value stays stored until replaced
    with pytest.raises(blocks.InvalidArgumentException):
        w.call("setWindowType", "no_such_window")


@pytest.mark.parametrize("n,fp,fs,wt", [(101, 0.125, 0.175, 11.5), (64, 0.1, 0.2, 1.0), (33, 0.2, 0.3, 10.0), (255, 0.05, 0.07, 3.0)])
def test_remez_designs_are_equiripple_like_scipy(n, fp, fs, wt):
    """REMEZ/LOW_PASS with a rectangular window is the bare Parks-McClellan design: compare with scipy.signal.remez"""
    from scipy.signal import remez
    from pothoscomms_b200 import blocks
    d = blocks.make("/comms/fir_designer")
    d.call("setFilterType", "REMEZ")
    d.call("setWindowType", "rectangular")
    d.call("setNumTaps", n)
    d.call("setFrequencyLower", (fp + fs) / 2)          # nominal edge: the transition band is centred on it
    d.call("setBandwidthTrans", fs - fp)
    delta_s = 1e-3
    pass_db = 20 * np.log10((1 + wt * delta_s) / (1 - wt * delta_s))   # makes remez_estimate_weight() == wt
    d.call("setPassDB", pass_db)
    d.call("setStopDB", 60.0)
    d.activate()
    taps, _ = d.last_signal()
    ref = remez(n, [0, fp, fs, 0.5], [1, 0], weight=[1, wt], fs=1.0, maxiter=100)
    assert np.max(np.abs(taps - ref)) < 1e-4 * np.max(np.abs(ref))
