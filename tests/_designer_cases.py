"""The reference's FIR designer test matrix (filter/TestFIRDesigner.cpp), shared by the CPU and GPU tests."""
FILTER_TYPES = ["SINC", "MAXFLAT", "GAUSSIAN", "REMEZ", "ROOT_RAISED_COSINE", "RAISED_COSINE"]
BAND_TYPES = ["LOW_PASS", "HIGH_PASS", "BAND_PASS", "BAND_STOP", "COMPLEX_BAND_PASS", "COMPLEX_BAND_STOP"]


def reference_matrix():
    """filter/TestFIRDesigner.cpp:255-273 incl. its skips"""
    for ft in FILTER_TYPES:
        for bt in BAND_TYPES:
            stop, high = "STOP" in bt, "HIGH" in bt
            if ft == "MAXFLAT" and stop:
                continue
            if ft == "GAUSSIAN":
                continue
            if ft in ("RAISED_COSINE", "ROOT_RAISED_COSINE") and (stop or high):
                continue
            yield ft, bt


def mask_points(band, rate, lo, hi):
    """(pass?, frequency) pairs of filter/TestFIRDesigner.cpp:191-230"""
    P, S = True, False
    return {
        "LOW_PASS": [(S, -(lo + rate / 2) / 2), (P, 0.0), (S, (lo + rate / 2) / 2)],
        "HIGH_PASS": [(P, -(lo + rate / 2) / 2), (S, 0.0), (P, (lo + rate / 2) / 2)],
        "BAND_PASS": [(S, -(hi + rate / 2) / 2), (P, -(lo + hi) / 2), (S, 0.0), (P, (lo + hi) / 2), (S, (hi + rate / 2) / 2)],
        "BAND_STOP": [(P, -(hi + rate / 2) / 2), (S, -(lo + hi) / 2), (P, 0.0), (S, (lo + hi) / 2), (P, (hi + rate / 2) / 2)],
        "COMPLEX_BAND_PASS": [(S, (lo - rate / 2) / 2), (P, (lo + hi) / 2), (S, (hi + rate / 2) / 2)],
        "COMPLEX_BAND_STOP": [(P, (lo - rate / 2) / 2), (S, (lo + hi) / 2), (P, (hi + rate / 2) / 2)],
    }[band]


def configure(designer, ft, bt, rate=1e6, lo=1.5e5, hi=3.0e5, ntaps=101):
    """the call sequence of filter/TestFIRDesigner.cpp:159-166"""
    designer.call("setSampleRate", rate)
    designer.call("setFilterType", ft)
    designer.call("setBandType", bt)
    designer.call("setFrequencyLower", lo)
    designer.call("setFrequencyUpper", hi)
    designer.call("setBandwidthTrans", rate / 20)
    designer.call("setNumTaps", ntaps)
