"""Synthetic workloads of BASELINE.json's configs: host-generated taps and tone+noise input.

Taps are plain numpy (the reference keeps tap design on the host: filter/FIRDesigner.cpp);
identical bytes are handed to the CPU oracle and to the GPU path.  The input follows the
reference's own test signal (filter/TestFIRFilter.cpp:72-74: 30 kHz complex sine at 1 MHz;
waveform/WaveformSource.cpp:221 for the sine, waveform/NoiseSource.cpp:17-18 for
independent re/im noise): x[n] = A*exp(j*2*pi*f0*n/fs) + sigma*(g_re + j*g_im).
"""
from __future__ import annotations

import numpy as np

FS = 1e6
F0 = 30e3


def hann(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(n) + 0.5) / n)


def sinc_lowpass(ntaps: int, cutoff: float) -> np.ndarray:
    """Hann-windowed sinc low-pass, cutoff as a fraction of the sample rate, unity DC gain."""
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = 2 * cutoff * np.sinc(2 * cutoff * t) * hann(ntaps)
    return h / h.sum()


def complex_bandpass(ntaps: int, center: float, halfwidth: float) -> np.ndarray:
    """Low-pass prototype shifted to `center` (fractions of fs): complex band-pass taps."""
    t = np.arange(ntaps) - (ntaps - 1) / 2.0
    return sinc_lowpass(ntaps, halfwidth) * np.exp(2j * np.pi * center * t)


def root_raised_cosine(ntaps: int, sps: float, alpha: float = 0.5) -> np.ndarray:
    """RRC pulse, `sps` samples per symbol, roll-off alpha (designer default 0.5,
    filter/FIRDesigner.cpp:108,156), unit energy."""
    t = (np.arange(ntaps) - (ntaps - 1) / 2.0) / sps
    h = np.empty(ntaps)
    for i, ti in enumerate(t):
        if abs(ti) < 1e-12:
            h[i] = 1.0 - alpha + 4 * alpha / np.pi
        elif abs(abs(4 * alpha * ti) - 1.0) < 1e-9:
            h[i] = (alpha / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))) / (np.pi * ti * (1 - (4 * alpha * ti) ** 2))
    return h / np.sqrt(np.sum(h * h))


def config_taps(name: str):
    """(taps, taps_type) for the named BASELINE.json config."""
    if name == "c1_real":
        return sinc_lowpass(64, 0.1), "REAL"
    if name == "c1":
        return sinc_lowpass(64, 0.1).astype(np.complex128), "COMPLEX"
    if name == "headline":   # cf32, 256 taps (the metric string)
        return complex_bandpass(256, F0 / FS, 0.05), "COMPLEX"
    if name == "c2":         # complex int16, 128 complex taps, sum|h| <= 1
        h = complex_bandpass(128, F0 / FS, 0.1)
        return h / np.abs(h).sum(), "COMPLEX"
    if name == "c3":         # L=3, M=2, 255-tap RRC, real taps
        return root_raised_cosine(255, 3.0, 0.5), "REAL"
    if name == "c3_i16":     # the same resampler scaled for fixed point: peak tap 0.33 (Q16 taps of two byte digits)
        return 0.5 * root_raised_cosine(255, 3.0, 0.5), "REAL"
    if name == "c5":
        return complex_bandpass(1024, F0 / FS, 0.02), "COMPLEX"
    # short-tap streams (north_star: "the short-tap path stays on FFMA and is reported as an
    # HBM-bound stream"): not BASELINE configs, bench-only
    if name == "short":      # 16 real taps
        return sinc_lowpass(16, 0.1), "REAL"
    if name == "short_cx":   # 16 complex taps
        return complex_bandpass(16, F0 / FS, 0.1), "COMPLEX"
    if name == "resamp_short":   # L=3, M=2, 48-tap RRC (16 taps per phase)
        return root_raised_cosine(48, 3.0, 0.5), "REAL"
    if name == "real64":     # real data, 64 real taps
        return sinc_lowpass(64, 0.1), "REAL"
    if name.startswith("sweep"):   # bench.py --ntaps N
        return complex_bandpass(int(name[5:]), F0 / FS, 0.1), "COMPLEX"
    raise KeyError(name)


def bank_taps(chan: int, nchan: int, ntaps: int = 1024) -> np.ndarray:
    """Channel `chan` of an `nchan`-channel channeliser: the low-pass prototype (half-width
    0.5/nchan of fs) shifted to the channel centre (chan/nchan - 0.5) * fs: complex taps."""
    return complex_bandpass(ntaps, chan / nchan - 0.5, 0.5 / nchan)


def tone_noise_numpy(dtype_code: int, n: int, seed: int, start: int = 0) -> np.ndarray:
    """Raw [n, ncomp] tone+noise for small CPU-side cases (numpy Philox, seeded)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    cls, cx = dtype_code >> 1, dtype_code & 1
    is_float = cls < 2
    amp, sigma = (1.0, 0.1) if is_float else {2: (100.0, 10.0), 3: (1000.0, 100.0)}.get(cls, (1000.0, 100.0))
    idx = np.arange(start, start + n, dtype=np.float64)
    ph = 2 * np.pi * (F0 / FS) * idx
    re = amp * np.cos(ph) + sigma * rng.standard_normal(n)
    im = amp * np.sin(ph) + sigma * rng.standard_normal(n)
    sc = {0: np.float32, 1: np.float64, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64}[cls]
    cols = [re, im] if cx else [re]
    out = np.stack(cols, axis=1)
    if not is_float:
        info = np.iinfo(sc)
        out = np.clip(np.rint(out), info.min, info.max)
    return np.ascontiguousarray(out.astype(sc))


def tone_noise_torch(dtype_code: int, n: int, seed: int, device, chunk: int = 1 << 24):
    """Raw [n, ncomp] tone+noise generated in HBM (torch's seeded Philox), chunked."""
    import torch
    cls, cx = dtype_code >> 1, dtype_code & 1
    is_float = cls < 2
    amp, sigma = (1.0, 0.1) if is_float else {2: (100.0, 10.0), 3: (1000.0, 100.0)}.get(cls, (1000.0, 100.0))
    tdt = {0: torch.float32, 1: torch.float64, 2: torch.int8, 3: torch.int16, 4: torch.int32, 5: torch.int64}[cls]
    nc = 2 if cx else 1
    out = torch.empty((n, nc), dtype=tdt, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    # phase is kept small by reducing the sample index mod the tone period (100/3 samples -> 100 samples = 3 periods)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        idx = (torch.arange(s, s + m, device=device, dtype=torch.int64) % 100).to(torch.float32)
        ph = idx * (2 * np.pi * F0 / FS)
        noise = torch.randn((m, nc), generator=gen, device=device, dtype=torch.float32) * sigma
        sig = torch.stack([torch.cos(ph), torch.sin(ph)][:nc], dim=1) * amp + noise
        if is_float:
            out[s:s + m] = sig.to(tdt)
        else:
            info = torch.iinfo(tdt)
            out[s:s + m] = torch.clamp(torch.round(sig), info.min, info.max).to(tdt)
    return out
