"""The algebra fir_os32x_kernel implements, restated in numpy (double) and checked
against the CPU oracle -- no GPU involved.  For interpolation L = 3 / decimation M = 2 the reference's nest
(filter/FIRFilter.cpp:286-302) equals, per block of 1024 inputs starting at buffer element P,
    X  = DFT_1024(in[P : P + 1024])
    W[k] = X[k mod 1024] H'[k] + X[(k + 512) mod 1024] H'[k + 1536],  k < 1536
    w  = 1536 * IDFT_1536(W),        y[m' + 3 (P - (K-1)) / 2] = w[m']  for m' >= m0
with H'[k] = DFT_3072(h)[k] exp(2 pi i k (M - 1) / 3072) / 3072, m0 = ceil((ntaps - 2) / 2) rounded up to a multiple of 3,
first block at P0 = (K - 1) - 2 m0 / 3 and a hop of (1536 - m0) outputs = (1536 - m0) 2 / 3 inputs.  The kernel's plan
(fir_os_configure) uses exactly these m0 / P0 formulas."""
import numpy as np
import pytest

L, M = 3, 2


def plan(ntaps):
    K = -(-ntaps // L)
    m0 = (-(-(ntaps - 2) // 2) + 2) // 3 * 3
    return K, m0, (K - 1) - 2 * (m0 // 3)


def spectral_resample(taps, x, n_out):
    ntaps = len(taps)
    K, m0, p0 = plan(ntaps)
    hop_out, hop_in = 1536 - m0, (1536 - m0) // 3 * 2
    k = np.arange(3072)
    H = np.exp(-2j * np.pi * np.outer(k, np.arange(ntaps)) / 3072) @ taps
    Hp = H * np.exp(2j * np.pi * k * (M - 1) / 3072) / 3072
    y = np.zeros(n_out, dtype=np.complex128)
    kap = np.arange(1536)
    for b in range(-(-n_out // hop_out)):
        P = p0 + b * hop_in
        idx = P + np.arange(1024)
        ok = (idx >= 0) & (idx < len(x))
        X = np.fft.fft(np.where(ok, x[np.clip(idx, 0, len(x) - 1)], 0))
        w = np.fft.ifft(X[kap % 1024] * Hp[kap] + X[(kap + 512) % 1024] * Hp[kap + 1536]) * 1536
        m = b * hop_out + np.arange(m0, 1536) - m0
        keep = m < n_out
        y[m[keep]] = w[m0:][keep]
    return y


@pytest.mark.parametrize("ntaps", [2, 7, 48, 100, 255, 301, 1200])
def test_replicate_multiply_fold_equals_the_reference_nest(oracle, ntaps):
    rng = np.random.default_rng(ntaps)
    taps = rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)
    n = 6000
    xr = oracle.to_raw(rng.standard_normal(n) + 1j * rng.standard_normal(n), oracle.CF64)
    y_ref, cons, prod = oracle.fir(oracle.CF64, True, taps, M, L, xr)
    y = spectral_resample(taps, xr[:, 0] + 1j * xr[:, 1], prod)
    ref = y_ref[:, 0] + 1j * y_ref[:, 1]
    assert np.max(np.abs(y - ref)) <= 1e-11 * np.max(np.abs(ref))
    K, m0, p0 = plan(ntaps)
    assert m0 % 3 == 0 and (p0 - (K - 1)) % 2 == 0 and 2 * m0 + 1 >= ntaps - 1     # alias free from m0, blocks on the output grid
