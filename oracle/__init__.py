"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the /comms/fir_filter + /comms/fft hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``pothoscomms_b200``) never
does; it fails loudly if its CUDA library is missing instead of falling back to this.

* ``fir(...)``      : C restatement of filter/FIRFilter.cpp:278-302,327-354 (``liboracle.so``), pinned
                      bit-for-bit against ``ref_fir`` (tests/test_oracle_fir.py).
* ``ref_fir(...)``, ``ref_fir_stream(...)`` : the REFERENCE's own filter/FIRFilter.cpp, compiled unmodified
                      from /root/reference into ``oracle/_ref/libfirref.so`` against the repo's Pothos API
                      subset and a RECALLED Pothos/Util/QFormat.hpp (the one external header; its rounding
                      is the only thing still unpinned, see oracle/ref_include/Pothos/Util/QFormat.hpp).
* ``fft(...)``      : C restatement of fft/kissfft.hh + fft/kiss_fft.c (``liboracle.so``), pinned
                      bit-for-bit against ``ref_fft``.
* ``ref_fft(...)``  : the REFERENCE's own kiss_fft sources compiled from /root/reference into
                      ``oracle/_ref/libkissref.so`` (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# dtype codes shared with include/b200comms.h
F32, CF32, F64, CF64, I8, CI8, I16, CI16, I32, CI32, I64, CI64 = range(12)
DTYPE_CODES = {
    "float32": F32, "complex_float32": CF32, "float64": F64, "complex_float64": CF64,
    "int8": I8, "complex_int8": CI8, "int16": I16, "complex_int16": CI16,
    "int32": I32, "complex_int32": CI32, "int64": I64, "complex_int64": CI64,
}
_SCALAR_NP = {0: np.float32, 1: np.float64, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64}


def scalar_np(dtype_code: int):
    return _SCALAR_NP[dtype_code >> 1]


def is_complex(dtype_code: int) -> bool:
    return bool(dtype_code & 1)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    lib = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fir_oracle.c", "fft_oracle.c", "math_oracle.c", "source_oracle.cpp", "qformat.h", "Makefile")]
    stale = force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs)
    have_checkout = os.path.exists("/root/reference/fft/kiss_fft.c")
    refs = [os.path.join(_HERE, "_ref", n) for n in ("libkissref.so", "libfirref.so")]
    ref_srcs = [os.path.join(_HERE, f) for f in ("ref_fir_wrap.cpp", "ref_wrap.cpp", "ref_include/Pothos/Util/QFormat.hpp")]
    ref_stale = have_checkout and any(not os.path.exists(r) or any(os.path.getmtime(s) > os.path.getmtime(r) for s in ref_srcs)
                                      for r in refs)
    if stale or ref_stale:
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True, capture_output=True)


_lib = None
_ref = None
_firref = None
DTYPE_NAMES = {v: k for k, v in DTYPE_CODES.items()}


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        sz, vp, i = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
        _lib.oracle_fir_run.argtypes = [i, i, vp, sz, sz, sz, vp, sz, vp, sz, i, vp, vp]
        _lib.oracle_fir_run_mt.argtypes = [i, i, i, vp, sz, sz, sz, vp, sz, vp, sz, vp, vp]
        _lib.oracle_fir_K.argtypes = [sz, sz]
        _lib.oracle_fir_K.restype = sz
        _lib.oracle_fir_phase_taps.argtypes = [i, i, vp, sz, sz, vp, vp]
        _lib.oracle_fft.argtypes = [i, sz, i, vp, vp, sz]
        _lib.oracle_fft_plan.argtypes = [i, i, vp, vp]
        _lib.oracle_scale.argtypes = [i, ctypes.c_double, vp, vp, sz]
        _lib.oracle_rotate.argtypes = [i, ctypes.c_double, vp, vp, sz]
        _lib.oracle_probe.argtypes = [i, i, vp, sz, vp]
        d, u64 = ctypes.c_double, ctypes.c_uint64
        _lib.oracle_waveform_table.argtypes = [i, ctypes.c_char_p, d, d, d, d, d, d, d, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(u64)]
        _lib.oracle_table_walk.argtypes = [i, vp, sz, u64, u64, vp, sz]
        _lib.oracle_table_walk.restype = None
        _lib.oracle_noise_stream.argtypes = [i, ctypes.c_char_p, d, d, d, d, d, d, ctypes.c_uint32, vp, vp, sz, vp, vp]
    return _lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libkissref.so"))


def ref():
    """The reference's own FFT sources, compiled (oracle/_ref)."""
    global _ref
    if _ref is None:
        build()
        _ref = ctypes.CDLL(os.path.join(_HERE, "_ref", "libkissref.so"))
        sz, vp, i = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
        _ref.ref_fft.argtypes = [i, sz, i, vp, vp, sz]
        _ref.ref_fft_mt.argtypes = [i, i, sz, i, vp, vp, sz]
    return _ref


def have_ref_fir() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libfirref.so"))


def firref():
    """The reference's own filter/FIRFilter.cpp, compiled (oracle/_ref/libfirref.so)."""
    global _firref
    if _firref is None:
        build()
        _firref = ctypes.CDLL(os.path.join(_HERE, "_ref", "libfirref.so"))
        sz, vp, i, cp = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
        _firref.firref_last_error.restype = cp
        _firref.firref_stream.argtypes = [cp, cp, vp, sz, sz, sz, vp, sz, vp, sz, sz, sz, i, vp, vp, vp]
        _firref.firref_work.argtypes = [cp, cp, vp, sz, sz, sz, vp, sz, vp, sz, vp, vp]
        _firref.firref_work_mt.argtypes = [i, cp, cp, vp, sz, sz, sz, vp, sz, vp, sz, vp, vp]
        _firref.firref_input_require.argtypes = [cp, cp, vp, sz, sz, sz]
        _firref.firref_input_require.restype = ctypes.c_long
    return _firref


class ReferenceError_(ValueError):
    """An exception thrown by the reference block (factory or setter), text preserved."""


def _ref_fir_args(dtype_code, taps_complex, taps, x_raw):
    t = _as_taps(taps, taps_complex)
    ntaps = t.size // (2 if taps_complex else 1)
    ncomp = 2 if is_complex(dtype_code) else 1
    x_raw = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code)).reshape(-1, ncomp)
    return t, ntaps, ncomp, x_raw, DTYPE_NAMES[dtype_code].encode(), (b"COMPLEX" if taps_complex else b"REAL")


def ref_fir(dtype_code: int, taps_complex: bool, taps, decim: int, interp: int, x_raw: np.ndarray,
            out_capacity: int | None = None, threads: int = 1):
    """ONE work() call of the reference block on the window ``x_raw`` (K-1 history then new data):
    same contract as ``fir(..., zero_tail=False)``.  threads > 1: that many block instances, each on
    its own equal share of x_raw ([threads, seg, ncomp]) -- the CPU baseline driver."""
    t, ntaps, ncomp, x_raw, dn, tn = _ref_fir_args(dtype_code, taps_complex, taps, x_raw)
    cons, prod = ctypes.c_size_t(0), ctypes.c_size_t(0)
    if threads > 1:
        seg = x_raw.shape[0] // threads
        cap = (seg // decim + 1) * interp
        out = np.zeros((threads, cap, ncomp), dtype=x_raw.dtype)
        rc = firref().firref_work_mt(threads, dn, tn, t.ctypes.data, ntaps, decim, interp, x_raw.ctypes.data, seg,
                                     out.ctypes.data, cap, ctypes.addressof(cons), ctypes.addressof(prod))
        if rc:
            raise ReferenceError_(firref().firref_last_error().decode())
        return out, cons.value, prod.value
    n_in = x_raw.shape[0]
    if out_capacity is None:
        out_capacity = (n_in // max(decim, 1) + 1) * interp
    out = np.zeros((max(out_capacity, 1), ncomp), dtype=x_raw.dtype)
    rc = firref().firref_work(dn, tn, t.ctypes.data, ntaps, decim, interp, x_raw.ctypes.data, n_in, out.ctypes.data,
                              out_capacity, ctypes.addressof(cons), ctypes.addressof(prod))
    if rc:
        raise ReferenceError_(firref().firref_last_error().decode())
    return out[: prod.value], cons.value, prod.value


def ref_fir_stream(dtype_code: int, taps_complex: bool, taps, decim: int, interp: int, x_raw: np.ndarray,
                   in_chunk: int = 0, out_chunk: int = 0, frame_end: bool = False, out_capacity: int | None = None):
    """Stream ``x_raw`` through one reference block instance the way feeder -> fir_filter -> collector does
    (filter/TestFIRFilter.cpp:19-53): ``in_chunk`` elements arrive per round, ``out_chunk`` output elements are
    offered per work(); ``frame_end`` labels the last element as the end of a burst (flush path :263-272).
    Returns (out_raw, consumed, produced, work_calls)."""
    t, ntaps, ncomp, x_raw, dn, tn = _ref_fir_args(dtype_code, taps_complex, taps, x_raw)
    n_in = x_raw.shape[0]
    if out_capacity is None:
        out_capacity = (n_in // decim + 2) * interp
    out = np.zeros((max(out_capacity, 1), ncomp), dtype=x_raw.dtype)
    cons, prod, calls = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = firref().firref_stream(dn, tn, t.ctypes.data, ntaps, decim, interp, x_raw.ctypes.data, n_in, out.ctypes.data,
                                out_capacity, in_chunk, out_chunk, int(frame_end), ctypes.addressof(cons),
                                ctypes.addressof(prod), ctypes.addressof(calls))
    if rc:
        raise ReferenceError_(firref().firref_last_error().decode())
    return out[: prod.value], cons.value, prod.value, calls.value


def ref_fir_input_require(dtype_code: int, taps_complex: bool, taps, decim: int, interp: int) -> int:
    """_inputRequire = M + K - 1 of the reference block (filter/FIRFilter.cpp:353), read back through the
    reserve it sets when starved (:248-252)."""
    t, ntaps, _, _, dn, tn = _ref_fir_args(dtype_code, taps_complex, taps, np.zeros((1, 2 if is_complex(dtype_code) else 1)))
    r = firref().firref_input_require(dn, tn, t.ctypes.data, ntaps, decim, interp)
    if r < 0:
        raise ReferenceError_(firref().firref_last_error().decode())
    return int(r)


def _as_taps(taps, taps_complex: bool) -> np.ndarray:
    if taps_complex:
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.complex128))
        return t.view(np.float64)
    return np.ascontiguousarray(np.asarray(taps, dtype=np.float64))


def to_raw(x: np.ndarray, dtype_code: int) -> np.ndarray:
    """Interleaved scalar view [n, ncomp] of a numpy array for ``dtype_code``."""
    sc = scalar_np(dtype_code)
    if is_complex(dtype_code):
        if np.iscomplexobj(x):
            if sc in (np.float32, np.float64):
                return np.ascontiguousarray(x.astype(np.complex64 if sc == np.float32 else np.complex128)).view(sc).reshape(-1, 2)
            out = np.empty((x.size, 2), dtype=sc)
            out[:, 0] = x.real
            out[:, 1] = x.imag
            return out
        x = np.ascontiguousarray(x, dtype=sc)
        return x.reshape(-1, 2)
    return np.ascontiguousarray(x, dtype=sc).reshape(-1, 1)


def fir_K(ntaps: int, interp: int) -> int:
    return int(lib().oracle_fir_K(ntaps, interp))


def fir(dtype_code: int, taps_complex: bool, taps, decim: int, interp: int, x_raw: np.ndarray,
        out_capacity: int | None = None, zero_tail: bool = False, threads: int = 1):
    """One work() call.  ``x_raw`` is [n, ncomp] scalars: K-1 history elements then new data.
    Returns (out_raw [produced, ncomp], consumed, produced)."""
    t = _as_taps(taps, taps_complex)
    ntaps = t.size // (2 if taps_complex else 1)
    ncomp = 2 if is_complex(dtype_code) else 1
    x_raw = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code)).reshape(-1, ncomp)
    n_in = x_raw.shape[0]
    K = fir_K(ntaps, interp)
    if out_capacity is None:
        out_capacity = ((n_in + (K - 1 if zero_tail else 0)) // decim + 1) * interp
    out = np.zeros((max(out_capacity, 1), ncomp), dtype=x_raw.dtype)
    cons, prod = ctypes.c_size_t(0), ctypes.c_size_t(0)
    if threads > 1 and not zero_tail:
        rc = lib().oracle_fir_run_mt(threads, dtype_code, int(taps_complex), t.ctypes.data, ntaps, decim, interp,
                                     x_raw.ctypes.data, n_in, out.ctypes.data, out_capacity,
                                     ctypes.addressof(cons), ctypes.addressof(prod))
    else:
        rc = lib().oracle_fir_run(dtype_code, int(taps_complex), t.ctypes.data, ntaps, decim, interp,
                                  x_raw.ctypes.data, n_in, out.ctypes.data, out_capacity, int(zero_tail),
                                  ctypes.addressof(cons), ctypes.addressof(prod))
    if rc != 0:
        raise ValueError("oracle_fir_run: invalid arguments")
    return out[: prod.value], cons.value, prod.value


def fir_phase_taps(dtype_code: int, taps_complex: bool, taps, interp: int):
    t = _as_taps(taps, taps_complex)
    tc = 2 if taps_complex else 1
    ntaps = t.size // tc
    K = fir_K(ntaps, interp)
    out = np.zeros((interp, K, tc), dtype=np.float64)
    nt = np.zeros(interp, dtype=np.uintp)
    rc = lib().oracle_fir_phase_taps(dtype_code, int(taps_complex), t.ctypes.data, ntaps, interp,
                                     out.ctypes.data, nt.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_fir_phase_taps: invalid arguments")
    return out, nt


def _fft_call(fn, dtype_code, nbins, inverse, x_raw, *pre):
    ncomp = 2
    x_raw = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code)).reshape(-1, ncomp)
    batch = x_raw.shape[0] // nbins
    out = np.zeros_like(x_raw[: batch * nbins])
    rc = fn(*pre, dtype_code, nbins, int(inverse), x_raw.ctypes.data, out.ctypes.data, batch)
    if rc != 0:
        raise ValueError("fft oracle: unsupported dtype")
    return out


def fft(dtype_code: int, nbins: int, inverse: bool, x_raw: np.ndarray) -> np.ndarray:
    """floor(len/nbins) transforms with the restated kiss_fft."""
    return _fft_call(lib().oracle_fft, dtype_code, nbins, inverse, x_raw)


def ref_fft(dtype_code: int, nbins: int, inverse: bool, x_raw: np.ndarray, threads: int = 1) -> np.ndarray:
    """floor(len/nbins) transforms with the reference's own compiled kiss_fft."""
    if threads > 1:
        return _fft_call(ref().ref_fft_mt, dtype_code, nbins, inverse, x_raw, threads)
    return _fft_call(ref().ref_fft, dtype_code, nbins, inverse, x_raw)


def fft_plan(nbins: int, fixed: bool):
    radix = (ctypes.c_int * 64)()
    rem = (ctypes.c_int * 64)()
    n = lib().oracle_fft_plan(nbins, int(fixed), ctypes.addressof(radix), ctypes.addressof(rem))
    return list(radix[:n]), list(rem[:n])


def scale(dtype_code: int, factor: float, x_raw: np.ndarray) -> np.ndarray:
    """/comms/scale (math/Scale.cpp:15-23): out = fromQ(floatToQ(factor) * Q(in)), any shape of raw scalars."""
    x = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code))
    out = np.empty_like(x)
    rc = lib().oracle_scale(dtype_code, ctypes.c_double(factor), x.ctypes.data, out.ctypes.data, x.size)
    assert rc == 0
    return out


def rotate(dtype_code: int, phase: float, x_raw: np.ndarray) -> np.ndarray:
    """/comms/rotate (math/Rotate.cpp:15-23): out = fromQ(floatToQ(polar(1, phase)) * Q(in)), raw [n, 2]."""
    x = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code)).reshape(-1, 2)
    out = np.empty_like(x)
    rc = lib().oracle_rotate(dtype_code, ctypes.c_double(phase), x.ctypes.data, out.ctypes.data, x.shape[0])
    assert rc == 0
    return out


PROBE_MODES = {"VALUE": 0, "RMS": 1, "MEAN": 2}


def probe(dtype_code: int, mode: str, x_raw: np.ndarray) -> complex:
    """/comms/signal_probe (utility/SignalProbe.cpp:140-160) over the whole of x_raw ([n, ncomp])."""
    nc = 2 if is_complex(dtype_code) else 1
    x = np.ascontiguousarray(x_raw, dtype=scalar_np(dtype_code)).reshape(-1, nc)
    v = (ctypes.c_double * 2)()
    rc = lib().oracle_probe(dtype_code, PROBE_MODES[mode], x.ctypes.data, x.shape[0], ctypes.addressof(v))
    assert rc == 0
    return complex(v[0], v[1])


# ------------------------------------------------- stream sources ---
def waveform_table(dtype_code: int, wave: str, freq: float, rate: float, res: float = 0.0, ampl: complex = 1.0,
                   offset: complex = 0.0):
    """WaveformSource::updateTable() (waveform/WaveformSource.cpp:184-260): (raw table [entries, ncomp], step).
    Raises ValueError where the reference throws InvalidArgumentException."""
    ampl, offset = complex(ampl), complex(offset)
    nc = 2 if is_complex(dtype_code) else 1
    table = np.zeros((1 << 20, nc), dtype=scalar_np(dtype_code))
    entries, step = ctypes.c_size_t(0), ctypes.c_uint64(0)
    rc = lib().oracle_waveform_table(dtype_code, wave.encode(), freq, rate, res, ampl.real, ampl.imag, offset.real, offset.imag,
                                     table.ctypes.data, table.shape[0], ctypes.byref(entries), ctypes.byref(step))
    if rc:
        raise ValueError(f"WaveformSource::updateTable(): invalid setting ({wave}, freq={freq}, rate={rate})")
    return table[: entries.value].copy(), step.value


def table_walk(dtype_code: int, table_raw: np.ndarray, index: int, step: int, n: int) -> np.ndarray:
    """WaveformSource::work() (waveform/WaveformSource.cpp:98-108): out[i] = table[(index + i*step) & mask]."""
    t = np.ascontiguousarray(table_raw)
    out = np.empty((n, t.shape[1]), dtype=t.dtype)
    lib().oracle_table_walk(dtype_code, t.ctypes.data, t.shape[0], index & (2**64 - 1), step & (2**64 - 1), out.ctypes.data, n)
    return out


def noise_stream(dtype_code: int, wave: str, mean: float, b: float, seed: int, work_elems, refill_before=None,
                 ampl: complex = 1.0, offset: complex = 0.0):
    """A NoiseSource (waveform/NoiseSource.cpp) seeded with `seed`: activate(), then one work() per entry of
    work_elems (a setter call first where refill_before[w] is set).  Returns (stream [sum, ncomp], table [4096, ncomp])."""
    ampl, offset = complex(ampl), complex(offset)
    nc = 2 if is_complex(dtype_code) else 1
    we = np.ascontiguousarray(work_elems, dtype=np.uintp)
    rb = np.ascontiguousarray(refill_before if refill_before is not None else np.zeros(we.size), dtype=np.int32)
    out = np.empty((int(we.sum()), nc), dtype=scalar_np(dtype_code))
    table = np.empty((4096, nc), dtype=scalar_np(dtype_code))
    rc = lib().oracle_noise_stream(dtype_code, wave.encode(), mean, b, ampl.real, ampl.imag, offset.real, offset.imag, seed,
                                   we.ctypes.data, rb.ctypes.data, we.size, out.ctypes.data, table.ctypes.data)
    if rc:
        raise ValueError(f"NoiseSource::setWaveform({wave}): unknown waveform setting")
    return out, table
