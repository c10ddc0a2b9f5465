"""ctypes driver of the C++ block layer (libb200comms_blocks.so) for tests.

Mirrors how the reference's tests talk to blocks through Pothos proxies
(``BlockRegistry::make(path, args...)`` then ``block.call("setTaps", taps)``,
filter/TestFIRFilter.cpp:28-31), with a single-block harness playing the roles of
/blocks/feeder_source, the scheduler and /blocks/collector_sink.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _abi
from .handles import dtype_code, ncomp, np_scalar

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200comms_blocks.so")
_lib = None


class PothosException(RuntimeError):
    pass


class InvalidArgumentException(PothosException, ValueError):
    pass


def lib():
    global _lib
    if _lib is None:
        _abi.lib()
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, i, ll, ull, cp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_char_p
        L.b200c_blk_last_error.restype = cp
        L.b200c_blk_registry_has.argtypes = [cp]
        L.b200c_blk_make.restype = vp
        L.b200c_blk_make.argtypes = [cp, cp, cp, sz, i, sz, sz, ctypes.POINTER(i)]
        L.b200c_blk_destroy.argtypes = [vp]
        L.b200c_blk_call_taps.argtypes = [vp, cp, vp, sz, i]
        L.b200c_blk_call_size.argtypes = [vp, cp, sz]
        L.b200c_blk_call_bool.argtypes = [vp, cp, i]
        L.b200c_blk_call_string.argtypes = [vp, cp, cp]
        L.b200c_blk_get_size.argtypes = [vp, cp, ctypes.POINTER(sz)]
        L.b200c_blk_get_bool.argtypes = [vp, cp, ctypes.POINTER(i)]
        L.b200c_blk_get_string.argtypes = [vp, cp, cp, sz]
        L.b200c_blk_get_taps.argtypes = [vp, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(i)]
        L.b200c_blk_has_call.argtypes = [vp, cp]
        L.b200c_blk_activate.argtypes = [vp]
        L.b200c_blk_post_label.argtypes = [vp, cp, i, sz, ctypes.c_double, ull, sz]
        L.b200c_blk_feed.restype = ll
        L.b200c_blk_feed.argtypes = [vp, vp, sz]
        L.b200c_blk_run.argtypes = [vp]
        L.b200c_blk_pending.restype = ll
        L.b200c_blk_pending.argtypes = [vp]
        L.b200c_blk_collect.restype = ll
        L.b200c_blk_collect.argtypes = [vp, vp, sz]
        L.b200c_blk_reserve.restype = sz
        L.b200c_blk_reserve.argtypes = [vp]
        L.b200c_blk_total_consumed.restype = ull
        L.b200c_blk_total_consumed.argtypes = [vp]
        L.b200c_blk_work_calls.restype = ull
        L.b200c_blk_work_calls.argtypes = [vp]
        L.b200c_blk_input_domain.argtypes = [vp, cp, sz]
        L.b200c_blk_num_out_labels.restype = ll
        L.b200c_blk_num_out_labels.argtypes = [vp]
        L.b200c_blk_out_label.argtypes = [vp, sz, cp, sz, ctypes.POINTER(ull), ctypes.POINTER(sz), ctypes.POINTER(i),
                                          ctypes.POINTER(ctypes.c_double)]
        L.b200c_blk_make_dtype.restype = vp
        L.b200c_blk_make_dtype.argtypes = [cp, cp, sz, sz, ctypes.POINTER(i)]
        L.b200c_blk_get_complex.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_double)]
        L.b200c_blk_last_signal_value.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(sz)]
        L.b200c_blk_make_noargs.restype = vp
        L.b200c_blk_make_noargs.argtypes = [cp, ctypes.POINTER(i)]
        L.b200c_blk_call_double.argtypes = [vp, cp, ctypes.c_double]
        L.b200c_blk_get_double.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_double)]
        L.b200c_blk_get_doubles.argtypes = [vp, cp, vp, sz, ctypes.POINTER(sz)]
        L.b200c_blk_connect_signal.argtypes = [vp, cp, vp, cp]
        L.b200c_blk_last_signal.argtypes = [vp, cp, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(i), ctypes.POINTER(sz)]
        L.b200c_blk_has_signal.argtypes = [vp, cp]
        L.b200c_blk_call_complex.argtypes = [vp, cp, ctypes.c_double, ctypes.c_double]
        L.b200c_blk_run_source.argtypes = [vp, sz, sz]
        L.b200c_blk_total_produced.restype = ull
        L.b200c_blk_total_produced.argtypes = [vp]
        L.b200c_blk_seam_windows.restype = ull
        L.b200c_blk_seam_windows.argtypes = [vp]
        L.b200c_blk_ring_bytes.restype = sz
        L.b200c_blk_ring_bytes.argtypes = [vp]
        L.b200c_blk_run_host_chain.restype = ll
        L.b200c_blk_run_host_chain.argtypes = [vp, vp, sz, sz, vp, sz, ctypes.POINTER(ull)]
        L.b200c_blk_stream_bench.restype = ctypes.c_double
        L.b200c_blk_stream_bench.argtypes = [vp, vp, sz, sz, sz]
        _lib = L
    return _lib


def _check(rc: int):
    if rc == 0:
        return
    text = lib().b200c_blk_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise InvalidArgumentException(text)
    raise PothosException(text)


def registry_has(path: str) -> bool:
    return bool(lib().b200c_blk_registry_has(path.encode()))


_SIZE_SETTERS = {"setDecimation", "setInterpolation", "setWindow"}
_BOOL_SETTERS = {"setWaitTaps", "setInverse"}
_STRING_SETTERS = {"setFrameStartId", "setFrameEndId", "setLabelId", "setMode", "setWaveform"}
_DOUBLE_SETTERS = {"setFactor", "setPhase", "setRate", "setFrequency", "setSampleRate", "setResolution", "setMean", "setB"}
_COMPLEX_SETTERS = {"setOffset", "setAmplitude"}
_COMPLEX_GETTERS = {"value", "getOffset", "getAmplitude"}
_SIZE_GETTERS = {"getDecimation", "getInterpolation", "getNumBins", "getWindow"}
_BOOL_GETTERS = {"getWaitTaps", "getInverse"}
_STRING_GETTERS = {"getFrameStartId", "getFrameEndId", "getLabelId", "getMode", "getWaveform"}
_DOUBLE_GETTERS = {"getFactor", "getPhase", "getRate", "getFrequency", "getSampleRate", "getResolution", "getMean", "getB"}


class Block:
    """One block instance inside the test harness (device buffer managers in HBM)."""

    def __init__(self, path: str, dtype: str, *args, in_bytes: int = 8 << 20, out_bytes: int = 16 << 20):
        self.dtype = dtype_code(dtype) if dtype in _abi.DTYPE_CODES else -1
        self.dtype_name = dtype
        status = ctypes.c_int(0)
        if len(args) == 0:   # /comms/scale, /comms/rotate, /comms/signal_probe, the two sources: factory (dtype)
            self._h = lib().b200c_blk_make_dtype(path.encode(), dtype.encode(), in_bytes, out_bytes, ctypes.byref(status))
        elif len(args) == 1:   # /comms/fir_filter(dtype, tapsType)
            self._h = lib().b200c_blk_make(path.encode(), dtype.encode(), str(args[0]).encode(), 0, 0, in_bytes, out_bytes,
                                           ctypes.byref(status))
        elif len(args) == 2:  # /comms/fft(dtype, numBins, inverse)
            self._h = lib().b200c_blk_make(path.encode(), dtype.encode(), None, int(args[0]), int(bool(args[1])), in_bytes,
                                           out_bytes, ctypes.byref(status))
        else:
            raise InvalidArgumentException("wrong number of factory arguments")
        _check(status.value)
        self._elems_fed = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().b200c_blk_destroy(self._h)
            self._h = None

    __del__ = close

    def call(self, name: str, *args):
        n = name.encode()
        L = lib()
        if name == "setTaps":
            taps = np.asarray(args[0])
            if np.iscomplexobj(taps):
                t = np.ascontiguousarray(taps, dtype=np.complex128).view(np.float64)
                _check(L.b200c_blk_call_taps(self._h, n, t.ctypes.data, taps.size, 1))
            else:
                t = np.ascontiguousarray(taps, dtype=np.float64)
                _check(L.b200c_blk_call_taps(self._h, n, t.ctypes.data, t.size, 0))
            return None
        if name == "getTaps":
            cnt, cx = ctypes.c_size_t(0), ctypes.c_int(0)
            buf = np.zeros(1 << 16, dtype=np.float64)
            _check(L.b200c_blk_get_taps(self._h, buf.ctypes.data, buf.size, ctypes.byref(cnt), ctypes.byref(cx)))
            return buf[: 2 * cnt.value].view(np.complex128).copy() if cx.value else buf[: cnt.value].copy()
        if name in _SIZE_SETTERS:
            _check(L.b200c_blk_call_size(self._h, n, int(args[0])))
            return None
        if name in _BOOL_SETTERS:
            _check(L.b200c_blk_call_bool(self._h, n, int(bool(args[0]))))
            return None
        if name in _STRING_SETTERS:
            _check(L.b200c_blk_call_string(self._h, n, str(args[0]).encode()))
            return None
        if name in _DOUBLE_SETTERS:
            _check(L.b200c_blk_call_double(self._h, n, float(args[0])))
            return None
        if name in _DOUBLE_GETTERS:
            v = ctypes.c_double(0)
            _check(L.b200c_blk_get_double(self._h, n, ctypes.byref(v)))
            return v.value
        if name in _COMPLEX_SETTERS:
            z = complex(args[0])
            _check(L.b200c_blk_call_complex(self._h, n, z.real, z.imag))
            return None
        if name in _COMPLEX_GETTERS:   # SignalProbe::value(): double or complex<double>; getOffset / getAmplitude
            v = (ctypes.c_double * 2)()
            _check(L.b200c_blk_get_complex(self._h, n, v))
            return complex(v[0], v[1])
        if name in _SIZE_GETTERS:
            v = ctypes.c_size_t(0)
            _check(L.b200c_blk_get_size(self._h, n, ctypes.byref(v)))
            return v.value
        if name in _BOOL_GETTERS:
            v = ctypes.c_int(0)
            _check(L.b200c_blk_get_bool(self._h, n, ctypes.byref(v)))
            return bool(v.value)
        if name in _STRING_GETTERS:
            buf = ctypes.create_string_buffer(256)
            _check(L.b200c_blk_get_string(self._h, n, buf, 256))
            return buf.value.decode()
        raise PothosException(f"Block.call({name}): no such registered call in the driver")

    def has_call(self, name: str) -> bool:
        return bool(lib().b200c_blk_has_call(self._h, name.encode()))

    def has_signal(self, name: str) -> bool:
        return bool(lib().b200c_blk_has_signal(self._h, name.encode()))

    def last_signal_value(self, signal: str = "valueChanged"):
        """(last payload of a value-carrying signal as complex, number of emissions)"""
        v, count = (ctypes.c_double * 2)(), ctypes.c_size_t(0)
        _check(lib().b200c_blk_last_signal_value(self._h, signal.encode(), v, ctypes.byref(count)))
        return complex(v[0], v[1]), count.value

    def activate(self):
        _check(lib().b200c_blk_activate(self._h))

    def post_label(self, label_id: str, index: int, data=None, width: int = 1):
        kind, sval, dval = 0, 0, 0.0
        if isinstance(data, float):
            kind, dval = 2, data
        elif data is not None:
            kind, sval = 1, int(data)
        _check(lib().b200c_blk_post_label(self._h, label_id.encode(), kind, sval, dval, index, width))

    def feed(self, x_raw: np.ndarray) -> int:
        nc = ncomp(self.dtype)
        x = np.ascontiguousarray(x_raw, dtype=np_scalar(self.dtype)).reshape(-1, nc)
        n = lib().b200c_blk_feed(self._h, x.ctypes.data, x.shape[0])
        if n < 0:
            _check(int(n))
        return int(n)

    def run(self):
        _check(lib().b200c_blk_run(self._h))

    def run_source(self, nwork: int = 1, elems: int = 0) -> np.ndarray:
        """A source block: `nwork` work() calls, each offered room for `elems` elements (0 = the whole output
        buffer); returns everything produced."""
        _check(lib().b200c_blk_run_source(self._h, nwork, elems))
        return self.collect()

    def collect(self) -> np.ndarray:
        nc = ncomp(self.dtype)
        n = int(lib().b200c_blk_pending(self._h))
        out = np.empty((n, nc), dtype=np_scalar(self.dtype))
        if n:
            got = lib().b200c_blk_collect(self._h, out.ctypes.data, n)
            assert got == n
        return out

    def push_through(self, x_raw: np.ndarray) -> np.ndarray:
        """feed everything (in ring-sized pieces), running the block in between; returns all output."""
        nc = ncomp(self.dtype)
        x = np.ascontiguousarray(x_raw, dtype=np_scalar(self.dtype)).reshape(-1, nc)
        outs, pos = [], 0
        while pos < x.shape[0]:
            n = self.feed(x[pos:])
            pos += n
            self.run()
            outs.append(self.collect())
            if n == 0 and outs[-1].shape[0] == 0:
                raise PothosException("harness made no progress (ring full and block idle)")
        return np.concatenate(outs) if outs else np.empty((0, nc), dtype=x.dtype)

    @property
    def reserve(self) -> int:
        return int(lib().b200c_blk_reserve(self._h))

    @property
    def total_consumed(self) -> int:
        return int(lib().b200c_blk_total_consumed(self._h))

    @property
    def work_calls(self) -> int:
        return int(lib().b200c_blk_work_calls(self._h))

    @property
    def total_produced(self) -> int:
        return int(lib().b200c_blk_total_produced(self._h))

    @property
    def seam_windows(self) -> int:
        """work() calls whose readable window (history + new) ran across the end of the ring's first mapping."""
        return int(lib().b200c_blk_seam_windows(self._h))

    @property
    def ring_bytes(self) -> int:
        return int(lib().b200c_blk_ring_bytes(self._h))

    def run_host_chain(self, x_raw: np.ndarray, chunk: int, out_capacity: int):
        """feeder (host) -> /b200c/host_to_hbm -> this block -> /b200c/hbm_to_host -> collector (host), the way a topology
        with host-memory neighbours is wired around the device blocks (blocks/Bridge.cpp).  Returns (output, bridge work() calls)."""
        nc = ncomp(self.dtype)
        x = np.ascontiguousarray(x_raw, dtype=np_scalar(self.dtype)).reshape(-1, nc)
        out = np.empty((out_capacity, nc), dtype=x.dtype)
        calls = ctypes.c_ulonglong(0)
        got = lib().b200c_blk_run_host_chain(self._h, x.ctypes.data, x.shape[0], chunk, out.ctypes.data, out_capacity, ctypes.byref(calls))
        if got < 0:
            _check(int(got))
        return out[:got], calls.value

    def stream_bench(self, pattern_raw: np.ndarray, chunk_elems: int, rounds: int) -> float:
        """Seconds for `rounds` rounds of `chunk_elems` new elements through work() between device neighbours
        (blocks/Harness.cpp streamBench: no host<->device copy inside the loop)."""
        nc = ncomp(self.dtype)
        x = np.ascontiguousarray(pattern_raw, dtype=np_scalar(self.dtype)).reshape(-1, nc)
        secs = lib().b200c_blk_stream_bench(self._h, x.ctypes.data, x.shape[0], chunk_elems, rounds)
        if secs < 0:
            _check(int(secs))
        return float(secs)

    @property
    def input_domain(self) -> str:
        buf = ctypes.create_string_buffer(64)
        lib().b200c_blk_input_domain(self._h, buf, 64)
        return buf.value.decode()

    def out_labels(self):
        L = lib()
        res = []
        for i in range(int(L.b200c_blk_num_out_labels(self._h))):
            idb = ctypes.create_string_buffer(128)
            index, width, kind, val = ctypes.c_ulonglong(0), ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_double(0)
            _check(L.b200c_blk_out_label(self._h, i, idb, 128, ctypes.byref(index), ctypes.byref(width), ctypes.byref(kind),
                                         ctypes.byref(val)))
            data = None if kind.value == 0 else (int(val.value) if kind.value == 1 else val.value)
            res.append({"id": idb.value.decode(), "index": index.value, "width": width.value, "data": data})
        return res


_D_STRING = {"setFilterType": "filterType", "setBandType": "bandType", "setWindowType": "windowType"}
_D_DOUBLE = {"setSampleRate": "sampleRate", "setFrequencyLower": "frequencyLower", "setFrequencyUpper": "frequencyUpper",
             "setBandwidthTrans": "bandwidthTrans", "setAlpha": "alpha", "setStopDB": "stopDB", "setPassDB": "passDB",
             "setGain": "gain"}
_D_SIZE = {"setNumTaps": "numTaps"}
_D_VECTOR = {"setWindowArgs": "windowArgs", "setFrequencies": None}


class HostBlock:
    """A port-less host block (/comms/fir_designer, /blocks/fir_designer, /comms/window_designer):
    registered calls, activate() and the "tapsChanged" signal (filter/FIRDesigner.cpp:143-169)."""

    def __init__(self, path: str):
        status = ctypes.c_int(0)
        self._h = lib().b200c_blk_make_noargs(path.encode(), ctypes.byref(status))
        _check(status.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().b200c_blk_destroy(self._h)
            self._h = None

    __del__ = close

    def has_call(self, name: str) -> bool:
        return bool(lib().b200c_blk_has_call(self._h, name.encode()))

    def has_signal(self, name: str) -> bool:
        return bool(lib().b200c_blk_has_signal(self._h, name.encode()))

    def activate(self):
        _check(lib().b200c_blk_activate(self._h))

    def call(self, name: str, *args):
        n, L = name.encode(), lib()
        if name in _D_STRING:
            return _check(L.b200c_blk_call_string(self._h, n, str(args[0]).encode()))
        if name in _D_DOUBLE:
            return _check(L.b200c_blk_call_double(self._h, n, float(args[0])))
        if name in _D_SIZE:
            return _check(L.b200c_blk_call_size(self._h, n, int(args[0])))
        if name in _D_VECTOR:
            v = np.ascontiguousarray(args[0], dtype=np.float64)
            return _check(L.b200c_blk_call_taps(self._h, n, v.ctypes.data, v.size, 0))
        if name in _D_STRING.values():
            buf = ctypes.create_string_buffer(256)
            _check(L.b200c_blk_get_string(self._h, n, buf, 256))
            return buf.value.decode()
        if name in _D_DOUBLE.values():
            v = ctypes.c_double(0)
            _check(L.b200c_blk_get_double(self._h, n, ctypes.byref(v)))
            return v.value
        if name in _D_SIZE.values():
            v = ctypes.c_size_t(0)
            _check(L.b200c_blk_get_size(self._h, n, ctypes.byref(v)))
            return v.value
        if name == "windowArgs":
            cnt, buf = ctypes.c_size_t(0), np.zeros(64)
            _check(L.b200c_blk_get_doubles(self._h, n, buf.ctypes.data, buf.size, ctypes.byref(cnt)))
            return buf[: cnt.value].copy()
        raise PothosException(f"HostBlock.call({name}): no such registered call in the driver")

    def connect(self, signal: str, dst, slot: str):
        """Topology::connect(self, signal, dst, slot): every emission calls dst.<slot>(payload)."""
        _check(lib().b200c_blk_connect_signal(self._h, signal.encode(), dst._h, slot.encode()))

    def last_signal(self, signal: str = "tapsChanged"):
        """(payload of the last emission or None, number of emissions so far)"""
        cnt, cx, count = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_size_t(0)
        buf = np.zeros(1 << 16, dtype=np.float64)
        _check(lib().b200c_blk_last_signal(self._h, signal.encode(), buf.ctypes.data, buf.size, ctypes.byref(cnt), ctypes.byref(cx),
                                           ctypes.byref(count)))
        if count.value == 0:
            return None, 0
        return (buf[: 2 * cnt.value].view(np.complex128).copy() if cx.value else buf[: cnt.value].copy()), count.value


def make(path: str, dtype: str | None = None, *args, **kw):
    """BlockRegistry::make(path, ...) -- /comms/fir_filter, /blocks/fir_filter, /comms/fft (dtype, ...);
    /comms/fir_designer, /blocks/fir_designer, /comms/window_designer (no arguments)."""
    if dtype is None:
        return HostBlock(path)
    return Block(path, dtype, *args, **kw)
