"""The C-ABI library must load on a GPU-less host, export every symbol include/b200comms.h
declares, and FAIL LOUDLY (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200comms.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200c_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pothoscomms_b200 import _abi
    lib = _abi.lib()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200comms.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_abi.SYMBOLS) == names


def test_abi_version_and_dtype_sizes():
    from pothoscomms_b200 import _abi
    lib = _abi.lib()
    assert lib.b200c_abi_version() == 2
    sizes = [lib.b200c_dtype_size(i) for i in range(12)]
    assert sizes == [4, 8, 8, 16, 1, 2, 2, 4, 4, 8, 8, 16]
    assert lib.b200c_dtype_size(99) == 0


def test_no_link_time_dependency_on_libcuda():
    import subprocess
    from pothoscomms_b200 import _abi
    out = subprocess.run(["readelf", "-d", _abi.LIB_PATH], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any(n.startswith("libcuda.so") for n in needed), needed
    assert not any("oracle" in n for n in needed), "product library must not link the oracle"


def test_argument_errors_do_not_need_a_device():
    from pothoscomms_b200 import _abi
    lib = _abi.lib()
    h = ctypes.c_void_p()
    # real data + COMPLEX taps is not in the factory table (filter/FIRFilter.cpp:373-383)
    assert lib.b200c_fir_create(ctypes.byref(h), _abi.F32, _abi.TAPS_COMPLEX, 0) == _abi.ERR_UNSUPPORTED
    assert b"unsupported types" in lib.b200c_last_error()
    assert lib.b200c_fir_create(ctypes.byref(h), 77, _abi.TAPS_REAL, 0) == _abi.ERR_UNSUPPORTED
    # FFTFactory only knows cf64, cf32, complex int16 (fft/FFT.cpp:89-92)
    assert lib.b200c_fft_create(ctypes.byref(h), _abi.F32, 1024, 0, 0) == _abi.ERR_UNSUPPORTED
    assert lib.b200c_fft_create(ctypes.byref(h), _abi.CI32, 1024, 0, 0) == _abi.ERR_UNSUPPORTED


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the GPU-less behaviour")
    from pothoscomms_b200 import _abi, FirFilter, Fft
    n = ctypes.c_int(-1)
    assert _abi.lib().b200c_device_count(ctypes.byref(n)) == 0 and n.value == 0
    with pytest.raises(_abi.B200CommsError) as e:
        FirFilter("complex_float32", "COMPLEX")
    assert e.value.code == _abi.ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_abi.B200CommsError):
        Fft("complex_float32", 1024, False)


def test_product_package_never_imports_the_oracle():
    # the oracle is test infrastructure; the product path must not route through it
    pkg = os.path.join(ROOT, "pothoscomms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "libkissref" not in src, f
