"""Block-layer checks that need no GPU: registry paths, factory errors, loud failure."""
import pytest


def test_registry_paths():
    from pothoscomms_b200 import blocks
    # filter/FIRFilter.cpp:385-389 and fft/FFT.cpp:94-95
    for path in ("/comms/fir_filter", "/blocks/fir_filter", "/comms/fft"):
        assert blocks.registry_has(path), path
    assert not blocks.registry_has("/comms/iir_filter")   # out of scope


def test_factory_rejects_unsupported_types():
    from pothoscomms_b200 import blocks
    with pytest.raises(blocks.InvalidArgumentException, match="unsupported types"):   # FIRFilter.cpp:383
        blocks.make("/comms/fir_filter", "float32", "COMPLEX")
    with pytest.raises(blocks.InvalidArgumentException, match="unsupported types"):
        blocks.make("/comms/fir_filter", "complex_float32", "BOGUS")
    with pytest.raises(blocks.InvalidArgumentException, match="unsupported type"):    # FFT.cpp:92
        blocks.make("/comms/fft", "float32", 1024, False)
    with pytest.raises(blocks.InvalidArgumentException, match="unsupported type"):
        blocks.make("/comms/fft", "complex_int32", 1024, False)
    with pytest.raises(blocks.PothosException):
        blocks.make("/comms/nope", "float32", "REAL")


def test_block_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pothoscomms_b200 import blocks
    with pytest.raises(blocks.PothosException, match="no CPU fallback"):
        blocks.make("/comms/fir_filter", "complex_float32", "COMPLEX")
    with pytest.raises(blocks.PothosException, match="no CPU fallback"):
        blocks.make("/comms/fft", "complex_float32", 1024, False)


def test_neighbour_blocks_registry_and_factory_errors():
    """/comms/scale (math/Scale.cpp:155-156), /comms/rotate (math/Rotate.cpp:159-160), /comms/signal_probe and
    its legacy path (utility/SignalProbe.cpp:188-192): registered; unsupported types rejected before any
    device is touched."""
    from pothoscomms_b200 import blocks
    for path in ("/comms/scale", "/comms/rotate", "/comms/signal_probe", "/blocks/stream_probe"):
        assert blocks.registry_has(path), path
    with pytest.raises(blocks.InvalidArgumentException, match="unsupported type"):   # Rotate.cpp:157: complex only
        blocks.make("/comms/rotate", "int16")
    with pytest.raises(blocks.PothosException):
        blocks.make("/comms/scale", "uint8")


def test_source_blocks_registry_and_factory_errors():
    """/comms/waveform_source + /blocks/waveform_source (waveform/WaveformSource.cpp:289-293) and
    /comms/noise_source + /blocks/noise_source (waveform/NoiseSource.cpp:285-289): registered; a type outside the
    twelve factory rows is rejected before any device is touched."""
    from pothoscomms_b200 import blocks
    for path in ("/comms/waveform_source", "/blocks/waveform_source", "/comms/noise_source", "/blocks/noise_source"):
        assert blocks.registry_has(path), path
    with pytest.raises(blocks.InvalidArgumentException):
        blocks.make("/comms/waveform_source", "uint8")
    with pytest.raises(blocks.InvalidArgumentException):
        blocks.make("/comms/noise_source", "uint8")


def test_source_blocks_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pothoscomms_b200 import blocks
    for path in ("/comms/waveform_source", "/comms/noise_source"):
        with pytest.raises(blocks.PothosException, match="no CPU fallback"):
            blocks.make(path, "complex_float32")


def test_table_source_argument_errors_and_loud_failure():
    import ctypes
    import torch
    from pothoscomms_b200 import _abi
    lib = _abi.lib()
    buf = (ctypes.c_char * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.b200c_table_source(77, p, 4096, 0, 1, p, 4, 0, None) == _abi.ERR_UNSUPPORTED
    assert lib.b200c_table_source(_abi.CF32, p, 4095, 0, 1, p, 4, 0, None) == _abi.ERR_INVALID     # mask needs a power of two
    assert lib.b200c_table_source(_abi.CF32, p, 0, 0, 1, p, 4, 0, None) == _abi.ERR_INVALID
    assert lib.b200c_table_source(_abi.CF32, p, 4096, 0, 1, p, 0, 0, None) == _abi.OK              # nothing to produce
    if not torch.cuda.is_available():
        assert lib.b200c_table_source(_abi.CF32, p, 4, 0, 1, p, 4, 0, None) == _abi.ERR_CUDA
        assert b"no CPU fallback" in lib.b200c_last_error()


def test_bridge_blocks_are_registered():
    """/b200c/host_to_hbm and /b200c/hbm_to_host (blocks/Bridge.cpp): the copy blocks a topology puts between a
    host-memory neighbour and the device blocks, which refuse host-domain peers (Pothos::PortDomainError)."""
    from pothoscomms_b200 import blocks
    for path in ("/b200c/host_to_hbm", "/b200c/hbm_to_host"):
        assert blocks.registry_has(path), path
