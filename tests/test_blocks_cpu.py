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
