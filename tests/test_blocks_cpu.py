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
