"""The driver's entry points: `__graft_entry__.smoke()` is what runs on the GPU box before the bench, so the suite runs it too
(its kernel-name assertions must follow the dispatch table)."""
import pytest


@pytest.mark.gpu
def test_smoke_entry_point(oracle, cuda_device):
    import __graft_entry__ as entry
    entry.smoke()
