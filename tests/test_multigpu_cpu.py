"""world_size-2 (and 3) gloo tests of the multi-GPU partitioning logic on CPU.

No GPU here, so the per-rank arithmetic is done by the CPU oracle; what is under test is the
host-side algebra the N>1 bench path uses: M-aligned contiguous segments + a K-1 sample halo
exchanged between neighbours must reproduce the single-stream result exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, code, tcx, taps, M, L, x_full, K, q, overlapped=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from pothoscomms_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total_new = x_full.shape[0] - (K - 1)
        start, stop = sharding.segment_bounds(total_new, world, M)[rank]
        # [K-1 halo | segment]; only rank 0 knows the stream's true first K-1 samples
        buf = torch.zeros((K - 1 + stop - start, x_full.shape[1]), dtype=torch.from_numpy(x_full).dtype)
        buf[K - 1:] = torch.from_numpy(x_full[K - 1 + start: K - 1 + stop])
        if rank == 0:
            buf[: K - 1] = torch.from_numpy(x_full[: K - 1])
        if overlapped:
            # what bench.py does at N > 1: start the P2P, compute everything that does not read the
            # halo while it is in flight, then the q0 halo-dependent blocks
            q0, in0, out0 = sharding.split_at_halo(K, M, L)
            works = sharding.start_halo_exchange(buf, K, rank, world)
            y1, c1, p1 = oracle.fir(code, tcx, taps, M, L, buf.numpy()[in0:])
            sharding.finish_halo_exchange(works)
            y0, c0, p0 = oracle.fir(code, tcx, taps, M, L, buf.numpy()[: in0 + K - 1])
            assert p0 == out0
            y, cons = np.concatenate([y0, y1]), c0 + c1
        else:
            sharding.exchange_halo(buf, K, rank, world)
            y, cons, prod = oracle.fir(code, tcx, taps, M, L, buf.numpy())
        assert cons == stop - start
        q.put((rank, y))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,overlapped", [(2, False), (3, False), (2, True)])
@pytest.mark.parametrize("case", ["cf32_resampler", "ci16_fir"])
def test_segments_with_halo_reproduce_single_stream(oracle, world, overlapped, case):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    if case == "cf32_resampler":
        code, tcx, M, L = oracle.CF32, False, 2, 3
        taps = rng.standard_normal(255) * 0.05
        x = rng.standard_normal((20011, 2)).astype(np.float32)
    else:
        code, tcx, M, L = oracle.CI16, True, 1, 1
        taps = (rng.standard_normal(128) + 1j * rng.standard_normal(128)) * 0.05
        x = rng.integers(-30000, 30000, size=(15000, 2), dtype=np.int16)
    K = oracle.fir_K(len(taps), L)
    y_ref, cons_ref, prod_ref = oracle.fir(code, tcx, taps, M, L, x)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, code, tcx, taps, M, L, x, K, q, overlapped)) for r in range(world)]
    for p in procs:
        p.start()
    parts = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    y = np.concatenate([parts[r] for r in range(world)])
    assert y.shape[0] == prod_ref
    assert np.array_equal(y, y_ref)   # same oracle arithmetic per output => exact, float included


def test_segment_bounds_are_M_aligned_and_cover():
    from pothoscomms_b200 import sharding
    for total, world, M in ((1 << 20, 8, 1), (1000003, 8, 2), (999, 4, 7), (5, 8, 1)):
        b = sharding.segment_bounds(total, world, M)
        assert b[0][0] == 0 and b[-1][1] == total // M * M
        for (s0, e0), (s1, _) in zip(b, b[1:]):
            assert e0 == s1
        assert all(s % M == 0 and e % M == 0 for s, e in b)
    assert [sharding.channel_range(1024, 8, r) for r in range(8)] == [(128 * r, 128 * (r + 1)) for r in range(8)]


def test_segments_shorter_than_the_halo_are_rejected():
    """A segment shorter than K-1 would forward halo samples it has only just received (its tail and its
    halo overlap in [halo | segment]): refused, never silently wrong."""
    from pothoscomms_b200 import sharding
    with pytest.raises(ValueError, match="shorter than the K-1"):
        sharding.segment_bounds(1000, 8, 1, K=256)
    with pytest.raises(ValueError, match="shorter than the K-1"):
        sharding.check_segment(100, 128, rank=1, world=2)
    sharding.check_segment(127, 128, rank=1, world=2)
    sharding.check_segment(5, 128, rank=0, world=1)      # a single rank has no neighbour
    assert len(sharding.segment_bounds(1 << 16, 8, 2, K=256)) == 8
