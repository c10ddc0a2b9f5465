"""Multi-GPU correctness ON HARDWARE (needs >= 2 visible GPUs; `gpurun --gpus 2 -- python -m pytest
tests/test_multigpu_gpu.py -m gpu`): the peer-memory halo exchange of include/b200comms.h
(b200c_peer_* + b200c_halo_exchange) and the segment algebra of SURVEY.md 8e against the oracle's
single-stream output (filter/FIRFilter.cpp:281,283: K-1 history, decimation phase restarts per call)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")


@pytest.mark.parametrize("case", ["cf32_headline", "cf32_resampler", "ci16_c2"])
def test_two_gpu_segments_with_peer_halo_match_single_stream(oracle, cuda_device, case):
    _need_two()
    import torch

    from pothoscomms_b200 import FirFilter, sharding
    from pothoscomms_b200 import workloads as wl
    cfg = {"cf32_headline": ("headline", oracle.CF32, 1, 1), "cf32_resampler": ("c3", oracle.CF32, 2, 3),
           "ci16_c2": ("c2", oracle.CI16, 1, 1)}[case]
    taps, tt = wl.config_taps(cfg[0])
    code, M, L = cfg[1], cfg[2], cfg[3]
    K = oracle.fir_K(len(taps), L)
    world, n_new = 2, 1 << 18
    x = wl.tone_noise_numpy(code, K - 1 + n_new, seed=99)
    y_ref, c_ref, p_ref = oracle.fir(code, tt == "COMPLEX", taps, M, L, x)
    bounds = sharding.segment_bounds(n_new, world, M, K=K)
    bufs, firs = [], []
    for r, (s0, s1) in enumerate(bounds):
        dev = torch.device("cuda", r)
        b = torch.zeros((K - 1 + s1 - s0, x.shape[1]), dtype=torch.from_numpy(x).dtype, device=dev)
        b[K - 1:] = torch.from_numpy(x[K - 1 + s0: K - 1 + s1]).to(dev)
        if r == 0:
            b[: K - 1] = torch.from_numpy(x[: K - 1]).to(dev)
        f = FirFilter(code, tt, device=r)
        f.set_taps(taps)
        f.set_rates(M, L)
        bufs.append(b); firs.append(f)
    # same-process form of the record exchange (one process drives both GPUs here; under torchrun it is one
    # all_gather_object): ranks are constructed in order, so rank r sees the records of ranks < r
    staged = []

    def gather(mine):
        staged.append(mine)
        return staged + [None] * (world - len(staged))
    links = [sharding.PeerHalo(bufs[r], K, r, world, r, gather=gather) for r in range(world)]
    parts = []
    for r in range(world):
        torch.cuda.synchronize(r)          # every tail is final (host-level ordering, as in the bench)
    for r in range(world):
        with torch.cuda.device(r):
            links[r].pull(wait=False)
            y, c, p = firs[r].run(bufs[r])
            torch.cuda.synchronize(r)
            assert c == bounds[r][1] - bounds[r][0]
            parts.append(y.cpu().numpy())
    # the halo rank 1 received is rank 0's tail, bit for bit
    assert np.array_equal(bufs[1][: K - 1].cpu().numpy(), x[bounds[1][0]: bounds[1][0] + K - 1])
    y = np.concatenate(parts)
    assert y.shape[0] == p_ref
    if code == oracle.CI16:
        assert np.array_equal(y, y_ref)
    else:
        err = np.sqrt(np.mean((y.astype(np.float64) - y_ref) ** 2)) / np.sqrt(np.mean(y_ref.astype(np.float64) ** 2))
        assert err < 1e-5, err          # north_star: 1e-5 of output RMS for float32
    for l in links:
        l.close()
