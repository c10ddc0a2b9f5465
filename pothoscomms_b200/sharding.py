"""Multi-GPU partitioning of the FIR path, host-side logic only.

One long stream is cut into contiguous per-rank segments whose starts are multiples of the
decimation M (the reference's decimation counter restarts at M on every work() call,
filter/FIRFilter.cpp:283,291-292, so any M-aligned cut reproduces the single-stream output).
The only cross-rank dependency is the K-1 samples of left history (:281,298): rank r sends the
last K-1 samples of its segment to rank r+1 -- one NCCL P2P over NVLink per pass (gloo in the
CPU tests).  Independent channels / FFT batches need no exchange at all: `channel_range`.
"""
from __future__ import annotations


def segment_bounds(total_new: int, world: int, decim: int):
    """[(start, stop)] over the `total_new` consumable samples (history excluded), every start a
    multiple of `decim`, lengths as even as possible."""
    blocks = total_new // decim
    bounds = []
    for r in range(world):
        b0, b1 = blocks * r // world, blocks * (r + 1) // world
        bounds.append((b0 * decim, b1 * decim))
    return bounds


def channel_range(num_channels: int, world: int, rank: int):
    """Contiguous block of channels (or FFT batches) owned by `rank`: zero exchange."""
    return num_channels * rank // world, num_channels * (rank + 1) // world


def start_halo_exchange(buf, K: int, rank: int, world: int):
    """Asynchronous form of `exchange_halo`: starts the grouped P2P and returns its work handles;
    `finish_halo_exchange` makes the current stream wait for them.  Between the two calls the caller
    may launch everything that does not read buf[:K-1] (see `split_at_halo`)."""
    import torch
    import torch.distributed as dist
    if world == 1 or K <= 1:
        return []
    ops = []
    n = buf.shape[0]
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, buf[n - (K - 1):].view(torch.uint8), rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, buf[: K - 1].view(torch.uint8), rank - 1))
    return dist.batch_isend_irecv(ops) if ops else []


def finish_halo_exchange(works):
    for w in works:
        w.wait()


def split_at_halo(K: int, decim: int, interp: int, align: int = 1):
    """Blocks q < q0 of a segment's output read the neighbour's halo, blocks q >= q0 read only the
    segment itself (its own first samples serve as their history): q0 = ceil((K-1) / M), rounded up
    to a multiple of `align`.  Returns
    (q0, first buffer element of the halo-free call, first output element of the halo-free call);
    the head call takes buf[: q0*M + K-1] and writes out[: q0*L]."""
    q0 = -(-(K - 1) // decim)
    q0 = -(-q0 // align) * align      # align = 16 keeps both sub-buffers 16-byte aligned for any element size
    return q0, q0 * decim, q0 * interp


def exchange_halo(buf, K: int, rank: int, world: int):
    """`buf` is [K-1 halo | segment] on every rank.  Sends this rank's last K-1 samples to
    rank+1 and receives rank-1's into buf[:K-1]; rank 0 keeps its own halo (the stream's true
    first K-1 samples, history only).  One grouped P2P; works with nccl (GPU) and gloo (CPU)."""
    import torch.distributed as dist
    if world == 1 or K <= 1:
        return
    import torch
    ops = []
    n = buf.shape[0]
    # the halo travels as raw bytes: NCCL has no int16 (the complex-int16 streams of config C2)
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, buf[n - (K - 1):].view(torch.uint8), rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, buf[: K - 1].view(torch.uint8), rank - 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
