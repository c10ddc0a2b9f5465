"""Multi-GPU partitioning of the FIR path, host-side logic only.

One long stream is cut into contiguous per-rank segments whose starts are multiples of the
decimation M (the reference's decimation counter restarts at M on every work() call,
filter/FIRFilter.cpp:283,291-292, so any M-aligned cut reproduces the single-stream output).
The only cross-rank dependency is the K-1 samples of left history (:281,298): rank r sends the
last K-1 samples of its segment to rank r+1 -- one NCCL P2P over NVLink per pass (gloo in the
CPU tests).  Independent channels / FFT batches need no exchange at all: `channel_range`.
"""
from __future__ import annotations


def check_segment(n_seg: int, K: int, rank: int = 0, world: int = 2) -> None:
    """A segment must hold at least K-1 samples of its own: its tail IS the neighbour's halo, and a
    shorter segment would have to forward samples it has only just received (tail and halo overlap in
    [halo | segment]).  Raised instead of silently producing wrong outputs."""
    if world > 1 and n_seg < K - 1:
        raise ValueError(f"rank {rank}: segment of {n_seg} samples is shorter than the K-1 = {K - 1} halo; "
                         f"use fewer ranks or a longer stream")


def segment_bounds(total_new: int, world: int, decim: int, K: int = 1):
    """[(start, stop)] over the `total_new` consumable samples (history excluded), every start a
    multiple of `decim`, lengths as even as possible.  With K given, partitions whose segments are
    shorter than the K-1 halo are rejected (`check_segment`)."""
    blocks = total_new // decim
    if K > 1 and world > 1 and (blocks // world) * decim < K - 1:
        raise ValueError(f"{total_new} samples over {world} ranks leaves segments shorter than the K-1 = {K - 1} halo")
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
    check_segment(n - (K - 1), K, rank, world)
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
    check_segment(n - (K - 1), K, rank, world)
    # the halo travels as raw bytes: NCCL has no int16 (the complex-int16 streams of config C2)
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, buf[n - (K - 1):].view(torch.uint8), rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, buf[: K - 1].view(torch.uint8), rank - 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class PeerHalo:
    """The halo exchange of the product path: rank r PULLS the last K-1 samples of rank r-1's segment
    straight out of the neighbour's HBM (CUDA IPC mapping, NVLink) with one copy enqueued on its own
    compute stream -- `b200c_halo_exchange` of include/b200comms.h.  No collective and no extra kernel
    sits in front of the FIR launch.  torch.distributed is used ONCE, at construction, to carry the two
    128-byte export records to the neighbour; a C++ Pothos host would use any IPC of its own.

    `buf` is this rank's [K-1 halo | segment] tensor (cudaMalloc-backed: torch's default allocator).
    `pull(wait=False)` enqueues the copy alone (the caller has ordered producer and consumer on the host);
    `mark_tail_ready()` / `pull(wait=True)` add the GPU-side edge through an interprocess event.  Note the
    CUDA semantics: a wait captures the owner's most recent record AT CALL TIME, so hosts that run several
    steps ahead of their GPUs must sequence record-before-wait themselves."""

    def __init__(self, buf, K: int, rank: int, world: int, device: int, gather=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _abi
        self._abi, self._ct = _abi, ctypes
        self.lib = _abi.lib()
        self.rank, self.world, self.device = rank, world, device
        self.active = world > 1 and K > 1
        self._mapping = self._peer_event = self._event = None
        if not self.active:
            return
        esz = buf.element_size() * buf.shape[1]
        n = buf.shape[0]
        check_segment(n - (K - 1), K, rank, world)
        self.nbytes = (K - 1) * esz
        self.dst = buf.data_ptr()
        tail_ptr = buf.data_ptr() + (n - (K - 1)) * esz
        mem = ctypes.create_string_buffer(_abi.PEER_RECORD_BYTES)
        evr = ctypes.create_string_buffer(_abi.PEER_RECORD_BYTES)
        _abi.check(self.lib.b200c_peer_export(ctypes.c_void_p(tail_ptr), self.nbytes, device, mem))
        ev = ctypes.c_void_p()
        _abi.check(self.lib.b200c_peer_event_create(ctypes.byref(ev), device, evr))
        self._event = ev
        mine = (bytes(mem.raw), bytes(evr.raw))
        if gather is None:
            records = [None] * world
            dist.all_gather_object(records, mine)
        else:
            records = gather(mine)
        self.peer_src = None
        if rank > 0:
            pmem, pev = records[rank - 1]
            ptr, mapping, pevent = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            _abi.check(self.lib.b200c_peer_open(ctypes.create_string_buffer(pmem, len(pmem)), device, ctypes.byref(ptr),
                                                ctypes.byref(mapping)))
            _abi.check(self.lib.b200c_peer_event_open(ctypes.create_string_buffer(pev, len(pev)), device, ctypes.byref(pevent)))
            self.peer_src, self._mapping, self._peer_event = ptr, mapping, pevent
        torch.cuda.synchronize(device)

    def _stream(self, stream):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return self._ct.c_void_p(s.cuda_stream)

    def mark_tail_ready(self, stream=None):
        if self.active:
            self._abi.check(self.lib.b200c_peer_event_record(self._event, self.device, self._stream(stream)))

    def pull(self, stream=None, wait: bool = True):
        """Enqueue: [wait for the neighbour's tail] -> copy it into buf[:K-1].  Rank 0 does nothing."""
        if not self.active or self.peer_src is None:
            return
        s = self._stream(stream)
        if wait:
            self._abi.check(self.lib.b200c_peer_event_wait(self._peer_event, self.device, s))
        self._abi.check(self.lib.b200c_halo_exchange(self._ct.c_void_p(self.dst), self.peer_src, self.nbytes, self.device, s))

    def close(self):
        if self._mapping is not None:
            self.lib.b200c_peer_close(self._mapping)
            self._mapping = None
        if self._peer_event is not None:
            self.lib.b200c_peer_event_destroy(self._peer_event, self.device)
            self._peer_event = None
        if self._event is not None:
            self.lib.b200c_peer_event_destroy(self._event, self.device)
            self._event = None
