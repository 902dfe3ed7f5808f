"""Peer-memory exchange of the pupil-sharded spot pass (``tl_peer_*`` of the C ABI).

One process per GPU; ``torch.distributed`` is used ONCE, to hand every rank the CUDA IPC
handles of the other ranks' windows.  After that the per-step SUM all-reduce of the moment
sums is a single kernel that stores into the peers' memory over NVLink / NVSwitch and adds
the contributions in rank order -- no NCCL call on the data path, graph-capturable, and
bit-identical on every rank.

    ex = PeerExchange(capacity=moments.numel())          # collective: every rank calls it
    rms, _ = tracer.spot_rms(specs, lens, shard=(rank, world), group=ex)

``ops.reduce_moments`` dispatches on the type of ``group``: a :class:`PeerExchange` uses
the kernel, a ``ProcessGroup`` (or None) uses ``torch.distributed.all_reduce`` (NCCL on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat


class PeerExchange:
    """A window on every rank of ``group`` (ranks must be GPUs of one node)."""

    def __init__(self, capacity, group=None, device=None):
        import torch.distributed as dist
        lib = nat.load()
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity = int(capacity)
        self._comm = ctypes.c_void_p()
        self._group = group
        nbytes = lib.tl_peer_handle_bytes()
        handle = ctypes.create_string_buffer(nbytes)
        # Every rank walks the same collective sequence (all_gather, all_reduce) whether or not
        # its local steps succeed, so a rank that cannot create or map a window (no CUDA IPC in
        # this container, no peer access) makes ALL ranks raise together instead of deadlocking.
        failure = None
        with torch.cuda.device(self.device):
            rc = lib.tl_peer_create(self.rank, self.world, self.capacity, ctypes.byref(self._comm), handle)
            if rc != 0:
                failure = f'tl_peer_create: {lib.tl_last_error().decode(errors="replace")}'
                self._comm = ctypes.c_void_p()
            if self.world > 1:
                mine = torch.frombuffer(bytearray(handle.raw) + bytearray([0 if failure else 1]),
                                        dtype=torch.uint8).to(self.device)
                everyone = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(everyone, mine, group=group)
                blobs = [bytes(h.cpu().numpy().tobytes()) for h in everyone]
                if failure is None and all(b[-1] == 1 for b in blobs):
                    rc = lib.tl_peer_connect(self._comm, b''.join(b[:-1] for b in blobs))
                    if rc != 0:
                        failure = f'tl_peer_connect: {lib.tl_last_error().decode(errors="replace")}'
                elif failure is None:
                    failure = 'a peer could not create its window'
                ok = torch.tensor([0 if failure else 1], dtype=torch.int32, device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also: every window is mapped
                if int(ok.item()) == 0 and failure is None:
                    failure = 'a peer could not map the windows'
            if failure is not None:
                if self._comm:
                    lib.tl_peer_destroy(self._comm)
                    self._comm = ctypes.c_void_p()
                raise nat.NativeLibraryError(f'peer-memory exchange unavailable on rank {self.rank}: {failure}')
        self._group = group

    def all_reduce(self, data):
        """Sum of ``data`` (float64, contiguous, CUDA) over the ranks; returns a new tensor.

        Every ``all_reduce`` of one ``PeerExchange`` must be issued on ONE stream, in the same order on every
        rank: the kernel reads its epoch from the window at the start and bumps it at the end, so two calls in
        flight on different streams would share an epoch (and a slot parity).  A peer that does not show up within
        4 s makes the result NaN on this rank, now and in every later step (``status()`` then reports 1)."""
        if data.dtype != torch.float64 or not data.is_contiguous():
            raise ValueError('PeerExchange.all_reduce needs a contiguous float64 tensor')
        nat.require_cuda(data, 'data')
        if data.numel() > self.capacity:
            raise ValueError(f'{data.numel()} values exceed the window capacity {self.capacity}')
        out = torch.empty_like(data)
        with torch.cuda.device(data.device):
            nat.check(nat.load().tl_peer_allreduce_f64(self._comm, data.data_ptr(), out.data_ptr(),
                                                       data.numel(), nat.stream_ptr(data.device)),
                      'tl_peer_allreduce_f64')
        return out

    def status(self):
        """(status, epoch): status 0 = ok, 1 = a peer did not show up within the spin limit.
        Synchronises the device."""
        st, ep = ctypes.c_int32(), ctypes.c_uint32()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            nat.check(nat.load().tl_peer_status(self._comm, ctypes.byref(st), ctypes.byref(ep)),
                      'tl_peer_status')
        return st.value, ep.value

    def close(self):
        """Collective: no rank may free its window while a peer can still write into it."""
        if self._comm:
            import torch.distributed as dist
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                if self.world > 1 and dist.is_initialized():
                    dist.barrier(group=self._group)
                nat.load().tl_peer_destroy(self._comm)
            self._comm = ctypes.c_void_p()
