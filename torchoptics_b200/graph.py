"""CUDA-graph capture of one lens evaluation: host prescription in, RMS spot size
and its gradients out, as ONE graph launch.

The reference evaluates a lens with ``trace_rays`` + ``compute_rms2d`` +
``.backward()`` -- about 2 200 eager kernel launches for a 7-surface lens
(SURVEY.md section 6).  Here the whole evaluation

    pinned host (c, t, nd, v)  --H2D-->  index model, paraxial pupil position, ray set
    (torch, ~60 tiny kernels)  -->  fused trace + adjoint kernel  -->  finalize  -->
    chain rule back to (c, t, nd, v)  --D2H-->  pinned host (rms, gradients)

is captured once and replayed, so a step costs one launch plus the kernels' own
time: CUDA streams and graphs instead of a tracing compiler.  Shapes are static
(the lens structure, ray counts, fields and wavelengths are fixed at capture).
"""
from __future__ import annotations

import numpy as np
import torch

from .lens_modeling import Lens, Specs


class GraphedSpotStep:
    """``step(c, t, nd, v) -> (rms [B], {'c','t','nd','v': gradient of sum(rms)})``.

    ``tracer`` / ``specs`` / ``lens`` fix the structure and sizes; the prescription
    values themselves are inputs of every call (CPU tensors of the lens' 2-D padded
    shapes, or None to keep the previous value).  With ``shard=(rank, world)`` every
    rank traces its slice of the pupil and the graph contains the all-reduce (NCCL, or the
    peer-memory exchange kernel when ``group`` is a :class:`~torchoptics_b200.peer.PeerExchange`).

    ``penalty_rate`` (e.g. 0.2): the step evaluates ``compute_loss_out``'s loss
    ``rms + penalty_rate * penalty`` (optics_simulator_lite.py:430-450) -- the fused spot pass plus
    the fused penalty pass -- returns the gradients of that loss, and ``host_penalty`` holds the
    penalty of every lens after a call.
    """

    PARAMS = ('c', 't', 'nd', 'v')

    def __init__(self, tracer, specs, lens, shard=(0, 1), group=None, warmup=3, penalty_rate=None):
        dev = torch.device(tracer.default_device)
        if dev.type != 'cuda':
            raise ValueError('GraphedSpotStep needs a CUDA tracer')
        self.device = dev
        self.tracer, self.shard, self.group = tracer, shard, group
        self.penalty_rate = penalty_rate
        self.structure = lens.structure
        # ONE pinned staging buffer each way: the four prescription tensors travel in one H2D copy,
        # gradients + rms (+ penalty) come back in one D2H copy (a step is latency-bound: every
        # separate small copy costs a few microseconds of a 0.3 ms step)
        shape = tuple(lens.c.shape)
        n_par, n_lens = len(self.PARAMS), len(lens)
        per = int(lens.c.numel())
        self._host_in_buf = torch.empty((n_par * per,), dtype=torch.float32).pin_memory()
        self._dev_in_buf = torch.empty((n_par * per,), dtype=torch.float32, device=dev)
        self.host_in = {k: self._host_in_buf[i * per:(i + 1) * per].view(shape) for i, k in enumerate(self.PARAMS)}
        self.dev_in = {k: self._dev_in_buf[i * per:(i + 1) * per].view(shape) for i, k in enumerate(self.PARAMS)}
        for k in self.PARAMS:
            self.host_in[k].copy_(getattr(lens, k).detach().to('cpu', torch.float32))
        self._dev_in_buf.copy_(self._host_in_buf)
        self._host_out_buf = torch.zeros((n_par * per + 2 * n_lens,), dtype=torch.float32).pin_memory()
        self._dev_out_buf = torch.zeros((n_par * per + 2 * n_lens,), dtype=torch.float32, device=dev)
        self.specs = Specs(specs.structure, specs.epd.detach().to(dev), specs.hfov.detach().to(dev),
                           specs.vig_up.detach().to(dev), specs.vig_down.detach().to(dev),
                           specs.vig_x.detach().to(dev))
        self.host_out = {k: self._host_out_buf[i * per:(i + 1) * per].view(shape)
                         for i, k in enumerate(self.PARAMS)}
        self.host_rms = self._host_out_buf[n_par * per:n_par * per + n_lens]
        self.host_penalty = self._host_out_buf[n_par * per + n_lens:]
        self._host_rms_np = self.host_rms.numpy()            # (shares the pinned memory)
        self._out_slices = (per, n_par * per, n_lens)
        self.h2d_bytes = self._host_in_buf.numel() * 4
        self.d2h_bytes = self._host_out_buf.numel() * 4
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):      # warm-up outside capture (allocator, lazy init)
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    def _body(self):
        self._dev_in_buf.copy_(self._host_in_buf, non_blocking=True)
        per, grads_end, n_lens = self._out_slices
        if self.penalty_rate is None:
            # the bare kernel sequence writing straight into the packed result buffer: no autograd
            # node, no elementwise glue, no per-tensor copies
            buf = self._dev_out_buf
            out = {'rms': buf[grads_end:grads_end + n_lens]}
            for i, name in enumerate(('gc', 'gt', 'gnd', 'gv')):
                out[name] = buf[i * per:(i + 1) * per].view(self.dev_in['c'].shape)
            lens = Lens(self.structure, *(self.dev_in[k] for k in self.PARAMS))
            self.tracer.spot_rms_and_grads(self.specs, lens, shard=self.shard, group=self.group, out=out)
            self._host_out_buf.copy_(buf, non_blocking=True)
            return
        leaves = {k: self.dev_in[k].detach().requires_grad_(True) for k in self.PARAMS}
        lens = Lens(self.structure, leaves['c'], leaves['t'], leaves['nd'], leaves['v'])
        res = self.tracer.loss_unsup(self.specs, lens, penalty_rate=self.penalty_rate,
                                     shard=self.shard, group=self.group)
        rms, loss = res['rms'], res['loss_unsup']
        self._dev_out_buf[grads_end + n_lens:].copy_(res['penalty'].detach())
        grads = torch.autograd.grad(loss.sum(), [leaves[k] for k in self.PARAMS], allow_unused=True)
        self._dev_out_buf[grads_end:grads_end + n_lens].copy_(rms.detach())
        for i, g in enumerate(grads):
            if g is not None:                    # (an unused parameter keeps its zeros)
                self._dev_out_buf[i * per:(i + 1) * per].copy_(g.reshape(-1))
        self._host_out_buf.copy_(self._dev_out_buf, non_blocking=True)

    def __call__(self, c=None, t=None, nd=None, v=None):
        # (the step is synchronous and ~0.24 ms long: the host-side staging is not hidden behind anything, and four
        # separate copy_ calls are 12 us of dispatch -- one fused call is 7)
        given = [(self.host_in[k], val) for k, val in (('c', c), ('t', t), ('nd', nd), ('v', v)) if val is not None]
        if given:
            torch._foreach_copy_([d for d, _ in given], [s for _, s in given])
        self.graph.replay()
        torch.cuda.current_stream(self.device).synchronize()
        self.check_exchange()
        return self.host_rms, self.host_out

    def check_exchange(self):
        """A peer that did not show up within the exchange kernel's spin limit poisons the sums with
        NaN (csrc/peer_exchange.cuh); turn that into an exception instead of a silently wrong step."""
        from .peer import PeerExchange
        # (numpy on the pinned buffer: ~1 us where torch.isnan(...).any() is ~6 us of dispatch, every step)
        if isinstance(self.group, PeerExchange) and bool(np.isnan(self._host_rms_np).any()):
            status, epoch = self.group.status()
            if status != 0:
                from ._native import NativeLibraryError
                raise NativeLibraryError(f'peer-memory exchange timed out on rank {self.group.rank} '
                                         f'(epoch {epoch}): a peer is late or dead; results are poisoned')
