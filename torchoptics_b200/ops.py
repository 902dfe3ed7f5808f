"""torch.autograd front end of the CUDA hot path.

Three differentiable operators, each a thin shim that lays out pointers for the
C ABI (include/torchoptics_b200.h) and launches on the current CUDA stream:

* :func:`trace`     -- ``trace_skew``            (reference rtl:594-675)
* :func:`spot_rms_from_rays` -- ``compute_rms2d`` (rtl:678-702), every lens
* :func:`spot_rms`  -- both fused, forward and backward in ONE pass over the rays

rtl = /root/reference/torchlens/ray_tracing_lite.py.  No CPU path exists.
"""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _native as nat

_GRAD_PER_RAY = ('x', 'y', 'z', 'cx', 'cy')


class _Layout:
    """Shapes and flat device views of one trace_skew argument set."""

    def __init__(self, x, y, z, cx, cy, c, t, mu, mask, k=None, a=None, sd=None):
        for name, v in (('x', x), ('y', y), ('z', z), ('cx', cx), ('cy', cy), ('c', c), ('t', t),
                        ('mu', mu), ('mask', mask)):
            nat.require_cuda(v, name)
            if name != 'mask' and v.dtype != torch.float32:
                raise TypeError(f'{name} must be float32 (got {v.dtype}); the kernels compute in fp32')
        for name, v in (('x', x), ('y', y), ('z', z), ('cx', cx), ('cy', cy)):
            if v.dim() != 4:
                raise ValueError(f'{name} must be 4-D [B,F,P,W]-broadcastable, got {tuple(v.shape)}')
        for name, v in (('c', c), ('t', t), ('mu', mu), ('mask', mask)):
            if v.dim() != 5:
                raise ValueError(f'{name} must be 5-D [B,1,1,W|1,S], got {tuple(v.shape)}')
        S = t.shape[-1]
        if not (c.shape[-1] == S and mu.shape[-1] == S and mask.shape[-1] == S):
            raise ValueError('c, t, mu and mask must agree on the number of surfaces')
        full = torch.broadcast_shapes(x.shape, y.shape, z.shape, cx.shape, cy.shape, c.shape[:-1],
                                      t.shape[:-1], mu.shape[:-1], mask.shape[:-1])
        B, F, P, W = full
        for name, v in (('c', c), ('t', t), ('mask', mask)):
            if v.shape[1] != 1 or v.shape[2] != 1 or v.shape[3] != 1:
                raise ValueError(f'{name} must have shape [B|1,1,1,1,S], got {tuple(v.shape)}')
        if mu.shape[1] != 1 or mu.shape[2] != 1:
            raise ValueError(f'mu must have shape [B|1,1,1,W|1,S], got {tuple(mu.shape)}')
        self.B, self.F, self.P, self.W, self.S = B, F, P, W, S
        self.shape = (B, F, P, W)
        self.device = y.device
        # tiny per-lens tables, contiguous
        self.c2 = torch.broadcast_to(c.detach(), (B, 1, 1, 1, S)).reshape(B, S).contiguous()
        self.t2 = torch.broadcast_to(t.detach(), (B, 1, 1, 1, S)).reshape(B, S).contiguous()
        self.mu3 = torch.broadcast_to(mu.detach(), (B, 1, 1, W, S)).reshape(B, W, S).contiguous()
        self.live = torch.broadcast_to(mask, (B, 1, 1, 1, S)).reshape(B, S).to(torch.uint8).contiguous()
        self.rays = [v.detach() for v in (x, y, z, cx, cy)]
        # extension surfaces (no reference behaviour): conic k [B,1,1,1,S], asphere coefficients
        # a [B,1,1,1,S,7], clear semi-diameter sd [B,1,1,1,S]; any of them selects the general kernels
        self.general = k is not None or a is not None or sd is not None
        self.k2 = self.a3 = self.sd2 = None
        if self.general:
            if S > nat.MAX_SURFACES_GEN:
                raise ValueError(f'general-surface lenses support at most {nat.MAX_SURFACES_GEN} surfaces')
            dev = self.device
            def table(v, tail, fill):
                if v is None:
                    return torch.full((B, S) + tail, fill, dtype=torch.float32, device=dev)
                nat.require_cuda(v, 'asphere table')
                return torch.broadcast_to(v.detach().to(torch.float32), (B, 1, 1, 1, S) + tail) \
                    .reshape((B, S) + tail).contiguous()
            self.k2 = table(k, (), 0.0)
            self.a3 = table(a, (nat.N_ASPHERE_TERMS,), 0.0)
            self.sd2 = table(sd, (), float('inf'))

    def problem(self, allow_backward_rays, arith, p_begin=0, p_end=None):
        pb = nat.TlProblem()
        for name, v in zip(_GRAD_PER_RAY, self.rays):
            setattr(pb, name, nat.strided(v, self.shape))
        pb.c = self.c2.data_ptr()
        pb.t = self.t2.data_ptr()
        pb.mu = self.mu3.data_ptr()
        pb.live = self.live.data_ptr()
        pb.B, pb.F, pb.P, pb.W, pb.S = self.B, self.F, self.P, self.W, self.S
        pb.allow_backward_rays = int(bool(allow_backward_rays))
        pb.arith = int(arith)
        pb.p_begin = int(p_begin)
        pb.p_end = int(self.P if p_end is None else p_end)
        if self.general:
            pb.k, pb.a, pb.sd = self.k2.data_ptr(), self.a3.data_ptr(), self.sd2.data_ptr()
        return pb


def _ptr(t):
    return None if t is None else t.data_ptr()


class _TraceSkew(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays, arith, k=None, a=None,
                sd=None, aggregate=False):
        lay = _Layout(x, y, z, cx, cy, c, t, mu, mask, k, a, sd)
        ctx.set_materialize_grads(False)      # unused outputs arrive as None, not as zero tensors
        ctx.aggregate = bool(aggregate)
        if aggregate and lay.general:
            raise ValueError('aggregate=True (the penalty stacks of rtl:641-657) exists for spherical '
                             'lenses only, not with the extension tables k / a / sd')
        if aggregate and lay.S > nat.MAX_SURFACES_BWD and any(ctx.needs_input_grad[:8]):
            raise ValueError(f'differentiable aggregate=True supports at most {nat.MAX_SURFACES_BWD} surfaces')
        if lay.general:
            ctx.save_for_backward(x, y, z, cx, cy, c, t, mu, mask,
                                  *(v for v in (k, a, sd) if v is not None))
            ctx.flags = (allow_backward_rays, arith)
            ctx.ext = (k is not None, a is not None, sd is not None)
            return _TraceSkew._forward_general(ctx, lay, allow_backward_rays, arith)
        if lay.S > nat.MAX_SURFACES_FWD:
            raise ValueError(f'at most {nat.MAX_SURFACES_FWD} surfaces are supported')
        lib = nat.load()
        with nat.on_device(lay.device):
            outs = [torch.empty(lay.shape, dtype=torch.float32, device=lay.device) for _ in range(4)]
            ok = torch.empty(lay.shape, dtype=torch.bool, device=lay.device)
            backward = torch.empty(lay.shape, dtype=torch.bool, device=lay.device)
            pb = lay.problem(allow_backward_rays, arith)
            stacks = []
            if aggregate:      # z_RELU, theta_norm, theta_prime_norm of every surface: [S,B,F,P,W]
                stacks = [torch.empty((lay.S,) + lay.shape, dtype=torch.float32, device=lay.device)
                          for _ in range(3)]
            out = nat.TlTraceOut(*[o.data_ptr() for o in outs], ok.data_ptr(), backward.data_ptr(), None,
                                 *[v.data_ptr() for v in stacks])
            nat.check(lib.tl_trace_fwd(ctypes.byref(pb), ctypes.byref(out), nat.stream_ptr(lay.device)),
                      'tl_trace_fwd')
        ctx.save_for_backward(x, y, z, cx, cy, c, t, mu, mask)
        ctx.flags = (allow_backward_rays, arith)
        ctx.mark_non_differentiable(ok, backward)
        return (*outs, ok, backward, *stacks)

    @staticmethod
    def _forward_general(ctx, lay, allow_backward_rays, arith):
        """Forward of a lens with extension surfaces: also returns the optical path length,
        differentiable like x, y, cx, cy (TlSeeds.gopl)."""
        lib = nat.load()
        with nat.on_device(lay.device):
            outs = [torch.empty(lay.shape, dtype=torch.float32, device=lay.device) for _ in range(4)]
            ok = torch.empty(lay.shape, dtype=torch.bool, device=lay.device)
            backward = torch.empty(lay.shape, dtype=torch.bool, device=lay.device)
            opl = torch.empty(lay.shape, dtype=torch.float32, device=lay.device)
            pb = lay.problem(allow_backward_rays, arith)
            out = nat.TlTraceOut(*[o.data_ptr() for o in outs], ok.data_ptr(), backward.data_ptr(),
                                 opl.data_ptr())
            nat.check(lib.tl_trace_fwd(ctypes.byref(pb), ctypes.byref(out), nat.stream_ptr(lay.device)),
                      'tl_trace_fwd')
        ctx.general = True
        ctx.mark_non_differentiable(ok, backward)
        return (*outs, ok, backward, opl)

    @staticmethod
    def backward(ctx, gx, gy, gcx, gcy, _gok, _gbw, *extra_grads):
        saved = ctx.saved_tensors
        stack_grads = extra_grads if ctx.aggregate else (None, None, None)
        k = a = sd = None
        g_opl = None
        if getattr(ctx, 'general', False):
            g_opl = extra_grads[0] if extra_grads else None
            extra = list(saved[9:])
            has_k, has_a, has_sd = ctx.ext
            k = extra.pop(0) if has_k else None
            a = extra.pop(0) if has_a else None
            sd = extra.pop(0) if has_sd else None
            saved = saved[:9]
        x, y, z, cx, cy, c, t, mu, mask = saved
        allow_backward_rays, arith = ctx.flags
        lay = _Layout(x, y, z, cx, cy, c, t, mu, mask, k, a, sd)
        if lay.S > (nat.MAX_SURFACES_GEN if lay.general else nat.MAX_SURFACES_BWD):
            raise ValueError(f'backward supports at most {nat.MAX_SURFACES_BWD} surfaces')
        lib = nat.load()
        dev = lay.device
        need = ctx.needs_input_grad
        with nat.on_device(dev):
            seeds = [None if g is None else g.to(torch.float32).expand(lay.shape).contiguous()
                     for g in (gx, gy, gcx, gcy, g_opl)]
            seed_opl = seeds.pop()
            gc = torch.empty((lay.B, lay.S), dtype=torch.float32, device=dev)
            gt = torch.empty_like(gc)
            gmu = torch.empty((lay.B, lay.W, lay.S), dtype=torch.float32, device=dev)
            gz_sum = torch.empty((lay.B,), dtype=torch.float32, device=dev)
            z_per_lens = z.numel() == z.shape[0]      # [B|1,1,1,1]
            per_ray = {}
            for i, name in enumerate(_GRAD_PER_RAY):
                if need[i] and not (name == 'z' and z_per_lens):
                    per_ray[name] = torch.empty(lay.shape, dtype=torch.float32, device=dev)
            pb = lay.problem(allow_backward_rays, arith)
            stack_seeds = [None if g is None else
                           g.to(torch.float32).expand((lay.S,) + lay.shape).contiguous() for g in stack_grads]
            sd = nat.TlSeeds(*[_ptr(s) for s in seeds], *[_ptr(s) for s in stack_seeds], _ptr(seed_opl))
            gk = ga = None
            if lay.general:
                gk = torch.empty_like(gc)
                ga = torch.empty((lay.B, lay.S, nat.N_ASPHERE_TERMS), dtype=torch.float32, device=dev)
            gr = nat.TlGrads(gc.data_ptr(), gt.data_ptr(), gmu.data_ptr(), gz_sum.data_ptr(),
                             *[_ptr(per_ray.get(n)) for n in _GRAD_PER_RAY], _ptr(gk), _ptr(ga))
            ws_bytes = lib.tl_trace_bwd_workspace(ctypes.byref(pb))
            if ws_bytes == 0:
                nat.check(-1, 'tl_trace_bwd_workspace')
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            nat.check(lib.tl_trace_bwd(ctypes.byref(pb), ctypes.byref(sd), ctypes.byref(gr),
                                       ws.data_ptr(), ws_bytes, nat.stream_ptr(dev)), 'tl_trace_bwd')
        grads = []
        for i, (name, v) in enumerate(zip(_GRAD_PER_RAY, (x, y, z, cx, cy))):
            if not need[i]:
                grads.append(None)
            elif name == 'z' and z_per_lens:
                g = gz_sum.reshape(lay.B, 1, 1, 1)
                grads.append(g.sum_to_size(z.shape))
            else:
                grads.append(per_ray[name].sum_to_size(v.shape))
        grads.append(gc.reshape(lay.B, 1, 1, 1, lay.S).sum_to_size(c.shape) if need[5] else None)
        grads.append(gt.reshape(lay.B, 1, 1, 1, lay.S).sum_to_size(t.shape) if need[6] else None)
        grads.append(gmu.reshape(lay.B, 1, 1, lay.W, lay.S).sum_to_size(mu.shape) if need[7] else None)
        g_k = g_a = None
        if lay.general and k is not None and need[11]:
            g_k = gk.reshape(lay.B, 1, 1, 1, lay.S).sum_to_size(k.shape)
        if lay.general and a is not None and need[12]:
            g_a = ga.reshape(lay.B, 1, 1, 1, lay.S, -1).sum_to_size(a.shape)
        return (*grads, None, None, None, g_k, g_a, None, None)


STACK_KEYS = ('z_RELU', 'theta_norm', 'theta_prime_norm')      # rtl:598


def trace(x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
          k=None, a=None, sd=None, aggregate=False):
    """CUDA ``trace_skew``: returns (x, y, cx, cy, ray_ok, ray_backward), all [B,F,P,W]; with any of
    the extension tables k / a / sd also the optical path length as a 7th output; with
    ``aggregate=True`` (rtl:641-657) a 7th output ``stacks``: a dict of three S-long lists of
    [B,F,P,W] tensors (views of one [S,B,F,P,W] tensor per key), differentiable like the rest."""
    full = torch.broadcast_shapes(x.shape, y.shape, z.shape, cx.shape, cy.shape, c.shape[:-1],
                                  t.shape[:-1], mu.shape[:-1], mask.shape[:-1])
    if 0 in full:       # empty ray set: nothing to launch, same (empty) results as the reference
        for name, v in (('x', x), ('y', y)):
            nat.require_cuda(v, name)
        zero = (x.sum() + y.sum() + z.sum() + cx.sum() + cy.sum() + c.sum() + t.sum() + mu.sum()) * 0
        outs = [zero.expand(full).clone() for _ in range(4)]
        flags = [torch.zeros(full, dtype=torch.bool, device=y.device) for _ in range(2)]
        flags[0] = ~flags[0]
        if aggregate:
            S = t.shape[-1]
            return (*outs, *flags, {key: [zero.expand(full).clone() for _ in range(S)] for key in STACK_KEYS})
        return (*outs, *flags)
    res = _TraceSkew.apply(x, y, z, cx, cy, c, t, mu, mask, bool(allow_backward_rays), int(arith),
                           k, a, sd, bool(aggregate))
    # provenance: lets rms_of_trace() recognise un-modified trace outputs and evaluate the RMS (and
    # its backward) with the fused pass on the trace's INPUTS instead of reading [B,F,P,W] tensors back
    tensors = [v for v in (x, y, z, cx, cy, c, t, mu, mask, k, a, sd) if v is not None]
    res[1]._tl_prov = {'kind': 'rays', 'args': (x, y, z, cx, cy, c, t, mu, mask),
                       'ext': {'k': k, 'a': a, 'sd': sd}, 'allow': bool(allow_backward_rays), 'arith': int(arith),
                       'tensors': tensors, 'versions': [v._version for v in tensors],
                       'ok': weakref.ref(res[4]), 'ok_version': res[4]._version, 'y_version': res[1]._version}
    if aggregate:
        return (*res[:6], {key: list(res[6 + i].unbind(0)) for i, key in enumerate(STACK_KEYS)})
    return res


class _RmsFromRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, ray_ok):
        nat.require_cuda(y, 'y')
        nat.require_cuda(ray_ok, 'ray_ok')
        if y.dim() != 4 or y.dtype != torch.float32:
            raise ValueError('y must be a float32 [B,F,P,W] tensor')
        B, F, P, W = y.shape
        lib = nat.load()
        dev = y.device
        with nat.on_device(dev):
            yc = y.detach().contiguous()
            okc = torch.broadcast_to(ray_ok, y.shape).to(torch.uint8).contiguous()
            rms = torch.empty((B,), dtype=torch.float32, device=dev)
            rms_field = torch.empty((B, F), dtype=torch.float32, device=dev)
            stats = torch.empty((B, F, 4), dtype=torch.float64, device=dev)
            ws_bytes = lib.tl_rms_workspace(B, F, P, W)
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            nat.check(lib.tl_rms_fwd(yc.data_ptr(), okc.data_ptr(), B, F, P, W, rms.data_ptr(),
                                     rms_field.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws_bytes,
                                     nat.stream_ptr(dev)), 'tl_rms_fwd')
        ctx.save_for_backward(yc, okc, stats)
        ctx.mark_non_differentiable(rms_field)
        return rms, rms_field

    @staticmethod
    def backward(ctx, grad_rms, _grad_field):
        yc, okc, stats = ctx.saved_tensors
        B, F, P, W = yc.shape
        lib = nat.load()
        dev = yc.device
        with nat.on_device(dev):
            g = grad_rms.to(torch.float32).contiguous()
            gy = torch.empty_like(yc)
            nat.check(lib.tl_rms_bwd(yc.data_ptr(), okc.data_ptr(), stats.data_ptr(), g.data_ptr(),
                                     B, F, P, W, gy.data_ptr(), nat.stream_ptr(dev)), 'tl_rms_bwd')
        return gy, None


def spot_rms_from_rays(y, ray_ok):
    """Per-lens mean-over-fields y-RMS spot size of already traced rays:
    returns (rms [B], rms_field [B,F])."""
    return _RmsFromRays.apply(y, ray_ok)


def _provenance_of(y, ray_ok):
    """The provenance record of `y` if (y, ray_ok) are the untouched outputs of one trace call whose
    inputs have not been modified in place since; else None."""
    prov = getattr(y, '_tl_prov', None)
    if prov is None or prov['ok']() is not ray_ok:
        return None
    if y._version != prov['y_version'] or ray_ok._version != prov['ok_version']:
        return None
    if any(t._version != v for t, v in zip(prov['tensors'], prov['versions'])):
        return None
    return prov


def rms_of_trace(y, ray_ok):
    """``compute_rms2d`` for every lens, [B] (rtl:678-702).  When (y, ray_ok) are the unmodified outputs
    of :func:`trace` (or ``RayTracer.trace_rays``) the value AND its gradient come from the fused
    spot pass evaluated on that trace's inputs -- one pass over the rays, nothing per-ray read back or
    written in backward -- which is what makes the reference's own call sequence ``trace_rays ->
    compute_rms2d -> backward`` fast without changing it.  Anything else (a modified, sliced or
    foreign y) takes the reduction over the materialised tensors."""
    prov = _provenance_of(y, ray_ok)
    if prov is not None:
        if prov['kind'] == 'lens':
            fused = prov['fused']()
            if fused is not None:
                return fused
        else:
            x, yy, z, cx, cy, c, t, mu, mask = prov['args']
            ext = prov['ext']
            general = any(v is not None for v in ext.values())
            wants = torch.is_grad_enabled() and any(v is not None and v.requires_grad
                                                    for v in (z, c, t, mu, ext['k'], ext['a']))
            limit = nat.MAX_SURFACES_GEN if general else (nat.MAX_SURFACES_SPOT if wants else nat.MAX_SURFACES_FWD)
            per_ray_grad = torch.is_grad_enabled() and any(v.requires_grad for v in (x, yy, cx, cy))
            if z.numel() == z.shape[0] and t.shape[-1] <= limit and not per_ray_grad:
                return spot_rms(x, yy, z, cx, cy, c, t, mu, mask, prov['allow'], prov['arith'], **ext)[0]
    return spot_rms_from_rays(y, ray_ok)[0]


def pupil_slice(n_pupil, rank, world):
    """Contiguous slice [begin, end) of the pupil axis traced by ``rank`` of ``world``."""
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world of {world}')
    return (n_pupil * rank) // world, (n_pupil * (rank + 1)) // world


def reduce_moments(moments, group=None):
    """The one data-path collective of the sharded spot pass: per-(lens, field,
    wavelength) sums are additive over pupil slices -> SUM all-reduce (fp64,
    a few KB).  ``group`` is a :class:`~torchoptics_b200.peer.PeerExchange` (one kernel storing
    into the peers' memory over NVLink, rank-ordered sum) or a torch process group / None
    (``all_reduce``: NCCL on GPUs, gloo in the CPU tests).  Returns the reduced tensor."""
    from .peer import PeerExchange
    if isinstance(group, PeerExchange):
        return group.all_reduce(moments)
    torch.distributed.all_reduce(moments, op=torch.distributed.ReduceOp.SUM, group=group)
    return moments


def _accumulate(lay, allow_backward_rays, arith, want_grad, p_begin, p_end):
    lib = nat.load()
    dev = lay.device
    count = lib.tl_spot_moment_count_general if lay.general else lib.tl_spot_moment_count
    n_acc = count(lay.S, int(want_grad))
    moments = torch.empty((lay.B, lay.F, lay.W, n_acc), dtype=torch.float64, device=dev)
    ref_y = torch.empty((lay.B, lay.F), dtype=torch.float32, device=dev)
    pb = lay.problem(allow_backward_rays, arith, p_begin, p_end)
    ws_bytes = lib.tl_spot_workspace(ctypes.byref(pb), int(want_grad))
    if ws_bytes == 0:
        nat.check(-1, 'tl_spot_workspace')
    ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
    nat.check(lib.tl_spot_accumulate(ctypes.byref(pb), int(want_grad), moments.data_ptr(),
                                     ref_y.data_ptr(), ws.data_ptr(), ws_bytes, nat.stream_ptr(dev)),
              'tl_spot_accumulate')
    return moments, ref_y


def spot_moments(x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
                 want_grad=True, shard=(0, 1), k=None, a=None, sd=None):
    """Raw additive sums of one pupil slice (no autograd): (moments [B,F,W,n], ref_y [B,F])."""
    lay = _Layout(x, y, z, cx, cy, c, t, mu, mask, k, a, sd)
    p_begin, p_end = pupil_slice(lay.P, *shard)
    with nat.on_device(lay.device):
        return _accumulate(lay, allow_backward_rays, arith, want_grad, p_begin, p_end)


def spot_kernel_name(x, y, z, cx, cy, c, t, mu, mask, want_grad=True, shard=(0, 1), k=None, a=None, sd=None):
    """Which kernel the fused spot pass launches for this problem (diagnostics)."""
    lay = _Layout(x, y, z, cx, cy, c, t, mu, mask, k, a, sd)
    p_begin, p_end = pupil_slice(lay.P, *shard)
    pb = lay.problem(True, nat.ARITH_GUARDED, p_begin, p_end)
    return nat.load().tl_spot_kernel_name(ctypes.byref(pb), int(want_grad)).decode()


def spot_kernel_runner(x, y, z, cx, cy, c, t, mu, mask, shard=(0, 1)):
    """A closure that launches the dominant kernel of the fused pass ALONE on the current stream
    (tl_spot_kernel_only; bench.py times it with CUDA events for its roofline line)."""
    lay = _Layout(x, y, z, cx, cy, c, t, mu, mask)
    p_begin, p_end = pupil_slice(lay.P, *shard)
    lib = nat.load()
    with nat.on_device(lay.device):
        _, ref_y = _accumulate(lay, True, nat.ARITH_GUARDED, True, p_begin, p_end)
        pb = lay.problem(True, nat.ARITH_GUARDED, p_begin, p_end)
        ws_bytes = lib.tl_spot_workspace(ctypes.byref(pb), 1)
        ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=lay.device)

    def run():
        nat.check(lib.tl_spot_kernel_only(ctypes.byref(pb), ref_y.data_ptr(), ws.data_ptr(), ws_bytes,
                                          nat.stream_ptr(lay.device)), 'tl_spot_kernel_only')
    run.keep = (lay, ref_y, ws)
    return run


class _SpotRms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays, arith, shard, group,
                k, a, sd, grad_on):
        if any(ctx.needs_input_grad[i] for i in (0, 1, 3, 4)):
            raise ValueError('the fused spot pass differentiates w.r.t. z, c, t, mu (and k, a) only; '
                             'use trace() + spot_rms_from_rays() for gradients of x, y, cx, cy')
        if ctx.needs_input_grad[15]:
            raise ValueError('the clear semi-diameter sd is not differentiable')
        lay = _Layout(x, y, z, cx, cy, c, t, mu, mask, k, a, sd)
        if z.numel() != z.shape[0]:
            raise ValueError('the fused spot pass needs a per-lens pupil position z of shape [B,1,1,1]')
        # (grad mode is sampled by the caller: inside forward() autograd is always off)
        want_grad = grad_on and any(ctx.needs_input_grad[i] for i in (2, 5, 6, 7, 13, 14))
        if lay.general:
            limit = nat.MAX_SURFACES_GEN
        else:
            limit = nat.MAX_SURFACES_SPOT if want_grad else nat.MAX_SURFACES_FWD
        if lay.S > limit:
            raise ValueError(f'the fused spot pass supports at most {limit} surfaces '
                             f'({"with" if want_grad else "without"} gradients)')
        lib = nat.load()
        dev = lay.device
        rank, world = shard
        p_begin, p_end = pupil_slice(lay.P, rank, world)
        if p_end <= p_begin:
            raise ValueError(f'pupil axis ({lay.P}) is too short to shard over {world} ranks')
        with nat.on_device(dev):
            moments, ref_y = _accumulate(lay, allow_backward_rays, arith, want_grad, p_begin, p_end)
            stream = nat.stream_ptr(dev)
            if world > 1:
                moments = reduce_moments(moments, group)
            rms = torch.empty((lay.B,), dtype=torch.float32, device=dev)
            rms_field = torch.empty((lay.B, lay.F), dtype=torch.float32, device=dev)
            saved = []
            if want_grad:
                gc = torch.empty((lay.B, lay.S), dtype=torch.float32, device=dev)
                gt = torch.empty_like(gc)
                gmu = torch.empty((lay.B, lay.W, lay.S), dtype=torch.float32, device=dev)
                gz = torch.empty((lay.B,), dtype=torch.float32, device=dev)
                saved = [gc, gt, gmu, gz]
                extra = [None, None]
                if lay.general:
                    gk = torch.empty_like(gc)
                    ga = torch.empty((lay.B, lay.S, nat.N_ASPHERE_TERMS), dtype=torch.float32, device=dev)
                    saved += [gk, ga]
                    extra = [gk.data_ptr(), ga.data_ptr()]
                out = nat.TlSpotOut(rms.data_ptr(), rms_field.data_ptr(), gc.data_ptr(), gt.data_ptr(),
                                    gmu.data_ptr(), gz.data_ptr(), *extra)
            else:
                out = nat.TlSpotOut(rms.data_ptr(), rms_field.data_ptr(), None, None, None, None, None, None)
            nat.check(lib.tl_spot_finalize(moments.data_ptr(), ref_y.data_ptr(), lay.B, lay.F, lay.W,
                                           lay.S, lay.P, int(want_grad), ctypes.byref(out), stream),
                      'tl_spot_finalize')
        if want_grad:
            ctx.save_for_backward(*saved)
        ctx.meta = (lay.B, lay.W, lay.S, z.shape, c.shape, t.shape, mu.shape,
                    None if k is None else k.shape, None if a is None else a.shape)
        ctx.mark_non_differentiable(rms_field)
        return rms, rms_field

    @staticmethod
    def backward(ctx, grad_rms, _grad_field):
        saved = ctx.saved_tensors
        if len(saved) < 4:      # forward ran without gradients (nothing it differentiates required one)
            return (None,) * 17
        gc, gt, gmu, gz = saved[:4]
        B, W, S, z_shape, c_shape, t_shape, mu_shape, k_shape, a_shape = ctx.meta
        need = ctx.needs_input_grad
        g = grad_rms.to(torch.float32).reshape(B)
        out = [None] * 17
        if need[2]:
            out[2] = (gz * g).reshape(B, 1, 1, 1).sum_to_size(z_shape)
        if need[5]:
            out[5] = (gc * g[:, None]).reshape(B, 1, 1, 1, S).sum_to_size(c_shape)
        if need[6]:
            out[6] = (gt * g[:, None]).reshape(B, 1, 1, 1, S).sum_to_size(t_shape)
        if need[7]:
            out[7] = (gmu * g[:, None, None]).reshape(B, 1, 1, W, S).sum_to_size(mu_shape)
        if need[13] and len(saved) > 4:
            out[13] = (saved[4] * g[:, None]).reshape(B, 1, 1, 1, S).sum_to_size(k_shape)
        if need[14] and len(saved) > 4:
            out[14] = (saved[5] * g[:, None, None]).reshape(B, 1, 1, 1, S, -1).sum_to_size(a_shape)
        return tuple(out)


def spot_rms(x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
             shard=(0, 1), group=None, k=None, a=None, sd=None):
    """Fused ``trace_skew`` -> ``compute_rms2d`` (and, when any of z, c, t, mu
    requires grad, its backward) in one pass over the rays.

    ``shard=(rank, world)`` traces only this rank's slice of the pupil axis and
    all-reduces the per-field sums over ``group``; every rank returns the same
    (rms [B], rms_field [B,F]) and, after ``.backward()``, the same gradients.
    """
    return _SpotRms.apply(x, y, z, cx, cy, c, t, mu, mask, bool(allow_backward_rays), int(arith),
                          (int(shard[0]), int(shard[1])), group, k, a, sd, torch.is_grad_enabled())


# ---------------------------------------------------------------------------
# Lens-level fused pass: ray-set staging kernels + spot pass + chain rule
# ---------------------------------------------------------------------------
class _PenaltySum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays, arith, shard, group, scale):
        if any(ctx.needs_input_grad[i] for i in (0, 1, 3, 4)):
            raise ValueError('the fused penalty pass differentiates w.r.t. z, c, t, mu only; use '
                             'trace(..., aggregate=True) for gradients of x, y, cx, cy')
        lay = _Layout(x, y, z, cx, cy, c, t, mu, mask)
        if z.numel() != z.shape[0]:
            raise ValueError('the fused penalty pass needs a per-lens pupil position z of shape [B,1,1,1]')
        if lay.S > nat.MAX_SURFACES_BWD:
            raise ValueError(f'the fused penalty pass supports at most {nat.MAX_SURFACES_BWD} surfaces')
        lib = nat.load()
        dev = lay.device
        rank, world = shard
        p_begin, p_end = pupil_slice(lay.P, rank, world)
        if p_end <= p_begin:
            raise ValueError(f'pupil axis ({lay.P}) is too short to shard over {world} ranks')
        with nat.on_device(dev):
            n_acc = lib.tl_penalty_moment_count(lay.S)
            moments = torch.empty((lay.B, lay.F, lay.W, n_acc), dtype=torch.float64, device=dev)
            pb = lay.problem(allow_backward_rays, arith, p_begin, p_end)
            ws_bytes = lib.tl_penalty_workspace(ctypes.byref(pb))
            if ws_bytes == 0:
                nat.check(-1, 'tl_penalty_workspace')
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            stream = nat.stream_ptr(dev)
            nat.check(lib.tl_penalty_accumulate(ctypes.byref(pb), moments.data_ptr(), ws.data_ptr(), ws_bytes,
                                                stream), 'tl_penalty_accumulate')
            if world > 1:
                moments = reduce_moments(moments, group)
            penalty = torch.empty((lay.B,), dtype=torch.float32, device=dev)
            gc = torch.empty((lay.B, lay.S), dtype=torch.float32, device=dev)
            gt = torch.empty_like(gc)
            gmu = torch.empty((lay.B, lay.W, lay.S), dtype=torch.float32, device=dev)
            gz = torch.empty((lay.B,), dtype=torch.float32, device=dev)
            out = nat.TlPenaltyOut(penalty.data_ptr(), gc.data_ptr(), gt.data_ptr(), gmu.data_ptr(),
                                   gz.data_ptr())
            nat.check(lib.tl_penalty_finalize(moments.data_ptr(), lay.B, lay.F, lay.W, lay.S, float(scale),
                                              ctypes.byref(out), stream), 'tl_penalty_finalize')
        ctx.save_for_backward(gc, gt, gmu, gz)
        ctx.meta = (lay.B, lay.W, lay.S, z.shape, c.shape, t.shape, mu.shape)
        return penalty

    @staticmethod
    def backward(ctx, grad_penalty):
        gc, gt, gmu, gz = ctx.saved_tensors
        B, W, S, z_shape, c_shape, t_shape, mu_shape = ctx.meta
        need = ctx.needs_input_grad
        g = grad_penalty.to(torch.float32).reshape(B)
        out = [None] * 14
        if need[2]:
            out[2] = (gz * g).reshape(B, 1, 1, 1).sum_to_size(z_shape)
        if need[5]:
            out[5] = (gc * g[:, None]).reshape(B, 1, 1, 1, S).sum_to_size(c_shape)
        if need[6]:
            out[6] = (gt * g[:, None]).reshape(B, 1, 1, 1, S).sum_to_size(t_shape)
        if need[7]:
            out[7] = (gmu * g[:, None, None]).reshape(B, 1, 1, W, S).sum_to_size(mu_shape)
        return tuple(out)


def penalty_sum(x, y, z, cx, cy, c, t, mu, mask, n_seq, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
                shard=(0, 1), group=None):
    """The penalty of ``compute_loss_out`` (optics_simulator_lite.py:430-450) for every lens, as one
    fused pass: ``sum over rays of Q``, ``Q = (sum_k theta_norm + sum_k theta_prime_norm +
    sum_k z_RELU) / n_seq`` with the per-surface terms of ``trace_skew(aggregate=True)``
    (rtl:641-657).  Returns a [B] tensor, differentiable w.r.t. z, c, t, mu; no stack is
    materialised.  ``shard`` / ``group`` as in :func:`spot_rms`."""
    return _PenaltySum.apply(x, y, z, cx, cy, c, t, mu, mask, bool(allow_backward_rays), int(arith),
                             (int(shard[0]), int(shard[1])), group, 1.0 / float(n_seq))


class LensTables:
    """Device-side constants of a (structure, specs, tracer) triple that the staging
    kernels read: masks, stop indices, fields, wavelengths.  Built once."""

    def __init__(self, structure, rel_fields, wavelengths, device):
        dev = torch.device(device)
        self.mask = structure.mask_torch.to(dev).to(torch.uint8).contiguous()
        self.mask_g = structure.mask_G_torch.to(dev).to(torch.uint8).contiguous()
        self.stop_idx = torch.as_tensor(structure.stop_idx, dtype=torch.int32).to(dev).contiguous()
        self.rel_fields = torch.tensor([float(f) for f in rel_fields], dtype=torch.float32, device=dev)
        self.wavelengths = torch.tensor([float(w) for w in wavelengths], dtype=torch.float32, device=dev)
        self.zero = torch.zeros((1, 1, 1, 1), dtype=torch.float32, device=dev)
        self.B, self.L = self.mask.shape
        self.F, self.W = self.rel_fields.numel(), self.wavelengths.numel()

    def lens_struct(self, c, t, nd, v, hfov, epd):
        ln = nat.TlLens()
        ln.c, ln.t, ln.nd, ln.v = c.data_ptr(), t.data_ptr(), nd.data_ptr(), v.data_ptr()
        ln.mask, ln.mask_g, ln.stop_idx = self.mask.data_ptr(), self.mask_g.data_ptr(), self.stop_idx.data_ptr()
        ln.hfov, ln.epd = hfov.data_ptr(), epd.data_ptr()
        ln.rel_fields, ln.wavelengths = self.rel_fields.data_ptr(), self.wavelengths.data_ptr()
        ln.B, ln.L, ln.F, ln.W = self.B, self.L, self.F, self.W
        return ln


def aim_table(c, t, nd, v, hfov, epd, tables, allow_backward_rays=True, vig=None, aim_mode=nat.AIM_REAL):
    """Ray aiming on the device (``RayTracer.ray_aiming`` rtl:129-208: one iteration, 'real' stop
    radius, no pupil vignetting function): returns ``aim [B,F,W,3] = (x_gain, y_gain, y_shift)`` of
    the affine pupil map ``x_rel * x_gain, y_rel * y_gain + y_shift`` (rtl:196-206).  Two kernel
    launches (tl_stage_fwd, tl_aim); no gradient, like the reference's (rtl:111 detaches the lens)."""
    for name, val in (('c', c), ('t', t), ('nd', nd), ('v', v), ('hfov', hfov), ('epd', epd)):
        nat.require_cuda(val, name)
        if val.dtype != torch.float32:
            raise TypeError(f'{name} must be float32')
    lib = nat.load()
    dev = c.device
    B, L, F, W = tables.B, tables.L, tables.F, tables.W
    if tuple(c.shape) != (B, L):
        raise ValueError(f'lens tensors must be [B={B}, L={L}], got {tuple(c.shape)}')
    if L > 64:
        raise ValueError('too many surfaces for the staging kernels')
    with nat.on_device(dev), torch.no_grad():
        cc, tt, ndd, vv = (a.detach().contiguous() for a in (c, t, nd, v))
        hf, ep = hfov.detach().contiguous(), epd.detach().contiguous()
        mu = torch.empty((B, W, L), dtype=torch.float32, device=dev)
        z = torch.empty((B,), dtype=torch.float32, device=dev)
        cy = torch.empty((B, F), dtype=torch.float32, device=dev)
        half_epd = torch.empty((B,), dtype=torch.float32, device=dev)
        aim = torch.empty((B, F, W, 3), dtype=torch.float32, device=dev)
        ln = tables.lens_struct(cc, tt, ndd, vv, hf, ep)
        stream = nat.stream_ptr(dev)
        nat.check(lib.tl_stage_fwd(ctypes.byref(ln), mu.data_ptr(), z.data_ptr(), cy.data_ptr(),
                                   half_epd.data_ptr(), stream), 'tl_stage_fwd')
        vig_c = None if vig is None else vig.detach().to(torch.float32).contiguous()
        nat.check(lib.tl_aim(ctypes.byref(ln), mu.data_ptr(), z.data_ptr(), cy.data_ptr(), half_epd.data_ptr(),
                             _ptr(vig_c), int(aim_mode), int(bool(allow_backward_rays)), aim.data_ptr(), stream),
                  'tl_aim')
    return aim


class _Staged:
    """Ray set of a lens batch built by the staging kernel (tl_stage_fwd, optionally tl_aim): the
    index ratios mu [B,W,L], pupil position z [B], field cosines cy [B,F], half EPD [B] and the
    TlProblem that points at them (relative pupil grid + xy_scale / aim applied on load).  Holds
    every buffer alive for as long as the problem is in use."""

    def __init__(self, c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed,
                 p_begin=0, p_end=None, max_surfaces=64, want_ref=False, vig=None, aim_mode=nat.AIM_REAL):
        for name, val in (('c', c), ('t', t), ('nd', nd), ('v', v), ('hfov', hfov), ('epd', epd),
                          ('x', x_rel), ('y', y_rel)):
            nat.require_cuda(val, name)
            if val.dtype != torch.float32:
                raise TypeError(f'{name} must be float32')
        lib = nat.load()
        dev = c.device
        B, L, F, W = tables.B, tables.L, tables.F, tables.W
        if tuple(c.shape) != (B, L):
            raise ValueError(f'lens tensors must be [B={B}, L={L}], got {tuple(c.shape)}')
        if L > max_surfaces:
            raise ValueError('too many surfaces for the staged lens pass')
        P = x_rel.shape[2]
        self.device, self.B, self.L, self.F, self.W, self.P = dev, B, L, F, W, P
        self.shape = (B, F, P, W)
        self.keep = [a.detach().contiguous() for a in (c, t, nd, v, hfov, epd)]
        cc, tt, ndd, vv, hf, ep = self.keep
        self.mu = torch.empty((B, W, L), dtype=torch.float32, device=dev)
        # one allocation: z, half_epd, cy, (reference heights), (aiming map)
        n_small = 2 * B + B * F + (B * F if want_ref else 0) + (B * F * W * 3 if aimed else 0)
        packed = torch.empty((n_small,), dtype=torch.float32, device=dev)
        self.z, self.half_epd, self.cy = packed[:B], packed[B:2 * B], packed[2 * B:2 * B + B * F].view(B, F)
        at = 2 * B + B * F
        self.ref_y = None
        if want_ref:
            self.ref_y = packed[at:at + B * F].view(B, F)
            at += B * F
        self.aim = packed[at:at + B * F * W * 3].view(B, F, W, 3) if aimed else None
        self.ln = tables.lens_struct(cc, tt, ndd, vv, hf, ep)
        self.xy = (x_rel.detach(), y_rel.detach())
        self.vig = None
        if vig is not None:      # [B,F,3] = (x_scale, y_scale, y_offset) of apply_vignetting, per (lens, field)
            nat.require_cuda(vig, 'vig')
            if tuple(vig.shape) != (B, F, 3):
                raise ValueError(f'vig must be [B={B}, F={F}, 3], got {tuple(vig.shape)}')
            self.vig = vig.detach().to(torch.float32).contiguous()
        pb = nat.TlProblem()
        pb.aim = _ptr(self.aim)
        pb.vig = _ptr(self.vig)
        pb.x = nat.strided(self.xy[0], self.shape)
        pb.y = nat.strided(self.xy[1], self.shape)
        pb.z = nat.strided(self.z.reshape(B, 1, 1, 1), self.shape)
        pb.cx = nat.strided(tables.zero, self.shape)
        pb.cy = nat.strided(self.cy.reshape(B, F, 1, 1), self.shape)
        pb.c, pb.t, pb.mu, pb.live = cc.data_ptr(), tt.data_ptr(), self.mu.data_ptr(), tables.mask.data_ptr()
        pb.B, pb.F, pb.P, pb.W, pb.S = B, F, P, W, L
        pb.allow_backward_rays, pb.arith = int(bool(allow_backward_rays)), int(arith)
        pb.p_begin, pb.p_end = int(p_begin), int(P if p_end is None else p_end)
        pb.xy_scale = self.half_epd.data_ptr()
        self.pb = pb
        self.tables = tables
        # staging, ray aiming (rtl:129-208; the map is applied on load) and the reference heights of the
        # fused pass: ONE launch
        nat.check(lib.tl_stage_ref(ctypes.byref(self.ln), ctypes.byref(pb), self.mu.data_ptr(), self.z.data_ptr(),
                                   self.cy.data_ptr(), self.half_epd.data_ptr(), _ptr(self.aim), _ptr(self.vig),
                                   int(aim_mode), int(bool(allow_backward_rays)), _ptr(self.ref_y),
                                   nat.stream_ptr(dev)), 'tl_stage_ref')

    def chain_rule(self, gmu, gz, gc, gt, gnd, gv):
        """ADDS the gradients induced through mu and z to gc, gt, gnd, gv [B,L] (tl_stage_bwd)."""
        nat.check(nat.load().tl_stage_bwd(ctypes.byref(self.ln), gmu.data_ptr(), gz.data_ptr(), gc.data_ptr(),
                                          gt.data_ptr(), gnd.data_ptr(), gv.data_ptr(), nat.stream_ptr(self.device)),
                  'tl_stage_bwd')


def stage_lens(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
               aimed=False, vig=None, aim_mode=nat.AIM_REAL):
    """The staged ray set of a lens batch (one ``tl_stage_ref`` launch), for callers that run several
    fused passes over it (``staged=`` of lens_spot_rms / lens_penalty)."""
    with nat.on_device(c.device):
        return _Staged(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed,
                       max_surfaces=nat.MAX_SURFACES_SPOT, want_ref=True, vig=vig, aim_mode=aim_mode)


def _lens_spot_core(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, shard, group,
                    want_grad, aimed, out=None, staged=None, vig=None, aim_mode=nat.AIM_REAL):
    """The staged fused pass itself (no autograd): staging kernel -> (ray aiming) -> chief rays ->
    fused trace+adjoint -> row reduction -> (all-reduce) -> finalize -> staging chain rule.
    Returns (rms [B], rms_field [B,F], gc, gt, gnd, gv) -- the four gradients of sum(rms) w.r.t. the
    padded [B,L] lens tensors, or None without ``want_grad``.  ``out``: optional dict of
    preallocated float32 tensors 'rms' [B] and 'gc', 'gt', 'gnd', 'gv' [B,L] to write into (a
    caller that owns a packed staging buffer passes views of it and saves the copies)."""
    lib = nat.load()
    out = out or {}
    rank, world = shard
    p_begin, p_end = pupil_slice(x_rel.shape[2], rank, world)

    def buffer(name, shape):
        given = out.get(name)
        if given is None:
            return torch.empty(shape, dtype=torch.float32, device=dev)
        if tuple(given.shape) != tuple(shape) or given.dtype != torch.float32 or not given.is_contiguous():
            raise ValueError(f'out[{name!r}] must be a contiguous float32 tensor of shape {tuple(shape)}')
        return given

    nat.require_cuda(c, 'c')
    dev = c.device
    with nat.on_device(dev):
        if staged is not None:      # the ray set a staged trace_rays of the same lens already built
            st = staged
            if want_grad and st.L > nat.MAX_SURFACES_SPOT:
                raise ValueError('too many surfaces for the staged lens pass')
            st.pb.p_begin, st.pb.p_end = int(p_begin), int(p_end)
        else:
            st = _Staged(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed, p_begin,
                         p_end, max_surfaces=nat.MAX_SURFACES_SPOT if want_grad else 64, want_ref=True, vig=vig,
                         aim_mode=aim_mode)
        B, L, F, W, P = st.B, st.L, st.F, st.W, st.P
        pb = st.pb
        stream = nat.stream_ptr(dev)
        n_acc = lib.tl_spot_moment_count(L, int(want_grad))
        moments = torch.empty((B, F, W, n_acc), dtype=torch.float64, device=dev)
        ws_bytes = lib.tl_spot_workspace(ctypes.byref(pb), int(want_grad))
        if ws_bytes == 0:
            nat.check(-1, 'tl_spot_workspace')
        ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
        if st.ref_y is not None:
            ref_y = st.ref_y
            nat.check(lib.tl_spot_accumulate_ref(ctypes.byref(pb), int(want_grad), moments.data_ptr(),
                                                 ref_y.data_ptr(), ws.data_ptr(), ws_bytes, stream),
                      'tl_spot_accumulate_ref')
        else:
            ref_y = torch.empty((B, F), dtype=torch.float32, device=dev)
            nat.check(lib.tl_spot_accumulate(ctypes.byref(pb), int(want_grad), moments.data_ptr(),
                                             ref_y.data_ptr(), ws.data_ptr(), ws_bytes, stream),
                      'tl_spot_accumulate')
        if world > 1:
            moments = reduce_moments(moments, group)
        rms = buffer('rms', (B,))
        rms_field = torch.empty((B, F), dtype=torch.float32, device=dev)
        gc = gt = gnd = gv = None
        if want_grad:      # finalize + chain rule to (c, t, nd, v): one launch
            gc, gt, gnd, gv = buffer('gc', (B, L)), buffer('gt', (B, L)), buffer('gnd', (B, L)), buffer('gv', (B, L))
            scratch = torch.empty((B * W * L + B,), dtype=torch.float32, device=dev)      # gmu, gz
            spot_out = nat.TlSpotOut(rms.data_ptr(), rms_field.data_ptr(), gc.data_ptr(), gt.data_ptr(),
                                     scratch.data_ptr(), scratch[B * W * L:].data_ptr())
            nat.check(lib.tl_lens_spot_finalize(moments.data_ptr(), ref_y.data_ptr(), ctypes.byref(st.ln), P,
                                                ctypes.byref(spot_out), gnd.data_ptr(), gv.data_ptr(), stream),
                      'tl_lens_spot_finalize')
        else:
            spot_out = nat.TlSpotOut(rms.data_ptr(), rms_field.data_ptr(), None, None, None, None)
            nat.check(lib.tl_spot_finalize(moments.data_ptr(), ref_y.data_ptr(), B, F, W, L, P, 0,
                                           ctypes.byref(spot_out), stream), 'tl_spot_finalize')
    return rms, rms_field, gc, gt, gnd, gv


class _LensTrace(torch.autograd.Function):
    """``RayTracer.trace_rays`` (rtl:80-127) as ONE autograd node over the lens tensors: the staging
    kernel builds the ray set (index model, pupil position, field cosines; ~60 eager tensor ops in
    the reference), tl_trace_fwd traces it; backward = tl_trace_bwd on the upstream gradients of
    x, y, cx, cy + the staging chain rule.  Outputs as trace_skew's: x, y, cx, cy, ray_ok, ray_backward."""

    @staticmethod
    def forward(ctx, c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed, vig=None,
                aim_mode=nat.AIM_REAL):
        if any(ctx.needs_input_grad[4:8]):
            raise ValueError('the staged trace does not differentiate w.r.t. hfov / epd / pupil coordinates')
        lib = nat.load()
        ctx.set_materialize_grads(False)
        dev = c.device
        with nat.on_device(dev):
            st = _Staged(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed,
                         max_surfaces=nat.MAX_SURFACES_FWD, vig=vig, aim_mode=aim_mode)
            outs = torch.empty((4,) + st.shape, dtype=torch.float32, device=dev)
            flags = torch.empty((2,) + st.shape, dtype=torch.bool, device=dev)
            out = nat.TlTraceOut(*[outs[i].data_ptr() for i in range(4)], flags[0].data_ptr(), flags[1].data_ptr(),
                                 None, None, None, None)
            nat.check(lib.tl_trace_fwd(ctypes.byref(st.pb), ctypes.byref(out), nat.stream_ptr(dev)), 'tl_trace_fwd')
        ctx.staged = st
        _LensTrace.last_staged = st      # (picked up by lens_trace right after apply)
        x, y, cx, cy = outs.unbind(0)
        ok, backward = flags.unbind(0)
        ctx.mark_non_differentiable(ok, backward)
        return x, y, cx, cy, ok, backward

    @staticmethod
    def backward(ctx, gx, gy, gcx, gcy, _gok, _gbw):
        st = ctx.staged
        if st.L > nat.MAX_SURFACES_BWD:
            raise ValueError(f'backward supports at most {nat.MAX_SURFACES_BWD} surfaces')
        lib = nat.load()
        dev = st.device
        B, L, W = st.B, st.L, st.W
        with nat.on_device(dev):
            seeds = [None if g is None else g.to(torch.float32).expand(st.shape).contiguous()
                     for g in (gx, gy, gcx, gcy)]
            grads = torch.zeros((4, B, L), dtype=torch.float32, device=dev)      # gc, gt, gnd, gv
            gmu = torch.empty((B, W, L), dtype=torch.float32, device=dev)
            gz = torch.empty((B,), dtype=torch.float32, device=dev)
            sd = nat.TlSeeds(*[_ptr(s) for s in seeds], None, None, None, None)
            gr = nat.TlGrads(grads[0].data_ptr(), grads[1].data_ptr(), gmu.data_ptr(), gz.data_ptr(),
                             None, None, None, None, None, None, None)
            ws_bytes = lib.tl_trace_bwd_workspace(ctypes.byref(st.pb))
            if ws_bytes == 0:
                nat.check(-1, 'tl_trace_bwd_workspace')
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            nat.check(lib.tl_trace_bwd(ctypes.byref(st.pb), ctypes.byref(sd), ctypes.byref(gr), ws.data_ptr(),
                                       ws_bytes, nat.stream_ptr(dev)), 'tl_trace_bwd')
            st.chain_rule(gmu, gz, grads[0], grads[1], grads[2], grads[3])
        need = ctx.needs_input_grad
        return (*[grads[i] if need[i] else None for i in range(4)], *([None] * 10))


def lens_trace(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays=True, arith=nat.ARITH_GUARDED,
               aimed=False, vig=None, aim_mode=nat.AIM_REAL):
    """Staged ``trace_rays``: (x, y, cx, cy, ray_ok, ray_backward), differentiable w.r.t. c, t, nd, v."""
    out = _LensTrace.apply(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, bool(allow_backward_rays), int(arith),
                           bool(aimed), vig, int(aim_mode))
    out[1]._tl_staged = _LensTrace.last_staged      # the ray set, reusable by the fused pass of the same lens
    _LensTrace.last_staged = None
    return out


class _LensPenalty(torch.autograd.Function):
    """``RayTracer.penalty`` as ONE autograd node over the lens tensors: staging kernel (or the ray set
    a staged pass of the same lens already built) -> fused penalty pass -> finalize -> staging chain
    rule.  Returns penalty [B] (sum over rays of Q, optics_simulator_lite.py:441-448)."""

    @staticmethod
    def forward(ctx, c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, shard, group, scale,
                aimed, staged, vig, aim_mode):
        if any(ctx.needs_input_grad[4:8]):
            raise ValueError('the staged penalty pass does not differentiate w.r.t. hfov / epd / pupil coordinates')
        lib = nat.load()
        nat.require_cuda(c, 'c')
        dev = c.device
        rank, world = shard
        p_begin, p_end = pupil_slice(x_rel.shape[2], rank, world)
        with nat.on_device(dev):
            if staged is not None:
                st = staged
                st.pb.p_begin, st.pb.p_end = int(p_begin), int(p_end)
            else:
                st = _Staged(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, aimed, p_begin,
                             p_end, max_surfaces=nat.MAX_SURFACES_BWD, vig=vig, aim_mode=aim_mode)
            if st.L > nat.MAX_SURFACES_BWD:
                raise ValueError(f'the fused penalty pass supports at most {nat.MAX_SURFACES_BWD} surfaces')
            B, L, F, W = st.B, st.L, st.F, st.W
            stream = nat.stream_ptr(dev)
            n_acc = lib.tl_penalty_moment_count(L)
            moments = torch.empty((B, F, W, n_acc), dtype=torch.float64, device=dev)
            ws_bytes = lib.tl_penalty_workspace(ctypes.byref(st.pb))
            if ws_bytes == 0:
                nat.check(-1, 'tl_penalty_workspace')
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            nat.check(lib.tl_penalty_accumulate(ctypes.byref(st.pb), moments.data_ptr(), ws.data_ptr(), ws_bytes,
                                                stream), 'tl_penalty_accumulate')
            if world > 1:
                moments = reduce_moments(moments, group)
            penalty = torch.empty((B,), dtype=torch.float32, device=dev)
            grads = torch.zeros((4, B, L), dtype=torch.float32, device=dev)       # gc, gt, gnd, gv
            gmu = torch.empty((B, W, L), dtype=torch.float32, device=dev)
            gz = torch.empty((B,), dtype=torch.float32, device=dev)
            out = nat.TlPenaltyOut(penalty.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(), gmu.data_ptr(),
                                   gz.data_ptr())
            nat.check(lib.tl_penalty_finalize(moments.data_ptr(), B, F, W, L, float(scale), ctypes.byref(out), stream),
                      'tl_penalty_finalize')
            st.chain_rule(gmu, gz, grads[0], grads[1], grads[2], grads[3])
        ctx.save_for_backward(grads)
        return penalty

    @staticmethod
    def backward(ctx, grad_penalty):
        (grads,) = ctx.saved_tensors
        g = grad_penalty.to(torch.float32).reshape(1, -1, 1)
        need = ctx.needs_input_grad
        scaled = grads * g
        return (*[scaled[i] if need[i] else None for i in range(4)], *([None] * 14))


def lens_penalty(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, n_seq, allow_backward_rays=True,
                 arith=nat.ARITH_GUARDED, shard=(0, 1), group=None, aimed=False, staged=None, vig=None,
                 aim_mode=nat.AIM_REAL):
    """Staged fused penalty pass: sum over rays of Q for every lens, [B], differentiable w.r.t. c, t, nd, v."""
    return _LensPenalty.apply(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, bool(allow_backward_rays), int(arith),
                              (int(shard[0]), int(shard[1])), group, 1.0 / float(n_seq), bool(aimed), staged, vig,
                              int(aim_mode))


class _LensSpotRms(torch.autograd.Function):
    """RayTracer.spot_rms as ONE autograd node over the lens tensors (see _lens_spot_core).
    Seven kernel launches (eight with ray aiming), no per-ray tensor."""

    @staticmethod
    def forward(ctx, c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays, arith, shard,
                group, grad_on, aimed=False, staged=None, vig=None, aim_mode=nat.AIM_REAL):
        if ctx.needs_input_grad[4] or ctx.needs_input_grad[5]:
            raise ValueError('the fused lens pass does not differentiate w.r.t. hfov / epd')
        if ctx.needs_input_grad[6] or ctx.needs_input_grad[7]:
            raise ValueError('the fused lens pass does not differentiate w.r.t. the pupil coordinates x_rel / '
                             'y_rel; use trace() + spot_rms_from_rays() for those gradients')
        want_grad = grad_on and any(ctx.needs_input_grad[:4])
        rms, rms_field, gc, gt, gnd, gv = _lens_spot_core(c, t, nd, v, hfov, epd, x_rel, y_rel, tables,
                                                          allow_backward_rays, arith, shard, group,
                                                          want_grad, aimed, staged=staged, vig=vig,
                                                          aim_mode=aim_mode)
        if want_grad:
            ctx.save_for_backward(gc, gt, gnd, gv)
        ctx.mark_non_differentiable(rms_field)
        return rms, rms_field

    @staticmethod
    def backward(ctx, grad_rms, _grad_field):
        if len(ctx.saved_tensors) < 4:      # forward ran without gradients
            return (None,) * 18
        gc, gt, gnd, gv = ctx.saved_tensors
        g = grad_rms.to(torch.float32).reshape(-1, 1)
        need = ctx.needs_input_grad
        grads = [(gc * g) if need[0] else None, (gt * g) if need[1] else None,
                 (gnd * g) if need[2] else None, (gv * g) if need[3] else None]
        return (*grads, *([None] * 14))


def lens_spot_rms_and_grads(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays=True,
                            arith=nat.ARITH_GUARDED, shard=(0, 1), group=None, aimed=False, out=None, vig=None,
                            aim_mode=nat.AIM_REAL):
    """:func:`lens_spot_rms` and the gradients of ``sum(rms)`` in one call, outside autograd:
    returns ``(rms [B], rms_field [B,F], {'c','t','nd','v': [B,L] gradients})``.  ``out`` may hold
    preallocated 'rms', 'gc', 'gt', 'gnd', 'gv' tensors (e.g. views of one staging buffer)."""
    rms, rms_field, gc, gt, gnd, gv = _lens_spot_core(c, t, nd, v, hfov, epd, x_rel, y_rel, tables,
                                                      bool(allow_backward_rays), int(arith),
                                                      (int(shard[0]), int(shard[1])), group, True,
                                                      bool(aimed), out, vig=vig, aim_mode=aim_mode)
    return rms, rms_field, {'c': gc, 't': gt, 'nd': gnd, 'v': gv}


def lens_spot_rms(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, allow_backward_rays=True,
                  arith=nat.ARITH_GUARDED, shard=(0, 1), group=None, aimed=False, staged=None, vig=None,
                  aim_mode=nat.AIM_REAL):
    """(rms [B], rms_field [B,F]) of a lens batch given as padded [B,L] tensors, differentiable
    w.r.t. c, t, nd, v.  x_rel, y_rel: relative pupil coordinates [1,1,P,1].  ``aimed``: with one
    iteration of 'real' ray aiming (rtl:129-208) done by tl_aim and applied inside the kernels."""
    return _LensSpotRms.apply(c, t, nd, v, hfov, epd, x_rel, y_rel, tables, bool(allow_backward_rays),
                              int(arith), (int(shard[0]), int(shard[1])), group, torch.is_grad_enabled(),
                              bool(aimed), staged, vig, int(aim_mode))


# ---------------------------------------------------------------------------
# Spot-diagram / PSF binning (compute_psf, ray_tracing.py:206-270 of the reference's TensorFlow original)
# ---------------------------------------------------------------------------
def psf_bin(x, y, y_target, x_incr, y_incr, x_size, y_size, n_bins):
    """Gaussian soft histogram of image-plane points: x, y [G, C, R] (grid = lens x field, colour
    channel, ray), per-grid centre / pitch / window [G].  Returns (sums [G, C, n_y, n_xh] float64 --
    un-normalised, non-negative half of the x bins -- and inside [G, C] float64, the rays inside the
    window).  One binning kernel + one fixed-order reduction; not differentiable."""
    for name, v in (('x', x), ('y', y), ('y_target', y_target), ('x_incr', x_incr), ('y_incr', y_incr),
                    ('x_size', x_size), ('y_size', y_size)):
        nat.require_cuda(v, name)
        if v.dtype != torch.float32:
            raise TypeError(f'{name} must be float32')
    if x.dim() != 3 or x.shape != y.shape:
        raise ValueError('x and y must be [G, C, R] tensors of one shape')
    G, C, R = x.shape
    n_x, n_y = int(n_bins[0]), int(n_bins[1])
    lib = nat.load()
    dev = x.device
    with nat.on_device(dev):
        keep = [v.detach().contiguous() for v in (x, y, y_target, x_incr, y_incr, x_size, y_size)]
        for v in keep[2:]:
            if v.numel() != G:
                raise ValueError('per-grid arguments must have G elements')
        p = nat.TlPsf(*[v.data_ptr() for v in keep], G, C, R, n_x, n_y)
        n_xh = n_x // 2 + 1 if n_x % 2 else n_x // 2
        ws_bytes = lib.tl_psf_workspace(ctypes.byref(p))
        if ws_bytes == 0:
            nat.check(-1, 'tl_psf_workspace')
        ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
        sums = torch.empty((G, C, n_y, n_xh), dtype=torch.float64, device=dev)
        inside = torch.empty((G, C), dtype=torch.float64, device=dev)
        nat.check(lib.tl_psf_bin(ctypes.byref(p), sums.data_ptr(), inside.data_ptr(), ws.data_ptr(), ws_bytes,
                                 nat.stream_ptr(dev)), 'tl_psf_bin')
    return sums, inside


# ---------------------------------------------------------------------------
# Paraxial front end of a padded lens batch (get_first_order rtl:772-794, compute_last_curvature
# rtl:725-769): one thread per lens, forward and hand-derived adjoint (csrc/paraxial.cuh)
# ---------------------------------------------------------------------------
class _Paraxial(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, t, n, live, glass, mode):
        for name, v in (('c', c), ('t', t), ('n', n)):
            nat.require_cuda(v, name)
            if v.dtype != torch.float32 or v.dim() != 2:
                raise TypeError(f'{name} must be a [B, L] float32 tensor')
        if not (c.shape == t.shape == n.shape == live.shape == glass.shape):
            raise ValueError('c, t, n, live, glass must share one [B, L] shape')
        B, L = c.shape
        if L > nat.PARAXIAL_MAX_SLOTS:
            raise ValueError(f'at most {nat.PARAXIAL_MAX_SLOTS} slots per lens')
        lib = nat.load()
        dev = c.device
        with nat.on_device(dev):
            keep = [v.detach().contiguous() for v in (c, t, n)] + [live, glass]
            p = nat.TlParaxial(*[v.data_ptr() for v in keep], B, L, mode)
            out = torch.empty((B, 2), dtype=torch.float32, device=dev)
            nat.check(lib.tl_paraxial_fwd(ctypes.byref(p), out.data_ptr(), nat.stream_ptr(dev)), 'tl_paraxial_fwd')
        ctx.keep, ctx.mode = keep, mode
        ctx.mark_non_differentiable(live, glass)
        return out

    @staticmethod
    def backward(ctx, gout):
        c, t, n, live, glass = ctx.keep
        B, L = c.shape
        lib = nat.load()
        dev = c.device
        with nat.on_device(dev):
            gout = gout.detach().to(torch.float32).contiguous()
            gc, gt, gn = (torch.empty((B, L), dtype=torch.float32, device=dev) for _ in range(3))
            p = nat.TlParaxial(c.data_ptr(), t.data_ptr(), n.data_ptr(), live.data_ptr(), glass.data_ptr(), B, L, ctx.mode)
            nat.check(lib.tl_paraxial_bwd(ctypes.byref(p), gout.data_ptr(), gc.data_ptr(), gt.data_ptr(), gn.data_ptr(),
                                          nat.stream_ptr(dev)), 'tl_paraxial_bwd')
        return gc, gt, gn, None, None, None


def _mask_bytes(structure):
    """(live, glass) uint8 [B,L] device tensors of a Structure, cached on it."""
    def build():
        dev = structure.mask_torch.device
        return (structure.mask_torch.to(torch.uint8).contiguous(), structure.mask_G_torch.to(torch.uint8).contiguous())
    return structure.device_tables('paraxial_masks', build)


def first_order(structure, c, t, n_after):
    """(EFL, BFL) [B] of padded lenses: c, t [B,L]; n_after [B,L] the index behind every slot (1 behind
    air and padding).  One launch; differentiable w.r.t. c, t, n_after (one more launch)."""
    live, glass = _mask_bytes(structure)
    out = _Paraxial.apply(c, t, n_after, live, glass, nat.PARAXIAL_FIRST_ORDER)
    return out[:, 0], out[:, 1]


def last_curvature(structure, c, t, n_after):
    """(solved [B], slot [B] int64): the curvature of the last glass-air surface that makes EFL = 1
    (rtl:725-769) of padded lenses; c at and behind the solved slot is not read."""
    live, glass = _mask_bytes(structure)
    out = _Paraxial.apply(c, t, n_after, live, glass, nat.PARAXIAL_LAST_CURVATURE)
    return out[:, 0], out[:, 1].detach().to(torch.int64)
