"""Timing of the aggregate=True path (penalty stacks, rtl:641-657) on the config-2 lens:
forward with the three [S,B,F,P,W] stacks, and forward + backward of rms + 0.2 * sum(Q)
(compute_loss_out, optics_simulator_lite.py:430-450).  `python tools/profile_penalty.py [n_side]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, prescriptions   # noqa: E402
from torchoptics_b200 import ray_tracing_lite as rt     # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
S = args[6].shape[-1]
rays = 16 * 3 * side * side
events = rays * S


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fwd():
    return rt.trace_skew(*args, aggregate=True)


def fwd_plain():
    return rt.trace_skew(*args)


leaves = [args[j].clone().requires_grad_(True) for j in (2, 5, 6, 7)]
call = list(args)
call[2], call[5], call[6], call[7] = leaves


def loss_step():
    out = rt.trace_skew(*call, aggregate=True)
    st = out[6]
    q = (torch.stack(st['theta_norm']).sum(0) + torch.stack(st['theta_prime_norm']).sum(0) +
         torch.stack(st['z_RELU']).sum(0)) / S
    loss = rt.compute_rms2d(out[0], out[1], out[4]) + 0.2 * q.sum()
    return torch.autograd.grad(loss, leaves)


def fused_step():
    lens_leaves = [getattr(lens, k).detach().clone().requires_grad_(True) for k in ('c', 't', 'nd')]
    from torchoptics_b200.lens_modeling import Lens
    res = tracer.loss_unsup(specs, Lens(lens.structure, *lens_leaves, lens.v))
    return torch.autograd.grad(res['loss_unsup'][0], lens_leaves)


def penalty_only():
    from torchoptics_b200 import ops
    return ops.penalty_sum(*args, S)


ms_plain = timed(fwd_plain)
ms_fwd = timed(fwd)
ms_step = timed(loss_step, 5)
bytes_fwd = rays * (12 * S + 18)
print(f'rays {rays}  S {S}')
print(f'trace_skew                  : {ms_plain:.3f} ms  {events / ms_plain / 1e6:.1f} G events/s')
print(f'trace_skew(aggregate=True)  : {ms_fwd:.3f} ms  {events / ms_fwd / 1e6:.1f} G events/s  '
      f'{bytes_fwd / ms_fwd / 1e6:.0f} GB/s of stores ({bytes_fwd / 1e6:.0f} MB)')
print(f'loss fwd+bwd (unfused, torch stack/sum glue included): {ms_step:.3f} ms  '
      f'{events / ms_step / 1e6:.1f} G events/s')
ms_pen = timed(penalty_only)
ms_fused = timed(fused_step, 5)
print(f'fused penalty pass (value + gradient, nothing materialised): {ms_pen:.3f} ms  '
      f'{events / ms_pen / 1e6:.1f} G events/s')
print(f'RayTracer.loss_unsup fwd+bwd (spot pass + penalty pass, eager front end): {ms_fused:.3f} ms  '
      f'{events / ms_fused / 1e6:.1f} G events/s')

from torchoptics_b200 import GraphedSpotStep   # noqa: E402
import time                                      # noqa: E402
step = GraphedSpotStep(tracer, specs, lens, penalty_rate=0.2)
host = {k: getattr(lens, k).detach().cpu() for k in ('c', 't', 'nd', 'v')}
for _ in range(3):
    step(**host)
t0 = time.perf_counter()
for _ in range(50):
    rms, grads = step(**host)
secs = (time.perf_counter() - t0) / 50
print(f'GraphedSpotStep(penalty_rate=0.2): host prescription in, loss + gradients out: {secs * 1e3:.3f} ms  '
      f'{events / secs / 1e9:.1f} G events/s  (rms {float(rms[0]):.6f}, penalty {float(step.host_penalty[0]):.3f})')
