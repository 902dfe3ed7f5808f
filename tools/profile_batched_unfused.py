"""Batched-lens workload through the UNFUSED drop-in calls: trace_skew forward (and backward
through compute_rms2d) for 1 024 lenses x 1 536 rays.  `python tools/profile_batched_unfused.py [B]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = 'cuda:0'
specs, lens = prescriptions.load_yaml('baseline_cooke.yml', dev)
tracer = RayTracer(mode='circular', n_rays=(8, 8), rel_fields=tuple(np.linspace(0, 1, 8).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
x, y, z, cx, cy, c, t, mu, mask = [a.detach() for a in tracer._ray_set(specs, lens)]
args = [x, y, z.expand(B, 1, 1, 1).contiguous(), cx, cy.expand(B, -1, 1, 1).contiguous(),
        c.expand(B, 1, 1, 1, -1).contiguous(), t.expand(B, 1, 1, 1, -1).contiguous(),
        mu.expand(B, 1, 1, -1, -1).contiguous(), mask.expand(B, 1, 1, 1, -1).contiguous()]
events = B * 8 * 3 * 64 * c.shape[-1]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: ops.trace(*args))
print(f'{B} lenses x 1536 rays: tl_trace_fwd {ms:.4f} ms -> {events / ms / 1e6:.1f} G events/s')
leaves = [args[j].clone().requires_grad_(True) for j in (2, 5, 6, 7)]
call = list(args)
call[2], call[5], call[6], call[7] = leaves


def drop_in():
    out = ops.trace(*call)
    rms, _ = ops.spot_rms_from_rays(out[1], out[4])
    return torch.autograd.grad(rms.sum(), leaves)


ms = timed(drop_in)
print(f'{B} lenses x 1536 rays: unfused trace_skew -> compute_rms2d -> backward {ms:.4f} ms -> '
      f'{events / ms / 1e6:.1f} G events/s')
