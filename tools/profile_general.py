"""Timing of the general-surface fused spot pass on the config-3 lens (12 even-asphere surfaces).
python tools/profile_general.py [n_side]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dev = 'cuda:0'
specs, lens = prescriptions.asphere_12(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
ext = {k: v.detach() for k, v in tracer._extension_tables(lens).items() if v is not None}
events = 16 * 3 * side * side * 12
for want_grad in (True, False):
    for _ in range(3):
        m, _ = ops.spot_moments(*args, want_grad=want_grad, **ext)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m, _ = ops.spot_moments(*args, want_grad=want_grad, **ext)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # Two conventions.  `flops`: the oracle's algorithm, 4 fixed Newton steps: 61 + 55 n + 35
    # forward, + 105 + 2 x 55 + 35 adjoint.  `done`: what the fast policy executes on THIS lens since the early exit
    # (two steps per event on average, csrc/trace_core_asph.cuh: newton_settled): 348 - 2 x 57 forward, + 271 adjoint.
    flops = 566 if want_grad else 316
    done = 505 if want_grad else 234
    print(f'general spot pass want_grad={want_grad}: {ms:.4f} ms -> {events / ms / 1e6:.1f} G events/s '
          f'({events * flops / ms / 1e9 / 74.45 * 100:.1f}% of 74.45 TFLOP/s credited at the oracle\'s {flops} flop/event, '
          f'{events * done / ms / 1e9 / 74.45 * 100:.1f}% at the ~{done} executed), '
          f'ok fraction {float(m[..., -1].sum()) / (16 * 3 * side * side):.4f}')
