"""Small driver for ncu: a few launches of the fused spot pass (config-2 lens,
4.2 M rays) with inputs resident in HBM.  `python tools/profile_spot.py [n_launches] [n_side]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
side = int(sys.argv[2]) if len(sys.argv) > 2 else 296
dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
for _ in range(n):
    m, _ = ops.spot_moments(*args)
torch.cuda.synchronize()
reps = max(n, 1) * 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):                       # back to back: the GPU, not Python, sets the pace
    m, _ = ops.spot_moments(*args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
events = 16 * 3 * side * side * 11
print(f'spot_moments: {ms:.4f} ms/launch -> {events / ms / 1e6:.1f} G events/s '
      f'({events * 166 / ms / 1e9 / 74.45 * 100:.1f}% of 74.45 TFLOP/s), n_ok={float(m[..., -1].sum()):.0f}')
