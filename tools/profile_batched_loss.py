"""Batched-lens workload: the fused penalty pass and the fused spot pass for B lenses x 1536 rays."""
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
S = c.shape[-1]
events = B * 8 * 3 * 64 * S


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


ms_p = timed(lambda: ops.penalty_sum(*args, S))
ms_s = timed(lambda: ops.spot_rms(*args))
print(f'{B} lenses x 1536 rays, S={S}: penalty pass {ms_p:.4f} ms ({events / ms_p / 1e6:.1f} G events/s), '
      f'spot pass {ms_s:.4f} ms ({events / ms_s / 1e6:.1f} G events/s)')
