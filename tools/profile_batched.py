"""The reference's real workload shape: many lenses, few rays each
(8 fields x 3 wavelengths x 8x8 pupil = 1 536 rays per lens).  Fused spot pass fwd+bwd.
`python tools/profile_batched.py [n_lenses]`"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
SIDE = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = 'cuda:0'
specs, lens = prescriptions.load_yaml('baseline_cooke.yml', dev)
tracer = RayTracer(mode='circular', n_rays=(SIDE, SIDE), rel_fields=tuple(np.linspace(0, 1, 8).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
x, y, z, cx, cy, c, t, mu, mask = [a.detach() for a in tracer._ray_set(specs, lens)]
g = torch.Generator(device='cpu').manual_seed(0)
jit = (1.0 + 0.01 * torch.randn((B, 1, 1, 1, c.shape[-1]), generator=g)).to(dev)
args = [x, y, z.expand(B, 1, 1, 1).contiguous(), cx, cy.expand(B, -1, 1, 1).contiguous(),
        (c * jit).contiguous(), t.expand(B, 1, 1, 1, -1).contiguous(),
        mu.expand(B, 1, 1, -1, -1).contiguous(), mask.expand(B, 1, 1, 1, -1).contiguous()]
S = c.shape[-1]
rays = B * 8 * 3 * SIDE * SIDE
events = rays * S
for _ in range(3):
    m, _ = ops.spot_moments(*args)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    ops.spot_moments(*args)
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(graph):
    m, _ = ops.spot_moments(*args)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f'{B} lenses x {8 * 3 * SIDE * SIDE} rays (pupil {SIDE * SIDE}), S={S}: fused spot pass {ms:.4f} ms -> {events / ms / 1e6:.1f} G events/s '
      f'({events * 166 / ms / 1e9 / 74.45 * 100:.1f}% of FP32 peak), n_ok={float(m[..., -1].sum()):.0f} of {rays}')
