"""A few launches of the fused penalty pass and of trace_skew(aggregate=True) (config-2 lens,
4.2 M rays) for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402
from torchoptics_b200 import ray_tracing_lite as rt          # noqa: E402

dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
tracer = RayTracer(mode='circular', n_rays=(296, 296), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
for _ in range(3):
    pen = ops.penalty_sum(*args, args[6].shape[-1])
    out = rt.trace_skew(*args, aggregate=True)
torch.cuda.synchronize()
print(float(pen[0]))
