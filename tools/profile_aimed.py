"""Config-2 lens with ray aiming on (the reference's default, osl:359): GraphedSpotStep host->host,
staged path with tl_aim applied on load vs the torch mirror of rtl:129-208 feeding the unstaged pass."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import GraphedSpotStep, RayTracer, prescriptions   # noqa: E402

dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
host = {k: getattr(lens, k).detach().cpu() for k in ('c', 't', 'nd', 'v')}
events = 16 * 3 * 296 * 296 * lens.c.shape[1]
for label, device_aiming in (('tl_aim, map applied on load (staged)', True), ('torch mirror of rtl:129-208 (unstaged)', False)):
    tracer = RayTracer(mode='circular', n_rays=(296, 296), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                       wavelengths=('C', 'd', 'F'), n_ray_aiming_iter=1, default_device=dev)
    tracer.device_aiming = device_aiming
    if device_aiming:
        step = GraphedSpotStep(tracer, specs, lens)
    else:       # the mirror's boolean-mask indexing (lens.up_to_stop) cannot be captured: eager calls
        def step(**_):
            rms, grads = tracer.spot_rms_and_grads(specs, lens)
            return rms.cpu(), {k: g.cpu() for k, g in grads.items()}
    for _ in range(3):
        step(**host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        rms, _ = step(**host)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f'aimed, {label}: {ms:.3f} ms/step host to host, {events / ms / 1e6:.1f} G events/s, rms {float(rms[0]):.7f}')
