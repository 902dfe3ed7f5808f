"""Small end-to-end case for compute-sanitizer: every kernel of the library once."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, prescriptions, ray_tracing_lite as rt   # noqa: E402

dev = 'cuda:0'
for name, loader in (('cooke', lambda: prescriptions.load_yaml('baseline_cooke.yml', dev, epd_scale=2.6)),
                     ('asphere', lambda: prescriptions.asphere_12(dev, f_number=4.0)),
                     ('zoom30', lambda: prescriptions.wide_zoom_30(dev))):
    specs, lens = loader()
    for k in ('c', 't', 'nd'):
        getattr(lens, k).requires_grad_(True)
    tracer = RayTracer(mode='circular', n_rays=(13, 11), rel_fields=(0., 0.6, 1.), wavelengths=('C', 'd', 'F'),
                       default_device=dev)
    out = tracer.trace_rays(specs, lens)
    if name != 'asphere':
        rms = rt.compute_rms2d(out[0], out[1], out[4])
        rms.backward()
    if lens.c.shape[1] <= 16:
        for k in ('c', 't', 'nd'):
            getattr(lens, k).grad = None
        rms2, _ = tracer.spot_rms(specs, lens)
        rms2[0].backward()
    with torch.no_grad():
        rms3, _ = tracer.spot_rms(specs, lens)
    torch.cuda.synchronize()
    print(name, 'ok', float(out[4].float().mean()), 'rms', float(rms3[0]))
