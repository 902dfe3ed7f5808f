"""Where the host time of the reference's own call sequence goes (trace_rays -> compute_rms2d -> backward, config 2):
cProfile over N iterations, top functions by own and cumulative time, plus the wall time per iteration with and
without the profiler.   python tools/profile_dropin.py [N]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, lens_modeling as lm, prescriptions, ray_tracing_lite as rt   # noqa: E402

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
tracer = RayTracer(mode='circular', n_rays=(296, 296), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
host_lens = {k: getattr(lens, k).detach().cpu().pin_memory() for k in ('c', 't', 'nd', 'v')}
S = lens.c.shape[1]
result_host = torch.empty((1 + 3 * S,), dtype=torch.float32).pin_memory()


def step():
    dl = {k: v.to(dev, non_blocking=True) for k, v in host_lens.items()}
    for k in ('c', 't', 'nd'):
        dl[k].requires_grad_(True)
    lens_i = lm.Lens(lens.structure, dl['c'], dl['t'], dl['nd'], dl['v'])
    out = tracer.trace_rays(specs, lens_i)
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    rms.backward()
    result_host.copy_(torch.cat([rms.detach().reshape(1)] + [dl[k].grad.reshape(-1) for k in ('c', 't', 'nd')]),
                      non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return float(result_host[0])


for _ in range(20):
    step()
t0 = time.perf_counter()
for _ in range(n_iter):
    step()
plain = (time.perf_counter() - t0) / n_iter
# the same with the launches left to run on (no sync): the host time alone
def step_nosync():
    dl = {k: v.to(dev, non_blocking=True) for k, v in host_lens.items()}
    for k in ('c', 't', 'nd'):
        dl[k].requires_grad_(True)
    lens_i = lm.Lens(lens.structure, dl['c'], dl['t'], dl['nd'], dl['v'])
    out = tracer.trace_rays(specs, lens_i)
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    rms.backward()
torch.cuda.synchronize()
prof = cProfile.Profile()
prof.enable()
for _ in range(n_iter):
    step()
prof.disable()
print(f'wall per iteration: {plain * 1e6:.1f} us  ({16 * 3 * 296 * 296 * 11 / plain / 1e9:.1f} G events/s)')
for key in ('tottime', 'cumulative'):
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).strip_dirs().sort_stats(key).print_stats(28)
    text = buf.getvalue()
    print(text[text.index('ncalls'):])
