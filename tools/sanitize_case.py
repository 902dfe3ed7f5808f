"""Small end-to-end case for compute-sanitizer: every kernel of the library once (round 2: the variants of
k_spot_rev incl. the forward-only and the 4-warp one, staging with vignetting / aiming, the warp-per-row kernels,
the PSF binning, the paraxial kernels, the batched loss).   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, prescriptions, ray_tracing_lite as rt   # noqa: E402
from torchoptics_b200.optical_loss import Optical_Loss                          # noqa: E402

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
        rms = rt.compute_rms2d(out[0], out[1], out[4])           # fused pass through the provenance (k_spot_rev)
        rms.backward()
    for k in ('c', 't', 'nd'):
        getattr(lens, k).grad = None
    rms2, _ = tracer.spot_rms(specs, lens)                        # cooke: tmem12; zoom30: tmem4; asphere: k_trace_gen
    rms2[0].backward()
    with torch.no_grad():
        rms3, _ = tracer.spot_rms(specs, lens)                    # forward-only variant
    torch.cuda.synchronize()
    print(name, 'ok', float(out[4].float().mean()), 'rms', float(rms3[0]))

# every gradient variant of k_spot_rev on the over-filled Cooke pupil (misses, exact re-traces, dead-lane mirroring)
specs, lens = prescriptions.load_yaml('baseline_cooke.yml', dev, epd_scale=2.6)
for variant in ('tmem12', 'tmem12c2', 'tmem16', 'tmem8', 'tmem4', 'reg8'):
    os.environ['TL_REV'] = variant
    lens.c.requires_grad_(True)
    tracer = RayTracer(mode='circular', n_rays=(24, 21), rel_fields=(0., 0.7, 1.), wavelengths=('C', 'd', 'F'), default_device=dev)
    rms, _ = tracer.spot_rms(specs, lens)
    rms[0].backward()
    torch.cuda.synchronize()
    print(variant, 'ok', float(rms[0]))
os.environ.pop('TL_REV')

# vignetting + ray aiming ('real' and 'paraxial') in the staging kernel, penalty pass, loss
for mode in ('real', 'paraxial'):
    specs, lens = prescriptions.load_yaml('baseline_tessar.yml', dev)
    lens.c.requires_grad_(True)
    tracer = RayTracer(mode='circular', n_rays=(8, 8), rel_fields=(0., 0.7, 1.), wavelengths=('C', 'd', 'F'),
                       vig_fn=lambda fields, vig: vig[:, None] * fields, n_ray_aiming_iter=1, ray_aiming_mode=mode,
                       default_device=dev)
    specs.vig_up = torch.full_like(specs.epd, 0.1)
    specs.vig_down = torch.full_like(specs.epd, 0.05)
    specs.vig_x = torch.full_like(specs.epd, 0.02)
    loss = tracer.loss_unsup(specs, lens)['loss_unsup']
    loss[0].backward()
    torch.cuda.synchronize()
    print('aimed', mode, 'ok', float(loss[0]))

# batched lenses: warp-per-row kernels, paraxial kernels
with np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'optical_loss',
                          'GAGA.npz')) as z:
    x = torch.from_numpy(np.tile(z['inputs'], (12, 1))).to(dev)
    y = torch.from_numpy(np.tile(z['outputs'], (12, 1))).to(dev).requires_grad_(True)
loss, rms, pen = Optical_Loss('GAGA').optical_loss_unsupervised(x, y, 0.2, dev)
loss.backward()
torch.cuda.synchronize()
print('batched loss ok', float(loss))

# PSF binning
g = torch.Generator(device='cpu').manual_seed(0)
px = (torch.randn((1, 3, 3, 5000), generator=g) * 0.006).abs().to(dev)
py = (torch.randn((1, 3, 3, 5000), generator=g) * 0.008).to(dev)
for bins in ((21, 21), (10, 12), (64, 64), (1, 1)):
    k = rt.compute_psf(px, py, bins, 0.002)[3]
    torch.cuda.synchronize()
    print('psf', bins, 'ok', float(k.sum()))
