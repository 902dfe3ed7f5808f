"""BASELINE.json configs 3, 4 and 5 at their FULL sizes on one B200, with the size-independent
properties the parity tests use at small sizes (ok fraction, additivity of pupil slices, finite
gradients, a decreasing loss).  Writes one JSON object to stdout.

    python tools/full_size_configs.py > gpurun_out/full_size_configs.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402
from torchoptics_b200.optimize import optimize_spot          # noqa: E402

dev = 'cuda:0'
FIELDS = tuple(np.linspace(0, 1, 16).tolist())
WL = ('C', 'd', 'F')
res = {}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def rel(a, b):
    denom = b.abs().amax(dim=(0, 1, 2)).clamp_min(1e-30)
    return float(((a - b).abs().amax(dim=(0, 1, 2)) / denom).max())


# ---- config 3: 12 even-asphere surfaces, 64 M rays, forward + backward -------------------------
side = 1183
specs, lens = prescriptions.asphere_12(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=FIELDS, wavelengths=WL, default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
ext = {k: v.detach() for k, v in tracer._extension_tables(lens).items() if v is not None}
rays = 16 * 3 * side * side
ms, (whole, ref_y) = timed(lambda: ops.spot_moments(*args, **ext), 3)
parts = [ops.spot_moments(*args, shard=(r, 2), **ext)[0] for r in range(2)]
for name in ('c', 'k', 'a'):
    getattr(lens, name).requires_grad_(True)
rms, _ = tracer.spot_rms(specs, lens)
grads = torch.autograd.grad(rms[0], [lens.c, lens.k, lens.a])
res['config3'] = {'lens': 'asphere_12 (12 even-asphere surfaces, a4..a16; oracle: 4 Newton steps, fast policy: early exit)', 'rays': rays,
                  'events': rays * 12, 'fwd_bwd_ms': ms, 'events_per_s': rays * 12 / (ms * 1e-3),
                  'ok_fraction': float(whole[..., -1].sum()) / rays,
                  'two_slices_vs_whole_rel': rel(parts[0] + parts[1], whole),
                  'rms': float(rms[0]), 'grads_finite': bool(all(torch.isfinite(g).all() for g in grads)),
                  'grad_c_norm': float(grads[0].norm())}
del whole, parts, args
torch.cuda.empty_cache()

# ---- config 4: ~30 surfaces, 1 B rays, forward spot sweep, sharded by pupil slice ------------
side = 4565
specs, lens = prescriptions.wide_zoom_30(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=FIELDS, wavelengths=WL, default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
S = args[6].shape[-1]
rays = 16 * 3 * side * side
ms, (whole, _) = timed(lambda: ops.spot_moments(*args, want_grad=False), 2)
slices = None
t_slices = []
for r in range(8):                              # what each of 8 ranks would trace
    ms_r, (m_r, _) = timed(lambda r=r: ops.spot_moments(*args, want_grad=False, shard=(r, 8)), 1)
    t_slices.append(ms_r)
    slices = m_r if slices is None else slices + m_r
rms_eval, rms_field = ops.spot_rms(*args)
res['config4'] = {'lens': f'wide_zoom_30 ({S} spherical surfaces)', 'rays': rays, 'events': rays * S,
                  'fwd_sweep_ms_one_gpu': ms, 'events_per_s_one_gpu': rays * S / (ms * 1e-3),
                  'ok_fraction': float(whole[..., -1].sum()) / rays,
                  'eight_slices_vs_whole_rel': rel(slices, whole),
                  'slice_ms_max': max(t_slices), 'events_per_s_8_slices_concurrent_estimate': rays * S / (max(t_slices) * 1e-3),
                  'rms': float(rms_eval[0]), 'rms_field_min_max': [float(rms_field.min()), float(rms_field.max())]}
del whole, slices, args
torch.cuda.empty_cache()

# ---- config 5: Adam on the asphere lens, one GPU's share (32 M rays / step) of 256 M rays ------
side = 816
specs, lens = prescriptions.asphere_12(dev)
tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=FIELDS, wavelengths=WL, default_device=dev)
rays = 16 * 3 * side * side
steps = 60
stamps = []


def stamp(step, loss):
    torch.cuda.synchronize()
    stamps.append(time.perf_counter())


best, history = optimize_spot(tracer, specs, lens, steps=steps, lr=5e-5, callback=stamp)
step_s = float(np.median(np.diff(stamps)))
res['config5'] = {'what': f'{steps} Adam steps (c, t, k, a; lr 5e-5) on asphere_12, {rays} rays/step = one GPU share of '
                          '256 M rays/step on 8 GPUs', 'rays_per_step': rays, 'median_s_per_step': step_s,
                  'events_per_s_incl_host_and_optimizer': rays * 12 / step_s,
                  'rms_first_min_last': [history[0], min(history), history[-1]],
                  'reduced_by': 1.0 - min(history[-5:]) / history[0]}
print(json.dumps(res, indent=1))
