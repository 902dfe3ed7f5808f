"""The drop-in boundary as a reference user meets it:

* the reference's own import lines, unchanged, against the `torchlens` alias package;
* INTEGRATION.md's patch B, extracted from the document and executed verbatim against the
  UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) -- the reference's own
  front end (RayTracer.trace_rays, ray aiming, compute_pupil_radius) then runs on the GPU kernels;
* the reference's call sequence trace_rays -> compute_rms2d -> backward takes the fused pass through
  the provenance of the trace outputs, and falls back to the materialised reduction the moment the
  outputs are touched.
"""
import os
import re
import sys

import numpy as np
import pytest
import torch

from tests.conftest import load_golden
from tests.test_gpu_parity import DEV, GRAD_TOL, RMS_TOL, _check_outputs, _close_or_no_worse_than_reference, _rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_lines_work_unchanged():
    import torchlens.lens_modeling as lm                  # the reference's spelling (lm:17, osl:7-8)
    import torchlens.ray_tracing_lite as rt
    import torchlens.ray_tracing as rt_tf
    import torchoptics_b200.ray_tracing_lite as ours
    assert rt is ours and rt_tf is ours
    rec = load_golden('cooke_8x8')
    structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device=DEV)
    lens = lm.Lens(structure, *[torch.from_numpy(rec[k]).to(DEV).requires_grad_(True)
                                for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    specs = lm.Specs(structure, torch.from_numpy(rec['epd']).to(DEV), torch.from_numpy(rec['hfov']).to(DEV))
    tracer = rt.RayTracer(mode='circular', n_rays=(8, 8), rel_fields=(0., 0.707, 1.), wavelengths=('C', 'd', 'F'),
                          default_device=DEV)
    x, y, cx, cy, ray_ok, ray_backward = tracer.trace_rays(specs, lens)
    rms = rt.compute_rms2d(x, y, ray_ok)
    rms.backward()
    assert abs(rms.item() - float(rec['rms'])) <= RMS_TOL * float(rec['rms'])
    for name in ('c', 't', 'nd'):
        _close_or_no_worse_than_reference(getattr(lens, name).grad.cpu().numpy(), rec['grad_' + name],
                                          rec['f64_grad_' + name], GRAD_TOL, f'alias d rms/d {name}')


@pytest.mark.parametrize('name', ['cooke_8x8', 'cooke_16x16_epd2.6', 'tessar_8x8_aimed', 'cooke_96x76_config1'])
def test_drop_in_sequence_takes_the_fused_pass(name, monkeypatch):
    """trace_rays -> compute_rms2d -> backward, exactly as the reference spells it: same numbers as the
    golden record, and the RMS is evaluated by the fused pass (no materialised reduction is launched)
    unless the outputs were modified."""
    from torchoptics_b200 import lens_modeling as lm, ops
    from torchoptics_b200 import ray_tracing_lite as rt
    rec = load_golden(name)
    rec['name'] = name
    structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device=DEV)

    def problem():
        lens = lm.Lens(structure, *[torch.from_numpy(rec[k]).to(DEV).requires_grad_(True)
                                    for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
        specs = lm.Specs(structure, torch.from_numpy(rec['epd']).to(DEV), torch.from_numpy(rec['hfov']).to(DEV))
        tracer = rt.RayTracer(mode='circular', n_rays=tuple(int(v) for v in rec['n_rays']),
                              rel_fields=tuple(float(v) for v in rec['rel_fields']),
                              wavelengths=tuple(float(v) for v in rec['wavelengths']),
                              n_ray_aiming_iter=1 if 'aimed' in name else 0,
                              allow_backward_rays=bool(rec['allow_backward_rays']), default_device=DEV)
        return tracer, specs, lens

    calls = []
    real = ops.spot_rms_from_rays
    monkeypatch.setattr(ops, 'spot_rms_from_rays', lambda *a, **k: (calls.append('materialised'), real(*a, **k))[1])
    tracer, specs, lens = problem()
    out = tracer.trace_rays(specs, lens)
    if 'aimed' not in name:
        _check_outputs([o.detach() for o in out], rec)
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    assert calls == [], 'untouched trace outputs must take the fused pass'
    _close_or_no_worse_than_reference(rms.item(), rec['rms'], rec['f64_rms'], RMS_TOL, f'{name} drop-in rms')
    rms.backward()
    fused_grads = [getattr(lens, k).grad.clone() for k in ('c', 't', 'nd')]
    for k, g in zip(('c', 't', 'nd'), fused_grads):
        _close_or_no_worse_than_reference(g.cpu().numpy(), rec['grad_' + k], rec['f64_grad_' + k], GRAD_TOL,
                                          f'{name} drop-in d rms/d {k}')
    # a touched y (here: a no-op arithmetic copy) loses its provenance: materialised reduction, same numbers
    tracer, specs, lens = problem()
    out = tracer.trace_rays(specs, lens)
    rms2 = rt.compute_rms2d(out[0], out[1] * 1.0, out[4])
    assert calls == ['materialised']
    rms2.backward()
    assert abs(rms2.item() - rms.item()) <= RMS_TOL * rms.item()
    for k, g in zip(('c', 't', 'nd'), fused_grads):
        assert _rel(getattr(lens, k).grad.cpu().numpy(), g.cpu().numpy()) <= GRAD_TOL, k
    # so does a lens edited in place between the trace and the reduction
    tracer, specs, lens = problem()
    out = tracer.trace_rays(specs, lens)
    with torch.no_grad():
        lens.t.mul_(1.0)
    rt.compute_rms2d(out[0], out[1], out[4])
    assert calls == ['materialised', 'materialised']


def _patch_b_source():
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    section = text[text.index('## B. Function-level patch'):]
    return re.search(r'```python\n(.*?)```', section, flags=re.S).group(1)


def test_integration_patch_b_verbatim():
    """INTEGRATION.md section B, executed as written, against the unmodified reference."""
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip('oracle/_ref is not staged (python oracle/make_ref.py where /root/reference exists)')
    saved_path = list(sys.path)
    saved_modules = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('torchlens', 'shapely')}
    try:
        rtl, lm = make_ref.load_reference()
        assert 'oracle/_ref' in rtl.__file__
        originals = (rtl.trace_skew, rtl.compute_rms2d)
        exec(compile(_patch_b_source(), 'INTEGRATION.md#B', 'exec'), {'__name__': 'torchlens._b200'})
        assert rtl.trace_skew is not originals[0] and rtl.compute_rms2d is not originals[1]
        for name in ('cooke_8x8', 'tessar_16x16_epd2.0', 'cooke_8x8_aimed'):
            rec = load_golden(name)
            rec['name'] = name
            aimed = 'aimed' in name
            structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device=DEV)
            lens = lm.Lens(structure, *[torch.from_numpy(rec[k]).to(DEV).requires_grad_(True)
                                        for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
            specs = lm.Specs(structure, torch.from_numpy(rec['epd']).to(DEV), torch.from_numpy(rec['hfov']).to(DEV))
            # the REFERENCE's RayTracer (its own ray-set construction, ray aiming and pupil radius) on cuda
            tracer = rtl.RayTracer(mode='circular', n_rays=tuple(int(v) for v in rec['n_rays']),
                                   rel_fields=tuple(float(v) for v in rec['rel_fields']),
                                   wavelengths=tuple(float(v) for v in rec['wavelengths']),
                                   n_ray_aiming_iter=1 if aimed else 0, default_device=DEV)
            out = tracer.trace_rays(specs, lens)
            if not aimed:
                _check_outputs([o.detach() for o in out], rec)
            rms = rtl.compute_rms2d(out[0], out[1], out[4])
            _close_or_no_worse_than_reference(rms.item(), rec['rms'], rec['f64_rms'], RMS_TOL, f'patch B {name} rms')
            grads = torch.autograd.grad(rms, [lens.c, lens.t, lens.nd])
            for k, g in zip(('c', 't', 'nd'), grads):
                _close_or_no_worse_than_reference(g.cpu().numpy(), rec['grad_' + k], rec['f64_grad_' + k], GRAD_TOL,
                                                  f'patch B {name} d rms/d {k}')
        rtl.trace_skew, rtl.compute_rms2d = originals
    finally:
        sys.path[:] = saved_path
        for k in [m for m in sys.modules if m.split('.')[0] in ('torchlens', 'shapely')]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
