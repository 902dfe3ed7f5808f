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

from tests.conftest import GOLDEN_DIR, load_golden
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


def _reference_front_end(monkeypatch):
    """The reference's OWN front-end files (optics_simulator_lite.py, optical_loss.py: staged unmodified under
    oracle/_ref) imported so that every hot-path name they use -- `torchlens.ray_tracing_lite`, `torchlens.
    lens_modeling` (osl:7-8) and the top-level `lens_modeling`, `ray_tracing_lite` (ol:7-8) -- binds to THIS
    repo's modules.  Stand-ins only for imports the image lacks (as in tests/golden/make_golden_optical_loss.py)."""
    import importlib.util
    import types
    from oracle import make_ref
    from torchoptics_b200 import lens_modeling, ray_tracing_lite
    import torchlens      # the alias package of this repo
    assert sys.modules['torchlens.ray_tracing_lite'] is ray_tracing_lite
    stubs = {name: types.ModuleType(name) for name in
             ('matplotlib', 'matplotlib.pyplot', 'utils', 'utils.w2rgb', 'preprocessing', 'preprocessing.process_dataframe')}
    stubs['matplotlib'].pyplot = stubs['matplotlib.pyplot']
    stubs['utils'].w2rgb = stubs['utils.w2rgb']
    stubs['utils.w2rgb'].wavelength_to_rgb = lambda *a, **k: (0, 0, 0)
    stubs['preprocessing'].process_dataframe = stubs['preprocessing.process_dataframe']
    from torchoptics_b200.optical_loss import sequence_decoder, sequence_encoder
    stubs['preprocessing.process_dataframe'].sequence_encoder = sequence_encoder
    stubs['preprocessing.process_dataframe'].sequence_decoder = sequence_decoder
    for name, module in {**stubs, 'lens_modeling': lens_modeling, 'ray_tracing_lite': ray_tracing_lite}.items():
        monkeypatch.setitem(sys.modules, name, module)
    real_loadtxt = np.loadtxt
    monkeypatch.setattr(np, 'loadtxt', lambda path, *a, **k: np.asarray([[1.5168, 64.17], [1.7847, 25.68]], np.float32)
                        if str(path).endswith('selected_ohara_glass.csv') else real_loadtxt(path, *a, **k))
    loaded = {}
    for name in ('optics_simulator_lite', 'optical_loss'):      # (osl first: optical_loss imports it by that name)
        path = os.path.join(make_ref.DEST, 'torchlens', name + '.py')
        spec = importlib.util.spec_from_file_location(name, path)
        module = importlib.util.module_from_spec(spec)
        monkeypatch.setitem(sys.modules, name, module)
        spec.loader.exec_module(module)
        loaded[name] = module
    assert loaded['optics_simulator_lite'].rt is ray_tracing_lite and loaded['optical_loss'].Structure is lens_modeling.Structure
    return loaded['optics_simulator_lite'], loaded['optical_loss']


@pytest.mark.parametrize('lens_type', ['GA', 'GAGA'])
def test_reference_front_end_runs_unchanged_on_the_cuda_path(lens_type, monkeypatch):
    """`RaytracedOptics.do_ray_tracing` (optics_simulator_lite.py:456-493) and the loop body of
    `Optical_Loss` (optical_loss.py:20-96) -- the reference's own files, unmodified -- on top of this repo's
    RayTracer / Lens / compute_last_curvature / compute_rms2d on `cuda`: same numbers as the reference end to end
    (golden records of tests/golden/make_golden_optical_loss.py)."""
    from oracle import make_ref
    if not make_ref.available() or not os.path.exists(os.path.join(make_ref.DEST, 'torchlens', 'optical_loss.py')):
        pytest.skip('oracle/_ref is not staged (python oracle/make_ref.py where /root/reference exists)')
    osl, ol = _reference_front_end(monkeypatch)
    with np.load(os.path.join(GOLDEN_DIR, 'optical_loss', lens_type + '.npz')) as z:
        g = {k: z[k] for k in z.files}
    loss_fn = ol.Optical_Loss(lens_type)                       # the REFERENCE's class
    for i in (0, 1, 4):
        x = torch.from_numpy(g['inputs'][i]).to(DEV)
        y = torch.from_numpy(g['outputs'][i]).to(DEV).requires_grad_(True)
        loss, rms, penalty = loss_fn.optical_loss_unsupervised_single(x, y, 0.2, device=DEV)
        assert abs(float(rms) - g['per_sample_rms'][i]) <= 5e-5 * g['per_sample_rms'][i], (i, float(rms))
        assert abs(float(penalty) - g['per_sample_penalty'][i]) <= 2e-5 * g['per_sample_penalty'][i]
        assert abs(float(loss) - g['per_sample_loss'][i]) <= 2e-5 * g['per_sample_loss'][i]
        grad, = torch.autograd.grad(loss, y)
        got, ref32, ref64 = grad.cpu().numpy().astype(np.float64), g['grad_outputs'][i].astype(np.float64), g['f64_grad_outputs'][i]
        assert np.isfinite(got).all()
        scale = np.abs(ref64).max()
        ours, theirs = np.abs(got - ref64).max() / scale, np.abs(ref32 - ref64).max() / scale
        print(f'{lens_type}[{i}] d loss / d output: ours-vs-fp64 {ours:.2e}, reference-fp32-vs-fp64 {theirs:.2e}')
        assert ours <= 1e-4 or ours <= 3 * theirs, (ours, theirs)
