"""Generate the golden vectors for the hot path by RUNNING the unmodified reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or golden files, so the pins are
outputs of its own functions ``RayTracer.trace_rays`` -> ``trace_skew`` ->
``compute_rms2d`` (ray_tracing_lite.py:80-127, :594-675, :678-702) and the autograd
gradients of the RMS w.r.t. the prescription, on the four lenses it ships
(``torchlens/data/*.yml``).  ``shapely`` is imported but unused by the reference
(ray_tracing_lite.py:15), so a dummy module stands in for it.

Every case stores the exact tensors handed to ``trace_skew`` (captured by wrapping
it), its six outputs, the RMS and the gradients; the GPU box never sees
/root/reference, only these ``.npz`` files.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'


def import_reference():
    shapely = types.ModuleType('shapely')
    geometry = types.ModuleType('shapely.geometry')
    geometry.Polygon = object
    shapely.geometry = geometry
    sys.modules.setdefault('shapely', shapely)
    sys.modules.setdefault('shapely.geometry', geometry)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torchlens.ray_tracing_lite as rtl
    import torchlens.lens_modeling as lm
    return rtl, lm


def load_lens(lm, path, dtype=torch.float32):
    with open(path) as fh:
        d = yaml.safe_load(fh)
    structure = lm.Structure(np.array(d['stop_idx']), sequence=np.array(d['sequence']),
                             default_device='cpu')
    lens = lm.Lens(structure,
                   torch.tensor(d['c'], dtype=dtype), torch.tensor(d['t'], dtype=dtype),
                   torch.tensor(d['nd'], dtype=dtype), torch.tensor(d['v'], dtype=dtype))
    return d, structure, lens


F64_KEYS = ('rms', 'grad_c', 'grad_t', 'grad_nd', 'grad_in_z', 'grad_in_c', 'grad_in_t', 'grad_in_mu',
            'in_x', 'in_y')


def run_case(rtl, lm, yml, n_rays, rel_fields, wavelengths, epd_scale=1.0, n_ray_aiming_iter=0,
             allow_backward_rays=True, mode='circular', vig=None, ray_aiming_mode='real'):
    """One golden record: the reference in fp32 (every key) and, under the prefix ``f64_``, the RMS,
    the gradients and the (possibly ray-aimed) pupil of the SAME reference code run in float64 --
    the yardstick for "no farther from the truth than the reference's own fp32"."""
    rec = _run_case(rtl, lm, yml, n_rays, rel_fields, wavelengths, epd_scale, n_ray_aiming_iter,
                    allow_backward_rays, mode, torch.float32, vig, ray_aiming_mode)
    torch.set_default_dtype(torch.float64)
    try:
        rec64 = _run_case(rtl, lm, yml, n_rays, rel_fields, wavelengths, epd_scale, n_ray_aiming_iter,
                          allow_backward_rays, mode, torch.float64, vig, ray_aiming_mode)
    finally:
        torch.set_default_dtype(torch.float32)
    assert np.array_equal(rec['out_ok'], rec64['out_ok']), 'fp32 and fp64 masks of the reference differ'
    for k in F64_KEYS + (('out_x', 'out_y') if n_ray_aiming_iter > 0 else ()):
        rec['f64_' + k] = rec64[k]
    if vig is not None:
        rec['vig'] = np.asarray(vig, dtype=np.float32)
    rec['ray_aiming_mode'] = np.asarray(ray_aiming_mode)
    return rec


def linear_vignetting(fields, vig):
    """The vignetting function of the golden cases: factor growing linearly with the relative field,
    [1,F] x [B] -> [B,F] (RayTracer calls ``vig_fn(fields, specs.vig_up)``, rtl:100-102)."""
    return vig[:, None] * fields


def restore_commented_out_helpers(rtl):
    """ray_tracing_lite.py calls two helpers whose definitions it only carries as comments -- and spells
    one dtype ``tf.float32`` -- so its own vignetting and 'paraxial' ray-aiming branches raise
    NameError (rtl:98-104, :138-140, :154-160).  To let the reference's control
    flow produce goldens for those branches, the two helpers are supplied as literal torch
    transcriptions of the commented text (rtl:483-494 = rt_tf:479-490, rtl:797-809 = rt_tf:765-777; both
    pinned separately in tests/test_tf_original_functions.py) and ``tf`` is bound to torch."""
    import torch as _torch

    def apply_vignetting(y, vig_up, vig_down):
        trailing = [1] * (len(y.shape) - len(vig_down.shape))
        vig_up = _torch.reshape(vig_up, (*vig_up.shape, *trailing))
        vig_down = _torch.reshape(vig_down, (*vig_down.shape, *trailing))
        scale = 1 - (vig_up + vig_down) / 2
        offset = (vig_down - vig_up) / 2
        return y * scale + offset

    def compute_magnification(lens):
        nd = _torch.cat((_torch.ones_like(lens.nd[:, 0:1]), lens.nd), dim=1)
        abcd = rtl.reduce_abcd(rtl.interface_propagation_abcd(lens.c, lens.t, nd))
        return abcd[:, 0, 0]

    rtl.apply_vignetting = apply_vignetting
    rtl.compute_magnification = compute_magnification
    rtl.tf = _torch


def _run_case(rtl, lm, yml, n_rays, rel_fields, wavelengths, epd_scale, n_ray_aiming_iter,
              allow_backward_rays, mode, dtype, vig=None, ray_aiming_mode='real'):
    d, structure, lens = load_lens(lm, yml, dtype)
    for name in ('c', 't', 'nd', 'v'):
        getattr(lens, name).requires_grad_(True)
    efl = lens.efl.detach()
    epd = efl / torch.tensor(d['f_number'], dtype=dtype) * epd_scale
    hfov = torch.deg2rad(torch.tensor(d['hfov'], dtype=dtype))
    if vig is None:
        specs = lm.Specs(structure, epd, hfov)
    else:
        specs = lm.Specs(structure, epd, hfov, *(torch.tensor([v], dtype=dtype) for v in vig))

    captured = {}
    real_trace = rtl.trace_skew

    def spy(x, y, z, cx, cy, c, t, mu, mask, aggregate=False, allow_backward_rays=True):
        captured['last'] = dict(x=x, y=y, z=z, cx=cx, cy=cy, c=c, t=t, mu=mu, mask=mask)
        return real_trace(x, y, z, cx, cy, c, t, mu, mask, aggregate, allow_backward_rays)

    rtl.trace_skew = spy
    try:
        tracer = rtl.RayTracer(mode=mode, n_rays=n_rays, rel_fields=rel_fields,
                               wavelengths=wavelengths, n_ray_aiming_iter=n_ray_aiming_iter,
                               allow_backward_rays=allow_backward_rays, default_device='cpu',
                               vig_fn=None if vig is None else linear_vignetting, ray_aiming_mode=ray_aiming_mode)
        x, y, cx, cy, ok, bw = tracer.trace_rays(specs, lens)
    finally:
        rtl.trace_skew = real_trace
    rms = rtl.compute_rms2d(x, y, ok)
    grads = torch.autograd.grad(rms, [lens.c, lens.t, lens.nd, lens.v], allow_unused=True)
    ins = captured['last']
    # gradient of the rms w.r.t. the tensors trace_skew itself receives
    rec = {k: v.detach().clone() for k, v in ins.items()}
    for k in ('z', 'c', 't', 'mu'):
        rec[k].requires_grad_(True)
    out2 = real_trace(rec['x'], rec['y'], rec['z'], rec['cx'], rec['cy'], rec['c'], rec['t'],
                      rec['mu'], rec['mask'], False, allow_backward_rays)
    rms2 = rtl.compute_rms2d(out2[0], out2[1], out2[4])
    g_z, g_c, g_t, g_mu = torch.autograd.grad(rms2, [rec['z'], rec['c'], rec['t'], rec['mu']])
    out = dict(
        efl=efl.numpy(), bfl=lens.bfl.detach().numpy(), epd=epd.numpy(), hfov=hfov.numpy(),
        rel_fields=np.asarray(rel_fields, dtype=np.float32),
        wavelengths=np.asarray(tracer.wavelengths, dtype=np.float64),
        n_rays=np.asarray(n_rays), allow_backward_rays=np.asarray(allow_backward_rays),
        lens_c=lens.c.detach().numpy(), lens_t=lens.t.detach().numpy(),
        lens_nd=lens.nd.detach().numpy(), lens_v=lens.v.detach().numpy(),
        stop_idx=np.asarray(d['stop_idx']), sequence=np.asarray(d['sequence']),
        in_x=ins['x'].detach().numpy(), in_y=ins['y'].detach().numpy(),
        in_z=ins['z'].detach().numpy(), in_cx=ins['cx'].detach().numpy(),
        in_cy=ins['cy'].detach().numpy(), in_c=ins['c'].detach().numpy(),
        in_t=ins['t'].detach().numpy(), in_mu=ins['mu'].detach().numpy(),
        in_mask=ins['mask'].numpy(),
        out_x=x.detach().numpy(), out_y=y.detach().numpy(), out_cx=cx.detach().numpy(),
        out_cy=cy.detach().numpy(), out_ok=ok.numpy(), out_backward=bw.numpy(),
        rms=rms.detach().numpy(),
        grad_c=grads[0].numpy(), grad_t=grads[1].numpy(), grad_nd=grads[2].numpy(),
        grad_v=np.nan_to_num(grads[3].numpy()),
        grad_in_z=g_z.numpy(), grad_in_c=g_c.numpy(), grad_in_t=g_t.numpy(), grad_in_mu=g_mu.numpy(),
    )
    return out


def export_lens_yaml(src, dst):
    """Re-emit a reference prescription (data, same schema) into the package."""
    with open(src) as fh:
        d = yaml.safe_load(fh)
    with open(dst, 'w') as fh:
        fh.write('# prescription data in the reference schema (stop_idx, sequence, hfov [deg],\n'
                 '# f_number, c, t, nd, v); values from the reference fixture of the same name\n')
        yaml.safe_dump(d, fh, default_flow_style=None, sort_keys=False)


def main():
    rtl, lm = import_reference()
    torch.manual_seed(0)
    data = os.path.join(REF_ROOT, 'torchlens', 'data')
    lens_dir = os.path.join(HERE, '..', '..', 'torchoptics_b200', 'lenses')
    os.makedirs(lens_dir, exist_ok=True)
    lenses = {'singlet': 'singlet_lens.yml', 'doublet': 'baseline_doublet.yml',
              'cooke': 'baseline_cooke.yml', 'tessar': 'baseline_tessar.yml'}
    for name, fn in lenses.items():
        export_lens_yaml(os.path.join(data, fn), os.path.join(lens_dir, fn))
    std = dict(n_rays=(8, 8), rel_fields=(0., 0.707, 1.), wavelengths=('C', 'd', 'F'))
    cases = {}
    for name, fn in lenses.items():
        cases[f'{name}_8x8'] = run_case(rtl, lm, os.path.join(data, fn), **std)
    cooke = os.path.join(data, lenses['cooke'])
    tessar = os.path.join(data, lenses['tessar'])
    big = dict(n_rays=(16, 16), rel_fields=(0., 0.707, 1.), wavelengths=('C', 'd', 'F'))
    for scale in (2.0, 2.6, 3.2):
        cases[f'cooke_16x16_epd{scale}'] = run_case(rtl, lm, cooke, epd_scale=scale, **big)
    cases['cooke_16x16_epd2.6_nobackward'] = run_case(rtl, lm, cooke, epd_scale=2.6,
                                                      allow_backward_rays=False, **big)
    cases['tessar_16x16_epd2.0'] = run_case(rtl, lm, tessar, epd_scale=2.0, **big)
    cases['cooke_8x8_aimed'] = run_case(rtl, lm, cooke, n_ray_aiming_iter=1, **std)
    cases['tessar_8x8_aimed'] = run_case(rtl, lm, tessar, n_ray_aiming_iter=1, **std)  # >1 iteration raises in the reference (rtl:170)
    cases['cooke_32x32'] = run_case(rtl, lm, cooke, n_rays=(32, 32), rel_fields=(0., 0.5, 0.707, 1.),
                                    wavelengths=(656.3, 587.6, 546.1, 486.1))
    # BASELINE.json configs[0] at its exact shape: Cooke triplet, 3 fields x 3 wavelengths, 96 x 76
    # pupil = 65 664 rays
    cases['cooke_96x76_config1'] = run_case(rtl, lm, cooke, n_rays=(96, 76), **{k: std[k] for k in ('rel_fields', 'wavelengths')})
    # pupil vignetting and the 'paraxial' stop radius: branches the lite reference cannot run as shipped
    # (see restore_commented_out_helpers); vig = (vig_up, vig_down, vig_x) at full field
    restore_commented_out_helpers(rtl)
    vig = (0.30, 0.15, 0.10)
    cases['cooke_8x8_vig'] = run_case(rtl, lm, cooke, vig=vig, **std)
    cases['cooke_8x8_vig_aimed'] = run_case(rtl, lm, cooke, vig=vig, n_ray_aiming_iter=1, **std)
    cases['tessar_8x8_vig_aimed'] = run_case(rtl, lm, tessar, vig=vig, n_ray_aiming_iter=1, **std)
    cases['cooke_8x8_paraxial_aimed'] = run_case(rtl, lm, cooke, n_ray_aiming_iter=1, ray_aiming_mode='paraxial', **std)
    cases['tessar_8x8_vig_paraxial_aimed'] = run_case(rtl, lm, tessar, vig=vig, n_ray_aiming_iter=1,
                                                      ray_aiming_mode='paraxial', **std)
    for name, rec in cases.items():
        np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **rec)
        print(f"{name:34s} S={rec['in_t'].shape[-1]} rays={rec['out_ok'].size:6d} "
              f"ok={int(rec['out_ok'].sum()):6d} bw={int(rec['out_backward'].sum()):5d} "
              f"rms={float(rec['rms']):.8f} efl={float(rec['efl'][0]):.6f}")


if __name__ == '__main__':
    main()
