"""Parity on the BASELINE.json configurations themselves ("weak" items 1-2):

* config 1 at its exact shape -- Cooke triplet, 3 fields x 3 wavelengths, 96 x 76 pupil = 65 664
  rays -- against the golden record the unmodified reference produced for it
  (tests/golden/cooke_96x76_config1.npz): masks bit-exact, points / cosines 1e-5, RMS 1e-5,
  gradients 1e-4, every entry point (trace_skew both policies, split backward, fused pass,
  RayTracer front end, graphed step);
* config 2 (Double-Gauss S=11, 16 fields x 3 wavelengths) -- the benchmarked configuration -- the
  GRADIENTS of the fused pass against autograd of the oracle evaluated on the device, at ~1 M rays
  and at the full 296^2 pupil (4.2 M rays), next to the oracle's own fp32-vs-fp64 distance.

Tolerances are north_star's (written below); where one cannot hold because the reference's fp32
value is itself farther from the truth, the bar is the float64 run of the same code, printed.
"""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden
from tests.test_gpu_parity import (DEV, GRAD_TOL, POINT_TOL, COS_TOL, RMS_TOL, _args, _check_outputs,
                                   _close_or_no_worse_than_reference, _double_gauss_problem, _inputs, _rel,
                                   oracle_lens_gradients)
from torchoptics_b200 import lens_modeling as lm
from torchoptics_b200 import ops
from torchoptics_b200 import ray_tracing_lite as rt

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------
# config 1
# ---------------------------------------------------------------------------
@pytest.fixture(scope='module')
def config1():
    rec = load_golden('cooke_96x76_config1')
    rec['name'] = 'cooke_96x76_config1'
    assert rec['out_ok'].shape == (1, 3, 96 * 76, 3)
    return rec


def _config1_problem(rec, requires_grad=True):
    structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device=DEV)
    lens = lm.Lens(structure, *[torch.from_numpy(rec[k]).to(DEV).requires_grad_(requires_grad)
                                for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    specs = lm.Specs(structure, torch.from_numpy(rec['epd']).to(DEV), torch.from_numpy(rec['hfov']).to(DEV))
    tracer = rt.RayTracer(mode='circular', n_rays=(96, 76), rel_fields=(0., 0.707, 1.),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    return tracer, specs, lens


def test_config1_trace_and_masks(config1):
    """trace_skew on the very tensors the reference handed to its own trace_skew."""
    i = _inputs(config1, DEV)
    for arith in ('guarded', 'exact'):
        out = rt.trace_skew(*_args(i), arith=arith)
        _check_outputs(out, config1)
    assert int(config1['out_backward'].sum()) == 1044 and bool(config1['out_ok'].all())


def test_config1_ray_set_is_the_references(config1):
    """RayTracer builds the same ray set (pupil grid, pupil position, mu) as the reference did."""
    tracer, specs, lens = _config1_problem(config1, requires_grad=False)
    args = tracer._ray_set(specs, lens)
    for key, got in zip(('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu'), args):
        want = config1['in_' + key]
        assert tuple(got.shape) == want.shape, key
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), key


def test_config1_loss_and_gradients_every_entry_point(config1):
    want_rms = float(config1['rms'])
    # (a) drop-in sequence through the front end: trace_rays -> compute_rms2d -> backward
    tracer, specs, lens = _config1_problem(config1)
    out = tracer.trace_rays(specs, lens)
    _check_outputs([o.detach() for o in out], config1)
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    assert abs(rms.item() - want_rms) <= RMS_TOL * want_rms
    grads = torch.autograd.grad(rms, [lens.c, lens.t, lens.nd])
    for name, g in zip(('c', 't', 'nd'), grads):
        _close_or_no_worse_than_reference(g.cpu().numpy(), config1['grad_' + name], config1['f64_grad_' + name],
                                          GRAD_TOL, f'config 1 drop-in d rms/d {name}')
    # (b) fused pass (the headline kernel), staged and unstaged front end
    for staged in (True, False):
        tracer, specs, lens = _config1_problem(config1)
        rms_f, _ = tracer.spot_rms(specs, lens, staged=staged)
        assert abs(rms_f[0].item() - want_rms) <= RMS_TOL * want_rms, staged
        grads_f = torch.autograd.grad(rms_f[0], [lens.c, lens.t, lens.nd])
        for name, g in zip(('c', 't', 'nd'), grads_f):
            _close_or_no_worse_than_reference(g.cpu().numpy(), config1['grad_' + name],
                                              config1['f64_grad_' + name], GRAD_TOL,
                                              f'config 1 fused (staged={staged}) d rms/d {name}')
    # (c) fused pass on the reference's own trace_skew inputs: gradients w.r.t. z, c, t, mu
    gpu = _inputs(config1, DEV, grad=('z', 'c', 't', 'mu'))
    rms_i, _ = ops.spot_rms(*_args(gpu))
    got = torch.autograd.grad(rms_i[0], [gpu[k] for k in ('z', 'c', 't', 'mu')])
    for name, g in zip(('z', 'c', 't', 'mu'), got):
        scale = max(float(np.abs(config1['f64_grad_in_z']).max()), float(np.abs(config1['f64_grad_in_t']).max())) \
            if name == 'z' else None
        _close_or_no_worse_than_reference(g.cpu().numpy(), config1['grad_in_' + name],
                                          config1['f64_grad_in_' + name], GRAD_TOL,
                                          f'config 1 fused d rms/d {name} (trace_skew inputs)', group_scale=scale)
    # (d) the graphed end-to-end step (host prescription in, host gradients out)
    from torchoptics_b200 import GraphedSpotStep
    tracer, specs, lens = _config1_problem(config1, requires_grad=False)
    step = GraphedSpotStep(tracer, specs, lens)
    host_rms, host_grads = step()
    assert abs(float(host_rms[0]) - want_rms) <= RMS_TOL * want_rms
    for name in ('c', 't', 'nd'):
        _close_or_no_worse_than_reference(host_grads[name].numpy(), config1['grad_' + name],
                                          config1['f64_grad_' + name], GRAD_TOL, f'config 1 graphed d rms/d {name}')


# ---------------------------------------------------------------------------
# config 2: gradients against the oracle on the device
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('n_side', [148, 296])
def test_config2_gradients_against_oracle_on_device(n_side):
    """Double-Gauss S=11, 16 fields x 3 wavelengths x n_side^2 pupil (1.05 M and the benchmarked
    4.2 M rays): RMS and d rms / d(c, t, nd) of the fused pass -- the kernel bench.py times --
    against autograd of the oracle on the same device in fp32 (north_star: 1e-5 / 1e-4), with the
    oracle's float64 run as the yardstick; and the drop-in sequence likewise."""
    tracer, specs, lens = _double_gauss_problem(n_side, requires_grad=True)
    ref_rms, ref_g, ref_out = oracle_lens_gradients(tracer, specs, lens, torch.float32)
    assert bool(ref_out[4].all())
    del ref_out
    torch.cuda.empty_cache()
    rms64, g64, _ = oracle_lens_gradients(tracer, specs, lens, torch.float64)
    torch.cuda.empty_cache()
    for staged in (True, False):
        for k in ('c', 't', 'nd'):
            getattr(lens, k).grad = None
        rms_f, _ = tracer.spot_rms(specs, lens, staged=staged)
        _close_or_no_worse_than_reference(rms_f[0].item(), ref_rms.item(), rms64.item(), RMS_TOL,
                                          f'config 2 ({n_side}^2, staged={staged}) rms')
        got = torch.autograd.grad(rms_f[0], [lens.c, lens.t, lens.nd])
        for name, g, r, r64 in zip(('c', 't', 'nd'), got, ref_g, g64):
            _close_or_no_worse_than_reference(g.cpu().numpy(), r.cpu().numpy(), r64.cpu().numpy(), GRAD_TOL,
                                              f'config 2 ({n_side}^2, staged={staged}) fused d rms/d {name}')
    # the unfused drop-in sequence on the same problem
    out = tracer.trace_rays(specs, lens)
    rms_s = rt.compute_rms2d(out[0], out[1], out[4])
    got = torch.autograd.grad(rms_s, [lens.c, lens.t, lens.nd])
    for name, g, r, r64 in zip(('c', 't', 'nd'), got, ref_g, g64):
        _close_or_no_worse_than_reference(g.cpu().numpy(), r.cpu().numpy(), r64.cpu().numpy(), GRAD_TOL,
                                          f'config 2 ({n_side}^2) drop-in d rms/d {name}')


def _equator_bundle(n=256, dtype=torch.float64, device=DEV):
    """A ball-like front surface (R = 2 mm) hit from outside by steep rays on its FAR half: the near
    root of the intersection lies beyond the equator (z_hit > R, c h_z > 1), the index step is weak,
    so the rays run on through a flat second surface to the image plane, all of them ok, none
    flagged backward, every predicate far from its threshold (fast path)."""
    c = torch.tensor([0.5, 0.0], dtype=dtype, device=device).reshape(1, 1, 1, 1, 2)
    t = torch.tensor([3.0, 5.0], dtype=dtype, device=device).reshape(1, 1, 1, 1, 2)
    mu = torch.tensor([1.0 / 1.05, 1.05], dtype=dtype, device=device).reshape(1, 1, 1, 1, 2)
    mask = torch.ones((1, 1, 1, 1, 2), dtype=torch.bool, device=device)
    y = torch.linspace(-7.0, -6.0, n, dtype=dtype, device=device).reshape(1, 1, n, 1)
    x = torch.full_like(y, 0.3)
    z = torch.full((1, 1, 1, 1), -1.0, dtype=dtype, device=device)
    cx = torch.zeros((1, 1, 1, 1), dtype=dtype, device=device)
    cy = torch.full((1, 1, 1, 1), 0.8, dtype=dtype, device=device)
    # first intersection with the sphere of centre (0, 0, 2), radius 2
    cz = (1 - cy ** 2).sqrt()
    ox, oy, oz = x, y, z - 2.0
    b = ox * cx + oy * cy + oz * cz
    disc = b * b - (ox * ox + oy * oy + oz * oz - 4.0)
    z_hit = z + (-b - disc.sqrt()) * cz
    assert bool((disc > 0).all()) and bool((z_hit > 2.0).all()), 'every ray must hit beyond the equator'
    return [x, y, z, cx, cy, c, t, mu], mask


def test_equator_hit_that_reaches_the_image_has_the_right_gradient():
    """VERDICT r1 "weak" 3: rays that hit a sphere BEYOND ITS EQUATOR (c h_z > 1) and still reach the
    image.  Round 1's sweep rebuilt n_z = +sqrt(1 - c^2 rho) and gave such a ray the adjoint of the
    wrong sag branch; the reversible sweep of the fused pass carries the true h_z (and the split
    backward carries the branch bit).  Against autograd of the oracle in float64 on the device."""
    from oracle import trace_oracle as oracle
    args64, mask = _equator_bundle()
    for v in args64[5:]:
        v.requires_grad_(True)
    ref = oracle.trace(*args64, mask)
    assert bool(ref[4].all()) and not bool(ref[5].any())
    rms64 = oracle.spot_rms_all_lenses(ref[1], ref[4])[0]
    want = torch.autograd.grad(rms64, args64[5:])
    # the oracle's own fp32 run: the yardstick for what fp32 can deliver on rays this oblique
    args32 = [v.detach().float() for v in args64]
    for v in args32[5:]:
        v.requires_grad_(True)
    ref32 = oracle.trace(*args32, mask)
    rms32 = oracle.spot_rms_all_lenses(ref32[1], ref32[4])[0]
    want32 = torch.autograd.grad(rms32, args32[5:])
    leaves = [v.detach().float().requires_grad_(True) for v in args64[5:]]
    rms_f, _ = ops.spot_rms(*[v.detach().float() for v in args64[:5]], *leaves, mask)
    _close_or_no_worse_than_reference(rms_f[0].item(), rms32.item(), rms64.item(), RMS_TOL, 'equator bundle rms')
    got = torch.autograd.grad(rms_f[0], leaves)
    for name, a, b32, b64 in zip(('c', 't', 'mu'), got, want32, want):
        _close_or_no_worse_than_reference(a.cpu().numpy(), b32.cpu().numpy(), b64.cpu().numpy(), GRAD_TOL,
                                          f'equator bundle d rms/d {name}')
        # and in any case nowhere near the wrong-branch gradient (round 1 was off by O(1) here)
        assert _rel(a.cpu().numpy(), b64.cpu().numpy()) <= 1e-3, name
