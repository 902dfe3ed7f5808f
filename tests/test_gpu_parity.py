"""Parity of the CUDA hot path (through the C ABI) with the oracle and with the
golden records of the reference.  Needs a B200: ``pytest -m gpu``.

Tolerances (BASELINE.json north_star / SURVEY.md section 8c): masks bit-exact;
points max|d(x,y)| / max|(x,y)| <= 1e-5; direction cosines max|d| <= 1e-5;
gradients ||dg|| / ||g|| <= 1e-4 per parameter group; RMS 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import trace_oracle as oracle
from tests.conftest import load_golden
from torchoptics_b200 import lens_modeling as lm
from torchoptics_b200 import ops, prescriptions
from torchoptics_b200 import ray_tracing_lite as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
POINT_TOL = 1e-5
COS_TOL = 1e-5
GRAD_TOL = 1e-4
RMS_TOL = 1e-5


def _inputs(rec, device, grad=()):
    t = {k[3:]: torch.from_numpy(rec[k]).to(device) for k in rec if k.startswith('in_')}
    for k in grad:
        t[k] = t[k].clone().requires_grad_(True)
    return t


def _args(i):
    return [i[k] for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'mask')]


def _rel(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)


def _close_or_no_worse_than_reference(got, ref32, ref64, tol, what, group_scale=None):
    """The north_star tolerance against the reference's fp32 value -- or, where that value is itself
    farther than `tol` from the truth, no farther from the reference run in float64 than the
    reference's own fp32 is.  Both distances are printed (pytest -s / on failure).
    `group_scale`: for a single heavily cancelled scalar (d rms / d z is ~1e-3 of d rms / d t, both
    axial shifts; its fp32 value is noise at the 5e-4 level on EITHER side, so "no farther than the
    reference" is a coin flip) the absolute error may instead be within `tol` of the scale of its
    parameter group."""
    got, ref32, ref64 = (np.nan_to_num(np.asarray(v, dtype=np.float64)) for v in (got, ref32, ref64))
    vs32, ours, theirs = _rel(got, ref32), _rel(got, ref64), _rel(ref32, ref64)
    print(f'{what}: ours-vs-reference-fp32 {vs32:.2e}, ours-vs-fp64 {ours:.2e}, reference-fp32-vs-fp64 {theirs:.2e}')
    if group_scale is not None and np.abs(got - ref64).max() <= tol * group_scale:
        return
    assert vs32 <= tol or ours <= theirs, (what, vs32, ours, theirs)


def oracle_lens_gradients(tracer, specs, lens, dtype=torch.float32):
    """(rms, [d rms/d c, d rms/d t, d rms/d nd]) of lens 0 from the ORACLE evaluated by torch on the
    device in `dtype` (autograd through the ray-set construction, the trace and the RMS)."""
    leaves = [getattr(lens, k).detach().to(dtype).clone().requires_grad_(True) for k in ('c', 't', 'nd')]
    lens_d = lm.Lens(lens.structure, *leaves, lens.v.detach().to(dtype))
    specs_d = lm.Specs(specs.structure, specs.epd.detach().to(dtype), specs.hfov.detach().to(dtype))
    out = oracle.trace(*tracer._ray_set(specs_d, lens_d))
    rms = oracle.spot_rms_all_lenses(out[1], out[4])[0]
    return rms.detach(), torch.autograd.grad(rms, leaves), out


def _check_outputs(out, rec, exact_ref=None):
    ok_shape = rec['out_ok'].shape
    scale = max(np.abs(rec['out_x']).max(), np.abs(rec['out_y']).max())
    x, y, cx, cy, ok, bw = [o.cpu().numpy() for o in out]
    assert np.array_equal(ok, rec['out_ok'])
    assert np.array_equal(bw, np.broadcast_to(rec['out_backward'], ok_shape))
    assert np.abs(x - rec['out_x']).max() <= POINT_TOL * scale
    assert np.abs(y - rec['out_y']).max() <= POINT_TOL * scale
    assert np.abs(cx - rec['out_cx']).max() <= COS_TOL
    assert np.abs(cy - rec['out_cy']).max() <= COS_TOL
    # failed rays are parked: all four outputs exactly zero
    dead = ~rec['out_ok']
    assert not x[dead].any() and not y[dead].any() and not cx[dead].any() and not cy[dead].any()
    if exact_ref is not None:
        for got, want in zip((x, y, cx, cy), exact_ref[:4]):
            want = np.broadcast_to(want.numpy(), ok_shape)
            assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))


def test_forward_exact_policy_bit_identical(golden):
    i = _inputs(golden, DEV)
    allow = bool(golden['allow_backward_rays'])
    out = rt.trace_skew(*_args(i), allow_backward_rays=allow, arith='exact')
    cpu = _inputs(golden, 'cpu')
    with oracle.ieee_sqrt():
        ref = oracle.trace(*_args(cpu), False, allow)
    _check_outputs(out, golden, exact_ref=ref)


def test_forward_guarded_policy(golden):
    i = _inputs(golden, DEV)
    out = rt.trace_skew(*_args(i), allow_backward_rays=bool(golden['allow_backward_rays']))
    _check_outputs(out, golden)


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
def test_forward_long_row_kernel_on_every_golden(golden, arith, monkeypatch):
    """The golden pupils are short (64 .. 1 024 points) and take the flattened forward kernel; with TL_NO_ROWS the
    same calls take the long-row kernel (k_spot_rev<out16>, the one BASELINE-size traces use): same checks -- masks,
    parked failed rays, tolerances, and the exact policy bit for bit -- misses, TIR and backward rays included."""
    monkeypatch.setenv('TL_NO_ROWS', '1')
    i = _inputs(golden, DEV)
    allow = bool(golden['allow_backward_rays'])
    out = rt.trace_skew(*_args(i), allow_backward_rays=allow, arith=arith)
    ref = None
    if arith == 'exact':
        cpu = _inputs(golden, 'cpu')
        with oracle.ieee_sqrt():
            ref = oracle.trace(*_args(cpu), False, allow)
    _check_outputs(out, golden, exact_ref=ref)


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
def test_backward_against_oracle_autograd(golden, arith):
    allow = bool(golden['allow_backward_rays'])
    gen = torch.Generator().manual_seed(1)
    shape = golden['out_ok'].shape
    # positive seeds: the summed gradients are then well conditioned (with random signs the
    # per-lens sums cancel to ~1e-4 of their terms and fp32 round-off of either side dominates)
    seeds = [0.5 + torch.rand(shape, generator=gen) for _ in range(4)]
    wrt = ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu')
    cpu = _inputs(golden, 'cpu', grad=wrt)
    ref_out = oracle.trace(*_args(cpu), False, allow)
    ref_loss = sum((s * o).sum() for s, o in zip(seeds, ref_out[:4]))
    ref = torch.autograd.grad(ref_loss, [cpu[k] for k in wrt])
    gpu = _inputs(golden, DEV, grad=wrt)
    out = rt.trace_skew(*_args(gpu), allow_backward_rays=allow, arith=arith)
    loss = sum((s.to(DEV) * o).sum() for s, o in zip(seeds, out[:4]))
    got = torch.autograd.grad(loss, [gpu[k] for k in wrt])
    for name, g, r in zip(wrt, got, ref):
        assert g.shape == r.shape, name
        assert _rel(g.cpu().numpy(), r.numpy()) <= GRAD_TOL, (name, _rel(g.cpu().numpy(), r.numpy()))


def test_rms_from_rays(golden):
    y = torch.from_numpy(golden['out_y']).to(DEV).requires_grad_(True)
    ok = torch.from_numpy(golden['out_ok']).to(DEV)
    rms = rt.compute_rms2d(None, y, ok)
    assert abs(rms.item() - float(golden['rms'])) <= RMS_TOL * float(golden['rms'])
    (gy,) = torch.autograd.grad(rms, [y])
    y_cpu = torch.from_numpy(golden['out_y']).requires_grad_(True)
    ref = oracle.spot_rms(None, y_cpu, torch.from_numpy(golden['out_ok']))
    (ref_gy,) = torch.autograd.grad(ref, [y_cpu])
    assert _rel(gy.cpu().numpy(), ref_gy.numpy()) <= GRAD_TOL


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
def test_fused_spot_pass(golden, arith):
    allow = bool(golden['allow_backward_rays'])
    gpu = _inputs(golden, DEV, grad=('z', 'c', 't', 'mu'))
    rms, rms_field = ops.spot_rms(*_args(gpu), allow, rt._arith_code(arith))
    want = float(golden['rms'])
    assert abs(rms[0].item() - want) <= RMS_TOL * want, (rms[0].item(), want)
    got = torch.autograd.grad(rms[0], [gpu[k] for k in ('z', 'c', 't', 'mu')])
    for name, g in zip(('z', 'c', 't', 'mu'), got):
        ref = golden['grad_in_' + name]
        assert g.shape == ref.shape
        # (d rms / d z is a single, heavily cancelled number, ~1e-3 of d rms / d t: where the
        # reference's own fp32 value is farther than 1e-4 from its float64 run, the bar is "no farther
        # from float64 than the reference")
        scale = max(float(np.abs(golden['f64_grad_in_z']).max()), float(np.abs(golden['f64_grad_in_t']).max())) \
            if name == 'z' else None
        _close_or_no_worse_than_reference(g.cpu().numpy(), ref, golden['f64_grad_in_' + name], GRAD_TOL,
                                          f"{golden['name']} d rms/d {name}", group_scale=scale)
    # the forward-only (no grad) variant gives the same value
    plain = _inputs(golden, DEV)
    rms2, _ = ops.spot_rms(*_args(plain), allow, rt._arith_code(arith))
    assert abs(rms2[0].item() - rms[0].item()) <= 1e-6 * want


def test_raytracer_end_to_end(golden):
    """RayTracer on the GPU (incl. ray aiming, which differentiates the trace
    w.r.t. the pupil coordinates) against the reference's outputs and lens gradients."""
    from tests.conftest import golden_problem
    tracer, specs, lens = golden_problem(golden, DEV)
    structure = lens.structure
    aimed = 'aimed' in golden['name']
    out = tracer.trace_rays(specs, lens)
    if aimed:
        # the aimed pupil goes through a division by a traced slope, which amplifies fp32 noise on
        # either side: north_star tolerance, or no farther from the reference's float64 run than its fp32
        assert np.array_equal(out[4].cpu().numpy(), golden['out_ok'])
        for j, key in ((0, 'out_x'), (1, 'out_y')):
            got = out[j].detach().cpu().numpy().astype(np.float64)
            scale = max(np.abs(golden['out_x']).max(), np.abs(golden['out_y']).max())
            vs32 = np.abs(got - golden[key]).max() / scale
            ours = np.abs(got - golden['f64_' + key]).max() / scale
            theirs = np.abs(golden[key] - golden['f64_' + key]).max() / scale
            print(f"{golden['name']} {key}: ours-vs-fp32 {vs32:.2e} ours-vs-fp64 {ours:.2e} reference-fp32-vs-fp64 {theirs:.2e}")
            assert vs32 <= POINT_TOL or ours <= theirs, (key, vs32, ours, theirs)
    else:
        _check_outputs([o.detach() for o in out], golden)
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    _close_or_no_worse_than_reference(rms.item(), golden['rms'], golden['f64_rms'], RMS_TOL, f"{golden['name']} rms")
    grads = torch.autograd.grad(rms, [lens.c, lens.t, lens.nd, lens.v])
    for name, g in zip(('c', 't', 'nd'), grads):
        # (the reference's nd / v gradients are 0 * NaN at air slots; ours are 0 there)
        _close_or_no_worse_than_reference(g.cpu().numpy(), golden['grad_' + name], golden['f64_grad_' + name],
                                          GRAD_TOL, f"{golden['name']} d rms/d {name}")
    assert _rel(np.nan_to_num(grads[3].cpu().numpy()), np.nan_to_num(golden['grad_v'])) <= (5e-4 if aimed else GRAD_TOL)
    # fused pass through the same front end
    lens2 = lm.Lens(structure, *[torch.from_numpy(golden[k]).to(DEV).requires_grad_(True)
                                 for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    rms_f, _ = tracer.spot_rms(specs, lens2)
    # (the fused front end stages mu, z and cy with its own kernel: same fp32 formulas in another
    # operation order, worth ~1e-7 in mu and a few 1e-6 of a 0.013 spot -- inside RMS_TOL)
    assert abs(rms_f[0].item() - rms.item()) <= RMS_TOL * rms.item() + 1e-9
    grads_f = torch.autograd.grad(rms_f[0], [lens2.c, lens2.t, lens2.nd])
    for name, g, r in zip(('c', 't', 'nd'), grads_f, grads):
        assert _rel(g.cpu().numpy(), r.cpu().numpy()) <= GRAD_TOL, name


# ---------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties
# ---------------------------------------------------------------------------
def _double_gauss_problem(n_side, fields=16, requires_grad=False):
    specs, lens = prescriptions.double_gauss(DEV)
    if requires_grad:
        for name in ('c', 't', 'nd'):
            getattr(lens, name).requires_grad_(True)
    tracer = rt.RayTracer(mode='circular', n_rays=(n_side, n_side),
                          rel_fields=tuple(np.linspace(0, 1, fields).tolist()),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    return tracer, specs, lens


def test_double_gauss_1m_rays_exact_equals_oracle_on_device():
    """Config-2 lens at ~1M rays: the exact policy is bit-identical to the oracle
    evaluated by torch on the same GPU (IEEE sqrt/div there), masks included, and
    the guarded policy has identical masks and values within tolerance."""
    tracer, specs, lens = _double_gauss_problem(148)
    args = tracer._ray_set(specs, lens)
    exact = rt.trace_skew(*args, arith='exact')
    ref = oracle.trace(*args)
    for got, want in zip(exact, ref):
        assert torch.equal(got, torch.broadcast_to(want, got.shape))
    fast = rt.trace_skew(*args)
    assert torch.equal(fast[4], exact[4]) and torch.equal(fast[5], exact[5])
    scale = float(exact[1].abs().max())
    assert float((fast[0] - exact[0]).abs().max()) <= POINT_TOL * scale
    assert float((fast[1] - exact[1]).abs().max()) <= POINT_TOL * scale
    assert float((fast[2] - exact[2]).abs().max()) <= COS_TOL
    assert float((fast[3] - exact[3]).abs().max()) <= COS_TOL
    assert bool(exact[4].all())                    # 100 % of rays pass this lens


def test_double_gauss_4m_rays_fused_equals_split_and_shards_add_up():
    """Config 2 at full size (16 fields x 3 wavelengths x 296^2 = 4.2M rays):
    fused pass == trace + compute_rms2d + autograd (linearity of the deferred-seed
    adjoint), and the moments of two pupil shards sum to the unsharded ones."""
    tracer, specs, lens = _double_gauss_problem(296, requires_grad=True)
    rms_f, field_f = tracer.spot_rms(specs, lens)
    g_f = torch.autograd.grad(rms_f[0], [lens.c, lens.t, lens.nd])
    out = tracer.trace_rays(specs, lens)
    assert out[0].numel() == 16 * 3 * 296 * 296
    rms_s, field_s = rt.compute_rms2d_all(out[1], out[4])
    g_s = torch.autograd.grad(rms_s[0], [lens.c, lens.t, lens.nd])
    assert abs(rms_f[0].item() - rms_s[0].item()) <= RMS_TOL * rms_s[0].item()
    assert torch.allclose(field_f, field_s, rtol=1e-5)
    for a, b in zip(g_f, g_s):
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= GRAD_TOL
    args = tracer._ray_set(specs, lens)
    whole, ref_y = ops.spot_moments(*args)
    m0, r0 = ops.spot_moments(*args, shard=(0, 2))
    m1, r1 = ops.spot_moments(*args, shard=(1, 2))
    assert torch.equal(ref_y, r0) and torch.equal(ref_y, r1)
    denom = whole.abs().amax(dim=(0, 1, 2)).clamp_min(1e-30)
    assert float(((m0 + m1 - whole).abs().amax(dim=(0, 1, 2)) / denom).max()) <= 1e-5
    assert float(whole[..., -1].sum()) == 16 * 3 * 296 * 296      # every ray counted once, all ok


def test_empty_and_unsupported_inputs_fail_loudly():
    from torchoptics_b200 import _native
    rec = load_golden('cooke_8x8')
    i = _inputs(rec, DEV)
    bad = dict(i)
    bad['x'] = i['x'].double()
    with pytest.raises(TypeError):
        rt.trace_skew(*_args(bad))
    empty = dict(i)
    empty['x'] = i['x'][:, :, :0]
    empty['y'] = i['y'][:, :, :0]
    out = rt.trace_skew(*_args(empty))          # empty ray set -> empty results, no launch
    assert all(o.shape == (1, 3, 0, 3) for o in out)
    with pytest.raises((ValueError, _native.NativeLibraryError)):
        ops.spot_rms(*_args(empty))


def test_batch_of_lenses_matches_single_lens_runs():
    """B > 1: every lens of a batch gets the result of tracing it alone
    (the reference's compute_rms2d only ever looks at lens 0, rtl:695)."""
    recs = [load_golden('cooke_8x8'), load_golden('cooke_16x16_epd2.0')]
    base = _inputs(recs[0], DEV)
    scale = torch.tensor([1.0, 0.97, 1.05], device=DEV).reshape(3, 1, 1, 1, 1)
    batch = dict(base)
    batch['c'] = (base['c'] * scale).requires_grad_(True)
    batch['t'] = base['t'].expand(3, -1, -1, -1, -1).contiguous().requires_grad_(True)
    batch['mu'] = base['mu'].expand(3, -1, -1, -1, -1).contiguous()
    batch['mask'] = base['mask'].expand(3, -1, -1, -1, -1).contiguous()
    batch['z'] = base['z'].expand(3, -1, -1, -1).contiguous()
    batch['cy'] = base['cy'].expand(3, -1, -1, -1).contiguous()
    rms, _ = ops.spot_rms(*_args(batch))
    gc, = torch.autograd.grad(rms.sum(), [batch['c']])
    for b in range(3):
        one = {k: (v[b:b + 1] if v.shape[0] == 3 else v) for k, v in batch.items()}
        one['c'] = one['c'].detach().requires_grad_(True)
        r1, _ = ops.spot_rms(*_args(one))
        g1, = torch.autograd.grad(r1[0], [one['c']])
        assert abs(r1[0].item() - rms[b].item()) <= 1e-6 * r1[0].item()
        assert _rel(gc[b].cpu().numpy(), g1[0].cpu().numpy()) <= 1e-5


def test_graphed_step_equals_eager_api():
    """GraphedSpotStep (one CUDA-graph launch from pinned host memory) returns what the
    eager RayTracer.spot_rms + autograd path returns, also after the inputs change."""
    from torchoptics_b200 import GraphedSpotStep
    specs, lens = prescriptions.load_yaml('baseline_tessar.yml', DEV)
    tracer = rt.RayTracer(mode='circular', n_rays=(24, 24), rel_fields=(0., 0.5, 0.707, 1.),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    step = GraphedSpotStep(tracer, specs, lens)
    host = {k: getattr(lens, k).detach().cpu() for k in ('c', 't', 'nd', 'v')}
    for scale in (1.0, 1.01):
        host_c = host['c'] * scale
        rms, grads = step(c=host_c, t=host['t'], nd=host['nd'], v=host['v'])
        leaves = {k: (host_c if k == 'c' else host[k]).to(DEV).requires_grad_(True) for k in host}
        ref_rms, _ = tracer.spot_rms(specs, lm.Lens(lens.structure, leaves['c'], leaves['t'],
                                                    leaves['nd'], leaves['v']))
        ref = torch.autograd.grad(ref_rms.sum(), [leaves['c'], leaves['t'], leaves['nd']])
        assert abs(float(rms[0]) - ref_rms[0].item()) <= 1e-6 * ref_rms[0].item()
        for k, r in zip(('c', 't', 'nd'), ref):
            assert _rel(grads[k].numpy(), r.cpu().numpy()) <= 1e-5, k
    assert step.h2d_bytes > 0 and step.d2h_bytes > 0


@pytest.mark.parametrize('lens_file', ['baseline_cooke.yml', 'baseline_tessar.yml', 'baseline_doublet.yml',
                                       'singlet_lens.yml'])
def test_staging_kernels_equal_torch_front_end(lens_file):
    """tl_stage_fwd / tl_stage_bwd (index model, pupil position, field cosines and their chain
    rule) against the same quantities computed by the host torch code and autograd."""
    specs, lens = prescriptions.load_yaml(lens_file, DEV)
    tracer = rt.RayTracer(mode='circular', n_rays=(16, 12), rel_fields=(0., 0.4, 0.707, 1.),
                          wavelengths=('C', 'd', 'F', 546.1), default_device=DEV)

    def run(staged):
        leaves = [getattr(lens, k).detach().clone().requires_grad_(True) for k in ('c', 't', 'nd', 'v')]
        rms, field = tracer.spot_rms(specs, lm.Lens(lens.structure, *leaves), staged=staged)
        return rms, field, torch.autograd.grad(rms.sum(), leaves)

    rms_a, field_a, g_a = run(True)
    rms_b, field_b, g_b = run(False)
    assert abs(rms_a[0].item() - rms_b[0].item()) <= 1e-5 * rms_b[0].item()   # RMS_TOL: z differs by the ABCD product order
    assert torch.allclose(field_a, field_b, rtol=1e-5)
    for name, a, b in zip(('c', 't', 'nd', 'v'), g_a, g_b):
        a, b = np.nan_to_num(a.cpu().numpy()), np.nan_to_num(b.cpu().numpy())
        assert _rel(a, b) <= 2e-5, (name, _rel(a, b))


def test_thirty_surface_lens_forward_sweep_and_backward():
    """BASELINE config-4 shape: 30 spherical surfaces.  Forward trace and the gradient-free spot
    sweep (any surface count), the split backward and the fused pass with gradients (S <= 32), against the
    oracle on device."""
    specs, lens = prescriptions.wide_zoom_30(DEV)
    assert lens.c.shape == (1, 30)
    tracer = rt.RayTracer(mode='circular', n_rays=(96, 96), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    args = tracer._ray_set(specs, lens)
    ref = oracle.trace(*args)
    for policy in ('exact', 'guarded'):
        out = rt.trace_skew(*args, arith=policy)
        assert torch.equal(out[4], ref[4]) and torch.equal(out[5], torch.broadcast_to(ref[5], out[5].shape))
        if policy == 'exact':
            assert torch.equal(out[1], torch.broadcast_to(ref[1], out[1].shape))
    assert bool(ref[4].all()) and 0 < int(ref[5].sum())
    with torch.no_grad():
        rms, field = tracer.spot_rms(specs, lens)            # forward-only fused sweep, S = 30
    want = oracle.spot_rms_all_lenses(ref[1], ref[4])
    assert abs(rms[0].item() - want[0].item()) <= RMS_TOL * want[0].item()
    # two pupil shards of the sweep add up
    m, _ = ops.spot_moments(*args, want_grad=False)
    m0, _ = ops.spot_moments(*args, want_grad=False, shard=(0, 2))
    m1, _ = ops.spot_moments(*args, want_grad=False, shard=(1, 2))
    denom = m.abs().amax(dim=(0, 1, 2)).clamp_min(1e-30)       # per slot: sum(y-y0) itself can cancel to ~0
    assert float(((m0 + m1 - m).abs().amax(dim=(0, 1, 2)) / denom).max()) <= 1e-5
    # split path with gradients: trace + rms + autograd through the 32-surface backward kernel
    leaves = [getattr(lens, k_).detach().clone().requires_grad_(True) for k_ in ('c', 't', 'nd', 'v')]
    lens_g = lm.Lens(lens.structure, *leaves)
    out = tracer.trace_rays(specs, lens_g)
    rms_s = rt.compute_rms2d(out[0], out[1], out[4])
    g = torch.autograd.grad(rms_s, leaves[:3])
    ref_leaves = [l_.detach().clone().requires_grad_(True) for l_ in leaves]
    ref_out = oracle.trace(*tracer._ray_set(specs, lm.Lens(lens.structure, *ref_leaves)))
    ref_g = torch.autograd.grad(oracle.spot_rms_all_lenses(ref_out[1], ref_out[4])[0], ref_leaves[:3])
    # 30 surfaces accumulate fp32 noise on either side: north_star tolerance against the oracle's fp32
    # gradients, or no farther from the oracle run in float64 than its fp32 run is
    g64 = oracle_lens_gradients(tracer, specs, lens, torch.float64)[1]
    for name, a_, b_, c_ in zip(('c', 't', 'nd'), g, ref_g, g64):
        _close_or_no_worse_than_reference(a_.cpu().numpy(), b_.cpu().numpy(), c_.cpu().numpy(), GRAD_TOL,
                                          f'wide_zoom_30 d rms/d {name}')
    # the FUSED pass with gradients at 30 surfaces (17..32: k_spot_rev<tmem4 / tmem8>, accumulators in tensor memory)
    assert ops.spot_kernel_name(*[a.detach() for a in args]).startswith('k_spot_rev<tmem')
    fused_leaves = [l_.detach().clone().requires_grad_(True) for l_ in leaves]
    rms_f, _ = tracer.spot_rms(specs, lm.Lens(lens.structure, *fused_leaves))
    assert abs(rms_f[0].item() - want[0].item()) <= RMS_TOL * want[0].item()
    g_f = torch.autograd.grad(rms_f[0], fused_leaves[:3])
    for name, a_, b_, c_ in zip(('c', 't', 'nd'), g_f, ref_g, g64):
        _close_or_no_worse_than_reference(a_.cpu().numpy(), b_.cpu().numpy(), c_.cpu().numpy(), GRAD_TOL,
                                          f'wide_zoom_30 fused d rms/d {name}')


def test_no_grad_sweep_with_grad_requiring_lens():
    """Regression: under torch.no_grad() the fused pass must take its gradient-free variant
    (any surface count) even when the lens tensors require grad."""
    specs, lens = prescriptions.wide_zoom_30(DEV)
    for k_ in ('c', 't', 'nd'):
        getattr(lens, k_).requires_grad_(True)
    tracer = rt.RayTracer(mode='circular', n_rays=(13, 11), rel_fields=(0., 0.6, 1.), wavelengths=('C', 'd', 'F'),
                          default_device=DEV)
    with torch.no_grad():
        rms, _ = tracer.spot_rms(specs, lens)
        rms_t, _ = tracer.spot_rms(specs, lens, staged=False)
    assert not rms.requires_grad and abs(rms[0].item() - rms_t[0].item()) <= 1e-5 * rms_t[0].item()
    # with gradients the 30-surface lens takes k_spot_rev's 4-warp TMEM variant (round 1: ValueError beyond 16)
    rms_g, _ = tracer.spot_rms(specs, lens)
    assert rms_g.requires_grad and abs(rms_g[0].item() - rms_t[0].item()) <= 1e-5 * rms_t[0].item()


# ---------------------------------------------------------------------------
# many short rows: the warp-per-row spot kernel
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['cooke_8x8', 'cooke_16x16_epd2.6', 'tessar_16x16_epd2.0'])
def test_rows_kernel_matches_cta_kernel_on_a_batch_of_lenses(name, monkeypatch):
    """B = 12 perturbed copies of a golden lens (rows = B*F*W >= 64, P <= 512) take the
    warp-per-row kernel; with TL_NO_ROWS the same call takes the CTA-per-row kernel.  Same sums
    (fp32 summation order differs), and lens 0 (unperturbed) reproduces the reference's RMS."""
    rec = load_golden(name)
    allow = bool(rec['allow_backward_rays'])
    B = 12
    g = torch.Generator(device='cpu').manual_seed(7)
    base = _inputs(rec, DEV)
    S = base['c'].shape[-1]
    jitter = 1.0 + 0.01 * torch.randn((B, 1, 1, 1, S), generator=g)
    jitter[0] = 1.0

    def run():
        i = dict(base)
        i['c'] = (base['c'] * jitter.to(DEV)).requires_grad_(True)
        i['t'] = base['t'].expand(B, 1, 1, 1, S).clone().requires_grad_(True)
        i['mu'] = base['mu'].expand(B, 1, 1, -1, S).clone().requires_grad_(True)
        i['z'] = base['z'].expand(B, 1, 1, 1).clone().requires_grad_(True)
        i['cy'] = base['cy'].expand(B, -1, 1, 1).contiguous()
        i['mask'] = base['mask'].expand(B, 1, 1, 1, S).contiguous()
        rms, rms_field = ops.spot_rms(*_args(i), allow_backward_rays=allow)
        grads = torch.autograd.grad(rms.sum(), [i['c'], i['t'], i['mu'], i['z']])
        no_grad_rms, _ = ops.spot_rms(*[a.detach() for a in _args(i)], allow_backward_rays=allow)
        return rms.detach(), rms_field, grads, no_grad_rms

    monkeypatch.delenv('TL_NO_ROWS', raising=False)
    rows = run()
    monkeypatch.setenv('TL_NO_ROWS', '1')
    cta = run()
    assert float(((rows[0] - cta[0]).abs() / cta[0]).max()) <= 2e-6
    assert float(((rows[1] - cta[1]).abs() / cta[1].clamp_min(1e-12)).max()) <= 2e-6
    assert float(((rows[3] - cta[3]).abs() / cta[3]).max()) <= 2e-6          # forward-only (EVAL) variant
    assert float(((rows[3] - rows[0]).abs() / rows[0]).max()) <= 2e-6
    for a, b in zip(rows[2], cta[2]):
        # (two adjoint formulations -- parked hit points vs the reversible walk -- on a bundle with rays at the
        # thresholds: a third of north_star's 1e-4)
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= 3e-5
    assert abs(float(rows[0][0]) - float(rec['rms'])) <= RMS_TOL * float(rec['rms'])


# ---------------------------------------------------------------------------
# ray aiming on the device
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['cooke_8x8_aimed', 'tessar_8x8_aimed', 'cooke_8x8_vig_aimed', 'tessar_8x8_vig_aimed',
                                  'cooke_8x8_paraxial_aimed', 'tessar_8x8_vig_paraxial_aimed'])
def test_device_ray_aiming_matches_reference_and_torch_path(name):
    """tl_aim (one kernel) against (1) the aimed pupil the reference handed to trace_skew (golden in_x /
    in_y) and (2) this package's torch mirror of rtl:129-208 (three nested traces + autograd), on the
    same lens -- with the 'real' and the 'paraxial' stop radius, without and with a pupil vignetting
    function (the branches the lite reference runs only with its two commented-out helpers restored,
    tests/golden/make_golden.py)."""
    from tests.conftest import golden_problem
    golden = load_golden(name)
    golden['name'] = name
    on_device, specs, lens = golden_problem(golden, DEV, requires_grad=False)
    mirror, _, _ = golden_problem(golden, DEV, requires_grad=False)
    mirror.device_aiming = False
    before = rt.ops.nat.launch_count()
    args_dev = on_device._ray_set(specs, lens)
    launched = rt.ops.nat.launch_count() - before
    args_ref = mirror._ray_set(specs, lens)
    assert launched == 2                                  # tl_stage_fwd + tl_aim, nothing else of ours
    scale = float(np.abs(golden['in_y']).max())
    for j, key in ((0, 'in_x'), (1, 'in_y')):
        got = args_dev[j].cpu().numpy()
        assert got.shape == golden[key].shape
        assert np.abs(got - golden[key]).max() <= 2e-5 * scale, key            # vs the reference
        assert np.abs(got - args_ref[j].cpu().numpy()).max() <= 2e-5 * scale, key    # vs the torch mirror
    out = on_device.trace_rays(specs, lens)
    assert np.array_equal(out[4].cpu().numpy(), golden['out_ok'])
    assert np.abs(out[1].cpu().numpy() - golden['out_y']).max() <= 5e-5 * np.abs(golden['out_y']).max()
    rms = rt.compute_rms2d(out[0], out[1], out[4])
    assert abs(rms.item() - float(golden['rms'])) <= 5e-5 * float(golden['rms'])


@pytest.mark.parametrize('name', ['cooke_8x8_aimed', 'tessar_8x8_aimed', 'cooke_8x8_vig_aimed', 'tessar_8x8_vig_aimed',
                                  'cooke_8x8_paraxial_aimed', 'tessar_8x8_vig_paraxial_aimed'] + ['cooke_8x8_vig'])
def test_staged_fused_pass_with_ray_aiming(name):
    """RayTracer.spot_rms with ray aiming (and / or pupil vignetting) stays on the staged path (one
    launch for staging + aiming + reference heights, then the fused kernel applying the maps on
    load): same RMS and gradients as the unstaged path that materialises the aimed [B,F,P,W] pupil,
    and the reference's RMS."""
    from tests.conftest import golden_problem
    golden = load_golden(name)
    golden['name'] = name
    tracer, specs, lens0 = golden_problem(golden, DEV)
    structure = lens0.structure
    assert tracer._staging(lens0)[0]

    def run(staged):
        lens = lm.Lens(structure, *[torch.from_numpy(golden[k]).to(DEV).requires_grad_(True)
                                    for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
        before = rt.ops.nat.launch_count()
        rms, _ = tracer.spot_rms(specs, lens, staged=staged)
        grads = torch.autograd.grad(rms[0], [lens.c, lens.t, lens.nd])
        return rms.detach(), grads, rt.ops.nat.launch_count() - before

    rms_s, g_s, n_s = run(True)
    rms_u, g_u, n_u = run(False)
    assert n_s <= 8                                   # stage, aim, chief, trace+adjoint, reduce, finalize, chain rule
    # the aimed pupil goes through a division by a traced slope, which amplifies fp32 noise: each path is
    # held to north_star against the reference -- or to "no farther from its float64 run than its own
    # fp32" -- and the two paths to each other within twice that noise
    for tag, r in (('staged', rms_s), ('unstaged', rms_u)):
        _close_or_no_worse_than_reference(float(r[0]), golden['rms'], golden['f64_rms'], RMS_TOL, f'{name} aimed {tag} rms')
    noise = abs(float(golden['rms']) - float(golden['f64_rms']))
    assert abs(float(rms_s[0]) - float(rms_u[0])) <= max(RMS_TOL * float(rms_u[0]), 2 * noise)
    for a, b in zip(g_s, g_u):
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= GRAD_TOL


def test_spot_rms_and_grads_matches_autograd_and_writes_into_given_buffers():
    """The autograd-free staged call (what GraphedSpotStep replays) against autograd over spot_rms,
    with and without caller-provided output views of one packed buffer."""
    specs, lens = prescriptions.double_gauss(DEV)
    tracer = rt.RayTracer(mode='circular', n_rays=(40, 40), rel_fields=(0., 0.5, 1.), wavelengths=('C', 'd', 'F'),
                          default_device=DEV)
    leaves = {k: getattr(lens, k).detach().clone().requires_grad_(True) for k in ('c', 't', 'nd', 'v')}
    rms, _ = tracer.spot_rms(specs, lm.Lens(lens.structure, *leaves.values()))
    want = torch.autograd.grad(rms.sum(), list(leaves.values()))
    got_rms, got = tracer.spot_rms_and_grads(specs, lens)
    assert torch.equal(got_rms, rms.detach())
    for k, w in zip(('c', 't', 'nd', 'v'), want):
        assert torch.equal(got[k], w), k
    per, B = lens.c.numel(), lens.c.shape[0]
    buf = torch.full((4 * per + B,), float('nan'), device=DEV)
    out = {name: buf[i * per:(i + 1) * per].view(lens.c.shape) for i, name in enumerate(('gc', 'gt', 'gnd', 'gv'))}
    out['rms'] = buf[4 * per:]
    tracer.spot_rms_and_grads(specs, lens, out=out)
    assert torch.equal(buf[4 * per:], rms.detach())
    for i, w in enumerate(want):
        assert torch.equal(buf[i * per:(i + 1) * per].view(lens.c.shape), w)


def test_new_entry_points_reject_what_they_do_not_cover():
    """Loud failures, never a silent wrong answer: the aiming map / penalty terms with extension
    surfaces, per-ray x/y gradients through an aiming map, undersized peer windows."""
    import ctypes
    from torchoptics_b200 import _native
    from torchoptics_b200.peer import PeerExchange
    rec = load_golden('cooke_8x8')
    i = _inputs(rec, DEV)
    S = i['c'].shape[-1]
    k = torch.zeros((1, 1, 1, 1, S), device=DEV)
    with pytest.raises(ValueError):                      # stacks are defined for spherical lenses only
        rt.trace_skew(*_args(i), aggregate=True, k=k)
    with pytest.raises(_native.NativeLibraryError):      # same at the C ABI
        lay = ops._Layout(*_args(i), k=k)
        pb = lay.problem(True, _native.ARITH_GUARDED)
        assert _native.load().tl_penalty_workspace(ctypes.byref(pb)) == 0
        mom = torch.empty(3 * 3 * (3 * S + 2), dtype=torch.float64, device=DEV)
        _native.check(_native.load().tl_penalty_accumulate(ctypes.byref(pb), mom.data_ptr(), mom.data_ptr(), 8,
                                                           None), 'tl_penalty_accumulate')
    i['x'], i['y'] = i['x'] * 0.25, i['y'] * 0.25     # inside the map's clamp to [-2, 2] (rtl:111)
    assert float(i['y'].abs().max()) < 2.0
    lay = ops._Layout(*_args(i))
    pb = lay.problem(True, _native.ARITH_GUARDED)
    aim = torch.ones((1, 3, 3, 3), device=DEV)
    pb.aim = aim.data_ptr()
    ws_bytes = _native.load().tl_trace_bwd_workspace(ctypes.byref(pb))
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=DEV)
    g = [torch.zeros(n, device=DEV) for n in (S, S, 3 * S, 1)]
    gx = torch.zeros(lay.shape, device=DEV)
    seeds = _native.TlSeeds()
    grads = _native.TlGrads(g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), gx.data_ptr())
    rc = _native.load().tl_trace_bwd(ctypes.byref(pb), ctypes.byref(seeds), ctypes.byref(grads), ws.data_ptr(),
                                     ws_bytes, None)
    assert rc == -1 and b'ray-aiming' in _native.load().tl_last_error()
    # an aiming map of ones with a zero shift is the identity
    out_plain = rt.trace_skew(*_args(i))
    aim[..., 2] = 0.0
    o = [torch.empty(lay.shape, device=DEV) for _ in range(4)] + \
        [torch.empty(lay.shape, dtype=torch.bool, device=DEV) for _ in range(2)]
    tout = _native.TlTraceOut(*[t.data_ptr() for t in o])
    _native.check(_native.load().tl_trace_fwd(ctypes.byref(pb), ctypes.byref(tout), None), 'tl_trace_fwd')
    torch.cuda.synchronize()
    for a, b in zip(o, out_plain):
        assert torch.equal(a, b)
    ex = PeerExchange(capacity=16)
    with pytest.raises(ValueError):
        ex.all_reduce(torch.zeros(17, dtype=torch.float64, device=DEV))
    with pytest.raises(ValueError):
        ex.all_reduce(torch.zeros(8, dtype=torch.float32, device=DEV))
    ex.close()


@pytest.mark.parametrize('aggregate', [False, True])
@pytest.mark.parametrize('name', ['cooke_8x8', 'cooke_16x16_epd2.6'])
def test_unfused_backward_rows_kernel_matches_cta_kernel(name, aggregate, monkeypatch):
    """trace_skew(...).backward() on a batch of 12 lenses (>= 64 short rows): the warp-per-row
    backward (with and without seeds on the aggregate=True stacks) against the CTA-per-row one,
    per-ray gradients of x, y, cx, cy included."""
    rec = load_golden(name)
    allow = bool(rec['allow_backward_rays'])
    B = 12
    shape = (B,) + rec['out_ok'].shape[1:]
    g = torch.Generator(device='cpu').manual_seed(3)
    base = _inputs(rec, DEV)
    S = base['c'].shape[-1]
    jitter = 1.0 + 0.01 * torch.randn((B, 1, 1, 1, S), generator=g)
    seeds = [torch.rand(shape, generator=g).to(DEV) + 0.1 for _ in range(4)]
    stack_seed = (torch.rand((S,) + shape, generator=g) + 0.1).to(DEV)

    def run():
        i = {}
        for key in ('x', 'y', 'cx'):
            i[key] = torch.broadcast_to(base[key], shape).contiguous().requires_grad_(True)
        i['cy'] = torch.broadcast_to(base['cy'], shape).contiguous().requires_grad_(True)
        i['z'] = base['z'].expand(B, 1, 1, 1).clone().requires_grad_(True)
        i['c'] = (base['c'] * jitter.to(DEV)).requires_grad_(True)
        i['t'] = base['t'].expand(B, 1, 1, 1, S).clone().requires_grad_(True)
        i['mu'] = base['mu'].expand(B, 1, 1, -1, S).clone().requires_grad_(True)
        i['mask'] = base['mask'].expand(B, 1, 1, 1, S).contiguous()
        out = rt.trace_skew(*_args(i), aggregate=aggregate, allow_backward_rays=allow)
        loss = sum((s * o).sum() for s, o in zip(seeds, out[:4]))
        if aggregate:
            loss = loss + sum((torch.stack(out[6][k]) * stack_seed).sum() for k in out[6])
        return torch.autograd.grad(loss, [i[k] for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu')])

    monkeypatch.delenv('TL_NO_ROWS', raising=False)
    rows = run()
    monkeypatch.setenv('TL_NO_ROWS', '1')
    cta = run()
    for k, a, b in zip(('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu'), rows, cta):
        assert torch.isfinite(a).all(), k
        assert float((a - b).norm()) <= 2e-5 * float(b.norm()), k
