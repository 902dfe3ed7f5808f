"""trace_skew(aggregate=True) on the GPU (rtl:641-657): the three penalty stacks and their
gradients, through the C ABI, against the oracle and the golden vectors made by the reference
(tests/golden/aggregate).

Tolerances: z_RELU like points (1e-5 of the lens scale); theta = acos(cos)/(pi/2) is compared
through its cosine (1e-6 abs: acos is ill-conditioned at normal incidence, where one ULP of the
cosine moves theta by 3e-4) and directly (1e-5) where sin(theta) > 0.05; masks (theta == 1) exact;
gradients ||dg|| / ||g|| <= 1e-4 per parameter group against the fp64 oracle -- or, where the
reference's own fp32 arithmetic is further than that from fp64 (the angle terms amplify rounding
by 1 / sin(theta)), no further than 1.5x what the reference's fp32 evaluation is.  Exact policy: z_RELU bit-identical, angles
within 4 ULP (two acos implementations on identical cos^2 bits).
"""
import numpy as np
import pytest
import torch

from oracle import trace_oracle as oracle
from tests.test_penalty_cpu import AGG_CASES, KEYS, load_aggregate
from torchoptics_b200 import ray_tracing_lite as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GRAD_TOL = 1e-4
PENALTY_RATE = 0.2


def _inputs(rec, device, broadcast, grad=(), dtype=torch.float32):
    shape = rec['out_ok'].shape
    t = {}
    for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu'):
        v = rec['in_' + k]
        if broadcast and k in ('x', 'y', 'cx', 'cy'):
            v = np.ascontiguousarray(np.broadcast_to(v, shape))
        t[k] = torch.from_numpy(v).to(device=device, dtype=dtype)
    t['mask'] = torch.from_numpy(rec['in_mask']).to(device)
    for k in grad:
        t[k] = t[k].clone().requires_grad_(True)
    return t


def _args(i):
    return [i[k] for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'mask')]


def _stack(stacks, shape):
    return {k: torch.stack([torch.broadcast_to(s, shape) for s in stacks[k]]) for k in KEYS}


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)


def _check_stacks(got, want, scale, exact):
    zr, zr_ref = got['z_RELU'], want['z_RELU']
    if exact:
        assert np.array_equal(zr.view(np.uint32), zr_ref.view(np.uint32))
    assert np.abs(zr - zr_ref).max() <= 1e-5 * scale
    for key in ('theta_norm', 'theta_prime_norm'):
        a, b = got[key].astype(np.float64), want[key].astype(np.float64)
        assert np.array_equal(a == 1.0, b == 1.0), key                    # failed-ray marks
        # (cosines of the angles: north_star holds direction cosines to 1e-5; these sit at ~1e-6)
        assert np.abs(np.cos(a * np.pi / 2) - np.cos(b * np.pi / 2)).max() <= 2e-6, key
        steep = np.sin(b * np.pi / 2) > 0.05
        assert np.abs(a - b)[steep].max(initial=0.0) <= 1e-5, key
        if exact:
            np.testing.assert_allclose(a, b, rtol=5e-7, atol=1e-7)


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
@pytest.mark.parametrize('name', AGG_CASES)
def test_stacks_match_reference_and_oracle(name, arith):
    rec, agg = load_aggregate(name)
    allow = bool(rec['allow_backward_rays'])
    shape = rec['out_ok'].shape
    scale = max(np.abs(rec['out_x']).max(), np.abs(rec['out_y']).max(), np.abs(rec['in_t']).max())
    # un-broadcast pupil grids as trace_rays hands them over (the reference raises there for W > 1)
    i = _inputs(rec, DEV, broadcast=False)
    out = rt.trace_skew(*_args(i), aggregate=True, allow_backward_rays=allow, arith=arith)
    assert len(out) == 7 and set(out[6]) == set(KEYS)
    assert all(len(out[6][k]) == rec['in_t'].shape[-1] and out[6][k][0].shape == shape for k in KEYS)
    got = {k: v.cpu().numpy() for k, v in _stack(out[6], shape).items()}
    assert np.array_equal(out[4].cpu().numpy(), rec['out_ok'])
    # the reference's own stacks (CPU sqrt: within the float32 budget)
    _check_stacks(got, {k: agg[k] for k in KEYS}, scale, exact=False)
    # the oracle with the correctly rounded sqrt: what the exact policy reproduces
    cpu = _inputs(rec, 'cpu', broadcast=True)
    with oracle.ieee_sqrt():
        ref = oracle.trace(*_args(cpu), True, allow)
    want = {k: v.numpy() for k, v in _stack(ref[6], shape).items()}
    _check_stacks(got, want, scale, exact=(arith == 'exact'))
    # the six regular outputs are those of aggregate=False
    plain = rt.trace_skew(*_args(i), allow_backward_rays=allow, arith=arith)
    for a, b in zip(out[:6], plain):
        assert torch.equal(a, b)


def _loss(out, oracle_side, n_seq):
    stacks = out[6]
    q = (torch.stack(stacks['theta_norm']).sum(0) + torch.stack(stacks['theta_prime_norm']).sum(0) +
         torch.stack(stacks['z_RELU']).sum(0)) / n_seq
    q = torch.where(torch.isnan(q), torch.zeros_like(q), q)
    penalty = q.sum()
    rms = oracle.spot_rms(out[0], out[1], out[4]) if oracle_side else rt.compute_rms2d(out[0], out[1], out[4])
    return rms, penalty


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
@pytest.mark.parametrize('name', AGG_CASES)
def test_penalty_and_loss_gradients(name, arith):
    """d penalty / d (z, c, t, mu) and d (rms + 0.2 penalty) / d (...) as compute_loss_out builds
    them (optics_simulator_lite.py:430-450)."""
    rec, agg = load_aggregate(name)
    allow = bool(rec['allow_backward_rays'])
    n_seq = int(agg['n_seq'])
    names = ('z', 'c', 't', 'mu')
    i = _inputs(rec, DEV, broadcast=False, grad=names)
    out = rt.trace_skew(*_args(i), aggregate=True, allow_backward_rays=allow, arith=arith)
    rms, penalty = _loss(out, False, n_seq)
    leaves = [i[k] for k in names]
    g_pen = torch.autograd.grad(penalty, leaves, retain_graph=True)
    g_loss = torch.autograd.grad(rms + PENALTY_RATE * penalty, leaves)
    assert abs(float(penalty) - float(agg['penalty'])) <= 2e-5 * abs(float(agg['penalty']))
    # fp64 oracle with finite penalty gradients: the truth both fp32 evaluations approximate
    cpu = _inputs(rec, 'cpu', broadcast=True, grad=names, dtype=torch.float64)
    with oracle.finite_penalty_gradients(), oracle.fp32_clamp_bound():
        ref = oracle.trace(*_args(cpu), True, allow)
    all_ok = bool(rec['out_ok'].all())
    same_masks = torch.equal(ref[4], torch.from_numpy(rec['out_ok']))
    rms64, pen64 = _loss(ref, True, n_seq)
    ref_leaves = [cpu[k] for k in names]
    want_pen = torch.autograd.grad(pen64, ref_leaves, retain_graph=True)
    want_loss = torch.autograd.grad(rms64 + PENALTY_RATE * pen64, ref_leaves)
    def rel(got, want, group_with=None):
        # d/dz is the heavily cancelled sum of two axial shifts: like the RMS
        # tests, that one scalar is held to 1e-4 of the {z, t} group scale
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        scale = np.linalg.norm(want) if group_with is None else \
            np.linalg.norm(np.concatenate([want.ravel(), np.asarray(group_with, np.float64).ravel()]))
        return np.linalg.norm(got - want) / max(scale, 1e-30)

    # The angle terms are ill-conditioned at normal incidence (d theta / d cos = 1 / sin theta), so
    # the REFERENCE's own fp32 gradient sits up to 4e-4 from its fp64 evaluation on these cases
    # (cooke_8x8: c 4.2e-4, tessar_8x8: t 2.8e-4).  The bar is therefore: within 1e-4 of the fp64
    # truth, or no further from it than 1.5x the reference's fp32 arithmetic is.
    cpu32 = _inputs(rec, 'cpu', broadcast=True, grad=names)
    with oracle.finite_penalty_gradients():
        ref32 = oracle.trace(*_args(cpu32), True, allow)
    rms32, pen32 = _loss(ref32, True, n_seq)
    leaves32 = [cpu32[k] for k in names]
    ref32_pen = torch.autograd.grad(pen32, leaves32, retain_graph=True)
    ref32_loss = torch.autograd.grad(rms32 + PENALTY_RATE * pen32, leaves32)
    for j, (k, got, want, want_l, got_l) in enumerate(zip(names, g_pen, want_pen, want_loss, g_loss)):
        assert torch.isfinite(got).all() and torch.isfinite(got_l).all(), k
        gw, gwl = (want_pen[2].numpy(), want_loss[2].numpy()) if k == 'z' else (None, None)
        if same_masks:       # fp64 takes the same branch for every ray
            bar = max(GRAD_TOL, 1.5 * rel(ref32_pen[j].numpy(), want.numpy(), gw))
            bar_l = max(GRAD_TOL, 1.5 * rel(ref32_loss[j].numpy(), want_l.numpy(), gwl))
            assert bar <= 1e-3 and bar_l <= 1e-3
            assert rel(got.cpu().numpy(), want.numpy(), gw) <= bar, (k, 'penalty vs fp64 oracle')
            assert rel(got_l.cpu().numpy(), want_l.numpy(), gwl) <= bar_l, (k, 'loss vs fp64 oracle')
            if all_ok:       # the reference's own gradient is finite only when no ray fails
                gwa, gwla = (agg['gpen_t'], agg['gloss_t']) if k == 'z' else (None, None)
                assert rel(got.cpu().numpy(), agg['gpen_' + k], gwa) <= 2 * bar, (k, 'penalty vs reference')
                assert rel(got_l.cpu().numpy(), agg['gloss_' + k], gwla) <= 2 * bar_l, (k, 'loss vs reference')
    if not same_masks:       # rays on a threshold: compare with the fp32 oracle, same masks by construction
        cpu32 = _inputs(rec, 'cpu', broadcast=True, grad=names)
        with oracle.finite_penalty_gradients(), oracle.ieee_sqrt():
            ref32 = oracle.trace(*_args(cpu32), True, allow)
        _, pen32 = _loss(ref32, True, n_seq)
        want32 = torch.autograd.grad(pen32, [cpu32[k] for k in names])
        for k, got, want in zip(names, g_pen, want32):
            assert _rel(got.cpu().numpy(), want.numpy()) <= 10 * GRAD_TOL, (k, 'penalty vs fp32 oracle')


def test_partial_seeds_and_per_ray_gradients():
    """Only one stack used, and gradients w.r.t. the per-ray inputs x, y (the ray-aiming
    contract, rtl:170-180) through the penalty terms."""
    rec, agg = load_aggregate('cooke_16x16_epd2.6')
    names = ('x', 'y', 'z', 'c', 't', 'mu')
    i = _inputs(rec, DEV, broadcast=True, grad=names)
    out = rt.trace_skew(*_args(i), aggregate=True)
    g = torch.Generator(device='cpu').manual_seed(5)
    S = rec['in_t'].shape[-1]
    seed = torch.rand((S,) + rec['out_ok'].shape, generator=g) + 0.2
    loss = (torch.stack(out[6]['theta_prime_norm']) * seed.to(DEV)).sum()
    got = torch.autograd.grad(loss, [i[k] for k in names])
    cpu = _inputs(rec, 'cpu', broadcast=True, grad=names, dtype=torch.float64)
    with oracle.finite_penalty_gradients():
        ref = oracle.trace(*_args(cpu), True)
    want = torch.autograd.grad((torch.stack(ref[6]['theta_prime_norm']) * seed.double()).sum(),
                               [cpu[k] for k in names])
    if torch.equal(ref[4], torch.from_numpy(rec['out_ok'])):
        for k, a, b in zip(names, got, want):
            assert torch.isfinite(a).all()
            assert _rel(a.cpu().numpy(), b.numpy()) <= GRAD_TOL, k


def test_large_pupil_penalty_matches_oracle_on_device():
    """Config-2 lens, 0.5 M rays: stacks against the oracle evaluated by torch on the same GPU
    (correctly rounded sqrt there), loss gradient against fp64."""
    from torchoptics_b200 import RayTracer, prescriptions
    specs, lens = prescriptions.double_gauss(DEV)
    tracer = RayTracer(mode='circular', n_rays=(104, 104), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                       wavelengths=('C', 'd', 'F'), default_device=DEV)
    args = [a.detach() for a in tracer._ray_set(specs, lens)]
    shape = (1, 16, 104 * 104, 3)
    full = [torch.broadcast_to(a, shape).contiguous() if j in (0, 1, 3, 4) else a for j, a in enumerate(args)]
    out = rt.trace_skew(*args, aggregate=True, arith='exact')
    ref = oracle.trace(*full, True)
    assert torch.equal(out[4], ref[4])
    for k in KEYS:
        a, b = torch.stack(out[6][k]), torch.stack([torch.broadcast_to(s, shape) for s in ref[6][k]])
        if k == 'z_RELU':
            assert torch.equal(a, b)
        else:
            assert float((torch.cos(a.double() * np.pi / 2) - torch.cos(b.double() * np.pi / 2)).abs().max()) <= 1e-6
    # gradients of the summed penalty w.r.t. c, t, mu: guarded policy against fp64 autograd
    leaves = [args[j].clone().requires_grad_(True) for j in (5, 6, 7)]
    call = list(args)
    call[5], call[6], call[7] = leaves
    o2 = rt.trace_skew(*call, aggregate=True)
    pen = sum(torch.stack(o2[6][k]).sum() for k in KEYS)
    got = torch.autograd.grad(pen, leaves)
    leaves64 = [args[j].double().clone().requires_grad_(True) for j in (5, 6, 7)]
    call64 = [a.double() if a.dtype == torch.float32 else a for a in full]
    call64[5], call64[6], call64[7] = leaves64
    with oracle.finite_penalty_gradients(), oracle.fp32_clamp_bound():
        r64 = oracle.trace(*call64, True)
    pen64 = sum(torch.stack([torch.broadcast_to(s, shape) for s in r64[6][k]]).sum() for k in KEYS)
    want = torch.autograd.grad(pen64, leaves64)
    # the bar of test_penalty_and_loss_gradients: the oracle in fp32 on the same device sets it
    leaves32 = [args[j].clone().requires_grad_(True) for j in (5, 6, 7)]
    call32 = list(full)
    call32[5], call32[6], call32[7] = leaves32
    with oracle.finite_penalty_gradients():
        r32 = oracle.trace(*call32, True)
    pen32 = sum(torch.stack([torch.broadcast_to(s, shape) for s in r32[6][k]]).sum() for k in KEYS)
    noise = torch.autograd.grad(pen32, leaves32)
    for label, a, b, c32 in zip(('c', 't', 'mu'), got, want, noise):
        bar = max(GRAD_TOL, 1.5 * _rel(c32.cpu().numpy(), b.cpu().numpy()))
        assert bar <= 1e-3
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= bar, (label, bar)


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
@pytest.mark.parametrize('name', AGG_CASES)
def test_fused_penalty_pass_matches_reference_and_unfused(name, arith):
    """ops.penalty_sum: value and gradients of sum(Q) in one pass with no stacks, against the
    reference's number (golden), the fp64 oracle, and the unfused aggregate=True path; sharded over
    two pupil slices the sums add up to the same result."""
    from torchoptics_b200 import _native, ops
    rec, agg = load_aggregate(name)
    allow = bool(rec['allow_backward_rays'])
    n_seq = int(agg['n_seq'])
    names = ('z', 'c', 't', 'mu')
    code = _native.ARITH_EXACT if arith == 'exact' else _native.ARITH_GUARDED
    i = _inputs(rec, DEV, broadcast=False, grad=names)
    pen = ops.penalty_sum(*_args(i), n_seq, allow, code)
    assert pen.shape == (1,)
    assert abs(float(pen[0]) - float(agg['penalty'])) <= 2e-5 * abs(float(agg['penalty']))
    got = torch.autograd.grad(pen[0], [i[k] for k in names])
    # unfused path of this package
    j = _inputs(rec, DEV, broadcast=False, grad=names)
    out = rt.trace_skew(*_args(j), aggregate=True, allow_backward_rays=allow, arith=arith)
    _, pen_u = _loss(out, False, n_seq)
    unfused = torch.autograd.grad(pen_u, [j[k] for k in names])
    assert abs(float(pen[0]) - float(pen_u)) <= 5e-6 * abs(float(pen_u))     # (vs the reference itself: 2e-5 above)
    group = np.linalg.norm(np.concatenate([unfused[0].cpu().numpy().ravel(), unfused[2].cpu().numpy().ravel()]))
    for k, a, b in zip(names, got, unfused):
        a, b = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64)
        scale = group if k == 'z' else np.linalg.norm(b)
        assert np.linalg.norm(a - b) <= 2e-5 * scale, (k, 'fused vs unfused')
    # fp64 oracle (fp32 clamp constant, finite gradients)
    cpu = _inputs(rec, 'cpu', broadcast=True, grad=names, dtype=torch.float64)
    with oracle.finite_penalty_gradients(), oracle.fp32_clamp_bound():
        ref = oracle.trace(*_args(cpu), True, allow)
    if torch.equal(ref[4], torch.from_numpy(rec['out_ok'])):
        _, pen64 = _loss(ref, True, n_seq)
        want = torch.autograd.grad(pen64, [cpu[k] for k in names])
        assert abs(float(pen[0]) - float(pen64)) <= 2e-5 * abs(float(pen64))
        group64 = np.linalg.norm(np.concatenate([want[0].numpy().ravel(), want[2].numpy().ravel()]))
        for k, a, b in zip(names, got, want):
            a, b = a.cpu().numpy().astype(np.float64), b.numpy()
            scale = group64 if k == 'z' else np.linalg.norm(b)
            assert np.linalg.norm(a - b) <= GRAD_TOL * scale, (k, 'fused vs fp64 oracle')
    # two pupil slices: additive moments -> same numbers
    P = rec['out_ok'].shape[2]
    if P >= 2:
        parts = []
        for rank in range(2):
            lo, hi = ops.pupil_slice(P, rank, 2)
            s = _inputs(rec, DEV, broadcast=True, grad=names)
            for key in ('x', 'y', 'cx', 'cy'):
                s[key] = s[key][:, :, lo:hi].contiguous()
            p_r = ops.penalty_sum(*_args(s), n_seq, allow, code)
            parts.append((p_r, torch.autograd.grad(p_r[0], [s[k] for k in names])))
        assert abs(float(parts[0][0] + parts[1][0]) - float(pen[0])) <= 2e-6 * abs(float(pen[0]))
        for idx, k in enumerate(names):
            both = parts[0][1][idx] + parts[1][1][idx]
            scale = group if k == 'z' else float(got[idx].norm())
            assert float((both - got[idx]).norm()) <= 2e-5 * scale, (k, 'sliced')


def test_loss_unsup_front_end_matches_compute_loss_out():
    """RayTracer.loss_unsup == the reference's compute_loss_out numbers on the Cooke triplet
    (rms 0.01862689, golden penalty), with gradients w.r.t. the lens through the front end."""
    from torchoptics_b200 import lens_modeling as lm
    from tests.conftest import load_golden
    golden = load_golden('cooke_8x8')
    _, agg = load_aggregate('cooke_8x8')
    structure = lm.Structure(golden['stop_idx'], sequence=golden['sequence'], default_device=DEV)
    lens = lm.Lens(structure, *[torch.from_numpy(golden[k]).to(DEV).requires_grad_(True)
                                for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    specs = lm.Specs(structure, torch.from_numpy(golden['epd']).to(DEV), torch.from_numpy(golden['hfov']).to(DEV))
    tracer = rt.RayTracer(mode='circular', n_rays=(8, 8), rel_fields=(0., 0.707, 1.), wavelengths=('C', 'd', 'F'),
                          default_device=DEV)
    res = tracer.loss_unsup(specs, lens, penalty_rate=PENALTY_RATE)
    assert abs(float(res['rms'][0]) - float(agg['rms'])) <= 1e-5 * float(agg['rms'])
    assert abs(float(res['penalty'][0]) - float(agg['penalty'])) <= 2e-5 * float(agg['penalty'])
    want = float(agg['rms']) + PENALTY_RATE * float(agg['penalty'])
    assert abs(float(res['loss_unsup'][0]) - want) <= 2e-5 * want
    grads = torch.autograd.grad(res['loss_unsup'][0], [lens.c, lens.t, lens.nd])
    # the same loss through the unfused API
    lens2 = lm.Lens(structure, *[torch.from_numpy(golden[k]).to(DEV).requires_grad_(True)
                                 for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    out = tracer.trace_rays(specs, lens2, aggregate=True)
    rms, pen = _loss(out, False, int(agg['n_seq']))
    ref = torch.autograd.grad(rms + PENALTY_RATE * pen, [lens2.c, lens2.t, lens2.nd])
    for a, b in zip(grads, ref):
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= GRAD_TOL


def test_graphed_loss_step_matches_eager():
    """GraphedSpotStep(penalty_rate=...): the CUDA-graph replay of compute_loss_out's loss gives
    the eager front end's numbers, replay after replay, and follows new prescriptions."""
    from torchoptics_b200 import GraphedSpotStep, RayTracer, prescriptions
    from torchoptics_b200.lens_modeling import Lens
    specs, lens = prescriptions.double_gauss(DEV)
    tracer = RayTracer(mode='circular', n_rays=(48, 48), rel_fields=(0., 0.6, 1.), wavelengths=('C', 'd', 'F'),
                       default_device=DEV)
    step = GraphedSpotStep(tracer, specs, lens, penalty_rate=PENALTY_RATE)
    host = {k: getattr(lens, k).detach().cpu().clone() for k in ('c', 't', 'nd', 'v')}
    for trial in range(3):
        if trial:
            host['c'] = host['c'] * (1.0 + 0.002 * trial)
        rms, grads = step(**host)
        leaves = {k: host[k].to(DEV).requires_grad_(True) for k in ('c', 't', 'nd')}
        res = tracer.loss_unsup(specs, Lens(lens.structure, leaves['c'], leaves['t'], leaves['nd'], host['v'].to(DEV)),
                                penalty_rate=PENALTY_RATE)
        want = torch.autograd.grad(res['loss_unsup'].sum(), list(leaves.values()))
        assert abs(float(rms[0]) - float(res['rms'][0])) <= 1e-6 * float(res['rms'][0])
        assert abs(float(step.host_penalty[0]) - float(res['penalty'][0])) <= 1e-6 * float(res['penalty'][0])
        for k, w in zip(('c', 't', 'nd'), want):
            assert _rel(grads[k].numpy(), w.cpu().numpy()) <= 1e-5, k


@pytest.mark.parametrize('name', ['cooke_8x8', 'cooke_16x16_epd2.6', 'tessar_16x16_epd2.0'])
def test_penalty_rows_kernel_matches_cta_kernel(name, monkeypatch):
    """A batch of 12 perturbed lenses (>= 64 short rows) takes the warp-per-row penalty kernel;
    with TL_NO_ROWS the CTA-per-row one.  Same value and gradients, failing rays included."""
    from torchoptics_b200 import ops
    rec, agg = load_aggregate(name)
    allow = bool(rec['allow_backward_rays'])
    n_seq = int(agg['n_seq'])
    B = 12
    g = torch.Generator(device='cpu').manual_seed(11)
    base = _inputs(rec, DEV, broadcast=False)
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
        pen = ops.penalty_sum(*_args(i), n_seq, allow)
        return pen.detach(), torch.autograd.grad(pen.sum(), [i['c'], i['t'], i['mu'], i['z']])

    monkeypatch.delenv('TL_NO_ROWS', raising=False)
    rows = run()
    monkeypatch.setenv('TL_NO_ROWS', '1')
    cta = run()
    assert float(((rows[0] - cta[0]).abs() / cta[0]).max()) <= 2e-6
    group = float(torch.cat([cta[1][1].reshape(-1), cta[1][3].reshape(-1)]).norm())
    for k, a, b in zip(('c', 't', 'mu', 'z'), rows[1], cta[1]):
        scale = group if k == 'z' else float(b.norm())
        assert float((a - b).norm()) <= 2e-5 * scale, k
    assert abs(float(rows[0][0]) - float(agg['penalty'])) <= 2e-5 * float(agg['penalty'])
