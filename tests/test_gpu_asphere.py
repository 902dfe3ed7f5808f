"""Extension surfaces (conic / even asphere / clip / OPL) on the GPU against the extension
oracle.  PARITY UNPINNED: the reference has no such surfaces; the oracle that defines them is
itself checked by tests/test_asphere_oracle.py.  Needs a B200: ``pytest -m gpu``."""
import numpy as np
import pytest
import torch

from oracle import asphere_oracle as gen
from oracle import trace_oracle as sph
from tests.conftest import load_golden
from tests.test_asphere_oracle import _asphere_problem
from torchoptics_b200 import ops
from torchoptics_b200 import ray_tracing_lite as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)


def _to(p, dev, grad=()):
    q = {k: v.to(dev) for k, v in p.items()}
    for k in grad:
        q[k] = q[k].clone().requires_grad_(True)
    return q


def _problem(n=400, clip=True):
    p = _asphere_problem(torch.float32, n=n)
    sd = torch.full_like(p['c'], float('inf'))
    if clip:
        sd[..., 2] = 1.7
    p['sd'] = sd
    return p


def _call(fn, q, **kw):
    return fn(q['x'], q['y'], q['z'], q['cx'], q['cy'], q['c'], q['t'], q['mu'], q['mask'], **kw)


def test_forward_exact_policy_bit_identical_incl_opl_and_clip():
    p = _problem()
    with sph.ieee_sqrt():
        ref = _call(gen.trace, p, k=p['k'], a=p['a'], sd=p['sd'])
    assert 0 < int(ref[4].sum()) < ref[4].numel()
    q = _to(p, DEV)
    out = _call(rt.trace_skew, q, arith='exact', k=q['k'], a=q['a'], sd=q['sd'])
    assert len(out) == 7
    for j in (0, 1, 2, 3, 6):
        assert np.array_equal(out[j].cpu().numpy().view(np.uint32), ref[j].numpy().view(np.uint32)), j
    assert torch.equal(out[4].cpu(), ref[4]) and torch.equal(out[5].cpu(), ref[5])


def test_forward_guarded_policy():
    p = _problem()
    ref = _call(gen.trace, p, k=p['k'], a=p['a'], sd=p['sd'])
    q = _to(p, DEV)
    out = _call(rt.trace_skew, q, k=q['k'], a=q['a'], sd=q['sd'])
    assert torch.equal(out[4].cpu(), ref[4]) and torch.equal(out[5].cpu(), ref[5])
    scale = float(ref[1].abs().max())
    for j, tol in ((0, 1e-5 * scale), (1, 1e-5 * scale), (2, 1e-5), (3, 1e-5), (6, 1e-5 * float(ref[6].max()))):
        assert float((out[j].cpu() - ref[j]).abs().max()) <= tol, j
    dead = ~ref[4]
    for j in range(4):
        assert not out[j].cpu()[dead].any()


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
@pytest.mark.parametrize('clip', [False, True])
def test_fused_spot_pass_general_surfaces(arith, clip):
    p = _problem(n=600, clip=clip)
    wrt = ('z', 'c', 't', 'mu', 'k', 'a')
    # truth: the oracle in fp64; the fp32 oracle tells how much fp32 noise to expect
    def oracle_run(dtype):
        o = {k: (v if k == 'mask' else v.to(dtype)) for k, v in p.items()}
        for k in wrt:
            o[k] = o[k].clone().requires_grad_(True)
        out = _call(gen.trace, o, k=o['k'], a=o['a'], sd=o['sd'])
        rms = sph.spot_rms(out[0], out[1], out[4])
        return rms, torch.autograd.grad(rms, [o[k] for k in wrt])
    ref_rms, ref = oracle_run(torch.float64)
    q = _to(p, DEV, grad=wrt)
    rms, field = _call(ops.spot_rms, q, arith=rt._arith_code(arith), k=q['k'], a=q['a'], sd=q['sd'])
    assert abs(rms[0].item() - ref_rms.item()) <= 2e-5 * ref_rms.item(), (rms[0].item(), ref_rms.item())
    got = torch.autograd.grad(rms[0], [q[k] for k in wrt])
    for name, g, r in zip(wrt, got, ref):
        assert g.shape == r.shape, name
        g, r = g.cpu().numpy(), r.numpy()
        if name == 'z':
            scale = max(abs(float(r.ravel()[0])), float(np.abs(ref[2].numpy()).max()))
            assert abs(float(g.ravel()[0]) - float(r.ravel()[0])) <= 2e-4 * scale
        elif name == 'a':
            for i in range(7):      # coefficients of rho^2 .. rho^8 live on very different scales
                assert _rel(g[..., i], r[..., i]) <= 2e-4, (name, i, _rel(g[..., i], r[..., i]))
        else:
            assert _rel(g, r) <= 2e-4, (name, _rel(g, r))


def test_general_kernels_reduce_to_the_spherical_ones():
    """k = 0, a = 0, sd = inf through the general kernels == the spherical kernels."""
    rec = load_golden('cooke_32x32')
    i = {k[3:]: torch.from_numpy(rec[k]).to(DEV) for k in rec if k.startswith('in_')}
    for k in ('z', 'c', 't', 'mu'):
        i[k] = i[k].clone().requires_grad_(True)
    args = [i[k] for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'mask')]
    rms_s, _ = ops.spot_rms(*args)
    g_s = torch.autograd.grad(rms_s[0], [i[k] for k in ('c', 't', 'mu')])
    zeros = torch.zeros_like(i['c'].detach())
    rms_g, _ = ops.spot_rms(*args, k=zeros)
    g_g = torch.autograd.grad(rms_g[0], [i[k] for k in ('c', 't', 'mu')])
    assert abs(rms_g[0].item() - rms_s[0].item()) <= 1e-5 * rms_s[0].item()
    for a_, b_ in zip(g_g, g_s):
        assert _rel(a_.cpu().numpy(), b_.cpu().numpy()) <= 1e-4
    out_s = rt.trace_skew(*[a_.detach() for a_ in args])
    out_g = rt.trace_skew(*[a_.detach() for a_ in args], k=zeros)
    assert torch.equal(out_s[4], out_g[4]) and torch.equal(out_s[5], out_g[5])
    assert float((out_s[1] - out_g[1]).abs().max()) <= 1e-5 * float(out_s[1].abs().max())


def test_config3_lens_at_3m_rays_against_the_oracle_on_device():
    """The synthetic 12-surface asphere lens (BASELINE config 3 shape) at 3.1 M rays: masks of the
    exact policy identical to the extension oracle evaluated by torch on the same GPU, fused RMS and
    its gradients within tolerance, >= 99 % of the rays traced."""
    from torchoptics_b200 import prescriptions
    # masks on the f/4 variant (0.3 % of the rays fail or are flagged) ...
    specs4, lens4 = prescriptions.asphere_12(DEV, f_number=4.0)
    tracer4 = rt.RayTracer(mode='circular', n_rays=(128, 128), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                           wavelengths=('C', 'd', 'F'), default_device=DEV)
    args4 = tracer4._ray_set(specs4, lens4)
    ext4 = tracer4._extension_tables(lens4)
    ref4 = gen.trace(*args4, **ext4)
    assert 0.99 <= float(ref4[4].float().mean()) < 1.0
    for policy in ('exact', 'guarded'):
        out4 = rt.trace_skew(*args4, arith=policy, **ext4)
        assert torch.equal(out4[4], ref4[4]) and torch.equal(out4[5], ref4[5]), policy
    # ... values and gradients on the default f/5 lens, where every ray traces
    specs, lens = prescriptions.asphere_12(DEV)
    for name in ('c', 't', 'k', 'a'):
        getattr(lens, name).requires_grad_(True)
    tracer = rt.RayTracer(mode='circular', n_rays=(256, 256), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    rms, _ = tracer.spot_rms(specs, lens)
    grads = torch.autograd.grad(rms[0], [lens.c, lens.t, lens.k, lens.a])
    args = tracer._ray_set(specs, lens)
    ext = tracer._extension_tables(lens)
    ref = gen.trace(*args, **ext)
    assert float(ref[4].float().mean()) >= 0.99
    out = rt.trace_skew(*[a_.detach() for a_ in args], arith='exact', **{k: v.detach() for k, v in ext.items() if v is not None})
    assert torch.equal(out[4], ref[4]) and torch.equal(out[5], ref[5])
    # truth for the values: the same oracle in fp64 on the device
    dbl = [a_.detach().double() if a_.is_floating_point() else a_ for a_ in args]
    leaves = {n_: getattr(lens, n_).detach().double().requires_grad_(True) for n_ in ('c', 't', 'k', 'a')}
    dbl[5] = leaves['c'].reshape(1, 1, 1, 1, -1)
    dbl[6] = leaves['t'].reshape(1, 1, 1, 1, -1)
    ref64 = gen.trace(*dbl, k=leaves['k'].reshape(1, 1, 1, 1, -1), a=leaves['a'].reshape(1, 1, 1, 1, 12, 7))
    assert torch.equal(ref64[4], ref[4])
    ref_rms = sph.spot_rms_all_lenses(ref64[1], ref64[4])[0]
    assert abs(rms[0].item() - ref_rms.item()) <= 2e-5 * ref_rms.item()
    ref_grads = torch.autograd.grad(ref_rms, [leaves[n_] for n_ in ('c', 't', 'k', 'a')])
    for name, g, r in zip(('c', 't', 'k'), grads, ref_grads):
        err = _rel(g.cpu().numpy(), r.cpu().numpy())
        assert err <= 2e-4, (name, err)
    for i in range(7):
        err = _rel(grads[3][..., i].cpu().numpy(), ref_grads[3][..., i].cpu().numpy())
        assert err <= 2e-4, ('a', i, err)


def test_adam_loop_reduces_the_spot_size():
    """Config-5 shape in miniature: Adam on c, t, k, a of the asphere lens through the fused pass."""
    from torchoptics_b200 import prescriptions
    from torchoptics_b200.optimize import optimize_spot
    specs, lens = prescriptions.asphere_12(DEV)
    tracer = rt.RayTracer(mode='circular', n_rays=(48, 48), rel_fields=tuple(np.linspace(0, 1, 8).tolist()),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    best, history = optimize_spot(tracer, specs, lens, steps=60, lr=5e-5)
    assert np.isfinite(history).all()
    assert min(history[-5:]) < 0.9 * history[0], (history[0], history[-5:])
    rms, _ = tracer.spot_rms(specs, best)
    assert abs(rms[0].item() - history[-1]) < 0.2 * history[0]
    assert best.a.shape == lens.a.shape and not torch.equal(best.k, lens.k)


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
def test_split_backward_general_surfaces(arith):
    """trace_skew with extension tables is differentiable w.r.t. x, y, z, cx, cy, c, t, mu, k, a."""
    p = _problem(n=300, clip=True)
    wrt = ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'k', 'a')
    gen_ = torch.Generator().manual_seed(2)
    o = {k: (v if k == 'mask' else v.double()) for k, v in p.items()}
    for k in wrt:
        o[k] = o[k].clone().requires_grad_(True)
    ref_out = _call(gen.trace, o, k=o['k'], a=o['a'], sd=o['sd'])
    seeds = [0.5 + torch.rand(ref_out[4].shape, generator=gen_) for _ in range(4)]
    ref_loss = sum((s.double() * v).sum() for s, v in zip(seeds, ref_out[:4]))
    ref = torch.autograd.grad(ref_loss, [o[k] for k in wrt])
    q = _to(p, DEV, grad=wrt)
    out = _call(rt.trace_skew, q, arith=arith, k=q['k'], a=q['a'], sd=q['sd'])
    loss = sum((s.to(DEV) * v).sum() for s, v in zip(seeds, out[:4]))
    got = torch.autograd.grad(loss, [q[k] for k in wrt])
    for name, g, r in zip(wrt, got, ref):
        assert g.shape == r.shape, name
        g, r = g.cpu().numpy(), r.numpy()
        if name == 'a':
            for i in range(7):
                assert _rel(g[..., i], r[..., i]) <= 2e-4, (name, i, _rel(g[..., i], r[..., i]))
        else:
            assert _rel(g, r) <= 2e-4, (name, _rel(g, r))


def _oracle_path_run(p, wrt, dtype, seed_opl, seed_opd, radius):
    """Loss = sum(seed_opl * opl) + sum(seed_opd * opd) on the extension oracle; returns opl, opd and gradients."""
    o = {k: (v if k == 'mask' else v.to(dtype)) for k, v in p.items()}
    for k in wrt:
        o[k] = o[k].clone().requires_grad_(True)
    out = _call(gen.trace, o, k=o['k'], a=o['a'], sd=o['sd'])
    n_image = (1 / o['mu']).prod(-1)
    opd = gen.opd(out[0], out[1], out[2], out[3], out[6], out[4], n_image, radius)
    loss = (seed_opl.to(dtype) * out[6]).sum() + (seed_opd.to(dtype) * opd).sum()
    return out, opd, torch.autograd.grad(loss, [o[k] for k in wrt])


@pytest.mark.parametrize('arith', ['guarded', 'exact'])
def test_optical_path_and_opd_values_and_gradients(arith):
    """Row A10: the optical path length is differentiable (TlSeeds.gopl) and compute_opd subtracts the reference
    sphere.  Truth = the extension oracle in fp64 (its OPD function has a closed-form test of its own); the fp32
    oracle next to it shows the fp32 noise of the same formulas.  Every ray of this problem traces."""
    p = _problem(n=600, clip=False)
    wrt = ('z', 'c', 't', 'mu', 'k', 'a')
    radius = 40.0
    gen_ = torch.Generator().manual_seed(11)
    shape = (1, 2, 600, 2)
    seed_opl = torch.randn(shape, generator=gen_)
    seed_opd = torch.randn(shape, generator=gen_)
    out64, opd64, ref = _oracle_path_run(p, wrt, torch.float64, seed_opl, seed_opd, radius)
    out32, opd32, ref32 = _oracle_path_run(p, wrt, torch.float32, seed_opl, seed_opd, radius)
    assert bool(out64[4].all())

    q = _to(p, DEV, grad=wrt)
    out = _call(rt.trace_skew, q, arith=arith, k=q['k'], a=q['a'], sd=q['sd'])
    assert out[6].requires_grad
    opd = rt.compute_opd(out[0], out[1], out[2], out[3], out[6], out[4], q['mu'], radius)
    # values: the path to 1e-6 of itself, the OPD to the same ABSOLUTE size (a difference of two paths)
    path_scale = float(out64[6].detach().abs().max())
    err_opl = float((out[6].detach().cpu().double() - out64[6]).abs().max())
    err_opd = float((opd.detach().cpu().double() - opd64).abs().max())
    noise_opd = float((opd32.double() - opd64).abs().max())
    print(f'opl err {err_opl:.3e} of {path_scale:.3f}; opd err {err_opd:.3e} (fp32 oracle: {noise_opd:.3e}), '
          f'opd range {float(opd64.abs().max()):.3e}')
    assert err_opl <= 2e-6 * path_scale
    assert err_opd <= max(2.0 * noise_opd, 2e-6 * path_scale)
    loss = (seed_opl.to(DEV) * out[6]).sum() + (seed_opd.to(DEV) * opd).sum()
    got = torch.autograd.grad(loss, [q[k] for k in wrt])
    for name, g, r, r32 in zip(wrt, got, ref, ref32):
        assert g.shape == r.shape, name
        g, r, r32 = g.cpu().numpy(), r.numpy(), r32.numpy()
        if name == 'a':
            for i in range(7):
                ours, theirs = _rel(g[..., i], r[..., i]), _rel(r32[..., i], r[..., i])
                assert ours <= max(2e-4, 2.0 * theirs), (name, i, ours, theirs)
        else:
            ours, theirs = _rel(g, r), _rel(r32, r)
            assert ours <= max(2e-4, 2.0 * theirs), (name, ours, theirs)


def test_no_seed_on_the_path_of_a_spherical_lens():
    """TlSeeds.gopl without extension tables is refused by the C ABI (a spherical trace has no opl output)."""
    import ctypes
    from torchoptics_b200 import _native as nat
    rec = load_golden('cooke_8x8')
    i = {k[3:]: torch.from_numpy(rec[k]).to(DEV) for k in rec if k.startswith('in_')}
    lay = ops._Layout(*[i[k] for k in ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'mask')])
    pb = lay.problem(True, nat.ARITH_GUARDED)
    seed = torch.ones(lay.shape, device=DEV)
    sd = nat.TlSeeds(None, None, None, None, None, None, None, seed.data_ptr())
    gc = torch.empty((lay.B, lay.S), device=DEV)
    gt, gz = torch.empty_like(gc), torch.empty((lay.B,), device=DEV)
    gmu = torch.empty((lay.B, lay.W, lay.S), device=DEV)
    gr = nat.TlGrads(gc.data_ptr(), gt.data_ptr(), gmu.data_ptr(), gz.data_ptr(), *([None] * 7))
    lib = nat.load()
    n = lib.tl_trace_bwd_workspace(ctypes.byref(pb))
    ws = torch.empty((n // 8,), dtype=torch.float64, device=DEV)
    rc = lib.tl_trace_bwd(ctypes.byref(pb), ctypes.byref(sd), ctypes.byref(gr), ws.data_ptr(), n, nat.stream_ptr(DEV))
    assert rc != 0 and b"general-surface lenses only" in lib.tl_last_error()


def test_newton_early_exit_matches_the_four_fixed_steps(f_number=5.0):
    """The fast policy leaves the Newton loop once a step is <= 1e-3 |tau| (csrc/trace_core_asph.cuh:
    newton_settled); the exact policy and the oracle run the four fixed steps.  On the config-3 lens -- eight aspheric surfaces (the last takes a first
    step of 0.18 |tau|) and four spherical ones -- the yardstick is the fp64 oracle: masks identical, and
    points, cosines and optical path of the early-exit policy within the north-star 1e-5 of their scale
    AND no further from fp64 than 1.5 x the four-step fp32 policy's own distance (this lens carries
    ~7e-6 of fp32 noise in y whichever policy runs, so the two fp32 policies differ from EACH OTHER by
    about the sum of both).  All three distances are printed.  (The f/4 variant, whose masks both policies
    reproduce in test_config3_lens_at_3m_rays_against_the_oracle_on_device, is no yardstick for VALUES: its
    rays within ~1e-3 of a miss are ill-conditioned in fp32 whatever the iteration count)"""
    from torchoptics_b200 import prescriptions
    specs, lens = prescriptions.asphere_12(DEV, f_number=f_number)
    tracer = rt.RayTracer(mode='circular', n_rays=(96, 96), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                          wavelengths=('C', 'd', 'F'), default_device=DEV)
    args = [a_.detach() for a_ in tracer._ray_set(specs, lens)]
    ext = {k: v.detach() for k, v in tracer._extension_tables(lens).items() if v is not None}
    exact = rt.trace_skew(*args, arith='exact', **ext)
    fast = rt.trace_skew(*args, arith='guarded', **ext)
    dbl = [a_.double() if a_.is_floating_point() else a_ for a_ in args]
    ref64 = gen.trace(*dbl, **{k: (v.double() if v.is_floating_point() else v) for k, v in ext.items()})
    assert torch.equal(fast[4], exact[4]) and torch.equal(fast[5], exact[5])
    ok = exact[4] & ref64[4]
    assert float(ok.float().mean()) >= 0.99
    xy_scale = max(float(ref64[0].abs().max()), float(ref64[1].abs().max()))
    for j, name in ((0, 'x'), (1, 'y'), (2, 'cx'), (3, 'cy'), (6, 'opl')):
        scale = 1.0 if name in ('cx', 'cy') else (xy_scale if j < 2 else float(ref64[j].abs().max()))
        d_fast_exact = float((fast[j] - exact[j])[ok].abs().max())
        err_fast = float((fast[j].double() - ref64[j])[ok].abs().max())
        err_exact = float((exact[j].double() - ref64[j])[ok].abs().max())
        print(f'f/{f_number} {name}: fast-exact {d_fast_exact / scale:.2e}, fast-fp64 {err_fast / scale:.2e}, '
              f'exact-fp64 {err_exact / scale:.2e} (of scale {scale:.3g})')
        assert err_fast <= 1e-5 * scale, name
        assert err_fast <= max(1.5 * err_exact, 1e-6 * scale), name


def test_four_and_two_rays_per_thread_agree(monkeypatch):
    """The fused general-surface pass runs four rays per thread when two CTAs of it fit an SM (up to 12
    surfaces), two otherwise (plan_gen, csrc/trace_kernels.cu); TL_GEN_LANES=2 forces the latter.  Same
    rays, same arithmetic per ray, different summation order: moments and gradients agree to fp32
    summation noise, on the config-3 lens and on a small clipped problem whose rows are little longer than
    one 512-ray group."""
    from torchoptics_b200 import prescriptions

    def run_config3():
        specs, lens = prescriptions.asphere_12(DEV)
        for name in ('c', 't', 'k', 'a'):
            getattr(lens, name).requires_grad_(True)
        tracer = rt.RayTracer(mode='circular', n_rays=(64, 64), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                              wavelengths=('C', 'd', 'F'), default_device=DEV)
        rms, _ = tracer.spot_rms(specs, lens)
        return rms[0].detach(), torch.autograd.grad(rms[0], [lens.c, lens.t, lens.k, lens.a])

    def run_small():
        q = _to(_problem(n=600, clip=True), DEV, grad=('z', 'c', 't', 'mu', 'k', 'a'))
        rms, _ = _call(ops.spot_rms, q, arith=rt._arith_code('guarded'), k=q['k'], a=q['a'], sd=q['sd'])
        return rms[0].detach(), torch.autograd.grad(rms[0], [q[k] for k in ('z', 'c', 't', 'mu', 'k', 'a')])

    for run in (run_config3, run_small):
        rms4, g4 = run()
        monkeypatch.setenv('TL_GEN_LANES', '2')
        rms2, g2 = run()
        monkeypatch.delenv('TL_GEN_LANES')
        assert abs(rms4.item() - rms2.item()) <= 2e-6 * abs(rms2.item()), run.__name__
        for j, (a_, b_) in enumerate(zip(g4, g2)):
            a_, b_ = a_.cpu().numpy(), b_.cpu().numpy()
            if j == len(g4) - 1:          # the asphere coefficients: a4 ... a16 live on very different scales
                for i in range(7):
                    assert _rel(a_[..., i], b_[..., i]) <= 2e-5, (run.__name__, 'a', i)
            else:
                assert _rel(a_, b_) <= 2e-5, (run.__name__, j)
