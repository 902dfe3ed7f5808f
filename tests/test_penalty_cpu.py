"""aggregate=True penalty stacks (rtl:641-657) on the CPU:

* the oracle reproduces the reference's stacks bit for bit and its penalty / loss gradients
  (NaN pattern included) on the golden vectors of tests/golden/aggregate,
* the exact-policy arithmetic of csrc/trace_core.cuh (compiled by tests/hostcore) gives the same
  stacks, and the fast policy + `sweep_sphere_pen` adjoint match autograd of the oracle in fp64.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import trace_oracle as oracle
from tests.conftest import GOLDEN_DIR, load_golden
from tests.hostcore import binding as hc

AGG_DIR = os.path.join(GOLDEN_DIR, 'aggregate')
AGG_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(AGG_DIR, '*.npz')))
KEYS = ('z_RELU', 'theta_norm', 'theta_prime_norm')


def load_aggregate(name):
    with np.load(os.path.join(AGG_DIR, name + '.npz')) as z:
        agg = {k: z[k] for k in z.files}
    return load_golden(name), agg


def full_inputs(rec, grad=False, dtype=torch.float32):
    shape = rec['out_ok'].shape
    ins = {k: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(rec['in_' + k], shape))).to(dtype)
           for k in ('x', 'y', 'cx', 'cy')}
    for k in ('z', 'c', 't', 'mu'):
        ins[k] = torch.from_numpy(rec['in_' + k]).to(dtype).clone().requires_grad_(grad)
    ins['mask'] = torch.from_numpy(rec['in_mask'])
    return ins


def oracle_penalty(i, allow, shape):
    out = oracle.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'], i['mu'], i['mask'],
                       True, allow)
    stacks = {k: torch.stack([torch.broadcast_to(s, shape) for s in out[6][k]]) for k in KEYS}
    n_seq = int(i['mask'].sum())
    q = (stacks['theta_norm'].sum(0) + stacks['theta_prime_norm'].sum(0) + stacks['z_RELU'].sum(0)) / n_seq
    q = torch.where(torch.isnan(q), torch.zeros_like(q), q)
    return out, stacks, q.sum()


@pytest.mark.parametrize('name', AGG_CASES)
def test_oracle_reproduces_reference_stacks_and_gradients(name):
    rec, agg = load_aggregate(name)
    i = full_inputs(rec, grad=True)
    out, stacks, penalty = oracle_penalty(i, bool(rec['allow_backward_rays']), rec['out_ok'].shape)
    for k in KEYS:
        assert np.array_equal(stacks[k].detach().numpy().view(np.uint32), agg[k].view(np.uint32)), k
    assert float(penalty) == float(agg['penalty'])
    rms = oracle.spot_rms(out[0], out[1], out[4])
    leaves = [i[k] for k in ('z', 'c', 't', 'mu')]
    g_pen = torch.autograd.grad(penalty, leaves, retain_graph=True)
    g_loss = torch.autograd.grad(rms + 0.2 * penalty, leaves)
    for k, gp, gl in zip(('z', 'c', 't', 'mu'), g_pen, g_loss):
        np.testing.assert_allclose(gp.numpy(), agg['gpen_' + k], rtol=1e-6, atol=1e-6, equal_nan=True)
        np.testing.assert_allclose(gl.numpy(), agg['gloss_' + k], rtol=1e-6, atol=1e-6, equal_nan=True)


@pytest.mark.parametrize('name', AGG_CASES)
def test_finite_gradient_switch_keeps_values_and_ok_ray_gradients(name):
    rec, agg = load_aggregate(name)
    i = full_inputs(rec, grad=True)
    with oracle.finite_penalty_gradients():
        _, stacks, penalty = oracle_penalty(i, bool(rec['allow_backward_rays']), rec['out_ok'].shape)
    for k in KEYS:
        assert np.array_equal(stacks[k].detach().numpy().view(np.uint32), agg[k].view(np.uint32)), k
    grads = torch.autograd.grad(penalty, [i[k] for k in ('z', 'c', 't', 'mu')])
    for k, g in zip(('z', 'c', 't', 'mu'), grads):
        assert torch.isfinite(g).all()
        want = agg['gpen_' + k]
        keep = np.isfinite(want)
        if bool(rec['out_ok'].all()):
            assert keep.all()
        # where the reference's own gradient is finite and no ray failed, the switch changes nothing
        if keep.all():
            np.testing.assert_allclose(g.numpy(), want, rtol=1e-6, atol=1e-6)


def _per_wavelength(rec):
    shape = rec['out_ok'].shape
    full = {k: np.broadcast_to(rec['in_' + k], shape) for k in ('x', 'y', 'z', 'cx', 'cy')}
    for w in range(shape[3]):
        rays = {k: np.ascontiguousarray(v[0, :, :, w]).ravel() for k, v in full.items()}
        yield w, rays, rec['in_c'][0, 0, 0, 0], rec['in_t'][0, 0, 0, 0], rec['in_mu'][0, 0, 0, w], \
            rec['in_mask'][0, 0, 0, 0]


@pytest.mark.parametrize('name', AGG_CASES)
def test_exact_policy_stacks_match_oracle(name):
    rec, agg = load_aggregate(name)
    allow = bool(rec['allow_backward_rays'])
    i = full_inputs(rec)
    with oracle.ieee_sqrt():
        out, stacks, _ = oracle_penalty(i, allow, rec['out_ok'].shape)
    S = rec['in_t'].shape[-1]
    for w, rays, c, t, mu, live in _per_wavelength(rec):
        (zr, th, thp), ok, bits = hc.trace_exact_pen(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'],
                                                     c, t, mu, live, allow)
        want = {k: stacks[k][:, 0, :, :, w].reshape(S, -1).numpy() for k in KEYS}
        assert np.array_equal(zr.view(np.uint32), want['z_RELU'].view(np.uint32))
        # angles: same cos^2 bit for bit, acos from two different libms (<= a few ULP)
        np.testing.assert_allclose(th, want['theta_norm'], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(thp, want['theta_prime_norm'], rtol=1e-6, atol=1e-7)
        assert np.array_equal(ok.astype(bool), out[4][0, :, :, w].reshape(-1).numpy())
        # ok bits: a ray is "ok behind surface k" exactly where its angles were not overwritten
        # with 1 (an ok ray's theta_prime is < 1: cos(theta') > sqrt(1e-6))
        for k in range(S):
            ok_k = ((bits >> k) & 1).astype(bool)
            assert np.array_equal(ok_k, want['theta_prime_norm'][k] < 1.0)
        # against the reference's own (CPU sqrt) stacks: within the float32 budget
        np.testing.assert_allclose(zr, agg['z_RELU'][:, 0, :, :, w].reshape(S, -1), rtol=0, atol=1e-5)


@pytest.mark.parametrize('name', ['singlet_8x8', 'cooke_8x8', 'tessar_8x8', 'cooke_8x8_aimed'])
def test_fast_policy_and_penalty_adjoint_match_autograd_fp64(name):
    rec, _ = load_aggregate(name)
    rng = np.random.default_rng(1)
    S = rec['in_t'].shape[-1]
    for w, rays, c, t, mu, live in _per_wavelength(rec):
        n = min(150, rays['x'].size)
        rays = {k: v[:n].astype(np.float64) for k, v in rays.items()}
        seeds = [rng.uniform(0.2, 1.0, (S, n)) for _ in range(3)]
        seed_y = rng.standard_normal(n)
        c64, t64, mu64 = (v.astype(np.float64) for v in (c, t, mu))
        t64 = t64.copy()
        t64[1] = -abs(t64[1]) * 0.05          # a negative gap: its z_RELU is active on part of the pupil
        r = hc.fast_pen(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c64, t64,
                        mu64, seed_y, *seeds)
        ti = {k: torch.tensor(v.reshape(1, 1, -1, 1), requires_grad=True) for k, v in rays.items()}
        tc = torch.tensor(c64.reshape(1, 1, 1, 1, -1), requires_grad=True)
        tt = torch.tensor(t64.reshape(1, 1, 1, 1, -1), requires_grad=True)
        tmu = torch.tensor(mu64.reshape(1, 1, 1, 1, -1), requires_grad=True)
        tmask = torch.tensor(live.reshape(1, 1, 1, 1, -1))
        out = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask, True)
        assert bool(out[4].all())
        stacks = {k: torch.stack(out[6][k]).reshape(S, n) for k in KEYS}
        for key, mine in (('z_RELU', 'z_relu'), ('theta_norm', 'theta'), ('theta_prime_norm', 'theta_prime')):
            assert np.abs(r[mine] - stacks[key].detach().numpy()).max() < 1e-7, key   # fp32 clamp constant
        assert (stacks['z_RELU'] > 0).any()
        loss = (torch.tensor(seed_y.reshape(1, 1, -1, 1)) * out[1]).sum()
        for key, sd in zip(KEYS, seeds):
            loss = loss + (torch.tensor(sd) * stacks[key]).sum()
        g = torch.autograd.grad(loss, [ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu])
        for got, want, label in zip((r['gx'], r['gy'], r['gz'], r['gcx'], r['gcy'], r['gc'], r['gt'], r['gmu']),
                                    g, ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu')):
            want = want.numpy().ravel()
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-3)
            assert err < 1e-8, (label, err)


@pytest.mark.parametrize('name', ['cooke_16x16_epd2.6', 'cooke_16x16_epd2.6_nobackward', 'tessar_16x16_epd2.0'])
def test_penalty_backward_with_failing_rays_matches_finite_oracle(name):
    """The penalty backward as the kernels run it for exact-policy rays (per-surface ok bits, failed
    lanes forced to zero, hits beyond the sphere's equator on the other branch of the sag), in fp64
    on the fp32-traced states, against autograd of the oracle with finite penalty gradients."""
    rec, agg = load_aggregate(name)
    S = rec['in_t'].shape[-1]
    n_seq = int(agg['n_seq'])
    allow = bool(rec['allow_backward_rays'])
    i = full_inputs(rec, grad=True, dtype=torch.float64)
    with oracle.finite_penalty_gradients():
        out, _, penalty = oracle_penalty(i, allow, rec['out_ok'].shape)
    assert torch.equal(out[4], torch.from_numpy(rec['out_ok']))
    assert not bool(out[4].all())
    want = torch.autograd.grad(penalty, [i[k] for k in ('z', 'c', 't', 'mu')])
    gz, gc, gt, gmu = 0.0, np.zeros(S), np.zeros(S), []
    for w, rays, c, t, mu, live in _per_wavelength(rec):
        n = rays['x'].size
        seed = np.full((S, n), 1.0 / n_seq)
        r = hc.exact_pen_adjoint(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu,
                                 live, allow, np.zeros(n), seed, seed, seed)
        gz += r['gz'].sum()
        gc += r['gc']
        gt += r['gt']
        gmu.append(r['gmu'])
    for got, ref in ((np.array([gz]), want[0]), (gc, want[1]), (gt, want[2]), (np.stack(gmu), want[3])):
        ref = ref.numpy().reshape(got.shape)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-5      # fp32 states, fp64 sweep


@pytest.mark.parametrize('seed', range(16))
def test_penalty_backward_on_random_systems(seed):
    """The kernels' penalty backward (ok bits per surface, failed lanes forced to zero, equator
    branch) on random systems with wild ray bundles, in fp64 on the fp32-traced states, against
    autograd of the oracle with finite penalty gradients.  Seeds whose fp64 oracle takes another
    branch than the fp32 trace for some ray (a ray on a threshold) compare the remaining rays."""
    from tests.test_core_cpu import _random_problem
    rng = np.random.default_rng(2000 + seed)
    n_surf = int(rng.integers(2, 8))
    rays, c, t, mu, live = _random_problem(rng, n_surf, 200)
    n = rays['x'].size
    ti = {k: torch.from_numpy(v.astype(np.float64)).reshape(1, 1, -1, 1).requires_grad_(True) for k, v in rays.items()}
    tc, tt, tmu = (torch.from_numpy(v.astype(np.float64)).reshape(1, 1, 1, 1, -1).requires_grad_(True)
                   for v in (c, t, mu))
    tmask = torch.from_numpy(live).reshape(1, 1, 1, 1, -1)
    with oracle.finite_penalty_gradients(), oracle.fp32_clamp_bound():
        ref = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask, True)
    (zr, th, thp), ok32, bits = hc.trace_exact_pen(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t,
                                                   mu, live, True)
    thp64 = torch.stack(ref[6]['theta_prime_norm']).reshape(n_surf, n).detach().numpy()
    same = np.all((thp64 < 1.0) == (thp < 1.0), axis=0)          # same ok history in fp32 and fp64
    # keep away from the clamp window and from grazing incidence, where fp32 states limit the match
    th64 = torch.stack(ref[6]['theta_norm']).reshape(n_surf, n).detach().numpy()
    calm = np.all((np.minimum(th64, thp64) > 2e-3) & (np.maximum(np.where(th64 < 1, th64, 0),
                                                                np.where(thp64 < 1, thp64, 0)) < 0.93), axis=0)
    keep = same & calm
    assert keep.sum() >= 20
    seed_w = np.zeros(n)
    seed_w[keep] = 1.0
    seeds = np.tile(seed_w, (n_surf, 1))
    r = hc.exact_pen_adjoint(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live,
                             True, np.zeros(n), seeds, seeds, seeds)
    w = torch.from_numpy(seeds)
    loss = sum((torch.stack(ref[6][k]).reshape(n_surf, n) * w).sum() for k in KEYS)
    want = torch.autograd.grad(loss, [ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu])
    for label, got, ref_g in zip(('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu'),
                                 (r['gx'], r['gy'], r['gz'], r['gcx'], r['gcy'], r['gc'], r['gt'], r['gmu']), want):
        ref_g = ref_g.numpy().ravel()
        scale = max(np.abs(ref_g).max(), 1e-6)
        assert np.abs(got - ref_g).max() <= 2e-4 * scale, (label, np.abs(got - ref_g).max() / scale)
