"""The per-ray arithmetic that the CUDA kernels execute (csrc/trace_core.cuh),
compiled for the CPU by tests/hostcore, against the oracle:

* exact policy  -> bit-identical to the golden vectors of the reference,
* fast policy   -> within 1e-5 of them on clearly-good rays,
* adjoint       -> fp64, against autograd of the oracle (~1e-10).
"""
import numpy as np
import torch

from oracle import trace_oracle as oracle
from tests.hostcore import binding as hc


def _per_wavelength(rec):
    """Yield the [B=1,F,P,W] problem of a golden record one wavelength at a time,
    every input broadcast to full ray shape."""
    shape = rec['out_ok'].shape
    full = {k: np.broadcast_to(rec['in_' + k], shape) for k in ('x', 'y', 'z', 'cx', 'cy')}
    for w in range(shape[3]):
        rays = {k: np.ascontiguousarray(v[0, :, :, w]).ravel() for k, v in full.items()}
        yield w, rays, rec['in_c'][0, 0, 0, 0], rec['in_t'][0, 0, 0, 0], rec['in_mu'][0, 0, 0, w], \
            rec['in_mask'][0, 0, 0, 0]


def test_exact_policy_is_bit_identical(golden):
    """Exact policy == oracle with the correctly rounded sqrt, bit for bit; and its
    masks == the reference's own (golden), values within the 1e-5 budget."""
    allow = bool(golden['allow_backward_rays'])
    i = {k[3:]: torch.from_numpy(golden[k]) for k in golden if k.startswith('in_')}
    with oracle.ieee_sqrt():
        ref = oracle.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'], i['mu'],
                           i['mask'], False, allow)
    ref = [torch.broadcast_to(r, ref[4].shape).numpy() for r in ref]
    scale = max(np.abs(golden['out_x']).max(), np.abs(golden['out_y']).max())
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        got = hc.trace_exact(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'],
                             c, t, mu, live, allow)
        for j, (key, tol) in enumerate((('out_x', 1e-5 * scale), ('out_y', 1e-5 * scale),
                                        ('out_cx', 1e-5), ('out_cy', 1e-5))):
            want = np.ascontiguousarray(ref[j][0, :, :, w]).ravel()
            assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32)), (key, w)
            assert np.abs(got[j] - golden[key][0, :, :, w].ravel()).max() <= tol, (key, w)
        for j, key in ((4, 'out_ok'), (5, 'out_backward')):
            assert np.array_equal(got[j].astype(bool), ref[j][0, :, :, w].ravel())
            # (the reference returns an un-broadcast all-False `ray_backward` when the flag is off)
            want = np.broadcast_to(golden[key], golden['out_ok'].shape)[0, :, :, w].ravel()
            assert np.array_equal(got[j].astype(bool), want)


def test_fast_policy_close_on_clear_rays(golden):
    scale = max(np.abs(golden['out_x']).max(), np.abs(golden['out_y']).max())
    n_clear = 0
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        r = hc.fast(np.float32, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live)
        length = np.abs(t).sum() + np.abs(rays['z']).max()
        clear = (r['min_cos2'] > 1e-6 + 1e-4) & (r['min_travel'] > 1e-5 * max(1.0, length))
        ok = golden['out_ok'][0, :, :, w].ravel()
        bw = np.broadcast_to(golden['out_backward'], golden['out_ok'].shape)[0, :, :, w].ravel()
        # a clearly-good ray must be ok and not flagged in the reference
        assert np.all(ok[clear]) and not np.any(bw[clear])
        n_clear += clear.sum()
        for key, ref_key, tol in (('x', 'out_x', 1e-5 * scale), ('y', 'out_y', 1e-5 * scale),
                                  ('cx', 'out_cx', 1e-5), ('cy', 'out_cy', 1e-5)):
            want = golden[ref_key][0, :, :, w].ravel()
            assert np.abs(r[key][clear] - want[clear]).max(initial=0.0) <= tol, (key, w)
    if golden['out_ok'].all() and not golden['out_backward'].any():
        assert n_clear == golden['out_ok'].size


def test_adjoint_matches_autograd_fp64(golden):
    rng = np.random.default_rng(0)
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        ok = golden['out_ok'][0, :, :, w].ravel()
        keep = np.nonzero(ok)[0][:200]
        if keep.size == 0:
            continue
        rays = {k: v[keep].astype(np.float64) for k, v in rays.items()}
        seeds = [rng.standard_normal(keep.size) for _ in range(4)]
        r = hc.fast(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'],
                    c.astype(np.float64), t.astype(np.float64), mu.astype(np.float64), live, seeds)
        ti = {k: torch.tensor(v.reshape(1, 1, -1, 1), requires_grad=True) for k, v in rays.items()}
        tc = torch.tensor(c.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
        tt = torch.tensor(t.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
        tmu = torch.tensor(mu.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
        tmask = torch.tensor(live.reshape(1, 1, 1, 1, -1))
        out = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask)
        assert bool(out[4].all())
        loss = sum((torch.tensor(s.reshape(1, 1, -1, 1)) * o).sum() for s, o in zip(seeds, out[:4]))
        g = torch.autograd.grad(loss, [ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu])
        for got, want in zip((r['gx'], r['gy'], r['gz'], r['gcx'], r['gcy'], r['gc'], r['gt'], r['gmu']), g):
            want = want.numpy().ravel()
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-3)
            assert err < 1e-9, err
        # the forward itself, in fp64, equals the oracle in fp64
        assert np.abs(r['y'] - out[1].detach().numpy().ravel()).max() < 1e-11


def test_forward_mode_pair_matches_autograd(golden):
    """D2 (value + derivatives along the pupil's x and y), the arithmetic of the on-device ray
    aiming (tl_aim): Jacobian d(image point)/d(pupil point) against autograd of the oracle in fp64."""
    if not golden['out_ok'].all():
        return
    shape = golden['out_ok'].shape
    full = {k: np.broadcast_to(golden['in_' + k], shape) for k in ('x', 'y', 'z', 'cx', 'cy')}
    w = 1
    rays = {k: np.ascontiguousarray(v[0, :, :, w]).ravel()[:96] for k, v in full.items()}
    c, t, mu = golden['in_c'][0, 0, 0, 0], golden['in_t'][0, 0, 0, 0], golden['in_mu'][0, 0, 0, w]
    ox, oy, jac = hc.forward_mode(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu)
    ti = {k: torch.tensor(v.astype(np.float64).reshape(1, 1, -1, 1), requires_grad=k in ('x', 'y'))
          for k, v in rays.items()}
    tc, tt, tmu = (torch.tensor(v.astype(np.float64).reshape(1, 1, 1, 1, -1)) for v in (c, t, mu))
    mask = torch.ones((1, 1, 1, 1, c.size), dtype=torch.bool)
    out = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, mask)
    scale = max(np.abs(golden['out_x']).max(), np.abs(golden['out_y']).max())
    assert np.abs(ox - out[0].detach().numpy().ravel()).max() <= 1e-5 * scale
    assert np.abs(oy - out[1].detach().numpy().ravel()).max() <= 1e-5 * scale
    want = []
    for o in (out[0], out[1]):
        gx, gy = torch.autograd.grad(o.sum(), [ti['x'], ti['y']], retain_graph=True)
        want += [gx.numpy().ravel(), gy.numpy().ravel()]
    want = np.stack(want, axis=1)                     # [n, 4]: dx/dxp, dx/dyp, dy/dxp, dy/dyp
    assert np.abs(jac - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


def _random_problem(rng, n_surf, n_rays):
    """A random spherical system (curvatures of both signs, flat and dummy surfaces, glass / air
    alternation with random indices) and a ray bundle wide enough that a good part of it misses,
    reflects totally or runs backward."""
    c = rng.uniform(-0.12, 0.12, n_surf).astype(np.float32)
    c[rng.random(n_surf) < 0.2] = 0.0
    t = rng.uniform(0.3, 6.0, n_surf).astype(np.float32)
    t[rng.random(n_surf) < 0.1] *= -0.2                      # a few negative gaps
    n = np.ones(n_surf + 1, np.float32)
    n[1:] = np.where(rng.random(n_surf) < 0.5, rng.uniform(1.45, 1.9, n_surf), 1.0).astype(np.float32)
    mu = (n[:-1] / n[1:]).astype(np.float32)
    live = (rng.random(n_surf) < 0.85)
    live[0] = True
    x = rng.uniform(-9, 9, n_rays).astype(np.float32)
    y = rng.uniform(-9, 9, n_rays).astype(np.float32)
    z = np.full(n_rays, rng.uniform(-3, 3), np.float32)
    cx = rng.uniform(-0.3, 0.3, n_rays).astype(np.float32)
    cy = rng.uniform(-0.5, 0.5, n_rays).astype(np.float32)
    return dict(x=x, y=y, z=z, cx=cx, cy=cy), c, t, mu, live


import pytest  # noqa: E402


@pytest.mark.parametrize('seed', range(24))
@pytest.mark.parametrize('allow', [True, False])
def test_exact_policy_bit_identical_on_random_systems(seed, allow):
    """Beyond the 12 golden cases: 24 random systems x {allow_backward_rays}, 300 wild rays each.
    The exact policy (masks, points, cosines, and the aggregate=True stacks) against the oracle
    with the correctly rounded sqrt, bit for bit (angles: two libms, a few ULP)."""
    rng = np.random.default_rng(1000 + seed)
    n_surf = int(rng.integers(1, 9))
    rays, c, t, mu, live = _random_problem(rng, n_surf, 300)
    ti = {k: torch.from_numpy(v).reshape(1, 1, -1, 1) for k, v in rays.items()}
    tc, tt, tmu = (torch.from_numpy(v).reshape(1, 1, 1, 1, -1) for v in (c, t, mu))
    tmask = torch.from_numpy(live).reshape(1, 1, 1, 1, -1)
    with oracle.ieee_sqrt():
        ref = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask, True, allow)
    got = hc.trace_exact(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live, allow)
    for j in range(4):
        want = ref[j].reshape(-1).numpy()
        assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32)), j
    assert np.array_equal(got[4].astype(bool), ref[4].reshape(-1).numpy())
    assert np.array_equal(got[5].astype(bool), torch.broadcast_to(ref[5], ref[4].shape).reshape(-1).numpy())
    (zr, th, thp), ok, bits = hc.trace_exact_pen(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu,
                                                 live, allow)
    stacks = {k: torch.stack(ref[6][k]).reshape(n_surf, -1).numpy() for k in ref[6]}
    assert np.array_equal(zr.view(np.uint32), stacks['z_RELU'].view(np.uint32))
    np.testing.assert_allclose(th, stacks['theta_norm'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(thp, stacks['theta_prime_norm'], rtol=1e-6, atol=1e-7)
    assert np.array_equal(ok.astype(bool), ref[4].reshape(-1).numpy())
    # the bundle really exercises the failure paths
    frac_ok = ref[4].float().mean().item()
    assert 0.0 <= frac_ok <= 1.0


@pytest.mark.parametrize('seed', range(16))
def test_fast_policy_and_adjoint_on_random_systems(seed):
    """Guarded policy on random systems: a ray the fast path calls clear (every predicate off its
    threshold by the guard band) is ok and not flagged in the oracle and within the value budget;
    its fp64 geometric adjoint equals autograd of the oracle (rays that hit a sphere beyond its
    equator excluded: the documented limit of sweep_sphere)."""
    rng = np.random.default_rng(3000 + seed)
    n_surf = int(rng.integers(1, 9))
    rays, c, t, mu, live = _random_problem(rng, n_surf, 400)
    ti = {k: torch.from_numpy(v).reshape(1, 1, -1, 1) for k, v in rays.items()}
    tc, tt, tmu = (torch.from_numpy(v).reshape(1, 1, 1, 1, -1) for v in (c, t, mu))
    tmask = torch.from_numpy(live).reshape(1, 1, 1, 1, -1)
    ref = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask)
    r = hc.fast(np.float32, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live)
    length = np.abs(t).sum() + np.abs(rays['z']).max()
    clear = (r['min_cos2'] > 1e-6 + 1e-4) & (r['min_travel'] > 1e-5 * max(1.0, length)) & \
        np.isfinite(r['x'] + r['y'] + r['cx'] + r['cy'])
    ok = ref[4].reshape(-1).numpy()
    bw = torch.broadcast_to(ref[5], ref[4].shape).reshape(-1).numpy()
    assert np.all(ok[clear]) and not np.any(bw[clear])
    if clear.sum() == 0:
        return
    scale = max(1.0, np.abs(ref[0].numpy()).max(), np.abs(ref[1].numpy()).max())
    for key, j, tol in (('x', 0, 2e-5 * scale), ('y', 1, 2e-5 * scale), ('cx', 2, 2e-5), ('cy', 3, 2e-5)):
        assert np.abs(r[key][clear] - ref[j].reshape(-1).numpy()[clear]).max() <= tol, key
    # adjoint in fp64 on (up to 120 of) the clear rays
    keep = np.nonzero(clear)[0][:120]
    r64 = {k: rays[k][keep].astype(np.float64) for k in rays}
    t64i = {k: torch.tensor(v.reshape(1, 1, -1, 1), requires_grad=True) for k, v in r64.items()}
    c64, tt64, mu64 = (torch.tensor(v.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True) for v in (c, t, mu))
    out = oracle.trace(t64i['x'], t64i['y'], t64i['z'], t64i['cx'], t64i['cy'], c64, tt64, mu64, tmask)
    seeds = [rng.standard_normal(keep.size) for _ in range(4)]
    rr = hc.fast(np.float64, r64['x'], r64['y'], r64['z'], r64['cx'], r64['cy'], c.astype(np.float64),
                 t.astype(np.float64), mu.astype(np.float64), live, seeds)
    loss = sum((torch.tensor(s.reshape(1, 1, -1, 1)) * o).sum() for s, o in zip(seeds, out[:4]))
    g = torch.autograd.grad(loss, [t64i['x'], t64i['y'], t64i['z'], t64i['cx'], t64i['cy']])
    per_ray_err = np.zeros(keep.size)
    for got, want in zip((rr['gx'], rr['gy'], rr['gz'], rr['gcx'], rr['gcy']), g):
        want = want.numpy().ravel()
        per_ray_err = np.maximum(per_ray_err, np.abs(got - want) / np.maximum(np.abs(want), 1e-3))
    bad = per_ray_err > 1e-8
    # the geometric adjoint is exact except on rays that hit a sphere beyond its equator, where it
    # is grossly off (not subtly): few such rays survive as clear rays even in these wild bundles
    assert bad.sum() <= 0.1 * keep.size, bad.sum()
    assert np.all(per_ray_err[bad] > 1e-4) if bad.any() else True


# ---------------------------------------------------------------------------
# Reversible formulation (fast_surface_rev / sweep_sphere_rev): what k_spot_rev runs
# ---------------------------------------------------------------------------
def _clear_rev(r, t, z):
    length = np.abs(t).sum() + np.abs(z).max()
    return (r['min_cos2'] > 1e-6 + 1e-4) & (r['min_travel'] > 1e-5 * max(1.0, length)) & \
        np.isfinite(r['x'] + r['y'] + r['cx'] + r['cy'])


def test_rev_fast_policy_close_on_clear_rays(golden):
    """The reversible forward (near-root marching distance, cz' from the refraction formula instead
    of a renormalising sqrt) against the reference's golden outputs, fp32."""
    scale = max(np.abs(golden['out_x']).max(), np.abs(golden['out_y']).max())
    n_clear = 0
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        r = hc.rev(np.float32, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live)
        clear = _clear_rev(r, t, rays['z'])
        ok = golden['out_ok'][0, :, :, w].ravel()
        bw = np.broadcast_to(golden['out_backward'], golden['out_ok'].shape)[0, :, :, w].ravel()
        assert np.all(ok[clear]) and not np.any(bw[clear])
        n_clear += clear.sum()
        for key, ref_key, tol in (('x', 'out_x', 1e-5 * scale), ('y', 'out_y', 1e-5 * scale),
                                  ('cx', 'out_cx', 1e-5), ('cy', 'out_cy', 1e-5)):
            want = golden[ref_key][0, :, :, w].ravel()
            assert np.abs(r[key][clear] - want[clear]).max(initial=0.0) <= tol, (key, w)
    if golden['out_ok'].all() and not golden['out_backward'].any():
        assert n_clear == golden['out_ok'].size


def _autograd_fp64(rays, c, t, mu, live, seeds):
    ti = {k: torch.tensor(v.astype(np.float64).reshape(1, 1, -1, 1), requires_grad=True) for k, v in rays.items()}
    tc = torch.tensor(c.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
    tt = torch.tensor(t.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
    tmu = torch.tensor(mu.astype(np.float64).reshape(1, 1, 1, 1, -1), requires_grad=True)
    tmask = torch.tensor(np.asarray(live).reshape(1, 1, 1, 1, -1))
    out = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask)
    loss = sum((torch.tensor(s.reshape(1, 1, -1, 1)) * o).sum() for s, o in zip(seeds, out[:4]))
    g = torch.autograd.grad(loss, [ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu])
    return out, [v.numpy().ravel() for v in g]


@pytest.mark.parametrize('two_comp', [False, True])
def test_rev_adjoint_matches_autograd_fp64(golden, two_comp):
    """Backward walk from (dist, cos, cos') -- or from (dist, cos) alone, cos' rebuilt -- in fp64 ==
    autograd of the oracle in fp64."""
    rng = np.random.default_rng(0)
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        ok = golden['out_ok'][0, :, :, w].ravel()
        keep = np.nonzero(ok)[0][:200]
        if keep.size == 0:
            continue
        rays = {k: v[keep].astype(np.float64) for k, v in rays.items()}
        seeds = [rng.standard_normal(keep.size) for _ in range(4)]
        r = hc.rev(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'],
                   c.astype(np.float64), t.astype(np.float64), mu.astype(np.float64), live, seeds,
                   two_comp=two_comp)
        out, g = _autograd_fp64(rays, c, t, mu, live, seeds)
        assert bool(out[4].all())
        for got, want in zip((r['gx'], r['gy'], r['gz'], r['gcx'], r['gcy'], r['gc'], r['gt'], r['gmu']), g):
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-3)
            assert err < 1e-9, err
        assert np.abs(r['y'] - out[1].detach().numpy().ravel()).max() < 1e-11
        assert np.abs(r['cx'] - out[2].detach().numpy().ravel()).max() < 1e-12


@pytest.mark.parametrize('two_comp', [False, True])
@pytest.mark.parametrize('exact_park', [False, True])
def test_rev_adjoint_fp32_within_budget(golden, exact_park, two_comp):
    """The same sweep in fp32 (as the kernel runs it; parked values from the fast forward or from
    the exact-policy re-trace) against fp64 autograd: parameter gradients to 1e-4 (north_star),
    positive seeds like the kernel's unit seed on y."""
    rng = np.random.default_rng(1)
    for w, rays, c, t, mu, live in _per_wavelength(golden):
        ok = golden['out_ok'][0, :, :, w].ravel()
        keep = np.nonzero(ok)[0][:400]
        if keep.size == 0:
            continue
        rays = {k: v[keep] for k, v in rays.items()}
        seeds = [np.zeros(keep.size), rng.uniform(0.5, 1.5, keep.size), np.zeros(keep.size), np.zeros(keep.size)]
        r = hc.rev(np.float32, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live,
                   [s.astype(np.float32) for s in seeds], exact_park=exact_park, two_comp=two_comp)
        out, g = _autograd_fp64(rays, c, t, mu, live, seeds)
        for name, got, want in zip(('c', 't', 'mu'), (r['gc'], r['gt'], r['gmu']), g[5:]):
            err = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-12)
            assert err < 1e-4, (name, w, err)


@pytest.mark.parametrize('seed', range(16))
def test_rev_adjoint_on_random_systems_including_equator_hits(seed):
    """Random systems with wild bundles: EVERY clear ray's reversible adjoint equals fp64 autograd of
    the oracle -- including the rays that hit a sphere beyond its equator, where sweep_sphere (which
    rebuilds n_z = +sqrt(1 - c^2 rho)) is grossly wrong: the backward walk carries the true h_z."""
    rng = np.random.default_rng(3000 + seed)
    n_surf = int(rng.integers(1, 9))
    rays, c, t, mu, live = _random_problem(rng, n_surf, 400)
    ti = {k: torch.from_numpy(v).reshape(1, 1, -1, 1) for k, v in rays.items()}
    tc, tt, tmu = (torch.from_numpy(v).reshape(1, 1, 1, 1, -1) for v in (c, t, mu))
    tmask = torch.from_numpy(live).reshape(1, 1, 1, 1, -1)
    ref = oracle.trace(ti['x'], ti['y'], ti['z'], ti['cx'], ti['cy'], tc, tt, tmu, tmask)
    r = hc.rev(np.float32, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], c, t, mu, live)
    clear = _clear_rev(r, t, rays['z'])
    ok = ref[4].reshape(-1).numpy()
    bw = torch.broadcast_to(ref[5], ref[4].shape).reshape(-1).numpy()
    assert np.all(ok[clear]) and not np.any(bw[clear])
    if clear.sum() == 0:
        return
    scale = max(1.0, np.abs(ref[0].numpy()).max(), np.abs(ref[1].numpy()).max())
    for key, j, tol in (('x', 0, 2e-5 * scale), ('y', 1, 2e-5 * scale), ('cx', 2, 2e-5), ('cy', 3, 2e-5)):
        assert np.abs(r[key][clear] - ref[j].reshape(-1).numpy()[clear]).max() <= tol, key
    keep = np.nonzero(clear)[0][:120]
    r64 = {k: rays[k][keep].astype(np.float64) for k in rays}
    seeds = [rng.standard_normal(keep.size) for _ in range(4)]
    rr = hc.rev(np.float64, r64['x'], r64['y'], r64['z'], r64['cx'], r64['cy'], c.astype(np.float64),
                t.astype(np.float64), mu.astype(np.float64), live, seeds)
    out, g = _autograd_fp64(r64, c, t, mu, live, seeds)
    for got, want in zip((rr['gx'], rr['gy'], rr['gz'], rr['gcx'], rr['gcy']), g[:5]):
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
        assert err.max() < 1e-8, err.max()
    for got, want in zip((rr['gc'], rr['gt'], rr['gmu']), g[5:]):
        assert np.abs(got - want).max() <= 1e-8 * max(1.0, np.abs(want).max())
