"""The extension oracle (conic / even asphere / clipping / OPL) has no reference to be
pinned to ("parity unpinned"), so it is checked through properties that
do not depend on it: reduction to the pinned spherical oracle, fp64 gradcheck, the stigmatic
conic (Cartesian ellipsoid) and the sag equation at every hit point."""
import numpy as np
import torch

from oracle import asphere_oracle as gen
from oracle import trace_oracle as sph


def _inputs(rec, dtype=torch.float32):
    t = {k[3:]: torch.from_numpy(rec[k]) for k in rec if k.startswith('in_')}
    return {k: (v if k == 'mask' else v.to(dtype)) for k, v in t.items()}


def test_reduces_to_the_spherical_oracle(golden):
    i = _inputs(golden)
    allow = bool(golden['allow_backward_rays'])
    want = sph.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'], i['mu'], i['mask'],
                     False, allow)
    got = gen.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'], i['mu'], i['mask'],
                    allow_backward_rays=allow)
    ok = torch.broadcast_to(want[4], got[4].shape)
    # identical masks except for rays the spherical form decides within 1e-5 of a threshold
    flips = int((got[4] != ok).sum())
    assert flips <= max(2, ok.numel() // 500), flips
    both = got[4] & ok
    scale = float(torch.broadcast_to(want[1], ok.shape)[both].abs().max())
    for g, w, tol in ((got[0], want[0], 2e-5 * scale), (got[1], want[1], 2e-5 * scale),
                      (got[2], want[2], 2e-5), (got[3], want[3], 2e-5)):
        w = torch.broadcast_to(w, ok.shape)
        assert float((g - w)[both].abs().max()) <= tol


def _asphere_problem(dtype=torch.float64, n=12):
    gen_ = torch.Generator().manual_seed(3)
    S = 4
    c = torch.tensor([0.05, -0.03, 0.04, -0.06], dtype=dtype).reshape(1, 1, 1, 1, S)
    k = torch.tensor([-0.8, 0.5, -1.2, 0.0], dtype=dtype).reshape(1, 1, 1, 1, S)
    a = torch.zeros(1, 1, 1, 1, S, 7, dtype=dtype)
    a[..., 0] = torch.tensor([2e-5, -1e-5, 3e-5, 0.0], dtype=dtype)
    a[..., 1] = torch.tensor([-1e-7, 2e-7, 0.0, 1e-7], dtype=dtype)
    a[..., 2] = 1e-10
    t = torch.tensor([3.0, 1.0, 2.5, 18.0], dtype=dtype).reshape(1, 1, 1, 1, S)
    mu = torch.tensor([[1 / 1.52, 1.52, 1 / 1.61, 1.61], [1 / 1.53, 1.53, 1 / 1.63, 1.63]],
                      dtype=dtype).reshape(1, 1, 1, 2, S)
    mask = torch.ones(1, 1, 1, 1, S, dtype=torch.bool)
    x = (torch.rand(1, 1, n, 1, generator=gen_, dtype=dtype) - 0.5) * 4
    y = (torch.rand(1, 1, n, 1, generator=gen_, dtype=dtype) - 0.5) * 4
    z = torch.tensor([-2.0], dtype=dtype).reshape(1, 1, 1, 1)
    cx = torch.tensor([0.02], dtype=dtype).reshape(1, 1, 1, 1)
    cy = torch.tensor([0.0, 0.12], dtype=dtype).reshape(1, 2, 1, 1)
    return dict(x=x, y=y, z=z, cx=cx, cy=cy, c=c, t=t, mu=mu, mask=mask, k=k, a=a)


def test_gradcheck_fp64():
    p = _asphere_problem()
    names = ('x', 'y', 'z', 'cx', 'cy', 'c', 't', 'mu', 'k', 'a')
    # the coefficients multiply rho^2 .. rho^8 (rho up to ~8 here): differentiate w.r.t. a
    # scaled copy so that gradcheck's finite step is a small change of the surface
    scale = torch.tensor([8.0 ** -i for i in range(2, 9)], dtype=torch.float64)
    p = dict(p, a=p['a'] / scale)
    leaves = [p[n].clone().requires_grad_(True) for n in names]

    def fn(*v):
        q = dict(zip(names, v))
        q['a'] = q['a'] * scale
        out = gen.trace(q['x'], q['y'], q['z'], q['cx'], q['cy'], q['c'], q['t'], q['mu'], p['mask'],
                        k=q['k'], a=q['a'])
        assert bool(out[4].all())
        return out[0], out[1], out[2], out[3], out[6]

    assert torch.autograd.gradcheck(fn, leaves, eps=1e-6, atol=1e-6, rtol=1e-5, nondet_tol=0.0)


def test_hit_points_satisfy_the_sag_equation():
    p = _asphere_problem()
    for s_i in range(4):
        cs, ks, as_ = p['c'][..., s_i], p['k'][..., s_i], p['a'][..., s_i, :]
        cz = torch.sqrt(1 - p['cx'] ** 2 - p['cy'] ** 2)
        _, tau = gen._intersect(cs, ks, as_, p['x'], p['y'], p['z'], p['cx'], p['cy'], cz)
        hx, hy, hz = p['x'] + tau * p['cx'], p['y'] + tau * p['cy'], p['z'] + tau * cz
        s, _, _ = gen.sag_and_slope(cs, ks, as_, hx * hx + hy * hy)
        assert float((hz - s).abs().max()) < 1e-12


def test_cartesian_ellipsoid_is_stigmatic_and_isochronous():
    """Collimated on-axis light refracted from air into index n by the conic k = -1/n^2
    focuses perfectly at t = R n / (n - 1), and every ray has the same optical path."""
    n_glass, radius = 1.5, 20.0
    dtype = torch.float64
    c = torch.tensor([1 / radius], dtype=dtype).reshape(1, 1, 1, 1, 1)
    k = torch.tensor([-1 / n_glass ** 2], dtype=dtype).reshape(1, 1, 1, 1, 1)
    t = torch.tensor([radius * n_glass / (n_glass - 1)], dtype=dtype).reshape(1, 1, 1, 1, 1)
    mu = torch.tensor([1 / n_glass], dtype=dtype).reshape(1, 1, 1, 1, 1)
    mask = torch.ones(1, 1, 1, 1, 1, dtype=torch.bool)
    r = torch.linspace(0, 8, 33, dtype=dtype)
    x = (r * np.cos(0.7)).reshape(1, 1, -1, 1)
    y = (r * np.sin(0.7)).reshape(1, 1, -1, 1)
    zero = torch.zeros(1, 1, 1, 1, dtype=dtype)
    out = gen.trace(x, y, zero - 5.0, zero, zero, c, t, mu, mask, k=k)
    assert bool(out[4].all())
    assert float(out[0].abs().max()) < 1e-9 and float(out[1].abs().max()) < 1e-9
    opl = out[6]
    assert float((opl - opl[0, 0, 0, 0]).abs().max()) < 1e-9
    # the sphere of the same vertex radius is not stigmatic
    sph_out = gen.trace(x, y, zero - 5.0, zero, zero, c, t, mu, mask)
    assert float(sph_out[1].abs().max()) > 1e-3


def test_semi_diameter_clip_and_parking():
    p = _asphere_problem(torch.float32, n=64)
    sd = torch.full_like(p['c'], float('inf'))
    sd[..., 1] = 2.0
    out = gen.trace(p['x'], p['y'], p['z'], p['cx'], p['cy'], p['c'], p['t'], p['mu'], p['mask'],
                    k=p['k'], a=p['a'], sd=sd)
    free = gen.trace(p['x'], p['y'], p['z'], p['cx'], p['cy'], p['c'], p['t'], p['mu'], p['mask'],
                     k=p['k'], a=p['a'])
    ok, ok_free = out[4], free[4]
    assert bool(ok_free.all()) and 0 < int(ok.sum()) < ok.numel()
    dead = ~ok
    for j in range(4):
        assert not out[j][dead].any()                  # parked rays output exact zeros
        assert torch.equal(out[j][ok], free[j][ok])    # survivors are untouched by the clip


def test_against_an_independent_bracketed_root_finder_fp64():
    """A second, independent statement of rows A9/A10 in scalar Python: the intersection by a BRACKETED root
    finder (scipy.optimize.brentq on F(tau) = z + tau cz - sag, no Newton, no base-sphere start), the normal
    from the analytic slope written with explicit powers, Snell's law in its textbook vector form
    d' = mu d + (sqrt(1 - mu^2 (1 - (n.d)^2)) - mu n.d) n, the optical path as the sum of n |segment|.
    1 000 random rays through the four-surface asphere: positions / cosines / OPL of the oracle (fp64,
    four Newton steps) agree with it to 1e-9 -- the fixed iteration count has converged and the oracle's
    conventions (shifted vertex coordinates, index bookkeeping) describe the same physical ray."""
    import math
    from scipy.optimize import brentq

    p = _asphere_problem(torch.float64, n=500)          # 500 pupil points x 2 fields x 2 wavelengths
    out = gen.trace(p['x'], p['y'], p['z'], p['cx'], p['cy'], p['c'], p['t'], p['mu'], p['mask'], k=p['k'], a=p['a'])
    assert bool(out[4].all())
    c = p['c'].reshape(-1).tolist()
    k = p['k'].reshape(-1).tolist()
    t = p['t'].reshape(-1).tolist()
    a = p['a'].reshape(4, 7).tolist()
    mu = p['mu'].reshape(2, 4).tolist()

    def sag(s_i, rho):
        base = c[s_i] * rho / (1.0 + math.sqrt(1.0 - (1.0 + k[s_i]) * c[s_i] ** 2 * rho))
        return base + sum(a[s_i][j] * rho ** (j + 2) for j in range(7))

    def slope(s_i, rho):                                  # d sag / d rho
        base = c[s_i] / (2.0 * math.sqrt(1.0 - (1.0 + k[s_i]) * c[s_i] ** 2 * rho))
        return base + sum((j + 2) * a[s_i][j] * rho ** (j + 1) for j in range(7))

    worst = 0.0
    checked = 0
    for f in range(2):
        for w in range(2):
            for q in range(0, 500, 2):                    # 250 x 4 = 1 000 rays
                pos = [float(p['x'][0, 0, q, 0]), float(p['y'][0, 0, q, 0]), float(p['z'][0, 0, 0, 0])]
                d = [float(p['cx'][0, 0, 0, 0]), float(p['cy'][0, f, 0, 0]), 0.0]
                d[2] = math.sqrt(1.0 - d[0] ** 2 - d[1] ** 2)
                index, path = 1.0, 0.0
                for s_i in range(4):
                    def gap(tau):
                        hx, hy = pos[0] + tau * d[0], pos[1] + tau * d[1]
                        return pos[2] + tau * d[2] - sag(s_i, hx * hx + hy * hy)
                    lo, hi = -4.0, (4.0 - pos[2]) / d[2]
                    assert gap(lo) < 0.0 < gap(hi)
                    tau = brentq(gap, lo, hi, xtol=1e-15, rtol=8.9e-16, maxiter=200)
                    pos = [pos[j] + tau * d[j] for j in range(3)]
                    path += index * tau
                    ds = slope(s_i, pos[0] ** 2 + pos[1] ** 2)
                    n = [-2.0 * pos[0] * ds, -2.0 * pos[1] * ds, 1.0]
                    norm = math.sqrt(sum(v * v for v in n))
                    n = [v / norm for v in n]
                    m = mu[w][s_i]
                    cos_in = sum(n[j] * d[j] for j in range(3))
                    cos_out = math.sqrt(1.0 - m * m * (1.0 - cos_in ** 2))
                    d = [m * d[j] + (cos_out - m * cos_in) * n[j] for j in range(3)]
                    norm = math.sqrt(sum(v * v for v in d))
                    d = [v / norm for v in d]
                    pos[2] -= t[s_i]
                    index /= m
                tau = -pos[2] / d[2]
                path += index * tau
                got = [float(out[j][0, f, q, w]) for j in (0, 1, 2, 3, 6)]
                want = [pos[0] + tau * d[0], pos[1] + tau * d[1], d[0], d[1], path]
                worst = max(worst, max(abs(g - v) for g, v in zip(got, want)))
                checked += 1
    assert checked == 1000
    assert worst < 1e-9, worst


def test_opd_of_a_defocused_perfect_wave_has_its_closed_form():
    """The stigmatic ellipsoid makes a perfect spherical wave converging on its focus; with the image plane moved
    delta behind the focus, the OPD against a reference sphere of radius R centred on the chief ray's image point
    is, exactly,  -n (sqrt(R^2 - delta^2 (1 - cz^2)) - delta cz - (R - delta))  for a ray of direction cosine cz
    (intersect the ray through the focus with the sphere) -- and 0 for every ray when delta = 0, whatever R."""
    n_glass, radius = 1.5, 20.0
    dtype = torch.float64
    focus = radius * n_glass / (n_glass - 1)
    c = torch.tensor([1 / radius], dtype=dtype).reshape(1, 1, 1, 1, 1)
    k = torch.tensor([-1 / n_glass ** 2], dtype=dtype).reshape(1, 1, 1, 1, 1)
    mu = torch.tensor([1 / n_glass], dtype=dtype).reshape(1, 1, 1, 1, 1)
    mask = torch.ones(1, 1, 1, 1, 1, dtype=torch.bool)
    r = torch.linspace(0, 8, 33, dtype=dtype)            # the first ray is the chief ray (on the axis)
    x = (r * np.cos(0.7)).reshape(1, 1, -1, 1)
    y = (r * np.sin(0.7)).reshape(1, 1, -1, 1)
    zero = torch.zeros(1, 1, 1, 1, dtype=dtype)
    n_image = (1 / mu).prod(-1)
    for delta in (0.0, 0.05, -0.08):
        t = torch.tensor([focus + delta], dtype=dtype).reshape(1, 1, 1, 1, 1)
        out = gen.trace(x, y, zero - 5.0, zero, zero, c, t, mu, mask, k=k)
        assert bool(out[4].all())
        for big_r in (30.0, 55.0):
            got = gen.opd(out[0], out[1], out[2], out[3], out[6], out[4], n_image, big_r)
            cz = torch.sqrt(1 - out[2] ** 2 - out[3] ** 2)
            want = -n_glass * (torch.sqrt(big_r ** 2 - delta ** 2 * (1 - cz ** 2)) - delta * cz - (big_r - delta))
            assert float((got - want).abs().max()) < 1e-10, (delta, big_r)
            if delta == 0.0:
                assert float(got.abs().max()) < 1e-10
            else:
                assert float(got.abs().max()) > 1e-5          # a real defocus term (~ n delta (1 - cz))
