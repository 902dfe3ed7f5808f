"""csrc/trace_core_asph.cuh (the per-ray arithmetic of the extension surfaces), compiled for
the CPU, against the extension oracle: exact policy bit-identical (with the correctly rounded
sqrt), fast policy and the geometric adjoint in fp64 against autograd."""
import numpy as np
import torch

from oracle import asphere_oracle as gen
from oracle import trace_oracle as sph
from tests.hostcore import binding as hc
from tests.test_asphere_oracle import _asphere_problem


def _flat(p, dtype, w):
    """One wavelength of the test problem with every ray input broadcast to [F*P]."""
    F, P = p['cy'].shape[1], p['x'].shape[2]
    shape = (1, F, P, 1)
    rays = {k: torch.broadcast_to(p[k], shape).reshape(-1).numpy().astype(dtype)
            for k in ('x', 'y', 'z', 'cx', 'cy')}
    S = p['c'].shape[-1]
    tabs = dict(c=p['c'].reshape(S).numpy(), k=p['k'].reshape(S).numpy(), a=p['a'].reshape(S, 7).numpy(),
                t=p['t'].reshape(S).numpy(), mu=p['mu'][0, 0, 0, w].numpy(),
                sd=np.full(S, np.inf), live=np.ones(S, np.uint8))
    return rays, tabs


def test_exact_policy_bit_identical_to_the_extension_oracle():
    p = _asphere_problem(torch.float32, n=96)
    sd = torch.full_like(p['c'], float('inf'))
    sd[..., 2] = 1.6
    with sph.ieee_sqrt():
        ref = gen.trace(p['x'], p['y'], p['z'], p['cx'], p['cy'], p['c'], p['t'], p['mu'], p['mask'],
                        k=p['k'], a=p['a'], sd=sd)
    assert 0 < int(ref[4].sum()) < ref[4].numel()
    for w in range(2):
        rays, tabs = _flat(p, np.float32, w)
        tabs['sd'] = sd.reshape(-1).numpy()
        got = hc.asph_exact(rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], tabs['c'], tabs['k'],
                            tabs['a'], tabs['t'], tabs['mu'], tabs['sd'], tabs['live'])
        for j in (0, 1, 2, 3, 6):
            want = np.ascontiguousarray(ref[j][0, :, :, w].numpy()).ravel()
            assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32)), j
        assert np.array_equal(got[4].astype(bool), ref[4][0, :, :, w].numpy().ravel())
        assert np.array_equal(got[5].astype(bool), ref[5][0, :, :, w].numpy().ravel())


def test_fast_policy_and_adjoint_fp64():
    p = _asphere_problem(torch.float64, n=24)
    names = ('x', 'y', 'z', 'cx', 'cy', 'c', 'k', 'a', 't', 'mu')
    rng = np.random.default_rng(5)
    for w in range(2):
        rays, tabs = _flat(p, np.float64, w)
        n = rays['x'].size
        seeds = [rng.standard_normal(n) for _ in range(5)]      # on x, y, cx, cy and on the optical path length
        r = hc.asph_fast(np.float64, rays['x'], rays['y'], rays['z'], rays['cx'], rays['cy'], tabs['c'],
                         tabs['k'], tabs['a'], tabs['t'], tabs['mu'], tabs['sd'], seeds)
        # oracle on the same flattened ray set, one wavelength
        q = {k: torch.tensor(v.reshape(1, 1, -1, 1), requires_grad=True) for k, v in rays.items()}
        S = tabs['c'].size
        tb = {k: torch.tensor(tabs[k].reshape(1, 1, 1, 1, S), requires_grad=True) for k in ('c', 'k', 't', 'mu')}
        ta = torch.tensor(tabs['a'].reshape(1, 1, 1, 1, S, 7), requires_grad=True)
        out = gen.trace(q['x'], q['y'], q['z'], q['cx'], q['cy'], tb['c'], tb['t'], tb['mu'],
                        torch.ones(1, 1, 1, 1, S, dtype=torch.bool), k=tb['k'], a=ta)
        assert bool(out[4].all())
        for key, j in (('x', 0), ('y', 1), ('cx', 2), ('cy', 3), ('opl', 6)):
            assert np.abs(r[key] - out[j].detach().numpy().ravel()).max() < 1e-10, key
        loss = sum((torch.tensor(s.reshape(1, 1, -1, 1)) * o).sum()
                   for s, o in zip(seeds, (out[0], out[1], out[2], out[3], out[6])))
        g = torch.autograd.grad(loss, [q['x'], q['y'], q['z'], q['cx'], q['cy'], tb['c'], tb['k'], ta,
                                       tb['t'], tb['mu']])
        got = (r['gx'], r['gy'], r['gz'], r['gcx'], r['gcy'], r['gp'][:, 0], r['gp'][:, 1], r['gp'][:, 2:],
               r['gt'], r['gmu'])
        for name, a_, b_ in zip(names, got, g):
            b_ = b_.numpy().reshape(np.shape(a_))
            err = np.abs(a_ - b_).max() / max(np.abs(b_).max(), 1e-3)
            assert err < 1e-8, (name, err)
