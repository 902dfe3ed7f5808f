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


def test_fast_policy_fp32_with_early_exit_on_the_config3_lens(tmp_path):
    """The fp32 fast policy as the kernels run it -- Newton loop left once a step is <= 1e-3 |tau|, hit-point
    evaluation skipped after a step <= 2e-6 |tau| (csrc/trace_core_asph.cuh: newton_settled, newton_is_fresh) --
    on the config-3 lens (eight aspheric surfaces, four flat ones), one wavelength at a time: within the north-star
    1e-5 of the fp64 oracle (four fixed steps), and within 1e-6 of scale (a few ulp of the image height; a tenth of the budget) of THE SAME header compiled with the exit
    disabled (-DTL_NEWTON_EXIT=0: four steps and the evaluation, always) -- what the exit itself changes.  The
    distances are printed next to the fp32 oracle's own."""
    import ctypes
    import os
    import subprocess
    from torchoptics_b200 import prescriptions, ray_tracing_lite as rt
    here = os.path.dirname(os.path.abspath(hc.__file__))
    fixed_so = str(tmp_path / 'hostcore_fixed4.so')
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-std=c++17', '-shared', '-fPIC',
                           '-DTL_NEWTON_EXIT=0.0f', '-x', 'c++', os.path.join(here, 'hostcore.cpp'), '-o', fixed_so])
    specs, lens = prescriptions.asphere_12('cpu')
    tracer = rt.RayTracer(mode='circular', n_rays=(24, 24), rel_fields=(0.0, 0.35, 0.7, 1.0),
                          wavelengths=('C', 'd', 'F'), default_device='cpu')
    args = [a_.detach() for a_ in tracer._ray_set(specs, lens)]
    ext = {k: v.detach() for k, v in tracer._extension_tables(lens).items() if v is not None}
    ref32 = gen.trace(*args, **ext)
    ref64 = gen.trace(*[a_.double() if a_.is_floating_point() else a_ for a_ in args],
                      **{k: (v.double() if v.is_floating_point() else v) for k, v in ext.items()})
    assert bool(ref64[4].all()) and bool(ref32[4].all())
    x, y, z, cx, cy, c, t, mu, mask = args
    F, W, S, P = cy.shape[1], mu.shape[3], c.shape[-1], x.shape[2]
    xy_scale = max(float(ref64[0].abs().max()), float(ref64[1].abs().max()))

    def run(w):
        rays = [torch.broadcast_to(v, (1, F, P, 1)).reshape(-1).numpy().astype(np.float32) for v in (x, y, z, cx, cy)]
        return hc.asph_fast(np.float32, *rays, c.reshape(S).numpy(), ext['k'].reshape(S).numpy(),
                            ext['a'].reshape(S, 7).numpy(), t.reshape(S).numpy(), mu[0, 0, 0, w].numpy(),
                            np.full(S, np.inf))

    hc.lib()
    shipped = hc._lib
    worst = {}
    for w in range(W):
        r = run(w)
        try:
            hc._lib = ctypes.CDLL(fixed_so)
            r4 = run(w)
        finally:
            hc._lib = shipped
        assert float(r['min_cos2'].min()) > 1e-3          # every ray clear of the guard bands: no exact re-trace
        for key, j in (('x', 0), ('y', 1), ('cx', 2), ('cy', 3), ('opl', 6)):
            want64 = ref64[j][0, :, :, w].numpy().ravel()
            want32 = ref32[j][0, :, :, w].numpy().ravel().astype(np.float64)
            scale = 1.0 if key in ('cx', 'cy') else (xy_scale if key in ('x', 'y') else float(np.abs(want64).max()))
            err_fast = float(np.abs(r[key].astype(np.float64) - want64).max()) / scale
            err_ref = float(np.abs(want32 - want64).max()) / scale
            d_exit = float(np.abs(r[key].astype(np.float64) - r4[key].astype(np.float64)).max()) / scale
            worst[key] = tuple(max(p_, q_) for p_, q_ in zip(worst.get(key, (0.0, 0.0, 0.0)), (err_fast, err_ref, d_exit)))
            assert err_fast <= 1e-5, (key, w, err_fast)
            assert d_exit <= 1e-6, (key, w, d_exit)
    print({k: 'fast-fp64 %.1e, fp32 oracle-fp64 %.1e, exit on/off %.1e' % v for k, v in worst.items()})
