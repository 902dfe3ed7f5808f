"""csrc/paraxial.cuh (the per-lens functions behind tl_paraxial_fwd / tl_paraxial_bwd) compiled for the
host, against the torch statements of get_first_order rtl:772-794 and compute_last_curvature
rtl:725-769 in this package (themselves pinned to the reference, tests/test_host_logic.py) evaluated in
float64, values and autograd gradients."""
import numpy as np
import pytest
import torch

from tests.hostcore import binding
from torchoptics_b200 import ray_tracing_lite as rt

SEQUENCES = ['GAGAGA', 'GGAGA', 'GA', 'GAGAA', 'GAGGAAGGAGA', 'AGA', 'GAAGA']


def random_batch(seed, n_lens=12):
    rng = np.random.default_rng(seed)
    seqs = [SEQUENCES[i % len(SEQUENCES)] for i in range(n_lens)]
    L = max(len(s) for s in seqs)
    live = np.zeros((n_lens, L), bool)
    glass = np.zeros((n_lens, L), bool)
    for i, s in enumerate(seqs):
        live[i, :len(s)] = True
        glass[i, :len(s)] = [ch == 'G' for ch in s]
    c = np.where(live, rng.uniform(-0.05, 0.05, (n_lens, L)), 0.0)
    t = np.where(live, rng.uniform(0.5, 6.0, (n_lens, L)), 0.0)
    n = np.where(glass, rng.uniform(1.45, 1.85, (n_lens, L)), 1.0)
    return live, glass, c.astype(np.float32), t.astype(np.float32), n.astype(np.float32)


def torch_first_order(c, t, n, live):
    nd = torch.cat((torch.ones_like(n[:, 0:1]), n), dim=1)
    last = live.sum(dim=1) - 1
    t = t.clone()
    t[torch.arange(c.shape[0]), last] = 0.
    system = rt.reduce_abcd(rt.interface_propagation_abcd(c, t, nd))
    return -1 / system[:, 1, 0], -system[:, 0, 0] / system[:, 1, 0]


def torch_last_curvature(c, t, n, live, glass):
    """the padded core of compute_last_curvature (rtl:735-762)"""
    B = c.shape[0]
    rows = torch.arange(B)
    n_surf = live.sum(dim=1)
    air_air = ~glass[rows, n_surf - 2]
    solve_at = n_surf - 1 - air_air.long()
    ahead = live.clone()
    ahead[rows, n_surf - 1] = False
    ahead[rows, solve_at] = False
    n2d = torch.cat((torch.ones_like(n[:, 0:1]), n), dim=1)
    abcd = rt.interface_propagation_abcd(c, t, n2d)
    eye = torch.eye(2, dtype=c.dtype).expand_as(abcd)
    system = rt.reduce_abcd(torch.where(ahead[..., None, None], abcd, eye))
    n_after = n2d[rows, solve_at]
    return -(1 + n_after * system[:, 1, 0]) / (system[:, 0, 0] * (n_after - 1)), solve_at


@pytest.mark.parametrize('seed', range(4))
def test_first_order_values_and_adjoint(seed):
    live, glass, c, t, n = random_batch(seed)
    gout = np.random.default_rng(100 + seed).uniform(0.5, 1.5, (c.shape[0], 2)).astype(np.float32)
    out, gc, gt, gn = binding.paraxial(0, c, t, n, live, glass, gout)
    leaves = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (c, t, n)]
    efl, bfl = torch_first_order(*leaves, torch.from_numpy(live))
    assert np.allclose(out[:, 0], efl.detach().numpy(), rtol=2e-6)
    assert np.allclose(out[:, 1], bfl.detach().numpy(), rtol=2e-6)
    g = torch.tensor(gout, dtype=torch.float64)
    want = torch.autograd.grad((efl * g[:, 0] + bfl * g[:, 1]).sum(), leaves)
    # (padding slots are identity matrices and the index behind an air slot is the constant 1, not a variable:
    # c, t are compared on live slots, n on glass slots)
    for got, ref, name, where in zip((gc, gt, gn), want, 'ctn', (live, live, glass)):
        ref = np.where(where, ref.numpy(), 0.0)
        assert np.abs(np.where(where, got, 0.0) - ref).max() <= 2e-6 * np.abs(ref).max(), name


@pytest.mark.parametrize('seed', range(4))
def test_last_curvature_values_and_adjoint(seed):
    live, glass, c, t, n = random_batch(10 + seed)
    gout = np.random.default_rng(200 + seed).uniform(0.5, 1.5, (c.shape[0], 2)).astype(np.float32)
    out, gc, gt, gn = binding.paraxial(1, c, t, n, live, glass, gout)
    leaves = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (c, t, n)]
    solved, slot = torch_last_curvature(*leaves, torch.from_numpy(live), torch.from_numpy(glass))
    assert np.array_equal(out[:, 1].astype(np.int64), slot.numpy())
    assert np.allclose(out[:, 0], solved.detach().numpy(), rtol=2e-6)
    want = torch.autograd.grad((solved * torch.tensor(gout[:, 0], dtype=torch.float64)).sum(), leaves)
    for got, ref, name, where in zip((gc, gt, gn), want, 'ctn', (live, live, glass)):
        ref = np.where(where, ref.numpy(), 0.0)
        assert np.abs(np.where(where, got, 0.0) - ref).max() <= 2e-6 * np.abs(ref).max(), name


def test_solved_lens_has_unit_focal_length():
    live, glass, c, t, n = random_batch(7)
    out = binding.paraxial(1, c, t, n, live, glass)
    c2 = c.copy()
    rows = np.arange(c.shape[0])
    slot = out[:, 1].astype(np.int64)
    c2[rows, slot] = out[:, 0]
    c2[np.arange(c.shape[1])[None, :] > slot[:, None]] = 0.0
    efl = binding.paraxial(0, c2, t, n, live, glass)[:, 0]
    assert np.allclose(efl, 1.0, atol=2e-5)
