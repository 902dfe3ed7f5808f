"""compute_psf on the device (tl_psf_bin) against golden vectors produced by
executing the reference's own source (ray_tracing.py:206-270, the TensorFlow original) under a numpy
stand-in for tensorflow (tests/golden/make_golden_tf.py), plus size-independent properties at a
spot-sweep size."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_DIR
from torchoptics_b200 import ray_tracing_lite as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def psf_golden():
    with np.load(os.path.join(GOLDEN_DIR, 'tf', 'psf.npz')) as z:
        return {k: z[k] for k in z.files}


def test_psf_matches_the_reference_source(psf_golden):
    cases = sorted(k[:-len('_kernels')] for k in psf_golden if k.endswith('_kernels'))
    assert len(cases) >= 7
    for tag in cases:
        incr = float(psf_golden[tag + '_increment'])
        target = psf_golden.get(tag + '_in_y_target')
        x = torch.from_numpy(psf_golden[tag + '_in_x']).to(DEV)
        y = torch.from_numpy(psf_golden[tag + '_in_y']).to(DEV)
        n_bins = tuple(int(v) for v in psf_golden[tag + '_n_bins'])
        x_size, y_size, y_target, kernels, accounted = rt.compute_psf(
            x, y, n_bins, None if np.isnan(incr) else incr, None if target is None else torch.from_numpy(target).to(DEV))
        if target is None and not np.array_equal(y_target.cpu().numpy(), psf_golden[tag + '_y_target']):
            # the default centre is an fp32 mean of ~1e4 heights of ~5 mm: its last bit depends on the summation
            # order (numpy pairwise here, a CUDA tree there, TF's own in the original), and ONE ulp of the centre
            # (5e-7 mm against sigma = 1.5e-3 mm) moves the kernels by 3e-5 -- so the bins are compared at the
            # golden record's centre, and the default centre itself to the ulp below
            assert np.abs(y_target.cpu().numpy() - psf_golden[tag + '_y_target']).max() <= 1e-6, tag
            _, _, _, kernels, accounted = rt.compute_psf(x, y, n_bins, None if np.isnan(incr) else incr,
                                                         torch.from_numpy(psf_golden[tag + '_y_target']).to(DEV))
        want = psf_golden[tag + '_kernels']
        got = kernels.cpu().numpy()
        assert got.shape == want.shape and got.dtype == np.float32, tag
        assert np.array_equal(np.isnan(got), np.isnan(want)), tag        # (a grid no ray hits: 0 / 0, as in the reference)
        err = np.abs(np.where(np.isnan(want), 0, got - want)).max()
        assert err <= 2e-6, (tag, err)                                   # unit-mass kernels, peak ~1e-2 .. 1e-1
        assert np.allclose(y_target.cpu().numpy(), psf_golden[tag + '_y_target'], atol=2e-6), tag
        assert np.allclose(np.asarray(torch.as_tensor(x_size).cpu(), np.float64), psf_golden[tag + '_x_size'], rtol=1e-5), tag
        assert np.allclose(np.asarray(torch.as_tensor(y_size).cpu(), np.float64), psf_golden[tag + '_y_size'], rtol=1e-5), tag
        assert np.abs(accounted.cpu().numpy() - psf_golden[tag + '_accounted']).max() <= 2e-3, tag   # (rays on a window edge)


def test_psf_properties_at_sweep_size():
    """2 M rays per (field, channel): unit mass, mirror symmetry in x, centroid on the target, linearity
    in the ray set (two halves of the rays add up), and agreement with the oracle on a subsample."""
    from oracle import psf_oracle
    gen = torch.Generator(device='cpu').manual_seed(0)
    n = 1 << 21
    x = (torch.randn((1, 2, 3, n), generator=gen) * 0.006).abs()
    y = torch.randn((1, 2, 3, n), generator=gen) * 0.008 + torch.tensor([3.0, 5.0]).reshape(1, 2, 1, 1)
    xd, yd = x.to(DEV), y.to(DEV)
    bins, incr = (21, 21), 0.002
    target = y.reshape(2, -1).mean(dim=1)
    _, _, yt, k, acc = rt.compute_psf(xd, yd, bins, incr, target.to(DEV))
    k64 = k.double().cpu().numpy()
    assert np.allclose(k64.sum(axis=(-1, -2)), 1.0, atol=1e-5)
    assert np.abs(k64 - k64[..., ::-1]).max() <= 1e-7
    rows = (np.arange(21) + 0.5 - 10.5) * incr
    assert np.abs((k64.sum(axis=-1) * rows).sum(axis=-1)).max() < 2e-5
    # linearity: un-normalised sums of the two halves of the ray set add up to the whole
    from torchoptics_b200 import ops
    def sums(xs, ys):
        g = torch.full((2,), incr, device=DEV)
        w = torch.full((2,), incr * 21, device=DEV)
        return ops.psf_bin(xs.reshape(2, 3, -1), ys.reshape(2, 3, -1), target.to(DEV), g, g, w, w, bins)
    whole, inside = sums(xd, yd)
    a, ia = sums(xd[..., :n // 2].contiguous(), yd[..., :n // 2].contiguous())
    b, ib = sums(xd[..., n // 2:].contiguous(), yd[..., n // 2:].contiguous())
    assert float(((a + b - whole).abs() / whole.abs().clamp_min(1e-30)).max()) <= 1e-5
    assert torch.equal(ia + ib, inside)
    # the oracle on the first 4096 rays of every (field, channel)
    sub = 4096
    _, _, _, k_ref, acc_ref = psf_oracle.compute_psf(x[..., :sub].numpy(), y[..., :sub].numpy(), bins, incr, target.numpy())
    _, _, _, k_sub, acc_sub = rt.compute_psf(xd[..., :sub].contiguous(), yd[..., :sub].contiguous(), bins, incr, target.to(DEV))
    assert np.abs(k_sub.cpu().numpy() - k_ref).max() <= 2e-6
    assert np.abs(acc_sub.cpu().numpy() - acc_ref).max() <= 1e-3
