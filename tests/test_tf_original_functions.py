"""The functions that exist only in the reference's TensorFlow original (ray_tracing.py): pupil
samplers rt_tf:358-476, apply_vignetting rt_tf:479-490, compute_magnification rt_tf:765-777 and
compute_psf rt_tf:206-270, against golden vectors produced by executing the reference's own source
with a numpy stand-in for tensorflow (tests/golden/make_golden_tf.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_DIR
from torchoptics_b200 import lens_modeling as lm
from torchoptics_b200 import ray_tracing_lite as rt


@pytest.fixture(scope='module')
def tf_golden():
    with np.load(os.path.join(GOLDEN_DIR, 'tf', 'samplers_vignetting_magnification.npz')) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope='module')
def psf_golden():
    with np.load(os.path.join(GOLDEN_DIR, 'tf', 'psf.npz')) as z:
        return {k: z[k] for k in z.files}


def _same(got, want):
    got = got.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max(initial=0.0) <= 1e-7


@pytest.mark.parametrize('n_r,n_i', [(1, 1), (3, 2), (4, 4), (8, 8), (5, 3)])
@pytest.mark.parametrize('name', ['skew_uniform_half_equidistant', 'skew_uniform_half_jittered'])
def test_half_pupil_samplers(tf_golden, name, n_r, n_i):
    x, y = getattr(rt, name)(None, n_r, n_i, device='cpu')
    assert x.shape == (1, 1, n_r * n_r * n_i, 1)
    _same(x, tf_golden[f'{name}_{n_r}_{n_i}_x'])
    _same(y, tf_golden[f'{name}_{n_r}_{n_i}_y'])
    # and through the RayTracer mode switch (the reference's RaytracedOptics default is the jittered one)
    tracer = rt.RayTracer(mode=name, n_rays=(n_r, n_i), default_device='cpu')
    _same(tracer.pupil_span(None)[0], tf_golden[f'{name}_{n_r}_{n_i}_x'])


@pytest.mark.parametrize('n_y', [2, 5, 8])
def test_inner_square_sampler(tf_golden, n_y):
    x, y = rt.skew_inner_square_half(None, n_y, None, device='cpu')
    _same(x, tf_golden[f'skew_inner_square_half_{n_y}_x'])
    _same(y, tf_golden[f'skew_inner_square_half_{n_y}_y'])


@pytest.mark.parametrize('n', [2, 7, 16])
def test_line_and_rim_samplers(tf_golden, n):
    for name, fn in (('meridional_uniform', rt.meridional_uniform), ('sagittal_uniform', rt.sagittal_uniform),
                     ('circle_outer_edge_uniform', rt.circle_outer_edge_uniform)):
        x, y = fn(None, n, device='cpu')
        assert np.abs(x.numpy() - tf_golden[f'{name}_{n}_x']).max() <= 2e-7, name
        assert np.abs(y.numpy() - tf_golden[f'{name}_{n}_y']).max() <= 2e-7, name
    for name, fn in (('chief', lambda: rt.chief(None, None, device='cpu')), ('tee', lambda: rt.tee(None, 'cpu'))):
        x, y = fn()
        _same(x, tf_golden[name + '_x'])
        _same(y, tf_golden[name + '_y'])


def test_apply_vignetting(tf_golden):
    got = rt.apply_vignetting(torch.from_numpy(tf_golden['vig_in_y']), torch.from_numpy(tf_golden['vig_up']),
                              torch.from_numpy(tf_golden['vig_down']))
    assert np.abs(got.numpy() - tf_golden['vig_out']).max() <= 1e-7


@pytest.mark.parametrize('name', ['baseline_cooke', 'baseline_tessar', 'baseline_doublet'])
def test_compute_magnification_is_the_a_element(tf_golden, name):
    c, t, nd = (torch.from_numpy(tf_golden[f'magnification_{name}_{k}']) for k in ('c', 't', 'nd'))
    n = c.shape[1]
    structure = lm.Structure(np.array([n]), mask=np.ones((1, n), bool), mask_G=(nd.numpy() != 1.0),
                             default_device='cpu')
    lens = lm.Lens(structure, c, t, nd, torch.full_like(nd, 50.0))
    got = rt.compute_magnification(lens)
    assert np.abs(got.numpy() - tf_golden[f'magnification_{name}']).max() <= 2e-6


def _psf_cases(psf_golden):
    return sorted(k[:-len('_kernels')] for k in psf_golden if k.endswith('_kernels'))


def test_psf_oracle_matches_the_reference_source(psf_golden):
    """oracle/psf_oracle.py (numpy restatement) == the reference's compute_psf executed from its own file."""
    from oracle import psf_oracle
    cases = _psf_cases(psf_golden)
    assert len(cases) >= 7
    for tag in cases:
        incr = float(psf_golden[tag + '_increment'])
        target = psf_golden.get(tag + '_in_y_target')
        x_size, y_size, y_target, kernels, accounted = psf_oracle.compute_psf(
            psf_golden[tag + '_in_x'], psf_golden[tag + '_in_y'], tuple(int(v) for v in psf_golden[tag + '_n_bins']),
            None if np.isnan(incr) else incr, target)
        want = psf_golden[tag + '_kernels']
        assert kernels.shape == want.shape, tag
        both_nan = np.isnan(kernels) & np.isnan(want)      # a grid no ray hits: 0 / 0 in the reference, too
        assert np.array_equal(np.isnan(kernels), np.isnan(want)), tag
        assert np.abs(np.where(both_nan, 0, kernels - want)).max() <= 1e-7, tag
        assert np.allclose(y_target, psf_golden[tag + '_y_target'], atol=1e-7), tag
        assert np.allclose(np.asarray(x_size, np.float64), psf_golden[tag + '_x_size'], rtol=1e-6), tag
        assert np.allclose(np.asarray(y_size, np.float64), psf_golden[tag + '_y_size'], rtol=1e-6), tag
        assert np.array_equal(accounted, psf_golden[tag + '_accounted']), tag


def test_psf_oracle_properties():
    """Reference-independent checks of the soft histogram: unit mass, mirror symmetry in x, centroid at
    the target, and the sigma -> 0 limit is a hard histogram."""
    from oracle import psf_oracle
    rng = np.random.default_rng(0)
    x = np.abs(rng.normal(0, 0.01, (1, 2, 3, 4000))).astype(np.float32)
    y = (rng.normal(0, 0.012, (1, 2, 3, 4000)) + np.array([1.0, 2.0])[None, :, None, None]).astype(np.float32)
    x_size, y_size, y_target, kernels, accounted = psf_oracle.compute_psf(x, y, (21, 21), increment=0.004)
    assert np.allclose(kernels.sum(axis=(-1, -2)), 1.0, atol=1e-5)
    assert np.allclose(kernels, kernels[..., ::-1], atol=1e-7)
    rows = (np.arange(21) + 0.5 - 10.5) * 0.004
    centroid = (kernels.sum(axis=-1) * rows).sum(axis=-1)
    assert np.abs(centroid).max() < 3e-4                   # centred on the mean of y
    assert np.allclose(y_target, y.reshape(2, -1).mean(axis=1), atol=1e-6)
