"""Host-side mirror of the reference API (ray-set construction, paraxial optics,
data model) against the golden records of the reference -- CPU only."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_DIR
from torchoptics_b200 import lens_modeling as lm
from torchoptics_b200 import prescriptions
from torchoptics_b200 import ray_tracing_lite as rt


def _lens_from(rec):
    structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device='cpu')
    lens = lm.Lens(structure, torch.from_numpy(rec['lens_c']), torch.from_numpy(rec['lens_t']),
                   torch.from_numpy(rec['lens_nd']), torch.from_numpy(rec['lens_v']))
    specs = lm.Specs(structure, torch.from_numpy(rec['epd']), torch.from_numpy(rec['hfov']))
    return specs, lens


def test_first_order_matches_reference(golden):
    _, lens = _lens_from(golden)
    efl, bfl = rt.get_first_order(lens)
    assert np.allclose(efl.numpy(), golden['efl'], rtol=2e-6)
    assert np.allclose(bfl.numpy(), golden['bfl'], rtol=2e-6)
    assert np.allclose(lens.efl.numpy(), golden['efl'], rtol=2e-6)


def test_ray_set_matches_reference(golden):
    if 'aimed' in golden['name']:
        pytest.skip('ray aiming traces rays: covered by the gpu tests')
    from tests.conftest import golden_problem
    tracer, specs, lens = golden_problem(golden, 'cpu', requires_grad=False)      # (incl. pupil vignetting)
    x, y, z, cx, cy, c, t, mu, mask = tracer._ray_set(specs, lens)
    for got, key in ((x, 'in_x'), (y, 'in_y'), (cx, 'in_cx'), (cy, 'in_cy'), (c, 'in_c'), (t, 'in_t')):
        assert got.shape == golden[key].shape, key
        assert np.array_equal(got.numpy(), golden[key]), key
    assert np.array_equal(mask.numpy(), golden['in_mask'])
    assert np.allclose(z.numpy(), golden['in_z'], rtol=1e-6, atol=1e-7)
    assert np.allclose(mu.numpy(), golden['in_mu'], rtol=1e-6)


def test_yaml_loader_and_shipped_lenses():
    for fn, efl in (('singlet_lens.yml', 17.156055), ('baseline_doublet.yml', 17.156054),
                    ('baseline_cooke.yml', 17.156055), ('baseline_tessar.yml', 17.154451)):
        specs, lens = prescriptions.load_yaml(fn, 'cpu')
        assert abs(float(lens.efl[0]) - efl) < 1e-4
        assert specs.epd.shape == (1,)
    specs, lens = prescriptions.double_gauss('cpu')
    assert lens.c.shape == (1, 11)
    assert abs(float(lens.efl[0]) - 49.75) < 0.05
    assert abs(float(lens.bfl[0]) - 28.749) < 0.05


def test_compute_last_curvature_normalises_efl():
    specs, lens = prescriptions.load_yaml('baseline_cooke.yml', 'cpu')
    s = lens.structure
    c = rt.compute_last_curvature(s, lens.flat_c_but_last, lens.flat_t, lens.flat_nd)
    assert abs(float(c[-1]) - (-1.6177329)) < 2e-5
    solved = lm.Lens(s, c, lens.flat_t, lens.flat_nd, lens.flat_v)
    assert abs(float(solved.efl[0]) - 1.0) < 1e-5


def test_structure_slicing_and_up_to_stop():
    s = lm.Structure(np.array([4, 2]), sequence=np.array(['GAGAAGA', 'GAGA']), default_device='cpu')
    assert s.mask.shape == (2, 7) and s.mask[1].sum() == 4
    front = s.up_to_stop()
    assert front.mask.shape == (2, 4)
    assert front.mask[0].all() and front.mask[1].tolist() == [True, True, False, False]
    assert s[1].mask.shape == (1, 4)
    assert s.last_g_idx.tolist() == [5, 2]


def test_glass_round_trip():
    n = torch.tensor([1.5168, 1.7552])
    v = torch.tensor([64.17, 27.51])
    n2, v2 = lm.n_v_from_g(lm.g_from_n_v(n, v))
    assert torch.allclose(n, n2, atol=1e-5) and torch.allclose(v, v2, atol=1e-3)


def test_pupil_samplers_shapes():
    x, y = rt.circle(None, 4, 6, 'cpu')
    assert x.shape == (1, 1, 24, 1) and float(x[0, 0, 0, 0]) == 0.0
    x, y = rt.tee(None, 'cpu')
    assert y.flatten().tolist() == [-1., 1., 0.]
    x, y = rt.circle_pseudo_random(torch.zeros(2, 1, 1, 1), 3, 5, device='cpu')
    assert x.shape == (2, 1, 15, 1) and float((x ** 2 + y ** 2).max()) <= 1.0 + 1e-6
    x, y = rt.circle_outer_edge_uniform(None, 8, 'cpu')
    assert torch.allclose(x ** 2 + y ** 2, torch.ones_like(x), atol=1e-6)


def test_cpu_tensors_are_rejected_loudly():
    from torchoptics_b200 import _native
    specs, lens = prescriptions.load_yaml('baseline_cooke.yml', 'cpu')
    tracer = rt.RayTracer(mode='circular', n_rays=(4, 4), default_device='cpu')
    with pytest.raises(_native.NativeLibraryError):
        tracer.trace_rays(specs, lens)


def test_staging_decision_of_the_front_end():
    """Which ray sets the staged kernels (tl_stage_fwd, tl_aim) build, decided on the host:
    (can_stage, aimed) for the tracer settings the reference offers."""
    from tests.conftest import load_golden
    specs, lens = _lens_from(load_golden('cooke_8x8'))          # stop inside the lens (stop_idx 4)
    base = dict(mode='circular', n_rays=(8, 8), rel_fields=(0., 1.), wavelengths=('d',), default_device='cpu')
    assert rt.RayTracer(**base)._staging(lens) == (True, False)
    assert rt.RayTracer(n_ray_aiming_iter=1, **base)._staging(lens) == (True, True)
    mirror = rt.RayTracer(n_ray_aiming_iter=1, **base)
    mirror.device_aiming = False
    assert mirror._staging(lens) == (False, False)
    # round 2: the 'paraxial' stop radius and pupil vignetting functions stay on the staged path too
    assert rt.RayTracer(n_ray_aiming_iter=1, ray_aiming_mode='paraxial', **base)._staging(lens) == (True, True)
    vig = rt.RayTracer(vig_fn=lambda fields, v: v[:, None] * fields, **base)
    assert vig._staging(lens)[0] is True and vig._staging(lens, use_vig=False)[0] is True
    kw = vig._staged_kwargs(specs, lens)
    assert kw['vig'].shape == (1, 2, 3) and vig._staged_kwargs(specs, lens, use_vig=False)['vig'] is None
    rnd = dict(base, mode='skew_random')
    assert rt.RayTracer(**rnd)._staging(lens)[0] is False
    # a stop in front of the lens needs no aiming at all (rtl:131-133)
    structure = lm.Structure(np.array([0]), sequence=np.array(['GA']), default_device='cpu')
    front = lm.Lens(structure, torch.tensor([[0.05, -0.05]]), torch.tensor([[2.0, 10.0]]),
                    torch.tensor([[1.5, 1.0]]), torch.tensor([[60.0, 0.0]]))
    assert rt.RayTracer(n_ray_aiming_iter=1, **base)._staging(front) == (True, False)


def test_stack_keys_and_exchange_module_import_without_a_gpu():
    from torchoptics_b200 import ops, peer
    assert ops.STACK_KEYS == ('z_RELU', 'theta_norm', 'theta_prime_norm')       # rtl:598
    assert hasattr(peer, 'PeerExchange')
    with pytest.raises(Exception):       # CPU tensors never reach a kernel
        ops.penalty_sum(*[torch.zeros(1, 1, 1, 1)] * 5, *[torch.zeros(1, 1, 1, 1, 2)] * 3,
                        torch.ones(1, 1, 1, 1, 2, dtype=torch.bool), 2)


def test_lens_tables_follow_the_structure_not_its_id():
    """Regression: the staged pass' device tables were cached under
    id(structure); CPython reuses ids, so a loop over freshly built structures of one shape got
    stale masks / stop indices.  They now live on the Structure, stamped with its content."""
    import gc
    from torchoptics_b200 import RayTracer
    from torchoptics_b200.lens_modeling import Lens, Structure
    tracer = RayTracer(mode='circular', n_rays=(4, 4), default_device='cpu')
    seqs = (np.array(['GAGAAGA']), np.array(['GAAGAGA']))
    stops = (np.array([4]), np.array([2]))
    for i in range(50):
        structure = Structure(stops[i % 2], sequence=seqs[i % 2], default_device='cpu')
        z = torch.zeros((1, 7))
        lens = Lens(structure, z, z, z + 1.5, z + 50.0)
        tab = tracer._tables(lens)
        assert np.array_equal(tab.mask_g.numpy().astype(bool), structure.mask_G), i
        assert int(tab.stop_idx[0]) == int(structure.stop_idx[0]), i
        assert tracer._tables(lens) is tab                       # cached while the structure lives
        del structure, lens, tab
        gc.collect()
    # an in-place edit of the structure invalidates both caches
    structure = Structure(np.array([4]), sequence=seqs[0], default_device='cpu')
    lens = Lens(structure, z, z, z + 1.5, z + 50.0)
    first, front = tracer._tables(lens), structure.up_to_stop()
    structure.stop_idx[0] = 2
    assert int(tracer._tables(lens).stop_idx[0]) == 2 and tracer._tables(lens) is not first
    assert structure.up_to_stop() is not front and structure.up_to_stop().mask.shape[1] == 2


def test_optical_loss_supervised_and_codes_match_the_reference():
    """optical_loss.py:136-176 (plain tensor arithmetic, runs anywhere) against the records of the reference's own
    class, and the inferred sequence codes (optical_loss.py:14-18)."""
    import glob
    from torchoptics_b200.optical_loss import Optical_Loss, sequence_decoder, sequence_encoder
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, 'optical_loss', '*.npz')))
    assert len(paths) == 4
    for path in paths:
        with np.load(path) as z:
            lens_type = str(z['lens_type'])
            loss = Optical_Loss(lens_type)
            got = loss.optical_loss_supervised(torch.from_numpy(z['supervised_labels']), torch.from_numpy(z['outputs']), device='cpu')
            assert abs(float(got) - float(z['supervised_loss'])) <= 1e-6 * abs(float(z['supervised_loss']))
            assert sequence_decoder(sequence_encoder(lens_type)) == lens_type
            assert z['outputs'].shape[1] == loss.numout and z['inputs'].shape[1] == loss.numin + 4
    t = torch.arange(4.0)
    assert torch.equal(Optical_Loss.t_converter(2, 'GAGA', t, torch.tensor([9.0])), torch.tensor([0., 9., 1., 2., 3.]))
    assert Optical_Loss.t_converter(2, 'GAGA', t, -1) is t and Optical_Loss.t_converter(1, 'GAGA', t, torch.tensor([9.0])) is t
