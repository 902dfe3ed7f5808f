"""tl_paraxial_fwd / tl_paraxial_bwd (get_first_order rtl:772-794 and
compute_last_curvature rtl:725-769 as one kernel each way) against the torch statements of the same
functions evaluated on the CPU -- which tests/test_host_logic.py pins to the reference -- and the batched
training-loss entry `Optical_Loss` against golden vectors produced by the reference's own optical_loss.py /
optics_simulator_lite.py run sample by sample (tests/golden/make_golden_optical_loss.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_DIR
from tests.test_paraxial_cpu import random_batch
from torchoptics_b200 import ops, prescriptions
from torchoptics_b200 import ray_tracing_lite as rt
from torchoptics_b200.lens_modeling import Lens, Structure
from torchoptics_b200.optical_loss import Optical_Loss, sequence_decoder, sequence_encoder

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name', ['baseline_cooke.yml', 'baseline_tessar.yml', 'baseline_doublet.yml'])
def test_first_order_kernel_on_the_shipped_lenses(name):
    _, lens_gpu = prescriptions.load_yaml(name, DEV)
    _, lens_cpu = prescriptions.load_yaml(name, 'cpu')
    efl, bfl = rt.get_first_order(lens_gpu)
    efl_ref, bfl_ref = rt.get_first_order(lens_cpu)
    assert torch.allclose(efl.cpu(), efl_ref, rtol=2e-6) and torch.allclose(bfl.cpu(), bfl_ref, rtol=2e-6)


@pytest.mark.parametrize('seed', range(3))
def test_kernels_match_torch_on_random_batches(seed):
    live, glass, c, t, n = random_batch(seed, n_lens=40)
    structure = Structure(np.zeros(c.shape[0], np.int64), live, glass, default_device=DEV)
    for mode in (0, 1):
        leaves = [torch.tensor(v, device=DEV, requires_grad=True) for v in (c, t, n)]
        if mode == 0:
            a, b = ops.first_order(structure, *leaves)
            value = a * 1.25 + b * 0.75
        else:
            value, slot = ops.last_curvature(structure, *leaves)
        got = torch.autograd.grad(value.sum(), leaves)
        from tests.hostcore import binding
        gout = np.tile(np.asarray([[1.25, 0.75]], np.float32), (c.shape[0], 1)) if mode == 0 else \
            np.ones((c.shape[0], 2), np.float32)
        out, gc, gt, gn = binding.paraxial(mode, c, t, n, live, glass, gout)      # (the same source on the host: fp64 inside)
        want_value = out[:, 0] * 1.25 + out[:, 1] * 0.75 if mode == 0 else out[:, 0]
        assert np.allclose(value.detach().cpu().numpy(), want_value, rtol=1e-6)
        if mode == 1:
            assert np.array_equal(slot.cpu().numpy(), out[:, 1].astype(np.int64))
        for g, ref in zip(got, (gc, gt, gn)):
            assert np.abs(g.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_compute_last_curvature_cuda_equals_cpu_path():
    sequence = np.array(['GAGAGA', 'GGAGA', 'GAGAA'])
    c = torch.tensor([0.9, -0.8, 0.5, -0.4, 0.7, 1.1, 0.2, -0.6, -0.9, 0.8, -0.3, 0.4, 0.1])      # all but each last
    t = torch.tensor([0.1, 0.05, 0.1, 0.04, 0.08, 0.5, 0.07, 0.1, 0.03, 0.06, 0.6, 0.09, 0.05, 0.1, 0.04, 0.5])
    nd = torch.tensor([1.6, 1.7, 1.5, 1.55, 1.65, 1.75, 1.62, 1.58])
    out = {}
    for dev in ('cpu', DEV):
        structure = Structure(np.zeros(3, np.int64), sequence=sequence, default_device=dev)
        leaves = [v.to(dev).requires_grad_(True) for v in (c, t, nd)]
        full = rt.compute_last_curvature(structure, *leaves)
        grads = torch.autograd.grad((full * torch.arange(1, full.numel() + 1, device=dev)).sum(), leaves)
        out[dev] = [full.detach().cpu()] + [g.cpu() for g in grads]
    for a, b in zip(out['cpu'], out[DEV]):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (a, b)


def test_sequence_codes_round_trip():
    for seq in ('GA', 'GGA', 'GAGA', 'GAGAGA'):
        assert sequence_decoder(sequence_encoder(seq)) == seq
    loss = Optical_Loss('GAGA')
    assert (loss.numsurf, loss.numglass, loss.numin, loss.numout) == (4, 2, 10, 11)      # ol:14-18


GOLDENS = sorted(glob.glob(os.path.join(GOLDEN_DIR, 'optical_loss', '*.npz')))


@pytest.mark.parametrize('path', GOLDENS, ids=[os.path.basename(p)[:-4] for p in GOLDENS])
def test_batched_optical_loss_matches_the_reference_loop(path):
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    lens_type = str(g['lens_type'])
    loss_fn = Optical_Loss(lens_type)
    x = torch.from_numpy(g['inputs']).to(DEV)
    y = torch.from_numpy(g['outputs']).to(DEV).requires_grad_(True)
    loss, rms, penalty = loss_fn.per_sample(x, y, float(g['penalty_rate']), device=DEV)
    # the aimed pupil differs from the reference's by ~1e-5 of its radius (tests/test_gpu_parity.py), hence 5e-5 on
    # the RMS, like the aimed golden cases; the penalty is a sum over 1 536 rays of O(1) terms
    assert np.allclose(rms.detach().cpu().numpy(), g['per_sample_rms'], rtol=5e-5), (rms, g['per_sample_rms'])
    assert np.allclose(penalty.detach().cpu().numpy(), g['per_sample_penalty'], rtol=2e-5)
    assert np.allclose(loss.detach().cpu().numpy(), g['per_sample_loss'], rtol=2e-5)
    grads, = torch.autograd.grad(loss.sum(), y)
    # The loss is dominated by the penalty (0.2 x ~250 against an RMS of ~2e-3), whose angle terms are
    # ill-conditioned in fp32 near normal incidence: the reference's own fp32 gradient is up
    # to 4e-3 from the same code run in float64 (printed by the generator; a single cancelled component, GA sample 3's
    # d/dc, is 11 % off).  Measured here: ours sits 10-1000 x CLOSER to the float64 run than the reference's fp32 on
    # 90 % of the (sample, group) pairs and within 1.5-3 x of it on the rest.  The bar, per sample and parameter group:
    # within 1e-4 of the float64 run, or no farther from it than 3 x the reference's own fp32 gradient -- and over the
    # samples of a file the median distance must be below the reference's.
    got, ref32, ref64 = grads.cpu().numpy().astype(np.float64), g['grad_outputs'].astype(np.float64), g['f64_grad_outputs']
    G, S = loss_fn.numglass, loss_fn.numsurf
    for name, cols in (('g', slice(0, 2 * G)), ('c', slice(2 * G, 2 * G + S - 1)), ('t', slice(2 * G + S - 1, None))):
        scale = np.abs(ref64[:, cols]).max(axis=1)
        ours = np.abs(got[:, cols] - ref64[:, cols]).max(axis=1) / scale
        theirs = np.abs(ref32[:, cols] - ref64[:, cols]).max(axis=1) / scale
        print(lens_type, 'd loss / d', name, '\n   ours-vs-fp64          ', ours, '\n   reference-fp32-vs-fp64', theirs)
        assert ((ours <= 1e-4) | (ours <= 3 * theirs)).all(), (name, ours, theirs)
        assert np.median(ours) <= np.median(theirs), (name, ours, theirs)
    mean = loss_fn.optical_loss_unsupervised(x, y.detach(), float(g['penalty_rate']), device=DEV)
    assert np.allclose([float(v) for v in mean], g['batch_mean'], rtol=2e-5)
    one = loss_fn.optical_loss_unsupervised_single(x[2], y.detach()[2], float(g['penalty_rate']), device=DEV)
    assert np.allclose([float(v) for v in one], [g['per_sample_loss'][2], g['per_sample_rms'][2], g['per_sample_penalty'][2]],
                       rtol=5e-5)
