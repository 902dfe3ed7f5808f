"""The CPU oracle is pinned to the reference's own outputs (tests/golden/*.npz,
made by tests/golden/make_golden.py from the unmodified reference): masks and
all six trace outputs bit-for-bit, RMS and gradients to float round-off."""
import numpy as np
import torch

from oracle import trace_oracle as oracle


def _inputs(rec, grad=False):
    t = {k[3:]: torch.from_numpy(rec[k]) for k in rec if k.startswith('in_')}
    if grad:
        for k in ('z', 'c', 't', 'mu'):
            t[k] = t[k].clone().requires_grad_(True)
    return t


def test_trace_bit_exact(golden):
    i = _inputs(golden)
    x, y, cx, cy, ok, bw = oracle.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'],
                                        i['mu'], i['mask'], False,
                                        bool(golden['allow_backward_rays']))
    assert np.array_equal(ok.numpy(), golden['out_ok'])
    assert np.array_equal(bw.numpy(), golden['out_backward'])
    for got, want in ((x, 'out_x'), (y, 'out_y'), (cx, 'out_cx'), (cy, 'out_cy')):
        got = torch.broadcast_to(got, ok.shape).numpy()
        assert np.array_equal(got.view(np.uint32), golden[want].view(np.uint32)), want


def test_rms_and_gradients(golden):
    i = _inputs(golden, grad=True)
    out = oracle.trace(i['x'], i['y'], i['z'], i['cx'], i['cy'], i['c'], i['t'], i['mu'],
                       i['mask'], False, bool(golden['allow_backward_rays']))
    rms = oracle.spot_rms(out[0], out[1], out[4])
    assert abs(float(rms) - float(golden['rms'])) <= 1e-7 * max(1.0, abs(float(golden['rms'])))
    g = torch.autograd.grad(rms, [i['z'], i['c'], i['t'], i['mu']])
    for got, want in zip(g, ('grad_in_z', 'grad_in_c', 'grad_in_t', 'grad_in_mu')):
        ref = golden[want]
        err = np.linalg.norm(got.numpy() - ref) / max(np.linalg.norm(ref), 1e-30)
        assert err < 1e-6, (want, err)


def test_rms_all_lenses_matches_lens0(golden):
    y = torch.from_numpy(golden['out_y'])
    ok = torch.from_numpy(golden['out_ok'])
    got = oracle.spot_rms_all_lenses(y, ok)[0]
    assert abs(float(got) - float(golden['rms'])) <= 2e-6 * max(1.0, abs(float(golden['rms'])))
