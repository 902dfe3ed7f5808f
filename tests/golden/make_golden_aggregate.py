"""Golden vectors of ``trace_skew(aggregate=True)`` (the penalty stacks, rtl:641-657) and of the
loss ``compute_loss_out`` builds on them (optics_simulator_lite.py:430-450), made by RUNNING the
unmodified reference on the inputs already stored in tests/golden/*.npz.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_aggregate.py

The reference's aggregate branch raises for the un-broadcast [1,1,P,1] pupil grids that
``trace_rays`` hands over when W > 1 (the boolean-mask assignment rtl:653 needs full shapes), so the
ray tensors are broadcast to [B,F,P,W] first -- exactly what ray aiming produces (rtl:129-208).
Stored per case (tests/golden/aggregate/<case>.npz): the three stacks as [S,B,F,P,W] arrays, the
penalty sum(Q), the RMS, and the gradients of ``penalty`` and of ``rms + 0.2 * penalty`` with
respect to the z, c, t, mu tensors of the call.
"""
import os

import numpy as np
import torch

from make_golden import HERE, import_reference

CASES = ('singlet_8x8', 'cooke_8x8', 'tessar_8x8', 'cooke_8x8_aimed', 'cooke_16x16_epd2.6',
         'cooke_16x16_epd2.6_nobackward', 'tessar_16x16_epd2.0')
PENALTY_RATE = 0.2      # optics_simulator_lite.py: penalty_rate default


def run_case(rtl, name):
    with np.load(os.path.join(HERE, name + '.npz')) as z:
        rec = {k: z[k] for k in z.files}
    shape = rec['out_ok'].shape
    ins = {k: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(rec['in_' + k], shape)))
           for k in ('x', 'y', 'cx', 'cy')}
    for k in ('z', 'c', 't', 'mu'):
        ins[k] = torch.from_numpy(rec['in_' + k]).clone().requires_grad_(True)
    mask = torch.from_numpy(rec['in_mask'])
    allow = bool(rec['allow_backward_rays'])
    out = rtl.trace_skew(ins['x'], ins['y'], ins['z'], ins['cx'], ins['cy'], ins['c'], ins['t'],
                         ins['mu'], mask, True, allow)
    x, y, cx, cy, ok, bw, stacks = out
    rms = rtl.compute_rms2d(x, y, ok)
    n_seq = int(mask.sum())                                  # len(sequence), osl:441
    q = (torch.stack(stacks['theta_norm'], dim=0).sum(dim=0) +
         torch.stack(stacks['theta_prime_norm'], dim=0).sum(dim=0) +
         torch.stack(stacks['z_RELU'], dim=0).sum(dim=0)) / n_seq
    q = torch.where(torch.isnan(q), torch.zeros_like(q), q)
    penalty = torch.sum(q)
    leaves = [ins[k] for k in ('z', 'c', 't', 'mu')]
    g_pen = torch.autograd.grad(penalty, leaves, retain_graph=True)
    g_loss = torch.autograd.grad(rms + PENALTY_RATE * penalty, leaves)
    res = dict(source=np.asarray(name), n_seq=np.asarray(n_seq), penalty=penalty.detach().numpy(),
               rms=rms.detach().numpy(), out_ok=ok.numpy(),
               out_y=y.detach().numpy(), out_x=x.detach().numpy())
    for key in ('z_RELU', 'theta_norm', 'theta_prime_norm'):
        res[key] = torch.stack([torch.broadcast_to(s, shape) for s in stacks[key]]).detach().numpy()
    for k, gp, gl in zip(('z', 'c', 't', 'mu'), g_pen, g_loss):
        res['gpen_' + k] = gp.numpy()
        res['gloss_' + k] = gl.numpy()
    return res


def main():
    rtl, _ = import_reference()
    os.makedirs(os.path.join(HERE, 'aggregate'), exist_ok=True)
    for name in CASES:
        res = run_case(rtl, name)
        np.savez_compressed(os.path.join(HERE, 'aggregate', name + '.npz'), **res)
        print(f"{name:32s} S={res['z_RELU'].shape[0]} rays={res['out_ok'].size:5d} ok={int(res['out_ok'].sum()):5d} "
              f"penalty={float(res['penalty']):.6f} rms={float(res['rms']):.8f} "
              f"|gpen_c|={np.abs(res['gpen_c']).max():.4f}")


if __name__ == '__main__':
    main()
