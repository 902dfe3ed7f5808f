"""Multi-rank host logic on the CPU (gloo, world_size 2): the pupil axis is dealt
to ranks in contiguous slices that tile it exactly, and the per-(lens, field,
wavelength) sums of the slices add up through the one SUM all-reduce of the data
path.  The additive sums themselves are produced here by the oracle on each rank's
slice (the CUDA kernels that produce them on GPUs are covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchoptics_b200 import ops


def test_pupil_slices_tile_the_axis():
    for n in (1, 7, 64, 87616, 1000003):
        for world in (1, 2, 3, 8):
            edges = [ops.pupil_slice(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ops.pupil_slice(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import trace_oracle as oracle
    from torchoptics_b200 import RayTracer, prescriptions
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    specs, lens = prescriptions.load_yaml('baseline_cooke.yml', 'cpu')
    tracer = RayTracer(mode='circular', n_rays=(12, 10), rel_fields=(0., 0.707, 1.),
                       wavelengths=('C', 'd', 'F'), default_device='cpu')
    x, y, z, cx, cy, c, t, mu, mask = tracer._ray_set(specs, lens)
    n_pupil = x.shape[2]
    lo, hi = ops.pupil_slice(n_pupil, rank, world)
    out = oracle.trace(x[:, :, lo:hi], y[:, :, lo:hi], z, cx, cy, c, t, mu, mask)
    yy, ok = out[1].double(), out[4]
    # additive per-(lens, field, wavelength) sums of this slice: [sum y, sum y^2, n_ok]
    okf = ok.double()
    moments = torch.stack(((yy * okf).sum(2), (yy * yy * okf).sum(2), okf.sum(2)), dim=-1)
    # the penalty of compute_loss_out (rtl:641-657) is additive over pupil slices as well: one
    # more column rides in the same all-reduce
    ray_shape = (mu.shape[0], cy.shape[1], hi - lo, mu.shape[3])                      # [B,F,p,W]
    full = [torch.broadcast_to(a, ray_shape).contiguous() for a in (x[:, :, lo:hi], y[:, :, lo:hi], cx, cy)]
    stacks = oracle.trace(full[0], full[1], z, full[2], full[3], c, t, mu, mask, True)[6]
    terms = sum(torch.stack(stacks[k]).double().sum(0) for k in stacks)           # [B,F,p,W]
    moments = torch.cat((moments, terms.sum(2)[..., None]), dim=-1)
    ops.reduce_moments(moments)                        # the data path's one collective
    np.save(os.path.join(out_dir, f'pen_{rank}.npy'), moments[..., 3].sum().numpy())
    s1, s2, n_ok, _ = moments.sum(2).unbind(-1)        # -> per (lens, field)
    n = float(n_pupil * yy.shape[3])
    mean = s1 / n                                      # failed rays sit at y = 0
    rms = torch.sqrt((s2 - 2 * mean * s1 + n_ok * mean * mean) / n).mean(1)
    np.save(os.path.join(out_dir, f'rms_{rank}.npy'), rms.numpy())
    dist.destroy_process_group()


def test_two_rank_moment_allreduce_matches_single_process(tmp_path):
    from oracle import trace_oracle as oracle
    from torchoptics_b200 import RayTracer, prescriptions
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f'rms_{r}.npy') for r in range(world)]
    assert np.array_equal(got[0], got[1])              # every rank ends with the same loss
    specs, lens = prescriptions.load_yaml('baseline_cooke.yml', 'cpu')
    tracer = RayTracer(mode='circular', n_rays=(12, 10), rel_fields=(0., 0.707, 1.),
                       wavelengths=('C', 'd', 'F'), default_device='cpu')
    out = oracle.trace(*tracer._ray_set(specs, lens))
    want = oracle.spot_rms(out[0], out[1], out[4]).item()
    assert abs(got[0][0] - want) <= 1e-5 * want
    pens = [float(np.load(tmp_path / f'pen_{r}.npy')) for r in range(world)]
    assert pens[0] == pens[1]
    args = tracer._ray_set(specs, lens)
    shape = out[4].shape
    full = [torch.broadcast_to(a, shape).contiguous() if j in (0, 1, 3, 4) else a for j, a in enumerate(args)]
    stacks = oracle.trace(*full, True)[6]
    whole = float(sum(torch.stack(stacks[k]).double().sum() for k in stacks))
    assert abs(pens[0] - whole) <= 1e-9 * whole
