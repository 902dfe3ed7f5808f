"""End-to-end lens optimisation on the fused spot pass (BASELINE.json config 5): Adam on the
curvatures / conic constants / asphere coefficients / thicknesses of a lens, every step one
fused trace -> RMS -> gradient pass over the whole ray set.

Data parallel over GPUs: every rank traces its slice of the pupil (``shard=(rank, world)``), the
per-field sums are all-reduced inside the pass, so loss and gradients -- hence the Adam state and
the lens -- are identical on every rank without any further communication.
"""
from __future__ import annotations

import torch

from .lens_modeling import Lens

# a_i multiplies rho^i (rho in mm^2): step the coefficients in units of 1 / height^(2 i)
_ASPHERE_UNIT_HEIGHT = 5.0


def optimize_spot(tracer, specs, lens, variables=('c', 't', 'k', 'a'), steps=100, lr=1e-3, shard=(0, 1),
                  group=None, keep_last_thickness=True, callback=None):
    """Minimise the mean RMS spot size of ``lens`` with Adam.  Returns (optimised Lens, loss history).

    ``variables`` names the optimised Lens fields; ``a`` is stepped in dimensionless units
    (a_i * h^(2 i), h = 5 mm) so that one learning rate fits all orders.  The last thickness
    (image distance) is part of ``t`` unless ``keep_last_thickness``."""
    dev = lens.c.device
    fields = {name: getattr(lens, name) for name in ('c', 't', 'nd', 'v', 'k', 'a', 'sd')}
    scale = {}
    params = {}
    for name in variables:
        value = fields[name]
        if value is None:
            raise ValueError(f'lens has no field {name!r} to optimise')
        if name == 'a':
            powers = torch.arange(2, 2 + value.shape[-1], device=dev, dtype=torch.float32)
            scale[name] = _ASPHERE_UNIT_HEIGHT ** (-2 * powers)
        else:
            scale[name] = torch.ones((), device=dev)
        params[name] = (value.detach() / scale[name]).clone().requires_grad_(True)
    frozen_t = lens.t.detach().clone()
    last = lens.structure.mask_torch.sum(dim=1) - 1
    rows = torch.arange(len(lens), device=dev)
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    history = []
    for step in range(steps):
        opt.zero_grad(set_to_none=True)
        current = dict(fields)
        for name, p in params.items():
            current[name] = p * scale[name]
        if 't' in params and keep_last_thickness:
            t = current['t'].clone()
            t[rows, last] = frozen_t[rows, last]
            current['t'] = t
        trial = Lens(lens.structure, current['c'], current['t'], current['nd'], current['v'],
                     current['k'], current['a'], current['sd'])
        rms, _ = tracer.spot_rms(specs, trial, shard=shard, group=group)
        loss = rms.mean()
        loss.backward()
        opt.step()
        history.append(float(loss.detach()))
        if history[-1] != history[-1]:      # NaN: a poisoned peer exchange (a late or dead rank) must not pass silently
            from .peer import PeerExchange
            if isinstance(group, PeerExchange) and group.status()[0] != 0:
                from ._native import NativeLibraryError
                raise NativeLibraryError(f'peer-memory exchange timed out on rank {group.rank} at step {step}')
        if callback is not None:
            callback(step, history[-1])
    final = dict(fields)
    for name, p in params.items():
        final[name] = (p * scale[name]).detach()
    if 't' in params and keep_last_thickness:
        final['t'][rows, last] = frozen_t[rows, last]
    return Lens(lens.structure, final['c'], final['t'], final['nd'], final['v'], final['k'], final['a'],
                final['sd']), history
