"""Lens prescriptions: YAML loader in the reference's schema and the synthetic
benchmark lenses.

Schema (reference fixtures ``torchlens/data/*.yml``): ``stop_idx``, ``sequence``
('G' = surface followed by glass, 'A' = followed by air), ``hfov`` [deg],
``f_number``, ``c``, ``t``, ``nd``, ``v`` -- each a list with one entry (or one
row) per lens.  The reference's own loader branch is dead code
(optics_simulator_lite.py:64-79 never sets the fields ``initialize()`` needs);
this one builds the ``Structure`` / ``Specs`` / ``Lens`` triple directly.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import yaml

from .lens_modeling import Lens, Specs, Structure

LENS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lenses')


def from_dict(d, device='cuda', epd_scale=1.0):
    """Build (specs, lens) for one lens from a prescription dict; EPD = EFL / f_number."""
    from . import ray_tracing_lite as rt
    structure = Structure(np.asarray(d['stop_idx']), sequence=np.asarray(d['sequence']),
                          default_device=device)
    as_t = lambda key: torch.tensor(d[key], dtype=torch.float32, device=device)
    lens = Lens(structure, as_t('c'), as_t('t'), as_t('nd'), as_t('v'))
    efl, _ = rt.get_first_order(lens)
    epd = efl.detach() / as_t('f_number') * epd_scale
    specs = Specs(structure, epd, torch.deg2rad(as_t('hfov')))
    return specs, lens


def load_yaml(path, device='cuda', epd_scale=1.0):
    if not os.path.isabs(path) and not os.path.exists(path):
        path = os.path.join(LENS_DIR, path)
    with open(path) as fh:
        return from_dict(yaml.safe_load(fh), device, epd_scale)


# Double-Gauss 50 mm f/3 (BASELINE.json config 2): classic
# six-element layout scaled by 0.5, 11 surfaces including the flat stop.
_DG_RADII = [54.153, 152.522, 35.951, np.inf, 22.270, np.inf, -25.685, np.inf, -36.980, 196.417, -67.148]
_DG_THICK = [8.747, 0.5, 14.0, 3.777, 14.253, 12.428, 3.777, 10.834, 0.5, 6.858, 57.315]
DOUBLE_GAUSS = {
    'stop_idx': [5],
    'sequence': ['GAGGAAGGAGA'],
    'hfov': [14.0],
    'f_number': [3.0],
    'c': [0.0 if np.isinf(r) else 1.0 / (0.5 * r) for r in _DG_RADII],
    't': [0.5 * t for t in _DG_THICK],
    'nd': [1.60738, 1.62041, 1.60342, 1.60342, 1.62041, 1.62041],
    'v': [56.65, 60.32, 38.03, 38.03, 60.32, 60.32],
}


def double_gauss(device='cuda'):
    return from_dict(DOUBLE_GAUSS, device)


def asphere_12(device='cuda', seed=0, amplitude=0.03, f_number=5.0):
    """Synthetic 12-surface even-asphere lens for BASELINE.json config 3:
    six air-spaced elements ('GAGAGAGAGAGA', stop at the first surface) obtained from the
    Double-Gauss by splitting its two cemented interfaces with 0.1 mm air gaps and dropping the
    stop surface, then given random conic constants in [-0.6, 0.6] and coefficients a4..a16
    sized so that each polynomial term contributes at most `amplitude`/7 of the base sag at the
    marginal height.  8 deg half field; at the default f/5 every ray of a 16-field x 3-wavelength
    x 160^2 set traces (extension oracle), at f/4 99.7 % do -- the remainder runs into the rim of
    the asphere terms, and the few rays next to it dominate (and ill-condition) the gradients."""
    gap = 0.1
    r, th = _DG_RADII, _DG_THICK
    nd, v = DOUBLE_GAUSS['nd'], DOUBLE_GAUSS['v']
    rows = [(r[0], th[0], 0), (r[1], th[1], None), (r[2], th[2], 1), (r[3], 2 * gap, None),
            (r[3], th[3], 2), (r[4], th[4] + th[5], None), (r[6], th[6], 3), (r[7], 2 * gap, None),
            (r[7], th[7], 4), (r[8], th[8], None), (r[9], th[9], 5), (r[10], th[10], None)]
    d = {'stop_idx': [0], 'sequence': [''.join('A' if g is None else 'G' for _, _, g in rows)],
         'hfov': [8.0], 'f_number': [f_number],
         'c': [0.0 if np.isinf(rr) else 1.0 / (0.5 * rr) for rr, _, _ in rows],
         't': [0.5 * tt for _, tt, _ in rows],
         'nd': [nd[g] for _, _, g in rows if g is not None],
         'v': [v[g] for _, _, g in rows if g is not None]}
    specs, lens = from_dict(d, device)
    rng = np.random.default_rng(seed)
    n_surf = len(rows)
    k = rng.uniform(-0.6, 0.6, n_surf)
    a = np.zeros((n_surf, 7))
    height = 7.0
    for s in range(n_surf):
        base = abs(d['c'][s]) * height * height / 2
        for i in range(7):
            draw = rng.uniform(-1, 1)
            a[s, i] = draw * amplitude * base / 7 / height ** (2 * (i + 2)) if base > 0 else 0.0
    lens.k = torch.tensor(k, dtype=torch.float32, device=device).reshape(1, n_surf)
    lens.a = torch.tensor(a, dtype=torch.float32, device=device).reshape(1, n_surf, 7)
    return specs, lens


def wide_zoom_30(device='cuda'):
    """Synthetic 30-surface spherical lens for BASELINE.json config 4: a weak,
    nearly afocal front group of 19 thin surfaces (alternating curvature +-0.005, BK7-like
    singlets 2 mm thick, 1.5 mm air gaps) ahead of the Double-Gauss; stop inside the rear group.
    12 deg half field, f/3.5: every ray traces, ~0.1 % are flagged backward in the front group."""
    n_front, curv = 19, 0.005
    c, t, seq, nd, v = [], [], '', [], []
    for i in range(n_front):
        glass = i % 2 == 0 and i < n_front - 1
        c.append((curv if (i // 2) % 2 == 0 else -curv) * (1.0 if i % 2 == 0 else 0.9))
        t.append(2.0 if glass else 1.5)
        seq += 'G' if glass else 'A'
        if glass:
            nd.append(1.5168)
            v.append(64.17)
    d = DOUBLE_GAUSS
    return from_dict({'stop_idx': [n_front + d['stop_idx'][0]], 'sequence': [seq + d['sequence'][0]],
                      'hfov': [12.0], 'f_number': [3.5], 'c': c + d['c'], 't': t + d['t'],
                      'nd': nd + d['nd'], 'v': v + d['v']}, device)
