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
