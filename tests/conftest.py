import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of failing
    inside them.  (On a GPU box nothing is skipped: a missing library must fail loudly.)"""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=golden_names())
def golden(request):
    rec = load_golden(request.param)
    rec['name'] = request.param
    return rec


def linear_vignetting(fields, vig):
    """The vignetting function of the golden cases (tests/golden/make_golden.py)."""
    return vig[:, None] * fields


def golden_problem(rec, device, requires_grad=True, **tracer_overrides):
    """(tracer, specs, lens) of a golden record exactly as the reference built them: pupil, fields,
    wavelengths, ray aiming (iterations and stop-radius mode) and pupil vignetting."""
    import torch
    from torchoptics_b200 import lens_modeling as lm
    from torchoptics_b200 import ray_tracing_lite as rt
    name = str(rec.get('name', ''))
    structure = lm.Structure(rec['stop_idx'], sequence=rec['sequence'], default_device=device)
    lens = lm.Lens(structure, *[torch.from_numpy(rec[k]).to(device).requires_grad_(requires_grad)
                                for k in ('lens_c', 'lens_t', 'lens_nd', 'lens_v')])
    vig = [torch.from_numpy(np.asarray([v], np.float32)).to(device) for v in rec['vig']] if 'vig' in rec else []
    specs = lm.Specs(structure, torch.from_numpy(rec['epd']).to(device), torch.from_numpy(rec['hfov']).to(device), *vig)
    kwargs = dict(mode='circular', n_rays=tuple(int(v) for v in rec['n_rays']),
                  rel_fields=tuple(float(v) for v in rec['rel_fields']),
                  wavelengths=tuple(float(v) for v in rec['wavelengths']),
                  n_ray_aiming_iter=1 if 'aimed' in name else 0,
                  ray_aiming_mode=str(rec['ray_aiming_mode']) if 'ray_aiming_mode' in rec else 'real',
                  vig_fn=linear_vignetting if 'vig' in rec else None,
                  allow_backward_rays=bool(rec['allow_backward_rays']), default_device=device)
    kwargs.update(tracer_overrides)
    return rt.RayTracer(**kwargs), specs, lens
