"""Stage the UNMODIFIED reference for the CPU arm of bench.py (`--impl reference`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under torchoptics_b200/ imports this or anything
it produces.

The reference (OceanT-shirt/TorchOptics) is pure Python with no package metadata (no setup.py /
pyproject.toml: `pip install --target baseline/_ref /root/reference` has nothing to build, see
DESIGN.md section 9), and `/root/reference` does not exist on the GPU box.  This script copies the
two files of the hot path, byte for byte,

    /root/reference/torchlens/ray_tracing_lite.py     (RayTracer, trace_skew, compute_rms2d)
    /root/reference/torchlens/lens_modeling.py        (Structure, Specs, Lens)

(and, for the drop-in test of the reference's own front end, optics_simulator_lite.py and optical_loss.py)
into the git-ignored directory `oracle/_ref/torchlens/` (kept out of history like the built `.so`,
but NOT gpurun-ignored, so it travels to the GPU box with the snapshot) and writes next to them a
stand-in for `shapely`, which ray_tracing_lite.py imports at line 15 and never uses (its one use
is commented out, ray_tracing_lite.py:692-694; shapely is not installed in this image).  The
checksums of what was copied go to `oracle/_ref/MANIFEST.json`; `load_reference()` refuses files
whose checksum differs, so the timed code is provably the reference's.

    python oracle/make_ref.py            # run in the build container (needs /root/reference)
"""
import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
DEST = os.path.join(HERE, '_ref')
FILES = ('torchlens/ray_tracing_lite.py', 'torchlens/lens_modeling.py',
         # the reference's own front end of the path (RaytracedOptics.do_ray_tracing, Optical_Loss): never timed, run by
         # tests/test_gpu_dropin.py ON TOP of this repo's CUDA path to show that it keeps working unchanged
         'torchlens/optics_simulator_lite.py', 'torchlens/optical_loss.py')
SHAPELY_STUB = '''"""Stand-in written by oracle/make_ref.py: the reference imports shapely.geometry.Polygon
(ray_tracing_lite.py:15) but its only use is commented out (ray_tracing_lite.py:692-694)."""


class Polygon:      # never instantiated by the live code path
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('shapely is not installed; the reference never calls this on its live path')
'''


def _sha256(path):
    with open(path, 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(force=False):
    """Copy the reference files (only when /root/reference is present).  Returns DEST or None."""
    if not os.path.isdir(REF_ROOT):
        return DEST if os.path.exists(os.path.join(DEST, 'MANIFEST.json')) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha256(src) != _sha256(dst):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha256(dst)
    # (the reference is a namespace package; an empty __init__.py, written here, makes the staged copy a
    # regular one so that it wins over any other `torchlens` on sys.path by path ORDER)
    with open(os.path.join(DEST, 'torchlens', '__init__.py'), 'w') as fh:
        fh.write('')
    os.makedirs(os.path.join(DEST, 'shapely'), exist_ok=True)
    with open(os.path.join(DEST, 'shapely', '__init__.py'), 'w') as fh:
        fh.write('"""see geometry.py"""\n')
    with open(os.path.join(DEST, 'shapely', 'geometry.py'), 'w') as fh:
        fh.write(SHAPELY_STUB)
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': REF_ROOT, 'sha256': manifest}, fh, indent=1)
    return DEST


def available():
    return os.path.exists(os.path.join(DEST, 'MANIFEST.json'))


def load_reference():
    """(ray_tracing_lite, lens_modeling) modules of the staged, checksum-verified reference."""
    if not available():
        raise RuntimeError('oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists')
    with open(os.path.join(DEST, 'MANIFEST.json')) as fh:
        manifest = json.load(fh)['sha256']
    for rel, digest in manifest.items():
        if _sha256(os.path.join(DEST, rel)) != digest:
            raise RuntimeError(f'oracle/_ref/{rel} does not match its manifest: re-stage it')
    if sys.path[0] != DEST:
        sys.path.insert(0, DEST)
    # a `torchlens` (or `shapely`) that came from elsewhere -- e.g. this repo's drop-in alias package --
    # must not stand in for the reference
    for name in [m for m in sys.modules if m.split('.')[0] in ('torchlens', 'shapely')]:
        if not (getattr(sys.modules[name], '__file__', None) or '').startswith(DEST):
            del sys.modules[name]
    rtl = importlib.import_module('torchlens.ray_tracing_lite')
    lm = importlib.import_module('torchlens.lens_modeling')
    assert rtl.__file__.startswith(DEST) and lm.__file__.startswith(DEST), (rtl.__file__, lm.__file__)
    return rtl, lm


if __name__ == '__main__':
    dest = stage(force='--force' in sys.argv)
    print(dest or 'nothing staged: /root/reference is absent and oracle/_ref does not exist')
