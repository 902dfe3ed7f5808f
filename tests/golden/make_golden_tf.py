"""Golden vectors for the functions that exist only in the reference's TensorFlow original
(`/root/reference/torchlens/ray_tracing.py`, "rt_tf"; commented out or missing in ray_tracing_lite.py):
the pupil samplers rt_tf:358-476, `apply_vignetting` rt_tf:479-490, `compute_magnification`
rt_tf:765-777 and the Gaussian soft-histogram PSF `compute_psf` rt_tf:206-270.

TensorFlow cannot be installed here (no network), so the pins are produced by executing the
reference's OWN SOURCE FILE with a small numpy stand-in registered as the `tensorflow` module: every
`tf.*` call these functions make (reshape, linspace, constant, cos, sin, exp, reduce_*, concat,
reverse, range, ...) maps one-to-one onto the numpy function of the same meaning, in float32 where TF
would compute in float32.  What is pinned is therefore the reference's algorithm statement by
statement; what is not is TF's own rounding of transcendental functions (a few ULP).

    python tests/golden/make_golden_tf.py        # build container only (needs /root/reference)
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'tf')
REF = '/root/reference/torchlens/ray_tracing.py'


def tf_shim():
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.float64, tf.int32, tf.bool = np.float32, np.float64, np.int32, np.bool_
    f32 = lambda v: np.asarray(v, dtype=np.float32) if np.asarray(v).dtype.kind == 'f' else np.asarray(v)

    def constant(value, dtype=None):
        arr = np.asarray(value)
        if dtype is not None:
            return arr.astype(dtype)
        return arr.astype(np.float32) if arr.dtype.kind == 'f' else arr
    tf.constant = constant
    tf.reshape = lambda t, shape: np.reshape(t, shape)
    tf.linspace = lambda a, b, n: np.linspace(a, b, n).astype(np.float32)
    tf.zeros_like, tf.ones_like = np.zeros_like, np.ones_like
    tf.zeros = lambda shape, dtype=np.float32: np.zeros(shape, dtype)
    tf.ones = lambda shape, dtype=np.float32: np.ones(shape, dtype)
    tf.range = lambda n, dtype=np.int32: np.arange(n, dtype=dtype)
    tf.cos, tf.sin, tf.sqrt, tf.exp, tf.abs, tf.tan = np.cos, np.sin, np.sqrt, np.exp, np.abs, np.tan
    tf.maximum, tf.minimum = np.maximum, np.minimum
    tf.reduce_mean = lambda t, axis=None, keepdims=False: np.mean(t, axis=axis, keepdims=keepdims, dtype=np.float32) \
        if np.asarray(t).dtype == np.float32 else np.mean(t, axis=axis, keepdims=keepdims)
    tf.reduce_sum = lambda t, axis=None, keepdims=False: np.sum(t, axis=axis, keepdims=keepdims)
    tf.reduce_min = lambda t, axis=None: np.min(t, axis=axis)
    tf.reduce_max = lambda t, axis=None: np.max(t, axis=axis)
    tf.reduce_prod = lambda t: int(np.prod(t))
    tf.concat = lambda parts, axis: np.concatenate(parts, axis=axis)
    tf.stack = lambda parts, axis=0: np.stack(parts, axis=axis)
    tf.reverse = lambda t, axis: np.flip(t, axis=axis)
    tf.cast = lambda t, dtype: np.asarray(t).astype(dtype)
    tf.squeeze = lambda t, axis=None: np.squeeze(t, axis=axis)
    tf.where = np.where
    rnd = types.SimpleNamespace(uniform=lambda shape: np.random.default_rng(0).random(shape, dtype=np.float32))
    tf.random = rnd
    tf.__dict__['_f32'] = f32
    return tf


def import_rt_tf():
    shapely = types.ModuleType('shapely')
    geometry = types.ModuleType('shapely.geometry')
    geometry.Polygon = object
    shapely.geometry = geometry
    saved = {k: sys.modules.get(k) for k in ('tensorflow', 'shapely', 'shapely.geometry')}
    sys.modules.update({'tensorflow': tf_shim(), 'shapely': shapely, 'shapely.geometry': geometry})
    try:
        spec = importlib.util.spec_from_file_location('rt_tf_reference', REF)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return module


def main():
    os.makedirs(OUT, exist_ok=True)
    rt = import_rt_tf()
    rec = {}
    # ---- pupil samplers (rt_tf:358-476) ----
    for n_r, n_i in ((1, 1), (3, 2), (4, 4), (8, 8), (5, 3)):
        for name in ('skew_uniform_half_equidistant', 'skew_uniform_half_jittered'):
            x, y = getattr(rt, name)(None, n_r, n_i)
            rec[f'{name}_{n_r}_{n_i}_x'], rec[f'{name}_{n_r}_{n_i}_y'] = x, y
    for n_y in (2, 5, 8):
        x, y = rt.skew_inner_square_half(None, n_y, None)
        rec[f'skew_inner_square_half_{n_y}_x'], rec[f'skew_inner_square_half_{n_y}_y'] = x, y
    for n in (1, 2, 7, 16):
        for name in ('meridional_uniform', 'sagittal_uniform', 'circle_outer_edge_uniform'):
            if name != 'circle_outer_edge_uniform' and n == 1:
                continue
            x, y = getattr(rt, name)(None, n)
            rec[f'{name}_{n}_x'], rec[f'{name}_{n}_y'] = x, y
    x, y = rt.chief(None, None)
    rec['chief_x'], rec['chief_y'] = x, y
    x, y = rt.tee(None)
    rec['tee_x'], rec['tee_y'] = x, y
    # ---- apply_vignetting (rt_tf:479-490) ----
    rng = np.random.default_rng(3)
    yy = rng.uniform(-1, 1, (2, 3, 9, 1)).astype(np.float32)
    vig_up = rng.uniform(0, 0.4, (2, 3)).astype(np.float32)
    vig_down = rng.uniform(0, 0.3, (2, 3)).astype(np.float32)
    rec['vig_in_y'], rec['vig_up'], rec['vig_down'] = yy, vig_up, vig_down
    rec['vig_out'] = rt.apply_vignetting(yy, vig_up, vig_down)
    # ---- compute_magnification (rt_tf:765-777) on the front groups of the four shipped lenses ----
    import yaml
    for name in ('baseline_cooke', 'baseline_tessar', 'baseline_doublet'):
        with open(f'/root/reference/torchlens/data/{name}.yml') as fh:
            d = yaml.safe_load(fh)
        seq = d['sequence'][0]
        stop = d['stop_idx'][0]
        c = np.asarray(d['c'], np.float32).reshape(1, -1)[:, :stop]
        t = np.asarray(d['t'], np.float32).reshape(1, -1)[:, :stop]
        nd_glass = iter(np.asarray(d['nd'], np.float32).ravel())
        nd = np.asarray([[next(nd_glass) if ch == 'G' else 1.0 for ch in seq]], np.float32)[:, :stop]
        lens = types.SimpleNamespace(c=c, t=t, nd=nd)
        rec[f'magnification_{name}'] = rt.compute_magnification(lens)
        rec[f'magnification_{name}_c'], rec[f'magnification_{name}_t'], rec[f'magnification_{name}_nd'] = c, t, nd
    np.savez_compressed(os.path.join(OUT, 'samplers_vignetting_magnification.npz'), **rec)
    print(f'{len(rec)} arrays -> tf/samplers_vignetting_magnification.npz')
    # ---- compute_psf (rt_tf:206-270) on traced spots of the reference itself ----
    psf = {}
    # (rt_tf:267 compares [lens, field, channel, ray] arrays with per-grid sizes of shape [n_grids]: without
    # `increment` the reference itself only runs for ONE grid; its consumer, sample_psfs
    # optics_simulator_lite.py:656-677, always passes increment and y_target)
    cases = (('cooke_32x32', None, dict(n_bins=(21, 21), increment=0.004)),
             ('cooke_32x32', None, dict(n_bins=(11, 15), increment=0.004, y_target='centroid+')),
             ('cooke_32x32', None, dict(n_bins=(10, 12), increment=0.003)),
             ('cooke_16x16_epd2.0', None, dict(n_bins=(8, 9), increment=0.05)),
             ('tessar_8x8', None, dict(n_bins=(5, 6), increment=0.01)),
             ('cooke_32x32', 2, dict(n_bins=(21, 21))),
             ('tessar_8x8', 1, dict(n_bins=(6, 7))))
    for case, field, kwargs in cases:
        with np.load(os.path.join(HERE, case + '.npz')) as g:
            x = np.transpose(g['out_x'], (0, 1, 3, 2)).copy()     # [lens, field, channel, ray]
            y = np.transpose(g['out_y'], (0, 1, 3, 2)).copy()
        if field is not None:
            x, y = x[:, field:field + 1], y[:, field:field + 1]
        kwargs = dict(kwargs)
        if kwargs.get('y_target') == 'centroid+':      # a caller-given target, as sample_psfs passes one
            kwargs['y_target'] = (y.reshape(x.shape[0] * x.shape[1], -1).mean(axis=1) + 0.001).astype(np.float32)
        tag = f"{case}_f{'all' if field is None else field}_bins{kwargs['n_bins'][0]}x{kwargs['n_bins'][1]}" + \
              ('_incr' if 'increment' in kwargs else '') + ('_target' if 'y_target' in kwargs else '')
        x_size, y_size, y_target, kernels, accounted = rt.compute_psf(x, y, **kwargs)
        psf[tag + '_in_x'], psf[tag + '_in_y'] = x, y
        psf[tag + '_n_bins'] = np.asarray(kwargs['n_bins'])
        psf[tag + '_increment'] = np.asarray(kwargs.get('increment', np.nan), np.float64)
        if 'y_target' in kwargs:
            psf[tag + '_in_y_target'] = kwargs['y_target']
        psf[tag + '_x_size'] = np.asarray(x_size, np.float64)
        psf[tag + '_y_size'] = np.asarray(y_size, np.float64)
        psf[tag + '_y_target'] = np.asarray(y_target)
        psf[tag + '_kernels'] = np.asarray(kernels)
        psf[tag + '_accounted'] = np.asarray(accounted)
        print(tag, 'kernels', np.asarray(kernels).shape, 'mass', float(np.asarray(kernels)[0].sum()),
              'accounted', np.asarray(accounted).ravel()[:3])
    np.savez_compressed(os.path.join(OUT, 'psf.npz'), **psf)


if __name__ == '__main__':
    main()
