"""ctypes binding of the CPU test aid (tests/hostcore/hostcore.cpp)."""
import ctypes

import numpy as np

from .build import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def trace_exact(x, y, z, cx, cy, c, t, mu, live, allow_backward=True):
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    out = [np.empty(n, np.float32) for _ in range(4)] + [np.empty(n, np.uint8) for _ in range(2)]
    lib().hc_trace_exact(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(c.size),
                         _p(c), _p(t), _p(mu), _p(live), ctypes.c_int(int(allow_backward)),
                         *[_p(o) for o in out])
    return out


def fast(dtype, x, y, z, cx, cy, c, t, mu, live, seeds=None):
    """Fast-policy forward (+ adjoint when seeds=(gx,gy,gcx,gcy) is given)."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=dtype)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    S = c.size
    outs = [np.zeros(n, dtype) for _ in range(6)]
    grads = [np.zeros(n, dtype) for _ in range(5)]
    pg = [np.zeros(S, np.float64) for _ in range(3)]
    sd = [None] * 4 if seeds is None else [f(s) for s in seeds]
    fn = lib().hc_fast_f32 if dtype == np.float32 else lib().hc_fast_f64
    fn(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S), _p(c), _p(t), _p(mu),
       _p(live), *[_p(s) for s in sd], *[_p(o) for o in outs], *[_p(g) for g in grads],
       *[_p(g) for g in pg])
    return dict(x=outs[0], y=outs[1], cx=outs[2], cy=outs[3], min_cos2=outs[4], min_travel=outs[5],
                gx=grads[0], gy=grads[1], gz=grads[2], gcx=grads[3], gcy=grads[4],
                gc=pg[0], gt=pg[1], gmu=pg[2])


def rev(dtype, x, y, z, cx, cy, c, t, mu, live, seeds=None, exact_park=False, two_comp=False):
    """Reversible formulation: fast_surface_rev forward (+ sweep_sphere_rev adjoint with seeds)."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=dtype)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    S = c.size
    outs = [np.zeros(n, dtype) for _ in range(6)]
    grads = [np.zeros(n, dtype) for _ in range(5)]
    pg = [np.zeros(S, np.float64) for _ in range(3)]
    sd = [None] * 4 if seeds is None else [f(s) for s in seeds]
    fn = lib().hc_rev_f32 if dtype == np.float32 else lib().hc_rev_f64
    fn(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S), _p(c), _p(t), _p(mu),
       _p(live), *[_p(s) for s in sd], *[_p(o) for o in outs], *[_p(g) for g in grads],
       *[_p(g) for g in pg], ctypes.c_int(int(exact_park) | (2 if two_comp else 0)))
    return dict(x=outs[0], y=outs[1], cx=outs[2], cy=outs[3], min_cos2=outs[4], min_travel=outs[5],
                gx=grads[0], gy=grads[1], gz=grads[2], gcx=grads[3], gcy=grads[4],
                gc=pg[0], gt=pg[1], gmu=pg[2])


def _asph_tables(dtype, c, k, a, t, mu, sd):
    f = lambda v: np.ascontiguousarray(v, dtype=dtype)
    sd2 = np.square(np.asarray(sd, dtype=np.float64)).astype(dtype)
    return f(c), f(k), f(a).reshape(-1), f(t), f(mu), np.ascontiguousarray(sd2)


def asph_exact(x, y, z, cx, cy, c, k, a, t, mu, sd, live, allow_backward=True):
    n = x.size
    f = lambda v: np.ascontiguousarray(v, dtype=np.float32)
    x, y, z, cx, cy = map(f, (x, y, z, cx, cy))
    c, k, a, t, mu, sd2 = _asph_tables(np.float32, c, k, a, t, mu, sd)
    live = np.ascontiguousarray(live, dtype=np.uint8)
    out = [np.empty(n, np.float32) for _ in range(4)] + [np.empty(n, np.uint8) for _ in range(2)] + \
          [np.empty(n, np.float32)]
    lib().hc_asph_exact(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(c.size),
                        _p(c), _p(k), _p(a), _p(t), _p(mu), _p(sd2), _p(live),
                        ctypes.c_int(int(allow_backward)), *[_p(o) for o in out])
    return out


def asph_fast(dtype, x, y, z, cx, cy, c, k, a, t, mu, sd, seeds=None):
    n = x.size
    f = lambda v: np.ascontiguousarray(v, dtype=dtype)
    x, y, z, cx, cy = map(f, (x, y, z, cx, cy))
    c, k, a, t, mu, sd2 = _asph_tables(dtype, c, k, a, t, mu, sd)
    S = c.size
    outs = [np.zeros(n, dtype) for _ in range(7)]
    grads = [np.zeros(n, dtype) for _ in range(5)]
    gp, gt, gmu = np.zeros(S * 9, np.float64), np.zeros(S, np.float64), np.zeros(S, np.float64)
    sd_ = [None] * 5 if seeds is None else [f(s) for s in seeds] + [None] * (5 - len(seeds))      # x, y, cx, cy[, opl]
    fn = lib().hc_asph_fast_f32 if dtype == np.float32 else lib().hc_asph_fast_f64
    fn(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S), _p(c), _p(k), _p(a),
       _p(t), _p(mu), _p(sd2), *[_p(s) for s in sd_], *[_p(o) for o in outs], *[_p(g) for g in grads],
       _p(gp), _p(gt), _p(gmu))
    return dict(x=outs[0], y=outs[1], cx=outs[2], cy=outs[3], opl=outs[4], min_cos2=outs[5],
                min_clip=outs[6], gx=grads[0], gy=grads[1], gz=grads[2], gcx=grads[3], gcy=grads[4],
                gp=gp.reshape(S, 9), gt=gt, gmu=gmu)


def trace_exact_pen(x, y, z, cx, cy, c, t, mu, live, allow_backward=True):
    """Exact-policy aggregate=True terms: z_relu, theta, theta_prime [S, n], ok [n], ok bits [n]."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    S = c.size
    stacks = [np.empty((S, n), np.float32) for _ in range(3)]
    ok = np.empty(n, np.uint8)
    bits = np.empty(n, np.uint32)
    lib().hc_trace_exact_pen(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S),
                             _p(c), _p(t), _p(mu), _p(live), ctypes.c_int(int(allow_backward)),
                             *[_p(s) for s in stacks], _p(ok), _p(bits))
    return stacks, ok, bits


def fast_pen(dtype, x, y, z, cx, cy, c, t, mu, seed_y, seed_zr, seed_th, seed_thp):
    """Fast-policy aggregate=True terms and the adjoint with per-surface seeds [S, n]."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=dtype)
    x, y, z, cx, cy, c, t, mu, seed_y, seed_zr, seed_th, seed_thp = map(
        f, (x, y, z, cx, cy, c, t, mu, seed_y, seed_zr, seed_th, seed_thp))
    S = c.size
    stacks = [np.zeros((S, n), dtype) for _ in range(3)]
    grads = [np.zeros(n, dtype) for _ in range(5)]
    pg = [np.zeros(S, np.float64) for _ in range(3)]
    fn = lib().hc_fast_pen_f32 if dtype == np.float32 else lib().hc_fast_pen_f64
    fn(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S), _p(c), _p(t), _p(mu),
       _p(seed_y), _p(seed_zr), _p(seed_th), _p(seed_thp), *[_p(s) for s in stacks],
       *[_p(g) for g in grads], *[_p(g) for g in pg])
    return dict(z_relu=stacks[0], theta=stacks[1], theta_prime=stacks[2], gx=grads[0], gy=grads[1],
                gz=grads[2], gcx=grads[3], gcy=grads[4], gc=pg[0], gt=pg[1], gmu=pg[2])


def exact_pen_adjoint(dtype, x, y, z, cx, cy, c, t, mu, live, allow_backward, seed_y, seed_zr, seed_th, seed_thp):
    """The PEN backward of the kernels for exact-policy rays (failed rays included)."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    d = lambda a: np.ascontiguousarray(a, dtype=dtype)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    seed_y, seed_zr, seed_th, seed_thp = map(d, (seed_y, seed_zr, seed_th, seed_thp))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    S = c.size
    grads = [np.zeros(n, dtype) for _ in range(5)]
    pg = [np.zeros(S, np.float64) for _ in range(3)]
    fn = lib().hc_exact_pen_adjoint_f32 if dtype == np.float32 else lib().hc_exact_pen_adjoint_f64
    fn(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(S), _p(c), _p(t), _p(mu),
       _p(live), ctypes.c_int(int(allow_backward)), _p(seed_y), _p(seed_zr), _p(seed_th), _p(seed_thp),
       *[_p(g) for g in grads], *[_p(g) for g in pg])
    return dict(gx=grads[0], gy=grads[1], gz=grads[2], gcx=grads[3], gcy=grads[4],
                gc=pg[0], gt=pg[1], gmu=pg[2])


def forward_mode(x, y, z, cx, cy, c, t, mu):
    """Fast-policy image point and its 2x2 Jacobian w.r.t. the pupil point via the D2 pair."""
    n = x.size
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    x, y, z, cx, cy, c, t, mu = map(f, (x, y, z, cx, cy, c, t, mu))
    ox, oy = np.empty(n, np.float32), np.empty(n, np.float32)
    jac = np.empty((n, 4), np.float32)
    lib().hc_forward_mode(ctypes.c_int64(n), _p(x), _p(y), _p(z), _p(cx), _p(cy), ctypes.c_int(c.size), _p(c),
                          _p(t), _p(mu), _p(ox), _p(oy), _p(jac))
    return ox, oy, jac


def paraxial(mode, c, t, n, live, glass, gout=None):
    """csrc/paraxial.cuh on the host: out [B,2] (and gc, gt, gn [B,L] when gout [B,2] is given)."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    c, t, n = map(f, (c, t, n))
    live = np.ascontiguousarray(live, dtype=np.uint8)
    glass = np.ascontiguousarray(glass, dtype=np.uint8)
    B, L = c.shape
    out = np.zeros((B, 2), np.float32)
    grads = [np.zeros((B, L), np.float32) for _ in range(3)]
    g = None if gout is None else f(gout)
    lib().hc_paraxial(ctypes.c_int(B), ctypes.c_int(L), ctypes.c_int(mode), _p(c), _p(t), _p(n), _p(live), _p(glass),
                      _p(g), _p(out), *[_p(v) for v in grads])
    return (out, *grads) if gout is not None else out
