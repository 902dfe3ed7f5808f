"""CPU oracle of the EXTENSION rows of the hot path:
conic / even-asphere surfaces solved by Newton iteration, vector Snell refraction
with the general surface normal, clear semi-diameter clipping and optical path
length accumulation.

TEST INFRASTRUCTURE ONLY (same rules as trace_oracle.py).

PARITY UNPINNED: the reference (/root/reference/torchlens) is spherical-only -- it
has no conic constant, no asphere coefficients, no iterative intersection, no OPD
and no aperture clipping anywhere -- so there is no reference behaviour, test or
golden vector to pin these rows to.  This file DEFINES the behaviour, in the
reference's conventions (vertex-centred coordinates, the `z -= t` shift of
rtl:639, the 1e-6 guards of rtl:530/:552, sticky ray_ok + parking of rtl:574-591,
backward-ray flag of rtl:626-632), and is itself checked by
tests/test_asphere_oracle.py through properties that do not depend on it:

* k = 0, a = 0, sd = inf  ->  the results of the pinned spherical oracle,
* fp64 gradcheck of the whole trace,
* a paraboloid (k = -1) focuses an on-axis collimated bundle to a point and all
  optical path lengths to the focus are equal,
* the sag equation is satisfied at every hit point.

Surface: z = s(rho), rho = x^2 + y^2,
    s(rho) = c rho / (1 + sqrt(1 - (1+k) c^2 rho)) + sum_{i=2..8} a_i rho^i
(a_2..a_8 multiply r^4..r^16), ds/drho = c / (2 sqrt(1 - (1+k) c^2 rho)) + sum i a_i rho^(i-1).
"""
from __future__ import annotations

import torch

from . import trace_oracle as sph

GUARD = sph.GUARD
N_COEF = 7            # a4, a6, ..., a16
N_NEWTON = 4          # fixed iteration count (deterministic; the start is the base-sphere hit)


def _poly(a, rho):
    """sum_{i=2..8} a_i rho^i and its rho-derivative, Horner form.  a: [..., 7]."""
    a4, a6, a8, a10, a12, a14, a16 = a.unbind(-1)
    inner = a16
    d_inner = 8 * a16
    for coef, mult in ((a14, 7), (a12, 6), (a10, 5), (a8, 4), (a6, 3), (a4, 2)):
        inner = inner * rho + coef
        d_inner = d_inner * rho + mult * coef
    return inner * rho * rho, d_inner * rho


def sag_and_slope(c, k, a, rho):
    """s(rho), ds/drho and the conic radicand 1 - (1+k) c^2 rho."""
    radicand = 1 - (1 + k) * c * c * rho
    safe = torch.where(radicand - GUARD < 0, torch.ones_like(radicand), radicand)
    root = sph._sqrt(safe)
    poly, dpoly = _poly(a, rho)
    s = c * rho / (1 + root) + poly
    ds = c / (2 * root) + dpoly
    return s, ds, radicand


def _intersect(c, k, a, px, py, pz, dx, dy, dz):
    """Distance along the ray to the surface: base-sphere closed form (rtl:525-545), then
    N_NEWTON Newton steps on F(tau) = pz + tau dz - s(rho(tau)).  The last step carries the
    autograd graph so that d tau / d(anything) is the implicit-function derivative."""
    missed, tau, _, _ = sph._march_to_sphere(c, px, py, pz, dx, dy, dz)
    tau = torch.where(missed, torch.zeros_like(tau), tau)
    for _ in range(N_NEWTON):
        tau = tau.detach()       # d tau comes from the last step alone: -dF/F' (implicit function theorem)
        hx = px + tau * dx
        hy = py + tau * dy
        rho = hx * hx + hy * hy
        s, ds, _ = sag_and_slope(c, k, a, rho)
        f = pz + tau * dz - s
        fp = dz - ds * 2 * (hx * dx + hy * dy)
        tau = tau - f / fp
    return missed, tau


def trace(x, y, z, cx, cy, c, t, mu, mask, k=None, a=None, sd=None, allow_backward_rays=True):
    """General-surface `trace_skew`.  c, t, mask, k, sd: [B,1,1,1,S]; mu: [B,1,1,W,S];
    a: [B,1,1,1,S,7].  Returns (x, y, cx, cy, ray_ok, ray_backward, opl) with opl the optical
    path length from the entrance point to the image plane."""
    n_surf = t.shape[-1]
    if k is None:
        k = torch.zeros_like(c)
    if a is None:
        a = torch.zeros(c.shape + (N_COEF,), dtype=c.dtype, device=c.device)
    if sd is None:
        sd = torch.full_like(c, float('inf'))
    ray_ok = torch.ones(torch.broadcast_shapes(x.shape, y.shape, z.shape, cx.shape, cy.shape,
                                               mu.shape[:-1]), dtype=torch.bool, device=y.device)
    ray_backward = torch.zeros_like(ray_ok)
    cz = sph._sqrt(1 - cx ** 2 - cy ** 2)
    index = torch.ones_like(mu[..., 0])                      # n of the current medium (air in front)
    opl = torch.zeros_like(ray_ok, dtype=y.dtype)
    for s_i in range(n_surf):
        cs, ks, as_, ts = c[..., s_i], k[..., s_i], a[..., s_i, :], t[..., s_i]
        mus, sds = mu[..., s_i], sd[..., s_i]
        missed, tau = _intersect(cs, ks, as_, x, y, z, cx, cy, cz)
        travel = tau * cz
        x = x + tau * cx
        y = y + tau * cy
        z = z + travel
        rho = x * x + y * y
        _, ds, radicand = sag_and_slope(cs, ks, as_, rho)
        failed = missed | (radicand - GUARD < 0) | (rho > sds * sds) | ~torch.isfinite(tau)
        ray_ok = ray_ok & ~failed
        opl = opl + torch.where(ray_ok, index * tau, torch.zeros_like(tau))
        x, y, z, cx, cy, cz = sph._park_failed(ray_ok, x, y, z, cx, cy, cz)
        ds = torch.where(ray_ok, ds, torch.zeros_like(ds))
        # unit normal (towards +z) and vector Snell
        nx, ny = -2 * x * ds, -2 * y * ds
        inv_norm = 1 / sph._sqrt(nx * nx + ny * ny + 1)
        nx, ny, nz = nx * inv_norm, ny * inv_norm, inv_norm
        cos_in = cx * nx + cy * ny + cz * nz
        cos2_out = 1 - mus ** 2 * (1 - cos_in ** 2)
        lost = cos2_out - GUARD < 0
        cos_out = sph._sqrt(torch.where(~lost, cos2_out, torch.ones_like(cos2_out)))
        g = cos_out - mus * cos_in
        cx = mus * cx + g * nx
        cy = mus * cy + g * ny
        cz2 = 1 - (cx ** 2 + cy ** 2)
        lost = lost | (cz2 - GUARD < 0)
        cz = sph._sqrt(torch.where(~lost, cz2, torch.ones_like(cz2)))
        if s_i > 0:
            counted = ray_ok & mask[..., s_i - 1]
            if allow_backward_rays:
                ray_backward = ray_backward | ((travel < 0) & counted)
            else:
                ray_ok = ray_ok & ~((travel < 0) & counted)
        ray_ok = ray_ok & ~lost
        x, y, z, cx, cy, cz = sph._park_failed(ray_ok, x, y, z, cx, cy, cz)
        z = z - ts
        index = index / mus
    travel = -z
    dist = travel / cz
    x = x + dist * cx
    y = y + dist * cy
    opl = opl + torch.where(ray_ok, index * dist, torch.zeros_like(dist))
    counted = ray_ok & mask[..., -1]
    if allow_backward_rays:
        ray_backward = ray_backward | ((travel < 0) & counted)
    else:
        ray_ok = ray_ok & ~((travel < 0) & counted)
    return x, y, cx, cy, ray_ok, ray_backward, opl


def opd(x, y, cx, cy, opl, ray_ok, n_image, radius, chief=0):
    """Optical path DIFFERENCE of row A10 from the outputs of `trace`: every ray's path is measured up to the
    reference sphere -- centre = the image point of the field's chief ray (pupil index `chief`), radius `radius`
    (> 0: the sphere lies in front of the image plane, normally through the exit pupil's axial point) -- and the
    chief ray's own path to that sphere is subtracted.

    A ray arrives at the image plane at P = (x, y, 0) with direction d; the sphere is met at P + s d, s < 0 the
    root of |P - C + s d| = R that lies in front of the plane; path to the sphere = opl + n_image s; for the chief
    ray itself s = -R.  x, y, cx, cy, opl, ray_ok: [B,F,P,W]; n_image: [B,1,1,W] (index of the image space:
    the product of 1 / mu over the surfaces); radius: broadcastable to [B,F,1,W].  Rays that are not ok -- and
    every ray of a field whose chief ray is not ok -- get 0."""
    cz = sph._sqrt(1 - cx ** 2 - cy ** 2)
    pick = slice(chief, chief + 1)
    dx, dy = x - x[:, :, pick], y - y[:, :, pick]
    along = dx * cx + dy * cy                                  # (P - C) . d
    disc = along * along - (dx * dx + dy * dy) + radius * radius
    s = -along - sph._sqrt(disc)
    out = (opl - opl[:, :, pick]) + n_image * (s + radius)
    good = ray_ok & ray_ok[:, :, pick]
    return torch.where(good, out, torch.zeros_like(out))
