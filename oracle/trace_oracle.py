"""CPU oracle for the sequential skew-ray trace and the spot/RMS reduction.

TEST INFRASTRUCTURE ONLY.  Nothing under ``torchoptics_b200/`` may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker (or as
the CPU arm that is timed *beside* the CUDA path, never instead of it).

It restates, in plain eager PyTorch on whatever dtype the inputs carry (fp32
for parity, fp64 for the "distance to truth" reports), the algorithm of the
reference tracer ``/root/reference/torchlens/ray_tracing_lite.py``:

* ``trace``              <- ``trace_skew``                         rtl:594-675
* ``_march_to_sphere``   <- ``find_marching_distance_spherical``   rtl:525-545
* ``_advance``           <- ``update_ray_coordinates``             rtl:514-522
* ``_refract``           <- ``apply_snell_spherical``              rtl:548-571
* ``_park_failed``       <- ``reset_bad_rays``                     rtl:574-591
* ``spot_rms``           <- ``compute_rms2d``                      rtl:678-702

Every arithmetic statement keeps the reference's operand order and grouping,
because the CUDA kernels' "exact" arithmetic policy reproduces exactly this
sequence of individually rounded IEEE-754 operations and is tested for
*bit-identical* outputs against this file.  Do not "simplify" expressions here.

Pinning: the reference ships no tests or golden vectors, so
the pins are outputs of the unmodified reference itself, generated in the build
container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds this oracle to
them bit-for-bit (masks, points, cosines) and to 1e-6 on gradients.  The
``aggregate=True`` branch (penalty stacks rtl:641-657 and the loss that
``compute_loss_out`` builds on them) is pinned the same way by
``tests/golden/make_golden_aggregate.py`` -> ``tests/golden/aggregate/*.npz`` and
``tests/test_penalty_cpu.py`` (stacks bit-for-bit, penalty / loss gradients to 1e-6,
the reference's NaN pattern included).  Three switches document where a test has to
leave the verbatim reference: ``ieee_sqrt`` (torch's CPU sqrt is not correctly
rounded), ``finite_penalty_gradients`` (the reference's penalty gradient is NaN once a
ray fails) and ``fp32_clamp_bound`` (fp64 runs clamp at 1 - 1e-7, fp32 runs at 1 - 2^-23).
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Tuple

import numpy as np
import torch

# torch's CPU ``sqrt`` is a vectorised <=0.5001-ULP routine: on ~0.5 % of inputs
# it differs by one ULP from the correctly rounded IEEE-754 square root that
# numpy, the host FPU and CUDA's ``sqrtf`` (hence the reference on its default
# device 'cuda', rtl:30) return.  The oracle follows torch by default, which is
# what pins it bit-for-bit to the golden vectors made on the CPU; inside
# ``with ieee_sqrt():`` it uses the correctly rounded root instead, which is
# the arithmetic the CUDA kernels' exact policy reproduces bit-for-bit.
_IEEE_SQRT = False


@contextlib.contextmanager
def ieee_sqrt():
    global _IEEE_SQRT
    previous, _IEEE_SQRT = _IEEE_SQRT, True
    try:
        yield
    finally:
        _IEEE_SQRT = previous


# rtl:646-654 computes sqrt(cos2) for every ray and only afterwards overwrites the failed rays'
# angles with 1; a failed ray's cos2 can be negative, and the backward of that masked-out NaN
# (0 / NaN) poisons every gradient of the penalty with NaN.  Inside
# ``with finite_penalty_gradients():`` the oracle feeds cos2 = 1 to the failed rays instead: the
# forward values and the gradients of the ok rays are unchanged, failed rays contribute 0.  This is
# the behaviour the CUDA path implements (and the only one a test can compare against when rays
# fail); without the switch the oracle is the reference verbatim, NaN included.
_FINITE_PENALTY = False


@contextlib.contextmanager
def finite_penalty_gradients():
    global _FINITE_PENALTY
    previous, _FINITE_PENALTY = _FINITE_PENALTY, True
    try:
        yield
    finally:
        _FINITE_PENALTY = previous


# rtl:645-647 clamps the cosines to +-(1 - 1e-7).  In the reference's fp32 run that bound is the
# float32 nearest to 0.9999999, i.e. 1 - 2^-23 = 0.99999988; an fp64 run of the same code uses
# 0.9999999.  Rays whose angle lies between the two bounds (4.47e-4 .. 4.88e-4 rad) are clamped
# (zero gradient) by one and not by the other, and an unclamped ray there carries a 1 / sin(theta)
# > 2000 gradient.  ``with fp32_clamp_bound():`` makes an fp64 run use the fp32 program's constant,
# so that it is the high-precision evaluation of the fp32 program.
_CLAMP_HI = None


@contextlib.contextmanager
def fp32_clamp_bound():
    global _CLAMP_HI
    previous, _CLAMP_HI = _CLAMP_HI, float(np.float32(1.0) - np.float32(1e-7))
    try:
        yield
    finally:
        _CLAMP_HI = previous


def _sqrt(v):
    if _IEEE_SQRT and v.device.type == 'cpu' and not v.requires_grad:
        return torch.from_numpy(np.sqrt(v.numpy()))
    return torch.sqrt(v)

# rtl:530 and rtl:552 -- the reference uses the same guard for "missed the
# sphere", "total internal reflection" and "direction lost normalisation".
GUARD = 1e-6


def _march_to_sphere(curv, px, py, pz, dx, dy, dz):
    """rtl:525-545.  Closed-form distance along (dx,dy,dz) from (px,py,pz) to the
    vertex-centred sphere of curvature ``curv`` (works for curv == 0)."""
    along = -(px * dx + py * dy + pz * dz)                       # rtl:531
    closest_z = pz + along * dz                                  # rtl:532
    perp2 = px ** 2 + py ** 2 + pz ** 2 - along ** 2             # rtl:533
    aux = curv * perp2 - 2 * closest_z                           # rtl:534
    cos2_inc = dz ** 2 - curv * aux                              # rtl:535
    missed = cos2_inc - GUARD < 0                                # rtl:540
    cos_inc = _sqrt(torch.where(~missed, cos2_inc, 1))      # rtl:541
    dist = along + aux / (dz + cos_inc)                          # rtl:543
    return missed, dist, cos_inc, cos2_inc


def _advance(px, py, pz, dx, dy, dz, dist):
    """rtl:514-522.  Move the point; also hand back the z travel."""
    dz_travel = dist * dz
    px = px + dist * dx
    py = py + dist * dy
    pz = pz + dz_travel
    return px, py, pz, dz_travel


def _refract(curv, ratio, px, py, dx, dy, cos_inc):
    """rtl:548-571.  Scalar-g spherical Snell; ``ratio`` = n / n'."""
    cos2_out = 1 - ratio ** 2 * (1 - cos_inc ** 2)               # rtl:553
    lost = cos2_out - GUARD < 0                                  # rtl:558 (TIR)
    cos_out = _sqrt(torch.where(~lost, cos2_out, 1))        # rtl:559
    g = cos_out - ratio * cos_inc                                # rtl:560
    dx = ratio * dx - g * curv * px                              # rtl:563
    dy = ratio * dy - g * curv * py                              # rtl:564
    dz2 = 1 - (dx ** 2 + dy ** 2)                                # rtl:566
    lost = lost | (dz2 - GUARD < 0)                              # rtl:567
    dz = _sqrt(torch.where(~lost, dz2, 1))                  # rtl:568
    return lost, dx, dy, dz, cos2_out


def _park_failed(alive, px, py, pz, dx, dy, dz):
    """rtl:574-591 (normalize=False branch, the only one trace_skew uses).
    Failed rays are parked on the axis: point (0,0,0), direction (0,0,1)."""
    px = torch.where(alive, px, 0)
    py = torch.where(alive, py, 0)
    pz = torch.where(alive, pz, 0)
    dx = torch.where(alive, dx, 0)
    dy = torch.where(alive, dy, 0)
    dz = torch.where(alive, dz, 1)
    return px, py, pz, dx, dy, dz


def trace(x, y, z, cx, cy, c, t, mu, mask, aggregate: bool = False,
          allow_backward_rays: bool = True):
    """rtl:594-675.  Same arguments, same returns as the reference ``trace_skew``.

    x, y, z, cx, cy : broadcastable to [B, F, P, W]
    c, t, mask      : [B, 1, 1, 1, S];  mu : [B, 1, 1, W, S]
    """
    penalties: Dict[str, List[torch.Tensor]] = {
        'z_RELU': [], 'theta_norm': [], 'theta_prime_norm': []}
    n_surf = t.shape[-1]
    curv = c.unbind(-1)
    gap = t.unbind(-1)
    ratio = mu.unbind(-1)
    live = mask.unbind(-1)

    ray_ok = torch.ones_like(y, dtype=torch.bool)                # rtl:605
    ray_backward = torch.zeros_like(y, dtype=torch.bool)         # rtl:606
    cz = _sqrt(1 - cx ** 2 - cy ** 2)                       # rtl:609

    for k in range(n_surf):                                      # rtl:611
        missed, dist, cos_inc, cos2_inc = _march_to_sphere(curv[k], x, y, z, cx, cy, cz)
        x, y, z, dz_travel = _advance(x, y, z, cx, cy, cz, dist)
        ray_ok = ray_ok & ~missed                                # rtl:619
        x, y, z, cx, cy, cz = _park_failed(ray_ok, x, y, z, cx, cy, cz)
        lost, cx, cy, cz, cos2_out = _refract(curv[k], ratio[k], x, y, cx, cy, cos_inc)
        if k > 0:                                                # rtl:626-632
            counted = ray_ok & live[k - 1]
            if allow_backward_rays:
                ray_backward = ray_backward | ((dz_travel < 0) & counted)
            else:
                ray_ok = ray_ok & ~((dz_travel < 0) & counted)
        ray_ok = ray_ok & ~lost                                  # rtl:635
        x, y, z, cx, cy, cz = _park_failed(ray_ok, x, y, z, cx, cy, cz)
        z = z - gap[k]                                           # rtl:639

        if aggregate:                                            # rtl:641-657
            z_pos = z.clone()
            z_pos[z_pos <= 0] = 0.
            tiny = 1e-7
            if _FINITE_PENALTY:
                cos2_inc = torch.where(ray_ok, cos2_inc, torch.ones_like(cos2_inc))
                cos2_out = torch.where(ray_ok, cos2_out, torch.ones_like(cos2_out))
            hi = 1.0 - tiny if _CLAMP_HI is None else _CLAMP_HI
            ang_in = torch.acos(torch.clamp(_sqrt(cos2_inc), min=-hi, max=hi))
            ang_out = torch.acos(torch.clamp(_sqrt(cos2_out), min=-hi, max=hi))
            ang_in = ang_in / (1 / 2 * math.pi)
            ang_out = ang_out / (1 / 2 * math.pi)
            ang_in[~ray_ok] = 1.
            ang_out[~ray_ok] = 1.
            full = (*x.shape[:3], ratio[0].shape[-1])
            penalties['z_RELU'].append(torch.broadcast_to(z_pos, full))
            penalties['theta_norm'].append(torch.broadcast_to(ang_in, full))
            penalties['theta_prime_norm'].append(torch.broadcast_to(ang_out, full))

    # image plane, rtl:660-670
    dz_travel = -z
    dist = dz_travel / cz
    x = x + dist * cx
    y = y + dist * cy
    counted = ray_ok & live[-1]
    if allow_backward_rays:
        ray_backward = ray_backward | ((dz_travel < 0) & counted)
    else:
        ray_ok = ray_ok & ~((dz_travel < 0) & counted)

    if aggregate:
        return x, y, cx, cy, ray_ok, ray_backward, penalties
    return x, y, cx, cy, ray_ok, ray_backward


def spot_rms(x, y, ray_ok):
    """rtl:678-702.  Mean over fields of the y-only RMS spot radius of lens 0.

    The centroid is the mean over *all* rays of the field (parked rays sit at
    y = 0 and are included, rtl:695-697); the squared deviations are summed over
    surviving rays only but divided by the full count P*W (rtl:699)."""
    n_field, n_pupil, n_wave = y.shape[1], y.shape[2], y.shape[3]
    total = 0.
    for f in range(n_field):
        centroid_sum = 0.
        for w in range(n_wave):
            centroid_sum = centroid_sum + torch.mean(y[0, f, :, w])
        centroid = centroid_sum / n_wave
        kept = y[0, f, :, :][ray_ok[0, f, :, :]]
        total = total + _sqrt(torch.sum((kept - centroid) ** 2) / (n_pupil * n_wave))
    return total / n_field


def spot_rms_all_lenses(y, ray_ok):
    """Vectorised restatement of :func:`spot_rms` for every lens of the batch
    (the reference hard-codes lens 0, rtl:695/:699).  Returns [B].  Used to check
    the batched CUDA reduction; lens 0 of the result equals ``spot_rms``
    up to summation order."""
    n_pupil, n_wave = y.shape[2], y.shape[3]
    centroid = y.mean(dim=2).mean(dim=2)                         # [B, F]
    dev2 = torch.where(ray_ok, (y - centroid[:, :, None, None]) ** 2, torch.zeros_like(y))
    per_field = _sqrt(dev2.sum(dim=(2, 3)) / (n_pupil * n_wave))
    return per_field.mean(dim=1)


def trace_and_rms(x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays: bool = True):
    """The fwd+bwd hot path as one call (trace -> RMS spot loss), for timing the
    CPU baseline: returns (rms, outputs)."""
    out = trace(x, y, z, cx, cy, c, t, mu, mask, False, allow_backward_rays)
    return spot_rms(out[0], out[1], out[4]), out
