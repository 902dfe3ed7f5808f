"""Sequential ray tracing of batched lenses -- host side of the CUDA hot path.

Drop-in for the reference module ``torchlens/ray_tracing_lite.py`` ("rtl"): the
same public names, arguments and tensor conventions

    dim 0 lenses (B), dim 1 fields (F), dim 2 pupil points (P), dim 3 wavelengths (W),
    dim 4 surfaces (S)                                                      rtl:4-9

but :func:`trace_skew` (rtl:594-675) and :func:`compute_rms2d` (rtl:678-702) run as
hand-written sm_100a kernels behind ``torchoptics_b200.ops``; this file only
prepares their (tiny) inputs: refractive-index ratios, paraxial pupil position,
pupil grids and field angles.  ``RayTracer.spot_rms`` is the fused
trace -> RMS -> gradient pass that the reference spells as
``trace_rays`` + ``compute_rms2d`` + ``.backward()``.

Only CUDA tensors are accepted by the traced functions (no CPU path).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _native as nat
from . import ops
from .lens_modeling import mask_replace  # noqa: F401  (re-exported like the reference, rtl:18)

WAVELENGTH_ALIASES = {'C': 656.3, 'd': 587.6, 'F': 486.1}   # rtl:71-75


# ---------------------------------------------------------------------------
# Pupil samplers.  Each returns (x, y) in relative pupil units, shape [1|n,1,P,1].
# ---------------------------------------------------------------------------
def tee(tensor=None, device='cuda'):
    """Lower / upper meridional and one sagittal ray (rtl:353-360)."""
    y = torch.tensor([-1., 1., 0.], device=device).reshape(1, 1, 3, 1)
    x = torch.tensor([0., 0., 1.], device=device).reshape(1, 1, 3, 1)
    return x, y


def chief(tensor=None, n_rays=None, device='cuda'):
    """The pupil-centre ray (rt_tf:378-385)."""
    zero = torch.zeros((1, 1, 1, 1), device=device)
    return zero, zero.clone()


def meridional_uniform(tensor, n_rays, device='cuda'):
    """``n_rays`` equidistant points on the pupil's y axis (rt_tf:358-365)."""
    y = torch.linspace(-1., 1., n_rays, device=device).reshape(1, 1, -1, 1)
    return torch.zeros_like(y), y


def sagittal_uniform(tensor, n_rays, device='cuda'):
    """``n_rays`` equidistant points on the POSITIVE half of the pupil's x axis, 0 ... 1
    (rt_tf:368-375; the system is symmetric about the meridional plane)."""
    x = torch.linspace(0., 1., n_rays, device=device).reshape(1, 1, -1, 1)
    return x, torch.zeros_like(x)


def _half_shells(n_r, n_i):
    """Shell index and polar angle of the (n_r ** 2) * n_i points both half-pupil samplers share
    (rt_tf:425-428, :440-445): shell i carries n_i * (2 i + 1) points, equidistant in angle over the
    right half circle (-pi/2, pi/2)."""
    per_shell = [n_i * (2 * i + 1) for i in range(n_r)]
    shell = np.array([i for i in range(n_r) for _ in range(per_shell[i])])
    theta = np.array([(j / n - 0.5) * np.pi for n in per_shell for j in (np.arange(n) + 0.5)])
    return shell, theta


def _points(r, theta, device):
    x = torch.from_numpy((r * np.cos(theta)).astype(np.float32)).to(device)
    y = torch.from_numpy((r * np.sin(theta)).astype(np.float32)).to(device)
    return x.reshape(1, 1, -1, 1), y.reshape(1, 1, -1, 1)


def skew_uniform_half_equidistant(tensor, n_r, n_i, device='cuda'):
    """(n_r ** 2) * n_i points spanning the right half of the pupil uniformly: shell i at radius
    (i + 1/2) / n_r (rt_tf:421-433)."""
    shell, theta = _half_shells(n_r, n_i)
    return _points(((np.arange(n_r) + 0.5) / n_r)[shell], theta, device)


def skew_uniform_half_jittered(tensor, n_r, n_i, device='cuda'):
    """The same shells with the radius alternating between the inner and the outer edge of each shell,
    so that the pupil rim is sampled (rt_tf:436-451) -- the default sampler of the reference's
    ``RaytracedOptics`` (optics_simulator_lite.py:360)."""
    shell, theta = _half_shells(n_r, n_i)
    inner = np.linspace(0, 1, n_r * 2)[::2]
    step = 1 / (2 * n_r - 1)
    r = inner[shell] + step * ((np.arange(len(shell)) + shell) % 2)
    return _points(r, theta, device)


def skew_inner_square_half(tensor, n_y, _=None, device='cuda'):
    """n_y x n_y grid on the right half of the square inscribed in the pupil (rt_tf:454-465):
    x in (0, 1/sqrt 2], y in [-1/sqrt 2, 1/sqrt 2], row-major in y."""
    x = (np.linspace(-1, 1, n_y * 2)[-n_y:] / np.sqrt(2)).astype(np.float32)
    y = (np.linspace(-1, 1, n_y) / np.sqrt(2)).astype(np.float32)
    xx = np.broadcast_to(x[None, :], (n_y, n_y)).reshape(-1).copy()
    yy = np.broadcast_to(y[:, None], (n_y, n_y)).reshape(-1).copy()
    return (torch.from_numpy(xx).to(device).reshape(1, 1, -1, 1),
            torch.from_numpy(yy).to(device).reshape(1, 1, -1, 1))


def circle(tensor, n_r, n_theta, default_device='cuda'):
    """Polar grid: radii i/n_r (the first ring is the centre), angles 2 pi j/n_theta,
    ring-major (rtl:412-422).  Built in fp32 exactly like the reference so that the
    pupil points are bit-identical."""
    radius = torch.from_numpy(np.linspace(0, 1.0, n_r, endpoint=False, dtype=np.float32))
    angle = torch.from_numpy(np.linspace(0, 2 * np.pi, n_theta, endpoint=False, dtype=np.float32))
    radius = radius.to(default_device)[:, None]
    angle = angle.to(default_device)[None, :]
    x = radius * torch.cos(angle)
    y = radius * torch.sin(angle)
    return x.reshape(1, 1, n_r * n_theta, 1), y.reshape(1, 1, n_r * n_theta, 1)


def circle_pseudo_random(tensor, n_r, n_theta, device='cuda'):
    """Jittered equal-area polar cells, one independent draw per element of
    ``tensor`` (rtl:393-410)."""
    n_sets = int(np.prod(tensor.shape))
    jitter_r2 = torch.rand((n_sets, n_r, n_theta)) / n_r
    jitter_angle = torch.rand((n_sets, n_r, n_theta)) / n_theta
    r2_start = torch.from_numpy(np.linspace(0, 1, n_r, endpoint=False, dtype=np.float32))[None, :, None]
    angle_start = torch.from_numpy(np.linspace(0, 1, n_theta, endpoint=False, dtype=np.float32))[None, None, :]
    radius = torch.sqrt(jitter_r2 + r2_start)
    angle = (jitter_angle + angle_start) * 2 * np.pi
    x = (radius * torch.cos(angle)).reshape(n_sets, 1, n_r * n_theta, 1)
    y = (radius * torch.sin(angle)).reshape(n_sets, 1, n_r * n_theta, 1)
    return x.to(device), y.to(device)


def circle_outer_edge_uniform(tensor, n_rays, device='cuda'):
    """``n_rays`` points on the pupil rim (rt_tf:468-476)."""
    angle = torch.from_numpy(np.linspace(0, 2 * np.pi, n_rays, endpoint=False, dtype=np.float32)).to(device)
    return torch.cos(angle).reshape(1, 1, -1, 1), torch.sin(angle).reshape(1, 1, -1, 1)


def apply_vignetting(y, vig_up, vig_down):
    """Squeeze relative pupil coordinates by the vignetting factors (rt_tf:479-490)."""
    vig_up = vig_up[..., None, None]
    vig_down = vig_down[..., None, None]
    return y * (1 - (vig_up + vig_down) / 2) + (vig_down - vig_up) / 2


def scale_to_epd(y, epd):
    """Relative pupil coordinate -> length units: y * EPD / 2 (rtl:497-507)."""
    return y * epd.reshape(-1, *([1] * (y.dim() - 1))) / 2


# ---------------------------------------------------------------------------
# Paraxial optics (per-lens scalars, plain torch)
# ---------------------------------------------------------------------------
def interface_propagation_abcd(c, t, n):
    """ABCD matrix of 'refract at curvature c, then travel t' for every surface:
    [[1 + t C, t D], [C, D]] with D = n/n', C = c (D - 1).  c, t: [B,S]; n: [B,S+1].
    Returns [B,S,2,2] (rtl:314-327)."""
    assert n.shape[-1] - 1 == c.shape[-1] == t.shape[-1]
    ratio = n[:, :-1] / n[:, 1:]
    power = c * (ratio - 1)
    rows = torch.stack((1 + power * t, ratio * t, power, ratio), dim=-1)
    return rows.reshape(n.shape[0], -1, 2, 2)


def reduce_abcd(abcd):
    """Ordered product M_{S-1} ... M_1 M_0 of [B,S,2,2] matrices -> [B,2,2], by
    pairwise (log-depth) multiplication (rtl:301-311)."""
    while abcd.shape[1] > 1:
        n_pairs = abcd.shape[1] // 2
        paired = abcd[:, 1:2 * n_pairs:2] @ abcd[:, 0:2 * n_pairs:2]
        abcd = paired if abcd.shape[1] % 2 == 0 else torch.cat((paired, abcd[:, -1:]), dim=1)
    return abcd[:, 0]


def compute_pupil_position(lens):
    """Paraxial entrance-pupil position relative to the first vertex: B/A of the
    system in front of the stop (rtl:330-350).  Differentiable w.r.t. the lens.

    Works on the padded [B, Lmax] tensors directly (slots at or behind the stop
    become identity matrices), so it launches no data-dependent-shape op and can be
    captured in a CUDA graph."""
    structure = lens.structure
    n_front = int(structure.stop_idx.max())
    if n_front == 0:
        return torch.zeros(len(lens), device=lens.c.device)
    front = structure.up_to_stop()
    keep, keep_g = front.mask_torch, front.mask_G_torch
    c = torch.where(keep, lens.c[:, :n_front], torch.zeros_like(lens.c[:, :n_front]))
    t = torch.where(keep, lens.t[:, :n_front], torch.zeros_like(lens.t[:, :n_front]))
    nd = torch.where(keep_g, lens.nd[:, :n_front], torch.ones_like(lens.nd[:, :n_front]))
    nd = torch.cat((torch.ones_like(nd[:, 0:1]), nd), dim=1)
    system = reduce_abcd(interface_propagation_abcd(c, t, nd))
    return system[:, 0, 1] / system[:, 0, 0]


def get_first_order(lens):
    """(EFL, BFL) of every lens (rtl:772-794).  CUDA lenses: one kernel, one thread per lens
    (``tl_paraxial_fwd``; its adjoint ``tl_paraxial_bwd`` in backward) instead of ~20 eager ops."""
    if lens.c.is_cuda and lens.c.shape[1] <= nat.PARAXIAL_MAX_SLOTS:
        return ops.first_order(lens.structure, lens.c, lens.t, lens.nd)
    nd = torch.cat((torch.ones_like(lens.nd[:, 0:1]), lens.nd), dim=1)
    last = lens.structure.mask_torch.sum(dim=1) - 1
    t = lens.t.clone()
    t[torch.arange(len(lens), device=t.device), last] = 0.
    system = reduce_abcd(interface_propagation_abcd(lens.c, t, nd))
    return -1 / system[:, 1, 0], -system[:, 0, 0] / system[:, 1, 0]


def extraction_from_indices(params, indices):
    """``params[i, j]`` for the single (i, j) row of ``indices`` (rtl:705-722)."""
    assert indices.dim() == 2 and indices.shape == (1, 2)
    assert params.dim() == 2 and params.shape[0] == 1
    indices = indices.long()
    return params[indices[:, 0], indices[:, 1]]


def compute_last_curvature(structures, c, t, nd):
    """Compact curvatures with the last free curvature solved so that EFL = 1
    (rtl:725-769).  ``c`` holds every curvature but the last surface's; a trailing
    air-air surface (e.g. a cover-glass back) keeps curvature as given and the
    solve moves one surface forward."""
    device = structures.mask_torch.device
    mask = structures.mask_torch
    B = mask.shape[0]
    if c.is_cuda and mask.shape[1] <= nat.PARAXIAL_MAX_SLOTS:
        # host-side masks (no device sync), then the padded solve as one kernel
        n_surf_h = structures.mask.sum(axis=1)
        given_h = structures.mask.copy()
        given_h[np.arange(B), n_surf_h - 1] = False
        c2d = mask_replace(given_h, torch.zeros(mask.shape, dtype=torch.float32, device=device), c)
        t2d = mask_replace(structures.mask, torch.zeros(mask.shape, dtype=torch.float32, device=device), t)
        n2d = mask_replace(structures.mask_G, torch.ones(mask.shape, dtype=torch.float32, device=device), nd)
        return compute_last_curvature_padded(structures, c2d, t2d, n2d)[mask]
    rows = torch.arange(B, device=device)
    n_surf = mask.sum(dim=1)
    ends_air_air = ~structures.mask_G_torch[rows, n_surf - 2]
    solve_at = n_surf - 1 - ends_air_air.long()
    given = mask.clone()
    given[rows, n_surf - 1] = False
    c2d = mask_replace(given.cpu().numpy(), torch.zeros(mask.shape, dtype=torch.float32, device=device), c)
    t2d = mask_replace(structures.mask, torch.zeros(mask.shape, dtype=torch.float32, device=device), t)
    n2d = mask_replace(structures.mask_G, torch.ones(mask.shape, dtype=torch.float32, device=device), nd)
    n2d = torch.cat((torch.ones_like(n2d[:, 0:1]), n2d), dim=1)
    ahead = given.clone()
    ahead[rows, solve_at] = False
    abcd = interface_propagation_abcd(c2d, t2d, n2d)
    eye = torch.eye(2, device=device).expand_as(abcd)
    system = reduce_abcd(torch.where(ahead[..., None, None], abcd, eye))
    n_after = n2d[rows, solve_at]
    solved = -(1 + n_after * system[:, 1, 0]) / (system[:, 0, 0] * (n_after - 1))
    c2d = c2d.clone()
    c2d[rows, solve_at] = solved
    return c2d[mask]


def compute_last_curvature_padded(structures, c2d, t2d, nd2d):
    """:func:`compute_last_curvature` on the padded ``[B, L]`` layout the batched front end works in
    (no compact <-> padded boolean indexing, hence no device sync): ``c2d`` with the solved curvature
    written into its slot (whatever that slot and the ones behind it held is ignored); ``nd2d`` = index
    behind every slot, 1 for air.  CUDA only: ``tl_paraxial_fwd`` / ``tl_paraxial_bwd``."""
    solved, slot = ops.last_curvature(structures, c2d, t2d, nd2d)
    behind = torch.arange(c2d.shape[1], device=c2d.device)[None, :] >= slot[:, None]
    c_given = torch.where(behind, torch.zeros_like(c2d), c2d)       # the reference leaves 0 behind the solved slot
    return torch.where(behind & (behind.cumsum(dim=1) == 1), solved[:, None], c_given)


def compute_magnification(lens):
    """First-order magnification of the system in front of the stop (rt_tf:765-777): height at the
    stop plane per unit height of a ray entering parallel to the axis -- the A element of the ABCD
    product (used by ``ray_aiming_mode='paraxial'``, rtl:138-140)."""
    if lens.structure.mask.shape[1] == 0:
        return torch.ones(len(lens), device=lens.c.device)
    nd = torch.cat((torch.ones_like(lens.nd[:, 0:1]), lens.nd), dim=1)
    system = reduce_abcd(interface_propagation_abcd(lens.c, lens.t, nd))
    return system[:, 0, 0]             # "the magnification corresponds to the A element" (rt_tf:774-775)


# ---------------------------------------------------------------------------
# The hot path
# ---------------------------------------------------------------------------
def trace_skew(x, y, z, cx, cy, c, t, mu, mask, aggregate=False, allow_backward_rays=True,
               arith=None, k=None, a=None, sd=None):
    """Trace rays from the entrance pupil to the image plane (rtl:594-675).

    Inputs broadcast to [B,F,P,W] (c, t, mask: [B,1,1,1,S]; mu: [B,1,1,W,S]).
    Returns ``(x, y, cx, cy, ray_ok, ray_backward)`` at the image plane, each
    [B,F,P,W].  Differentiable w.r.t. x, y, z, cx, cy, c, t and mu.  ``arith``
    selects the arithmetic policy (default: guarded fast path; ``'exact'`` =
    bit-identical to the reference's fp32 evaluation order).

    Extension (no reference behaviour): ``k`` [B,1,1,1,S] conic constants, ``a``
    [B,1,1,1,S,7] even-asphere coefficients a4..a16, ``sd`` [B,1,1,1,S] clear
    semi-diameters.  With any of them the surfaces are intersected by Newton
    iteration, rays outside ``sd`` fail, and a 7th output, the optical path length
    [B,F,P,W], is returned -- differentiable like the other four (see :func:`compute_opd`);
    gradients also flow to ``k`` and ``a``.

    ``aggregate=True`` (rtl:641-657, spherical lenses): also returns ``stacks``, the dict of
    per-surface penalty terms ``z_RELU``, ``theta_norm``, ``theta_prime_norm`` (S-long lists of
    [B,F,P,W] tensors) that ``compute_loss_out`` sums into its penalty Q
    (optics_simulator_lite.py:430-450).  Unlike the reference it accepts un-broadcast pupil grids
    with W > 1 (rtl:653 raises there) and its gradients stay finite when rays fail (the
    reference's turn NaN: it takes the square root of a failed ray's negative cos^2 before
    masking it); failed rays contribute 0.
    """
    return ops.trace(x, y, z, cx, cy, c, t, mu, mask, allow_backward_rays, _arith_code(arith), k, a, sd,
                     aggregate=aggregate)


def compute_rms2d(x, y, ray_ok):
    """Mean over fields of the y-RMS spot radius of lens 0 (rtl:678-702): per field
    the centroid is the mean over *all* rays, deviations are summed over
    surviving rays and divided by P*W.  ``x`` is unused, as in the reference.

    When ``y`` and ``ray_ok`` are the untouched outputs of :func:`trace_skew` /
    ``RayTracer.trace_rays`` the value and its gradient come from the fused spot pass on that
    trace's inputs (``ops.rms_of_trace``): the reference's own call sequence ``trace_rays ->
    compute_rms2d -> backward`` then costs one forward trace plus one fused pass, with no per-ray
    tensor read back or written in backward."""
    return ops.rms_of_trace(y, ray_ok)[0]


def compute_opd(x, y, cx, cy, opl, ray_ok, mu, radius, chief=0):
    """Optical path difference against a reference sphere (extension row A10; the reference has no OPD).

    ``x, y, cx, cy, opl, ray_ok`` are the outputs of :func:`trace_skew` with extension tables ([B,F,P,W]);
    ``mu`` the index ratios it was given ([B,1,1,W,S]: the image-space index is the product of ``1 / mu``);
    ``radius`` (> 0, broadcastable to [B,F,1,W]) the radius of the reference sphere, which is centred on the
    image point of each field's chief ray -- pupil index ``chief`` -- and lies in front of the image plane.
    A ray meets the image plane at P with direction d and the sphere at P + s d, s < 0;
    ``opd = (opl + n s) - (opl_chief - n radius)``.  Rays that are not ok, and fields whose chief ray is not
    ok, get 0.  Differentiable through all five float inputs, i.e. back to c, t, mu, k, a of the trace: a
    handful of elementwise CUDA ops on top of the trace kernels (the optical path itself, and its adjoint,
    are computed inside them: ``TlTraceOut.opl`` / ``TlSeeds.gopl``).  fp32: the optical path is a sum of
    ~10^1 mm carried to ~1e-7 of itself, so an OPD is good to ~1e-5 mm absolute, not to 1e-5 of itself."""
    nat.require_cuda(opl, 'opl')
    n_image = (1.0 / mu).prod(-1)
    pick = slice(chief, chief + 1)
    dx, dy = x - x[:, :, pick], y - y[:, :, pick]
    along = dx * cx + dy * cy
    reach = torch.sqrt(along * along - (dx * dx + dy * dy) + radius * radius)
    out = (opl - opl[:, :, pick]) + n_image * ((radius - along) - reach)
    good = ray_ok & ray_ok[:, :, pick]
    return torch.where(good, out, torch.zeros_like(out))


def compute_psf(x, y, n_bins=(21, 21), increment=None, y_target=None):
    """Spot diagram as a Gaussian soft histogram (the reference's ``compute_psf``, ray_tracing.py:206-270
    of its TensorFlow original; same arguments and return values).

    x, y: [n_lens, n_fields, n_channels, n_rays] image-plane points (the caller concatenates the rays'
    x-mirror images, optics_simulator_lite.py:667-669: only the non-negative half of the x bins is
    evaluated and mirrored).  One grid per (lens, field), centred on ``y_target`` (default: the mean
    of y) in y and on 0 in x; bin pitch ``increment`` (then the extent is ``increment * n_x_bins`` in
    BOTH directions, as rt_tf:224-226 has it), or -- single grid only, like the reference, whose line
    267 does not broadcast otherwise -- fitted to the ray extent.  Returns ``(x_size, y_size, y_target
    [n_grids], kernels [n_grids, n_channels, n_y_bins, n_x_bins] of unit mass, accounted
    [n_lens, n_fields])`` -- the last the fraction of rays inside the grid.  The binning runs as one
    CUDA kernel (``tl_psf_bin``); it is not differentiable."""
    nat.require_cuda(x, 'x')
    nat.require_cuda(y, 'y')
    if x.dim() != 4 or x.shape != y.shape:
        raise ValueError('x and y must be [n_lens, n_fields, n_channels, n_rays]')
    x = x.detach().to(torch.float32)
    y = y.detach().to(torch.float32)
    n_lens, n_fields, n_channels, n_rays = x.shape
    n_grids = n_lens * n_fields
    n_x_bins, n_y_bins = int(n_bins[0]), int(n_bins[1])
    flat_y = y.reshape(n_grids, -1)
    if y_target is None:
        y_target = flat_y.mean(dim=1)                                                  # rt_tf:218
    y_target = torch.as_tensor(y_target, dtype=torch.float32, device=x.device).reshape(n_grids)
    if increment is not None:                                                          # rt_tf:223-226
        x_incr = y_incr = torch.full((n_grids,), float(increment), dtype=torch.float32, device=x.device)
        x_size = y_size = increment * n_x_bins
        x_win = y_win = torch.full((n_grids,), float(increment * n_x_bins), dtype=torch.float32, device=x.device)
    else:                                                                              # rt_tf:228-235
        if n_grids != 1:
            raise ValueError('compute_psf without `increment` handles a single (lens, field) grid, like the '
                             'reference (ray_tracing.py:267 does not broadcast over several)')
        centred = flat_y - y_target[:, None]
        x_size = x_win = x.reshape(n_grids, -1).amax(dim=1)
        y_size = y_win = 2 * torch.maximum(centred.amax(dim=1) - y_target, y_target - centred.amin(dim=1))
        x_incr, y_incr = x_size / n_x_bins, y_size / n_y_bins
    sums, inside = ops.psf_bin(x.reshape(n_grids, n_channels, n_rays), y.reshape(n_grids, n_channels, n_rays),
                               y_target, x_incr, y_incr, x_win, y_win, (n_x_bins, n_y_bins))
    if n_x_bins % 2 == 1:                                                              # rt_tf:257-260
        full = torch.cat((sums[..., 1:].flip(-1), sums), dim=-1)
    else:
        full = torch.cat((sums.flip(-1), sums), dim=-1)
    kernels = (full / full.sum(dim=(-1, -2), keepdim=True)).to(torch.float32)          # rt_tf:263
    accounted = (inside.sum(dim=1) / (n_channels * n_rays)).to(torch.float32).reshape(n_lens, n_fields)
    return x_size, y_size, y_target, kernels, accounted


def compute_rms2d_all(y, ray_ok):
    """:func:`compute_rms2d` for every lens of the batch: (rms [B], rms_field [B,F])."""
    return ops.spot_rms_from_rays(y, ray_ok)


def _arith_code(arith):
    if arith is None or arith == 'guarded' or arith == nat.ARITH_GUARDED:
        return nat.ARITH_GUARDED
    if arith == 'exact' or arith == nat.ARITH_EXACT:
        return nat.ARITH_EXACT
    raise ValueError(f'unknown arithmetic policy {arith!r}')


class RayTracer:
    """Builds the ray set of a (specs, lens) batch and traces it (rtl:26-208).

    Same constructor arguments as the reference.  ``mode`` selects the pupil
    sampler; ``n_ray_aiming_iter`` > 0 enables real ray aiming to the stop.
    """

    def __init__(self, mode='skew_random', n_rays=(8, 8), rel_fields=(0., 0.707, 1.), vig_fn=None,
                 double_precision=False, wavelengths=(656.3, 587.6, 486.1), n_ray_aiming_iter=0,
                 ray_aiming_mode='real', allow_backward_rays=True, default_device='cuda', arith=None):
        self.mode = mode
        self.default_device = default_device
        dev = default_device
        samplers = {
            'skew_random': lambda ref: circle_pseudo_random(ref, *n_rays, device=dev),
            'circular': lambda ref: circle(ref, *n_rays, dev),
            'tee': lambda ref: tee(ref, dev),
            'chief': lambda ref: chief(ref, n_rays, dev),
            'meridional_uniform': lambda ref: meridional_uniform(ref, n_rays, dev),
            'sagittal_uniform': lambda ref: sagittal_uniform(ref, n_rays, dev),
            'skew_outer_edge_uniform': lambda ref: circle_outer_edge_uniform(ref, n_rays, dev),
            'skew_uniform_half_equidistant': lambda ref: skew_uniform_half_equidistant(ref, *n_rays, device=dev),
            'skew_uniform_half_jittered': lambda ref: skew_uniform_half_jittered(ref, *n_rays, device=dev),
            'skew_inner_square_half': lambda ref: skew_inner_square_half(ref, *n_rays, device=dev),
        }
        if mode not in samplers:
            raise ValueError(f'Ray tracing mode must be one of {sorted(samplers)}, got {mode!r}')
        if mode in ('skew_random', 'circular', 'skew_uniform_half_equidistant', 'skew_uniform_half_jittered'):
            assert len(n_rays) == 2
        self.pupil_span = samplers[mode]
        self.n_rays = n_rays
        self.rel_fields = rel_fields
        self.vig_fn = vig_fn
        self.n_ray_aiming_iter = n_ray_aiming_iter
        self.ray_aiming_mode = ray_aiming_mode
        self.allow_backward_rays = allow_backward_rays
        # (the reference's consumer passes a device tensor, optical_loss.py:83: one read at construction)
        self.wavelengths = [WAVELENGTH_ALIASES[w] if isinstance(w, str) else float(w) for w in wavelengths]
        if double_precision:
            raise NotImplementedError('the CUDA ray-trace kernels compute in fp32 only')
        self.double_precision = False
        self.arith = arith
        self.device_aiming = True      # ray aiming as one CUDA kernel where it applies (see ray_aiming)
        self._cache = {}

    def _fields(self):
        if 'fields' not in self._cache:
            self._cache['fields'] = torch.tensor(self.rel_fields, dtype=torch.float32,
                                                 device=self.default_device)
        return self._cache['fields']

    def _pupil(self, z):
        """Relative pupil coordinates of the configured sampler.  Deterministic grids
        are built once per tracer (they depend on nothing but the constructor
        arguments); the random sampler draws afresh on every call like the reference."""
        if self.mode == 'skew_random':
            return self.pupil_span(z)
        if 'pupil' not in self._cache:
            self._cache['pupil'] = self.pupil_span(z)
        return self._cache['pupil']

    # -- ray-set construction (rtl:80-124) ---------------------------------
    def _ray_set(self, specs, lens, use_vig=True, xy=None, up_to_stop=False):
        dev = self.default_device
        n = lens.get_refractive_indices(self.wavelengths)                 # [B,S,W]
        n = torch.cat((torch.ones_like(n[:, 0:1, :]), n), dim=1)          # air in front
        n = n.transpose(1, 2).reshape(n.shape[0], 1, 1, n.shape[2], -1)   # [B,1,1,W,S+1]
        z = compute_pupil_position(lens).reshape(-1, 1, 1, 1)
        xp_rel, yp_rel = self._pupil(z) if xy is None else xy
        fields = self._fields()
        if use_vig and self.vig_fn is not None and self.mode != 'chief':
            xp_rel, yp_rel = self._vignette(specs, fields, xp_rel, yp_rel)
        if self.n_ray_aiming_iter > 0 and not up_to_stop:
            aim = self.ray_aiming(specs, lens.detach(), use_vig)
            xp_rel, yp_rel = (torch.clamp(v, -2, 2).to(dev).detach() for v in aim(xp_rel, yp_rel))
        xp = scale_to_epd(xp_rel, specs.epd)
        yp = scale_to_epd(yp_rel, specs.epd)
        cy = torch.sin(specs.hfov[:, None] * fields[None, :])[..., None, None]   # [B,F,1,1]
        if 'cx' not in self._cache:
            self._cache['cx'] = torch.zeros((1, 1, 1, 1), device=dev)
        cx = self._cache['cx']
        c = lens.c.reshape(lens.c.shape[0], 1, 1, 1, -1)
        t = lens.t.reshape(lens.t.shape[0], 1, 1, 1, -1)
        mu = n[..., :-1] / n[..., 1:]
        mask = lens.structure.mask_torch.reshape(lens.c.shape[0], 1, 1, 1, -1)
        return xp, yp, z, cx, cy, c, t, mu, mask

    def _vignette(self, specs, fields, xp_rel, yp_rel):
        fields = fields[None, :]
        vig_up = self.vig_fn(fields, specs.vig_up)
        vig_down = self.vig_fn(fields, specs.vig_down)
        vig_x = self.vig_fn(fields, specs.vig_x)
        return apply_vignetting(xp_rel, vig_x, vig_x), apply_vignetting(yp_rel, vig_up, vig_down)

    @staticmethod
    def _extension_tables(lens):
        """k / a / sd of a lens as [B,1,1,1,S(,7)] views (None where the lens has none)."""
        def view(v, tail=()):
            return None if v is None else v.reshape(v.shape[0], 1, 1, 1, v.shape[1], *tail)
        a = getattr(lens, 'a', None)
        return dict(k=view(getattr(lens, 'k', None)), a=view(a, (a.shape[-1],)) if a is not None else None,
                    sd=view(getattr(lens, 'sd', None)))

    def trace_rays(self, specs, lens, use_vig=True, aggregate=False, xy=None, up_to_stop=False, staged=True):
        """Trace the configured ray set; returns what :func:`trace_skew` returns.

        ``staged`` (default): where the staging kernels can build the ray set (see :meth:`spot_rms`)
        the whole call is ONE autograd node over the lens tensors -- staging kernel (+ ray-aiming
        kernel) + trace kernel, three launches instead of the ~60 eager tensor ops of rtl:80-124 --
        and its outputs remember where they came from, so that ``compute_rms2d`` on them runs the
        fused pass (:func:`compute_rms2d`)."""
        can_stage, aimed = self._staging(lens, use_vig)
        if (staged and can_stage and not aggregate and xy is None and not up_to_stop
                and lens.c.is_cuda and lens.c.shape[1] <= nat.MAX_SURFACES_FWD and lens.c.shape[1] <= 64):
            x_rel, y_rel = self._pupil(None)
            out = ops.lens_trace(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd, x_rel, y_rel,
                                 self._tables(lens), self.allow_backward_rays, _arith_code(self.arith), aimed,
                                 **self._staged_kwargs(specs, lens, use_vig))
            self._remember(out, specs, lens, use_vig)
            return out
        args = self._ray_set(specs, lens, use_vig, xy, up_to_stop)
        return trace_skew(*args, aggregate, self.allow_backward_rays, arith=self.arith,
                          **self._extension_tables(lens))

    def _remember(self, out, specs, lens, use_vig):
        """Provenance of a staged trace (ops.rms_of_trace): the fused pass of the same (specs, lens)."""
        import weakref
        tensors = [lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd]
        ray_set = getattr(out[1], '_tl_staged', None)      # (not `out` itself: no reference cycle through y)

        def fused():
            if torch.is_grad_enabled() and any(v.requires_grad for v in tensors[:4]) \
                    and lens.c.shape[1] > nat.MAX_SURFACES_SPOT:
                return None
            return self.spot_rms(specs, lens, use_vig, _ray_set=ray_set)[0]
        out[1]._tl_prov = {'kind': 'lens', 'fused': fused, 'tensors': tensors,
                           'versions': [v._version for v in tensors], 'ok': weakref.ref(out[4]),
                           'ok_version': out[4]._version, 'y_version': out[1]._version}

    def penalty(self, specs, lens, use_vig=True, shard=(0, 1), group=None, n_seq=None, staged=True, _ray_set=None):
        """The ray-angle / ray-path penalty ``sum(Q)`` that ``compute_loss_out`` adds to the RMS
        (optics_simulator_lite.py:430-450), for every lens [B], fused: equal to summing
        ``(sum_k theta_norm + sum_k theta_prime_norm + sum_k z_RELU) / n_seq`` over the stacks of
        ``trace_rays(..., aggregate=True)`` but without materialising them.  ``n_seq`` defaults to
        the number of surfaces of the (first) lens' sequence, as in the reference (osl:441).
        ``staged`` as in :meth:`spot_rms`: the ray set comes from the staging kernel."""
        if any(v is not None for v in self._extension_tables(lens).values()):
            raise ValueError('the penalty terms exist for spherical lenses only')
        if n_seq is None:
            n_seq = int(np.asarray(lens.structure.mask)[0].sum())      # host-side constant: no sync
        can_stage, aimed = self._staging(lens, use_vig)
        if staged and can_stage and lens.c.shape[1] <= nat.MAX_SURFACES_BWD:
            x_rel, y_rel = self._pupil(None)
            return ops.lens_penalty(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd, x_rel, y_rel,
                                    self._tables(lens), n_seq, self.allow_backward_rays, _arith_code(self.arith), shard,
                                    group, aimed=aimed, staged=_ray_set, **self._staged_kwargs(specs, lens, use_vig))
        args = self._ray_set(specs, lens, use_vig)
        return ops.penalty_sum(*args, n_seq, self.allow_backward_rays, _arith_code(self.arith), shard, group)

    def loss_unsup(self, specs, lens, penalty_rate=0.2, use_vig=True, shard=(0, 1), group=None, n_seq=None):
        """``compute_loss_out`` (optics_simulator_lite.py:430-450) for EVERY lens of the batch as two
        fused passes over one staged ray set: ``{'loss_unsup': rms + penalty_rate * penalty, 'rms':
        rms, 'penalty': penalty}``, each a [B] tensor.  This is the batched form of the reference's
        per-sample loop ``Optical_Loss.optical_loss_unsupervised`` (optical_loss.py:99-122), which
        builds a RaytracedOptics per sample and traces 8 fields x 64 pupil points x 3 wavelengths one
        lens at a time (the reference evaluates lens 0's RMS and sums the penalty over its one-lens
        batch; a one-lens batch here gives its numbers)."""
        ray_set = None
        can_stage, aimed = self._staging(lens, use_vig)
        if can_stage and lens.c.is_cuda and lens.c.shape[1] <= nat.MAX_SURFACES_SPOT:
            # ONE staging launch (index model, pupil position, field cosines, ray aiming, reference heights) feeds
            # both fused passes
            x_rel, y_rel = self._pupil(None)
            ray_set = ops.stage_lens(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd, x_rel, y_rel,
                                     self._tables(lens), self.allow_backward_rays, _arith_code(self.arith), aimed,
                                     **self._staged_kwargs(specs, lens, use_vig))
        rms, _ = self.spot_rms(specs, lens, use_vig, shard, group, _ray_set=ray_set)
        pen = self.penalty(specs, lens, use_vig, shard, group, n_seq, _ray_set=ray_set)
        return {'loss_unsup': rms + penalty_rate * pen, 'rms': rms, 'penalty': pen}

    def spot_rms(self, specs, lens, use_vig=True, shard=(0, 1), group=None, staged=True, _ray_set=None):
        """RMS spot size of every lens -- ``compute_rms2d(*trace_rays(...))`` fused
        into one pass that also produces the gradients w.r.t. the lens
        (no [B,F,P,W] tensor is ever materialised).  Returns (rms [B], rms_field [B,F]).
        With ``shard=(rank, world)`` the pupil axis is split over ranks.

        ``staged`` (default): the ray set (index model, pupil position, field cosines) and its
        chain rule also run as two small CUDA kernels; it applies when there is no pupil
        vignetting function, a deterministic pupil sampler and either no ray aiming or the aiming
        the device kernel covers (one iteration, 'real' stop radius: tl_aim, applied on load inside
        the trace kernels); otherwise the torch front end of :meth:`trace_rays` feeds the fused pass."""
        can_stage, aimed = self._staging(lens, use_vig)
        ext = self._extension_tables(lens)
        if staged and can_stage:
            x_rel, y_rel = self._pupil(None)
            return ops.lens_spot_rms(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd, x_rel,
                                     y_rel, self._tables(lens), self.allow_backward_rays,
                                     _arith_code(self.arith), shard, group, aimed=aimed, staged=_ray_set,
                                     **self._staged_kwargs(specs, lens, use_vig))
        args = self._ray_set(specs, lens, use_vig)
        return ops.spot_rms(*args, self.allow_backward_rays, _arith_code(self.arith), shard, group, **ext)

    def _staging(self, lens, use_vig=True):
        """(can the staged kernels build this ray set?, with the ray-aiming kernel?)"""
        # ray aiming stays on the staged path when the device kernel covers it (one iteration, 'real' or
        # 'paraxial' stop radius); a lens batch whose stops are all in front needs none (rtl:131-133)
        stops_in_front = bool((np.asarray(lens.structure.stop_idx) == 0).all())
        aimed = (self.n_ray_aiming_iter == 1 and self.ray_aiming_mode in ('real', 'paraxial') and self.device_aiming
                 and not stops_in_front)
        plain = self.n_ray_aiming_iter == 0 or aimed or stops_in_front
        general = any(v is not None for v in self._extension_tables(lens).values())
        vignetted = self.vig_fn is not None and use_vig
        ok = plain and not general and self.mode != 'skew_random' and lens.c.shape[1] <= 64 \
            and not (vignetted and self.mode == 'chief')
        return ok, aimed

    def _staged_kwargs(self, specs, lens, use_vig=True):
        """Pupil vignetting table and stop-radius mode for the staged kernels."""
        vig = self._vig_table(specs) if (self.vig_fn is not None and use_vig and self.mode != 'chief') else None
        mode = nat.AIM_PARAXIAL if self.ray_aiming_mode == 'paraxial' else nat.AIM_REAL
        return dict(vig=vig, aim_mode=mode)

    def _vig_table(self, specs):
        """[B,F,3] = (x_scale, y_scale, y_offset) of apply_vignetting (rt_tf:479-490) with the factors
        the user's vignetting function gives per (lens, field) (rtl:98-104) -- the form the kernels
        apply on load, each step individually rounded like the reference's eager ops."""
        fields = self._fields()[None, :]
        vig_up = self.vig_fn(fields, specs.vig_up)
        vig_down = self.vig_fn(fields, specs.vig_down)
        vig_x = self.vig_fn(fields, specs.vig_x)
        B, F = specs.epd.shape[0], fields.shape[1]
        table = torch.stack((1 - (vig_x + vig_x) / 2, 1 - (vig_up + vig_down) / 2, (vig_down - vig_up) / 2), dim=-1)
        return torch.broadcast_to(table.to(torch.float32), (B, F, 3)).to(self.default_device).contiguous()

    def spot_rms_and_grads(self, specs, lens, use_vig=True, shard=(0, 1), group=None, out=None):
        """:meth:`spot_rms` together with the gradients of ``sum(rms)`` w.r.t. the lens tensors, as
        one call outside autograd: ``(rms [B], {'c','t','nd','v': [B,L]})``.  On the staged path this
        is the bare kernel sequence (no autograd node, no elementwise glue) and ``out`` may supply
        the result tensors (see ``ops.lens_spot_rms_and_grads``); otherwise it falls back to
        autograd over :meth:`spot_rms`."""
        can_stage, aimed = self._staging(lens, use_vig)
        if can_stage and lens.c.shape[1] <= nat.MAX_SURFACES_SPOT:
            x_rel, y_rel = self._pupil(None)
            rms, _, grads = ops.lens_spot_rms_and_grads(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd,
                                                        x_rel, y_rel, self._tables(lens), self.allow_backward_rays,
                                                        _arith_code(self.arith), shard, group, aimed, out,
                                                        **self._staged_kwargs(specs, lens, use_vig))
            return rms, grads
        leaves = {k: getattr(lens, k).detach().requires_grad_(True) for k in ('c', 't', 'nd', 'v')}
        from .lens_modeling import Lens
        trial = Lens(lens.structure, leaves['c'], leaves['t'], leaves['nd'], leaves['v'],
                     *(getattr(lens, k, None) for k in ('k', 'a', 'sd')))
        with torch.enable_grad():
            rms, _ = self.spot_rms(specs, trial, use_vig, shard, group)
            grads = torch.autograd.grad(rms.sum(), list(leaves.values()), allow_unused=True)
        grads = {k: (torch.zeros_like(leaves[k]) if g is None else g) for k, g in zip(leaves, grads)}
        if out:
            for name, key in (('gc', 'c'), ('gt', 't'), ('gnd', 'nd'), ('gv', 'v')):
                if out.get(name) is not None:
                    out[name].copy_(grads[key])
            if out.get('rms') is not None:
                out['rms'].copy_(rms.detach())
        return rms.detach(), grads

    def _tables(self, lens):
        # cached ON the structure (content-stamped), not under id(structure) in the tracer: ids are
        # reused once a structure is freed, and `tracer.spot_rms(specs[i], lens[i])` builds a fresh
        # Structure per index
        structure = lens.structure
        key = (tuple(float(f) for f in self.rel_fields), tuple(float(w) for w in self.wavelengths),
               str(self.default_device))
        return structure.device_tables(key, lambda: ops.LensTables(structure, self.rel_fields, self.wavelengths,
                                                                   self.default_device))

    # -- ray aiming (rtl:129-208) ---------------------------------------------
    def ray_aiming(self, specs, lens, use_vig):
        """One affine correction of the relative pupil coordinates per (lens, field,
        wavelength) so that the 'tee' rays land where they should on the stop.
        Returns a function (xp_rel, yp_rel) -> corrected (xp_rel, yp_rel)."""
        if (lens.structure.stop_idx == 0).all():
            return lambda xp_rel, yp_rel: (xp_rel, yp_rel)
        dev = self.default_device
        on_device = (self.device_aiming and self.ray_aiming_mode in ('real', 'paraxial') and self.n_ray_aiming_iter == 1
                     and lens.c.is_cuda and lens.c.shape[1] <= 64
                     and getattr(lens, 'k', None) is None and getattr(lens, 'a', None) is None
                     and getattr(lens, 'sd', None) is None)
        if on_device:
            # one kernel (tl_aim) instead of three nested traces and two backward calls
            gains = ops.aim_table(lens.c, lens.t, lens.nd, lens.v, specs.hfov, specs.epd, self._tables(lens),
                                  self.allow_backward_rays, **self._staged_kwargs(specs, lens, use_vig))   # [B,F,W,3]
            x_gain, y_gain, y_shift = (gains[..., j].unsqueeze(2) for j in range(3))   # [B,F,1,W]
            return lambda xp_rel, yp_rel: (xp_rel * x_gain, yp_rel * y_gain + y_shift)
        specs2stop = specs.up_to_stop()
        lens2stop = lens.up_to_stop()
        if self.ray_aiming_mode == 'paraxial':
            stop_radius = compute_magnification(lens2stop) * specs2stop.epd / 2
        elif self.ray_aiming_mode == 'real':
            stop_radius = compute_pupil_radius(specs2stop, lens2stop, default_device=dev)
        else:
            raise ValueError(self.ray_aiming_mode)
        stop_radius = stop_radius.reshape(-1, 1, 1, 1)

        x_tee, y_tee = tee(None, dev)
        shape = (len(lens), len(self.rel_fields), x_tee.shape[2], len(self.wavelengths))
        x_tee = x_tee.expand(shape).contiguous()
        y_tee = y_tee.expand(shape).contiguous()
        if use_vig and self.vig_fn:
            x_tee, y_tee = self._vignette(specs, self._fields(), x_tee, y_tee)
        x_target, y_target = x_tee.clone(), y_tee.clone()

        correct = None
        for _ in range(self.n_ray_aiming_iter):
            if correct is not None:
                x_tee, y_tee = correct(x_tee, y_tee)
            x_tee = x_tee.detach().requires_grad_(True)
            y_tee = y_tee.detach().requires_grad_(True)
            with torch.enable_grad():
                xs, ys, *_ = self.trace_rays(specs2stop, lens2stop, up_to_stop=True, use_vig=False,
                                             xy=(x_tee, y_tee))
                xs_rel = xs / stop_radius
                ys_rel = ys / stop_radius
            # d(stop)/d(pupil) summed over outputs, as the reference's two backward calls do
            slope_x, slope_y = torch.autograd.grad(
                [xs_rel, ys_rel], [x_tee, y_tee],
                [torch.ones_like(xs_rel), torch.ones_like(ys_rel)])
            step_x = -(xs_rel.detach() - x_target) / slope_x
            step_y = -(ys_rel.detach() - y_target) / slope_y
            step_x = torch.where(torch.isfinite(step_x), step_x, torch.zeros_like(step_x))
            step_y = torch.where(torch.isfinite(step_y), step_y, torch.zeros_like(step_y))
            x_tee, y_tee = x_tee.detach(), y_tee.detach()
            # affine map through the sagittal ray (x) and the two meridional rays (y)
            x_sag, dx_sag = x_tee[..., -1:, :], step_x[..., -1:, :]
            y_lo, y_hi = y_tee[..., 0:1, :], y_tee[..., 1:2, :]
            dy_lo, dy_hi = step_y[..., 0:1, :], step_y[..., 1:2, :]
            x_gain = (x_sag + dx_sag) / x_sag
            y_gain = (y_hi + dy_hi - (y_lo + dy_lo)) / (y_hi - y_lo)
            y_shift = (y_lo * dy_hi - y_hi * dy_lo) / (y_lo - y_hi)

            def correct(xp_rel, yp_rel, x_gain=x_gain, y_gain=y_gain, y_shift=y_shift):
                return xp_rel * x_gain, yp_rel * y_gain + y_shift
        return correct


def compute_pupil_radius(specs, lens2stop, default_device='cuda'):
    """Stop radius = height of the on-axis marginal ray at the stop (rtl:834-844)."""
    x = torch.zeros((1, 1, 1, 1), device=default_device)
    y = torch.ones((1, 1, 1, 1), device=default_device)
    tracer = RayTracer(mode='tee', rel_fields=[0.], vig_fn=None, wavelengths=['d'],
                       default_device=default_device)
    _, yp, *_ = tracer.trace_rays(specs, lens2stop, xy=(x, y), use_vig=False)
    return yp.reshape(yp.shape[0])
