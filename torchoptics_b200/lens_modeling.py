"""Batched lens prescriptions: ``Structure`` / ``Specs`` / ``Lens`` and the glass model.

Host-side mirror of the reference data model (``/root/reference/torchlens/
lens_modeling.py``): same class names, constructor arguments, attributes and
properties, so code written against ``torchlens.lens_modeling`` keeps working
(Structure lm:151-213, Specs lm:216-252, Lens lm:255-386, glass helpers
lm:29-104).  Everything here is tiny ``[B, Lmax]`` bookkeeping that feeds the
CUDA trace; it is device-agnostic torch/numpy code.

Storage convention (lm:1-12): prescriptions are padded 2-D ``[B, Lmax]`` tensors
-- curvature and thickness padded with 0, refractive index padded with 1 -- plus
boolean masks saying which slots are real surfaces (``mask``) and which of those
are followed by glass (``mask_G``).  The ``flat_*`` views are the compact 1-D
forms used as optimisation variables.

Extensions that the reference does not have (all optional, default = spherical
reference behaviour): per-surface conic constant ``k``, even-asphere
coefficients ``a`` (``[B, Lmax, 7]`` for rho^2 ... rho^8, rho = x^2 + y^2, i.e.
a4 ... a16) and clear semi-diameter ``sd``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

N_ASPHERE_TERMS = 7  # a4, a6, ..., a16

# Fraunhofer lines used by the two-term dispersion model (lm:362-364), in nm.
LINE_C, LINE_D, LINE_F = 656.3, 587.6, 486.1


def mask_replace(mask: np.ndarray, src: torch.Tensor, dst: torch.Tensor):
    """Scatter the 1-D values ``dst`` into ``src`` (in place) where ``mask`` is set.

    Argument naming follows the reference (lm:21-26): ``src`` is the padded 2-D
    tensor that receives the values, ``dst`` holds the compact values."""
    if src.shape != mask.shape:
        raise AssertionError(f'padded tensor {tuple(src.shape)} does not match mask {mask.shape}')
    if src.dtype != dst.dtype or src.device != dst.device:
        raise AssertionError('padded and compact tensors must agree in dtype and device')
    if dst.dim() != 1:
        raise AssertionError('compact values must be 1-D')
    where = torch.from_numpy(np.ascontiguousarray(mask)).to(dst.device)
    return src.masked_scatter_(where, dst)


# --------------------------------------------------------------------------
# Glass model: (nd, v) <-> normalised 2-vector g (lm:29-46)
# --------------------------------------------------------------------------
_GLASS_CENTRE = (1.6426209211349487, 48.8505973815918)
_NV_TO_G = ((-7.497527849096219, -7.49752916467739),
            (0.07842101471405442, -0.07842100095362642))
_G_TO_NV = ((-0.06668863644654068, 6.3758429552417315),
            (-0.0666886481483064, -6.375841836481304))


_GLASS_CONSTANTS = {}


def _glass_constant(like: torch.Tensor, name: str):
    """The glass-map constants as tensors of ``like``'s device and dtype, built once per (device,
    dtype): a fresh ``new_tensor`` per call is a pageable host -> device copy, which a CUDA graph
    capture of the caller refuses."""
    key = (name, like.device, like.dtype)
    if key not in _GLASS_CONSTANTS:
        value = {'centre': [_GLASS_CENTRE], 'nv_to_g': _NV_TO_G, 'g_to_nv': _G_TO_NV}[name]
        _GLASS_CONSTANTS[key] = torch.tensor(value, dtype=like.dtype, device=like.device)
    return _GLASS_CONSTANTS[key]


def g_from_n_v(n: torch.Tensor, v: torch.Tensor):
    """(nd, Abbe) -> whitened glass coordinates ``g`` of shape [N, 2] (lm:29-38)."""
    assert n.dim() == 1 and v.dim() == 1
    assert n.device == v.device and n.dtype == v.dtype
    centred = torch.stack((n, v), dim=-1) - _glass_constant(n, 'centre')
    return centred @ _glass_constant(n, 'nv_to_g')


def n_v_from_g(g: torch.Tensor):
    """Inverse of :func:`g_from_n_v`; returns the tuple (nd, v) (lm:41-46)."""
    assert g.dim() == 2 and g.shape[1] == 2
    nv = g @ _glass_constant(g, 'g_to_nv') + _glass_constant(g, 'centre')
    return torch.unbind(nv, dim=1)


def map_glass_to_closest(g, catalog_g):
    """Nearest catalogue glass in g-space (lm:101-104).  Returns the reference's
    pair ``(gathered, catalog_g)`` -- including its quirk of gathering along
    dim 0 with a 1-D index."""
    nearest = torch.argmin(torch.norm(g[:, None, :] - catalog_g[None, :, :], dim=-1), dim=1)
    return torch.gather(catalog_g, 0, nearest), catalog_g


def find_valid_curvatures(sequence):
    """Which surfaces may carry a free curvature (lm:49-53): glass-fronted ones,
    or ones directly behind glass (except the last air gap)."""
    behind_glass = np.concatenate(
        (np.zeros_like(sequence.mask_G[:, 0:1]), sequence.mask_G[:, :-1]), axis=1)
    return sequence.mask_G | behind_glass & sequence.mask_except_last & sequence.mask


# --------------------------------------------------------------------------
# Structure
# --------------------------------------------------------------------------
class Structure:
    """Which slots of the padded ``[B, Lmax]`` layout are surfaces / glass.

    Built either from ``sequence`` -- a numpy array of strings such as
    ``np.array(['GAGAAGA'])`` where ``G`` = surface followed by glass and ``A`` =
    followed by air -- or from explicit ``mask`` / ``mask_G`` arrays (lm:152-178).
    """

    def __init__(self, stop_idx, mask: np.ndarray = None, mask_G: np.ndarray = None,
                 sequence=None, default_device='cuda'):
        self.stop_idx = np.asarray(stop_idx)
        assert self.stop_idx.ndim == 1
        if sequence is not None:
            assert mask is None and mask_G is None
            assert isinstance(sequence, np.ndarray)
            letters = sequence.astype(str).view('U1').reshape(sequence.shape[0], -1)
            mask = np.array(letters != '')
            mask_G = np.array(letters == 'G')
        assert mask is not None and mask_G is not None
        assert mask.ndim == 2 and mask_G.ndim == 2
        self.mask = mask
        self.mask_G = mask_G
        self.default_device = default_device
        self.mask_torch = torch.from_numpy(np.ascontiguousarray(mask)).to(default_device)
        self.mask_G_torch = torch.from_numpy(np.ascontiguousarray(mask_G)).to(default_device)

    def __len__(self):
        return self.mask.shape[0]

    def up_to_stop(self):
        """Sub-structure in front of the aperture stop (lm:185-192)."""
        # built once per value of stop_idx (a fresh host->device copy per call could not be
        # graph-captured); an in-place edit of stop_idx rebuilds it, like the reference's per-call build
        key = self.stop_idx.tobytes()
        cached = getattr(self, '_front', None)
        if cached is None or cached[0] != key:
            n_keep = int(self.stop_idx.max())
            before_stop = np.arange(n_keep)[None, :] < self.stop_idx[:, None]
            front = Structure(self.stop_idx.copy(), self.mask[:, :n_keep] & before_stop,
                              self.mask_G[:, :n_keep] & before_stop,
                              default_device=self.default_device)
            cached = self._front = (key, front)
        return cached[1]

    def device_tables(self, key, build):
        """Per-structure cache of device-side constants (the staged kernels' masks / stop indices):
        lives and dies with this object -- never keyed by ``id()``, which CPython reuses -- and is
        rebuilt when the masks or stop indices were edited in place."""
        stamp = (self.stop_idx.tobytes(), self.mask.tobytes(), self.mask_G.tobytes())
        cache = self.__dict__.setdefault('_device_tables', {})
        hit = cache.get(key)
        if hit is None or hit[0] != stamp:
            hit = cache[key] = (stamp, build())
        return hit[1]

    def clone(self):
        return Structure(self.stop_idx.copy(), self.mask.copy(), self.mask_G.copy(),
                         default_device=self.default_device)

    def __getitem__(self, index):
        if isinstance(index, int):
            index = slice(index, index + 1)
        n_keep = self.mask[index].sum(axis=1).max()
        return Structure(self.stop_idx[index], self.mask[index, :n_keep],
                         self.mask_G[index, :n_keep], default_device=self.default_device)

    @property
    def last_g_idx(self):
        slot = np.broadcast_to(np.arange(self.mask.shape[1], dtype=self.stop_idx.dtype),
                               self.mask.shape)
        return np.where(self.mask_G, slot, 0).argmax(axis=1)

    @property
    def mask_except_last(self):
        trimmed = self.mask.copy()
        trimmed[np.arange(len(self)), self.last_g_idx + 1] = 0
        return trimmed


# --------------------------------------------------------------------------
# Specs
# --------------------------------------------------------------------------
@dataclass
class Specs:
    """First-order requirements per lens: entrance-pupil diameter, half field of
    view [rad] and optional pupil-vignetting factors (lm:216-252)."""
    structure: Structure
    epd: torch.Tensor
    hfov: torch.Tensor
    vig_up: torch.Tensor = None
    vig_down: torch.Tensor = None
    vig_x: torch.Tensor = None

    def __post_init__(self):
        assert self.epd.dim() == 1, 'EPD should be 1-dimensional'
        assert self.hfov.dim() == 1, 'HFOV should be 1-dimensional'
        if self.vig_up is None or self.vig_down is None:
            self.vig_up = torch.zeros_like(self.epd)
            self.vig_down = torch.zeros_like(self.epd)
            self.vig_x = torch.zeros_like(self.epd)

    def __len__(self):
        return len(self.structure)

    def scale(self, factor):
        return Specs(self.structure, self.epd * factor, self.hfov,
                     self.vig_up, self.vig_down, self.vig_x)

    def up_to_stop(self):
        return Specs(self.structure.up_to_stop(), self.epd, self.hfov,
                     self.vig_up, self.vig_down, self.vig_x)

    def __getitem__(self, index):
        if isinstance(index, int):
            index = slice(index, index + 1)
        return Specs(self.structure[index], self.epd[index], self.hfov[index],
                     self.vig_up[index], self.vig_down[index], self.vig_x[index])


_WL_CACHE = {}


def _wavelength_tensor(wavelengths, dtype, device):
    """Device tensor of the wavelength list, built once (a host->device copy per call
    would also make the index model impossible to capture in a CUDA graph)."""
    key = (wavelengths, dtype, str(device))
    if key not in _WL_CACHE:
        _WL_CACHE[key] = torch.tensor(wavelengths, dtype=dtype, device=device)
    return _WL_CACHE[key]


# --------------------------------------------------------------------------
# Lens
# --------------------------------------------------------------------------
def _pad_from_flat(structure_mask: np.ndarray, like_mask: torch.Tensor, flat: torch.Tensor,
                   fill: float):
    padded = torch.full(like_mask.shape, fill, dtype=flat.dtype, device=flat.device)
    return mask_replace(structure_mask, padded, flat)


@dataclass
class Lens:
    """Prescription of a batch of lenses (lm:255-386).

    ``c``, ``t``: curvature / thickness after each surface, over ``structure.mask``.
    ``nd``, ``v``: d-line index / Abbe number of the medium after each surface,
    meaningful over ``structure.mask_G``.  Each may be given 2-D padded or 1-D
    compact.  Optional extension fields ``k``, ``a``, ``sd`` are 2-D padded
    (``a``: ``[B, Lmax, 7]``) or ``None``.
    """
    structure: Structure
    c: torch.Tensor
    t: torch.Tensor
    nd: torch.Tensor
    v: torch.Tensor
    k: Optional[torch.Tensor] = None
    a: Optional[torch.Tensor] = None
    sd: Optional[torch.Tensor] = None

    def __post_init__(self):
        s = self.structure
        if self.c.dim() == 1:
            self.c = _pad_from_flat(s.mask, s.mask_torch, self.c, 0.)
        if self.t.dim() == 1:
            self.t = _pad_from_flat(s.mask, s.mask_torch, self.t, 0.)
        if self.nd.dim() == 1:
            self.nd = _pad_from_flat(s.mask_G, s.mask_torch, self.nd, 1.)
        if self.v.dim() == 1:
            self.v = _pad_from_flat(s.mask_G, s.mask_torch, self.v, float('nan'))
        if self.a is not None:
            assert self.a.dim() == 3 and self.a.shape[-1] == N_ASPHERE_TERMS, \
                f'asphere coefficients must be [B, Lmax, {N_ASPHERE_TERMS}] (a4..a16)'

    def __len__(self):
        return len(self.structure)

    @property
    def is_aspheric(self):
        return self.k is not None or self.a is not None

    def scale(self, factor):
        # lengths scale by `factor`; a_(2i) has units length^(1-2i)
        a = None
        if self.a is not None:
            powers = torch.tensor([1 - 2 * i for i in range(2, 2 + N_ASPHERE_TERMS)],
                                  dtype=self.a.dtype, device=self.a.device)
            a = self.a * factor ** powers
        return Lens(self.structure, self.c / factor, self.t * factor, self.nd, self.v,
                    self.k, a, None if self.sd is None else self.sd * factor)

    def up_to_stop(self):
        front = self.structure.up_to_stop()
        n_keep = front.mask.shape[1]

        def cut(field):
            if field is None:
                return None
            return torch.where(front.mask_torch[(...,) + (None,) * (field.dim() - 2)],
                               field[:, :n_keep], torch.zeros_like(field[:, :n_keep]))
        return Lens(front,
                    self.c[:, :n_keep][front.mask_torch],
                    self.t[:, :n_keep][front.mask_torch],
                    self.nd[:, :n_keep][front.mask_G_torch],
                    self.v[:, :n_keep][front.mask_G_torch],
                    cut(self.k), cut(self.a),
                    None if self.sd is None else self.sd[:, :n_keep])

    def __getitem__(self, index):
        if isinstance(index, int):
            index = slice(index, index + 1)
        sub = self.structure[index]
        n_keep = sub.mask.shape[1]

        def cut(field):
            return None if field is None else field[index, :n_keep]
        return Lens(sub, self.c[index, :n_keep], self.t[index, :n_keep],
                    self.nd[index, :n_keep], self.v[index, :n_keep],
                    cut(self.k), cut(self.a), cut(self.sd))

    def detach(self):
        def cut(field):
            return None if field is None else field.detach()
        return Lens(self.structure, self.c.detach(), self.t.detach(), self.nd.detach(),
                    self.v.detach(), cut(self.k), cut(self.a), cut(self.sd))

    # ---- compact <-> padded views (lm:317-353) ---------------------------
    @property
    def flat_c(self):
        return self.c[self.structure.mask_torch]

    @flat_c.setter
    def flat_c(self, values):
        self.c = mask_replace(self.structure.mask, self.c, values)

    @property
    def flat_c_but_last(self):
        keep = self.structure.mask.copy()
        keep[np.arange(len(self)), self.structure.mask.sum(axis=1) - 1] = False
        return self.c[keep]

    @property
    def flat_t(self):
        return self.t[self.structure.mask_torch]

    @flat_t.setter
    def flat_t(self, values):
        self.t = mask_replace(self.structure.mask, self.t, values)

    @property
    def flat_nd(self):
        return self.nd[self.structure.mask_G_torch]

    @flat_nd.setter
    def flat_nd(self, values):
        self.nd = mask_replace(self.structure.mask_G, self.nd, values)

    @property
    def flat_v(self):
        return self.v[self.structure.mask_G_torch]

    @flat_v.setter
    def flat_v(self, values):
        self.v = mask_replace(self.structure.mask_G, self.v, values)

    # ---- dispersion (lm:355-374) ----------------------------------------
    def get_refractive_indices(self, wavelengths):
        """n(lambda) = A + B / lambda^2 with A, B fixed by (nd, v); [B, Lmax, n_wl].

        Air slots give 1.  Glasses with v == 0 are treated as dispersion-free
        (n = nd).  The reference's version of that last fix-up (lm:372-373) only
        broadcasts for a batch of one lens; this one is the same for B = 1 and
        well defined for B > 1."""
        wl = _wavelength_tensor(tuple(float(w) for w in wavelengths), self.nd.dtype, self.nd.device)
        glass = self.structure.mask_G_torch
        # air slots carry v = NaN padding: keep it out of the arithmetic so that the
        # gradients of nd / v are 0 there instead of the reference's 0 * NaN
        v = torch.where(glass, self.v, torch.ones_like(self.v))
        dispersive = v != 0
        v = torch.where(dispersive, v, torch.ones_like(v))
        slope = (self.nd - 1) / (v * (LINE_F ** -2 - LINE_C ** -2))
        offset = self.nd - slope / LINE_D ** 2
        n = offset[..., None] + slope[..., None] / wl[None, None, :] ** 2
        n = torch.where(glass[..., None], n, torch.ones_like(n))
        return torch.where(dispersive[..., None], n, self.nd[..., None].expand_as(n))

    # ---- first-order properties (lm:376-386) ------------------------------
    @property
    def efl(self):
        from . import ray_tracing_lite as rt
        return rt.get_first_order(self)[0]

    @property
    def bfl(self):
        from . import ray_tracing_lite as rt
        return rt.get_first_order(self)[1]

    @property
    def entrance_pupil_position(self):
        from . import ray_tracing_lite as rt
        return rt.compute_pupil_position(self)
