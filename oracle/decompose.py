"""Wave / balanced projections restated in NumPy (oracle only -- test infrastructure, never on the product path).

rsw/RSWUtils.jl: `wave_balanced_decomposition` :9-22, `compute_balanced_wave_bases` :24-49,
`compute_balanced_wave_weights` :51-57, `compute_rsw_fields` :59-64.
thomasyamada/TYUtils.jl: `compute_balanced_basis` :10-19, `compute_wave_bases` :21-38, `decompose_balanced_wave` :40-51.
Pinned on K10 (geostrophic part non-divergent / wave part without linear PV) and K11 (G + W reproduces the baroclinic state).
"""
from __future__ import annotations

import numpy as np


def wave_balanced_decomposition(sol, grid, p):
    """((ugh, vgh, etagh), (uwh, vwh, etawh)) stacked as two (nkr, nl, 3) arrays.  rsw/RSWUtils.jl:9-22."""
    uh, vh, eh = sol[:, :, 0], sol[:, :, 1], sol[:, :, 2]
    Kd2 = p.f ** 2 / p.Cg2
    qh = 1j * grid.kr * vh - 1j * grid.l * uh - p.f * eh
    psih = -qh / (grid.Krsq + Kd2)
    bal = np.stack([-1j * grid.l * psih, 1j * grid.kr * psih, p.f / p.Cg2 * psih], axis=-1)
    return bal, sol - bal


def rsw_bases(grid, p):
    """Phi0, Phi+, Phi- (nkr, nl, 3) in (u, v, Cg eta) coordinates.  rsw/RSWUtils.jl:24-49."""
    Cg = np.sqrt(p.Cg2)
    w = np.sqrt(p.f ** 2 + p.Cg2 * grid.Krsq)
    kr, l = np.broadcast_to(grid.kr, w.shape), np.broadcast_to(grid.l, w.shape)
    s = np.sqrt(grid.invKrsq / 2)
    P0 = np.stack([-1j * l * Cg / w, 1j * kr * Cg / w, -p.f / w + 0j], axis=-1)
    Pp = np.stack([(w * kr + 1j * p.f * l) * s / w, (w * l - 1j * p.f * kr) * s / w, Cg * grid.Krsq * s / w + 0j], axis=-1)
    Pm = np.stack([(-w * kr + 1j * p.f * l) * s / w, (-w * l - 1j * p.f * kr) * s / w, Cg * grid.Krsq * s / w + 0j], axis=-1)
    P0[0, 0] = (0, 0, 1)
    Pp[0, 0] = np.array([1j, 1, 0]) / np.sqrt(2)
    Pm[0, 0] = np.array([-1j, 1, 0]) / np.sqrt(2)
    return P0, Pp, Pm


def rsw_weights(sol, bases, p):
    """c0, c+, c-.  rsw/RSWUtils.jl:51-57."""
    Cg = np.sqrt(p.Cg2)
    X = np.stack([sol[:, :, 0], sol[:, :, 1], Cg * sol[:, :, 2]], axis=-1)
    return tuple((X * np.conj(B)).sum(axis=-1) for B in bases)


def rsw_fields(c, bases, p):
    """Inverse of `rsw_weights`: (uh, vh, etah).  rsw/RSWUtils.jl:59-64."""
    X = sum(ci[:, :, None] * B for ci, B in zip(c, bases))
    return X[:, :, 0], X[:, :, 1], X[:, :, 2] / np.sqrt(p.Cg2)


def ty_bases(grid):
    """thomasyamada/TYUtils.jl:10-38 (non-dimensional f = c = 1)."""
    w = np.sqrt(1 + grid.Krsq)
    kr, l = np.broadcast_to(grid.kr, w.shape), np.broadcast_to(grid.l, w.shape)
    s = np.sqrt(grid.invKrsq / 2)
    P0 = np.stack([1j * l / w, -1j * kr / w, -1 / w + 0j], axis=-1)
    Pp = np.stack([(w * kr + 1j * l) * s / w, (w * l - 1j * kr) * s / w, (w * w - 1) * s / w + 0j], axis=-1)
    Pm = np.stack([(-w * kr + 1j * l) * s / w, (-w * l - 1j * kr) * s / w, (w * w - 1) * s / w + 0j], axis=-1)
    P0[0, 0] = (0, 0, 1)
    Pp[0, 0] = np.array([1j, 1, 0]) / np.sqrt(2)
    Pm[0, 0] = np.array([1j, -1, 0]) / np.sqrt(2)
    return P0, Pp, Pm


def ty_decompose(sol, grid, bases=None):
    """(Gh, Wh) of the baroclinic components sol[:, :, 1:4].  thomasyamada/TYUtils.jl:40-51."""
    P0, Pp, Pm = ty_bases(grid) if bases is None else bases
    b = sol[:, :, 1:4]
    proj = lambda B: (b * np.conj(B)).sum(axis=-1, keepdims=True) * B
    return proj(P0), proj(Pp) + proj(Pm)
