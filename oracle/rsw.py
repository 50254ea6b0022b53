"""Rotating-shallow-water family restated in NumPy (oracle; test infrastructure only).

Follows, statement by statement:
  rsw/RotatingShallowWater.jl   calcN! :140-230, populate_L! :262-274, updatevars! :101-116,
                                enforce_reality_condition! :118-133, energies :323-336
  rsw/ModifiedShallowWater.jl   extra pressure term :209-226, L :282-284
  rsw/LinborgShallowWater.jl    rotational advecting velocity :155-156, eta equation :223-237
  rsw/RSWRaytracingDriver.jl    get_streamfunction! :63-67, set_initial_condition! :15-54
State layout: sol[nkr, nl, 3] = (uh, vh, etah), complex128.
"""
from __future__ import annotations

import numpy as np

from .grid import TwoDGrid, parsevalsum2

RSW, MODIFIED, LINDBORG, QUADHEIGHT = "rsw", "modified", "lindborg", "quadheight"


class Params:
    def __init__(self, nu, nnu, f, Cg):
        self.nu, self.nnu, self.f, self.Cg2 = float(nu), int(nnu), float(f), float(Cg) ** 2


def populate_L(grid: TwoDGrid, p: Params, variant=RSW):
    """L[nkr, nl, 3, 3]; rsw/RotatingShallowWater.jl:262-274 (CPU method). L is NOT dealiased."""
    D = -p.nu * grid.Krsq ** p.nnu
    k = np.broadcast_to(grid.kr, D.shape)
    l = np.broadcast_to(grid.l, D.shape)
    L = np.zeros(D.shape + (3, 3), dtype=np.complex128)
    L[..., 0, 0] = D
    L[..., 0, 1] = p.f
    L[..., 1, 0] = -p.f
    L[..., 1, 1] = D
    if variant != QUADHEIGHT:                 # QuadHeightModifiedShallowWater.jl:280-281: no longer linear in m
        L[..., 2, 0] = -1j * k
        L[..., 2, 1] = -1j * l
    L[..., 2, 2] = D
    if variant in (RSW, LINDBORG):
        L[..., 0, 2] = -1j * k * p.Cg2
        L[..., 1, 2] = -1j * l * p.Cg2
    # MODIFIED: the pressure term is no longer linear in eta (ModifiedShallowWater.jl:268,272)
    return L


def calcN(sol, grid: TwoDGrid, p: Params, variant=RSW, Fh=None):
    """N = calcN!(sol); dealiases ``sol`` IN PLACE first, like the reference (:141).
    `Fh` (nkr, nl): vars.Fh as the user's calcF! left it; addforcing! (:234-240) ends calcN! with `@. N += vars.Fh`, which
    broadcasts the 2-D field over the three components of N.

    Every rfft/irfft of the reference is kept as a separate transform (no linear merging),
    so this is the arithmetic the reference performs, in its order.
    """
    g = grid
    g.dealias(sol)
    uh, vh, eh = sol[:, :, 0], sol[:, :, 1], sol[:, :, 2]
    ik, il = 1j * g.kr, 1j * g.l
    N = np.empty_like(sol)

    if variant == LINDBORG:
        rot = (g.kr * vh - g.l * uh) * g.invKrsq
        a = g.irfft2(-g.l * rot)          # ur
        b = g.irfft2(g.kr * rot)          # vr
    else:
        a = g.irfft2(uh)                  # u
        b = g.irfft2(vh)                  # v

    N[:, :, 0] = -g.rfft2(g.irfft2(ik * uh) * a)        # u ux
    N[:, :, 1] = -g.rfft2(g.irfft2(il * vh) * b)        # v vy
    N[:, :, 0] += -g.rfft2(g.irfft2(il * uh) * b)       # v uy
    N[:, :, 1] += -g.rfft2(g.irfft2(ik * vh) * a)       # u vx

    if variant == LINDBORG:
        N[:, :, 2] = -g.rfft2(g.irfft2(ik * eh) * a)
        N[:, :, 2] += -g.rfft2(g.irfft2(il * eh) * b)
        if Fh is not None:
            N += Fh[:, :, None]
        return N

    eta = g.irfft2(eh)
    if variant in (MODIFIED, QUADHEIGHT):
        # QuadHeight carries m = 1/(1+eta) in the third slot: F = 1.5 - 0.5 m^2 (QuadHeightModifiedShallowWater.jl:219-225)
        Ph = g.rfft2(1.5 - 0.5 / (1 + eta) ** 2 if variant == MODIFIED else 1.5 - 0.5 * eta ** 2)
        N[:, :, 0] += -1j * p.Cg2 * g.kr * Ph
        N[:, :, 1] += -1j * p.Cg2 * g.l * Ph
    N[:, :, 2] = -ik * g.rfft2(a * eta)
    N[:, :, 2] += -il * g.rfft2(b * eta)
    if Fh is not None:
        N += Fh[:, :, None]                  # addforcing! :237
    return N


def updatevars(sol, grid: TwoDGrid, p: Params):
    """rsw/RotatingShallowWater.jl:101-116: dealias sol in place, return (u, v, eta, zeta)."""
    g = grid
    g.dealias(sol)
    uh, vh, eh = sol[:, :, 0], sol[:, :, 1], sol[:, :, 2]
    zh = 1j * g.kr * vh - 1j * g.l * uh - p.f * eh
    return g.irfft2(uh), g.irfft2(vh), g.irfft2(eh), g.irfft2(zh)


def enforce_reality_condition(sol, grid: TwoDGrid, p: Params):
    """:118-133.  NB the reference writes the round-tripped fields into vars.*h, not into sol;
    what the caller observes afterwards is a dealiased sol.  We return the round-tripped copy too."""
    u, v, eta, _ = updatevars(sol, grid, p)
    return np.stack([grid.rfft2(u), grid.rfft2(v), grid.rfft2(eta)], axis=-1)


def kinetic_energy(sol, grid):
    return (parsevalsum2(sol[:, :, 0], grid) + parsevalsum2(sol[:, :, 1], grid)) / (2 * grid.Lx * grid.Ly)


def potential_energy(sol, grid, p: Params):
    return 0.5 * p.Cg2 * parsevalsum2(sol[:, :, 2], grid) / (grid.Lx * grid.Ly)


def get_streamfunction(sol, grid: TwoDGrid, p: Params):
    """rsw/RSWRaytracingDriver.jl:63-67: balanced streamfunction from linear PV."""
    Kd2 = p.f ** 2 / p.Cg2
    psih = 1j * grid.kr * sol[:, :, 1] - 1j * grid.l * sol[:, :, 0] - p.f * sol[:, :, 2]
    return psih / (-(grid.Krsq + Kd2))


def initial_condition(grid: TwoDGrid, p: Params, Kg, ag, Kw, aw, rng: np.random.Generator):
    """rsw/RSWRaytracingDriver.jl:15-54 recipe with OUR random stream (Julia's RNG cannot be
    reproduced; the same arrays feed both sides of every parity check).  Returns sol[nkr,nl,3].
    Normalisation pins K15: max|u_g| = ag and max|u_w| = aw exactly."""
    g = grid
    shape = (g.nkr, g.nl)
    geo = (Kg[0] ** 2 <= g.Krsq) & (g.Krsq <= Kg[1] ** 2)
    wav = (Kw[0] ** 2 <= g.Krsq) & (g.Krsq <= Kw[1] ** 2) & (g.Krsq > 0)
    phase = 2 * np.pi * rng.random(shape)
    sgn = np.sign(rng.random(shape) - 0.5)
    shift = np.exp(1j * phase)
    z = lambda: np.zeros(shape, dtype=np.complex128)
    ugh, vgh, egh, uwh, vwh, ewh = z(), z(), z(), z(), z(), z()
    egh[geo] = (0.5 * shift)[geo]
    ugh[geo] = (-0.5j * p.Cg2 / p.f * g.l * shift)[geo]
    vgh[geo] = (0.5j * p.Cg2 / p.f * g.kr * shift)[geo]
    s = ag / np.abs(g.irfft2(ugh)).max()
    ugh, vgh, egh = ugh * s, vgh * s, egh * s
    wK = sgn * np.sqrt(p.f ** 2 + p.Cg2 * g.Krsq)
    ewh[wav] = (0.5 * shift)[wav]
    uwh[wav] = (g.invKrsq * (0.5 * g.kr * wK * shift + 0.5j * p.f * g.l * shift))[wav]
    vwh[wav] = (g.invKrsq * (0.5 * g.l * wK * shift - 0.5j * p.f * g.kr * shift))[wav]
    s = aw / np.abs(g.irfft2(uwh)).max()
    uwh, vwh, ewh = uwh * s, vwh * s, ewh * s
    return np.stack([ugh + uwh, vgh + vwh, egh + ewh], axis=-1), (ugh, vgh, egh), (uwh, vwh, ewh)


def load_from_snapshot(snapshot, grid: TwoDGrid):
    """rsw/RSWDriver.jl:16-36: spectral zero-pad (or the same slicing when shrinking is not
    supported by the reference) of a (nkr', nl', 3) snapshot onto `grid`, scaled nl^2/nl'^2 (K5)."""
    snkr, snl = snapshot.shape[0], snapshot.shape[1]
    half_nl = snkr - 1
    scale = grid.nl ** 2 / snl ** 2
    new = np.zeros((grid.nkr, grid.nl) + snapshot.shape[2:], dtype=np.complex128)
    new[:snkr, :half_nl] = scale * snapshot[:, :half_nl]
    new[:snkr, grid.nl - half_nl:] = scale * snapshot[:, half_nl:]
    return new


def quadheight_set_solution(u0h, v0h, eta0h, grid: TwoDGrid):
    """QuadHeightModifiedShallowWater.set_solution! :333-347: the third variable is m = 1/(1+eta), through physical space."""
    m0h = grid.rfft2(1.0 / (1.0 + grid.irfft2(eta0h)))
    return np.stack([u0h, v0h, m0h], axis=-1)
