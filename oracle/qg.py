"""Quasi-geostrophic models restated in NumPy (oracle; test infrastructure only).

  swqg/SWQG.jl        single-layer shallow-water QG: inversion :101-107, calcN! :140-170, L :181-191,
                      energies :205-250
  swqg/TwoLayerQG.jl  equal-depth two-layer QG: inversion :92-111, calcN! :152-182, L_kernel! :184-198
                      (the reference's Complex{Float32} temporaries inside L_kernel! are a bug, SURVEY App. B #5;
                      T = Float64 is used throughout here), energies :221-250
State: SWQG sol[nkr, nl] (a trailing axis of length 1 is accepted), TwoLayerQG sol[nkr, nl, 2].
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .grid import TwoDGrid, parsevalsum, parsevalsum2


# ------------------------------------------------------------------------------------ SWQG
def swqg_streamfunction(qh, grid, Kd2):
    return -qh / (grid.Krsq + Kd2)


def swqg_L(grid, nu, nnu):
    return -nu * grid.Krsq ** nnu                       # real, diagonal (swqg/SWQG.jl:181-191)


def swqg_calcN(sol, grid, Kd2):
    """swqg/SWQG.jl:140-170; `sol` is (nkr, nl) and is dealiased in place."""
    g = grid
    g.dealias(sol)
    psih = swqg_streamfunction(sol, g, Kd2)
    q = g.irfft2(sol)
    N = -1j * g.l * g.rfft2(g.irfft2(1j * g.kr * psih) * q)
    N += 1j * g.kr * g.rfft2(g.irfft2(1j * g.l * psih) * q)
    return N


def swqg_energies(sol, grid, Kd2):
    psih = swqg_streamfunction(sol, grid, Kd2)
    ke = parsevalsum2(np.sqrt(grid.Krsq) * psih, grid) / (2 * grid.Lx * grid.Ly)
    pe = Kd2 * parsevalsum2(psih, grid) / (2 * grid.Lx * grid.Ly)
    return ke, pe


# ------------------------------------------------------------------------------------ two-layer QG
def twolayer_streamfunction(qh, grid, F):
    """swqg/TwoLayerQG.jl:101-111."""
    q1, q2 = qh[:, :, 0], qh[:, :, 1]
    psih = np.empty_like(qh)
    psih[:, :, 0] = -(grid.Krsq * q1 + F * (q1 + q2))
    psih[:, :, 1] = -(grid.Krsq * q2 + F * (q1 + q2))
    psih /= (grid.Krsq + 2 * F)[:, :, None]
    psih *= grid.invKrsq[:, :, None]
    return psih


def twolayer_L(grid, F, U, mu, nu, nnu):
    """L[nkr, nl, 2, 2], swqg/TwoLayerQG.jl:184-198 (L[a,b] = psi_terms[a] * Sinv[a,b], then the diagonal)."""
    K2 = grid.Krsq
    k = np.broadcast_to(grid.kr, K2.shape)
    D = -nu * K2 ** nnu
    K2inv = grid.invKrsq
    pv = np.stack([-2j * k * F * U, 2j * k * F * U], axis=-1)
    drag = np.stack([np.zeros_like(K2), mu * K2], axis=-1)
    psi_terms = pv + drag
    Sinv = np.empty(K2.shape + (2, 2))
    Sinv[..., 0, 0] = -K2 - F
    Sinv[..., 0, 1] = -F
    Sinv[..., 1, 0] = -F
    Sinv[..., 1, 1] = -K2 - F
    Sinv = Sinv / (K2 + 2 * F)[..., None, None] * K2inv[..., None, None]
    L = psi_terms[..., :, None] * Sinv
    L[..., 0, 0] += -1j * k * U + D
    L[..., 1, 1] += 1j * k * U + D
    return L


def twolayer_calcN(sol, grid, F):
    """swqg/TwoLayerQG.jl:152-182 (batched over the two layers); dealiases `sol` in place."""
    g = grid
    g.dealias(sol)
    psih = twolayer_streamfunction(sol, g, F)
    N = np.empty_like(sol)
    for j in range(2):
        q = g.irfft2(sol[:, :, j])
        N[:, :, j] = -1j * g.l * g.rfft2(g.irfft2(1j * g.kr * psih[:, :, j]) * q)
        N[:, :, j] += 1j * g.kr * g.rfft2(g.irfft2(1j * g.l * psih[:, :, j]) * q)
    return N


def multilayer2_calcN(sol, grid, F, U1, U2, beta, mu):
    """GeophysicalFlows MultiLayerQG.calcN! for two equal layers (third party, un-vendored; restated from SURVEY App. C --
    PARITY UNPINNED): N_j = -F[(u_j+U_j) Qx] - F[v_j Qy_j] - i k F[(u_j+U_j) q_j] - i l F[v_j q_j], Qx = 0,
    Qy = beta +- F (U1 - U2), and the bottom drag N_2 += mu K^2 psih_2.  Called by raytracing/TwoLayerRaytracing.jl:129-130
    with aliased_fraction = 0 (:174), so the physical-space products alias exactly as written here."""
    g = grid
    g.dealias(sol)
    psih = twolayer_streamfunction(sol, g, F)
    N = np.empty_like(sol)
    Qy = (beta + F * (U1 - U2), beta - F * (U1 - U2))
    for j, Uj in enumerate((U1, U2)):
        u = g.irfft2(-1j * g.l * psih[:, :, j]) + Uj
        v = g.irfft2(1j * g.kr * psih[:, :, j])
        q = g.irfft2(sol[:, :, j])
        N[:, :, j] = -g.rfft2(v * Qy[j]) - 1j * g.kr * g.rfft2(u * q) - 1j * g.l * g.rfft2(v * q)
    N[:, :, 1] += mu * g.Krsq * psih[:, :, 1]
    return N


def twolayer_energies(sol, grid, F):
    psih = twolayer_streamfunction(sol, grid, F)
    A = grid.Lx * grid.Ly
    ke = [parsevalsum(grid.Krsq * np.abs(psih[:, :, j]) ** 2, grid) / A for j in range(2)]
    pe = 1 / (2 * A) * F * parsevalsum(np.abs(psih[:, :, 0] - psih[:, :, 1]) ** 2, grid)
    return ke, pe


def expm2x2_closed_form(L, dt):
    """exp(L dt) for 2x2 blocks: e^s [cosh q I + sinh(q)/q (A - s I)], s = tr A / 2, q^2 = ((a-d)/2)^2 + b c
     -- what the CUDA host code tabulates; cross-checked against scipy.linalg.expm in the tests."""
    A = L * dt
    a, b, c, d = A[..., 0, 0], A[..., 0, 1], A[..., 1, 0], A[..., 1, 1]
    s = (a + d) / 2
    q = np.sqrt(((a - d) / 2) ** 2 + b * c + 0j)
    small = np.abs(q) < 1e-8
    qs = np.where(small, 1, q)
    sh = np.where(small, 1 + q * q / 6, np.sinh(qs) / qs)
    ch = np.cosh(q)
    E = np.empty_like(A)
    es = np.exp(s)
    E[..., 0, 0] = es * (ch + sh * (a - s))
    E[..., 0, 1] = es * sh * b
    E[..., 1, 0] = es * sh * c
    E[..., 1, 1] = es * (ch + sh * (d - s))
    return E


# ------------------------------------------------------------------------------------ FilteredAB3 (FourierFlows)
class FilteredAB3:
    """FourierFlows' FilteredAB3 (recalled) -- PARITY UNPINNED (no recorded reference value):
    RHS = calcN!(sol) + L .* sol; Euler while step < 3, then AB3 on the full RHS; then sol *= filter."""

    def __init__(self, L, dt, calcN, filt):
        self.L, self.dt, self.calcN, self.filter = L, float(dt), calcN, filt
        self.r1 = self.r2 = None
        self.t, self.step = 0.0, 0

    def stepforward(self, sol):
        rhs = self.calcN(sol) + self.L * sol
        if self.step < 3:
            sol += self.dt * rhs
        else:
            sol += self.dt * (23 / 12 * rhs - 16 / 12 * self.r1 + 5 / 12 * self.r2)
        sol *= self.filter
        self.t += self.dt
        self.step += 1
        self.r2 = self.r1 if self.r1 is not None else np.zeros_like(rhs)
        self.r1 = rhs
        return sol
