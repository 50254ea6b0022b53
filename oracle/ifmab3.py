"""Integrating-factor (matrix exponential) AB3 stepper restated in NumPy (oracle only).

Follows utils/IFMAB3.jl: getexpLs :26-30 (general matrix exponential per wavenumber, for
dt and 2dt), IFMAB3TimeStepper :68-88, mvmul! :125-127 (y_a = sum_b A[a,b] x_b, K2),
IFMAB3update! :129-140 (Euler while clock.step < 3, i.e. for the first THREE calls),
stepforward! :157-169 (calcN!, update, filter, clock, history rotation).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

AB3H1, AB3H2, AB3H3 = 23 / 12, 16 / 12, 5 / 12


def getexpLs(L, dt):
    """exp(L dt), exp(2 L dt) with a general matrix exponential (scipy.linalg.expm batches
    over leading axes), as `mapslices(exp, L*dt, dims=(3,4))` does."""
    if L.ndim == 4:
        return scipy.linalg.expm(L * dt), scipy.linalg.expm(L * (2 * dt))
    return np.exp(L * dt), np.exp(2 * L * dt)       # diagonal=true branch :72-74


def expL_closed_form(grid, params, dt, variant="rsw"):
    """Closed form used by the CUDA path: L0 = L - D I satisfies
    L0^3 = -w^2 L0, so exp(L dt) = e^{D dt} [I + sin(w dt)/w L0 + (1-cos(w dt))/w^2 L0^2].
    Kept in the oracle to cross-check against the general exponential (K3)."""
    from .rsw import populate_L

    L = populate_L(grid, params, variant)
    D = L[..., 0, 0].real
    L0 = L.copy()
    for a in range(3):
        L0[..., a, a] = 0
    w2 = params.f ** 2 + (params.Cg2 * grid.Krsq if variant in ("rsw", "lindborg") else 0.0)
    w2 = np.broadcast_to(w2, D.shape)
    w = np.sqrt(w2)
    x = w * dt
    s = np.where(x > 1e-8, np.sin(x) / np.where(w > 0, w, 1), dt)
    c = np.where(x > 1e-4, (1 - np.cos(x)) / np.where(w2 > 0, w2, 1), dt * dt * (0.5 - x * x / 24))
    I = np.eye(3)
    L02 = L0 @ L0   # (QuadHeight: third row of L0 is zero, L0^3 = -f^2 L0 still holds)
    return np.exp(D * dt)[..., None, None] * (I + s[..., None, None] * L0 + c[..., None, None] * L02)


def mvmul(A, x):
    """y[i,j,a] = sum_b A[i,j,a,b] x[i,j,b]   (utils/IFMAB3.jl:125-127; orientation pinned by K2)."""
    if A.ndim == x.ndim:
        return A * x
    return np.einsum("ijab,ijb->ija", A, x)


class IFMAB3:
    """One object = (timestepper, clock).  `calcN(sol)` must dealias `sol` in place."""

    def __init__(self, L, dt, calcN, filt=None):
        self.dt = float(dt)
        self.expLdt, self.exp2Ldt = getexpLs(L, self.dt)
        self.calcN = calcN
        self.filter = filt
        self.Nm1 = None
        self.Nm2 = None
        self.t = 0.0
        self.step = 0

    def stepforward(self, sol):
        N = self.calcN(sol)
        if self.step < 3:
            sol += self.dt * N
            sol[...] = mvmul(self.expLdt, sol)
        else:
            A = mvmul(self.expLdt, self.Nm1)
            B = mvmul(self.exp2Ldt, self.Nm2)
            sol += self.dt * (AB3H1 * N - AB3H2 * A + AB3H3 * B)
            sol[...] = mvmul(self.expLdt, sol)
        if self.filter is not None:
            sol *= self.filter if self.filter.ndim == sol.ndim else self.filter[..., None]
        self.t += self.dt
        self.step += 1
        self.Nm2 = self.Nm1 if self.Nm1 is not None else np.zeros_like(N)
        self.Nm1 = N
        return sol
