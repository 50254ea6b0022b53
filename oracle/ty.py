"""Thomas-Yamada barotropic/baroclinic model and the FourierFlows ETDRK4 / FilteredRK4 steppers, restated in NumPy
(oracle; test infrastructure only).

  thomasyamada/ThomasYamada.jl   calcN! :129-166, calcN_vorticity! :169-205, calcN_baroclinic! :207-249,
                                 calcN_pressure! :251-270, Equation :277-290 (diagonal hyperviscous L, linear f-plane
                                 wave terms live in N :143-146), updatevars! :76-101
  FourierFlows ETDRK4 / RK4      third party, un-vendored: recalled -- PARITY UNPINNED.
State sol[nkr, nl, 4] = (zeta_t, u_c, v_c, p_c).
"""
from __future__ import annotations

import numpy as np


def ty_L(grid, nu, nnu):
    D = -nu * grid.Krsq ** nnu
    return np.repeat(D[:, :, None], 4, axis=2)


def ty_calcN(sol, grid, Ro):
    g = grid
    g.dealias(sol)
    zth, uch, vch, pch = (sol[:, :, i] for i in range(4))
    ik, il = 1j * g.kr, 1j * g.l
    psith = -zth * g.invKrsq
    uth, vth = -il * psith, ik * psith
    N = np.empty_like(sol)
    N[:, :, 0] = 0.0
    N[:, :, 1] = vch - ik * pch
    N[:, :, 2] = -uch - il * pch
    N[:, :, 3] = -ik * uch - il * vch
    zt, ut, vt, uc, vc = g.irfft2(zth), g.irfft2(uth), g.irfft2(vth), g.irfft2(uch), g.irfft2(vch)
    # vorticity equation :169-205
    N[:, :, 0] += -Ro * (il * g.rfft2(vt * zt) + ik * g.rfft2(ut * zt))
    N[:, :, 0] += -Ro * (-g.kr ** 2 + g.l ** 2) * g.rfft2(uc * vc)
    N[:, :, 0] += -Ro * (-g.kr * g.l * g.rfft2(vc * vc) + g.kr * g.l * g.rfft2(uc * uc))
    # baroclinic momentum :207-249
    N[:, :, 1] += -Ro * (ik * g.rfft2(ut * uc))
    N[:, :, 2] += -Ro * (il * g.rfft2(vt * vc))
    N[:, :, 1] += -Ro * (g.rfft2(g.irfft2(il * uch) * vt) + g.rfft2(g.irfft2(il * uth) * vc))
    N[:, :, 2] += -Ro * (g.rfft2(g.irfft2(ik * vch) * ut) + g.rfft2(g.irfft2(ik * vth) * uc))
    # pressure :251-270
    N[:, :, 3] += -Ro * (g.rfft2(g.irfft2(ik * pch) * ut) + g.rfft2(g.irfft2(il * pch) * vt))
    return N


def etdrk4_coeffs(dt, L, ncirc=32, rcirc=1.0):
    """FourierFlows.getetdcoeffs (recalled): contour means on a circle of radius rcirc around dt L."""
    circ = rcirc * np.exp(2j * np.pi / ncirc * (np.arange(ncirc) + 0.5))
    zc = (dt * L)[..., None] + circ
    M = lambda x: np.mean(x, axis=-1)
    zeta = dt * M((np.exp(zc / 2) - 1) / zc)
    alpha = dt * M((-4 - zc + np.exp(zc) * (4 - 3 * zc + zc ** 2)) / zc ** 3)
    beta = dt * M((2 + zc + np.exp(zc) * (-2 + zc)) / zc ** 3)
    gamma = dt * M((-4 - 3 * zc - zc ** 2 + np.exp(zc) * (4 - zc)) / zc ** 3)
    if np.isrealobj(L):
        zeta, alpha, beta, gamma = zeta.real, alpha.real, beta.real, gamma.real
    return zeta, alpha, beta, gamma


class ETDRK4:
    """FourierFlows ETDRK4TimeStepper + stepforward! (Cox-Matthews / Kassam-Trefethen)."""

    def __init__(self, L, dt, calcN, filt=None):
        """`filt`: FilteredETDRK4TimeStepper -- the same step followed by `sol *= filter` (raytracing/TestParameters.jl:6)."""
        self.dt, self.calcN, self.filter = float(dt), calcN, filt
        self.expLdt, self.exphLdt = np.exp(L * dt), np.exp(L * dt / 2)
        self.zeta, self.alpha, self.beta, self.gamma = etdrk4_coeffs(self.dt, L)
        self.t, self.step = 0.0, 0

    def stepforward(self, sol):
        N1 = self.calcN(sol)
        s1 = self.exphLdt * sol + self.zeta * N1
        N2 = self.calcN(s1)
        s2 = self.exphLdt * sol + self.zeta * N2
        N3 = self.calcN(s2)
        s2 = self.exphLdt * s1 + self.zeta * (2 * N3 - N1)
        N4 = self.calcN(s2)
        sol[...] = self.expLdt * sol + self.alpha * N1 + 2 * self.beta * (N2 + N3) + self.gamma * N4
        if self.filter is not None:
            sol *= self.filter
        self.t += self.dt
        self.step += 1
        return sol


class FilteredRK4:
    """FourierFlows (Filtered)RK4: classical RK4 on RHS = calcN!(sol) + L .* sol, then sol *= filter."""

    def __init__(self, L, dt, calcN, filt=None):
        self.L, self.dt, self.calcN, self.filter = L, float(dt), calcN, filt
        self.t, self.step = 0.0, 0

    def _rhs(self, s):
        n = self.calcN(s)          # dealiases s in place first, like the reference
        return n + self.L * s

    def stepforward(self, sol):
        dt = self.dt
        r1 = self._rhs(sol)
        s = sol + dt / 2 * r1
        r2 = self._rhs(s)
        s = sol + dt / 2 * r2
        r3 = self._rhs(s)
        s = sol + dt * r3
        r4 = self._rhs(s)
        sol += dt * (r1 / 6 + r2 / 3 + r3 / 3 + r4 / 6)
        if self.filter is not None:
            sol *= self.filter if self.filter.ndim == sol.ndim else self.filter[..., None]
        self.t += dt
        self.step += 1
        return sol
