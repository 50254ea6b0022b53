"""Shared builders for parity tests (oracle side = checker only)."""
import numpy as np

from oracle import ifmab3, raytrace, rsw
from oracle.grid import TwoDGrid


def random_state(nx, ny=None, seed=0, amp=0.1, slope=2.0, Lx=2 * np.pi, Ly=None):
    """A smooth random dealiased, Hermitian-consistent RSW state sol[nkr, nl, 3]."""
    g = TwoDGrid(nx, Lx, ny, Ly)
    rng = np.random.default_rng(seed)
    sol = np.empty((g.nkr, g.nl, 3), dtype=np.complex128)
    for v in range(3):
        fh = g.rfft2(rng.standard_normal((g.nx, g.ny)))
        fh *= 1.0 / (1.0 + g.Krsq) ** (slope / 2 + 0.5)
        f = g.irfft2(g.dealias(fh))
        fh = g.rfft2(f * (amp / np.abs(f).max()))
        sol[:, :, v] = g.dealias(fh)
    return g, sol


def config2_setup(nx, seed=1234):
    """BASELINE config 2 recipe (rsw/RSWRaytracingParameters.jl + raytracing/RaytracingDriver.jl:49-85) at any nx."""
    L, f, Cg, nnu, nutune, cfltune = 2 * np.pi, 3.0, 1.0, 4, 1.0, 0.1
    ag, aw, Kg, Kw = 1.5, 0.1, (10, 13), (0, 5)
    dx = L / nx
    kmax = nx / 2 - 1
    dt = cfltune / (ag + aw) * dx
    nu = nutune * 2 * np.pi / nx / (kmax ** (2 * nnu)) / dt
    g = TwoDGrid(nx, L)
    p = rsw.Params(nu, nnu, f, Cg)
    sol, _, _ = rsw.initial_condition(g, p, Kg, ag, Kw, aw, np.random.default_rng(seed))
    g.dealias(sol)
    return g, p, sol, dict(L=L, f=f, Cg=Cg, nnu=nnu, nu=nu, dt=dt, k0=np.sqrt((2 * f) ** 2 - f ** 2) / Cg)


def oracle_steps(g, p, sol, dt, nsteps, variant=rsw.RSW, filt=None):
    ts = ifmab3.IFMAB3(ifmab3.expL_closed_form(g, p, dt, variant), dt, lambda s: rsw.calcN(s, g, p, variant), filt)
    # closed form for exp(L dt) and exp(2 L dt) (cross-checked against the general exponential in K3)
    ts.expLdt = ifmab3.expL_closed_form(g, p, dt, variant)
    ts.exp2Ldt = ifmab3.expL_closed_form(g, p, 2 * dt, variant)
    sol = sol.copy()
    for _ in range(nsteps):
        ts.stepforward(sol)
    return g.dealias(sol)


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
