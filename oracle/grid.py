"""FourierFlows ``TwoDGrid`` conventions restated in NumPy (oracle; test infrastructure only).

FourierFlows.jl is a third-party, un-vendored dependency of the reference (no
Project.toml/Manifest.toml => no pinned version).  What is restated here is what the
reference's call sites rely on (``rsw/RotatingShallowWater.jl:87,141``;
``utils/IFMAB3.jl:81``; ``raytracing/RaytracingDriver.jl:132-154``) and is pinned by
the recorded values K3/K4 (``Krsq`` extrema through ``|D|``), K5 (``parsevalsum2``)
and the grid printout of ``Notebooks/FFTInterpTest.ipynb`` (x in [-pi, pi-dx], kr = 0..n/2).

Index convention: arrays are indexed like the Julia ones, ``f[i, j]`` with ``i`` the
x / kr index and ``j`` the y / l index.  Memory order is irrelevant to the oracle; the
C-ABI boundary converts to column-major explicitly.
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.fft as sfft

_WORKERS = int(os.environ.get("SWRT_ORACLE_WORKERS", os.cpu_count() or 1))


def alias_ranges(n: int, nkr: int, aliased_fraction: float):
    """``getaliasedwavenumbers`` of FourierFlows (recalled).

    Returns 0-based half-open ranges ``(l_lo, l_hi), (kr_lo, kr_hi)`` of the modes that
    ``dealias!`` zeroes.  Float arithmetic is done exactly as in Julia so that the
    floor/ceil land on the same integers (512 -> 171:342 / 171:257 one-based).
    """
    if not aliased_fraction < 1:
        raise ValueError("`aliased_fraction` must be less than 1")
    if aliased_fraction > 0:
        L = (1 - aliased_fraction) / 2
        R = (1 + aliased_fraction) / 2
        iL = math.floor(L * n) + 1          # one-based, inclusive
        iR = math.ceil(R * n)               # one-based, inclusive
        return (iL - 1, iR), (iL - 1, nkr)
    # aliased_fraction == 0: only the Nyquist row / column
    return (n // 2, n // 2 + 1), (nkr - 1, nkr)


class TwoDGrid:
    """Doubly periodic grid; see SURVEY App. A.1."""

    def __init__(self, nx, Lx=2 * np.pi, ny=None, Ly=None, aliased_fraction=1 / 3, x0=None, y0=None):
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        self.nx, self.ny, self.Lx, self.Ly = int(nx), int(ny), float(Lx), float(Ly)
        self.dx, self.dy = self.Lx / nx, self.Ly / ny
        self.nk, self.nl, self.nkr = nx, ny, nx // 2 + 1
        x0 = -self.Lx / 2 if x0 is None else x0
        y0 = -self.Ly / 2 if y0 is None else y0
        self.x = x0 + self.dx * np.arange(nx)
        self.y = y0 + self.dy * np.arange(ny)
        self.kr = (2 * np.pi / self.Lx) * np.arange(self.nkr, dtype=np.float64).reshape(-1, 1)
        self.l = (2 * np.pi / self.Ly) * (np.fft.fftfreq(ny) * ny).reshape(1, -1)
        self.Krsq = self.kr ** 2 + self.l ** 2
        with np.errstate(divide="ignore"):
            self.invKrsq = 1.0 / self.Krsq
        self.invKrsq[0, 0] = 0.0
        self.aliased_fraction = aliased_fraction
        (self.l_alias, self.kr_alias) = alias_ranges(ny, self.nkr, aliased_fraction)
        # x-direction truncation is computed from nx (kralias = iL(nx):nkr)
        (_, self.kr_alias) = alias_ranges(nx, self.nkr, aliased_fraction)

    # -- transforms: mul!(fh, rfftplan, f) and ldiv!(f, rfftplan, fh) ---------------
    def rfft2(self, f):
        """Unnormalised forward transform, real axis = x (first index)."""
        return sfft.rfft2(f, axes=(1, 0), workers=_WORKERS)

    def irfft2(self, fh):
        """Backward transform scaled by 1/(nx ny); c2r along x is done last, like FFTW/cuFFT."""
        return sfft.irfft2(fh, s=(self.ny, self.nx), axes=(1, 0), workers=_WORKERS)

    # -- dealias!(fh, grid) ---------------------------------------------------------
    def dealias(self, fh):
        """In-place square truncation; ``fh`` is (nkr, nl[, ...])."""
        fh[self.kr_alias[0]:self.kr_alias[1], ...] = 0
        fh[:, self.l_alias[0]:self.l_alias[1], ...] = 0
        return fh

    def dealias_mask(self):
        m = np.ones((self.nkr, self.nl), dtype=bool)
        m[self.kr_alias[0]:self.kr_alias[1], :] = False
        m[:, self.l_alias[0]:self.l_alias[1]] = False
        return m


def parsevalsum2(uh, grid: TwoDGrid):
    """``FourierFlows.parsevalsum2`` for an rfft-shaped array (copy at
    ``thomasyamada/ThomasYamada.jl:319-331``)."""
    a = np.abs(uh) ** 2
    U = a[0, :].sum() + a[grid.nkr - 1, :].sum() + 2 * a[1:grid.nkr - 1, :].sum()
    norm = grid.Lx * grid.Ly / (grid.nx ** 2 * grid.ny ** 2)
    return float(norm * U)


def parsevalsum(uh, grid: TwoDGrid):
    """``FourierFlows.parsevalsum``: same weights applied to ``uh`` itself (real part)."""
    U = uh[0, :].sum() + uh[grid.nkr - 1, :].sum() + 2 * uh[1:grid.nkr - 1, :].sum()
    norm = grid.Lx * grid.Ly / (grid.nx ** 2 * grid.ny ** 2)
    return float(np.real(norm * U))


def makefilter(grid: TwoDGrid, order=4, innerK=2 / 3, outerK=1.0, tol=1e-15):
    """``FourierFlows.makefilter`` (recalled) -- PARITY UNPINNED.

    No recorded reference value exercises the filter; it is restated from the published
    formula: exp(-decay (K-innerK)^order) on non-dimensional K = sqrt((kr dx/pi)^2+(l dy/pi)^2).
    """
    K = np.sqrt((grid.kr * grid.dx / np.pi) ** 2 + (grid.l * grid.dy / np.pi) ** 2)
    decay = -np.log(tol) / (outerK - innerK) ** order
    filt = np.exp(-decay * np.clip(K - innerK, 0, None) ** order)
    filt[K < innerK] = 1.0
    return filt
