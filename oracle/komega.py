"""k-omega post-processing restated in NumPy (oracle only -- test infrastructure, never on the product path).

thomasyamada/TY_k_omega.jl: `hann` :11-17, series extraction and windowed transforms in `write_fourier_data` :46-110.
rsw/fourier-analysis/mrsw/FourierRSW.jl: `demean` :17-20, `linear_least_squares` :22-31, `detrend` :33-36, `clean_fft` :38-41,
the twelve linear series of `write_fourier_data` :76-160.  `k_idx` is 0-based here (the reference's `k_idx - 1`).
"""
from __future__ import annotations

import numpy as np

from . import decompose as od


def hann(L):
    """Periodic Hann window of length L (TY_k_omega.jl:11-17)."""
    n = np.arange(L + 1)
    return (0.5 * (1 - np.cos(2 * np.pi * n / L)))[:-1]


def demean(data):
    return data - data.sum(axis=0) / data.shape[0]


def linear_least_squares(t, data):
    tsum, t2sum = t.sum(), (t ** 2).sum()
    txsum = (t[:, None] * data).sum(axis=0)
    N = t.shape[0]
    slope = (N * txsum) / (N * t2sum - tsum ** 2)
    return slope, -slope * tsum / N


def detrend(t, data):
    m, b = linear_least_squares(t, demean(data))
    return data - m * t[:, None] - b


def clean_fft(t, data, window):
    return np.fft.fft(window[:, None] * detrend(t, data), axis=0)


def ty_series(sol, grid, k_idx):
    """One frame of the six Thomas-Yamada series (ut, vt, ug, vg, uw, vw), each (nl,).  TY_k_omega.jl:72-86."""
    Gh, Wh = od.ty_decompose(sol, grid)
    ut = (-1j * grid.l * sol[:, :, 0])[k_idx]
    vt = (1j * grid.kr * sol[:, :, 0])[k_idx]
    return np.stack([ut, vt, Gh[k_idx, :, 0], Gh[k_idx, :, 1], Wh[k_idx, :, 0], Wh[k_idx, :, 1]])


def ty_spectra(series):
    """series (T, 6, nl) -> the nine transforms the reference stores: ut, vt, ug, vg, uw, vw, U_balanced, U_wave, U_total."""
    T = series.shape[0]
    w = hann(T)[:, None]
    ut, vt, ug, vg, uw, vw = (series[:, j] for j in range(6))
    out = [np.fft.fft(w * x, axis=0) for x in (ut, vt, ug, vg, uw, vw)]
    out.append(np.fft.fft(w * ((ut + ug) + 1j * (vt + vg)), axis=0))
    out.append(np.fft.fft(w * (uw + 1j * vw), axis=0))
    out.append(np.fft.fft(w * ((uw + ug + ut) + 1j * (vw + vg + vt)), axis=0))
    return out


def rsw_series(sol, grid, p, k_idx):
    """One frame of the twelve linear RSW series: ut, vt, etat, ug, vg, etag, uw, vw, etaw, c0, c+, c-  (mrsw/FourierRSW.jl:118-137)."""
    bal, wav = od.wave_balanced_decomposition(sol, grid, p)
    c = od.rsw_weights(sol, od.rsw_bases(grid, p), p)
    rows = [sol[k_idx, :, j] for j in range(3)] + [bal[k_idx, :, j] for j in range(3)] + [wav[k_idx, :, j] for j in range(3)]
    return np.stack(rows + [ci[k_idx] for ci in c])
