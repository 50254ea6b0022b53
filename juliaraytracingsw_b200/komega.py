"""k-omega spectra accumulated on the device while the flow runs.

Mirrors `write_fourier_data(file_indices, k_idx)` of thomasyamada/TY_k_omega.jl:46-110 and
rsw/fourier-analysis/mrsw/FourierRSW.jl:76-160, which post-process stored snapshots with one task per kr index: here
`append()` takes a frame from the live problem (no snapshot files, no D2H of the state) and `spectrum()` runs the windowed
(and, for RSW, detrended) transform in time on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

SERIES_TY, SERIES_RSW = 0, 1
TY_NAMES = ("ut", "vt", "ug", "vg", "uw", "vw", "U_balanced", "U_wave", "U_total")
RSW_NAMES = ("ut", "vt", "ηt", "ugt", "vgt", "ηgt", "uwt", "vwt", "ηwt", "c0t", "c+t", "c-t")


class KOmega:
    def __init__(self, prob, k_idx, max_frames, kind=None):
        """`k_idx` is the reference's 1-based kr index (`grid.kr[k_idx]`)."""
        self.prob = prob
        self.kind = (SERIES_TY if prob.desc.model == 6 else SERIES_RSW) if kind is None else kind
        self.names = TY_NAMES if self.kind == SERIES_TY else RSW_NAMES
        self.k = prob.grid.kr[k_idx - 1, 0]
        self._h = C.c_void_p()
        check(lib().swrt_series_create(prob._h, self.kind, int(k_idx) - 1, int(max_frames), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().swrt_series_destroy(self._h)
            self._h = None

    __del__ = close

    def append(self):
        check(lib().swrt_series_append(self._h))

    @property
    def nframes(self):
        n = C.c_longlong()
        check(lib().swrt_series_frames(self._h, C.byref(n)))
        return n.value

    @property
    def t(self):
        out = np.empty(self.nframes)
        check(lib().swrt_series_times(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def _fetch(self, fn, which):
        out = np.empty((self.nframes, self.prob.grid.nl), dtype=np.complex128, order="F")
        check(fn(self._h, int(which), out.ctypes.data_as(C.c_void_p)))
        return out

    def series(self, which):
        """Raw time series (nframes, nl): `ut_series` ... of the reference's output file."""
        return self._fetch(lib().swrt_series_get, which)

    def spectrum(self, which):
        """`fft(window .* series, 1)` (Thomas-Yamada) or `clean_fft(t, series, window)` (RSW); `which` indexes `self.names`."""
        if isinstance(which, str):
            which = self.names.index(which)
        return self._fetch(lib().swrt_series_spectrum, which)

    def write(self, out):
        """The keys of `radial_data_k=%03d.jld2` into a dict-like `out`."""
        out["k"], out["t"] = self.k, self.t
        for j, name in enumerate(self.names):
            out[name] = self.spectrum(j)
        return out
