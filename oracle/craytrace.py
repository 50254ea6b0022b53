"""ctypes loader of the C restatement of the oracle's RK4 tracer (oracle/c/raytrace_oracle.c) -- test infrastructure only.

`raytrace(xk, sign, t0, t1, F_old, F_new, grid, f, Cg, nsub, lerp, threads)` has the signature and semantics of
`oracle.raytrace.raytrace` (bilinear mode); `available()` says whether the shared object has been built
(`make -C oracle/c`, done by `__graft_entry__.build()`)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libraytrace_oracle.so")
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.oracle_raytrace_rk4.restype = None
        _lib.oracle_raytrace_rk4.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                             C.c_longlong, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_double, C.c_int, C.c_int, C.c_int]
        _lib.oracle_raytrace_threads.restype = C.c_int
        _lib.oracle_raytrace_threads.argtypes = [C.c_int]
    return _lib


def threads_used(threads=0):
    """OpenMP threads the tracer runs on when asked for `threads` (0 = the OpenMP default, i.e. OMP_NUM_THREADS or all cores)."""
    return int(_load().oracle_raytrace_threads(int(threads)))


def raytrace(xk, sign, t0, t1, F_old, F_new, grid, f, Cg, nsub=1, lerp=0, threads=0):
    """In place on a C-contiguous (N, 4) float64 array; `threads` > 0 sets the OpenMP thread count explicitly (0 = OpenMP's
    default, which follows OMP_NUM_THREADS -- torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    assert xk.flags.c_contiguous and xk.dtype == np.float64 and xk.shape[1] == 4
    sign = np.ascontiguousarray(sign, dtype=np.float64)
    Fo, Fn = np.ascontiguousarray(F_old, dtype=np.float64), np.ascontiguousarray(F_new, dtype=np.float64)
    assert Fo.shape == (grid.nx, grid.ny, 5) and Fn.shape == Fo.shape
    _load().oracle_raytrace_rk4(xk.ctypes.data, sign.ctypes.data, xk.shape[0], float(t0), float(t1), Fo.ctypes.data, Fn.ctypes.data,
                                grid.nx, grid.ny, float(grid.x[0]), float(grid.y[0]), float(grid.dx), float(grid.dy), float(f), float(Cg),
                                int(nsub), int(lerp), int(threads))
    return xk
