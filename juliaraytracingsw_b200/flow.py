"""Host-side mirror of the reference's flow API over the C ABI.

Names and argument meaning follow rsw/RotatingShallowWater.jl (`Problem`, `set_solution!`,
`enforce_reality_condition!`, `updatevars!`, `kinetic_energy`, `potential_energy`) and
FourierFlows' `stepforward!(prob, diags, n)`; Julia's `!` is dropped.  Arrays cross the boundary
in the Julia layout: `sol[nkr, nl, nvar]` complex128 and fields `(nx, ny)` float64, i.e. NumPy
arrays indexed `[i, j(, v)]` in Fortran order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FlowDesc, check, lib

MODELS = {"RotatingShallowWater": 0, "ModifiedShallowWater": 1, "LinborgShallowWater": 2, "SWQG": 4,
          "TwoLayerQG": 5, "ThomasYamada": 6}
STEPPERS = {"IFMAB3": 0, "FilteredAB3": 1, "ETDRK4": 2, "FilteredRK4": 3}
FIELD_U, FIELD_V, FIELD_ETA, FIELD_ZETA = 0, 1, 2, 16


class Grid:
    """The subset of FourierFlows' TwoDGrid the reference's drivers read."""

    def __init__(self, nx, ny, Lx, Ly, aliased_fraction):
        self.nx, self.ny, self.Lx, self.Ly = nx, ny, Lx, Ly
        self.nk, self.nl, self.nkr = nx, ny, nx // 2 + 1
        self.dx, self.dy = Lx / nx, Ly / ny
        self.x = -Lx / 2 + self.dx * np.arange(nx)
        self.y = -Ly / 2 + self.dy * np.arange(ny)
        self.kr = (2 * np.pi / Lx) * np.arange(self.nkr, dtype=np.float64).reshape(-1, 1)
        self.l = (2 * np.pi / Ly) * (np.fft.fftfreq(ny) * ny).reshape(1, -1)
        self.Krsq = self.kr ** 2 + self.l ** 2
        with np.errstate(divide="ignore"):
            self.invKrsq = 1.0 / self.Krsq
        self.invKrsq[0, 0] = 0.0
        self.aliased_fraction = aliased_fraction


class Clock:
    def __init__(self, prob):
        self._p = prob
        self.dt = prob.dt

    def _get(self):
        t, s = C.c_double(), C.c_longlong()
        check(lib().swrt_flow_clock(self._p._h, C.byref(t), C.byref(s)))
        return t.value, s.value

    @property
    def t(self):
        return self._get()[0]

    @property
    def step(self):
        return self._get()[1]


class Vars:
    """`prob.vars`: physical fields computed on demand from the device state (u, v, η, ζ)."""

    def __init__(self, prob):
        self._p = prob

    def _field(self, which):
        out = np.empty((self._p.grid.nx, self._p.grid.ny), dtype=np.float64, order="F")
        check(lib().swrt_flow_get_field(self._p._h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    u = property(lambda s: s._field(FIELD_U))
    v = property(lambda s: s._field(FIELD_V))
    η = property(lambda s: s._field(FIELD_ETA))
    eta = η
    ζ = property(lambda s: s._field(FIELD_ZETA))
    zeta = ζ


class Problem:
    """`RotatingShallowWater.Problem(dev; nx, ny, Lx, Ly, ν, nν, f, Cg, stepper, dt, aliased_fraction,
    T, use_filter, stepper_kwargs...)` (rsw/RotatingShallowWater.jl:70-99) on one B200."""

    def __init__(self, dev=0, *, model="RotatingShallowWater", nx=128, ny=None, Lx=2 * np.pi, Ly=None, ν=1.0e-16,
                 nν=4, f=1.0, Cg=1.0, stepper="IFMAB3", dt=5e-2, aliased_fraction=1 / 3, T=np.float64,
                 use_filter=False, order=4, innerK=2 / 3, outerK=1.0, tol=1e-15, nu=None, nnu=None):
        if T not in (np.float64, float, "Float64"):
            raise _lib.SwrtError("only T=Float64 is implemented (the north star's arithmetic)")
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        ν = ν if nu is None else nu
        nν = nν if nnu is None else nnu
        d = FlowDesc(model=MODELS[model], stepper=STEPPERS[stepper], nx=nx, ny=ny, nnu=nν, use_filter=int(use_filter),
                     filter_order=order, device=int(dev), Lx=Lx, Ly=Ly, dt=dt, nu=ν, f=f, Cg=Cg,
                     aliased_fraction=aliased_fraction, filter_innerK=innerK, filter_outerK=outerK, filter_tol=tol)
        self._h = C.c_void_p()
        check(lib().swrt_flow_create(C.byref(d), C.byref(self._h)))
        self.desc, self.dt, self.nvar = d, dt, 3
        self.grid = Grid(nx, ny, Lx, Ly, aliased_fraction)
        self.clock = Clock(self)
        self.vars = Vars(self)
        self.params = type("Params", (), dict(ν=ν, nν=nν, f=f, Cg2=Cg * Cg))()

    def close(self):
        if getattr(self, "_h", None):
            lib().swrt_flow_destroy(self._h)
            self._h = None

    __del__ = close

    # -- prob.sol ------------------------------------------------------------------
    @property
    def sol(self):
        out = np.empty((self.grid.nkr, self.grid.nl, self.nvar), dtype=np.complex128, order="F")
        check(lib().swrt_flow_get_solution(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @sol.setter
    def sol(self, value):
        a = np.asfortranarray(value, dtype=np.complex128)
        if a.shape != (self.grid.nkr, self.grid.nl, self.nvar):
            raise ValueError(f"sol must have shape {(self.grid.nkr, self.grid.nl, self.nvar)}")
        check(lib().swrt_flow_set_solution(self._h, a.ctypes.data_as(C.c_void_p)))

    def sync(self):
        check(lib().swrt_flow_sync(self._h))

    def launch_count(self):
        n = C.c_longlong()
        check(lib().swrt_flow_launch_count(self._h, C.byref(n)))
        return n.value

    def profile(self, enable=2):
        """Per-kernel CUDA-event timing on the handle's stream (2 = enable and clear, 0 = off)."""
        check(lib().swrt_flow_profile(self._h, int(enable)))

    def profile_report(self):
        out = {}
        for i in range(11):
            ms, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
            check(lib().swrt_flow_profile_get(self._h, i, C.byref(ms), C.byref(n), C.byref(name)))
            if n.value:
                out[name.value.decode()] = dict(ms_total=ms.value, launches=n.value, ms_avg=ms.value / n.value)
        return out

    def timer_start(self):
        check(lib().swrt_flow_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().swrt_flow_timer_stop(self._h, C.byref(ms)))
        return ms.value


def set_solution(prob, u0h, v0h, η0h):
    """set_solution!(prob, u0h, v0h, η0h)  rsw/RotatingShallowWater.jl:309-321"""
    prob.sol = np.stack([np.asarray(u0h), np.asarray(v0h), np.asarray(η0h)], axis=-1)


def enforce_reality_condition(prob):
    """enforce_reality_condition!(prob)  :118-133"""
    check(lib().swrt_flow_enforce_reality(prob._h))


def updatevars(prob):
    """updatevars!(prob) :101-116 -- fields are materialised lazily by `prob.vars`; nothing to do eagerly."""
    return None


def stepforward(prob, diags=(), nsteps=1):
    """FourierFlows.stepforward!(prob, diags, nsteps) with the IFMAB3 stepper (utils/IFMAB3.jl:157-169).
    `diags`: objects with `.freq` and `.increment(prob)`; sampled when step % freq == 0 like FourierFlows."""
    if not diags:
        check(lib().swrt_flow_step(prob._h, int(nsteps)))
        return
    for _ in range(int(nsteps)):
        check(lib().swrt_flow_step(prob._h, 1))
        step = prob.clock.step
        for d in diags:
            if step % d.freq == 0:
                d.increment(prob)


def kinetic_energy(prob):
    ke, pe = C.c_double(), C.c_double()
    check(lib().swrt_flow_energies(prob._h, C.byref(ke), C.byref(pe)))
    return ke.value


def potential_energy(prob):
    ke, pe = C.c_double(), C.c_double()
    check(lib().swrt_flow_energies(prob._h, C.byref(ke), C.byref(pe)))
    return pe.value


def max_abs_uv(prob):
    """maximum(abs.(vars.u)), maximum(abs.(vars.v))  (CFL log, raytracing/RaytracingDriver.jl:244)"""
    a, b = C.c_double(), C.c_double()
    check(lib().swrt_flow_max_abs_uv(prob._h, C.byref(a), C.byref(b)))
    return a.value, b.value


def has_nan(prob):
    """any(isnan.(prob.vars.uh))  raytracing/RaytracingDriver.jl:282"""
    f = C.c_int()
    check(lib().swrt_flow_has_nan(prob._h, C.byref(f)))
    return bool(f.value)
