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

MODELS = {"RotatingShallowWater": 0, "ModifiedShallowWater": 1, "LinborgShallowWater": 2, "QuadHeightModifiedShallowWater": 3, "SWQG": 4,
          "TwoLayerQG": 5, "ThomasYamada": 6, "MultiLayerQG": 7}
STEPPERS = {"IFMAB3": 0, "FilteredAB3": 1, "ETDRK4": 2, "FilteredRK4": 3, "FilteredETDRK4": 4}
FIELD_U, FIELD_V, FIELD_ETA, FIELD_ZETA = 0, 1, 2, 16
FIELD_QG_PSI, FIELD_QG_U, FIELD_QG_V, FIELD_QG_ZETA = 32, 40, 48, 56
NVAR = {0: 3, 1: 3, 2: 3, 3: 3, 4: 1, 5: 2, 6: 4, 7: 2}


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
    """`prob.vars`: physical fields computed on demand from the device state.
    RSW family: u, v, η, ζ (rsw/RotatingShallowWater.jl:101-116); QG models: q, ψ, ζ, u, v
    (swqg/SWQG.jl:109-125; swqg/TwoLayerQG.jl:113-129 with a trailing layer axis)."""

    def __init__(self, prob):
        self._p = prob

    def _field(self, which):
        out = np.empty((self._p.grid.nx, self._p.grid.ny), dtype=np.float64, order="F")
        check(lib().swrt_flow_get_field(self._p._h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def _qg(self, base):
        n = self._p.nvar
        if n == 1:
            return self._field(base)
        return np.stack([self._field(base + j) for j in range(n)], axis=-1)

    def __getattr__(self, name):
        qg = self._p.desc.model in (4, 5, 7)
        if self._p.desc.model == 6:     # Thomas-Yamada state fields (thomasyamada/ThomasYamada.jl:76-88)
            ty = {"ζt": 0, "zetat": 0, "uc": 1, "vc": 2, "pc": 3}
            if name not in ty:
                raise AttributeError(name)
            return self._field(ty[name])
        if self._p.desc.model == 3 and name in ("m",):
            return self._field(FIELD_ETA)
        table = ({"q": 0, "ψ": FIELD_QG_PSI, "psi": FIELD_QG_PSI, "u": FIELD_QG_U, "v": FIELD_QG_V, "ζ": FIELD_QG_ZETA,
                  "zeta": FIELD_QG_ZETA} if qg else
                 {"u": FIELD_U, "v": FIELD_V, "η": FIELD_ETA, "eta": FIELD_ETA, "ζ": FIELD_ZETA, "zeta": FIELD_ZETA})
        if name not in table:
            raise AttributeError(name)
        return self._qg(table[name]) if qg else self._field(table[name])


class Problem:
    """`RotatingShallowWater.Problem(dev; nx, ny, Lx, Ly, ν, nν, f, Cg, stepper, dt, aliased_fraction,
    T, use_filter, stepper_kwargs...)` (rsw/RotatingShallowWater.jl:70-99) on one B200."""

    def __init__(self, dev=0, *, model="RotatingShallowWater", nx=128, ny=None, Lx=2 * np.pi, Ly=None, ν=1.0e-16,
                 nν=4, f=1.0, Cg=1.0, stepper="IFMAB3", dt=5e-2, aliased_fraction=1 / 3, T=np.float64,
                 use_filter=False, order=4, innerK=2 / 3, outerK=1.0, tol=1e-15, nu=None, nnu=None,
                 U=0.5, μ=1e-2, f0=None, δρρ0=0.2, mu=None, Ro=0.2, slab=None, H=None, b=None, β=0.0, beta=None):
        """Two-layer QG (swqg/TwoLayerQG.jl:55-72) takes U, μ, f0, Cg, δρρ0 (F = 2 f0²/Cg²/δρρ0); SWQG takes f, Cg (Kd2 = f²/Cg²).
        `model="MultiLayerQG"` mirrors `MultiLayerQG.Problem(2, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction)`
        (raytracing/TwoLayerRaytracing.jl:174) for two equal layers: U = (U₁, U₂), F = f₀²/((b₁ - b₂) H_j)."""
        if T not in (np.float64, float, "Float64"):
            raise _lib.SwrtError("only T=Float64 is implemented (the north star's arithmetic)")
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        ν = ν if nu is None else nu
        nν = nν if nnu is None else nnu
        μ = μ if mu is None else mu
        f0 = f if f0 is None else f0
        F = 2 * f0 ** 2 / Cg ** 2 / δρρ0
        U2 = 0.0
        β = β if beta is None else beta
        if model == "MultiLayerQG":
            H = (0.5, 0.5) if H is None else tuple(float(x) for x in H)
            b = (2.0, 1.0) if b is None else tuple(float(x) for x in b)
            if len(H) != 2 or len(b) != 2 or H[0] != H[1]:
                raise _lib.SwrtError("MultiLayerQG is implemented for two layers of equal depth")
            F = f0 ** 2 / ((b[0] - b[1]) * H[0])
            U, U2 = (float(U[0]), float(U[1])) if np.ndim(U) else (float(U), -float(U))
        d = FlowDesc(model=MODELS[model], stepper=STEPPERS[stepper], nx=nx, ny=ny, nnu=nν, use_filter=int(use_filter),
                     filter_order=order, device=int(dev), Lx=Lx, Ly=Ly, dt=dt, nu=ν, f=f, Cg=Cg,
                     aliased_fraction=aliased_fraction, filter_innerK=innerK, filter_outerK=outerK, filter_tol=tol,
                     U=U, mu=μ, F=F, Ro=Ro, U2=U2, beta=β, slab_rank=0 if slab is None else int(slab[0]),
                     slab_size=0 if slab is None else int(slab[1]))
        self._h = C.c_void_p()
        check(lib().swrt_flow_create(C.byref(d), C.byref(self._h)))
        self.desc, self.dt, self.nvar = d, dt, NVAR[d.model]
        self.F = F
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
        return out[:, :, 0] if self.desc.model == 4 else out        # SWQG's sol is (nkr, nl)

    @sol.setter
    def sol(self, value):
        a = np.asarray(value, dtype=np.complex128)
        if a.ndim == 2:
            a = a[:, :, None]
        a = np.asfortranarray(a)
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
        for i in range(13):
            ms, n, name = C.c_double(), C.c_longlong(), C.c_char_p()
            check(lib().swrt_flow_profile_get(self._h, i, C.byref(ms), C.byref(n), C.byref(name)))
            if n.value:
                out[name.value.decode()] = dict(ms_total=ms.value, launches=n.value, ms_avg=ms.value / n.value)
        return out

    def ray_kernel_name(self):
        """Name of the ray kernel the last `raytrace` call on this flow's packets launched (the profile label)."""
        name = C.c_char_p()
        check(lib().swrt_flow_profile_get(self._h, 6, None, None, C.byref(name)))
        return name.value.decode()

    def timer_start(self):
        check(lib().swrt_flow_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().swrt_flow_timer_stop(self._h, C.byref(ms)))
        return ms.value


def set_solution(prob, *fields):
    """set_solution!(prob, u0h, v0h, η0h) rsw/RotatingShallowWater.jl:309-321; (prob, ζ0h, u0h, v0h, p0h)
    thomasyamada/ThomasYamada.jl:292-317; (prob, q0h) swqg/TwoLayerQG.jl:211-219."""
    if len(fields) == 1:
        prob.sol = np.asarray(fields[0])
    elif prob.desc.model == 3:
        # rsw/QuadHeightModifiedShallowWater.jl:333-347: the third variable is m = 1/(1+eta), formed in physical space
        prob.sol = np.stack([np.asarray(f) for f in fields], axis=-1)
        set_field_physical(prob, 2, 1.0 / (1.0 + prob.vars._field(FIELD_ETA)))
    else:
        prob.sol = np.stack([np.asarray(f) for f in fields], axis=-1)


def set_field_physical(prob, var, field):
    """mul!(varh, grid.rfftplan, field): overwrite state variable `var` with the (dealiased) transform of a physical field."""
    a = np.asfortranarray(field, dtype=np.float64)
    assert a.shape == (prob.grid.nx, prob.grid.ny)
    check(lib().swrt_flow_set_field_physical(prob._h, int(var), a.ctypes.data_as(C.c_void_p)))


def enforce_reality_condition(prob):
    """enforce_reality_condition!(prob)  :118-133"""
    check(lib().swrt_flow_enforce_reality(prob._h))


def updatevars(prob):
    """updatevars!(prob) :101-116 -- fields are materialised lazily by `prob.vars`; nothing to do eagerly."""
    return None


def set_forcing(prob, Fh):
    """vars.Fh of the forcing hook (rsw/RotatingShallowWater.jl:228-240): the (nkr, nl) complex field the caller's calcF! produced.
    calcN! adds it to every component of N from now on; `None` removes it."""
    if Fh is None:
        check(lib().swrt_flow_set_forcing(prob._h, None))
        return
    a = np.asfortranarray(Fh, dtype=np.complex128)
    assert a.shape == (prob.grid.nkr, prob.grid.nl), a.shape
    check(lib().swrt_flow_set_forcing(prob._h, a.ctypes.data_as(C.c_void_p)))


def stepforward(prob, diags=(), nsteps=1, calcF=None):
    """FourierFlows.stepforward!(prob, diags, nsteps) with the IFMAB3 stepper (utils/IFMAB3.jl:157-169).
    `diags`: objects with `.freq` and `.increment(prob)`; sampled when step % freq == 0 like FourierFlows.
    `calcF(Fh, t, clock)`: the `calcF!` of `Problem(...; calcF!)` for the one-calcN!-per-step steppers (IFMAB3, FilteredAB3): called
    before every step with a host (nkr, nl) complex array to fill, which then goes to the device (set_forcing)."""
    if calcF is not None:
        Fh = np.zeros((prob.grid.nkr, prob.grid.nl), dtype=np.complex128, order="F")
        for _ in range(int(nsteps)):
            calcF(Fh, prob.clock.t, prob.clock)
            set_forcing(prob, Fh)
            stepforward(prob, diags, 1)
        return
    if not diags:
        check(lib().swrt_flow_step(prob._h, int(nsteps)))
        return
    for _ in range(int(nsteps)):
        check(lib().swrt_flow_step(prob._h, 1))
        step = prob.clock.step
        for d in diags:
            if step % d.freq == 0:
                d.increment(prob)


def wave_balanced_decomposition(prob):
    """`wave_balanced_decomposition(prob)` (rsw/RSWUtils.jl:5-22) / `decompose_balanced_wave(sol, grid)`
    (thomasyamada/TYUtils.jl:40-51): (balanced, wave) as complex (nkr, nl, 3) arrays, projected on the device."""
    shape = (prob.grid.nkr, prob.grid.nl, 3)
    bal, wav = (np.empty(shape, dtype=np.complex128, order="F") for _ in range(2))
    check(lib().swrt_flow_wave_balanced_decomposition(prob._h, bal.ctypes.data_as(C.c_void_p), wav.ctypes.data_as(C.c_void_p)))
    return bal, wav


def compute_balanced_wave_weights(prob):
    """`compute_balanced_wave_weights(uh, vh, ηh, Φ₀, Φ₊, Φ₋, params)` with the bases of rsw/RSWUtils.jl:24-49 -> (c₀, c₊, c₋)."""
    c = [np.empty((prob.grid.nkr, prob.grid.nl), dtype=np.complex128, order="F") for _ in range(3)]
    check(lib().swrt_flow_wave_balanced_weights(prob._h, *(a.ctypes.data_as(C.c_void_p) for a in c)))
    return tuple(c)


def wave_geostrophic_energy(prob):
    """`wave_geostrophic_energy(prob)` (thomasyamada/ThomasYamada.jl:355-367): ((KE_w, PE_w), (KE_g, PE_g)), reduced on the device."""
    out = (C.c_double * 4)()
    check(lib().swrt_flow_wave_balanced_energies(prob._h, out))
    return (out[0], out[1]), (out[2], out[3])


def barotropic_energy(prob):
    """`barotropic_energy(prob)` (thomasyamada/ThomasYamada.jl:343-350)."""
    v = C.c_double()
    check(lib().swrt_flow_barotropic_energy(prob._h, C.byref(v)))
    return v.value


def kinetic_energy(prob):
    """kinetic_energy(prob); the two-layer model returns (KE_1, KE_2) like swqg/TwoLayerQG.jl:221-233."""
    if prob.desc.model in (5, 7):
        out = []
        for layer in range(2):
            v = C.c_double()
            check(lib().swrt_flow_layer_kinetic_energy(prob._h, layer, C.byref(v)))
            out.append(v.value)
        return tuple(out)
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
