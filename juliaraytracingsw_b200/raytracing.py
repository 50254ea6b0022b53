"""Host-side mirror of raytracing/GPURaytracing.jl over the C ABI.

`Velocity` / `VelocityGradient` are handles on the flow's two device-resident snapshot slots
(0 = old, 1 = new) instead of bundles of CuArrays; `raytrace`, `interpolate_velocity`,
`interpolate_gradients` keep the reference's argument order.  Packets are `(N, 4)` float64 arrays
with columns x, y, k, l (Fortran order at the boundary).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PacketsDesc, check, lib

PSI_RSW_BALANCED, PSI_SWQG, PSI_TWOLAYER_BAROCLINIC, PSI_TWOLAYER_MEAN = 0, 1, 2, 3
LERP_PHYSICAL, LERP_REFERENCE_GPU = 0, 1
INTERP_BILINEAR, INTERP_HERMITE_BICUBIC, INTERP_BSPLINE2 = 0, 1, 2
INTEG_RK4, INTEG_IMPLICIT_MIDPOINT = 0, 1


class Velocity:
    """Velocity(u, v)  raytracing/GPURaytracing.jl:6-9 -- here: snapshot slot of `prob`."""

    def __init__(self, prob, slot):
        self.prob, self.slot = prob, slot

    def _arr(self):
        g = self.prob.grid
        nf = C.c_int()
        check(lib().swrt_flow_snapshot_fields(self.prob._h, C.byref(nf)))
        out = np.empty((g.nx, g.ny, nf.value), dtype=np.float64, order="F")
        check(lib().swrt_flow_get_snapshot(self.prob._h, self.slot, out.ctypes.data_as(C.c_void_p)))
        return out

    u = property(lambda s: s._arr()[:, :, 0])
    v = property(lambda s: s._arr()[:, :, 1])


class VelocityGradient(Velocity):
    """VelocityGradient(ux, uy, vx, vy)  :11-16"""
    ux = property(lambda s: s._arr()[:, :, 2])
    uy = property(lambda s: s._arr()[:, :, 3])
    vx = property(lambda s: s._arr()[:, :, 4])
    vy = property(lambda s: -s._arr()[:, :, 2])


def get_velocity_info(prob, slot, psi_kind=PSI_RSW_BALANCED):
    """get_streamfunction! + get_velocity_info (rsw/RSWRaytracingDriver.jl:56-67,
    raytracing/RaytracingDriver.jl:132-154) into snapshot `slot`; returns (Velocity, VelocityGradient)."""
    check(lib().swrt_flow_velocity_snapshot(prob._h, psi_kind, slot))
    return Velocity(prob, slot), VelocityGradient(prob, slot)


def set_interpolation(prob, interp):
    """Choose the node data the flow's snapshots hold: INTERP_BILINEAR (5 fields) or INTERP_HERMITE_BICUBIC (7 fields)."""
    check(lib().swrt_flow_set_interp(prob._h, int(interp)))


def set_velocity_info(prob, slot, fields):
    """Load (nx, ny, 5) = u, v, ux, uy, vx (or (nx, ny, 7) with uxy, vxy in Hermite mode) host fields into a slot."""
    a = np.asfortranarray(fields, dtype=np.float64)
    nf = C.c_int()
    check(lib().swrt_flow_snapshot_fields(prob._h, C.byref(nf)))
    assert a.shape == (prob.grid.nx, prob.grid.ny, nf.value), a.shape
    check(lib().swrt_flow_set_snapshot(prob._h, slot, a.ctypes.data_as(C.c_void_p)))


def swap_snapshots(prob, alias=False):
    """old_velocity = new_velocity; old_grad_v = new_grad_v  (RaytracingDriver.jl:269-270)."""
    check(lib().swrt_flow_swap_snapshots(prob._h, int(alias)))


class Packets:
    """Device-resident wave packets = `create_template_ode(packets)` + the packet arrays."""

    def __init__(self, prob, n, f, Cg, nsub=1, time_lerp=LERP_PHYSICAL, sort_every=16, interp=INTERP_BILINEAR,
                 integrator=INTEG_RK4):
        self.prob, self.n = prob, int(n)
        d = PacketsDesc(n=self.n, interp=int(interp), integrator=int(integrator), nsub=int(nsub), time_lerp=int(time_lerp), sort_every=int(sort_every), f=f, Cg=Cg)
        self._h = C.c_void_p()
        check(lib().swrt_packets_create(C.byref(d), prob._h, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().swrt_packets_destroy(self._h)
            self._h = None

    __del__ = close

    def set(self, xk, frequency_sign=None):
        a = np.asfortranarray(xk, dtype=np.float64)
        assert a.shape == (self.n, 4)
        s = None if frequency_sign is None else np.ascontiguousarray(frequency_sign, dtype=np.float64)
        check(lib().swrt_packets_set(self._h, a.ctypes.data_as(C.c_void_p), None if s is None else s.ctypes.data_as(C.c_void_p)))

    def get(self, out=None):
        """Array(packets); `out` may be a caller-owned (N, 4) Fortran-ordered buffer (e.g. pinned memory)."""
        if out is None:
            out = np.empty((self.n, 4), dtype=np.float64, order="F")
        assert out.shape == (self.n, 4) and out.flags.f_contiguous and out.dtype == np.float64
        check(lib().swrt_packets_get(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def generate(self, L, k0, sqrtN, first=0):
        check(lib().swrt_packets_generate(self._h, L, k0, int(sqrtN), int(first)))

    def kcutoff_reset(self, k_cutoff, k0):
        n = C.c_longlong()
        check(lib().swrt_packets_kcutoff_reset(self._h, k_cutoff, k0, C.byref(n)))
        return n.value


def generate_initial_wavepackets(prob, L, k0, Npackets, sqrtNpackets, f, Cg, nsub=1, first=0, time_lerp=LERP_PHYSICAL,
                                 sort_every=16):
    """raytracing/RaytracingDriver.jl:27-47; `first`/`Npackets` select a contiguous shard of the lattice."""
    p = Packets(prob, Npackets, f, Cg, nsub=nsub, time_lerp=time_lerp, sort_every=sort_every)
    p.generate(L, k0, sqrtNpackets, first)
    return p


def create_template_ode(packets):
    """raytracing/GPURaytracing.jl:111-113 -- nothing to precompute for fixed-step RK4."""
    return packets


def raytrace(ode_template, velocity1, velocity2, gradient1, gradient2, grid, wavepacket_array, dt, tspan, params=None):
    """raytrace!(tmpl, v_old, v_new, g_old, g_new, grid, packets, dt, (t0, t1), params)  :115-142.
    The velocity arguments name the flow's snapshot slots; like the reference, `dt` is unused."""
    check(lib().swrt_packets_raytrace(wavepacket_array._h, float(tspan[0]), float(tspan[1])))


def interpolate_velocity(velocity, packets, output_U=None):
    """interpolate_velocity! :67-82 -> (N, 2) array of u, v at the packet positions."""
    U = np.empty((packets.n, 2), dtype=np.float64, order="F") if output_U is None else output_U
    check(lib().swrt_packets_sample(packets._h, velocity.slot, U.ctypes.data_as(C.c_void_p), None))
    return U


def interpolate_gradients(gradient, packets, output_G=None, output_U=None):
    """interpolate_gradients! :84-109 -> (N, 4) array of ux, uy, vx, vy at the packet positions
    (the velocity is sampled in the same pass and returned in `output_U` when given)."""
    U = np.empty((packets.n, 2), dtype=np.float64, order="F") if output_U is None else output_U
    G = np.empty((packets.n, 4), dtype=np.float64, order="F") if output_G is None else output_G
    check(lib().swrt_packets_sample(packets._h, gradient.slot, U.ctypes.data_as(C.c_void_p), G.ctypes.data_as(C.c_void_p)))
    return G
