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
INTERP_BILINEAR, INTERP_HERMITE_BICUBIC, INTERP_BSPLINE2, INTERP_BILINEAR_F32, INTERP_BSPLINE3, INTERP_NUFFT = 0, 1, 2, 3, 4, 5
INTEG_RK4, INTEG_IMPLICIT_MIDPOINT = 0, 1
RAYKERNEL_AUTO, RAYKERNEL_CACHED, RAYKERNEL_TILE, RAYKERNEL_TILE3, RAYKERNEL_PIPE = -1, 0, 1, 2, 3


class Velocity:
    """Velocity(u, v)  raytracing/GPURaytracing.jl:6-9 -- here: snapshot slot of `prob`."""

    def __init__(self, prob, slot):
        self.prob, self.slot = prob, slot

    def _arr(self):
        if getattr(self.prob, "world", 1) > 1:
            return self.prob.gather_snapshot(self.slot)             # team mode: assembled from the bands of all ranks
        g = self.prob.grid
        nf = C.c_int()
        check(lib().swrt_flow_snapshot_fields(self.prob._h, C.byref(nf)))
        nx, ny = C.c_int(), C.c_int()
        check(lib().swrt_flow_snapshot_dims(self.prob._h, C.byref(nx), C.byref(ny)))   # the node grid may be refined
        out = np.empty((nx.value, ny.value, nf.value), dtype=np.float64, order="F")
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
    if getattr(prob, "world", 1) > 1:
        prob.velocity_snapshot(slot, psi_kind)                     # team mode: this rank's band of rows
    else:
        check(lib().swrt_flow_velocity_snapshot(prob._h, psi_kind, slot))
    return Velocity(prob, slot), VelocityGradient(prob, slot)


def set_interpolation(prob, interp):
    """Choose the node data the flow's snapshots hold: INTERP_BILINEAR (5 fields) or INTERP_HERMITE_BICUBIC (7 fields)."""
    check(lib().swrt_flow_set_interp(prob._h, int(interp)))


def set_nufft_width(prob, nw):
    """Kernel width of INTERP_NUFFT (nodes of the 2x oversampled grid); the sampling error is ~ 10^(1 - nw)."""
    check(lib().swrt_flow_set_nufft_width(prob._h, int(nw)))


def set_snapshot_refinement(prob, refine):
    """"FFT interpolation": the snapshots' node grid becomes `refine` (1 or 2) times finer than the flow's by spectral zero
    padding, i.e. the exact trigonometric interpolant sampled on the finer grid (what raytracing/NUFFTRaytracing.jl:68-84 aims at
    and Notebooks/FFTInterpTest.ipynb does on the host).  Call before creating packets."""
    check(lib().swrt_flow_set_snapshot_refinement(prob._h, int(refine)))


def set_velocity_info(prob, slot, fields):
    """Load (nx, ny, 5) = u, v, ux, uy, vx (or (nx, ny, 7) with uxy, vxy in Hermite mode) host fields into a slot."""
    if getattr(prob, "world", 1) > 1:
        return prob.set_snapshot(slot, fields)
    a = np.asfortranarray(fields, dtype=np.float64)
    nf = C.c_int()
    check(lib().swrt_flow_snapshot_fields(prob._h, C.byref(nf)))
    nx, ny = C.c_int(), C.c_int()
    check(lib().swrt_flow_snapshot_dims(prob._h, C.byref(nx), C.byref(ny)))
    assert a.shape == (nx.value, ny.value, nf.value), a.shape
    check(lib().swrt_flow_set_snapshot(prob._h, slot, a.ctypes.data_as(C.c_void_p)))


def swap_snapshots(prob, alias=False):
    """old_velocity = new_velocity; old_grad_v = new_grad_v  (RaytracingDriver.jl:269-270)."""
    check(lib().swrt_flow_swap_snapshots(prob._h, int(alias)))


class Packets:
    """Device-resident wave packets = `create_template_ode(packets)` + the packet arrays."""

    def __init__(self, prob, n, f, Cg, nsub=1, time_lerp=LERP_PHYSICAL, sort_every=16, interp=INTERP_BILINEAR,
                 integrator=INTEG_RK4, first=None, capacity=None, overlap=True):
        """On a slab-decomposed problem (slab.SlabProblem, team mode) the packets are sharded by y-band: `n` is this rank's
        caller-order block, rows [first, first + n) of the ensemble (default: blocks in rank order), `capacity` the packets a
        rank can host (default 1.5 x the mean + 4096, the same on every rank); every method becomes a collective call.
        `overlap` puts the band packets on their own stream, so the ray tracing of step n runs beside the flow step n + 1."""
        self.prob, self.n = prob, int(n)
        self.band = getattr(prob, "world", 1) > 1
        if self.band:
            counts = prob._gather(self.n)
            if first is None:
                first = sum(counts[:prob.rank])
            if capacity is None:
                capacity = max(int(1.5 * -(-sum(counts) // prob.world)) + 4096, max(counts))
            if len(set(prob._gather(int(capacity)))) != 1:
                raise ValueError("band packets: `capacity` must be the same on every rank")
        self.first = int(first or 0)
        d = PacketsDesc(n=self.n, interp=int(interp), integrator=int(integrator), nsub=int(nsub), time_lerp=int(time_lerp), sort_every=int(sort_every), f=f, Cg=Cg,
                        band_first=self.first, band_capacity=int(capacity or 0))
        self._h = C.c_void_p()
        check(lib().swrt_packets_create(C.byref(d), prob._h, C.byref(self._h)))
        if self.band:                                               # one exchange of 64-byte IPC handles
            buf = C.create_string_buffer(64)
            check(lib().swrt_packets_ipc_handle(self._h, buf))
            for r, h in enumerate(prob._gather(buf.raw)):
                check(lib().swrt_packets_ipc_open(self._h, r, h))
            if overlap:
                self.use_own_stream()
            prob.dist.barrier()

    def resident(self):
        """Packets currently hosted by this rank (band mode; = n otherwise)."""
        n = C.c_longlong()
        check(lib().swrt_packets_resident(self._h, C.byref(n)))
        return n.value

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

    def set_kernel(self, kernel):
        """RAYKERNEL_AUTO (-1), RAYKERNEL_CACHED (0), RAYKERNEL_TILE (1), RAYKERNEL_TILE3 (2) or RAYKERNEL_PIPE (3): see swrt_packets_set_kernel."""
        check(lib().swrt_packets_set_kernel(self._h, int(kernel)))

    def generate(self, L, k0, sqrtN, first=0):
        check(lib().swrt_packets_generate(self._h, L, k0, int(sqrtN), int(first)))

    # -- overlapped I/O: the handle's own stream, asynchronous copies from / to page-locked row blocks -------------------
    def use_own_stream(self):
        check(lib().swrt_packets_use_own_stream(self._h))

    def _cols(self, a, ncol):
        """(pointer, leading dimension) of an (n, ncol) float64 block whose columns are contiguous (a row block of a
        Fortran-ordered array qualifies)."""
        assert a.dtype == np.float64 and a.shape == (self.n, ncol) and (self.n == 1 or a.strides[0] == 8), (a.shape, a.strides)
        return a.ctypes.data_as(C.c_void_p), (a.strides[1] // 8 if ncol > 1 else self.n)

    def set_async(self, xk, frequency_sign=None):
        p, ld = self._cols(xk, 4)
        s = None
        if frequency_sign is not None:
            assert frequency_sign.dtype == np.float64 and frequency_sign.flags.c_contiguous and frequency_sign.shape == (self.n,)
            s = frequency_sign.ctypes.data_as(C.c_void_p)
        check(lib().swrt_packets_set_async(self._h, p, ld, s))

    def get_async(self, out):
        p, ld = self._cols(out, 4)
        check(lib().swrt_packets_get_async(self._h, p, ld))

    def sample_async(self, slot, out_U, out_G=None):
        pu, ld = self._cols(out_U, 2)
        pg = None
        if out_G is not None:
            pg, ldg = self._cols(out_G, 4)
            assert ldg == ld, "velocity and gradient blocks must share the leading dimension"
        check(lib().swrt_packets_sample_async(self._h, int(slot), pu, pg, ld))

    def sync(self):
        check(lib().swrt_packets_sync(self._h))

    def kcutoff_reset(self, k_cutoff, k0):
        n = C.c_longlong()
        check(lib().swrt_packets_kcutoff_reset(self._h, k_cutoff, k0, C.byref(n)))
        return sum(self.prob._gather(n.value)) if self.band else n.value


def generate_initial_wavepackets(prob, L, k0, Npackets, sqrtNpackets, f, Cg, nsub=1, first=0, time_lerp=LERP_PHYSICAL,
                                 sort_every=16):
    """raytracing/RaytracingDriver.jl:27-47; `first`/`Npackets` select a contiguous shard of the lattice."""
    p = Packets(prob, Npackets, f, Cg, nsub=nsub, time_lerp=time_lerp, sort_every=sort_every, first=first)
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


class PacketPipeline:
    """An ensemble split over `nchunks` packet handles, each on its own stream: per chunk
    upload -> sort + ray-trace -> download, so the uploads, kernels and downloads of different chunks overlap on the two copy
    engines and the SMs.  Host arrays are (N, ncol) Fortran-ordered and should be page-locked; chunk c owns the contiguous
    rows [c N / nchunks, (c+1) N / nchunks), so row order is the caller's throughout."""

    def __init__(self, prob, n, f, Cg, nchunks=8, **kw):
        """Equal row blocks: a step is a two-stage pipeline (upload of block c + 1 beside the download of block c, both PCIe
        directions busy), so what cannot overlap is one block's upload and one block's download -- more, equal blocks shorten
        exactly that; blocks that grow towards the middle were tried and lose (the download of a block always waits for the
        upload of a larger one: 14.4 ms against 12.5 ms per step at 16.8 M packets, profiles/r02_s)."""
        self.prob, self.n = prob, int(n)
        from .parallel import shard_range
        self.bounds = [b for b in (shard_range(self.n, c, nchunks) for c in range(nchunks)) if b[1] > b[0]]   # the rank-shard rule; empty blocks dropped
        self.chunks = [Packets(prob, hi - lo, f, Cg, **kw) for lo, hi in self.bounds]
        for p in self.chunks:
            p.use_own_stream()

    def close(self):
        for p in self.chunks:
            p.close()

    def step(self, xk_in, sign, tspan, xk_out, out_U=None, out_G=None, after_raytrace=None, sample_slot=0):
        """set -> raytrace over tspan -> get, block by block, -> (after_raytrace(): e.g. swap_snapshots) [-> sample], all asynchronous;
        returns after every block's results are on the host.  The download of a block is enqueued right behind its kernels (not
        after every block's kernels have been enqueued: the launches of 8-32 blocks take the host 1-2 ms, during which the
        download engine would sit idle)."""
        for p, (lo, hi) in zip(self.chunks, self.bounds):
            p.set_async(xk_in[lo:hi], None if sign is None else sign[lo:hi])
            check(lib().swrt_packets_raytrace(p._h, float(tspan[0]), float(tspan[1])))
            p.get_async(xk_out[lo:hi])
        if after_raytrace is not None:
            after_raytrace()
        if out_U is not None:
            for p, (lo, hi) in zip(self.chunks, self.bounds):
                p.sample_async(sample_slot, out_U[lo:hi], None if out_G is None else out_G[lo:hi])
        for p in self.chunks:
            p.sync()
