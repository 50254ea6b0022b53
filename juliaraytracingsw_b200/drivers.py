"""Host-side mirror of the reference's coupled flow + packet driver (the caller of both hot paths).

Follows raytracing/RaytracingDriver.jl (`initialize_problem` :49-85, `start_raytracing!` :156-292,
`savepacketdata!` :96-124) and rsw/RSWRaytracingDriver.jl (`set_initial_condition!` :15-54,
`estimate_max_U` :69-71) with the call sequence of rsw/SingleWaveRSWRaytracingDriver.jl:154-299
(the committed RaytracingDriver.jl is mid-refactor).  All transforms and
reductions run on the device through libswrt; nothing here touches the oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import flow, raytracing
from .outputs import SequencedOutput


@dataclass
class Parameters:
    """rsw/RSWRaytracingParameters.jl (module Parameters), same names and defaults."""
    L: float = 2 * np.pi
    nx: int = 512
    background_Cg: float = 1.0
    f: float = 3.0
    nν: int = 4
    νtune: float = 1.0
    cfltune: float = 0.1
    filter_order: int = 8
    aliased_fraction: float = 1 / 3
    packet_spinup_T: float = 1000.0
    spinup_T: float = 1000.0
    T: float = 2000.0
    output_dt: float = 10.0 / 3.0
    diag_dt: float = 0.5
    max_writes: int = 300
    base_filename: str = "rsw"
    Kg: tuple = (10, 13)
    ag: float = 1.5
    Kw: tuple = (0, 5)
    aw: float = 0.1
    packet_base_filename: str = "packets"
    use_stationary_background_flow: bool = False
    write_gradients: bool = True
    packet_max_writes: int = 300
    packet_output_dt: float = 1.0
    sqrtNpackets: int = 128
    ω0: float = 6.0
    nsub: int = 1                       # RK4 sub-steps per flow step (the reference integrates adaptively)
    seed: int = 1234

    @property
    def use_filter(self):
        return self.νtune == 0

    @property
    def Npackets(self):
        return self.sqrtNpackets ** 2

    @property
    def packet_Cg(self):
        return self.background_Cg

    @property
    def Cg(self):
        return self.background_Cg


def estimate_max_U(P: Parameters):
    return P.ag + P.aw


def timestep_and_viscosity(P: Parameters, umax=None):
    """raytracing/RaytracingDriver.jl:52-63."""
    umax = estimate_max_U(P) if umax is None else umax
    dx = P.L / P.nx
    kmax = P.nx / 2 - 1
    dt = P.cfltune / umax * dx
    ν = P.νtune * 2 * np.pi / P.nx / (kmax ** (2 * P.nν)) / dt
    return dt, ν


def create_fourier_flows_problem(dev, P: Parameters, dt, ν):
    """rsw/RSWRaytracingDriver.jl:73-75 (T = Float64: the north star's arithmetic)."""
    return flow.Problem(dev, nx=P.nx, Lx=P.L, dt=dt, f=P.f, Cg=P.Cg, ν=ν, nν=P.nν, order=P.filter_order,
                        use_filter=P.use_filter, aliased_fraction=P.aliased_fraction)


def set_initial_condition(prob, P: Parameters, rng: np.random.Generator):
    """rsw/RSWRaytracingDriver.jl:15-54: random-phase geostrophic band + wave band, scaled so that
    max|u_g| = ag and max|u_w| = aw (K15).  Only the random numbers are drawn on the host (NumPy's stream, not Julia's);
    the mode arithmetic, the two normalising inverse transforms and the maxima run on the device
    (swrt_flow_set_rsw_initial_condition)."""
    import ctypes as C
    from ._lib import check, lib
    g = prob.grid
    shape = (g.nkr, g.nl)
    phase = np.asfortranarray(2 * np.pi * rng.random(shape))
    sgn = np.asfortranarray(np.sign(rng.random(shape) - 0.5))
    scales = (C.c_double * 2)()
    check(lib().swrt_flow_set_rsw_initial_condition(prob._h, phase.ctypes.data_as(C.c_void_p), sgn.ctypes.data_as(C.c_void_p),
                                                    float(P.Kg[0]), float(P.Kg[1]), float(P.ag), float(P.Kw[0]), float(P.Kw[1]),
                                                    float(P.aw), scales))
    return scales[0], scales[1]


def initialize_problem(P: Parameters, dev=0):
    """raytracing/RaytracingDriver.jl:49-85 -> (prob, cadence integers)."""
    dt, ν = timestep_and_viscosity(P)
    prob = create_fourier_flows_problem(dev, P, dt, ν)
    cad = dict(
        spinup_step=math.floor(P.spinup_T / dt),
        packet_spinup_step=math.floor(P.packet_spinup_T / dt),
        packet_output_freq=max(math.floor(P.packet_output_dt / dt), 1),
        diags_freq=max(math.floor(P.diag_dt / dt), 1),
        nsteps=math.ceil(P.T / dt),
    )
    cad["output_per_packet_freq"] = max(math.floor(P.output_dt / P.packet_output_dt), 1)
    cad["output_freq"] = cad["output_per_packet_freq"] * cad["packet_output_freq"]
    set_initial_condition(prob, P, np.random.default_rng(P.seed))
    return prob, cad


def coupled_step(prob, packets, old_t):
    """Body of the hot loop, raytracing/RaytracingDriver.jl:256-270: one flow step, new velocity snapshot,
    ray-trace all packets across it, new becomes old.  Returns new_t."""
    if getattr(prob, "world", 1) > 1:
        prob.stepforward(1)                                        # team mode: slab-decomposed step
    else:
        flow.stepforward(prob, (), 1)
        flow.updatevars(prob)
    new_velocity, new_grad_v = raytracing.get_velocity_info(prob, 1)
    new_t = prob.clock.t
    raytracing.raytrace(packets, None, new_velocity, None, new_grad_v, prob.grid, packets, prob.dt, (old_t, new_t))
    raytracing.swap_snapshots(prob, alias=False)
    return new_t


def coupled_steps(prob, packets, nsteps, psi_kind=raytracing.PSI_RSW_BALANCED, k_cutoff=0.0, k0=0.0):
    """`nsteps` iterations of `coupled_step` in one library call (same kernels, same order): returns the new time."""
    from ._lib import check, lib
    check(lib().swrt_packets_coupled_steps(packets._h, int(psi_kind), int(nsteps), float(k_cutoff), float(k0)))
    return prob.clock.t


@dataclass
class Frame:
    step: int
    t: float
    x: np.ndarray
    k: np.ndarray
    u: np.ndarray
    g: np.ndarray | None = None


def savepacketdata(out: SequencedOutput, frames: list, prob, packets, slot, write_gradients):
    """savepacketdata! + write_packets! (raytracing/RaytracingDriver.jl:96-124): sample the background at the
    packet positions, copy to the host, and emit the keys p/t, p/x, p/k, p/u[, p/g] in the reference's order."""
    xk = packets.get()
    step, t = prob.clock.step, prob.clock.t
    if write_gradients:
        U = np.empty((packets.n, 2), order="F")
        G = raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, slot), packets, output_U=U)
    else:
        U, G = raytracing.interpolate_velocity(raytracing.Velocity(prob, slot), packets), None
    fr = Frame(step, t, xk[:, 0:2], xk[:, 2:4], U, G)
    out[f"p/t/{step}"] = t
    out[f"p/x/{step}"] = fr.x
    out[f"p/k/{step}"] = fr.k
    out[f"p/u/{step}"] = fr.u
    if write_gradients:
        out[f"p/g/{step}"] = fr.g
    frames.append(fr)


def start_raytracing(P: Parameters, dev=0, max_frames=None, shard=(0, 1)):
    """start_raytracing!() (raytracing/RaytracingDriver.jl:156-292) for the RSW model.  `shard = (rank, world)`
    gives this process the contiguous packet block [rank N/world, (rank+1) N/world); the flow is replicated.
    Returns (prob, packets, packet_output, frames)."""
    prob, cad = initialize_problem(P, dev)
    flow.enforce_reality_condition(prob)
    k0 = math.sqrt(P.ω0 ** 2 - P.f ** 2) / P.background_Cg
    rank, world = shard
    lo, hi = rank * P.Npackets // world, (rank + 1) * P.Npackets // world
    packets = raytracing.generate_initial_wavepackets(prob, P.L, k0, hi - lo, P.sqrtNpackets, P.f, P.packet_Cg,
                                                      nsub=P.nsub, first=lo)
    packet_output = SequencedOutput(P.packet_base_filename, P.packet_max_writes)
    for key in ("f0", "Cg", "dt", "N", "k0", "ωsign"):            # savepacketproblem! :87-94
        packet_output["params/" + key] = None
    frames: list = []
    raytracing.get_velocity_info(prob, 0)
    savepacketdata(packet_output, frames, prob, packets, 0, P.write_gradients)
    nframes = round(cad["nsteps"] / cad["packet_output_freq"])
    if max_frames is not None:
        nframes = min(nframes, max_frames)
    t = prob.clock.t
    for _ in range(nframes + 1):
        if prob.clock.step < cad["packet_spinup_step"]:
            flow.stepforward(prob, (), cad["packet_output_freq"] * cad["output_per_packet_freq"])
            raytracing.get_velocity_info(prob, 0)
            t = prob.clock.t
        else:
            for _ in range(cad["packet_output_freq"]):
                t = coupled_step(prob, packets, t)
            savepacketdata(packet_output, frames, prob, packets, 0, P.write_gradients)
        if flow.has_nan(prob):
            break
    return prob, packets, packet_output, frames
