"""Host-side mirror of the CPU two-layer packet driver, raytracing/TwoLayerRaytracing.jl (BASELINE config 1).

`set_up_problem` :162-182 (MultiLayerQG.Problem with aliased_fraction = 0, state from a streamfunction snapshot),
`generate_initial_wavepackets` :10-22, `simulate!` :66-160 (flow `nsubs` steps -> layer-mean streamfunction -> velocity
info -> `Raytracing.solve!` -> k-cutoff reset -> old = new) and the parameters of raytracing/CPUParameters.jl.  The packet
integrator is the CPU tracer's: quadratic B-spline fields, linear in time, implicit midpoint with step `dt`
(raytracing/Raytracing.jl:91-170).  Everything numerical runs on the device through libswrt.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import flow, raytracing
from .drivers import Frame


@dataclass
class Parameters:
    """raytracing/CPUParameters.jl (module Parameters) + the flow parameters the IC file carries
    (simulation/Parameters.jl:33-49)."""
    stepper: str = "FilteredAB3"
    total_time: float = 8000.0
    L: float = 2 * np.pi
    nsubs: int = 1
    npacketsubs: int = 50
    max_writes: int = 1000
    packetSpinUpDelay: int = 0
    sqrtNpackets: int = 20
    corFactor: float = 2.0
    k_cutoff: float = 60.0
    packetStepsPerBackgroundStep: int = 1
    # simulation/Parameters.jl
    nx: int = 512
    f0: float = 1.0
    deformation_radius: float = 1 / 15
    intervortex_radius: float = 1 / 2
    avg_U: float = 0.1
    H0: float = 1.0
    b2: float = 1.0
    β: float = 0.0
    nν: int = 1
    ν: float = 0.0
    # ours
    interp: int = raytracing.INTERP_BSPLINE2
    integrator: int = raytracing.INTEG_IMPLICIT_MIDPOINT

    @property
    def Npackets(self):
        return self.sqrtNpackets ** 2

    def compute_parameters(self):
        """simulation/Parameters.jl:6-24 -> (μ, b1, U)."""
        l_star = self.intervortex_radius / self.deformation_radius
        kappa_star = 0.36 / math.log(l_star / 3.2)
        U = self.avg_U / l_star
        μ = 2 * U * kappa_star / self.deformation_radius
        b1 = 4 * self.f0 ** 2 * self.deformation_radius ** 2 / self.H0 + self.b2
        return μ, b1, U

    @property
    def dt(self):
        return 0.02 * (self.L / self.nx) / self.avg_U


def generate_initial_wavepackets(L, k0, Npackets, sqrtNpackets):
    """raytracing/TwoLayerRaytracing.jl:10-22 as (N, 4) rows x, y, k, l (row r-1 = packet (i-1) s + j)."""
    s = int(sqrtNpackets)
    r = np.arange(1, s * s + 1)
    i, j = (r - 1) // s + 1, (r - 1) % s + 1
    offset = L / s / 2
    ang = 2 * np.pi * r / Npackets
    return np.stack([i * L / s - L / 2 - offset, j * L / s - L / 2 - offset, k0 * np.cos(ang), k0 * np.sin(ang)], axis=1)


def set_up_problem(P: Parameters, ψh=None, qh=None, dev=0):
    """`set_up_problem` :162-182: ψh is the (nkr, nl, 2) streamfunction of the IC file (`pvfromstreamfunction!`), or pass qh."""
    μ, b1, U = P.compute_parameters()
    prob = flow.Problem(dev, model="MultiLayerQG", stepper=P.stepper, nx=P.nx, Lx=P.L, f0=P.f0, H=(P.H0 / 2, P.H0 / 2),
                        b=(b1, P.b2), U=(U, -U), μ=μ, β=P.β, dt=P.dt, ν=P.ν, nν=P.nν, aliased_fraction=0)
    if qh is None:
        K2 = prob.grid.Krsq[:, :, None]
        F = prob.F
        ψh = np.asarray(ψh, dtype=np.complex128)
        qh = -K2 * ψh + F * (ψh[:, :, ::-1] - ψh)          # q_j = lap psi_j + F (psi_other - psi_j)
    prob.sol = qh
    return prob


def savepackets(out, frames, prob, packets, step):
    """`savepackets!` :24-32: p/t, and per packet p/x, p/k, p/u (here one array per frame)."""
    xk = packets.get()
    U = raytracing.interpolate_velocity(raytracing.Velocity(prob, 0), packets)
    fr = Frame(step, prob.clock.t, xk[:, 0:2], xk[:, 2:4], U)
    if out is not None:
        out[f"p/t/{step}"] = fr.t
        out[f"p/x/{step}"], out[f"p/k/{step}"], out[f"p/u/{step}"] = fr.x, fr.k, fr.u
    frames.append(fr)


def start(P: Parameters, ψh=None, qh=None, dev=0, max_frames=None, out=None):
    """`start!` :184-221 + `simulate!` :66-160.  Returns (prob, packets, frames, step): `step` is the driver's own counter --
    the reference bumps `clock.step` once more per packet frame and per output frame (:146,:148), which only labels keys."""
    prob = set_up_problem(P, ψh, qh, dev)
    raytracing.set_interpolation(prob, P.interp)
    f, Cg = P.f0, 1.0
    k0 = math.sqrt(P.corFactor ** 2 - 1) * f / Cg
    pdt = prob.dt / P.packetStepsPerBackgroundStep
    nsub = max(1, round(P.nsubs * prob.dt / pdt))
    packets = raytracing.Packets(prob, P.Npackets, f, Cg, nsub=nsub, interp=P.interp, integrator=P.integrator)
    packets.set(generate_initial_wavepackets(P.L, k0, P.Npackets, P.sqrtNpackets), np.ones(P.Npackets))
    nsteps = int(math.ceil(P.total_time / prob.dt))
    frames: list = []
    step = prob.clock.step
    raytracing.get_velocity_info(prob, 0, raytracing.PSI_TWOLAYER_MEAN)
    savepackets(out, frames, prob, packets, step)
    nframes = round(nsteps / P.npacketsubs) + 1
    if max_frames is not None:
        nframes = min(nframes, max_frames)
    old_t = prob.clock.t
    for _ in range(nframes):
        for _ in range(round(P.npacketsubs / P.nsubs)):
            flow.stepforward(prob, (), P.nsubs)
            new_velocity, new_grad = raytracing.get_velocity_info(prob, 1, raytracing.PSI_TWOLAYER_MEAN)
            new_t = prob.clock.t
            raytracing.raytrace(packets, None, new_velocity, None, new_grad, prob.grid, packets, pdt, (old_t, new_t))
            packets.kcutoff_reset(P.k_cutoff, k0)
            raytracing.swap_snapshots(prob, alias=False)
            old_t = new_t
            step += P.nsubs + 1
        step += 1
        savepackets(out, frames, prob, packets, step)
    return prob, packets, frames, step
