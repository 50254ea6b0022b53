"""torchrun script: the slab-decomposed flow step over all ranks against the single-GPU step on every rank's own GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tests/multigpu/slab_parity.py
Exits non-zero on any mismatch.  (The single-GPU path itself is checked against the oracle by tests/test_gpu_parity.py.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import juliaraytracingsw_b200 as swrt  # noqa: E402
from juliaraytracingsw_b200 import flow, raytracing  # noqa: E402
from juliaraytracingsw_b200.slab import SlabProblem  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def smooth_state(nx, nvar, seed):
    rng = np.random.default_rng(seed)
    k = np.arange(nx // 2 + 1)[:, None]
    l = (np.fft.fftfreq(nx) * nx)[None, :]
    sol = np.empty((nx // 2 + 1, nx, nvar), dtype=np.complex128)
    for v in range(nvar):
        f = rng.standard_normal((nx, nx))
        fh = np.fft.fft(np.fft.rfft(f, axis=0), axis=1) / (1.0 + k * k + l * l) ** 1.5
        sol[:, :, v] = fh * (0.2 * nx * nx / np.abs(fh).max() / 50)
    return sol


def main():
    same = os.environ.get("SWRT_TEAM_SAME_GPU", "0") == "1"          # every rank on cuda:0 (single-GPU box): gloo + host barrier
    local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    barrier = "host" if same else "device"
    worst = 0.0
    cases = [("RotatingShallowWater", 3, dict(f=3.0, Cg=1.0), raytracing.PSI_RSW_BALANCED),
             ("TwoLayerQG", 2, dict(U=0.5, mu=1e-2, f0=3.0, Cg=1.0), raytracing.PSI_TWOLAYER_BAROCLINIC),
             ("SWQG", 1, dict(f=3.0, Cg=1.0), raytracing.PSI_SWQG)]
    for nx, p2p in ((128, True), (512, True)) + (() if same else ((256, False),)):      # (False: NCCL all_to_all_single between the phases)
        for model, nvar, kw, psi in cases:
            dt = 0.05 * 2 * np.pi / nx
            nu = 2 * np.pi / nx / ((nx / 2 - 1) ** 8) / dt
            sol0 = smooth_state(nx, nvar, 7 + nx)
            ref = swrt.Problem(local, model=model, nx=nx, dt=dt, nu=nu, nnu=4, **kw)
            sp = SlabProblem(dist, local, p2p=p2p, barrier=barrier, model=model, nx=nx, dt=dt, nu=nu, nnu=4, **kw)
            assert sp.p2p == p2p
            ref.sol = sol0 if nvar > 1 else sol0[:, :, 0]
            sp.sol = sol0 if nvar > 1 else sol0[:, :, 0]
            for n in (1, 3, 8):
                flow.stepforward(ref, (), n)
                sp.stepforward(n)
                e = rel(sp.gather_solution(), ref.sol)
                worst = max(worst, e)
                assert e < 1e-12, (model, nx, n, e)
            if p2p:
                sp.velocity_snapshot(1, psi)
                vel, _ = raytracing.get_velocity_info(ref, 1, psi)
                e = rel(raytracing.Velocity(sp, 1)._arr(), vel._arr())
                worst = max(worst, e)
                assert e < 1e-12, (model, nx, "snapshot", e)
            ke, pe = sp.energies()
            ke_r = flow.kinetic_energy(ref)
            ke_r = sum(ke_r) if isinstance(ke_r, tuple) else ke_r
            assert abs(ke / ke_r - 1) < 1e-11 and abs(pe / flow.potential_energy(ref) - 1) < 1e-11, (model, ke, ke_r)
            assert sp.clock.step == ref.clock.step == 12
            sp.close(); ref.close()
    allw = [None] * world
    dist.all_gather_object(allw, worst)
    if rank == 0:
        print(f"slab parity ok on {world} ranks: worst relative L2 {max(allw):.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
