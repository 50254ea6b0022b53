"""torchrun script: team mode on config 4 (RSW nx^2 + sq^2 packets) -- per-step time of the native coupled loop, of the flow step alone,
of the band snapshot alone, of the ray trace alone, the cost of one device barrier, and the per-kernel CUDA-event profile of rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 profiles/team_step_time.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juliaraytracingsw_b200 import drivers, raytracing  # noqa: E402
from juliaraytracingsw_b200.slab import SlabProblem  # noqa: E402

nx = int(os.environ.get("NX", 2048))
sq = int(os.environ.get("SQ", 4096))
same = os.environ.get("SWRT_TEAM_SAME_GPU", "0") == "1"
local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P = drivers.Parameters(nx=nx, sqrtNpackets=sq)
prob0, _ = drivers.initialize_problem(P, dev=local)
sol0 = prob0.sol
dt, nu = drivers.timestep_and_viscosity(P)
prob0.close()
sp = SlabProblem(dist, local, barrier="host" if same else None, nx=P.nx, Lx=P.L, dt=dt, f=P.f, Cg=P.Cg, ν=nu, nν=P.nν)
sp.sol = sol0
N = P.Npackets
lo, hi = rank * N // world, (rank + 1) * N // world
pk = raytracing.Packets(sp, hi - lo, P.f, P.Cg, first=lo, overlap=int(os.environ.get("OVERLAP", 1)) != 0)
pk.generate(P.L, 5.196, sq, lo)
if not int(os.environ.get("LATTICE", 0)):
    xk = pk.get()
    xk[:, 0:2] = np.random.default_rng(1000 + rank).uniform(-np.pi, np.pi, size=(hi - lo, 2))
    pk.set(xk)
    del xk
raytracing.get_velocity_info(sp, 0)


def timed(fn, n):
    pk.sync(); sp.sync(); dist.barrier()
    sp.timer_start()
    fn(n)
    pk.sync()                       # (packets on their own stream: the host waits for them before the flow-stream timer stops)
    ms = sp.timer_stop()
    parts = [None] * world
    dist.all_gather_object(parts, ms)
    return max(parts) / n


drivers.coupled_steps(sp, pk, 20)
res = {"ranks": world, "nx": nx, "packets": N, "overlap": int(os.environ.get("OVERLAP", 1)), "slab_mode": sp.mode, "resident": sp._gather(pk.resident())}
res["coupled_ms"] = timed(lambda n: drivers.coupled_steps(sp, pk, n), 64)
res["flow_ms"] = timed(lambda n: sp.stepforward(n), 64)
res["snapshot_ms"] = timed(lambda n: [sp.velocity_snapshot(1, 0) for _ in range(n)], 32)
res["barrier_ms"] = timed(lambda n: [sp.team_barrier() for _ in range(n)], 200)
t0 = sp.clock.t


def trace(n):
    for _ in range(n):
        raytracing.raytrace(pk, None, None, None, None, sp.grid, pk, dt, (0.0, dt))


pk_sort = 16
res["raytrace_ms_incl_sorts"] = timed(trace, 32)
sp.profile(2)
drivers.coupled_steps(sp, pk, 32)
sp.sync()
rep = sp.profile_report()
sp.profile(0)
res["kernels_rank0_ms_per_step"] = {k: round(v["ms_total"] / 32, 5) for k, v in rep.items()}
res["launches_per_step"] = sum(v["launches"] for v in rep.values()) / 32
res["value"] = N / (res["coupled_ms"] * 1e-3)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
pk.close(); sp.close()
dist.destroy_process_group()
