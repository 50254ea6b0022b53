"""torchrun script: flow-only steps/s of the slab-decomposed step (BASELINE config 5 flow: two-layer QG 4096^2, fp64).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 profiles/slab_bench.py [--nx 4096]
With one rank it times the single-GPU path.  Device timing with CUDA events on torch's stream, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import juliaraytracingsw_b200 as swrt  # noqa: E402
from juliaraytracingsw_b200 import flow  # noqa: E402
from juliaraytracingsw_b200.slab import SlabProblem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=4096)
ap.add_argument("--model", default="TwoLayerQG")
ap.add_argument("--no-p2p", action="store_true")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nx = a.nx
nvar = {"TwoLayerQG": 2, "RotatingShallowWater": 3, "SWQG": 1}[a.model]
# swqg/TwoLayerParameters.jl recipe (f=3, Cg=1, ug=0.025, cfltune=0.025, nutune=40, nnu=4)
dt = 0.025 / 0.025 * (2 * np.pi / nx) * 0.025
nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** 8) / dt
kw = dict(model=a.model, nx=nx, dt=dt, nu=nu, nnu=4, f=3.0, Cg=1.0)
if a.model == "TwoLayerQG":
    kw.update(U=0.025, mu=1e-2, f0=3.0)
rng = np.random.default_rng(0)
sol = np.zeros((nx // 2 + 1, nx, nvar), dtype=np.complex128)
sol[1:24, :24] = (rng.standard_normal((23, 24, nvar)) + 1j * rng.standard_normal((23, 24, nvar))) * nx * nx * 1e-3
prob = SlabProblem(dist, local, p2p=not a.no_p2p, **kw) if world > 1 else swrt.Problem(local, **kw)
prob.sol = sol if nvar > 1 else sol[:, :, 0]
step = (lambda n: prob.stepforward(n)) if world > 1 else (lambda n: flow.stepforward(prob, (), n))
step(5)
torch.cuda.synchronize(); prob.sync(); dist.barrier(); torch.cuda.synchronize()
if world > 1:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(a.steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
else:
    prob.timer_start(); step(a.steps); ms = prob.timer_stop()
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"{a.model} {nx}^2 flow step, slab-decomposed over {world} GPU(s)", "n_gpus": world, "steps": a.steps,
                      "ms_per_step": float(t) / a.steps, "steps_per_s": 1e3 * a.steps / float(t)}))
dist.destroy_process_group()
