"""torchrun script: per-phase device time of the slab-decomposed step (stage kernels vs all-to-all)."""
import argparse, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juliaraytracingsw_b200._lib import check, lib
from juliaraytracingsw_b200.slab import SlabProblem, A_RECV, A_SEND, B_RECV, B_SEND

ap = argparse.ArgumentParser(); ap.add_argument("--nx", type=int, default=4096); ap.add_argument("--model", default="TwoLayerQG"); ap.add_argument("--no-p2p", action="store_true")
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nx = a.nx; nvar = {"TwoLayerQG": 2, "RotatingShallowWater": 3}[a.model]
dt = 0.025 * (2 * np.pi / nx); nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** 8) / dt
kw = dict(model=a.model, nx=nx, dt=dt, nu=nu, nnu=4, f=3.0, Cg=1.0)
if a.model == "TwoLayerQG": kw.update(U=0.025, mu=1e-2, f0=3.0)
rng = np.random.default_rng(0)
sol = np.zeros((nx // 2 + 1, nx, nvar), dtype=np.complex128)
sol[1:24, :24] = (rng.standard_normal((23, 24, nvar)) + 1j * rng.standard_normal((23, 24, nvar))) * nx * nx * 1e-3
p = SlabProblem(dist, local, p2p=not a.no_p2p, **kw); p.sol = sol
p.stepforward(5)
L = lib(); names = ["stage_a", "a2a_A", "stage_b", "a2a_B", "stage_c"]; acc = dict.fromkeys(names, 0.0)
def timed(name, fn):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); acc[name] += e0.elapsed_time(e1)
n = 10
for _ in range(n):
    timed("stage_a", lambda: check(L.swrt_slab_stage_a(p._h)))
    timed("a2a_A", lambda: p._a2a(A_RECV, A_SEND, p.njobs_a))
    timed("stage_b", lambda: check(L.swrt_slab_stage_b(p._h)))
    timed("a2a_B", lambda: p._a2a(B_RECV, B_SEND, p.njobs_b))
    timed("stage_c", lambda: check(L.swrt_slab_stage_c(p._h)))
bytes_a = world * p.njobs_a * p.yrows * p.chunk * 16; bytes_b = world * p.njobs_b * p.yrows * p.chunk * 16
if rank == 0:
    print(json.dumps({"world": world, "nx": nx, "model": a.model, "ms": {k: round(v / n, 4) for k, v in acc.items()},
                      "a2a_A_MB_per_rank": bytes_a / 1e6, "a2a_B_MB_per_rank": bytes_b / 1e6}))
dist.destroy_process_group()
