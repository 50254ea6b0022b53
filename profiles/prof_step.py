"""Short profiling driver: RSW nx^2 coupled steps with randomised packet positions (worst-case gathers).
Used under `ncu` (one GPU); prints nothing that is a bench value."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, raytracing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=2048)
ap.add_argument("--sqrt-packets", type=int, default=2048)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--lattice", action="store_true")
ap.add_argument("--kernel", default="auto", choices=["auto", "cached", "tile", "tile3", "pipe"])
a = ap.parse_args()
P = drivers.Parameters(nx=a.nx, sqrtNpackets=a.sqrt_packets)
prob, _ = drivers.initialize_problem(P)
pk = raytracing.generate_initial_wavepackets(prob, P.L, 5.196, P.Npackets, P.sqrtNpackets, P.f, P.Cg)
pk.set_kernel({"auto": -1, "cached": 0, "tile": 1, "tile3": 2, "pipe": 3}[a.kernel])
if not a.lattice:
    xk = pk.get()
    xk[:, 0:2] = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(P.Npackets, 2))
    pk.set(xk)
raytracing.get_velocity_info(prob, 0)
t = 0.0
for _ in range(a.steps):
    t = drivers.coupled_step(prob, pk, t)
prob.sync()
print("done", prob.launch_count())
