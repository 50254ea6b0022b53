"""Time the ray kernel variants selected by SWRT_RAYTRACE_CACHE (read once per process): run as
`SWRT_RAYTRACE_CACHE=k python profiles/ray_variants.py` -- prints the average launch time from the library's CUDA-event profile."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, raytracing  # noqa: E402

nx = int(os.environ.get("NX", 2048))
sq = int(os.environ.get("SQ", 4096))
P = drivers.Parameters(nx=nx, sqrtNpackets=sq)
prob, _ = drivers.initialize_problem(P)
pk = raytracing.generate_initial_wavepackets(prob, P.L, 5.196, P.Npackets, P.sqrtNpackets, P.f, P.Cg)
xk = pk.get()
xk[:, 0:2] = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(P.Npackets, 2))
pk.set(xk)
raytracing.get_velocity_info(prob, 0)
t = 0.0
for _ in range(4):
    t = drivers.coupled_step(prob, pk, t)
prob.sync()
prob.profile(2)
for _ in range(20):
    t = drivers.coupled_step(prob, pk, t)
prob.sync()
rep = prob.profile_report()
r = rep["raytrace_rk4_kernel"]
print("variant", os.environ.get("SWRT_RAYTRACE_CACHE", "default"), "sgrid", os.environ.get("SWRT_RAYTRACE_SGRID", "1"),
      "raytrace ms %.4f" % r["ms_avg"], "checksum %.15e" % float(np.abs(pk.get()).sum()))
