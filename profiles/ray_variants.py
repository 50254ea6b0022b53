"""A/B of the ray kernels on the bench workload (RSW nx^2, sq^2 packets at uniformly random positions): per kernel selected
with swrt_packets_set_kernel, the average launch time from the library's CUDA-event profile over 32 coupled steps (two sort
periods) and a checksum (cached and tile are bit-identical, tile3 agrees to rounding).  `python profiles/ray_variants.py [cached tile ...]`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, raytracing  # noqa: E402

nx = int(os.environ.get("NX", 2048))
sq = int(os.environ.get("SQ", 4096))
lattice = int(os.environ.get("LATTICE", 0))
sort_every = int(os.environ.get("SORT_EVERY", 16))
names = sys.argv[1:] or ["cached", "tile", "tile3", "pipe"]
P = drivers.Parameters(nx=nx, sqrtNpackets=sq)
prob, _ = drivers.initialize_problem(P)
for name in names:
    pk = raytracing.generate_initial_wavepackets(prob, P.L, 5.196, P.Npackets, P.sqrtNpackets, P.f, P.Cg, sort_every=sort_every)
    pk.set_kernel({"cached": raytracing.RAYKERNEL_CACHED, "tile": raytracing.RAYKERNEL_TILE, "tile3": raytracing.RAYKERNEL_TILE3, "pipe": raytracing.RAYKERNEL_PIPE, "auto": raytracing.RAYKERNEL_AUTO}[name])
    if not lattice:
        xk = pk.get()
        xk[:, 0:2] = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(P.Npackets, 2))
        pk.set(xk)
        del xk
    raytracing.get_velocity_info(prob, 0)
    t = prob.clock.t
    for _ in range(4):
        t = drivers.coupled_step(prob, pk, t)
    prob.sync()
    prob.profile(2)
    for _ in range(32):
        t = drivers.coupled_step(prob, pk, t)
    prob.sync()
    rep = prob.profile_report()
    prob.profile(0)
    r = next(v for k, v in rep.items() if k.startswith("raytrace_rk4"))
    print("kernel", name, "sort_every", sort_every, "nx", nx, "packets", P.Npackets, "raytrace ms %.4f" % r["ms_avg"], "sort ms/launch %.4f x %d" %
          (rep["packet_sort_kernels"]["ms_avg"], rep["packet_sort_kernels"]["launches"]), "checksum %.15e" % float(np.abs(pk.get()).sum()), flush=True)
    pk.close()
