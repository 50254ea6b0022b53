import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from juliaraytracingsw_b200 import drivers, raytracing
P = drivers.Parameters(nx=2048, sqrtNpackets=4096)
prob, _ = drivers.initialize_problem(P)
for se in (16, 4, 64):
    pk = raytracing.generate_initial_wavepackets(prob, P.L, 5.196, P.Npackets, P.sqrtNpackets, P.f, P.Cg, sort_every=se)
    xk = pk.get(); xk[:, 0:2] = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(P.Npackets, 2)); pk.set(xk); del xk
    raytracing.get_velocity_info(prob, 0)
    t = prob.clock.t; ts = []
    for s in range(40):
        prob.profile(2)
        t = drivers.coupled_step(prob, pk, t)
        r = prob.profile_report()
        ts.append((round(next(v for k, v in r.items() if k.startswith("raytrace_rk4"))["ms_avg"], 3), round(r.get("packet_sort_kernels", {"ms_total": 0})["ms_total"], 3)))
    print("sort_every", se, ts)
    pk.close()
