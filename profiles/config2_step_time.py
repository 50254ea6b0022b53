"""BASELINE config 2 (RSW 512^2 + 65,536 packets) coupled step: Python per-call loop vs the fused swrt_packets_coupled_steps entry."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, raytracing  # noqa: E402

nx, sq = int(os.environ.get("NX", 512)), int(os.environ.get("SQ", 256))
P = drivers.Parameters(nx=nx, sqrtNpackets=sq)
prob, _ = drivers.initialize_problem(P)
k0 = (P.ω0 ** 2 - P.f ** 2) ** 0.5 / P.background_Cg
pk = raytracing.generate_initial_wavepackets(prob, P.L, k0, P.Npackets, P.sqrtNpackets, P.f, P.packet_Cg)
raytracing.get_velocity_info(prob, 0)
t = prob.clock.t
for _ in range(20):
    t = drivers.coupled_step(prob, pk, t)
prob.sync()
n = 300
w0 = time.perf_counter()
for _ in range(n):
    t = drivers.coupled_step(prob, pk, t)
prob.sync()
loop_ms = (time.perf_counter() - w0) * 1e3 / n
drivers.coupled_steps(prob, pk, 60)      # captures and instantiates the two six-step graphs (a few ms each, once)
prob.sync()
n = 1200
w0 = time.perf_counter()
drivers.coupled_steps(prob, pk, n)
prob.sync()
fused_ms = (time.perf_counter() - w0) * 1e3 / n
print("nx", nx, "packets", P.Npackets, "per-call loop %.4f ms/step" % loop_ms, "fused %.4f ms/step" % fused_ms,
      "packet-steps/s %.3e" % (P.Npackets / (fused_ms * 1e-3)))
