"""Flow-only step time (CUDA events on the handle's stream): `SWRT_PDL=0|1 python profiles/flow_step_time.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, flow  # noqa: E402

nx = int(os.environ.get("NX", 2048))
P = drivers.Parameters(nx=nx, sqrtNpackets=64)
prob, _ = drivers.initialize_problem(P)
flow.stepforward(prob, (), 20)
prob.sync()
best = 1e9
for _ in range(5):
    prob.timer_start()
    flow.stepforward(prob, (), 200)
    best = min(best, prob.timer_stop() / 200)
F = 8.0 * nx * nx
print("nx", nx, "PDL", os.environ.get("SWRT_PDL", "1"), "ms/step %.5f" % best, "42F frac %.4f" % (42 * F / (best * 1e-3) / 6459.9e9),
      "KE %.15e" % flow.kinetic_energy(prob))
