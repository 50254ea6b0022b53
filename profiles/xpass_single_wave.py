"""x-pass latency in the single-wave regime of team mode, measured on ONE GPU: an RSW flow on an nx x ny grid with ny = 2048 / P rows
(what one of P ranks owns) steps through the plain (non-slab) kernels; per-kernel CUDA-event times.  `python profiles/xpass_single_wave.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import juliaraytracingsw_b200 as swrt  # noqa: E402
from juliaraytracingsw_b200 import flow  # noqa: E402

for ny in (2048, 1024, 512, 256):
    nx = 2048
    prob = swrt.Problem(nx=nx, ny=ny, Lx=2 * np.pi, Ly=2 * np.pi * ny / nx, dt=1e-4, f=3.0, Cg=1.0, nu=1e-20, nnu=4)
    rng = np.random.default_rng(0)
    sol = np.zeros((nx // 2 + 1, ny, 3), dtype=np.complex128)
    sol[1:20, 1:20] = 1e-3 * nx * ny * (rng.standard_normal((19, 19, 3)) + 1j * rng.standard_normal((19, 19, 3)))
    prob.sol = sol
    flow.stepforward(prob, (), 5)
    prob.sync()
    prob.profile(2)
    flow.stepforward(prob, (), 20)
    prob.sync()
    rep = prob.profile_report()
    prob.profile(0)
    print(f"nx {nx} ny {ny}: " + ", ".join(f"{k.split('<')[0].replace('_kernel', '')}<{k.split('<')[1] if '<' in k else ''} {v['ms_avg'] * 1e3:.1f} us" for k, v in rep.items()), flush=True)
    prob.close()
