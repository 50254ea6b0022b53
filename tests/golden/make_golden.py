"""Generates the committed golden vectors from the oracle (run from the repository root: `python tests/golden/make_golden.py`).

The reference is Julia and cannot run in the build image, so these are outputs of `oracle/` -- the restatement pinned on the
reference's recorded known-answer values in `tests/golden/reference_kats.json` / `tests/test_oracle_kats.py` -- on seeded inputs.
They anchor both the oracle (CPU suite: it must keep reproducing them) and the CUDA path (GPU suite) against silent drift.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import config2_setup, oracle_steps  # noqa: E402
from oracle import raytrace as oray, rsw as orsw  # noqa: E402


def main():
    g, p, sol0, c = config2_setup(64)
    sol10 = oracle_steps(g, p, sol0, c["dt"], 10)
    sol13 = oracle_steps(g, p, sol0, c["dt"], 13)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol10, g, p), g)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(sol13, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 16)
    xk[:, 0:2] += np.random.default_rng(11).uniform(-10, 10, size=(xk.shape[0], 2))
    t0, t1 = 10 * c["dt"], 13 * c["dt"]
    xk1 = oray.raytrace(xk.copy(), sign, t0, t1, Fo, Fn, g, c["f"], c["Cg"], nsub=3)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rsw64_config2.npz"),
                        sol0=sol0, sol10=sol10, sol13=sol13, snapshot10=Fo, xk0=xk, sign=sign, xk1=xk1,
                        params=np.array([c["L"], c["dt"], c["f"], c["Cg"], c["nu"], c["nnu"], c["k0"]]))


if __name__ == "__main__":
    main()
