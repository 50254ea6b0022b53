"""e2e leg of bench.py by row-block count: host-resident packets (pinned), per step flow step + snapshot, upload -> sort + ray trace ->
download through raytracing.PacketPipeline.  Prints ms per step (wall clock around 10 steps, 3 repeats) and the host time spent
enqueueing one step.  `python profiles/e2e_chunks.py 8 16 32 64`.  Not a bench value."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juliaraytracingsw_b200 import drivers, flow, raytracing  # noqa: E402
from juliaraytracingsw_b200._lib import check, lib  # noqa: E402

counts = [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]
P = drivers.Parameters(nx=2048, sqrtNpackets=4096)
prob, _ = drivers.initialize_problem(P)
n = P.Npackets
pin = lambda *shape: torch.empty(shape[::-1], dtype=torch.float64, pin_memory=True).numpy().T
h_xk, h_out = pin(n, 4), pin(n, 4)
rng = np.random.default_rng(1)
h_xk[:, 0:2] = rng.uniform(-np.pi, np.pi, size=(n, 2))
h_xk[:, 2] = 5.196
h_xk[:, 3] = 0.0
h_sign = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
h_sign[:] = np.where(np.arange(n) % 2 == 0, -1.0, 1.0)
raytracing.get_velocity_info(prob, 0)
for nch in counts:
    pipe = raytracing.PacketPipeline(prob, n, P.f, P.packet_Cg, nchunks=nch, nsub=P.nsub)
    t = prob.clock.t
    enq = []

    def step(first=False):
        global t
        flow.stepforward(prob, (), 1)
        raytracing.get_velocity_info(prob, 1)
        t1 = prob.clock.t
        w0 = time.perf_counter()
        for p, (lo, hi) in zip(pipe.chunks, pipe.bounds):
            p.set_async(h_xk[lo:hi], h_sign[lo:hi] if first else None)
            check(lib().swrt_packets_raytrace(p._h, float(t), float(t1)))
            p.get_async(h_out[lo:hi])
        enq.append(time.perf_counter() - w0)
        raytracing.swap_snapshots(prob, alias=False)
        for p in pipe.chunks:
            p.sync()
        t = t1
    step(True)
    step()
    res = []
    for rep in range(3):
        prob.sync()
        w0 = time.perf_counter()
        for _ in range(10):
            step()
        prob.sync()
        res.append((time.perf_counter() - w0) * 100)
    print(f"chunks {nch:4d}: ms/step " + " ".join(f"{r:6.2f}" for r in res) + f"   host enqueue ms/step {1e3 * np.mean(enq[2:]):5.2f}", flush=True)
    if os.environ.get("PHASES"):
        # the phases of one step on their own (10 repetitions each): what the full step overlaps
        def timed(fn):
            prob.sync()
            w0 = time.perf_counter()
            for _ in range(10):
                fn()
                for p in pipe.chunks:
                    p.sync()
            prob.sync()
            return (time.perf_counter() - w0) * 100

        def flow_only():
            flow.stepforward(prob, (), 1)
            raytracing.get_velocity_info(prob, 1)
            raytracing.swap_snapshots(prob, alias=False)

        def uploads():
            for p, (lo, hi) in zip(pipe.chunks, pipe.bounds):
                p.set_async(h_xk[lo:hi], None)

        def kernels():
            for p in pipe.chunks:
                check(lib().swrt_packets_raytrace(p._h, float(t), float(t + prob.dt)))

        def downloads():
            for p, (lo, hi) in zip(pipe.chunks, pipe.bounds):
                p.get_async(h_out[lo:hi])
        print(f"   phases alone, ms: flow step + snapshot {timed(flow_only):5.2f} | uploads (0.54 GB) {timed(uploads):5.2f} | sort + ray trace of the "
              f"{nch} blocks {timed(kernels):5.2f} | unpermute + downloads (0.54 GB) {timed(downloads):5.2f}", flush=True)
    pipe.close()
