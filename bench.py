#!/usr/bin/env python
"""Benchmark of the coupled hot path: RSW 2048^2 IFMAB3 flow step + velocity snapshot + RK4 ray tracing
of 16,777,216 wave packets (BASELINE.json config 4), one process per GPU, packets sharded in contiguous
blocks over the ranks with a replicated flow (no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl swrt|reference]

One JSON line on stdout (rank 0).  `value` = packet-steps/s with everything resident in HBM; `e2e` = the same
step through the public API with pinned HOST packet buffers copied in and the output frame copied out every
step; `roofline` = the dominant kernel; `spectral_step` = the flow-only step against the 42 F contract of
SURVEY.md section 8(d); `cpu_baseline` = the NumPy/SciPy oracle on the host cores (bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="swrt", choices=["swrt", "reference"])
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--sqrt-packets", type=int, default=4096)
    ap.add_argument("--nsub", type=int, default=1)
    ap.add_argument("--workload", default="config4", choices=["config4", "config5"],
                    help="config4 (default): RSW 2048^2 + 16.8M packets, flow replicated; config5: two-layer QG 4096^2 slab-decomposed "
                         "over the ranks + 67M packets (BASELINE.json configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=64)
    ap.add_argument("--parallel", default="team", choices=["team", "replicated"],
                    help="N > 1: `team` (default) = flow slab-decomposed over the ranks + packets sharded by y-band, all native over CUDA "
                         "IPC; `replicated` = round-1 scheme (every rank repeats the flow step, packets sharded by index)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config2 / config3 / config5 extra keys")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


L2_NOTE = "inputs_exceed_l2 (2 x 168 MB snapshot fields, 0.67 GB packets, 0.47 GB spectral work set vs 126 MB L2 at the default sizes)"
WORKLOAD = "RSW {nx}^2 IFMAB3 flow step + velocity snapshot + RK4 ray tracing of {n} wave packets (BASELINE config 4)"


# ----------------------------------------------------------------------------------------------- CPU / reference arm
def cpu_sample(args, steps, warmup, cores):
    """The oracle (NumPy/SciPy restatement of the reference algorithm; the reference itself needs Julia, absent
    here) on the host cores: the full 2048^2 flow step + velocity info, and RK4 ray tracing of a 2^18-packet sample
    threaded over all cores; the packet cost is scaled to the full packet count."""
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import config2_setup
    from oracle import craytrace, ifmab3, raytrace as oray, rsw as orsw

    nx, ntot = args.nx, args.sqrt_packets ** 2
    use_c = craytrace.available()        # compiled, OpenMP-threaded restatement (bit-identical to the NumPy tracer): all packets
    # the thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers) and what OpenMP grants is recorded
    omp_threads = craytrace.threads_used(cores) if use_c else cores
    assert omp_threads == cores, (omp_threads, cores)
    nsample = ntot if use_c else min(ntot, 1 << 18)
    g, p, sol, c = config2_setup(nx)
    ts = ifmab3.IFMAB3(np.zeros((1, 1, 3, 3)), c["dt"], lambda s: orsw.calcN(s, g, p))
    ts.expLdt = ifmab3.expL_closed_form(g, p, c["dt"])
    ts.exp2Ldt = ifmab3.expL_closed_form(g, p, 2 * c["dt"])
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], int(round(nsample ** 0.5)))
    nsample = xk.shape[0]
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol, g, p), g)
    chunks = np.array_split(np.arange(nsample), cores)
    pool = ThreadPoolExecutor(cores)
    t_flow = t_pk = 0.0
    told = 0.0
    for it in range(warmup + steps):
        a = time.perf_counter()
        ts.stepforward(sol)
        Fn = oray.get_velocity_info(orsw.get_streamfunction(sol, g, p), g)
        b = time.perf_counter()
        tnew = told + c["dt"]

        def work(idx):
            z = xk[idx].copy()
            oray.raytrace(z, sign[idx], told, tnew, Fo, Fn, g, c["f"], c["Cg"], nsub=args.nsub)
            xk[idx] = z
        if use_c:
            craytrace.raytrace(xk, sign, told, tnew, Fo, Fn, g, c["f"], c["Cg"], nsub=args.nsub, threads=cores)
        else:
            list(pool.map(work, chunks))
        cend = time.perf_counter()
        Fo, told = Fn, tnew
        if it >= warmup:
            t_flow += b - a
            t_pk += cend - b
    pool.shutdown()
    t_step = t_flow / steps + (t_pk / steps) * (ntot / nsample)
    return dict(value=ntot / t_step, ms_per_step=1e3 * t_step, flow_ms=1e3 * t_flow / steps, omp_threads=omp_threads,
                packet_ms_sample=1e3 * t_pk / steps, nsample=nsample,
                sample=(f"per step: full {nx}^2 oracle flow step + velocity info (NumPy + threaded scipy.fft), RK4 of {nsample} packets "
                        + ("with the C/OpenMP restatement of the oracle tracer" if use_c else "with the NumPy tracer")
                        + f" on {cores} threads" + ("" if nsample == ntot else f" scaled x{ntot / nsample:.0f} to {ntot} packets")
                        + f"; {steps} steps after {warmup} warm-up"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, min(args.steps, 30)), max(3, min(args.warmup, 5))   # a step costs ~1 s on the host: bounded; warm-up >= 3 like the GPU arm
    r = cpu_sample(args, steps, warm, cores)
    ntot = args.sqrt_packets ** 2
    line = {
        "impl": "reference", "metric": "packet-steps/s", "value": r["value"], "unit": "packet-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(nx=args.nx, n=ntot), "nx": args.nx, "packets": ntot, "nsub": args.nsub,
                   "integrator": "RK4", "interp": "bilinear", "l2": L2_NOTE},
        "cpu_baseline": {"value": r["value"], "unit": "packet-steps/s", "cores": cores, "omp_threads": r["omp_threads"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "packet-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm (oracle/); the reference itself is Julia and cannot run here",
    }
    print(json.dumps(line), flush=True)



def gpu_local_cpus(dev):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- measured-elsewhere evidence
def load_ncu_traffic():
    """DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum) of the kernels this bench names, from the
    committed profiles/ncu_traffic.json (written by profiles/summarize_ncu.py from a `ncu --set full` capture; each entry says
    which capture and which commit it came from).  Absent file or kernel -> None: nothing is hard-coded here."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------- the other BASELINE configs
def extra_config2(dev):
    """BASELINE config 2: RSW 512^2 + 65 536 packets (rsw/RSWRaytracingMain), the driver's hot loop in one library call per 120 steps
    (six-step CUDA graphs between the cell sorts)."""
    from juliaraytracingsw_b200 import drivers, raytracing
    P = drivers.Parameters(nx=512, sqrtNpackets=256)
    prob, _ = drivers.initialize_problem(P, dev=dev)
    k0 = (P.ω0 ** 2 - P.f ** 2) ** 0.5 / P.background_Cg
    pk = raytracing.generate_initial_wavepackets(prob, P.L, k0, P.Npackets, P.sqrtNpackets, P.f, P.packet_Cg)
    raytracing.get_velocity_info(prob, 0)
    drivers.coupled_steps(prob, pk, 120)
    prob.sync()
    n = 1200
    l0 = prob.launch_count()
    prob.timer_start()
    drivers.coupled_steps(prob, pk, n)
    ms = prob.timer_stop()
    out = {"workload": "RSW 512^2 IFMAB3 + 65536 packets, swrt_packets_coupled_steps (BASELINE config 2)", "steps": n, "ms_per_step": ms / n,
           "value": P.Npackets * n / (ms * 1e-3), "unit": "packet-steps/s", "flow_steps_per_s": 1e3 * n / ms,
           "gpu_launches_per_step": (prob.launch_count() - l0) / n}
    pk.close(); prob.close()
    return out


def extra_config3(dev):
    """BASELINE config 3: Thomas-Yamada 1024^2, Lx = 6 pi, ETDRK4 (thomasyamada/gpu-setup/Parameters.jl), free evolution with the
    k-omega accumulator appending one frame per step (thomasyamada/TY_k_omega.jl as a streaming device-side series)."""
    import juliaraytracingsw_b200 as swrt
    from juliaraytracingsw_b200 import flow, komega
    nx, Lx, dt, nnu = 1024, 6 * np.pi, 5e-3, 8
    nu = 5e-34 * (Lx / (2 * np.pi)) ** 16
    prob = swrt.Problem(dev, model="ThomasYamada", stepper="ETDRK4", nx=nx, Lx=Lx, dt=dt, nu=nu, nnu=nnu, Ro=1.0)
    rng = np.random.default_rng(5678)
    g = prob.grid
    sol = np.zeros((g.nkr, g.nl, 4), dtype=np.complex128)
    band = (g.Krsq > 0) & (g.Krsq <= (13 / 3) ** 2)
    for v in range(4):
        ph = np.exp(2j * np.pi * rng.random((g.nkr, g.nl)))
        sol[:, :, v][band] = (0.05 * nx * nx / 40.0 * ph)[band]
    sol[0] = 0
    prob.sol = sol
    flow.stepforward(prob, (), 5)
    prob.sync()
    n = 40
    prob.timer_start()
    flow.stepforward(prob, (), n)
    ms = prob.timer_stop()
    kw = komega.KOmega(prob, k_idx=9, max_frames=n + 4)
    kw.append()
    prob.sync()
    prob.timer_start()
    for _ in range(n):
        flow.stepforward(prob, (), 1)
        kw.append()
    ms_kw = prob.timer_stop()
    F = 8.0 * nx * nx
    out = {"workload": "Thomas-Yamada 1024^2 ETDRK4 free evolution (BASELINE config 3)", "steps": n, "ms_per_step": ms / n, "steps_per_s": 1e3 * n / ms,
           "with_komega_frame_every_step": {"ms_per_step": ms_kw / n, "steps_per_s": 1e3 * n / ms_kw, "append_ms": (ms_kw - ms) / n},
           "algorithmic_bytes_per_step": 268 * F, "achieved_gbs": 268 * F / (ms / n * 1e-3) / 1e9,
           "contract": "B_step ~ 268 F (SURVEY.md 8d, 4 calcN! of 20 min. transforms + 11 pointwise + 16 stepper streams each)",
           "finite": bool(np.isfinite(flow.kinetic_energy(prob)))}
    kw.close(); prob.close()
    return out


def extra_config5(dist, local, world, steps):
    """BASELINE config 5: two-layer QG 4096^2 slab-decomposed over the ranks + 8192^2 packets sharded by y-band (team mode)."""
    from juliaraytracingsw_b200 import drivers, raytracing
    from juliaraytracingsw_b200.slab import SlabProblem
    nx, sq = 4096, 8192
    f, Cg, ug = 3.0, 1.0, 0.025
    dt = 0.025 * (2 * np.pi / nx)
    nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** 8) / dt
    prob = SlabProblem(dist, local, model="TwoLayerQG", nx=nx, dt=dt, nu=nu, nnu=4, f=f, Cg=Cg, U=ug, mu=1e-2, f0=f)
    rng = np.random.default_rng(0)
    sol = np.zeros((nx // 2 + 1, nx, 2), dtype=np.complex128)
    sol[1:24, :24] = (rng.standard_normal((23, 24, 2)) + 1j * rng.standard_normal((23, 24, 2))) * nx * nx * 1e-3
    prob.sol = sol
    del sol
    ntot = sq * sq
    rank = dist.get_rank()
    lo, hi = rank * ntot // world, (rank + 1) * ntot // world
    k0 = (3.0 ** 0.5) * f / Cg
    pk = raytracing.generate_initial_wavepackets(prob, 2 * np.pi, k0, hi - lo, sq, f, Cg, first=lo)
    psi = raytracing.PSI_TWOLAYER_BAROCLINIC
    raytracing.get_velocity_info(prob, 0, psi)
    drivers.coupled_steps(prob, pk, 3, psi_kind=psi)
    prob.sync(); dist.barrier()
    prob.timer_start()
    drivers.coupled_steps(prob, pk, steps, psi_kind=psi)
    ms = prob.timer_stop()
    prob.sync(); dist.barrier()
    prob.timer_start()
    prob.stepforward(steps)
    ms_flow = prob.timer_stop()
    both = [None] * world
    dist.all_gather_object(both, (ms, ms_flow))
    ms, ms_flow = max(b[0] for b in both), max(b[1] for b in both)
    out = {"workload": f"two-layer QG 4096^2 IFMAB3 slab-decomposed over {world} GPUs (NVLink peer stores + device barrier) + RK4 ray tracing of "
                       f"{ntot} packets sharded by y-band (BASELINE config 5)", "n_gpus": world, "steps": steps, "ms_per_step": ms / steps,
           "value": ntot * steps / (ms * 1e-3), "unit": "packet-steps/s", "flow_only_ms_per_step": ms_flow / steps,
           "flow_steps_per_s": 1e3 * steps / ms_flow, "packet_positions": "reference lattice"}
    pk.close(); prob.close()
    return out

# ----------------------------------------------------------------------------------------------- GPU arm
def run_swrt(args):
    import torch

    import juliaraytracingsw_b200 as swrt
    from juliaraytracingsw_b200 import drivers, flow, raytracing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl swrt needs a CUDA device (no CPU fallback)")
    # SWRT_TEAM_SAME_GPU=1: functional check of the N > 1 path on a one-GPU box -- all ranks on cuda:0, gloo, host team barrier
    same_gpu = os.environ.get("SWRT_TEAM_SAME_GPU", "0") == "1"
    if same_gpu:
        local = 0
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if same_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        if same_gpu:
            parts = [None] * world
            dist.all_gather_object(parts, float(x))
            return max(parts)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    P = drivers.Parameters(nx=args.nx, sqrtNpackets=args.sqrt_packets, nsub=args.nsub)
    ntot = P.Npackets
    lo, hi = rank * ntot // world, (rank + 1) * ntot // world
    nloc = hi - lo
    team = world > 1 and args.parallel == "team"
    prob, _ = drivers.initialize_problem(P, dev=local)
    if team:
        # the synthetic initial condition is built once on a single-GPU problem (setup only), then handed to the team:
        # rank r keeps its kr columns of the state and traces the packets of its band of ny / world rows
        from juliaraytracingsw_b200.slab import SlabProblem
        sol0 = prob.sol
        dt, ν = drivers.timestep_and_viscosity(P)
        prob.close()
        prob = SlabProblem(dist, local, barrier="host" if same_gpu else None, nx=P.nx, Lx=P.L, dt=dt, f=P.f, Cg=P.Cg, ν=ν, nν=P.nν, order=P.filter_order,
                           use_filter=P.use_filter, aliased_fraction=P.aliased_fraction)
        prob.sol = sol0
        del sol0
    k0 = (P.ω0 ** 2 - P.f ** 2) ** 0.5 / P.background_Cg
    packets = raytracing.generate_initial_wavepackets(prob, P.L, k0, nloc, P.sqrtNpackets, P.f, P.packet_Cg, nsub=P.nsub, first=lo)
    raytracing.get_velocity_info(prob, 0)
    t = prob.clock.t
    K, W = args.steps, max(args.warmup, 3)

    clocks = ClockSampler(local)

    def timed(nsteps, t):
        packets.sync(); prob.sync(); barrier()
        l0 = prob.launch_count()
        prob.timer_start()
        if team:                                     # the whole loop natively (swrt_packets_coupled_steps): ~15 small launches per step
            t = drivers.coupled_steps(prob, packets, nsteps)
        else:
            for _ in range(nsteps):
                t = drivers.coupled_step(prob, packets, t)
        ms = prob.timer_stop()
        barrier()
        return max_over_ranks(ms), prob.launch_count() - l0, t

    # ---- packets on the reference's initial lattice (t = 0: neighbouring packets share cache lines)
    for _ in range(W):
        t = drivers.coupled_step(prob, packets, t)
    ms_lat, _, t = timed(K, t)
    # ---- headline: positions pre-randomised U(-L/2, L/2), the fully mixed state (worst-case gathers)
    xk0 = packets.get()
    xk0[:, 0:2] = np.random.default_rng(1000 + rank).uniform(-P.L / 2, P.L / 2, size=(nloc, 2))
    packets.set(xk0)
    del xk0
    for _ in range(W):
        t = drivers.coupled_step(prob, packets, t)
    ms, launches, t = timed(K, t)
    value = ntot * K / (ms * 1e-3)

    # ---- same K steps with per-kernel CUDA events (roofline of the dominant kernel)
    prob.profile(2)
    for _ in range(K):
        t = drivers.coupled_step(prob, packets, t)
    prob.sync()
    kern = prob.profile_report()
    prob.profile(0)

    # ---- flow-only step (BASELINE: "RSW 2048^2 spectral steps/s (HBM %roofline)")
    prob.sync(); barrier()
    prob.timer_start()
    if team:
        prob.stepforward(K)
    else:
        flow.stepforward(prob, (), K)
    ms_flow = max_over_ranks(prob.timer_stop()) / K

    # ---- optional fp32 packet mode (north star: "reported separately"): Float32 node data + fp32 right-hand side, fp64 state
    fp32 = None
    if not args.no_fp32 and not team:
        raytracing.set_interpolation(prob, raytracing.INTERP_BILINEAR_F32)
        p32 = raytracing.Packets(prob, nloc, P.f, P.packet_Cg, nsub=P.nsub, interp=raytracing.INTERP_BILINEAR_F32)
        p32.set(packets.get(), np.where((np.arange(lo, hi) % 2) == 0, -1.0, 1.0))
        raytracing.get_velocity_info(prob, 0)
        for _ in range(W):
            t = drivers.coupled_step(prob, p32, t)
        prob.sync(); barrier()
        prob.timer_start()
        for _ in range(K):
            t = drivers.coupled_step(prob, p32, t)
        ms32 = max_over_ranks(prob.timer_stop())
        prob.profile(2)
        for _ in range(K):
            t = drivers.coupled_step(prob, p32, t)
        prob.sync()
        k32 = prob.profile_report()
        prob.profile(0)
        fp32 = {"value": ntot * K / (ms32 * 1e-3), "unit": "packet-steps/s", "ms_per_step": ms32 / K,
                "raytrace_ms": next(v for k, v in k32.items() if k.startswith("raytrace_rk4"))["ms_avg"],
                "what": "same coupled step; the tracer samples Float32 node records with an fp32 right-hand side (packet state, "
                        "cell coordinate and RK4 combination in fp64).  Not part of `value`."}
        p32.close()
        raytracing.set_interpolation(prob, raytracing.INTERP_BILINEAR)
        raytracing.get_velocity_info(prob, 0)

    # ---- end to end: pinned host packets in, output frame out, every step, through the public API
    e2e = None
    if not args.no_e2e:
        # page-locked buffers are allocated (first touched) by a thread running on the GPU's own NUMA node
        old_aff, near = os.sched_getaffinity(0), gpu_local_cpus(local)
        if near:
            os.sched_setaffinity(0, near)
        pin = lambda *shape: torch.empty(shape[::-1], dtype=torch.float64, pin_memory=True).numpy().T  # Fortran-ordered view
        h_xk, h_out, h_U, h_G = pin(nloc, 4), pin(nloc, 4), pin(nloc, 2), pin(nloc, 4)
        h_sign = torch.empty(nloc, dtype=torch.float64, pin_memory=True).numpy()
        packets.get(out=h_xk)
        h_sign[:] = np.where((np.arange(lo, hi) % 2) == 0, -1.0, 1.0)
        Ke = max(3, min(K, 10))
        nchunks = max(4, min(args.e2e_chunks, nloc // 131072))   # >= 4 row blocks at every N so that uploads, kernels and downloads overlap
        # Host-resident packets every step are served by the chunked pipeline API (raytracing.PacketPipeline): row blocks of this
        # rank's packets on their own streams against a flow that every rank steps itself -- PCIe, not the flow step, bounds
        # this path, so at N > 1 it runs beside the team (which keeps its packets on the GPUs) on a replicated problem.
        eprob = prob
        if team:
            eprob, _ = drivers.initialize_problem(P, dev=local)
            raytracing.get_velocity_info(eprob, 0)
            t = eprob.clock.t
        pipe = raytracing.PacketPipeline(eprob, nloc, P.f, P.packet_Cg, nchunks=nchunks, nsub=P.nsub)

        def e2e_step(t, frame, first=False):
            # the packets of this step arrive from the host and go back to it, chunk by chunk on the chunks' own streams:
            # uploads, sort + ray-trace kernels and downloads of different chunks overlap (raytracing.PacketPipeline).
            # frame=True also samples velocity and gradients at the new positions and copies them back (savepacketdata!).
            flow.stepforward(eprob, (), 1)
            raytracing.get_velocity_info(eprob, 1)
            new_t = eprob.clock.t
            pipe.step(h_xk, h_sign if first else None, (t, new_t), h_out, h_U if frame else None, h_G if frame else None,
                      after_raytrace=lambda: raytracing.swap_snapshots(eprob, alias=False))
            return new_t

        def e2e_time(frame):
            nonlocal t
            t = e2e_step(t, frame, first=True)   # the frequency signs are a parameter of the ensemble: uploaded once
            eprob.sync(); barrier()
            w0 = time.perf_counter()
            eprob.timer_start()
            for _ in range(Ke):
                t = e2e_step(t, frame)
            ms = eprob.timer_stop()
            wall = (time.perf_counter() - w0) * 1e3
            barrier()
            return max_over_ranks(max(ms, wall))
        ms_e, ms_f = e2e_time(False), e2e_time(True)
        # headline: what the reference arm's step does -- packets in, flow step + velocity info + ray trace, packets out
        e2e = {"value": ntot * Ke / (ms_e * 1e-3), "unit": "packet-steps/s", "h2d_bytes_per_step": int(8 * 4 * nloc),
               "d2h_bytes_per_step": int(8 * 4 * nloc), "steps": Ke, "ms_per_step": ms_e / Ke, "chunks": nchunks,
               "what": "per step: flow step + snapshot; packets (N,4) from pinned host memory -> set -> sort + raytrace -> "
                       "packets (N,4) copied back, in `chunks` row blocks on their own streams so copies and kernels overlap",
               "with_output_frame": {"value": ntot * Ke / (ms_f * 1e-3), "ms_per_step": ms_f / Ke, "d2h_bytes_per_step": int(8 * 10 * nloc),
                                     "what": "the same plus velocity (N,2) and gradients (N,4) sampled at the new positions and copied "
                                             "back every step (savepacketdata! with write_gradients)"}}
        pipe.close()
        if team:
            eprob.close()
        e2e["flow"] = "replicated on every rank (chunked pipeline API)" if world > 1 else "single GPU"
        e2e["numa_local_cpus"] = len(near) if near else None
        os.sched_setaffinity(0, old_aff)
    clk = clocks.stop()
    # ---- the other BASELINE configs as extra keys of the same line (single-GPU ones on rank 0's GPU at N = 1; config 5 at N = 8)
    extra_cfg = {}
    if not args.no_extra_configs and args.nx == 2048:
        try:
            packets.close(); prob.close()
            if world == 1:
                extra_cfg["config2"] = extra_config2(local)
                extra_cfg["config3"] = extra_config3(local)
            elif world == 8 and team:
                extra_cfg["config5"] = extra_config5(dist, local, world, 5)
        except Exception as ex:          # an extra key must never cost the headline line
            extra_cfg["extra_configs_error"] = repr(ex)[:300]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    F = 8.0 * args.nx * args.nx
    # dominant kernel by accumulated device time over the K profiled steps
    kern_k = {k: v for k, v in kern.items() if k not in ("packet_sort_kernels", "other")}
    name, rec = max(kern_k.items(), key=lambda kv: kv[1]["ms_total"])
    share = rec["ms_total"] / sum(v["ms_total"] for v in kern.values())
    if name.startswith("raytrace_rk4"):
        # Compulsory HBM bytes of one launch: packet state in and out (x,y,k,l,sign read; x,y,k,l written = 72 B) plus
        # every grid point of the two-level snapshot field once (80 B per point).  The 1280 B per packet-step that the
        # four RK4 stages GATHER are served from registers/L1/L2 once packets are cell-sorted and the
        # stencil is cached, so they are reported separately (gather_*), not as HBM traffic.
        by = 72.0 * nloc * args.nsub + 80.0 * args.nx * args.nx
        what = "72 B packet state per packet-step + 80 B per grid point of the two-level snapshot field once per launch"
        traffic = None
        extra = {"gather_bytes_per_launch": 1280.0 * nloc * args.nsub,
                 "gather_achieved_gbs": 1280.0 * nloc * args.nsub / (rec["ms_avg"] * 1e-3) / 1e9,
                 "gather_note": "SURVEY 8d figure (4 stages x 2 levels x 5 fields x 4 taps x 8 B); on-chip after sort + stencil cache",
                 }
        nt = load_ncu_traffic().get("raytrace")
        if nt and args.nx == nt.get("nx") and nloc == nt.get("packets") and name == nt.get("kernel"):   # only for the kernel and configuration the capture was taken on
            traffic = nt["dram_bytes_per_launch"]
            extra.update({"traffic_source": nt["source"], "binding_unit": nt.get("binding_unit")})
    else:
        flow_bytes = {"ypass_inv_kernel<RswLoaderA>": 8 * F, "xpass_kernel<RswXOp>": 9 * F, "ypass_fwd_kernel<RswCombiner>": 7 * F,
                      "ifmab3_update_rsw_kernel": 15 * F, "ypass_inv_kernel<PsiLoader>": 4 * F, "xpass_kernel<SnapshotXOp>": 8 * F}
        by, what, extra, traffic = flow_bytes.get(name, 0.0), "F-units of this stage of the 42 F (+10 F snapshot) contract, SURVEY 8d", {}, None
    ach = by / (rec["ms_avg"] * 1e-3) / 1e9
    roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": by, "per_unit": what, "avg_launch_ms": rec["ms_avg"], "peak_source": peak_src,
                "share_of_step": share, **extra}
    spectral = {"steps_per_s": 1e3 / ms_flow, "ms_per_step": ms_flow, "algorithmic_bytes_per_step": 42 * F,
                "achieved": 42 * F / (ms_flow * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": 42 * F / (ms_flow * 1e-3) / 1e9 / peak,
                "contract": "B_step = 42 F, F = 8 nx^2 bytes (SURVEY.md 8d, RSW + IFMAB3)",
                "traffic": None, "traffic_source": None}
    ft = load_ncu_traffic().get("spectral_step")
    if ft and args.nx == ft.get("nx") and world == 1:
        spectral.update({"traffic": ft["dram_bytes_per_step"], "traffic_source": ft["source"]})
    line = {
        "metric": "packet-steps/s", "value": value, "unit": "packet-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # `config` is exactly the reference arm's (same workload, same keys); what only describes this arm goes to config_detail
        "config": {"workload": WORKLOAD.format(nx=args.nx, n=ntot), "nx": args.nx, "packets": ntot, "nsub": args.nsub,
                   "integrator": "RK4", "interp": "bilinear", "l2": L2_NOTE},
        "config_detail": {"packets_per_gpu": nloc,
                          "packet_positions": "uniform random over the domain (fully mixed; the lattice start is value_lattice_t0)",
                          "parallelism": (f"team x{world}: flow slab-decomposed (NVLink peer stores + device barrier), packets sharded by y-band" if team
                                          else f"packets sharded x{world}, flow replicated"),
                          },
        "clocks": clk, "e2e": e2e, "fp32_packet_mode": fp32, "gpu_launches": int(launches), "roofline": roofline, "spectral_step": spectral,
        "value_lattice_t0": ntot * K / (ms_lat * 1e-3),
        "kernels": {k: {"ms_avg": round(v["ms_avg"], 5), "launches": v["launches"]} for k, v in kern.items()},
        **extra_cfg,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r = cpu_sample(args, 3, 1, cores)
        line["cpu_baseline"] = {"value": r["value"], "unit": "packet-steps/s", "cores": cores, "omp_threads": r["omp_threads"], "kind": "port", "sample": r["sample"],
                                "flow_ms": r["flow_ms"], "spectral_steps_per_s": 1e3 / r["flow_ms"]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE config 5: two-layer QG 4096^2, flow slab-decomposed over the ranks (direct NVLink transposes), 8192^2 packets
    sharded over the ranks.  Separate, smaller JSON line (not the driver's default workload)."""
    import torch
    import torch.distributed as dist

    import juliaraytracingsw_b200 as swrt
    from juliaraytracingsw_b200 import flow, raytracing
    from juliaraytracingsw_b200.slab import SlabProblem

    for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29577"), ("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")):
        os.environ.setdefault(k, v)
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = 4096 if args.nx == 2048 else args.nx
    sq = 8192 if args.sqrt_packets == 4096 else args.sqrt_packets
    ntot = sq * sq
    lo, hi = rank * ntot // world, (rank + 1) * ntot // world
    # swqg/TwoLayerParameters.jl recipe: f=3, Cg=1, rd=1/6, l=1, ug=0.025, cfltune=0.025, nutune=40, nnu=4
    f, Cg, ug = 3.0, 1.0, 0.025
    dt = 0.025 / ug * (2 * np.pi / nx) * ug
    nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** 8) / dt
    kw = dict(model="TwoLayerQG", nx=nx, dt=dt, nu=nu, nnu=4, f=f, Cg=Cg, U=ug, mu=1e-2, f0=f)
    prob = SlabProblem(dist, local, **kw) if world > 1 else swrt.Problem(local, **kw)
    rng = np.random.default_rng(0)
    sol = np.zeros((nx // 2 + 1, nx, 2), dtype=np.complex128)
    sol[1:24, :24] = (rng.standard_normal((23, 24, 2)) + 1j * rng.standard_normal((23, 24, 2))) * nx * nx * 1e-3
    prob.sol = sol
    k0 = (3.0 ** 0.5) * f / Cg
    packets = raytracing.generate_initial_wavepackets(prob, 2 * np.pi, k0, hi - lo, sq, f, Cg, nsub=args.nsub, first=lo)
    psi = raytracing.PSI_TWOLAYER_BAROCLINIC

    def snapshot(slot):
        if world > 1:
            prob.velocity_snapshot(slot, psi)
        else:
            raytracing.get_velocity_info(prob, slot, psi)

    def step(t):
        if world > 1:
            prob.stepforward(1)
        else:
            flow.stepforward(prob, (), 1)
        snapshot(1)
        tn = prob.clock.t
        raytracing.raytrace(packets, None, None, None, None, prob.grid, packets, dt, (t, tn))
        raytracing.swap_snapshots(prob)
        return tn

    snapshot(0)
    t = prob.clock.t
    K, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        t = step(t)
    torch.cuda.synchronize(); prob.sync(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        e0.record()
        for _ in range(K):
            t = step(t)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        prob.timer_start()
        for _ in range(K):
            t = step(t)
        ms = prob.timer_stop()
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(tt)
        print(json.dumps({"metric": "packet-steps/s", "value": ntot * K / (ms * 1e-3), "unit": "packet-steps/s", "n_gpus": world, "steps": K,
                          "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"two-layer QG {nx}^2 IFMAB3, flow slab-decomposed over {world} GPU(s) with direct NVLink "
                                                 f"transposes, + RK4 ray tracing of {ntot} packets sharded over the ranks (BASELINE config 5)",
                                     "nx": nx, "packets": ntot, "packet_positions": "reference lattice", "parallelism": f"slab x{world}"}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "config5":
        run_config5(a)
    else:
        run_swrt(a)
