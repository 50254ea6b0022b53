"""torchrun script: team mode (slab-decomposed flow step + y-band-sharded packets, all native over CUDA IPC) against the ORACLE.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tests/multigpu/team_parity.py

With SWRT_TEAM_SAME_GPU=1 every rank uses cuda:0 (CUDA IPC works between processes on one device), the process group is gloo and
the team barrier is the host variant -- this is how the single-GPU test box runs it; otherwise rank r uses GPU r and the barrier
is the device one (flag words over NVLink).  Exits non-zero on any mismatch.  Reference semantics: the reference's hot loop
raytracing/RaytracingDriver.jl:256-270 (flow step, velocity info, raytrace!, old = new) and its models' calcN!."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import juliaraytracingsw_b200 as swrt  # noqa: E402
from juliaraytracingsw_b200 import drivers, flow, raytracing  # noqa: E402
from juliaraytracingsw_b200.slab import SlabProblem  # noqa: E402
from helpers import config2_setup, rel_l2  # noqa: E402
from oracle import craytrace, ifmab3 as oif, qg as oqg, raytrace as oray, rsw as orsw  # noqa: E402
from oracle.grid import TwoDGrid  # noqa: E402


def main():
    same = os.environ.get("SWRT_TEAM_SAME_GPU", "0") == "1"
    local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    barrier = os.environ.get("SWRT_TEAM_BARRIER", "host" if same else "device")
    trace = craytrace.raytrace if craytrace.available() else oray.raytrace
    worst = {}

    # ---- 1. RSW coupled loop: flow parity, packet parity through several hand-overs, output frame, k-cutoff, set()
    nx = int(os.environ.get("SWRT_TEAM_NX", max(128, 16 * world)))
    side, nsteps, sort_every = 64, 41, 8
    g, p, sol0, c = config2_setup(nx)
    sp = SlabProblem(dist, local, barrier=barrier, nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    sp.sol = sol0
    np.testing.assert_array_equal(sp.gather_solution(), sol0)
    N = side * side
    lo, hi = rank * N // world, (rank + 1) * N // world
    xk0, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], side)
    rng = np.random.default_rng(17)
    xk0[:, 0:2] += rng.uniform(-9, 9, size=(N, 2))                   # anywhere, also far outside the domain (never wrapped)
    pk = raytracing.Packets(sp, hi - lo, c["f"], c["Cg"], sort_every=sort_every, first=lo)
    if os.environ.get("SWRT_TEAM_KERNEL"):      # e.g. "2": the three-level tile kernel on the band buffers (AUTO keeps the cached kernel at this packet density)
        pk.set_kernel(int(os.environ["SWRT_TEAM_KERNEL"]))
    pk.set(xk0[lo:hi], sign[lo:hi])
    assert sum(sp._gather(pk.resident())) == N
    np.testing.assert_array_equal(pk.get(), xk0[lo:hi])               # scatter to the band owners and back: bit exact
    raytracing.get_velocity_info(sp, 0)
    want_F0 = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    worst["snapshot"] = rel_l2(sp.gather_snapshot(0), want_F0)
    assert worst["snapshot"] < 1e-12, worst
    # the oracle's loop
    ts = oif.IFMAB3(np.zeros((1, 1, 3, 3)), c["dt"], lambda s: orsw.calcN(s, g, p))
    ts.expLdt = oif.expL_closed_form(g, p, c["dt"])
    ts.exp2Ldt = oif.expL_closed_form(g, p, 2 * c["dt"])
    sol, want = sol0.copy(), np.ascontiguousarray(xk0)
    Fo, t = want_F0, 0.0
    for _ in range(nsteps):
        ts.stepforward(sol)
        Fn = oray.get_velocity_info(orsw.get_streamfunction(g.dealias(sol.copy()), g, p), g)
        trace(want, sign, t, ts.t, Fo, Fn, g, c["f"], c["Cg"], nsub=1)
        Fo, t = Fn, ts.t
    # the team: first half step by step from Python, second half in one library call
    tt = sp.clock.t
    for _ in range(20):
        tt = drivers.coupled_step(sp, pk, tt)
    drivers.coupled_steps(sp, pk, nsteps - 20)
    assert sp.clock.step == nsteps
    if os.environ.get("SWRT_TEAM_KERNEL"):
        want_name = {"1": "raytrace_rk4_tile_kernel<3>", "2": "raytrace_rk4_tile3_kernel", "3": "raytrace_rk4_pipe_kernel"}[os.environ["SWRT_TEAM_KERNEL"]]
        assert sp.ray_kernel_name() == want_name, sp.ray_kernel_name()
    worst["flow"] = rel_l2(sp.gather_solution(), g.dealias(sol.copy()))
    assert worst["flow"] < 1e-10, worst
    got = pk.get()
    worst["packets"] = float(np.abs(got - want[lo:hi]).max() / np.abs(want).max())
    assert worst["packets"] < 1e-8, worst
    assert sum(sp._gather(pk.resident())) == N
    U, G = oray.interpolate_velocity(Fo, want[:, 0:2], g)
    Ug = np.empty((hi - lo, 2), order="F")
    Gg = raytracing.interpolate_gradients(raytracing.VelocityGradient(sp, 0), pk, output_U=Ug)
    np.testing.assert_allclose(Ug, U[lo:hi], rtol=0, atol=1e-9)
    np.testing.assert_allclose(Gg, G[lo:hi], rtol=0, atol=1e-8)
    kc = float(np.sqrt((want[:, 2] ** 2 + want[:, 3] ** 2)).mean())
    want2 = want.copy()
    nreset = oray.kcutoff_reset(want2, kc, c["k0"])
    assert pk.kcutoff_reset(kc, c["k0"]) == nreset and nreset > 0
    got2 = pk.get()
    assert np.abs(got2 - want2[lo:hi]).max() / np.abs(want2).max() < 1e-8
    pk.set(want2[lo:hi])                                              # positions only: the signs survive in the caller's order
    tt = drivers.coupled_step(sp, pk, sp.clock.t)
    ts.stepforward(sol)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(g.dealias(sol.copy()), g, p), g)
    trace(want2, sign, t, ts.t, Fo, Fn, g, c["f"], c["Cg"], nsub=1)
    assert np.abs(pk.get() - want2[lo:hi]).max() / np.abs(want2).max() < 1e-8
    ke, pe = sp.energies()
    assert abs(ke / orsw.kinetic_energy(g.dealias(sol.copy()), g) - 1) < 1e-9
    pk.close(); sp.close()

    # ---- 2. lattice generator on the team: bit-exact rows, every rank's block
    sp = SlabProblem(dist, local, barrier=barrier, nx=max(64, 16 * world), dt=1e-3, f=3.0)
    n_side = 24
    wantl, _ = oray.generate_initial_wavepackets(2 * np.pi, 5.196152422706632, n_side)
    Nl = n_side * n_side
    lo, hi = rank * Nl // world, (rank + 1) * Nl // world
    pl = raytracing.generate_initial_wavepackets(sp, 2 * np.pi, 5.196152422706632, hi - lo, n_side, 3.0, 1.0, first=lo)
    gotl = pl.get()
    np.testing.assert_array_equal(gotl[:, 0:2], wantl[lo:hi, 0:2])
    np.testing.assert_allclose(gotl[:, 2:4], wantl[lo:hi, 2:4], rtol=0, atol=4e-15)
    pl.close(); sp.close()

    # ---- 3. two-layer QG (config 5's model) and SWQG: flow steps + band snapshot against the oracle
    nx2, U, mu, f0, Cg, drr, nnu, dt = max(128, 16 * world), 0.5, 1e-2, 3.0, 1.0, 0.2, 4, 1e-3
    nu = 40 * 2 * np.pi / nx2 / ((nx2 / 2 - 1) ** (2 * nnu)) / dt
    F = 2 * f0 ** 2 / Cg ** 2 / drr
    from helpers import random_state
    g2, s3 = random_state(nx2, seed=9, amp=1.0, slope=1.0)
    q0 = np.ascontiguousarray(s3[:, :, :2])
    sp = SlabProblem(dist, local, barrier=barrier, model="TwoLayerQG", nx=nx2, dt=dt, U=U, mu=mu, f0=f0, Cg=Cg, δρρ0=drr, nu=nu, nnu=nnu)
    sp.sol = q0
    ts2 = oif.IFMAB3(oqg.twolayer_L(g2, F, U, mu, nu, nnu), dt, lambda s: oqg.twolayer_calcN(s, g2, F))
    w2 = q0.copy()
    sp.stepforward(12)
    for _ in range(12):
        ts2.stepforward(w2)
    w2 = g2.dealias(w2)
    worst["twolayer"] = rel_l2(sp.gather_solution(), w2)
    assert worst["twolayer"] < 1e-10, worst
    raytracing.get_velocity_info(sp, 1, raytracing.PSI_TWOLAYER_BAROCLINIC)
    psih = oqg.twolayer_streamfunction(w2, g2, F)
    worst["twolayer_snapshot"] = rel_l2(sp.gather_snapshot(1), oray.get_velocity_info(0.5 * (psih[:, :, 0] - psih[:, :, 1]), g2))
    assert worst["twolayer_snapshot"] < 1e-11, worst
    sp.close()
    Kd2 = 9.0
    sp = SlabProblem(dist, local, barrier=barrier, model="SWQG", nx=nx2, dt=2e-3, f=3.0, Cg=1.0, nu=nu, nnu=nnu)
    sp.sol = q0[:, :, 0]
    ts1 = oif.IFMAB3(oqg.swqg_L(g2, nu, nnu), 2e-3, lambda s: oqg.swqg_calcN(s, g2, Kd2))
    w1 = q0[:, :, 0].copy()
    sp.stepforward(12)
    for _ in range(12):
        ts1.stepforward(w1)
    worst["swqg"] = rel_l2(sp.gather_solution(), g2.dealias(w1))
    assert worst["swqg"] < 1e-10, worst
    sp.close()

    # ---- 4. the other slabbed models / steppers: MultiLayerQG-2 + FilteredAB3 (BASELINE config 1's flow, aliased_fraction = 0,
    #         raytracing/TwoLayerRaytracing.jl:174) and Modified RSW + IFMAB3 (rsw/ModifiedShallowWater.jl), flow steps against the oracle
    from test_gpu_parity import _config1_setup
    from helpers import oracle_steps
    nx1 = max(128, 16 * world)
    g1, sol1, c1 = _config1_setup(nx1)
    nnu1, nu1 = 4, 1e-16
    sp = SlabProblem(dist, local, barrier=barrier, model="MultiLayerQG", stepper="FilteredAB3", nx=nx1, dt=c1["dt"], f0=c1["f0"], H=c1["H"], b=c1["b"],
                     U=c1["U"], mu=c1["mu"], beta=c1["beta"], nu=nu1, nnu=nnu1, aliased_fraction=0)
    sp.sol = sol1
    Ld = (-nu1 * g1.Krsq ** nnu1)[:, :, None] * np.ones(2)
    tsm = oqg.FilteredAB3(Ld, c1["dt"], lambda s_: oqg.multilayer2_calcN(s_, g1, c1["F"], c1["U"][0], c1["U"][1], c1["beta"], c1["mu"]), c1["filt"][:, :, None])
    wm = sol1.copy()
    sp.stepforward(14)
    for _ in range(14):
        tsm.stepforward(wm)
    worst["multilayerqg_filteredab3"] = rel_l2(sp.gather_solution(), g1.dealias(wm.copy()))
    assert worst["multilayerqg_filteredab3"] < 1e-10, worst
    sp.close()
    gm, pm, solm, cm = config2_setup(nx1)
    sp = SlabProblem(dist, local, barrier=barrier, model="ModifiedShallowWater", nx=nx1, Lx=cm["L"], dt=cm["dt"], f=cm["f"], Cg=cm["Cg"], nu=cm["nu"], nnu=cm["nnu"])
    sp.sol = solm
    sp.stepforward(14)
    worst["modified_rsw"] = rel_l2(sp.gather_solution(), oracle_steps(gm, pm, solm, cm["dt"], 14, variant=orsw.MODIFIED))
    assert worst["modified_rsw"] < 1e-10, worst
    raytracing.get_velocity_info(sp, 1)
    sp.close()
    sp = SlabProblem(dist, local, barrier=barrier, model="LinborgShallowWater", nx=nx1, Lx=cm["L"], dt=cm["dt"], f=cm["f"], Cg=cm["Cg"], nu=cm["nu"], nnu=cm["nnu"])
    sp.sol = solm
    pkl = raytracing.Packets(sp, 256, cm["f"], cm["Cg"], first=256 * rank)            # (also through the fused loop: 8 + 3 y-jobs)
    xl, sl = oray.generate_initial_wavepackets(cm["L"], cm["k0"], 16 * int(np.sqrt(world)) if int(np.sqrt(world)) ** 2 == world else 16)
    pkl.set(np.resize(xl, (256 * world, 4))[256 * rank:256 * (rank + 1)], np.resize(sl, 256 * world)[256 * rank:256 * (rank + 1)])
    raytracing.get_velocity_info(sp, 0)
    drivers.coupled_steps(sp, pkl, 14)
    worst["lindborg_rsw"] = rel_l2(sp.gather_solution(), oracle_steps(gm, pm, solm, cm["dt"], 14, variant=orsw.LINDBORG))
    assert worst["lindborg_rsw"] < 1e-10, worst
    pkl.close(); sp.close()

    # ---- 5. multi-stage steppers on the team (every calcN! of a stage runs the three slab passes and their two barriers):
    #         Thomas-Yamada + ETDRK4 (BASELINE config 3's flow, thomasyamada/ThomasYamada.jl:129-274) and two-layer MultiLayerQG + FilteredRK4
    from oracle import ty as oty
    from oracle.grid import makefilter
    Lx_t, Ro, nnu_t, dt_t = 6 * np.pi, 1.0, 8, 5e-3
    nu_t = 5e-34 * (Lx_t / (2 * np.pi)) ** 16 * 1e12 * (64 / nx1) ** 16
    gt, s3t = random_state(nx1, seed=21, amp=0.3, slope=1.0, Lx=Lx_t)
    _, s1t = random_state(nx1, seed=22, amp=0.3, slope=1.0, Lx=Lx_t)
    solt = np.concatenate([s3t, s1t[:, :, :1]], axis=-1)
    sp = SlabProblem(dist, local, barrier=barrier, model="ThomasYamada", stepper="ETDRK4", nx=nx1, Lx=Lx_t, dt=dt_t, nu=nu_t, nnu=nnu_t, Ro=Ro)
    sp.sol = solt
    tst = oty.ETDRK4(oty.ty_L(gt, nu_t, nnu_t), dt_t, lambda s_: oty.ty_calcN(s_, gt, Ro))
    wt = solt.copy()
    sp.stepforward(6)
    for _ in range(6):
        tst.stepforward(wt)
    worst["thomasyamada_etdrk4"] = rel_l2(sp.gather_solution(), gt.dealias(wt.copy()))
    assert worst["thomasyamada_etdrk4"] < 1e-10, worst
    sp.close()
    sp = SlabProblem(dist, local, barrier=barrier, model="MultiLayerQG", stepper="FilteredRK4", nx=nx1, dt=c1["dt"], f0=c1["f0"], H=c1["H"], b=c1["b"],
                     U=c1["U"], mu=c1["mu"], beta=c1["beta"], nu=nu1, nnu=nnu1, aliased_fraction=0)
    sp.sol = sol1
    tsr = oty.FilteredRK4(Ld, c1["dt"], lambda s_: oqg.multilayer2_calcN(s_, g1, c1["F"], c1["U"][0], c1["U"][1], c1["beta"], c1["mu"]), c1["filt"])
    wr = sol1.copy()
    sp.stepforward(6)
    for _ in range(6):
        tsr.stepforward(wr)
    worst["multilayerqg_filteredrk4"] = rel_l2(sp.gather_solution(), g1.dealias(wr.copy()))
    assert worst["multilayerqg_filteredrk4"] < 1e-10, worst
    sp.close()

    allw = [None] * world
    dist.all_gather_object(allw, worst)
    if rank == 0:
        mx = {k: max(w[k] for w in allw) for k in worst}
        print(f"team parity ok on {world} ranks ({'one GPU, host barrier' if same else 'one GPU each'}, barrier={barrier}): " +
              ", ".join(f"{k} {v:.2e}" for k, v in mx.items()), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
