"""Parity at the BASELINE.json sizes (the template instantiations the benchmark actually runs: radix 16.16.2/4/8/16, the
prefetching y-passes, the CUDA-graph replay of the coupled loop, the cell sort), against the oracle on identical initial
conditions.  Tolerances are BASELINE.json's: spectral state <= 1e-10 relative L2 after 100 steps, packets (x, k) <= 1e-8
relative.  Reference semantics: rsw/RotatingShallowWater.jl:140-230, utils/IFMAB3.jl:129-169,
raytracing/RaytracingDriver.jl:256-270, thomasyamada/ThomasYamada.jl:129-274, swqg/TwoLayerQG.jl:152-198."""
import numpy as np
import pytest

import juliaraytracingsw_b200 as swrt
from juliaraytracingsw_b200 import drivers, flow, raytracing
from oracle import craytrace
from oracle import ifmab3 as oif
from oracle import raytrace as oray
from oracle import rsw as orsw
from oracle.grid import TwoDGrid, makefilter

from helpers import config2_setup, random_state, rel_l2

pytestmark = pytest.mark.gpu


def _oracle_coupled_loop(g, p, sol0, c, xk, sign, nsteps, nsub=1):
    """The reference's hot loop (raytracing/RaytracingDriver.jl:256-270) on the oracle: flow step, new velocity info,
    ray trace across the step (compiled C restatement of the oracle tracer when built, else NumPy), new becomes old."""
    ts = oif.IFMAB3(np.zeros((1, 1, 3, 3)), c["dt"], lambda s: orsw.calcN(s, g, p))
    ts.expLdt = oif.expL_closed_form(g, p, c["dt"])
    ts.exp2Ldt = oif.expL_closed_form(g, p, 2 * c["dt"])
    sol = sol0.copy()
    xk = np.ascontiguousarray(xk)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol, g, p), g)
    trace = craytrace.raytrace if craytrace.available() else oray.raytrace
    t = 0.0
    for _ in range(nsteps):
        ts.stepforward(sol)
        Fn = oray.get_velocity_info(orsw.get_streamfunction(g.dealias(sol.copy()), g, p), g)
        trace(xk, sign, t, ts.t, Fo, Fn, g, c["f"], c["Cg"], nsub=nsub)
        Fo, t = Fn, ts.t
    return g.dealias(sol), xk


def test_config4_rsw_2048_100_steps_and_16384_packets():
    """BASELINE config 4's flow at full size: 100 coupled steps at 2048^2 against the oracle, with 16 384 packets
    (128^2 lattice) ray-traced across every one of them through swrt_packets_coupled_steps."""
    nx, side, nsteps = 2048, 128, 100
    g, p, sol0, c = config2_setup(nx)
    prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    pk = raytracing.generate_initial_wavepackets(prob, c["L"], c["k0"], side * side, side, c["f"], c["Cg"])
    raytracing.get_velocity_info(prob, 0)
    drivers.coupled_steps(prob, pk, nsteps)
    xk0, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], side)
    want_sol, want_xk = _oracle_coupled_loop(g, p, sol0, c, xk0, sign, nsteps)
    err = rel_l2(prob.sol, want_sol)
    assert err < 1e-10, err
    assert prob.clock.step == nsteps
    got = pk.get()
    perr = np.abs(got - want_xk).max() / np.abs(want_xk).max()
    assert perr < 1e-8, perr


def test_config2_rsw_512_65536_packets_coupled_graph_loop():
    """BASELINE config 2 exactly: RSW 512^2 + 65 536 packets (256^2 lattice), 112 coupled steps in ONE library call -- the
    six-step CUDA graphs of swrt_packets_coupled_steps with >= 6 cell sorts (sort_every = 16) in between."""
    nx, side, nsteps = 512, 256, 112
    g, p, sol0, c = config2_setup(nx)
    prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    pk = raytracing.generate_initial_wavepackets(prob, c["L"], c["k0"], side * side, side, c["f"], c["Cg"], sort_every=16)
    raytracing.get_velocity_info(prob, 0)
    l0 = prob.launch_count()
    drivers.coupled_steps(prob, pk, nsteps)
    assert prob.launch_count() - l0 >= 8 * nsteps          # every step's kernels are accounted for, replayed or not
    xk0, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], side)
    want_sol, want_xk = _oracle_coupled_loop(g, p, sol0, c, xk0, sign, nsteps)
    err = rel_l2(prob.sol, want_sol)
    assert err < 1e-10, err
    got = pk.get()
    perr = np.abs(got - want_xk).max() / np.abs(want_xk).max()
    assert perr < 1e-8, perr
    # output frame at the end of the loop (savepacketdata!): velocity and gradients at the packets, caller's row order
    Fn = oray.get_velocity_info(orsw.get_streamfunction(want_sol, g, p), g)
    U, G = oray.interpolate_velocity(Fn, want_xk[:, 0:2], g)
    np.testing.assert_allclose(raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 0), pk), G, rtol=0, atol=1e-9)


def test_config3_thomasyamada_1024_etdrk4():
    """BASELINE config 3's flow: Thomas-Yamada 1024^2, Lx = 6 pi, ETDRK4, dt = 5e-3 (thomasyamada/gpu-setup/Parameters.jl),
    26 steps (= 104 calcN! evaluations of 24 transforms each) against the oracle."""
    from oracle import ty as oty
    nx, Lx, Ro, nnu, dt = 1024, 6 * np.pi, 1.0, 8, 5e-3
    nu = 5e-34 * (Lx / (2 * np.pi)) ** 16
    g, s3 = random_state(nx, seed=21, amp=0.3, slope=1.0, Lx=Lx)
    _, s1 = random_state(nx, seed=22, amp=0.3, slope=1.0, Lx=Lx)
    sol0 = np.concatenate([s3, s1[:, :, :1]], axis=-1)
    prob = swrt.Problem(model="ThomasYamada", stepper="ETDRK4", nx=nx, Lx=Lx, dt=dt, nu=nu, nnu=nnu, Ro=Ro)
    flow.set_solution(prob, *(sol0[:, :, i] for i in range(4)))
    ts = oty.ETDRK4(oty.ty_L(g, nu, nnu), dt, lambda s: oty.ty_calcN(s, g, Ro))
    want = sol0.copy()
    for n in (1, 25):
        flow.stepforward(prob, (), n)
        for _ in range(n):
            ts.stepforward(want)
        err = rel_l2(prob.sol, g.dealias(want.copy()))
        assert err < 1e-10, (n, err)
    assert rel_l2(prob.vars.uc, g.irfft2(g.dealias(want.copy())[:, :, 1])) < 1e-10


def test_config1_multilayerqg_256_filteredab3_100_steps():
    """BASELINE config 1's flow at its size: two-layer MultiLayerQG 256^2, FilteredAB3, aliased_fraction = 0
    (raytracing/TwoLayerRaytracing.jl:174), 100 steps."""
    from oracle import qg as oqg
    from test_gpu_parity import _config1_setup
    nx = 256
    g, sol0, c = _config1_setup(nx)
    nnu, nu = 4, 1e-16
    prob = swrt.Problem(model="MultiLayerQG", stepper="FilteredAB3", nx=nx, dt=c["dt"], f0=c["f0"], H=c["H"], b=c["b"], U=c["U"],
                        mu=c["mu"], beta=c["beta"], nu=nu, nnu=nnu, aliased_fraction=0)
    prob.sol = sol0
    L = (-nu * g.Krsq ** nnu)[:, :, None] * np.ones(2)
    ts = oqg.FilteredAB3(L, c["dt"], lambda s: oqg.multilayer2_calcN(s, g, c["F"], c["U"][0], c["U"][1], c["beta"], c["mu"]),
                         c["filt"][:, :, None])
    want = sol0.copy()
    flow.stepforward(prob, (), 100)
    for _ in range(100):
        ts.stepforward(want)
    err = rel_l2(prob.sol, g.dealias(want.copy()))
    assert err < 1e-10, err


def test_config5_twolayerqg_4096_steps():
    """BASELINE config 5's flow on one GPU: two-layer QG 4096^2 with IFMAB3 (swqg/TwoLayerParameters.jl recipe), 5 steps
    (3 Euler start-up calls + 2 AB3 steps) against the oracle's general matrix exponential; then the baroclinic
    velocity snapshot the packets sample."""
    from oracle import qg as oqg
    nx, f, Cg, ug, nnu = 4096, 3.0, 1.0, 0.025, 4
    dt = 0.025 * (2 * np.pi / nx)
    nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** (2 * nnu)) / dt
    mu, drr = 1e-2, 1.0
    F = 2 * f ** 2 / Cg ** 2 / drr
    g = TwoDGrid(nx)
    rng = np.random.default_rng(0)
    sol0 = np.zeros((g.nkr, g.nl, 2), dtype=np.complex128)
    sol0[1:24, :24] = (rng.standard_normal((23, 24, 2)) + 1j * rng.standard_normal((23, 24, 2))) * nx * nx * 1e-3
    sol0[1:24, -23:] = (rng.standard_normal((23, 23, 2)) + 1j * rng.standard_normal((23, 23, 2))) * nx * nx * 1e-3
    prob = swrt.Problem(model="TwoLayerQG", nx=nx, dt=dt, U=ug, mu=mu, f0=f, Cg=Cg, δρρ0=drr, nu=nu, nnu=nnu)
    assert abs(prob.desc.F / F - 1) < 1e-15
    prob.sol = sol0
    ts = oif.IFMAB3(oqg.twolayer_L(g, F, ug, mu, nu, nnu), dt, lambda s: oqg.twolayer_calcN(s, g, F))
    want = sol0.copy()
    flow.stepforward(prob, (), 5)
    for _ in range(5):
        ts.stepforward(want)
    want = g.dealias(want)
    err = rel_l2(prob.sol, want)
    assert err < 1e-10, err
    vel, _ = raytracing.get_velocity_info(prob, 1, raytracing.PSI_TWOLAYER_BAROCLINIC)
    psih = oqg.twolayer_streamfunction(want, g, F)
    ref = oray.get_velocity_info(0.5 * (psih[:, :, 0] - psih[:, :, 1]), g)
    assert rel_l2(vel._arr(), ref) < 1e-11


def test_set_stream_after_packets_create():
    """A packet handle attached before swrt_flow_set_stream follows the flow onto the caller's stream (the handle resolves the
    stream at every use instead of caching the one it was created with)."""
    import ctypes as C
    import torch
    from juliaraytracingsw_b200._lib import check, lib
    g, p, sol0, c = config2_setup(128)
    outs = []
    for move in (False, True):
        prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
        prob.sol = sol0
        pk = raytracing.generate_initial_wavepackets(prob, c["L"], c["k0"], 900, 30, c["f"], c["Cg"], sort_every=4)
        if move:
            st = torch.cuda.Stream()
            check(lib().swrt_flow_set_stream(prob._h, C.c_void_p(st.cuda_stream)))
        raytracing.get_velocity_info(prob, 0)
        drivers.coupled_steps(prob, pk, 20)
        pk.kcutoff_reset(5.0, c["k0"])
        outs.append((prob.sol, pk.get()))
        pk.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_max_abs_uv_propagates_nan():
    """maximum(abs.(vars.u)) of a blown-up field is NaN in the reference's CFL log (raytracing/RaytracingDriver.jl:244)."""
    g, p, sol0, c = config2_setup(64)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    bad = sol0.copy()
    bad[3, 4, 0] = np.nan
    prob.sol = bad
    um, vm = flow.max_abs_uv(prob)
    assert np.isnan(um) and np.isfinite(vm)


@pytest.mark.parametrize("nsub,shift", [(1, 0.0), (3, 0.0), (2, 7.3)])
def test_tile_kernel_is_bit_identical_to_the_cached_kernel(nsub, shift):
    """The TMA-staged tile kernel (one CTA per 16x16-cell sort tile, node records in shared memory) does the same arithmetic as
    the stencil-cached kernel: bit-identical packets, and both within 1e-8 of the oracle.  Packets sit anywhere in (and, with
    `shift`, far outside) the domain, so boundary tiles (global path), interior tiles (shared-memory path) and packets that have
    drifted beyond the staged margin since the last sort (`sort_every` = 40 steps at |U| dt ~ 0.1 cells -> ~4 cells) all occur."""
    nx = 128
    g, c, Fo, Fn, xk, sign = _tile_case(nx, 160, shift)
    outs = []
    for kernel in (raytracing.RAYKERNEL_CACHED, raytracing.RAYKERNEL_TILE):
        prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
        raytracing.set_velocity_info(prob, 0, Fo)
        raytracing.set_velocity_info(prob, 1, Fn)
        pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=nsub, sort_every=40)
        pk.set_kernel(kernel)
        pk.set(xk, sign)
        t = 0.0
        for _ in range(45):
            raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t, t + c["dt"]))
            t += c["dt"]
        outs.append(pk.get())
    np.testing.assert_array_equal(outs[0], outs[1])
    want = np.ascontiguousarray(xk.copy())
    t = 0.0
    trace = craytrace.raytrace if craytrace.available() else oray.raytrace
    for _ in range(45):
        trace(want, sign, t, t + c["dt"], Fo, Fn, g, c["f"], c["Cg"], nsub=nsub)
        t += c["dt"]
    assert np.abs(outs[1] - want).max() / np.abs(want).max() < 1e-8


@pytest.mark.parametrize("lerp,shift", [(0, 0.0), (1, 0.0), (0, 7.3)])
def test_three_level_tile_kernel_against_cached_kernel_and_oracle(lerp, shift):
    """nsub = 1: the three-level tile kernels (one CTA per tile, and the persistent producer/consumer pipeline: first level, mean of the levels, last level staged in shared memory, one patch per
    RK4 stage time) against the stencil-cached kernel (<= 1e-12 relative after 45 steps: bilinear(mean) = mean(bilinear), only
    the association differs) and against the oracle (<= 1e-8), for both time-lerp conventions.  Interior tiles (TMA), tiles whose
    patch wraps around the domain edge (filled by the CTA's threads) and packets beyond the staged margin (global path) occur."""
    nx = 128
    g, c, Fo, Fn, xk, sign = _tile_case(nx, 160, shift)
    outs = []
    names = {raytracing.RAYKERNEL_CACHED: "raytrace_rk4_cached_kernel<4>", raytracing.RAYKERNEL_TILE3: "raytrace_rk4_tile3_kernel",
             raytracing.RAYKERNEL_PIPE: "raytrace_rk4_pipe_kernel"}
    for kernel in names:
        prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
        raytracing.set_velocity_info(prob, 0, Fo)
        raytracing.set_velocity_info(prob, 1, Fn)
        pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=1, sort_every=40, time_lerp=lerp)
        pk.set_kernel(kernel)
        pk.set(xk, sign)
        t = 0.0
        for _ in range(45):
            raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t, t + c["dt"]))
            t += c["dt"]
        outs.append(pk.get())
        assert prob.ray_kernel_name() == names[kernel]
    np.testing.assert_array_equal(outs[1], outs[2])       # persistent warp-specialised kernel == one CTA per tile, bit for bit
    scale = np.abs(outs[0]).max()
    assert np.abs(outs[0] - outs[1]).max() / scale < 1e-12
    want = np.ascontiguousarray(xk.copy())
    t = 0.0
    trace = craytrace.raytrace if craytrace.available() else oray.raytrace
    for _ in range(45):
        trace(want, sign, t, t + c["dt"], Fo, Fn, g, c["f"], c["Cg"], nsub=1, lerp=lerp)
        t += c["dt"]
    assert np.abs(outs[1] - want).max() / np.abs(want).max() < 1e-8


@pytest.mark.parametrize("lerp,shift", [(0, 0.0), (1, 7.3)])
def test_fp32_three_level_tile_kernel(lerp, shift):
    """fp32 packet mode through its staged kernel (Float32 patches: first level, mean, last level; fp32 right-hand side, fp64 state)
    against the stencil-cached fp32 kernel and against the fp64 oracle fed the same Float32-rounded fields (2e-6 relative per the
    fp32 mode's contract, over 30 steps with one re-sort)."""
    nx = 128
    g, c, Fo, Fn, xk, sign = _tile_case(nx, 160, shift)
    Fo32, Fn32 = Fo.astype(np.float32), Fn.astype(np.float32)
    _, _, sol0, _ = config2_setup(nx)
    outs = []
    names = {raytracing.RAYKERNEL_CACHED: "raytrace_rk4_f32_kernel", raytracing.RAYKERNEL_AUTO: "raytrace_rk4_tile3_f32_kernel"}
    for kernel in names:
        prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
        raytracing.set_interpolation(prob, raytracing.INTERP_BILINEAR_F32)
        prob.sol = sol0                                        # (fp32 snapshots are built from the flow state only: the two levels of _tile_case)
        raytracing.get_velocity_info(prob, 0)
        flow.stepforward(prob, (), 1)
        raytracing.get_velocity_info(prob, 1)
        pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=1, sort_every=20, time_lerp=lerp, interp=raytracing.INTERP_BILINEAR_F32)
        pk.set_kernel(kernel)
        pk.set(xk, sign)
        t = 0.0
        for _ in range(30):
            raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t, t + c["dt"]))
            t += c["dt"]
        outs.append(pk.get())
        assert prob.ray_kernel_name() == names[kernel]
    want = np.ascontiguousarray(xk.copy())
    t = 0.0
    trace = craytrace.raytrace if craytrace.available() else oray.raytrace
    for _ in range(30):
        trace(want, sign, t, t + c["dt"], Fo32.astype(np.float64), Fn32.astype(np.float64), g, c["f"], c["Cg"], nsub=1, lerp=lerp)
        t += c["dt"]
    scale = np.abs(want).max()
    assert np.abs(outs[0] - want).max() / scale < 2e-6
    assert np.abs(outs[1] - want).max() / scale < 2e-6
    assert np.abs(outs[0] - outs[1]).max() / scale < 2e-6


def _tile_case(nx, n_side, shift):
    g, p, sol0, c = config2_setup(nx)
    from helpers import oracle_steps
    sol1 = oracle_steps(g, p, sol0, c["dt"], 1)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(sol1, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], n_side)
    rng = np.random.default_rng(5)
    xk[:, 0:2] = rng.uniform(-c["L"] / 2, c["L"] / 2, size=(xk.shape[0], 2)) + shift * c["L"]
    return g, c, Fo, Fn, xk, sign


@pytest.mark.parametrize("nx", [128, 512])
def test_device_initial_condition_generator_K15(nx):
    """set_initial_condition! (rsw/RSWRaytracingDriver.jl:15-54) built on the device from the host's random numbers: the state
    equals the oracle's restatement fed the same numbers, and K15 holds: max|u_g| = ag, max|u_w| = aw."""
    g, p, want, c = config2_setup(nx)              # the oracle's recipe with default_rng(1234)
    P = drivers.Parameters(nx=nx)
    prob, _ = drivers.initialize_problem(P)        # drivers.set_initial_condition with default_rng(P.seed = 1234)
    got = prob.sol
    assert rel_l2(got, want) < 1e-13
    _, (ugh, vgh, egh), (uwh, vwh, ewh) = orsw.initial_condition(g, p, P.Kg, P.ag, P.Kw, P.aw, np.random.default_rng(1234))
    for part, amp in (((ugh, vgh, egh), P.ag), ((uwh, vwh, ewh), P.aw)):
        prob.sol = np.stack(part, axis=-1)
        assert abs(flow.max_abs_uv(prob)[0] / amp - 1) < 1e-12


@pytest.mark.parametrize("nw,tol", [(6, 3e-5), (8, 3e-7), (12, 2e-10)])
def test_nufft_sampler_against_the_exact_trigonometric_sum(nw, tol):
    """f4: type-2 NUFFT sampling (2x oversampled grid of the deconvolved spectrum + nw x nw exponential-of-semicircle kernel) against
    the exact trigonometric sum at arbitrary points (what raytracing/NUFFTRaytracing.jl:68-84 aims at with nufft2d2, tol 1e-5),
    and a short ray trace against the oracle's RK4 with the exact sampler."""
    nx = 64
    g, p, sol0, c = config2_setup(nx)
    from helpers import oracle_steps
    sol1 = oracle_steps(g, p, sol0, c["dt"], 2)
    prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_interpolation(prob, raytracing.INTERP_NUFFT)
    raytracing.set_nufft_width(prob, nw)
    with pytest.raises(swrt._lib.SwrtError):
        raytracing.get_velocity_info(prob, 0)                  # needs the 2x oversampled node grid
    raytracing.set_snapshot_refinement(prob, 2)
    prob.sol = sol0
    raytracing.get_velocity_info(prob, 0)
    flow.stepforward(prob, (), 2)
    raytracing.get_velocity_info(prob, 1)
    assert rel_l2(prob.sol, sol1) < 1e-12
    psi0, psi1 = orsw.get_streamfunction(sol0, g, p), orsw.get_streamfunction(sol1, g, p)
    rng = np.random.default_rng(4)
    n = 300
    xk = np.zeros((n, 4))
    xk[:, 0:2] = rng.uniform(-3 * np.pi, 3 * np.pi, size=(n, 2))         # also outside the domain (periodic)
    xk[:, 2:4] = c["k0"] * rng.standard_normal((n, 2))
    sign = np.where(np.arange(n) % 2 == 0, -1.0, 1.0)
    pk = raytracing.Packets(prob, n, c["f"], c["Cg"], nsub=2, interp=raytracing.INTERP_NUFFT)
    pk.set(xk, sign)
    exact = oray.sample_trigonometric(psi0, xk[:, 0], xk[:, 1], g)
    U = np.empty((n, 2), order="F")
    G = raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 0), pk, output_U=U)
    scale_u, scale_g = np.abs(exact[:, 0:2]).max(), np.abs(exact[:, 2:5]).max()
    assert np.abs(U - exact[:, 0:2]).max() / scale_u < tol
    assert np.abs(G[:, 0:3] - exact[:, 2:5]).max() / scale_g < tol
    np.testing.assert_array_equal(G[:, 3], -G[:, 0])
    # RK4 over the two flow steps with the exact sampler on the oracle side
    t1 = 2 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (0.0, t1))
    want, h = xk.copy(), t1 / 2
    rhs = lambda z, al: oray.rhs_sampler(z, sign, al, oray.sample_trigonometric(psi0, z[:, 0], z[:, 1], g),
                                         oray.sample_trigonometric(psi1, z[:, 0], z[:, 1], g), c["f"], c["Cg"])
    for s_ in range(2):
        a0 = s_ * h / t1
        k1 = rhs(want, a0); k2 = rhs(want + 0.5 * h * k1, a0 + 0.5 * h / t1); k3 = rhs(want + 0.5 * h * k2, a0 + 0.5 * h / t1)
        k4 = rhs(want + h * k3, a0 + h / t1)
        want = want + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    d = pk.get() - want
    assert np.abs(d).max() / np.abs(want).max() < tol          # the packets move by O(dt): the sampling error barely shows
