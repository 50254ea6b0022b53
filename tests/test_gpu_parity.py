"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on identical inputs.
Tolerances: spectral state <= 1e-10 relative L2 after 100 steps; packets <= 1e-8 relative (north star);
integer / indexing results bit-exact."""
import numpy as np
import pytest

import juliaraytracingsw_b200 as swrt
from juliaraytracingsw_b200 import drivers, flow, raytracing
from oracle import raytrace as oray
from oracle import rsw as orsw
from oracle.grid import TwoDGrid, makefilter

from helpers import config2_setup, oracle_steps, random_state, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny", [(32, 32), (64, 64), (128, 64), (64, 256), (512, 512), (1024, 1024)])
def test_inverse_transform_matches_irfft2(nx, ny):
    g, sol = random_state(nx, ny, seed=nx + ny)
    prob = swrt.Problem(nx=nx, ny=ny, f=3.0, dt=1e-3)
    prob.sol = sol
    np.testing.assert_array_equal(prob.sol, sol)                  # pack/unpack is lossless on a dealiased state
    for which, name in ((0, "u"), (1, "v"), (2, "eta")):
        want = g.irfft2(sol[:, :, which])
        got = getattr(prob.vars, name)
        assert rel_l2(got, want) < 2e-14, (name, rel_l2(got, want))
    p = orsw.Params(0, 4, 3.0, 1.0)
    assert rel_l2(prob.vars.zeta, orsw.updatevars(sol.copy(), g, p)[3]) < 2e-14


def test_set_solution_dealiases_and_ignores_nonhermitian_dc_column():
    g = TwoDGrid(64)
    rng = np.random.default_rng(5)
    sol = rng.standard_normal((g.nkr, g.nl, 3)) + 1j * rng.standard_normal((g.nkr, g.nl, 3))   # aliased + non-Hermitian
    prob = swrt.Problem(nx=64, f=3.0, dt=1e-3)
    prob.sol = sol
    want = g.dealias(sol.copy())
    np.testing.assert_array_equal(prob.sol, want)
    assert rel_l2(prob.vars.u, g.irfft2(want[:, :, 0])) < 2e-14    # c2r semantics: Re of the kr=0 column after the l-transform


@pytest.mark.parametrize("model,variant", [("RotatingShallowWater", orsw.RSW), ("ModifiedShallowWater", orsw.MODIFIED),
                                           ("LinborgShallowWater", orsw.LINDBORG)])
@pytest.mark.parametrize("nx", [64, 128])
def test_rsw_100_steps_parity(model, variant, nx):
    g, p, sol0, c = config2_setup(nx)
    prob = swrt.Problem(model=model, nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    for nsteps, done in ((1, 1), (3, 4), (96, 100)):               # Euler start-up (3 calls), first AB3 step, then 100
        flow.stepforward(prob, (), nsteps)
        want = oracle_steps(g, p, sol0, c["dt"], done, variant)
        err = rel_l2(prob.sol, want)
        assert err < 1e-10, (done, err)
        assert prob.clock.step == done and abs(prob.clock.t - done * c["dt"]) < 1e-12


def test_rsw_filter_parity():
    nx = 64
    g, p, sol0, c = config2_setup(nx)
    prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=0.0, nnu=c["nnu"], use_filter=True, order=8)
    prob.sol = sol0
    flow.stepforward(prob, (), 20)
    p0 = orsw.Params(0.0, c["nnu"], c["f"], c["Cg"])
    want = oracle_steps(g, p0, sol0, c["dt"], 20, filt=makefilter(g, order=8))
    assert rel_l2(prob.sol, want) < 1e-10


@pytest.mark.parametrize("model,variant", [("RotatingShallowWater", "rsw"), ("ModifiedShallowWater", "modified"), ("LinborgShallowWater", "lindborg")])
def test_forcing_hook_parity(model, variant):
    """addforcing! (rsw/RotatingShallowWater.jl:228-240): `calcF!` fills vars.Fh, calcN! ends with `@. N += vars.Fh` (the 2-D field
    broadcast over the three equations).  A time-dependent calcF! through stepforward(..., calcF=) over the Euler start-up and
    AB3 steps (64^2: the CUDA-graph replay must stay out of the way), then the forcing is removed again."""
    from oracle import ifmab3 as oif
    nx, nsteps = 64, 12
    g, p, sol0, c = config2_setup(nx)
    rng = np.random.default_rng(9)
    shape = g.dealias((rng.standard_normal((g.nkr, g.nl)) + 1j * rng.standard_normal((g.nkr, g.nl)))[:, :, None].copy())[:, :, 0]
    shape *= 0.05 * np.abs(sol0).max() / c["dt"] / np.abs(shape).max() * (g.Krsq < 8 ** 2)
    shape[0, :] = 0.0

    def calcF(Fh, t, clock):
        Fh[:] = shape * np.cos(3.0 * t)

    prob = swrt.Problem(model=model, nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    flow.stepforward(prob, (), nsteps, calcF=calcF)
    Fh = np.zeros_like(shape)
    ts = oif.IFMAB3(oif.expL_closed_form(g, p, c["dt"], variant), c["dt"], lambda s_: orsw.calcN(s_, g, p, variant, Fh=Fh))
    ts.expLdt = oif.expL_closed_form(g, p, c["dt"], variant)
    ts.exp2Ldt = oif.expL_closed_form(g, p, 2 * c["dt"], variant)
    want = sol0.copy()
    for _ in range(nsteps):
        calcF(Fh, ts.t, None)
        ts.stepforward(want)
    assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-12
    unforced = oracle_steps(g, p, sol0, c["dt"], nsteps, variant=variant)
    assert rel_l2(prob.sol, unforced) > 1e-4                     # the forcing did act
    # cleared: the next steps equal the oracle's without Fh (and the replayed graphs are back)
    flow.set_forcing(prob, None)
    flow.stepforward(prob, (), 7)
    Fh[:] = 0.0
    for _ in range(7):
        ts.stepforward(want)
    assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-12
    qg = swrt.Problem(model="SWQG", nx=nx, dt=c["dt"])
    with pytest.raises(swrt._lib.SwrtError):
        flow.set_forcing(qg, shape)                             # swqg/SWQG.jl defines the hook but its calcN! never calls it


def test_energies_and_diagnostics():
    g, p, sol0, c = config2_setup(128)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    assert abs(flow.kinetic_energy(prob) / orsw.kinetic_energy(sol0, g) - 1) < 1e-13
    assert abs(flow.potential_energy(prob) / orsw.potential_energy(sol0, g, p) - 1) < 1e-13
    u, v, _, _ = orsw.updatevars(sol0.copy(), g, p)
    um, vm = flow.max_abs_uv(prob)
    assert abs(um / np.abs(u).max() - 1) < 1e-13 and abs(vm / np.abs(v).max() - 1) < 1e-13
    assert flow.has_nan(prob) is False
    bad = sol0.copy(); bad[3, 4, 0] = np.nan
    prob.sol = bad
    assert flow.has_nan(prob) is True


def test_velocity_snapshot_parity():
    g, p, sol0, c = config2_setup(128)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    vel, grad = raytracing.get_velocity_info(prob, 1)
    want = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    got = vel._arr()
    for c_ in range(5):
        assert rel_l2(got[:, :, c_], want[:, :, c_]) < 1e-13, c_
    np.testing.assert_array_equal(grad.vy, -got[:, :, 2])


def _packet_case(nx, n_side, seed):
    g, p, sol0, c = config2_setup(nx)
    sol1 = oracle_steps(g, p, sol0, c["dt"], 1)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(sol1, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], n_side)
    rng = np.random.default_rng(seed)
    xk[:, 0:2] += rng.uniform(-40, 40, size=(xk.shape[0], 2))     # packets are never wrapped into the domain
    return g, c, Fo, Fn, xk, sign


@pytest.mark.parametrize("lerp", [0, 1])
@pytest.mark.parametrize("nsub", [1, 4])
def test_raytrace_parity(lerp, nsub):
    g, c, Fo, Fn, xk, sign = _packet_case(128, 48, 3)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_velocity_info(prob, 0, Fo)
    raytracing.set_velocity_info(prob, 1, Fn)
    pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=nsub, time_lerp=lerp)
    pk.set(xk, sign)
    t0, t1 = 0.3, 0.3 + 25 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t0, t1))
    want = oray.raytrace(xk.copy(), sign, t0, t1, Fo, Fn, g, c["f"], c["Cg"], nsub=nsub, lerp=lerp)
    got = pk.get()
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-8
    assert rel_l2(got, want) < 1e-12
    # sampler for output frames
    U, G = oray.interpolate_velocity(Fn, want[:, 0:2], g)
    pk.set(want, sign)
    np.testing.assert_allclose(raytracing.interpolate_velocity(raytracing.Velocity(prob, 1), pk), U, rtol=0, atol=1e-13)
    np.testing.assert_allclose(raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 1), pk), G, rtol=0, atol=1e-12)


def test_cell_sort_keeps_caller_order_bit_exact():
    """The device copy is re-sorted by grid cell for locality; every host-visible array keeps the caller's rows."""
    g, c, Fo, Fn, xk, sign = _packet_case(64, 40, 11)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_velocity_info(prob, 0, Fo)
    raytracing.set_velocity_info(prob, 1, Fn)
    outs = []
    for sort_every in (0, 1, 3):
        pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=2, sort_every=sort_every)
        pk.set(xk, sign)
        t = 0.0
        for _ in range(7):
            raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t, t + c["dt"]))
            t += c["dt"]
        n_reset = pk.kcutoff_reset(5.3, c["k0"])
        outs.append((pk.get(), raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 1), pk), n_reset))
        pk.set(outs[-1][0])                                       # positions only: the signs must survive in caller order
        raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t, t + c["dt"]))
        outs[-1] += (pk.get(),)
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[1], outs[0][1])
        assert o[2] == outs[0][2]
        np.testing.assert_array_equal(o[3], outs[0][3])


def test_sampler_exact_at_nodes_K6():
    g = TwoDGrid(64)
    rng = np.random.default_rng(0)
    F = rng.standard_normal((64, 64, 5))
    prob = swrt.Problem(nx=64, f=3.0, dt=1e-3)
    raytracing.set_velocity_info(prob, 0, F)
    ii, jj = np.meshgrid(np.arange(0, 64, 16), np.arange(0, 64, 16), indexing="ij")
    xk = np.zeros((16, 4)); xk[:, 0] = g.x[ii.ravel()]; xk[:, 1] = g.y[jj.ravel()]
    pk = raytracing.Packets(prob, 16, 3.0, 1.0)
    pk.set(xk, np.ones(16))
    U = raytracing.interpolate_velocity(raytracing.Velocity(prob, 0), pk)
    # the kernel multiplies by 1/dx instead of dividing: a node can land one ulp beside the cell edge, which moves
    # the (continuous) interpolant by O(1e-16) -- floating-point tolerance, not index arithmetic
    np.testing.assert_allclose(U, F[ii.ravel(), jj.ravel(), 0:2], rtol=0, atol=1e-14)


def test_packet_generation_and_sharding_bit_exact_lattice():
    prob = swrt.Problem(nx=64, f=3.0, dt=1e-3)
    n_side, L, k0 = 12, 2 * np.pi, 5.196152422706632
    want, sign = oray.generate_initial_wavepackets(L, k0, n_side)
    N = n_side * n_side
    parts = []
    for r in range(3):                                            # three contiguous shards, like ranks
        lo, hi = r * N // 3, (r + 1) * N // 3
        pk = raytracing.generate_initial_wavepackets(prob, L, k0, hi - lo, n_side, 3.0, 1.0, first=lo)
        parts.append(pk.get())
    got = np.concatenate(parts, axis=0)
    np.testing.assert_array_equal(got[:, 0:2], want[:, 0:2])      # lattice: IEEE mul/div/sub only -> bit exact
    np.testing.assert_allclose(got[:, 2:4], want[:, 2:4], rtol=0, atol=4e-15)


def test_kcutoff_reset_bit_exact():
    prob = swrt.Problem(nx=64, f=3.0, dt=1e-3)
    rng = np.random.default_rng(2)
    xk = rng.standard_normal((1000, 4)) * 40
    xk[0, 2:4] = (30.0, 40.0)                                     # exactly on the cutoff circle -> reset (>=)
    pk = raytracing.Packets(prob, 1000, 3.0, 1.0)
    pk.set(xk, np.ones(1000))
    want = xk.copy()
    n_want = oray.kcutoff_reset(want, 50.0, 5.2)
    assert pk.kcutoff_reset(50.0, 5.2) == n_want
    np.testing.assert_array_equal(pk.get(), want)


def test_snapshot_aliasing_quirk_flag():
    g, p, sol0, c = config2_setup(64)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    raytracing.get_velocity_info(prob, 0)
    flow.stepforward(prob, (), 5)
    raytracing.get_velocity_info(prob, 1)
    old0, new0 = raytracing.Velocity(prob, 0)._arr(), raytracing.Velocity(prob, 1)._arr()
    assert np.abs(old0 - new0).max() > 0
    raytracing.swap_snapshots(prob, alias=False)
    np.testing.assert_array_equal(raytracing.Velocity(prob, 0)._arr(), new0)
    np.testing.assert_array_equal(raytracing.Velocity(prob, 1)._arr(), old0)
    raytracing.swap_snapshots(prob, alias=True)                   # reference rebinding: both names -> same buffer
    np.testing.assert_array_equal(raytracing.Velocity(prob, 0)._arr(), raytracing.Velocity(prob, 1)._arr())


def test_full_size_properties_2048():
    """BASELINE config-4 size: size-independent properties instead of an oracle run."""
    nx = 2048
    g, p, sol0, c = config2_setup(nx)
    sol0 = orsw.enforce_reality_condition(sol0, g, p)            # Hermitian-consistent kr=0 column (Parseval needs it)
    g.dealias(sol0)
    prob = swrt.Problem(nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    # Parseval: sum u^2 dx dy == parsevalsum2(uh)
    u = prob.vars.u
    ke = flow.kinetic_energy(prob)
    v = prob.vars.v
    assert abs(0.5 * ((u ** 2).sum() + (v ** 2).sum()) * g.dx * g.dy / (g.Lx * g.Ly) / ke - 1) < 1e-12
    assert rel_l2(u, g.irfft2(sol0[:, :, 0])) < 1e-13
    # one step against the oracle (a single 2048^2 oracle step takes a few seconds), then invariants over 20 more
    flow.stepforward(prob, (), 1)
    assert rel_l2(prob.sol, oracle_steps(g, p, sol0, c["dt"], 1)) < 1e-11
    e0 = flow.kinetic_energy(prob) + flow.potential_energy(prob)
    flow.stepforward(prob, (), 20)
    e1 = flow.kinetic_energy(prob) + flow.potential_energy(prob)
    assert abs(e1 / e0 - 1) < 1e-3 and not flow.has_nan(prob)
    sol = prob.sol
    assert np.all(sol[g.kr_alias[0]:] == 0) and np.all(sol[:, g.l_alias[0]:g.l_alias[1]] == 0)   # stays dealiased


# ------------------------------------------------------------------------------------------------ QG models
def _qg_state(nx, nlayers, seed):
    g, sol3 = random_state(nx, seed=seed, amp=1.0, slope=1.0)
    return g, np.ascontiguousarray(sol3[:, :, :nlayers])


@pytest.mark.parametrize("stepper", ["IFMAB3", "FilteredAB3"])
def test_swqg_parity(stepper):
    from oracle import ifmab3 as oif, qg as oqg
    nx, f, Cg, nnu, dt = 128, 3.0, 1.0, 4, 2e-3
    nu = 2 * np.pi / nx / ((nx / 2 - 1) ** (2 * nnu)) / dt
    g, sol0 = _qg_state(nx, 1, 5)
    sol0 = sol0[:, :, 0]
    Kd2 = f * f / (Cg * Cg)
    prob = swrt.Problem(model="SWQG", stepper=stepper, nx=nx, dt=dt, f=f, Cg=Cg, nu=nu, nnu=nnu)
    prob.sol = sol0
    np.testing.assert_array_equal(prob.sol, sol0)
    psi = g.irfft2(oqg.swqg_streamfunction(sol0, g, Kd2))
    assert rel_l2(prob.vars.ψ, psi) < 1e-13
    assert rel_l2(prob.vars.u, g.irfft2(-1j * g.l * oqg.swqg_streamfunction(sol0, g, Kd2))) < 1e-13
    ke, pe = oqg.swqg_energies(sol0, g, Kd2)
    assert abs(flow.kinetic_energy(prob) / ke - 1) < 1e-13 and abs(flow.potential_energy(prob) / pe - 1) < 1e-13
    L = oqg.swqg_L(g, nu, nnu)
    calcN = lambda s: oqg.swqg_calcN(s, g, Kd2)
    ts = oif.IFMAB3(L, dt, calcN) if stepper == "IFMAB3" else oqg.FilteredAB3(L, dt, calcN, makefilter(g))
    want = sol0.copy()
    for n in (1, 3, 46):
        flow.stepforward(prob, (), n)
        for _ in range(n):
            ts.stepforward(want)
        assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10, n
    vel, grad = raytracing.get_velocity_info(prob, 0, raytracing.PSI_SWQG)
    ref = oray.get_velocity_info(oqg.swqg_streamfunction(g.dealias(want.copy()), g, Kd2), g)
    assert rel_l2(vel._arr(), ref) < 1e-12


def test_twolayerqg_parity():
    from oracle import ifmab3 as oif, qg as oqg
    nx, U, mu, f0, Cg, drr, nnu, dt = 128, 0.5, 1e-2, 3.0, 1.0, 0.2, 4, 1e-3
    nu = 40 * 2 * np.pi / nx / ((nx / 2 - 1) ** (2 * nnu)) / dt
    F = 2 * f0 ** 2 / Cg ** 2 / drr
    g, sol0 = _qg_state(nx, 2, 9)
    prob = swrt.Problem(model="TwoLayerQG", nx=nx, dt=dt, U=U, mu=mu, f0=f0, Cg=Cg, δρρ0=drr, nu=nu, nnu=nnu)
    prob.sol = sol0
    psih = oqg.twolayer_streamfunction(sol0, g, F)
    assert rel_l2(prob.vars.ψ, np.stack([g.irfft2(psih[:, :, j]) for j in range(2)], axis=-1)) < 1e-13
    (ke1, ke2), pe = oqg.twolayer_energies(sol0, g, F)
    k1, k2 = flow.kinetic_energy(prob)
    assert abs(k1 / ke1 - 1) < 1e-12 and abs(k2 / ke2 - 1) < 1e-12 and abs(flow.potential_energy(prob) / pe - 1) < 1e-12
    L = oqg.twolayer_L(g, F, U, mu, nu, nnu)
    ts = oif.IFMAB3(L, dt, lambda s: oqg.twolayer_calcN(s, g, F))      # general matrix exponential, like the reference
    want = sol0.copy()
    for n in (1, 3, 46):
        flow.stepforward(prob, (), n)
        for _ in range(n):
            ts.stepforward(want)
        assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10, n
    for kind, comb in ((raytracing.PSI_TWOLAYER_BAROCLINIC, lambda p: 0.5 * (p[:, :, 0] - p[:, :, 1])),
                       (raytracing.PSI_TWOLAYER_MEAN, lambda p: (p[:, :, 0] + p[:, :, 1]) / 2)):
        vel, _ = raytracing.get_velocity_info(prob, 1, kind)
        ref = oray.get_velocity_info(comb(oqg.twolayer_streamfunction(g.dealias(want.copy()), g, F)), g)
        assert rel_l2(vel._arr(), ref) < 1e-12


# ------------------------------------------------------------------------------------------------ Thomas-Yamada + multi-stage steppers
@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredRK4", "IFMAB3"])
def test_thomasyamada_parity(stepper):
    from oracle import ifmab3 as oif, ty as oty
    nx, Lx, Ro, nnu, dt = 64, 6 * np.pi, 1.0, 8, 5e-3
    nu = 5e-34 * (Lx / (2 * np.pi)) ** 16 * 1e12          # thomasyamada/gpu-setup/Parameters.jl recipe, scaled so that it acts at 64^2
    g, s3 = random_state(nx, seed=21, amp=0.3, slope=1.0, Lx=Lx)
    _, s1 = random_state(nx, seed=22, amp=0.3, slope=1.0, Lx=Lx)
    sol0 = np.concatenate([s3, s1[:, :, :1]], axis=-1)
    prob = swrt.Problem(model="ThomasYamada", stepper=stepper, nx=nx, Lx=Lx, dt=dt, nu=nu, nnu=nnu, Ro=Ro)
    flow.set_solution(prob, *(sol0[:, :, i] for i in range(4)))
    np.testing.assert_array_equal(prob.sol, sol0)
    assert rel_l2(prob.vars.uc, g.irfft2(sol0[:, :, 1])) < 1e-13
    L = oty.ty_L(g, nu, nnu)
    calcN = lambda s: oty.ty_calcN(s, g, Ro)
    ts = {"ETDRK4": lambda: oty.ETDRK4(L, dt, calcN), "FilteredRK4": lambda: oty.FilteredRK4(L, dt, calcN, makefilter(g)),
          "IFMAB3": lambda: oif.IFMAB3(L, dt, calcN)}[stepper]()
    want = sol0.copy()
    for n in (1, 3, 26):
        flow.stepforward(prob, (), n)
        for _ in range(n):
            ts.stepforward(want)
        assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10, n
    ke, pe = flow.kinetic_energy(prob), flow.potential_energy(prob)
    from oracle.grid import parsevalsum2
    w = g.dealias(want.copy())
    assert abs(ke / (parsevalsum2(w[:, :, 1], g) + parsevalsum2(w[:, :, 2], g)) - 1) < 1e-9
    assert abs(pe / parsevalsum2(w[:, :, 3], g) - 1) < 1e-9


@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredRK4", "FilteredETDRK4"])
def test_swqg_multistage_steppers(stepper):
    from oracle import qg as oqg, ty as oty
    nx, f, Cg, nnu, dt = 64, 3.0, 1.0, 4, 4e-3
    nu = 2 * np.pi / nx / ((nx / 2 - 1) ** (2 * nnu)) / dt
    g, sol0 = _qg_state(nx, 1, 15)
    sol0 = sol0[:, :, 0]
    prob = swrt.Problem(model="SWQG", stepper=stepper, nx=nx, dt=dt, f=f, Cg=Cg, nu=nu, nnu=nnu)
    prob.sol = sol0
    L = oqg.swqg_L(g, nu, nnu)
    calcN = lambda s: oqg.swqg_calcN(s, g, f * f / (Cg * Cg))
    ts = {"ETDRK4": lambda: oty.ETDRK4(L, dt, calcN), "FilteredRK4": lambda: oty.FilteredRK4(L, dt, calcN, makefilter(g)),
          "FilteredETDRK4": lambda: oty.ETDRK4(L, dt, calcN, makefilter(g))}[stepper]()
    want = sol0.copy()
    flow.stepforward(prob, (), 25)
    for _ in range(25):
        ts.stepforward(want)
    assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10


def test_unsupported_combinations_fail_loudly():
    with pytest.raises(swrt.SwrtError):
        swrt.Problem(model="RotatingShallowWater", stepper="ETDRK4", nx=64)      # matrix L: FourierFlows steppers need a diagonal L
    with pytest.raises(swrt.SwrtError):
        swrt.Problem(nx=48)                                                      # not a power of two


# ------------------------------------------------------------------------------------------------ Hermite-bicubic interpolant
@pytest.mark.parametrize("nsub", [1, 3])
def test_hermite_bicubic_mode_parity(nsub):
    g, p, sol0, c = config2_setup(128)
    sol1 = oracle_steps(g, p, sol0, c["dt"], 4)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_interpolation(prob, raytracing.INTERP_HERMITE_BICUBIC)
    prob.sol = sol0
    vel, _ = raytracing.get_velocity_info(prob, 0)
    Fo = oray.get_velocity_info_cubic(orsw.get_streamfunction(sol0, g, p), g)
    got = vel._arr()
    assert got.shape[-1] == 7
    for k in range(7):
        assert rel_l2(got[:, :, k], Fo[:, :, k]) < 1e-12, k
    flow.stepforward(prob, (), 4)
    raytracing.get_velocity_info(prob, 1)
    Fn = oray.get_velocity_info_cubic(orsw.get_streamfunction(sol1, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 40)
    xk[:, 0:2] += np.random.default_rng(8).uniform(-20, 20, size=(xk.shape[0], 2))
    pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=nsub, interp=raytracing.INTERP_HERMITE_BICUBIC)
    pk.set(xk, sign)
    t0, t1 = 0.0, 4 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t0, t1))
    want = oray.raytrace(xk.copy(), sign, t0, t1, Fo, Fn, g, c["f"], c["Cg"], nsub=nsub)
    assert np.abs(pk.get() - want).max() / np.abs(want).max() < 1e-8
    assert rel_l2(pk.get(), want) < 1e-11
    U, G = oray.interpolate_velocity(Fn, want[:, 0:2], g)
    pk.set(want, sign)
    np.testing.assert_allclose(raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 1), pk), G, rtol=0, atol=1e-10)
    # a bilinear packet set on a flow holding Hermite node data must be refused, not silently mis-sampled
    pk2 = raytracing.Packets(prob, 16, c["f"], c["Cg"])
    with pytest.raises(swrt.SwrtError):
        raytracing.raytrace(pk2, None, None, None, None, prob.grid, pk2, c["dt"], (t0, t1))


# ------------------------------------------------------------------------------------------------ edge cases
def test_single_packet_and_far_outside_domain():
    """generate_single_wavepacket (raytracing/RaytracingDriver.jl:16-25): N = 1; packets are never wrapped (they drift to |x| ~ 60)."""
    g, c, Fo, Fn, _, _ = _packet_case(64, 4, 1)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_velocity_info(prob, 0, Fo)
    raytracing.set_velocity_info(prob, 1, Fn)
    for x0 in ((0.3, -1.2), (-61.7, 59.9), (1e4 + 0.25, -1e4 - 0.75)):
        xk = np.array([[x0[0], x0[1], 2.0, -4.5]])
        pk = raytracing.Packets(prob, 1, c["f"], c["Cg"], nsub=3)
        pk.set(xk, np.ones(1))
        raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (0.0, 0.2))
        want = oray.raytrace(xk.copy(), np.ones(1), 0.0, 0.2, Fo, Fn, g, c["f"], c["Cg"], nsub=3)
        np.testing.assert_allclose(pk.get(), want, rtol=1e-9, atol=1e-9)


def test_zero_flow_packets_move_with_group_velocity_only():
    prob = swrt.Problem(nx=64, f=3.0, Cg=1.0, dt=1e-2)
    xk, sign = oray.generate_initial_wavepackets(2 * np.pi, 5.0, 8)
    pk = raytracing.Packets(prob, 64, 3.0, 1.0, nsub=1)
    pk.set(xk, sign)
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, 1e-2, (0.0, 0.5))
    got = pk.get()
    w = sign * np.sqrt(9.0 + xk[:, 2] ** 2 + xk[:, 3] ** 2)
    np.testing.assert_allclose(got[:, 0], xk[:, 0] + 0.5 * xk[:, 2] / w, rtol=0, atol=1e-13)
    np.testing.assert_allclose(got[:, 1], xk[:, 1] + 0.5 * xk[:, 3] / w, rtol=0, atol=1e-13)
    np.testing.assert_array_equal(got[:, 2:4], xk[:, 2:4])          # k is exactly conserved without a background flow


def test_zero_state_stays_zero_and_aliased_fraction_zero():
    prob = swrt.Problem(nx=64, f=3.0, dt=1e-2)
    flow.stepforward(prob, (), 5)
    assert np.all(prob.sol == 0) and prob.clock.step == 5
    # aliased_fraction = 0 (all MultiLayerQG runs, raytracing/TwoLayerRaytracing.jl:174): only the Nyquist row/column is dropped
    g = TwoDGrid(64, aliased_fraction=0)
    rng = np.random.default_rng(3)
    sol = rng.standard_normal((g.nkr, g.nl, 3)) + 1j * rng.standard_normal((g.nkr, g.nl, 3))
    p0 = swrt.Problem(nx=64, f=3.0, dt=1e-3, aliased_fraction=0)
    p0.sol = sol
    want = g.dealias(sol.copy())
    np.testing.assert_array_equal(p0.sol, want)
    assert np.count_nonzero(want[:, :, 0]) == (g.nkr - 1) * (g.nl - 1)
    assert rel_l2(p0.vars.u, g.irfft2(want[:, :, 0])) < 2e-14


def test_rectangular_grid_step_parity():
    from oracle import ifmab3 as oif
    nx, ny, Lx, Ly = 128, 64, 2 * np.pi, np.pi
    g, sol0 = random_state(nx, ny, seed=33, amp=0.2, Lx=Lx, Ly=Ly)
    p = orsw.Params(1e-9, 4, 3.0, 1.0)
    prob = swrt.Problem(nx=nx, ny=ny, Lx=Lx, Ly=Ly, dt=2e-3, f=3.0, Cg=1.0, nu=1e-9, nnu=4)
    prob.sol = sol0
    flow.stepforward(prob, (), 12)
    assert rel_l2(prob.sol, oracle_steps(g, p, sol0, 2e-3, 12)) < 1e-11


def test_4096_transform_properties():
    """Largest supported size: Parseval and a round trip through the physical field (size-independent properties)."""
    nx = 4096
    g = TwoDGrid(nx)
    rng = np.random.default_rng(4)
    sol = np.zeros((g.nkr, g.nl, 3), dtype=np.complex128)
    sol[:40, :40] = rng.standard_normal((40, 40, 3)) + 1j * rng.standard_normal((40, 40, 3))
    sol[:40, -40:] = rng.standard_normal((40, 40, 3)) + 1j * rng.standard_normal((40, 40, 3))
    sol[0, :, :] = 0                                              # keep the kr = 0 column trivially Hermitian
    prob = swrt.Problem(nx=nx, f=3.0, dt=1e-4)
    prob.sol = sol
    u = prob.vars.u
    assert rel_l2(u, g.irfft2(sol[:, :, 0])) < 1e-13
    ke = flow.kinetic_energy(prob)
    v = prob.vars.v
    assert abs(0.5 * ((u ** 2).sum() + (v ** 2).sum()) * g.dx * g.dy / (g.Lx * g.Ly) / ke - 1) < 1e-12
    flow.stepforward(prob, (), 2)
    assert not flow.has_nan(prob)


def test_quadheight_variant_and_physical_forward_transform():
    nx = 64
    g, p, sol0, c = config2_setup(nx)
    sol0 = sol0 * 0.3                                            # keep 1 + eta well away from zero
    prob = swrt.Problem(model="QuadHeightModifiedShallowWater", nx=nx, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    flow.set_solution(prob, sol0[:, :, 0], sol0[:, :, 1], sol0[:, :, 2])
    want0 = g.dealias(orsw.quadheight_set_solution(sol0[:, :, 0], sol0[:, :, 1], sol0[:, :, 2], g))
    assert rel_l2(prob.sol, want0) < 1e-13                       # exercises the physical -> spectral path (R2C x-pass + forward y-pass)
    assert abs(flow.potential_energy(prob) - 0.5 * p.Cg2 * want0[0, 0, 2].real / (g.Lx * g.Ly)) < 1e-12
    flow.stepforward(prob, (), 30)
    assert rel_l2(prob.sol, oracle_steps(g, p, want0, c["dt"], 30, orsw.QUADHEIGHT)) < 1e-10
    # round trip of an arbitrary physical field through set_field_physical
    f = np.random.default_rng(1).standard_normal((nx, nx))
    flow.set_field_physical(prob, 0, f)
    assert rel_l2(prob.sol[:, :, 0], g.dealias(g.rfft2(f))) < 1e-13


# ------------------------------------------------------------------------------------------------ CPU-tracer semantics
@pytest.mark.parametrize("interp,integ", [(2, 0), (2, 1), (0, 1), (1, 1), (4, 0), (4, 1)])
def test_bspline_and_implicit_midpoint_modes(interp, integ):
    """Quadratic B-spline sampling and the implicit-midpoint integrator of raytracing/Raytracing.jl (a18), against the oracle."""
    g, p, sol0, c = config2_setup(128)
    sol1 = oracle_steps(g, p, sol0, c["dt"], 3)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_interpolation(prob, interp)
    prob.sol = sol0
    vel, _ = raytracing.get_velocity_info(prob, 0)
    flow.stepforward(prob, (), 3)
    raytracing.get_velocity_info(prob, 1)
    psi0, psi1 = orsw.get_streamfunction(sol0, g, p), orsw.get_streamfunction(sol1, g, p)
    if interp == 1:
        Fo, Fn, sampler = oray.get_velocity_info_cubic(psi0, g), oray.get_velocity_info_cubic(psi1, g), oray.sample_hermite
    elif interp == 2:
        Fo, Fn = (oray.bspline2_prefilter(oray.get_velocity_info(q, g), g) for q in (psi0, psi1))
        sampler = oray.sample_bspline2
    elif interp == 4:
        Fo, Fn = (oray.bspline3_prefilter(oray.get_velocity_info(q, g), g) for q in (psi0, psi1))
        sampler = oray.sample_bspline3
    else:
        Fo, Fn, sampler = oray.get_velocity_info(psi0, g), oray.get_velocity_info(psi1, g), oray.sample_bilinear
    assert rel_l2(vel._arr(), Fo) < 1e-12                         # the snapshot holds spline coefficients in mode 2
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 24)
    xk[:, 0:2] += np.random.default_rng(5).uniform(-20, 20, size=(xk.shape[0], 2))
    pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=2, interp=interp, integrator=integ)
    pk.set(xk, sign)
    t0, t1 = 0.0, 3 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (t0, t1))
    if integ == 1:
        want = oray.raytrace_midpoint(xk.copy(), sign, t0, t1, Fo, Fn, g, c["f"], c["Cg"], nsub=2, sampler=sampler)
    else:
        want = xk.copy()
        h = (t1 - t0) / 2
        for s_ in range(2):                                       # RK4 with the B-spline sampler
            def f_(z, al):
                return oray.rhs_sampler(z, sign, al, sampler(Fo, z[:, 0], z[:, 1], g), sampler(Fn, z[:, 0], z[:, 1], g), c["f"], c["Cg"])
            a0 = (s_ * h) / (t1 - t0)
            k1 = f_(want, a0); k2 = f_(want + 0.5 * h * k1, a0 + 0.5 * h / (t1 - t0)); k3 = f_(want + 0.5 * h * k2, a0 + 0.5 * h / (t1 - t0))
            k4 = f_(want + h * k3, a0 + h / (t1 - t0))
            want = want + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    got = pk.get()
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-8
    assert rel_l2(got, want) < 1e-10


def _config1_setup(nx, seed=1234):
    """BASELINE config 1 recipe (simulation/Parameters.jl:33-49, simulation/TwoLayerSimulation.jl:42-47) at a small size."""
    from oracle import qg as oqg
    f0, rd, lv, avg_U, H0, b2 = 1.0, 1 / 15, 1 / 2, 0.1, 1.0, 1.0
    l_star = lv / rd
    kappa = 0.36 / np.log(l_star / 3.2)
    U = avg_U / l_star
    mu = 2 * U * kappa / rd
    b1 = 4 * f0 ** 2 * rd ** 2 / H0 + b2
    g = TwoDGrid(nx, aliased_fraction=0)
    dt = 0.02 * g.dx / avg_U
    F = f0 ** 2 / ((b1 - b2) * H0 / 2)
    filt = makefilter(g)
    rng = np.random.default_rng(seed)
    q0 = 1e-2 * avg_U * rng.standard_normal((nx, nx, 2))
    sol0 = np.stack([g.rfft2(q0[:, :, j]) * filt for j in range(2)], axis=-1)
    # give the flow an O(avg_U) large-scale part so the nonlinear terms matter within a few steps
    sol0[1:6, 1:6] += 1.0 * nx * nx * avg_U * (rng.standard_normal((5, 5, 2)) + 1j * rng.standard_normal((5, 5, 2)))
    g.dealias(sol0)
    return g, sol0, dict(f0=f0, H=(H0 / 2, H0 / 2), b=(b1, b2), U=(U, -U), mu=mu, beta=0.0, dt=dt, F=F, filt=filt)


@pytest.mark.parametrize("stepper", ["FilteredAB3", "FilteredRK4", "ETDRK4"])
def test_multilayerqg2_parity(stepper):
    """GeophysicalFlows MultiLayerQG, two equal layers, aliased_fraction = 0 (config 1: raytracing/TwoLayerRaytracing.jl:174)."""
    from oracle import qg as oqg, ty as oty
    nx = 128
    g, sol0, c = _config1_setup(nx)
    nnu, nu = 4, 1e-14
    prob = swrt.Problem(model="MultiLayerQG", stepper=stepper, nx=nx, dt=c["dt"], f0=c["f0"], H=c["H"], b=c["b"], U=c["U"],
                        mu=c["mu"], beta=c["beta"], nu=nu, nnu=nnu, aliased_fraction=0)
    assert abs(prob.desc.F / c["F"] - 1) < 1e-15
    prob.sol = sol0
    np.testing.assert_array_equal(prob.sol, sol0)
    psih = oqg.twolayer_streamfunction(sol0, g, c["F"])
    assert rel_l2(prob.vars.ψ, np.stack([g.irfft2(psih[:, :, j]) for j in range(2)], axis=-1)) < 1e-13
    L = (-nu * g.Krsq ** nnu)[:, :, None] * np.ones(2)
    calcN = lambda s: oqg.multilayer2_calcN(s, g, c["F"], c["U"][0], c["U"][1], c["beta"], c["mu"])
    if stepper == "FilteredAB3":
        ts = oqg.FilteredAB3(L, c["dt"], calcN, c["filt"][:, :, None])
    elif stepper == "FilteredRK4":
        ts = oty.FilteredRK4(L, c["dt"], calcN, c["filt"])
    else:
        ts = oty.ETDRK4(L, c["dt"], calcN)
    want = sol0.copy()
    for n in (1, 3, 26):
        flow.stepforward(prob, (), n)
        for _ in range(n):
            ts.stepforward(want)
        assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10, n
    # the packet driver samples the layer-mean streamfunction (raytracing/TwoLayerRaytracing.jl:122)
    vel, _ = raytracing.get_velocity_info(prob, 0, raytracing.PSI_TWOLAYER_MEAN)
    psim = oqg.twolayer_streamfunction(g.dealias(want.copy()), g, c["F"]).mean(axis=-1)
    assert rel_l2(vel._arr(), oray.get_velocity_info(psim, g)) < 1e-11


def test_config1_twolayer_cpu_driver_parity():
    """BASELINE config 1 end to end at a small size: raytracing/TwoLayerRaytracing.jl's loop (MultiLayerQG + FilteredAB3,
    layer-mean streamfunction, quadratic B-spline x linear-in-time fields, implicit midpoint, k-cutoff reset)."""
    from oracle import qg as oqg
    from juliaraytracingsw_b200 import twolayer
    nx = 64
    g, sol0, c = _config1_setup(nx)
    P = twolayer.Parameters(nx=nx, sqrtNpackets=8, npacketsubs=5, total_time=1.0, k_cutoff=1.75)
    assert abs(P.dt - c["dt"]) < 1e-18
    prob, packets, frames, step = twolayer.start(P, qh=sol0, max_frames=2)
    assert len(frames) == 3 and prob.clock.step == 10 and step == 2 * (5 * 2 + 1)
    # oracle loop
    L = np.zeros((g.nkr, g.nl, 2))
    ts = oqg.FilteredAB3(L, c["dt"], lambda s: oqg.multilayer2_calcN(s, g, c["F"], c["U"][0], c["U"][1], 0.0, c["mu"]),
                         c["filt"][:, :, None])
    k0 = np.sqrt(3.0)
    xk, sign = oray.generate_initial_wavepackets_twolayer(P.L, k0, 8)
    np.testing.assert_array_equal(frames[0].x, xk[:, 0:2])
    np.testing.assert_array_equal(frames[0].k, xk[:, 2:4])
    fields = lambda s: oray.bspline2_prefilter(oray.get_velocity_info(oqg.twolayer_streamfunction(s, g, c["F"]).mean(axis=-1), g), g)
    want = sol0.copy()
    Fo, t, nreset = fields(want), 0.0, 0
    for fr in range(2):
        for _ in range(5):
            ts.stepforward(want)
            Fn = fields(g.dealias(want.copy()))
            xk = oray.raytrace_midpoint(xk, sign, t, ts.t, Fo, Fn, g, 1.0, 1.0, nsub=1, sampler=oray.sample_bspline2)
            nreset += oray.kcutoff_reset(xk, 1.75, k0)
            Fo, t = Fn, ts.t
        got = np.concatenate([frames[fr + 1].x, frames[fr + 1].k], axis=1)
        assert np.abs(got - xk).max() / np.abs(xk).max() < 1e-8, fr
        assert rel_l2(frames[fr + 1].u, oray.sample_bspline2(Fo, xk[:, 0], xk[:, 1], g)[:, 0:2]) < 1e-8
    assert rel_l2(prob.sol, g.dealias(want.copy())) < 1e-10
    assert nreset > 0                                              # the cutoff was exercised


def test_wave_balanced_projections_on_device():
    """rsw/RSWUtils.jl:5-64 and thomasyamada/TYUtils.jl:10-51 on the device, against the oracle restatement;
    K10 / K11 properties on the device results."""
    from oracle import decompose as od
    g, p, sol0, c = config2_setup(128)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    bal, wav = flow.wave_balanced_decomposition(prob)
    wb, ww = od.wave_balanced_decomposition(sol0, g, p)
    assert rel_l2(bal, wb) < 1e-14 and rel_l2(wav, ww) < 1e-13
    scale = np.abs(sol0).max()
    assert np.abs(1j * g.kr * bal[:, :, 0] + 1j * g.l * bal[:, :, 1]).max() < 1e-12 * scale * g.kr.max()          # K10: non-divergent
    assert np.abs(1j * g.kr * wav[:, :, 1] - 1j * g.l * wav[:, :, 0] - p.f * wav[:, :, 2]).max() < 1e-12 * scale * g.kr.max()   # K10: no PV
    (kw, pw), (kg, pg) = flow.wave_geostrophic_energy(prob)
    assert abs(kw / orsw.kinetic_energy(ww, g) - 1) < 1e-12 and abs(pw / orsw.potential_energy(ww, g, p) - 1) < 1e-12
    assert abs(kg / orsw.kinetic_energy(wb, g) - 1) < 1e-12 and abs(pg / orsw.potential_energy(wb, g, p) - 1) < 1e-12
    cw = od.rsw_weights(sol0, od.rsw_bases(g, p), p)
    for got, want in zip(flow.compute_balanced_wave_weights(prob), cw):
        assert rel_l2(got, g.dealias(want.copy())) < 1e-13
    # Thomas-Yamada
    from oracle.grid import parsevalsum2
    rng = np.random.default_rng(8)
    gt = TwoDGrid(64, 6 * np.pi)
    s4 = gt.dealias(rng.standard_normal((gt.nkr, gt.nl, 4)) + 1j * rng.standard_normal((gt.nkr, gt.nl, 4)))
    pt = swrt.Problem(model="ThomasYamada", stepper="ETDRK4", nx=64, Lx=6 * np.pi, dt=5e-3, Ro=1.0, nu=1e-20, nnu=8)
    pt.sol = s4
    G, W = flow.wave_balanced_decomposition(pt)
    Go, Wo = od.ty_decompose(s4, gt)
    assert rel_l2(G, Go) < 1e-13 and rel_l2(W, Wo) < 1e-13
    assert np.abs(G + W - s4[:, :, 1:4]).max() < 1e-12                                   # K11: the projection is complete
    (kw, pw), (kg, pg) = flow.wave_geostrophic_energy(pt)
    assert abs(kw / (parsevalsum2(Wo[:, :, 0], gt) + parsevalsum2(Wo[:, :, 1], gt)) - 1) < 1e-12
    assert abs(pg / parsevalsum2(Go[:, :, 2], gt) - 1) < 1e-12
    assert abs(flow.barotropic_energy(pt) / parsevalsum2(np.sqrt(gt.invKrsq) * s4[:, :, 0], gt) - 1) < 1e-12


def test_packet_pipeline_matches_single_handle():
    """Chunked, stream-overlapped packet I/O (transfers on the flow's upload / download streams, kernels on the
    blocks' own streams) gives bit-identical results to one synchronous handle."""
    import torch
    g, p, sol0, c = config2_setup(128)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 50)         # 2500 packets, 7 ragged chunks
    xk[:, 0:2] += np.random.default_rng(2).uniform(-3, 3, size=(xk.shape[0], 2))
    n = xk.shape[0]
    pin = lambda *shape: torch.empty(shape[::-1], dtype=torch.float64).pin_memory().numpy().T
    h_in, h_out, h_U, h_G = pin(n, 4), pin(n, 4), pin(n, 2), pin(n, 4)
    h_sign = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    h_in[:], h_sign[:] = xk, sign
    single = raytracing.Packets(prob, n, c["f"], c["Cg"], nsub=2)
    pipe = raytracing.PacketPipeline(prob, n, c["f"], c["Cg"], nchunks=7, nsub=2)
    assert pipe.bounds[0][0] == 0 and pipe.bounds[-1][1] == n and all(a[1] == b[0] for a, b in zip(pipe.bounds, pipe.bounds[1:]))
    raytracing.get_velocity_info(prob, 0)
    t = 0.0
    for step in range(4):
        flow.stepforward(prob, (), 1)
        raytracing.get_velocity_info(prob, 1)
        t1 = prob.clock.t
        single.set(h_in.copy(), sign)
        raytracing.raytrace(single, None, None, None, None, prob.grid, single, c["dt"], (t, t1))
        pipe.step(h_in, h_sign, (t, t1), h_out, h_U, h_G, after_raytrace=lambda: raytracing.swap_snapshots(prob), sample_slot=0)
        want = single.get()
        np.testing.assert_array_equal(h_out, want)
        G = raytracing.interpolate_gradients(raytracing.VelocityGradient(prob, 0), single, output_U=(U := np.empty((n, 2), order="F")))
        np.testing.assert_array_equal(h_U, U)
        np.testing.assert_array_equal(h_G, G)
        h_in[:] = h_out
        t = t1
    pipe.close()


def test_komega_accumulator_parity():
    """SURVEY 8f.3 (config 3's k-omega output): device-resident series + windowed transforms in time against the oracle's
    restatement of thomasyamada/TY_k_omega.jl and rsw/fourier-analysis/mrsw/FourierRSW.jl."""
    from oracle import komega as okw, ty as oty
    from juliaraytracingsw_b200 import komega
    # Thomas-Yamada, 13 frames (a length that is no power of two) at kr index 4
    nx, Lx, dt, nnu, nu, Ro = 64, 6 * np.pi, 5e-3, 8, 1e-20, 1.0
    gt = TwoDGrid(nx, Lx)
    rng = np.random.default_rng(12)
    s4 = np.zeros((gt.nkr, gt.nl, 4), dtype=np.complex128)
    s4[:8, :8] = 20.0 * (rng.standard_normal((8, 8, 4)) + 1j * rng.standard_normal((8, 8, 4)))
    s4[:8, -8:] = 20.0 * (rng.standard_normal((8, 8, 4)) + 1j * rng.standard_normal((8, 8, 4)))
    s4[0] = 0
    pt = swrt.Problem(model="ThomasYamada", stepper="ETDRK4", nx=nx, Lx=Lx, dt=dt, Ro=Ro, nu=nu, nnu=nnu)
    pt.sol = s4
    kw = komega.KOmega(pt, k_idx=5, max_frames=32)
    ts = oty.ETDRK4(oty.ty_L(gt, nu, nnu), dt, lambda s: oty.ty_calcN(s, gt, Ro))
    want, rows, times = s4.copy(), [], []
    for fr in range(13):
        kw.append()
        rows.append(okw.ty_series(gt.dealias(want.copy()), gt, 4))
        times.append(ts.t)
        flow.stepforward(pt, (), 2)
        for _ in range(2):
            ts.stepforward(want)
    series = np.stack(rows)                                      # (T, 6, nl)
    assert kw.nframes == 13 and np.allclose(kw.t, times, rtol=0, atol=1e-15)
    for j in range(6):
        assert rel_l2(kw.series(j), series[:, j]) < 1e-10, j
    for j, ref in enumerate(okw.ty_spectra(series)):
        assert rel_l2(kw.spectrum(j), ref) < 1e-10, kw.names[j]
    # RSW, twelve detrended series at kr index 3
    g, p, sol0, c = config2_setup(64)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    prob.sol = sol0
    kr = komega.KOmega(prob, k_idx=4, max_frames=10)
    rows, times = [], []
    for fr in range(10):
        flow.stepforward(prob, (), 3)
        kr.append()
        times.append(prob.clock.t)
        rows.append(okw.rsw_series(prob.sol, g, p, 3))          # the flow itself is checked elsewhere: project the device state
    series, t = np.stack(rows), np.array(times)
    w = okw.hann(10)
    for j in range(12):
        assert rel_l2(kr.series(j), series[:, j]) < 1e-12, j
        assert rel_l2(kr.spectrum(j), okw.clean_fft(t, series[:, j], w)) < 1e-10, kr.names[j]


def test_problems_with_different_dealiasing_in_one_process():
    """The staged y-passes size their shared-memory buffer by the number of retained rows; two problems of one grid size but
    different aliased fractions must both launch (regression: the opt-in size was cached from the first problem)."""
    g3, sol3 = random_state(128, seed=4, amp=0.1)
    pa = swrt.Problem(nx=128, f=3.0, dt=1e-3, aliased_fraction=1 / 3)
    pa.sol = sol3
    flow.stepforward(pa, (), 2)
    raytracing.get_velocity_info(pa, 0)
    pb = swrt.Problem(nx=128, f=3.0, dt=1e-3, aliased_fraction=0)
    pb.sol = sol3
    flow.stepforward(pb, (), 2)
    raytracing.get_velocity_info(pb, 0)
    pc = swrt.Problem(nx=128, f=3.0, dt=1e-3, aliased_fraction=1 / 2)
    pc.sol = sol3
    flow.stepforward(pc, (), 2)
    raytracing.get_velocity_info(pc, 0)
    flow.stepforward(pa, (), 1)
    assert not (flow.has_nan(pa) or flow.has_nan(pb) or flow.has_nan(pc))


def test_fp32_packet_mode():
    """North star's optional fp32 packet mode: Float32 node data and right-hand side, fp64 state.  Tolerance 2e-6 relative
    against the fp64 oracle fed the same Float32-rounded fields (fp32 arithmetic: ~1e-7 per evaluation)."""
    g, p, sol0, c = config2_setup(128)
    sol1 = oracle_steps(g, p, sol0, c["dt"], 3)
    prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_interpolation(prob, raytracing.INTERP_BILINEAR_F32)
    prob.sol = sol0
    vel, _ = raytracing.get_velocity_info(prob, 0)
    flow.stepforward(prob, (), 3)
    raytracing.get_velocity_info(prob, 1)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g).astype(np.float32)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(sol1, g, p), g).astype(np.float32)
    got_F = vel._arr()
    assert np.abs(got_F - Fo).max() <= 2.0 ** -23 * np.abs(Fo).max()          # the snapshot holds the fields rounded to fp32
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 32)
    xk[:, 0:2] += np.random.default_rng(7).uniform(-20, 20, size=(xk.shape[0], 2))
    pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=2, interp=raytracing.INTERP_BILINEAR_F32)
    pk.set(xk, sign)
    t1 = 3 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (0.0, t1))
    want = oray.raytrace(xk.copy(), sign, 0.0, t1, Fo.astype(np.float64), Fn.astype(np.float64), g, c["f"], c["Cg"], nsub=2)
    got = pk.get()
    d = got - want
    d[:, 0:2] = (d[:, 0:2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() / np.abs(want).max() < 2e-6
    U = raytracing.interpolate_velocity(raytracing.Velocity(prob, 1), pk)
    Uo, _ = oray.interpolate_velocity(Fn.astype(np.float64), got[:, 0:2], g)
    assert np.abs(U - Uo).max() < 1e-6 * np.abs(Uo).max()
    # mixing modes is refused
    pk64 = raytracing.Packets(prob, 16, c["f"], c["Cg"])
    with pytest.raises(swrt._lib.SwrtError):
        raytracing.raytrace(pk64, None, None, None, None, prob.grid, pk64, c["dt"], (0.0, t1))


@pytest.mark.parametrize("interp", [0, 2, 3])
def test_refined_snapshots_fft_interpolation(interp):
    """"FFT interpolation": snapshots on a node grid twice finer than the flow's (spectral zero padding) against the oracle's
    zero-padded transform; ray tracing on the refined grid; and the refined interpolant is closer to the exact trigonometric one."""
    g, p, sol0, c = config2_setup(64)
    sol1 = oracle_steps(g, p, sol0, c["dt"], 3)
    prob = swrt.Problem(nx=64, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    raytracing.set_interpolation(prob, interp)
    raytracing.set_snapshot_refinement(prob, 2)
    prob.sol = sol0
    vel, _ = raytracing.get_velocity_info(prob, 0)
    flow.stepforward(prob, (), 3)
    raytracing.get_velocity_info(prob, 1)
    F = []
    for s in (sol0, sol1):
        psif, gf = oray.refine_streamfunction(orsw.get_streamfunction(s, g, p), g, 2)
        Ff = oray.get_velocity_info(psif, gf)
        F.append(oray.bspline2_prefilter(Ff, gf) if interp == 2 else Ff)
    Fo, Fn = F
    got = vel._arr()
    assert got.shape == (128, 128, 5)
    tol = 1e-6 if interp == 3 else 1e-12
    assert rel_l2(got, Fo) < tol
    # the refined nodes that coincide with the flow's grid carry the unrefined values
    if interp == 0:
        assert rel_l2(got[::2, ::2], oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)) < 1e-12
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 20)
    xk[:, 0:2] += np.random.default_rng(3).uniform(-20, 20, size=(xk.shape[0], 2))
    pk = raytracing.Packets(prob, xk.shape[0], c["f"], c["Cg"], nsub=2, interp=interp)
    pk.set(xk, sign)
    t1 = 3 * c["dt"]
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, c["dt"], (0.0, t1))
    if interp == 2:
        want, h = xk.copy(), t1 / 2
        for s_ in range(2):
            f_ = lambda z, al: oray.rhs_sampler(z, sign, al, oray.sample_bspline2(Fo, z[:, 0], z[:, 1], gf), oray.sample_bspline2(Fn, z[:, 0], z[:, 1], gf), c["f"], c["Cg"])
            a0 = s_ * h / t1
            k1 = f_(want, a0); k2 = f_(want + 0.5 * h * k1, a0 + 0.5 * h / t1); k3 = f_(want + 0.5 * h * k2, a0 + 0.5 * h / t1)
            k4 = f_(want + h * k3, a0 + h / t1)
            want = want + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    else:
        cast = (lambda A: A.astype(np.float32).astype(np.float64)) if interp == 3 else (lambda A: A)
        want = oray.raytrace(xk.copy(), sign, 0.0, t1, cast(Fo), cast(Fn), gf, c["f"], c["Cg"], nsub=2)
    d = pk.get() - want
    assert np.abs(d).max() / np.abs(want).max() < (2e-6 if interp == 3 else 1e-8)
    if interp == 0:
        # exact trigonometric value of u at random points vs bilinear on the coarse and on the refined grid
        rng = np.random.default_rng(9)
        px, py = rng.uniform(-np.pi, np.pi, 200), rng.uniform(-np.pi, np.pi, 200)
        psih = orsw.get_streamfunction(sol0, g, p)
        uh = -1j * g.l * psih
        w = np.where((np.arange(g.nkr) == 0) | (np.arange(g.nkr) == g.nkr - 1), 1.0, 2.0)[:, None]
        phase = np.exp(1j * (g.kr[:, :, None] * (px - g.x[0])[None, None, :] + g.l[:, :, None] * (py - g.y[0])[None, None, :]))
        exact = ((w * uh)[:, :, None] * phase).real.sum(axis=(0, 1)) / (g.nx * g.ny)
        Fc = oray.get_velocity_info(psih, g)
        e_coarse = np.abs(oray.sample_bilinear(Fc, px, py, g)[:, 0] - exact).max()
        e_fine = np.abs(oray.sample_bilinear(Fo, px, py, gf)[:, 0] - exact).max()
        assert e_fine < 0.4 * e_coarse
        with pytest.raises(swrt._lib.SwrtError):               # 4096 is the largest transform
            raytracing.set_snapshot_refinement(swrt.Problem(nx=4096, f=3.0, dt=1e-4), 2)


def test_refined_snapshot_large_grid_property():
    """2048^2 flow, 4096^2 node grid (the largest transform): the refined nodes that coincide with the flow's grid carry the
    unrefined snapshot exactly to round-off -- a size-independent property of spectral zero padding."""
    g, p, sol0, c = config2_setup(2048)
    pa = swrt.Problem(nx=2048, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
    pa.sol = sol0
    coarse = raytracing.get_velocity_info(pa, 0)[0]._arr()
    raytracing.set_snapshot_refinement(pa, 2)
    fine = raytracing.get_velocity_info(pa, 0)[0]._arr()
    assert fine.shape == (4096, 4096, 5)
    assert rel_l2(fine[::2, ::2], coarse) < 1e-12
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 64)
    pk = raytracing.Packets(pa, xk.shape[0], c["f"], c["Cg"])
    pk.set(xk, sign)
    raytracing.get_velocity_info(pa, 1)
    raytracing.raytrace(pk, None, None, None, None, pa.grid, pk, c["dt"], (0.0, c["dt"]))
    U = raytracing.interpolate_velocity(raytracing.Velocity(pa, 0), pk)
    assert np.isfinite(pk.get()).all() and np.abs(U).max() <= np.abs(fine[:, :, 0:2]).max() * (1 + 1e-12)


def test_committed_golden_vectors():
    """The CUDA path against the committed fixture tests/golden/rsw64_config2.npz (spectral state 1e-10, packets 1e-8)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rsw64_config2.npz"))
    L_, dt, f, Cg, nu, nnu, k0 = G["params"]
    prob = swrt.Problem(nx=64, Lx=L_, dt=dt, f=f, Cg=Cg, nu=nu, nnu=int(nnu))
    prob.sol = G["sol0"]
    flow.stepforward(prob, (), 10)
    assert rel_l2(prob.sol, G["sol10"]) < 1e-10
    vel, _ = raytracing.get_velocity_info(prob, 0)
    assert rel_l2(vel._arr(), G["snapshot10"]) < 1e-11
    pk = raytracing.Packets(prob, G["xk0"].shape[0], f, Cg, nsub=3)
    pk.set(G["xk0"], G["sign"])
    flow.stepforward(prob, (), 3)
    assert rel_l2(prob.sol, G["sol13"]) < 1e-10
    raytracing.get_velocity_info(prob, 1)
    raytracing.raytrace(pk, None, None, None, None, prob.grid, pk, dt, (10 * dt, 13 * dt))
    assert np.abs(pk.get() - G["xk1"]).max() / np.abs(G["xk1"]).max() < 1e-8


@pytest.mark.parametrize("nsteps,sort_every", [(11, 4), (47, 16), (30, 0)])
def test_coupled_steps_entry_point_is_the_python_loop(nsteps, sort_every):
    """swrt_packets_coupled_steps = n x drivers.coupled_step: flow state and clock bit for bit; packets to rounding (the fused loop
    traces every step over (0, dt) instead of the absolute (old_t, new_t)).  The longer runs go through the six-step CUDA graphs,
    with sorts (which swap the packet buffers) in between."""
    g, p, sol0, c = config2_setup(128)
    res = []
    for fused in (False, True):
        prob = swrt.Problem(nx=128, Lx=c["L"], dt=c["dt"], f=c["f"], Cg=c["Cg"], nu=c["nu"], nnu=c["nnu"])
        prob.sol = sol0
        pk = raytracing.generate_initial_wavepackets(prob, c["L"], c["k0"], 900, 30, c["f"], c["Cg"], sort_every=sort_every)
        raytracing.get_velocity_info(prob, 0)
        t = prob.clock.t
        if fused:
            t = drivers.coupled_steps(prob, pk, nsteps)
        else:
            for _ in range(nsteps):
                t = drivers.coupled_step(prob, pk, t)
        res.append((prob.sol, pk.get(), t, prob.clock.step))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-11 * np.abs(res[0][1]).max()
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3] == nsteps


def test_bench_line_contract_small_grid():
    """`bench.py` end to end on a small grid: one JSON line carrying every key of the measurement contract."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--nx", "256", "--sqrt-packets", "256", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["fp32_packet_mode"]["value"] > 0
