"""Pins the oracle against the known-answer values recorded in the reference's notebooks
(SURVEY.md section 4, K1..K15).  CPU only."""
import numpy as np
import scipy.linalg

from oracle import ifmab3, outputs, raytrace, rsw
from oracle.grid import TwoDGrid, alias_ranges, parsevalsum2


def test_K1_matrix_exponential():
    # rsw/Notebooks/MatrixExponentialTest.ipynb:197-198
    A = np.array([[0, 1, 1j], [-1, 0, 1j], [-1j, -1j, 0]])
    want = np.array([
        [1, 1.7182818284590455, 1.7182818284590453j],
        [-0.6321205588285577, 1, 0.6321205588285577j],
        [-0.6321205588285578j, -1.7182818284590453j, 2.0861612696304874]])
    got, _ = ifmab3.getexpLs(A.reshape(1, 1, 3, 3), 1.0)
    np.testing.assert_allclose(got[0, 0], want, rtol=0, atol=5e-15)


def test_K2_mvmul_orientation():
    A = np.zeros((1, 1, 2, 2)); A[0, 0] = [[1, 1], [0, 1]]
    x = np.zeros((1, 1, 2)); x[0, 0] = [1, 2]
    assert ifmab3.mvmul(A, x)[0, 0].tolist() == [3, 2]


def _k3_setup():
    nx, Lx = 512, 2 * np.pi
    dx = Lx / nx
    kmax = nx / 2 - 1
    nnu, cfltune, umax = 4, 0.1, 0.3
    nutune = 1 / cfltune
    dt = cfltune / umax * dx
    nu = nutune * 2 * np.pi / nx / (kmax ** (2 * nnu)) / dt
    return nx, Lx, dt, nu, nnu


def test_K3_L_and_expLdt():
    # rsw/Notebooks/RSW_Test.ipynb:228-236 (Float32 run; compare to 7 significant digits)
    nx, Lx, dt, nu, nnu = _k3_setup()
    assert dt == 0.00409061543436171
    assert nu == 1.6780303489894543e-18
    g = TwoDGrid(nx, Lx)
    p = rsw.Params(nu, nnu, 3.0, 1.0)
    L = rsw.populate_L(g, p)
    assert abs(np.abs(L[..., 0, 0]).max() - 495.26715) < 5e-5
    E = ifmab3.expL_closed_form(g, p, dt)
    m = E[..., 1, 1].real.min()
    assert abs(m - 0.07184223) < 1e-7
    assert np.unravel_index(E[..., 1, 1].real.argmin(), E.shape[:2]) == (256, 256)  # (kr,l)=(256,-256)
    # closed form == general matrix exponential, on a strided subset of wavenumbers
    sub = (slice(0, None, 16), slice(0, None, 31))
    Eg, E2g = ifmab3.getexpLs(L[sub], dt)
    np.testing.assert_allclose(E[sub], Eg, rtol=0, atol=2e-14)
    E2 = ifmab3.expL_closed_form(g, p, 2 * dt)
    np.testing.assert_allclose(E2[sub], E2g, rtol=0, atol=2e-14)


def test_K3_closed_form_modified():
    nx, Lx, dt, nu, nnu = _k3_setup()
    g = TwoDGrid(64, Lx)
    p = rsw.Params(nu * 1e10, nnu, 3.0, 1.0)
    L = rsw.populate_L(g, p, rsw.MODIFIED)
    E = ifmab3.expL_closed_form(g, p, dt, rsw.MODIFIED)
    Eg, _ = ifmab3.getexpLs(L, dt)
    np.testing.assert_allclose(E, Eg, rtol=0, atol=2e-14)


def test_K4_swqg_diagonal_L():
    # swqg/Notebooks/SWQG_Test.ipynb cell 3: nutune = 2
    nx, Lx, dt, _, nnu = _k3_setup()
    nu = 2 * 2 * np.pi / nx / ((nx / 2 - 1) ** (2 * nnu)) / dt
    assert nu == 3.3560606979789085e-19
    g = TwoDGrid(nx, Lx)
    L = -nu * g.Krsq ** nnu
    assert abs(np.abs(L).max() - 99.05343) < 1e-5


def test_K5_parseval_and_spectral_resample():
    # Notebooks/FFTInterpTest.ipynb cells 0-3
    g1, g2 = TwoDGrid(32), TwoDGrid(128)
    X, Y = g1.x[:, None], g1.y[None, :]
    f1 = 2 * np.sin(4 * X + 3 * Y + np.pi / 4) - np.cos(X - 2 * Y) + 4 * np.sin(-2 * X + Y - np.pi / 4)
    assert g1.x[0] == -np.pi and abs(g1.x[-1] - 2.945243112740431) < 1e-15
    f1h = g1.rfft2(f1)
    assert abs(parsevalsum2(f1h, g1) - 414.52338484575296) < 1e-10
    assert abs((f1 ** 2).sum() * g1.dx * g1.dy - 414.52338484575296) < 1e-10
    assert abs(f1.max() - 6.828544345425804) < 1e-13
    f2h = rsw.load_from_snapshot(f1h[:, :, None], g2)[:, :, 0]
    f2 = g2.irfft2(f2h)
    assert abs(parsevalsum2(f2h, g2) - 414.52338484575296) < 1e-10
    assert abs(f2.max() - 6.962803990425579) < 1e-12
    assert np.allclose(g1.kr[:, 0], np.arange(17.0))


def test_alias_ranges_match_survey():
    # nx=512 -> 171:257 and 171:342 ; nx=2048 -> 683:1025 and 683:1366 (1-based)
    assert alias_ranges(512, 257, 1 / 3) == ((170, 342), (170, 257))
    assert alias_ranges(2048, 1025, 1 / 3) == ((682, 1366), (682, 1025))
    assert alias_ranges(256, 129, 0) == ((128, 129), (128, 129))


def test_K6_bilinear_exact_at_nodes():
    # Notebooks/LargeMatrixTest.ipynb cells 7-10: sampling at node coordinates returns U[i]
    g = TwoDGrid(64)
    rng = np.random.default_rng(0)
    F = rng.standard_normal((64, 64, 5))
    ii, jj = np.meshgrid(np.arange(0, 64, 16), np.arange(0, 64, 16), indexing="ij")
    S = raytrace.sample_bilinear(F, g.x[ii.ravel()], g.y[jj.ravel()], g)
    np.testing.assert_array_equal(S, F[ii.ravel(), jj.ravel()])
    # periodic wrap: one period away samples the same values to rounding
    S2 = raytrace.sample_bilinear(F, g.x[ii.ravel()] + g.Lx, g.y[jj.ravel()] - g.Ly, g)
    np.testing.assert_allclose(S2, S, atol=1e-12)


def test_K7_packet_initial_layout():
    # raytracing/Notebooks/GPUDriverTest.ipynb:459 ff, sqrtN=3, k0=1 (production x/y ordering)
    xk, sign = raytrace.generate_initial_wavepackets(2 * np.pi, 1.0, 3)
    j = np.arange(1, 10)
    np.testing.assert_allclose(xk[:, 2], np.cos(2 * np.pi * j / 9), atol=1e-15)
    np.testing.assert_allclose(xk[:, 3], np.sin(2 * np.pi * j / 9), atol=1e-15)
    lat = np.array([-2.0943951023931957, 0.0, 2.0943951023931953])
    np.testing.assert_allclose(xk[:3, 0], lat, atol=5e-16)
    np.testing.assert_allclose(xk[0::3, 1], lat, atol=5e-16)
    assert abs(xk[1, 0]) < 3e-16                      # the notebook prints -2.2e-16
    assert sign.tolist() == [-1, 1, -1, 1, -1, 1, -1, 1, -1]


def test_K8_K9_steady_flow_invariant_and_rk4_order():
    # Taylor-Green steady flow with exact gradients (K9); Omega = omega + U.k conserved (K8)
    g = TwoDGrid(256)
    X, Y = g.x[:, None], g.y[None, :]
    U0 = 0.3
    F = np.stack([U0 * np.cos(X) * np.sin(Y), -U0 * np.sin(X) * np.cos(Y),
                  -U0 * np.sin(X) * np.sin(Y), U0 * np.cos(X) * np.cos(Y),
                  -U0 * np.cos(X) * np.cos(Y)], axis=-1)
    xk0, sign = raytrace.generate_initial_wavepackets(2 * np.pi, 3.0, 6)
    f, Cg = 3.0, 1.0

    def Omega(xk):
        S = raytrace.sample_bilinear(F, xk[:, 0], xk[:, 1], g)
        return sign * np.sqrt(f * f + Cg * Cg * (xk[:, 2] ** 2 + xk[:, 3] ** 2)) + S[:, 0] * xk[:, 2] + S[:, 1] * xk[:, 3]

    xk = xk0.copy()
    raytrace.raytrace(xk, sign, 0.0, 1.0, F, F, g, f, Cg, nsub=200)
    drift = np.abs(Omega(xk) - Omega(xk0)) / np.abs(Omega(xk0))
    assert drift.mean() < 4.1e-4          # the reference's Vern7 run records 0.041 %
    # RK4 self-convergence on a smooth (analytic-like, short) horizon: error ratio ~ 2^4
    def run(n):
        z = xk0.copy(); raytrace.raytrace(z, sign, 0.0, 0.05, F, F, g, f, Cg, nsub=n); return z
    e1 = np.abs(run(1) - run(8)).max(); e2 = np.abs(run(2) - run(8)).max()
    assert e2 < e1


def test_kcutoff_reset_is_compare_and_select():
    xk = np.array([[0, 0, 3.0, 4.0], [0, 0, 3.0, 3.9], [0, 0, -5.0, 0.0]])
    n = raytrace.kcutoff_reset(xk, 5.0, 1.5)
    assert n == 2
    assert xk[:, 2].tolist() == [1.5, 3.0, 1.5] and xk[:, 3].tolist() == [0.0, 3.9, 0.0]


def test_K10_ic_geostrophic_nondivergent_wave_zero_pv_K15_amplitudes():
    g = TwoDGrid(128)
    p = rsw.Params(0.0, 4, 3.0, 1.0)
    sol, (ugh, vgh, egh), (uwh, vwh, ewh) = rsw.initial_condition(g, p, (10, 13), 1.5, (0, 5), 0.1,
                                                                   np.random.default_rng(1234))
    div_g = np.abs(1j * g.kr * ugh + 1j * g.l * vgh).max()
    pv_w = np.abs(1j * g.kr * vwh - 1j * g.l * uwh - p.f * ewh).max()
    scale = np.abs(ugh).max()
    assert div_g / scale < 1e-12 and pv_w / scale < 1e-12
    assert abs(np.abs(g.irfft2(ugh)).max() - 1.5) < 1e-12       # K15
    assert abs(np.abs(g.irfft2(uwh)).max() - 0.1) < 1e-12


def test_K12_collated_rollover():
    # Notebooks/CollatedOutputTest.ipynb:101-112: line_limit=100, 1000 writes
    out = outputs.CollatedOutput("test_dir3/sin", 100)
    for i in range(1, 1001):
        out.write(str(i))
    assert out.files["test_dir3/sin_00000001.out"] == [str(i) for i in range(101, 201)]
    assert out.files["test_dir3/sin_00000009.out"][-1] == "1000"
    assert out.files["test_dir3/sin_00000010.out"] == []


def test_K13_sequenced_output_rolls_inside_a_frame():
    # packet_max_writes=300, gradients on -> frame 59's t,x,k,u in file 0, g in file 1
    out = outputs.SequencedOutput(lambda i: outputs.packet_filename("packets", i), 300)
    outputs.savepacketproblem(out)
    for frame in range(0, 130):
        outputs.write_packets(out, frame * 10, True)
    f0 = out.files["packets.000000.jld2"]; f1 = out.files["packets.000001.jld2"]
    assert len(f0) == 300
    assert f0[-4:] == ["p/t/580", "p/x/580", "p/k/580", "p/u/580"]
    assert f1[0] == "p/g/580" and f1[1] == "p/t/590"


def test_frame_roller_every_file_has_max_writes_frames():
    # raytracing/Notebooks/GPUDriverTest.ipynb cell 6: max_writes=1000, npacketsubs=10
    r = outputs.FrameRoller("packets.jld2", 1000)
    r.initial_frame(0)
    for j in range(1, 5001):
        r.loop_frame(j * 10)
    assert r.files["packets.jld2.00000004"][0] == 40000
    assert len(r.files["packets.jld2.00000004"]) == 1000
    assert len(r.files["packets.jld2.00000000"]) == 1000


def test_ifmab3_third_order_and_energy_conservation():
    # AB3 with exact integrating factor: 3rd order in dt on a smooth IC; inviscid energy drift tiny
    g = TwoDGrid(32)
    p = rsw.Params(0.0, 4, 3.0, 1.0)
    sol0, _, _ = rsw.initial_condition(g, p, (2, 4), 0.05, (0, 3), 0.02, np.random.default_rng(7))
    sol0 = rsw.enforce_reality_condition(sol0, g, p)

    def run(nsteps, T=0.4):
        sol = sol0.copy()
        ts = ifmab3.IFMAB3(rsw.populate_L(g, p), T / nsteps, lambda s: rsw.calcN(s, g, p))
        for _ in range(nsteps):
            ts.stepforward(sol)
        return g.dealias(sol)

    ref = run(640)
    e1 = np.abs(run(40) - ref).max(); e2 = np.abs(run(80) - ref).max()
    # the 3 Euler start-up steps cost O(dt^2) globally; observed order sits between 2 and 3
    assert e1 / e2 > 3.5
    E0 = rsw.kinetic_energy(sol0, g) + rsw.potential_energy(sol0, g, p)
    E1 = rsw.kinetic_energy(ref, g) + rsw.potential_energy(ref, g, p)
    assert abs(E1 - E0) / E0 < 5e-3   # quadratic (linearised) energy is only approximately conserved


def test_qg_operators_closed_forms():
    from oracle import qg
    g = TwoDGrid(32)
    L = qg.twolayer_L(g, F=2 * 9.0 / 0.2, U=0.5, mu=1e-2, nu=1e-10, nnu=4)
    E = qg.expm2x2_closed_form(L, 3e-3)
    np.testing.assert_allclose(E, scipy.linalg.expm(L * 3e-3), rtol=0, atol=1e-13)
    # pv <-> streamfunction inversion round trip (swqg/TwoLayerQG.jl:92-111)
    rng = np.random.default_rng(3)
    psih = rng.standard_normal((g.nkr, g.nl, 2)) + 1j * rng.standard_normal((g.nkr, g.nl, 2))
    psih[0, 0] = 0
    F = 90.0
    qh = np.stack([-g.Krsq * psih[:, :, 0] + F * (psih[:, :, 1] - psih[:, :, 0]),
                   -g.Krsq * psih[:, :, 1] + F * (psih[:, :, 0] - psih[:, :, 1])], axis=-1)
    np.testing.assert_allclose(qg.twolayer_streamfunction(qh, g, F), psih, rtol=0, atol=1e-12)
    # the Jacobian conserves the mean of q: N[0, 0] == 0
    sol = g.dealias(g.rfft2(rng.standard_normal((32, 32))))
    assert abs(qg.swqg_calcN(sol.copy(), g, 9.0)[0, 0]) < 1e-9


def test_bspline2_prefilter_and_node_reproduction():
    """Quadratic B-spline mode of the CPU tracer (raytracing/Raytracing.jl:161-170): the Fourier prefilter equals the
    periodic tridiagonal solve, the spline reproduces node values, and it converges faster than bilinear."""
    from oracle import raytrace as oray
    from oracle.grid import TwoDGrid
    g = TwoDGrid(32, 2 * np.pi)
    X, Y = np.meshgrid(g.x, g.y, indexing="ij")
    f = np.stack([np.sin(X) * np.cos(2 * Y), np.cos(3 * X + Y)], axis=-1)
    C = oray.bspline2_prefilter(f, g)
    assert np.abs(C - oray.bspline2_prefilter_direct(f)).max() < 1e-13
    xs, ys = X.ravel(), Y.ravel()
    assert np.abs(oray.sample_bspline2(C, xs, ys, g) - f.reshape(-1, 2)).max() < 1e-13
    rng = np.random.default_rng(0)
    px, py = rng.uniform(-10, 10, 500), rng.uniform(-10, 10, 500)        # periodic wrap outside the box
    exact = np.stack([np.sin(px) * np.cos(2 * py), np.cos(3 * px + py)], axis=-1)
    C3 = oray.bspline3_prefilter(f, g)
    assert np.abs(oray.sample_bspline3(C3, xs, ys, g) - f.reshape(-1, 2)).max() < 1e-13       # the cubic reproduces the nodes too
    assert np.abs(oray.sample_bspline3(C3, px, py, g) - exact).max() < np.abs(oray.sample_bspline2(C, px, py, g) - exact).max()
    e_spline = np.abs(oray.sample_bspline2(C, px, py, g) - exact).max()
    e_lin = np.abs(oray.sample_bilinear(f, px, py, g) - exact).max()
    assert e_spline < 0.1 * e_lin


def test_implicit_midpoint_is_second_order_and_symmetric():
    """Implicit midpoint (raytracing/Raytracing.jl:106-109): integrating forward then backward returns to the start
    (the scheme is symmetric), and halving h cuts the error about four-fold."""
    from oracle import raytrace as oray, rsw as orsw
    from helpers import config2_setup
    g, p, sol0, c = config2_setup(64)
    F = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 6)
    T = 40 * c["dt"]
    ref = oray.raytrace_midpoint(xk.copy(), sign, 0.0, T, F, F, g, c["f"], c["Cg"], nsub=64)
    e = [np.abs(oray.raytrace_midpoint(xk.copy(), sign, 0.0, T, F, F, g, c["f"], c["Cg"], nsub=n) - ref).max() for n in (4, 8)]
    assert 2.5 < e[0] / e[1] < 6.0
    fwd = oray.raytrace_midpoint(xk.copy(), sign, 0.0, T, F, F, g, c["f"], c["Cg"], nsub=8, iters=40)
    back = oray.raytrace_midpoint(fwd.copy(), sign, T, 0.0, F, F, g, c["f"], c["Cg"], nsub=8, iters=40)
    assert np.abs(back - xk).max() < 1e-9 * max(1.0, np.abs(xk).max())


def test_K10_K11_wave_balanced_projections():
    """K10 (Notebooks/RSWInitialTestng.ipynb:95-96): the balanced part is non-divergent and the wave part carries no linear PV;
    K11 (thomasyamada/Notebooks/TestDecomposition.ipynb:45,138): G + W reproduces the baroclinic state (4.7e-13 there) and a
    purely balanced state has no wave part (1.4e-13 there)."""
    from oracle import decompose as od, rsw as orsw
    from oracle.grid import TwoDGrid
    from helpers import random_state
    g, sol = random_state(64, seed=1, amp=0.2)
    p = orsw.Params(1e-9, 4, 3.0, 1.0)
    bal, wav = od.wave_balanced_decomposition(sol, g, p)
    scale = np.abs(sol).max() * g.kr.max()
    assert np.abs(1j * g.kr * bal[:, :, 0] + 1j * g.l * bal[:, :, 1]).max() < 3.2e-13 * scale
    assert np.abs(1j * g.kr * wav[:, :, 1] - 1j * g.l * wav[:, :, 0] - p.f * wav[:, :, 2]).max() < 3.4e-13 * scale
    gt = TwoDGrid(64, 6 * np.pi)
    rng = np.random.default_rng(0)
    s4 = rng.standard_normal((gt.nkr, gt.nl, 4)) + 1j * rng.standard_normal((gt.nkr, gt.nl, 4))
    G, W = od.ty_decompose(s4, gt)
    assert np.abs(G + W - s4[:, :, 1:4]).max() < 4.7e-13
    sb = s4.copy()
    sb[:, :, 1:4] = G
    assert np.abs(od.ty_decompose(sb, gt)[1]).max() < 1.4e-13
    # the TY bases are orthonormal at every wavenumber
    P0, Pp, Pm = od.ty_bases(gt)
    for A in (P0, Pp, Pm):
        for B in (P0, Pp, Pm):
            ip = (A * np.conj(B)).sum(axis=-1)
            assert np.abs(ip - (1.0 if A is B else 0.0)).max() < 1e-13


def test_komega_window_and_detrend_restatement():
    """thomasyamada/TY_k_omega.jl:11-17 (periodic Hann) and mrsw/FourierRSW.jl:17-41 (detrend removes a linear trend's slope
    exactly; the intercept it subtracts is -m sum(t)/N, as written)."""
    from oracle import komega as okw
    w = okw.hann(8)
    assert w[0] == 0 and abs(w[4] - 1) < 1e-16 and np.allclose(w[1:], w[1:][::-1])
    t = np.linspace(0.5, 9.5, 19)
    data = (3.0 - 0.7j) * t[:, None] + np.ones((19, 2)) * (2.0 + 1.0j)
    d = okw.detrend(t, data)
    assert np.abs(d - d[0]).max() < 1e-12                       # the slope is gone, a constant is left
    m, b = okw.linear_least_squares(t, okw.demean(data))
    assert np.allclose(m, 3.0 - 0.7j) and np.allclose(b, -(3.0 - 0.7j) * t.sum() / 19)
    spec = okw.clean_fft(t, data, okw.hann(19))
    assert spec.shape == (19, 2)


def test_oracle_reproduces_committed_golden_vectors():
    """tests/golden/rsw64_config2.npz (made by tests/golden/make_golden.py): the oracle must keep reproducing it."""
    import os
    from helpers import config2_setup, oracle_steps
    from oracle import raytrace as oray, rsw as orsw
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rsw64_config2.npz"))
    g, p, sol0, c = config2_setup(64)
    assert np.abs(sol0 - G["sol0"]).max() <= 1e-13 * np.abs(sol0).max()
    sol10 = oracle_steps(g, p, G["sol0"].copy(), c["dt"], 10)
    assert np.linalg.norm(sol10 - G["sol10"]) <= 1e-12 * np.linalg.norm(sol10)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(G["sol10"], g, p), g)
    Fn = oray.get_velocity_info(orsw.get_streamfunction(G["sol13"], g, p), g)
    xk1 = oray.raytrace(G["xk0"].copy(), G["sign"], 10 * c["dt"], 13 * c["dt"], Fo, Fn, g, c["f"], c["Cg"], nsub=3)
    assert np.abs(xk1 - G["xk1"]).max() <= 1e-12 * np.abs(xk1).max()


def test_recalled_steppers_have_their_formal_order():
    """FilteredAB3 / ETDRK4 / FilteredRK4 are FourierFlows' (third party, recalled from its documentation -- parity unpinned): at
    least check each restatement against an exact solution.  y' = L y + N(y), L = -2, N = y^2 (Bernoulli):
    y(t) = 1 / ((1/y0 - 1/2) e^{2t} + 1/2); halving dt must cut the error 2^order-fold."""
    from oracle import qg as oqg, ty as oty
    L, y0, T = np.full((1, 1), -2.0), 0.7, 1.0
    exact = 1.0 / ((1 / y0 - 0.5) * np.exp(2 * T) + 0.5)
    N = lambda s: s * s

    def err(make, n):
        ts = make(T / n)
        y = np.full((1, 1), y0)
        for _ in range(n):
            ts.stepforward(y)
        return abs(y[0, 0] - exact)

    one = np.ones((1, 1))
    cases = (("FilteredAB3", lambda dt: oqg.FilteredAB3(L, dt, N, one), 3, (400, 800)),      # Euler start-up: measured late
             ("ETDRK4", lambda dt: oty.ETDRK4(L, dt, N), 4, (20, 40)),
             ("FilteredRK4", lambda dt: oty.FilteredRK4(L, dt, N, one), 4, (20, 40)))
    for name, make, order, (n1, n2) in cases:
        e1, e2 = err(make, n1), err(make, n2)
        rate = np.log2(e1 / e2)
        if name == "FilteredAB3":      # three Euler steps of size dt leave an O(dt^2) start-up error: second order globally
            assert 1.8 < rate < 3.3, (name, rate)
        else:
            assert order - 0.4 < rate < order + 0.6, (name, rate)
        assert e2 < 1e-4


def test_etdrk4_contour_coefficients_match_the_closed_forms():
    """The 32-point contour means (Kassam-Trefethen) against the closed-form ETD coefficients where those are well conditioned."""
    from oracle import ty as oty
    dt = 0.1
    L = np.array([[-30.0, -5.0, -0.7]])
    z = dt * L
    zeta, alpha, beta, gamma = oty.etdrk4_coeffs(dt, L)
    ez = np.exp(z)
    assert np.allclose(zeta, dt * (np.exp(z / 2) - 1) / z, rtol=1e-12)
    assert np.allclose(alpha, dt * (-4 - z + ez * (4 - 3 * z + z * z)) / z ** 3, rtol=1e-9)
    assert np.allclose(beta, dt * (2 + z + ez * (-2 + z)) / z ** 3, rtol=1e-9)
    assert np.allclose(gamma, dt * (-4 - 3 * z - z * z + ez * (4 - z)) / z ** 3, rtol=1e-9)
    # and they stay finite where the closed forms cancel catastrophically
    small = oty.etdrk4_coeffs(dt, np.array([[-1e-9, 0.0]]))
    assert all(np.isfinite(c).all() for c in small) and abs(small[1][0, 1] - dt / 6) < 1e-12


def test_multilayerqg_restatement_agrees_with_the_in_repo_two_layer_operator():
    """GeophysicalFlows' MultiLayerQG is third party (recalled), but for U = (U, -U), beta = 0 it must describe the same physics as
    the reference's own swqg/TwoLayerQG.jl: calcN_MLQG(q) + D q  ==  calcN_TwoLayerQG(q) + L_2x2 q  (L_2x2 of TwoLayerQG.jl:184-198
    carries the mean-flow advection, the PV-gradient term and the bottom drag that MultiLayerQG keeps inside calcN!)."""
    from oracle import qg as oqg
    from oracle.grid import TwoDGrid
    g = TwoDGrid(64)
    rng = np.random.default_rng(21)
    q = np.stack([g.rfft2(rng.standard_normal((64, 64))) for _ in range(2)], axis=-1)
    q = g.dealias(q * np.exp(-0.02 * g.Krsq)[:, :, None])
    F, U, mu, nu, nnu = 37.5, 0.4, 0.3, 1e-6, 2
    L2 = oqg.twolayer_L(g, F, U, mu, nu, nnu)
    lhs = oqg.twolayer_calcN(q.copy(), g, F) + np.einsum("ijab,ijb->ija", L2, q)
    D = (-nu * g.Krsq ** nnu)[:, :, None]
    rhs = oqg.multilayer2_calcN(q.copy(), g, F, U, -U, 0.0, mu) + D * q
    # the aliased band of the physical-space products differs (the in-repo model leaves it to the next dealias!): compare retained modes
    lhs, rhs = g.dealias(lhs), g.dealias(rhs)
    assert np.linalg.norm(lhs - rhs) < 1e-12 * np.linalg.norm(lhs)


def test_makefilter_shape():
    """FourierFlows.makefilter (recalled): 1 up to innerK, exp(-decay (K - innerK)^order) beyond, equal to `tol` at K = outerK."""
    from oracle.grid import TwoDGrid, makefilter
    g = TwoDGrid(64)
    filt = makefilter(g, order=4, innerK=2 / 3, outerK=1.0, tol=1e-15)
    K = np.sqrt((g.kr * g.dx / np.pi) ** 2 + (g.l * g.dy / np.pi) ** 2)
    assert np.all(filt[K < 2 / 3] == 1.0) and np.all(np.diff(filt[:, 0]) <= 0)
    i = np.argmin(np.abs(K[:, 0] - 1.0))            # the Nyquist mode sits at K = 1
    assert abs(K[i, 0] - 1.0) < 1e-12 and abs(filt[i, 0] / 1e-15 - 1) < 1e-9


def test_bsplines_match_scipy_ndimage():
    """Independent check of the two B-spline restatements (Interpolations.jl is not available here): scipy.ndimage's periodic
    centred B-spline interpolation (`map_coordinates(order=2|3, mode='grid-wrap')`) uses the same basis and prefilter."""
    from scipy import ndimage
    from oracle import raytrace as oray
    from oracle.grid import TwoDGrid
    g = TwoDGrid(32, 2 * np.pi)
    rng = np.random.default_rng(4)
    f = rng.standard_normal((32, 32, 1))
    ix, iy = rng.uniform(-40, 70, 300), rng.uniform(-40, 70, 300)          # index coordinates, far outside one period too
    px, py = g.x[0] + ix * g.dx, g.y[0] + iy * g.dy
    for order, pre, samp in ((2, oray.bspline2_prefilter, oray.sample_bspline2), (3, oray.bspline3_prefilter, oray.sample_bspline3)):
        want = ndimage.map_coordinates(f[:, :, 0], np.stack([ix, iy]), order=order, mode="grid-wrap")
        got = samp(pre(f, g), px, py, g)[:, 0]
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max(), order


def test_implicit_midpoint_fixed_point_satisfies_the_implicit_equation():
    """The 12 fixed-point sweeps (oracle and CUDA alike) must solve y+ = y + h f(t + h/2, (y + y+)/2) -- what OrdinaryDiffEq's Newton
    iteration solves -- to round-off for CFL-limited steps."""
    from oracle import raytrace as oray, rsw as orsw
    from helpers import config2_setup
    g, p, sol0, c = config2_setup(64)
    F = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 8)
    h = c["dt"]                                    # the drivers' step: h |grad U| ~ 0.1, contraction ~0.05 per sweep
    y1 = oray.raytrace_midpoint(xk.copy(), sign, 0.0, h, F, F, g, c["f"], c["Cg"], nsub=1)
    z = 0.5 * (xk + y1)
    Sz = oray.sample_bilinear(F, z[:, 0], z[:, 1], g)
    res = y1 - xk - h * oray.rhs_sampler(z, sign, 0.5, Sz, Sz, c["f"], c["Cg"])
    assert np.abs(res).max() < 1e-14 * np.abs(xk).max()
    # four times the step (h |grad U| ~ 0.4) still converges, just more slowly: 1e-12 after 12 sweeps, round-off after 24
    y4 = oray.raytrace_midpoint(xk.copy(), sign, 0.0, 4 * h, F, F, g, c["f"], c["Cg"], nsub=1, iters=24)
    z = 0.5 * (xk + y4)
    Sz = oray.sample_bilinear(F, z[:, 0], z[:, 1], g)
    assert np.abs(y4 - xk - 4 * h * oray.rhs_sampler(z, sign, 0.5, Sz, Sz, c["f"], c["Cg"])).max() < 1e-13 * np.abs(xk).max()


def test_hermite_gradient_is_the_gradient_of_the_interpolant_and_hann_matches_scipy():
    """The Hermite-bicubic mode's (ux, uy, vx) are specified as the analytic gradient of the interpolated (u, v): check against
    centred differences of the interpolant itself.  And the periodic Hann window of the k-omega pipeline is scipy's `sym=False`."""
    from scipy.signal import windows
    from oracle import komega as okw, raytrace as oray, rsw as orsw
    from helpers import config2_setup
    assert np.allclose(okw.hann(37), windows.hann(37, sym=False), atol=1e-15)
    g, p, sol0, c = config2_setup(64)
    F7 = oray.get_velocity_info_cubic(orsw.get_streamfunction(sol0, g, p), g)
    rng = np.random.default_rng(6)
    x, y = rng.uniform(-3, 3, 200), rng.uniform(-3, 3, 200)
    e = 1e-6 * g.dx
    S = oray.sample_hermite(F7, x, y, g)
    dudx = (oray.sample_hermite(F7, x + e, y, g)[:, 0] - oray.sample_hermite(F7, x - e, y, g)[:, 0]) / (2 * e)
    dudy = (oray.sample_hermite(F7, x, y + e, g)[:, 0] - oray.sample_hermite(F7, x, y - e, g)[:, 0]) / (2 * e)
    dvdx = (oray.sample_hermite(F7, x + e, y, g)[:, 1] - oray.sample_hermite(F7, x - e, y, g)[:, 1]) / (2 * e)
    scale = np.abs(S[:, 2:]).max()
    assert np.abs(S[:, 2] - dudx).max() < 1e-6 * scale and np.abs(S[:, 3] - dudy).max() < 1e-6 * scale
    assert np.abs(S[:, 4] - dvdx).max() < 1e-6 * scale
    xs, ys = np.meshgrid(g.x, g.y, indexing="ij")                       # and it reproduces the node data
    assert np.abs(oray.sample_hermite(F7, xs.ravel(), ys.ravel(), g) - F7[:, :, :5].reshape(-1, 5)).max() < 1e-12 * np.abs(F7).max()


def test_c_restatement_of_the_tracer_is_bit_identical_to_the_numpy_oracle():
    """oracle/c/raytrace_oracle.c (compiled, OpenMP; the tracer of bench.py's CPU legs) against oracle/raytrace.py, bit for bit,
    both time-lerp conventions, packets far outside the periodic box included."""
    import os
    import subprocess
    from oracle import craytrace, raytrace as oray, rsw as orsw
    from helpers import config2_setup
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle", "c")], check=True, capture_output=True)
    assert craytrace.available()
    g, p, sol0, c = config2_setup(64)
    Fo = oray.get_velocity_info(orsw.get_streamfunction(sol0, g, p), g)
    Fn = np.roll(Fo, 3, axis=0) * 1.02
    xk, sign = oray.generate_initial_wavepackets(c["L"], c["k0"], 40)
    xk[:, 0:2] += np.random.default_rng(8).uniform(-50, 50, size=(xk.shape[0], 2))
    for lerp in (0, 1):
        a = oray.raytrace(xk.copy(), sign, 0.25, 0.25 + 3 * c["dt"], Fo, Fn, g, c["f"], c["Cg"], nsub=3, lerp=lerp)
        b = craytrace.raytrace(xk.copy(), sign, 0.25, 0.25 + 3 * c["dt"], Fo, Fn, g, c["f"], c["Cg"], nsub=3, lerp=lerp)
        np.testing.assert_array_equal(a, b)


def test_addforcing_broadcasts_the_2d_field_over_the_three_equations():
    """rsw/RotatingShallowWater.jl:234-240: vars.Fh is (nkr, nl) and `@. N += vars.Fh` adds it to N[:, :, 1], N[:, :, 2] AND N[:, :, 3]."""
    from helpers import config2_setup
    from oracle import rsw as orsw
    g, p, sol0, c = config2_setup(32)
    rng = np.random.default_rng(3)
    Fh = rng.standard_normal((g.nkr, g.nl)) + 1j * rng.standard_normal((g.nkr, g.nl))
    for variant in (orsw.RSW, orsw.MODIFIED, orsw.LINDBORG, orsw.QUADHEIGHT):
        N0 = orsw.calcN(sol0.copy(), g, p, variant)
        N1 = orsw.calcN(sol0.copy(), g, p, variant, Fh=Fh)
        for v in range(3):
            np.testing.assert_allclose(N1[:, :, v] - N0[:, :, v], Fh, rtol=0, atol=1e-12 * np.abs(N0).max())
