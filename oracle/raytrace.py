"""Wave-packet ray tracer restated in NumPy (oracle; test infrastructure only).

Ray equations and sampling follow raytracing/GPURaytracing.jl:18-65 (RHS `dxkdt`,
texture-coordinate map, dispersion relation with `pos_neg` sign) with the two quirks of
SURVEY App. B made explicit flags; the integrator is the north star's fixed-step classical
RK4 (the reference integrates with adaptive Vern7 / implicit midpoint, OrdinaryDiffEq,
un-vendored => the integrator is OUR spec and parity for it is against this file).

  get_velocity_info      raytracing/RaytracingDriver.jl:132-154
  generate_initial_wavepackets  raytracing/RaytracingDriver.jl:27-47   (K7)
  k-cutoff reset         raytracing/GPUTwoLayerRaytracing.jl:136-138
  bilinear sampler       GPURaytracing.jl:18-20 + CUDA texture semantics (exact at nodes, K6),
                         with `floor` instead of `unsafe_trunc` (App. B #6)
  hermite bicubic        utils/CUDAInterpolations.jl:39-53,71-108
Packets: array (N, 4) with columns x, y, k, l; omega_sign (N,).
"""
from __future__ import annotations

import numpy as np

LERP_PHYSICAL = 0   # W = (1-a) old + a new      (Raytracing.jl:163-168, the CPU tracer)
LERP_REFERENCE_GPU = 1  # W = a old + (1-a) new  (GPURaytracing.jl:33,53, App. B #2)


def get_velocity_info(psih, grid):
    """u,v,ux,uy,vx (vy = -ux) from a streamfunction; RaytracingDriver.jl:132-154."""
    k, l = grid.kr, grid.l
    u = grid.irfft2(-1j * l * psih)
    v = grid.irfft2(1j * k * psih)
    ux = grid.irfft2(k * l * psih)
    uy = grid.irfft2(l * l * psih)
    vx = grid.irfft2(-k * k * psih)
    return np.stack([u, v, ux, uy, vx], axis=-1)      # (nx, ny, 5)


def generate_initial_wavepackets(L, k0, sqrtN):
    """RaytracingDriver.jl:27-47: lattice (x = repeat outer, y = repeat inner), ring of
    wavevectors, odd (1-based) packets get omega_sign = -1."""
    N = sqrtN * sqrtN
    offset = L / sqrtN / 2
    I = np.arange(1, sqrtN + 1, dtype=np.float64)
    J = np.arange(1, N + 1, dtype=np.float64)
    diag = I * L / sqrtN - L / 2 - offset
    phase = 2 * np.pi * J / N
    xk = np.empty((N, 4))
    xk[:, 0] = np.tile(diag, sqrtN)        # repeat(diagonal, outer=n)
    xk[:, 1] = np.repeat(diag, sqrtN)      # repeat(diagonal, inner=n)
    xk[:, 2] = k0 * np.cos(phase)
    xk[:, 3] = k0 * np.sin(phase)
    sign = np.ones(N)
    sign[0::2] *= -1
    return xk, sign


def kcutoff_reset(xk, kcut, k0):
    """GPUTwoLayerRaytracing.jl:136-138.  Returns number of packets reset (bit-exact compare+select)."""
    mask = xk[:, 2] ** 2 + xk[:, 3] ** 2 >= kcut ** 2
    xk[mask, 2] = k0
    xk[mask, 3] = 0.0
    return int(mask.sum())


def cell_index(x, x0, dx, n):
    """Texture addressing restated: s = (x - x0)/dx, i = floor(s) mod n, a = s - floor(s)."""
    s = (x - x0) / dx
    fl = np.floor(s)
    i = np.mod(fl, n).astype(np.int64)
    return i, s - fl


def sample_bilinear(fields, x, y, grid):
    """fields: (nx, ny, C).  Returns (N, C).  Exact at nodes (K6)."""
    i, a = cell_index(x, grid.x[0], grid.dx, grid.nx)
    j, b = cell_index(y, grid.y[0], grid.dy, grid.ny)
    i1, j1 = (i + 1) % grid.nx, (j + 1) % grid.ny
    a, b = a[:, None], b[:, None]
    bottom = (1 - a) * fields[i, j] + a * fields[i1, j]
    top = (1 - a) * fields[i, j1] + a * fields[i1, j1]
    return (1 - b) * bottom + b * top


def _cubic(al, f0, f1, m0, m1):
    """utils/CUDAInterpolations.jl:39-44."""
    return f0 + m0 * al + (-3 * f0 + 3 * f1 - 2 * m0 - m1) * al ** 2 + (2 * f0 - 2 * f1 + m0 + m1) * al ** 3


def sample_bicubic_hermite(f, fx, fy, fxy, x, y, grid):
    """utils/CUDAInterpolations.jl:71-108 for one scalar field with its derivatives."""
    i, a = cell_index(x, grid.x[0], grid.dx, grid.nx)
    j, b = cell_index(y, grid.y[0], grid.dy, grid.ny)
    i1, j1 = (i + 1) % grid.nx, (j + 1) % grid.ny
    dx = grid.dx
    f0 = _cubic(a, f[i, j], f[i1, j], fx[i, j] * dx, fx[i1, j] * dx)
    f1 = _cubic(a, f[i, j1], f[i1, j1], fx[i, j1] * dx, fx[i1, j1] * dx)
    g0 = _cubic(a, fy[i, j] * dx, fy[i1, j] * dx, fxy[i, j] * dx * dx, fxy[i1, j] * dx * dx)
    g1 = _cubic(a, fy[i, j1] * dx, fy[i1, j1] * dx, fxy[i, j1] * dx * dx, fxy[i1, j1] * dx * dx)
    return _cubic(b, f0, f1, g0, g1)


def _dcubic(al, f0, f1, m0, m1):
    """d/d(alpha) of the Hermite cubic above."""
    return m0 + 2 * (-3 * f0 + 3 * f1 - 2 * m0 - m1) * al + 3 * (2 * f0 - 2 * f1 + m0 + m1) * al ** 2


def get_velocity_info_cubic(psih, grid):
    """Node data of the Hermite-bicubic mode: u, v, ux, uy, vx, uxy, vxy from a streamfunction (nx, ny, 7)."""
    k, l = grid.kr, grid.l
    F5 = get_velocity_info(psih, grid)
    uxy = grid.irfft2(1j * k * l * l * psih)          # -psi_xyy
    vxy = grid.irfft2(-1j * k * k * l * psih)         # psi_xxy
    return np.concatenate([F5, uxy[:, :, None], vxy[:, :, None]], axis=-1)


def sample_hermite(F7, x, y, grid):
    """Hermite-bicubic interpolation of u and v from (f, f_x, f_y, f_xy) node data, utils/CUDAInterpolations.jl:71-108,
    and the ANALYTIC gradient of the interpolant (our specification of the cubic mode: the gradient that enters
    dk/dt is consistent with the velocity that enters dx/dt).  Returns (N, 5) = u, v, ux, uy, vx."""
    i, a = cell_index(x, grid.x[0], grid.dx, grid.nx)
    j, b = cell_index(y, grid.y[0], grid.dy, grid.ny)
    i1, j1 = (i + 1) % grid.nx, (j + 1) % grid.ny
    dx, dy = grid.dx, grid.dy
    out = np.empty((x.shape[0], 5))
    defs = ((F7[:, :, 0], F7[:, :, 2], F7[:, :, 3], F7[:, :, 5]),       # u, ux, uy, uxy
            (F7[:, :, 1], F7[:, :, 4], -F7[:, :, 2], F7[:, :, 6]))      # v, vx, vy = -ux, vxy
    for n, (f, fx, fy, fxy) in enumerate(defs):
        c = lambda A, s: (A[i, j] * s, A[i1, j] * s, A[i, j1] * s, A[i1, j1] * s)
        f00, f10, f01, f11 = c(f, 1.0)
        x00, x10, x01, x11 = c(fx, dx)
        y00, y10, y01, y11 = c(fy, dy)
        m00, m10, m01, m11 = c(fxy, dx * dy)
        f0, f1 = _cubic(a, f00, f10, x00, x10), _cubic(a, f01, f11, x01, x11)
        g0, g1 = _cubic(a, y00, y10, m00, m10), _cubic(a, y01, y11, m01, m11)
        val = _cubic(b, f0, f1, g0, g1)
        d0, d1 = _dcubic(a, f00, f10, x00, x10), _dcubic(a, f01, f11, x01, x11)
        e0, e1 = _dcubic(a, y00, y10, m00, m10), _dcubic(a, y01, y11, m01, m11)
        ddx = _cubic(b, d0, d1, e0, e1) / dx
        ddy = _dcubic(b, f0, f1, g0, g1) / dy
        if n == 0:
            out[:, 0], out[:, 2], out[:, 3] = val, ddx, ddy
        else:
            out[:, 1], out[:, 4] = val, ddx
    return out


def rhs(xk, sign, t, t0, t1, F_old, F_new, grid, f, Cg, lerp=LERP_PHYSICAL):
    """dxkdt of GPURaytracing.jl:32-65.  F_* are (nx, ny, 5) = u, v, ux, uy, vx (bilinear mode) or (nx, ny, 7) with
    uxy, vxy appended (Hermite-bicubic mode)."""
    alpha = (t - t0) / (t1 - t0)
    x, y, k, l = xk[:, 0], xk[:, 1], xk[:, 2], xk[:, 3]
    w = sign * np.sqrt(f * f + Cg * Cg * (k * k + l * l))
    cgx, cgy = Cg * Cg * k / w, Cg * Cg * l / w
    sampler = sample_bilinear if F_old.shape[-1] == 5 else sample_hermite
    So = sampler(F_old, x, y, grid)
    Sn = sampler(F_new, x, y, grid)
    if lerp == LERP_PHYSICAL:
        W = (1 - alpha) * So + alpha * Sn
    else:
        W = alpha * So + (1 - alpha) * Sn
    out = np.empty_like(xk)
    out[:, 0] = W[:, 0] + cgx
    out[:, 1] = W[:, 1] + cgy
    out[:, 2] = -(W[:, 2] * k + W[:, 4] * l)
    out[:, 3] = -(W[:, 3] * k - W[:, 2] * l)       # vy = -ux
    return out


def raytrace(xk, sign, t0, t1, F_old, F_new, grid, f, Cg, nsub=1, lerp=LERP_PHYSICAL):
    """Advance packets in place from t0 to t1 with `nsub` classical RK4 steps."""
    h = (t1 - t0) / nsub
    for s in range(nsub):
        t = t0 + s * h
        a = (xk, sign, t0, t1, F_old, F_new, grid, f, Cg, lerp)
        k1 = rhs(xk, sign, t, *a[2:])
        k2 = rhs(xk + 0.5 * h * k1, sign, t + 0.5 * h, *a[2:])
        k3 = rhs(xk + 0.5 * h * k2, sign, t + 0.5 * h, *a[2:])
        k4 = rhs(xk + h * k3, sign, t + h, *a[2:])
        xk += (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    return xk


def interpolate_velocity(F, pos, grid):
    """interpolate_velocity!/interpolate_gradients! (GPURaytracing.jl:67-109): u,v and
    ux,uy,vx,vy at packet positions.  Returns (N,2), (N,4)."""
    S = (sample_bilinear if F.shape[-1] == 5 else sample_hermite)(F, pos[:, 0], pos[:, 1], grid)
    return S[:, 0:2].copy(), np.stack([S[:, 2], S[:, 3], S[:, 4], -S[:, 2]], axis=1)


# ------------------------------------------------------------------------------------ CPU-tracer semantics (raytracing/Raytracing.jl)
# Quadratic B-spline interpolation (Interpolations.jl BSpline(Quadratic(Periodic(OnGrid()))), :161-170) and the implicit-midpoint
# integrator (:106-109).  Interpolations.jl / OrdinaryDiffEq are third party and un-vendored: restated from their documented
# formulas -- PARITY UNPINNED.
def bspline2_prefilter(fields, grid):
    """Spline coefficients c with c_{i-1}/8 + 3 c_i/4 + c_{i+1}/8 = f_i in x and y (periodic).  The prefilter is a
    convolution, i.e. a division by (3/4 + cos(k dx)/4)(3/4 + cos(l dy)/4) in Fourier space."""
    px = 0.75 + 0.25 * np.cos(grid.kr * grid.dx)
    py = 0.75 + 0.25 * np.cos(grid.l * grid.dy)
    out = np.empty_like(fields)
    for c in range(fields.shape[-1]):
        out[:, :, c] = grid.irfft2(grid.rfft2(fields[:, :, c]) / (px * py))
    return out


def bspline2_prefilter_direct(f):
    """Same coefficients by solving the periodic tridiagonal systems (checks the Fourier shortcut)."""
    def solve(a, axis):
        n = a.shape[axis]
        M = np.zeros((n, n))
        idx = np.arange(n)
        M[idx, idx] = 0.75
        M[idx, (idx - 1) % n] = 0.125
        M[idx, (idx + 1) % n] = 0.125
        return np.moveaxis(np.linalg.solve(M, np.moveaxis(a, axis, 0).reshape(n, -1)).reshape(np.moveaxis(a, axis, 0).shape), 0, axis)
    return solve(solve(f, 0), 1)


def sample_bspline2(C, x, y, grid):
    """Evaluate the quadratic B-spline with coefficients C (nx, ny, F): weights (1/2)(d-1/2)^2, 3/4 - d^2, (1/2)(d+1/2)^2,
    d = s - round(s)."""
    def axis(pos, p0, dp, n):
        s = (pos - p0) / dp
        r = np.floor(s + 0.5)
        d = s - r
        i = np.mod(r, n).astype(np.int64)
        return i, np.stack([0.5 * (d - 0.5) ** 2, 0.75 - d * d, 0.5 * (d + 0.5) ** 2], axis=0)
    i, wx = axis(x, grid.x[0], grid.dx, grid.nx)
    j, wy = axis(y, grid.y[0], grid.dy, grid.ny)
    out = np.zeros((x.shape[0], C.shape[-1]))
    for a in range(3):
        for b in range(3):
            out += (wx[a] * wy[b])[:, None] * C[(i + a - 1) % grid.nx, (j + b - 1) % grid.ny]
    return out


def bspline3_prefilter(fields, grid):
    """Cubic B-spline coefficients, c_{i-1}/6 + 2 c_i/3 + c_{i+1}/6 = f_i (periodic) in x and y: the steady-flow interpolant
    `BSpline(Cubic(Periodic(OnCell())))` of raytracing/Raytracing.jl:152-159 -- PARITY UNPINNED."""
    px = 2 / 3 + np.cos(grid.kr * grid.dx) / 3
    py = 2 / 3 + np.cos(grid.l * grid.dy) / 3
    out = np.empty_like(fields)
    for c in range(fields.shape[-1]):
        out[:, :, c] = grid.irfft2(grid.rfft2(fields[:, :, c]) / (px * py))
    return out


def sample_bspline3(C, x, y, grid):
    """Cubic B-spline with coefficients C (nx, ny, F): weights (1-d)^3/6, (3d^3-6d^2+4)/6, (-3d^3+3d^2+3d+1)/6, d^3/6 on nodes
    floor(s)-1 .. floor(s)+2, d = s - floor(s)."""
    def axis(pos, p0, dp, n):
        s = (pos - p0) / dp
        fl = np.floor(s)
        d = s - fl
        w = np.stack([(1 - d) ** 3 / 6, (3 * d ** 3 - 6 * d ** 2 + 4) / 6, (-3 * d ** 3 + 3 * d ** 2 + 3 * d + 1) / 6, d ** 3 / 6], axis=0)
        return np.mod(fl, n).astype(np.int64), w
    i, wx = axis(x, grid.x[0], grid.dx, grid.nx)
    j, wy = axis(y, grid.y[0], grid.dy, grid.ny)
    out = np.zeros((x.shape[0], C.shape[-1]))
    for a in range(4):
        for b in range(4):
            out += (wx[a] * wy[b])[:, None] * C[(i + a - 1) % grid.nx, (j + b - 1) % grid.ny]
    return out


def rhs_sampler(xk, sign, alpha, S_old, S_new, f, Cg, lerp=LERP_PHYSICAL):
    """Ray RHS from already sampled (N, 5) fields of the two time levels."""
    k, l = xk[:, 2], xk[:, 3]
    w = sign * np.sqrt(f * f + Cg * Cg * (k * k + l * l))
    W = (1 - alpha) * S_old + alpha * S_new if lerp == LERP_PHYSICAL else alpha * S_old + (1 - alpha) * S_new
    out = np.empty_like(xk)
    out[:, 0] = W[:, 0] + Cg * Cg * k / w
    out[:, 1] = W[:, 1] + Cg * Cg * l / w
    out[:, 2] = -(W[:, 2] * k + W[:, 4] * l)
    out[:, 3] = -(W[:, 3] * k - W[:, 2] * l)
    return out


def raytrace_midpoint(xk, sign, t0, t1, F_old, F_new, grid, f, Cg, nsub=1, sampler=sample_bilinear, iters=12):
    """Implicit midpoint y+ = y + h f(t + h/2, (y + y+)/2) by fixed-point iteration on the midpoint z (a fixed number of
    sweeps: the map contracts by ~h |grad U| per sweep, so 12 sweeps reach round-off for CFL-limited steps)."""
    h = (t1 - t0) / nsub
    for s in range(nsub):
        alpha = (t0 + (s + 0.5) * h - t0) / (t1 - t0)
        z = xk.copy()
        for _ in range(iters):
            fz = rhs_sampler(z, sign, alpha, sampler(F_old, z[:, 0], z[:, 1], grid), sampler(F_new, z[:, 0], z[:, 1], grid), f, Cg)
            z = xk + 0.5 * h * fz
        fz = rhs_sampler(z, sign, alpha, sampler(F_old, z[:, 0], z[:, 1], grid), sampler(F_new, z[:, 0], z[:, 1], grid), f, Cg)
        xk += h * fz
    return xk


def refine_streamfunction(psih, grid, r):
    """Spectral zero padding of psih (nkr, nl) onto a grid r times finer (rsw/RSWDriver.jl:16-36 does the same to restart at a
    higher resolution; Notebooks/FFTInterpTest.ipynb uses it to interpolate): returns (psih_fine, grid_fine).  The trigonometric
    interpolant is unchanged; only its sampling grid is finer."""
    from .grid import TwoDGrid
    gf = TwoDGrid(r * grid.nx, grid.Lx, r * grid.ny, grid.Ly, aliased_fraction=0)
    half = grid.nl // 2
    new = np.zeros((gf.nkr, gf.nl), dtype=np.complex128)
    new[:grid.nkr - 1, :half] = psih[:grid.nkr - 1, :half]
    new[:grid.nkr - 1, gf.nl - half + 1:] = psih[:grid.nkr - 1, half + 1:]          # Nyquist row and column carry nothing (dealiased)
    return new * (r * r), gf


def generate_initial_wavepackets_twolayer(L, k0, sqrtN):
    """raytracing/TwoLayerRaytracing.jl:10-22 (the CPU driver's lattice): packet (i-1) s + j sits at
    (i L/s - L/2 - L/2s, j L/s - L/2 - L/2s) with wavevector angle 2 pi ((i-1) s + j)/N; all frequency signs +1."""
    s = int(sqrtN)
    N = s * s
    xk = np.empty((N, 4))
    offset = L / s / 2
    for i in range(1, s + 1):
        for j in range(1, s + 1):
            r = (i - 1) * s + j
            xk[r - 1] = (i * L / s - L / 2 - offset, j * L / s - L / 2 - offset,
                         k0 * np.cos(2 * np.pi * r / N), k0 * np.sin(2 * np.pi * r / N))
    return xk, np.ones(N)


def sample_trigonometric(psih, x, y, grid):
    """Exact spectral evaluation of u, v, ux, uy, vx at arbitrary points: what raytracing/NUFFTRaytracing.jl:68-84 approximates with
    nufft2d2 (tol 1e-5) of the spectral fields -i l psih, i k psih, k l psih, l^2 psih, -k^2 psih.  Direct sum over the rfft
    half plane with c2r semantics (weight 1 on kr = 0 and the Nyquist column, 2 elsewhere; real part).  O(N nkr nl): small cases."""
    k, l = grid.kr, grid.l
    specs = (-1j * l * psih, 1j * k * psih, k * l * psih, l * l * psih, -k * k * psih)
    w = np.where((np.arange(grid.nkr) == 0) | (np.arange(grid.nkr) == grid.nkr - 1), 1.0, 2.0)[:, None]
    out = np.empty((x.shape[0], 5))
    ex = np.exp(1j * k[:, 0][None, :] * (x - grid.x[0])[:, None])            # (N, nkr)
    ey = np.exp(1j * l[0, :][None, :] * (y - grid.y[0])[:, None])            # (N, nl)
    for c, fh in enumerate(specs):
        out[:, c] = np.einsum("nk,kl,nl->n", ex, w * fh, ey).real / (grid.nx * grid.ny)
    return out
