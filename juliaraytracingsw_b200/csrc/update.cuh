// Fused IFMAB3 update: one coalesced pass over the retained spectral state.
// Reference: utils/IFMAB3.jl:129-169 (IFMAB3update! + filter + history), closed-form exp(L dt)
// per SURVEY App. A.4.  Included by api.cu only (non-template kernels live in one TU).
#pragma once
#include "models.cuh"

namespace swrt {

// ---------------------------------------------------------------- IFMAB3 update (RSW family)
// coef[l][kr] = { e^{D dt}, sin(w dt)/w, (1-cos(w dt))/w^2, filter }
// exp(L dt) x = e^{D dt} [ x + s L0 x + c L0 L0 x ],  L0 = L - D I  
struct RswLin {
    double f, c2;  // c2 = Cg^2 for RSW/Lindborg, 0 for Modified / QuadHeight (no -i k c2 coupling)
    double w2c;    // w^2 = f^2 + w2c K^2
    double d3;     // 1: third row (-i k, -i l, 0); 0: QuadHeight (third row of L0 vanishes)
    __device__ __forceinline__ void L0(const double2 (&x)[3], double k, double l, double2 (&y)[3]) const {
        // [0, f, -i k c2; -f, 0, -i l c2; -i k, -i l, 0]
        y[0] = make_double2(f * x[1].x + k * c2 * x[2].y, f * x[1].y - k * c2 * x[2].x);
        y[1] = make_double2(-f * x[0].x + l * c2 * x[2].y, -f * x[0].y - l * c2 * x[2].x);
        y[2] = make_double2(d3 * (k * x[0].y + l * x[1].y), -d3 * (k * x[0].x + l * x[1].x));
    }
    __device__ __forceinline__ void expmul(const double2 (&x)[3], double k, double l, double eD, double s, double c,
                                           double2 (&y)[3]) const {
        double2 a[3], b[3];
        L0(x, k, l, a);
        L0(a, k, l, b);
#pragma unroll
        for (int v = 0; v < 3; ++v)
            y[v] = make_double2(eD * (x[v].x + s * a[v].x + c * b[v].x), eD * (x[v].y + s * a[v].y + c * b[v].y));
    }
};

struct UpdateArgs {
    double2* sol;
    const double2* N;
    const double2* Nm1;
    const double2* Nm2;
    const double4* coef;
    double dt;
    int euler;  // clock.step < 3
};

__global__ void __launch_bounds__(256) ifmab3_update_rsw_kernel(UpdateArgs a, RswLin lin, SpecLayout L) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        const double kw = (L.kr_off + kr) * L.dk, lw = wave_l(L, l);
        // streaming loads for what this step never touches again (history, coefficients): the state written below should be the
        // thing that stays in L2 for the next step's first pass
        const double2 cf01 = __ldcs(reinterpret_cast<const double2*>(a.coef + off)), cf23 = __ldcs(reinterpret_cast<const double2*>(a.coef + off) + 1);
        const double4 cf = make_double4(cf01.x, cf01.y, cf23.x, cf23.y);
        const double eD = cf.x, s = cf.y, c = cf.z;
        double2 x[3], n[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            x[v] = __ldcs(a.sol + v * L.vs + off);
            n[v] = __ldcs(a.N + v * L.vs + off);
        }
        if (a.euler) {
#pragma unroll
            for (int v = 0; v < 3; ++v) x[v] = make_double2(x[v].x + a.dt * n[v].x, x[v].y + a.dt * n[v].y);
        } else {
            double2 n1[3], n2[3], A[3], B[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                n1[v] = __ldcs(a.Nm1 + v * L.vs + off);
                n2[v] = __ldcs(a.Nm2 + v * L.vs + off);
            }
            lin.expmul(n1, kw, lw, eD, s, c, A);
            const double w2 = lin.f * lin.f + lin.w2c * (kw * kw + lw * lw);
            const double cw = 1.0 - c * w2;
            lin.expmul(n2, kw, lw, eD * eD, 2.0 * s * cw, 2.0 * s * s, B);
            const double h1 = 23.0 / 12.0, h2 = 16.0 / 12.0, h3 = 5.0 / 12.0;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                x[v].x += a.dt * (h1 * n[v].x - h2 * A[v].x + h3 * B[v].x);
                x[v].y += a.dt * (h1 * n[v].y - h2 * A[v].y + h3 * B[v].y);
            }
        }
        double2 y[3];
        lin.expmul(x, kw, lw, eD, s, c, y);
#pragma unroll
        for (int v = 0; v < 3; ++v) a.sol[v * L.vs + off] = make_double2(cf.w * y[v].x, cf.w * y[v].y);
    }
}


// ---------------------------------------------------------------- IFMAB3 with tabulated exp(L dt), exp(2 L dt)
// General NV x NV complex blocks (two-layer QG: swqg/TwoLayerQG.jl:184-198 has no skew structure to exploit).
// Tables are SoA: E[(a NV + b)][l][kr_pad]; y_a = sum_b E[a,b] x_b (utils/IFMAB3.jl:90-100, orientation K2).
template <int NV>
__global__ void __launch_bounds__(256) ifmab3_update_table_kernel(UpdateArgs a, const double2* __restrict__ E,
                                                                  const double2* __restrict__ E2, SpecLayout L) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        double2 e[NV][NV], x[NV], n[NV];
#pragma unroll
        for (int p = 0; p < NV; ++p)
#pragma unroll
            for (int q = 0; q < NV; ++q) e[p][q] = __ldcs(E + (p * NV + q) * L.vs + off);   // tables and history: streaming
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            x[v] = __ldcs(a.sol + v * L.vs + off);
            n[v] = __ldcs(a.N + v * L.vs + off);
        }
        if (a.euler) {
#pragma unroll
            for (int v = 0; v < NV; ++v) x[v] = make_double2(x[v].x + a.dt * n[v].x, x[v].y + a.dt * n[v].y);
        } else {
            const double h1 = 23.0 / 12.0, h2 = 16.0 / 12.0, h3 = 5.0 / 12.0;
            double2 n1[NV], n2[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                n1[v] = __ldcs(a.Nm1 + v * L.vs + off);
                n2[v] = __ldcs(a.Nm2 + v * L.vs + off);
            }
#pragma unroll
            for (int p = 0; p < NV; ++p) {
                double2 A = make_double2(0.0, 0.0), B = make_double2(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const double2 e2 = __ldcs(E2 + (p * NV + q) * L.vs + off);
                    A = cadd(A, cmul(e[p][q], n1[q]));
                    B = cadd(B, cmul(e2, n2[q]));
                }
                x[p].x += a.dt * (h1 * n[p].x - h2 * A.x + h3 * B.x);
                x[p].y += a.dt * (h1 * n[p].y - h2 * A.y + h3 * B.y);
            }
        }
        const double filt = a.coef[off].w;
#pragma unroll
        for (int p = 0; p < NV; ++p) {
            double2 y = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < NV; ++q) y = cadd(y, cmul(e[p][q], x[q]));
            a.sol[p * L.vs + off] = make_double2(filt * y.x, filt * y.y);
        }
    }
}

// ---------------------------------------------------------------- diagonal (real) L: IFMAB3(diagonal=true) and FilteredAB3
// coef = { e^{D dt}, D, -, filter }.  utils/IFMAB3.jl:72-74,102-108 ; FourierFlows FilteredAB3:
// RHS = N + L .* sol kept in the history ring (written back into `Nrw`), sol += dt (23/12 RHS - 16/12 RHS1 + 5/12 RHS2), filter.
template <int NV, bool FILTERED_AB3>
__global__ void __launch_bounds__(256) update_diag_kernel(UpdateArgs a, double2* __restrict__ Nrw, SpecLayout L) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    const double h1 = 23.0 / 12.0, h2 = 16.0 / 12.0, h3 = 5.0 / 12.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        const double4 cf = a.coef[off];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const long long o = v * L.vs + off;
            double2 x = a.sol[o], n = Nrw[o];
            if (FILTERED_AB3) {
                n = make_double2(n.x + cf.y * x.x, n.y + cf.y * x.y);
                Nrw[o] = n;
                if (a.euler) x = make_double2(x.x + a.dt * n.x, x.y + a.dt * n.y);
                else {
                    const double2 n1 = __ldcs(a.Nm1 + o), n2 = __ldcs(a.Nm2 + o);   // history: streaming
                    x.x += a.dt * (h1 * n.x - h2 * n1.x + h3 * n2.x);
                    x.y += a.dt * (h1 * n.y - h2 * n1.y + h3 * n2.y);
                }
                a.sol[o] = make_double2(cf.w * x.x, cf.w * x.y);
            } else {
                if (a.euler) x = make_double2(x.x + a.dt * n.x, x.y + a.dt * n.y);
                else {
                    const double2 n1 = __ldcs(a.Nm1 + o), n2 = __ldcs(a.Nm2 + o);   // history: streaming
                    const double e1 = cf.x, e2 = cf.x * cf.x;
                    x.x += a.dt * (h1 * n.x - h2 * e1 * n1.x + h3 * e2 * n2.x);
                    x.y += a.dt * (h1 * n.y - h2 * e1 * n1.y + h3 * e2 * n2.y);
                }
                a.sol[o] = make_double2(cf.w * cf.x * x.x, cf.w * cf.x * x.y);
            }
        }
    }
}

// ---------------------------------------------------------------- multi-stage steppers on a diagonal (real) L
// FourierFlows ETDRK4 and (Filtered)RK4 (third party, parity unpinned).
// coef = { e^{D dt}, D, e^{D dt/2}, filter }, coef2 = { zeta, alpha, beta, Gamma } (contour-integral ETD coefficients).
enum { ST_ETD_SUB12 = 0, ST_ETD_SUB3 = 1, ST_ETD_UPDATE = 2, ST_RK4_STAGE = 3, ST_RK4_FINAL = 4 };
struct StageArgs {
    double2* out;        // state written by this stage
    const double2* x;    // state the stage starts from
    double2* n1;         // N buffers (n1 may be updated in place: RK4 stores RHS = N + D x_stage there)
    const double2* n2;
    const double2* n3;
    const double2* n4;
    const double2* xs;   // RK4: the stage state N was evaluated at
    const double4* coef;
    const double4* coef2;
    double c;            // RK4: stage weight * dt
    int mode, nvar;
};
__global__ void __launch_bounds__(256) diag_stage_kernel(StageArgs a, SpecLayout L) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        const double4 cf = a.coef[off];
        const double4 c2 = a.coef2 ? a.coef2[off] : make_double4(0, 0, 0, 0);
        for (int v = 0; v < a.nvar; ++v) {
            const long long o = v * L.vs + off;
            if (a.mode == ST_ETD_SUB12) {                 // out = e^{L dt/2} x + zeta N
                const double2 x = a.x[o], n = a.n1[o];
                a.out[o] = make_double2(cf.z * x.x + c2.x * n.x, cf.z * x.y + c2.x * n.y);
            } else if (a.mode == ST_ETD_SUB3) {           // out = e^{L dt/2} x + zeta (2 N3 - N1)
                const double2 x = a.x[o], n1 = a.n1[o], n3 = a.n3[o];
                a.out[o] = make_double2(cf.z * x.x + c2.x * (2.0 * n3.x - n1.x), cf.z * x.y + c2.x * (2.0 * n3.y - n1.y));
            } else if (a.mode == ST_ETD_UPDATE) {         // sol = e^{L dt} sol + alpha N1 + 2 beta (N2 + N3) + Gamma N4
                const double2 x = a.x[o], n1 = a.n1[o], n2 = a.n2[o], n3 = a.n3[o], n4 = a.n4[o];
                // (cf.w = the filter of FilteredETDRK4 / use_filter, exactly 1 otherwise: `sol *= filter` after the update)
                a.out[o] = make_double2(cf.w * (cf.x * x.x + c2.y * n1.x + 2.0 * c2.z * (n2.x + n3.x) + c2.w * n4.x),
                                        cf.w * (cf.x * x.y + c2.y * n1.y + 2.0 * c2.z * (n2.y + n3.y) + c2.w * n4.y));
            } else if (a.mode == ST_RK4_STAGE) {          // RHS = N + D x_stage (kept); out = sol + c RHS
                const double2 xs = a.xs[o], x = a.x[o];
                double2 r = a.n1[o];
                r = make_double2(r.x + cf.y * xs.x, r.y + cf.y * xs.y);
                a.n1[o] = r;
                a.out[o] = make_double2(x.x + a.c * r.x, x.y + a.c * r.y);
            } else {                                      // RK4 final: sol += dt (R1/6 + R2/3 + R3/3 + R4/6); filter
                const double2 xs = a.xs[o], x = a.x[o], r1 = a.n1[o], r2 = a.n2[o], r3 = a.n3[o];
                double2 r4 = a.n4[o];
                r4 = make_double2(r4.x + cf.y * xs.x, r4.y + cf.y * xs.y);
                const double sx = x.x + a.c * (r1.x / 6.0 + r2.x / 3.0 + r3.x / 3.0 + r4.x / 6.0);
                const double sy = x.y + a.c * (r1.y / 6.0 + r2.y / 3.0 + r3.y / 3.0 + r4.y / 6.0);
                a.out[o] = make_double2(cf.w * sx, cf.w * sy);
            }
        }
    }
}

// ---------------------------------------------------------------- streamfunction for the packet snapshot
// psih[l][kr] = PsiLoader::psi_of(...)  (get_streamfunction!, rsw/RSWRaytracingDriver.jl:56-67 and the QG variants), materialised
// once so that the three y-transform jobs of the snapshot read one plain field (prefetchable, see passes.cuh)
// `lshift` > 0: psih is laid out for a finer node grid (spectral zero padding): rows of negative l move up by that many rows
__global__ void __launch_bounds__(256) psi_kernel(PsiLoader ld, SpecLayout L, double2* __restrict__ psih, int lshift = 0) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        psih[(long long)(l < L.lz0 ? l : l + lshift) * L.kr_pad + kr] = ld.psi_of((L.kr_off + kr) * L.dk, wave_l(L, l), off);
    }
}

// ---------------------------------------------------------------- RSW initial condition on the device
// set_initial_condition! of rsw/RSWRaytracingDriver.jl:15-54: a random-phase geostrophic band Kg and a wave band Kw.  `rnd`
// holds the host's random numbers in the reference layout, (phase, sgn) per mode of the (nkr, nl) array.
// part 0: geostrophic modes, part 1: wave modes, each UNSCALED into sol (the caller measures max|u|, :39-52);
// part 2: sol = sg * geostrophic + sw * wave.
__global__ void __launch_bounds__(256) rsw_ic_kernel(const double2* __restrict__ rnd, int nkr, SpecLayout L, int part, double f, double Cg2,
                                                     double Kg0, double Kg1, double Kw0, double Kw1, double sg, double sw, double2* __restrict__ sol) {
    const int rows = L.ny - (L.lz1 - L.lz0);
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)rows * L.kr_keep) return;
    const int kr = (int)(i % L.kr_keep), r = (int)(i / L.kr_keep), l = r < L.lz0 ? r : r + (L.lz1 - L.lz0);
    const double kw = (L.kr_off + kr) * L.dk, lw = wave_l(L, l), K2 = kw * kw + lw * lw;
    const double2 pr = rnd[(long long)l * nkr + L.kr_off + kr];
    double sn, cs;
    sincos(pr.x, &sn, &cs);                                       // shift = exp(i phase)
    double2 u = make_double2(0.0, 0.0), v = u, e = u;
    if (part != 1 && Kg0 * Kg0 <= K2 && K2 <= Kg1 * Kg1) {        // eta = 0.5 shift, u = -0.5 i Cg2/f l shift, v = 0.5 i Cg2/f k shift
        const double a = 0.5 * Cg2 / f * (part == 2 ? sg : 1.0), h = 0.5 * (part == 2 ? sg : 1.0);
        e = make_double2(h * cs, h * sn);
        u = make_double2(a * lw * sn, -a * lw * cs);
        v = make_double2(-a * kw * sn, a * kw * cs);
    }
    if (part != 0 && Kw0 * Kw0 <= K2 && K2 <= Kw1 * Kw1 && K2 > 0.0) {
        const double s = part == 2 ? sw : 1.0, wK = pr.y * sqrt(f * f + Cg2 * K2), inv = s / K2;
        // u = (0.5 k wK shift + 0.5 i f l shift)/K2 ; v = (0.5 l wK shift - 0.5 i f k shift)/K2 ; eta = 0.5 shift
        u.x += inv * (0.5 * kw * wK * cs - 0.5 * f * lw * sn);
        u.y += inv * (0.5 * kw * wK * sn + 0.5 * f * lw * cs);
        v.x += inv * (0.5 * lw * wK * cs + 0.5 * f * kw * sn);
        v.y += inv * (0.5 * lw * wK * sn - 0.5 * f * kw * cs);
        e.x += 0.5 * s * cs;
        e.y += 0.5 * s * sn;
    }
    const long long off = (long long)l * L.kr_pad + kr;
    sol[off] = u;
    sol[L.vs + off] = v;
    sol[2 * L.vs + off] = e;
}

// ---------------------------------------------------------------- wave / balanced projections
// DEC_RSW:     wave_balanced_decomposition, rsw/RSWUtils.jl:9-22 -- balanced part from the linear PV, wave part = rest.
// DEC_TY:      decompose_balanced_wave, thomasyamada/TYUtils.jl:10-51 -- projections of (u_c, v_c, p_c) on Phi0 and Phi+-.
// DEC_RSW_WTS: compute_balanced_wave_weights with the bases of rsw/RSWUtils.jl:24-57 -> A = (c0, c+, c-); the sign of the
//              eta component of Phi0 is the reference's (:33), which makes Phi0 non-orthogonal to Phi+- -- kept as written.
enum { DEC_RSW = 0, DEC_TY = 1, DEC_RSW_WTS = 2 };
struct cd { double x, y; };
__device__ __forceinline__ cd cmul(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cmulc(cd a, cd b) { return cd{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }   // a conj(b)
__global__ void __launch_bounds__(256) decompose_kernel(const double2* __restrict__ sol, SpecLayout L, int mode, double f, double Cg2,
                                                        double2* __restrict__ A, double2* __restrict__ B) {
    const int nlk = L.ny - (L.lz1 - L.lz0);
    const long long total = (long long)nlk * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int lr = (int)(i / L.kr_keep), kr = (int)(i - (long long)lr * L.kr_keep);
        const int l = lr < L.lz0 ? lr : lr + (L.lz1 - L.lz0);
        const long long off = (long long)l * L.kr_pad + kr;
        const double kw = (L.kr_off + kr) * L.dk, lw = wave_l(L, l), K2 = kw * kw + lw * lw;
        const int v0 = mode == DEC_TY ? 1 : 0;
        const double2 s0 = sol[(v0 + 0) * L.vs + off], s1 = sol[(v0 + 1) * L.vs + off], s2 = sol[(v0 + 2) * L.vs + off];
        if (mode == DEC_RSW) {
            // qh = i k vh - i l uh - f etah ; psih = -qh / (K^2 + f^2/Cg2)
            const double inv = -1.0 / (K2 + f * f / Cg2);
            const cd psi{inv * (-(kw * s1.y - lw * s0.y) - f * s2.x), inv * ((kw * s1.x - lw * s0.x) - f * s2.y)};
            const double2 g0 = make_double2(lw * psi.y, -lw * psi.x), g1 = make_double2(-kw * psi.y, kw * psi.x);
            const double2 g2 = make_double2(f / Cg2 * psi.x, f / Cg2 * psi.y);
            A[off] = g0; A[L.vs + off] = g1; A[2 * L.vs + off] = g2;
            B[off] = make_double2(s0.x - g0.x, s0.y - g0.y);
            B[L.vs + off] = make_double2(s1.x - g1.x, s1.y - g1.y);
            B[2 * L.vs + off] = make_double2(s2.x - g2.x, s2.y - g2.y);
            continue;
        }
        const bool ty = mode == DEC_TY;
        const double ff = ty ? 1.0 : f, Cg = ty ? 1.0 : sqrt(Cg2);
        const double w = sqrt(ff * ff + Cg * Cg * K2), sq = K2 > 0.0 ? sqrt(0.5 / K2) : 0.0;
        cd P0[3], Pp[3], Pm[3];
        if (K2 > 0.0) {
            if (ty) { P0[0] = cd{0, lw / w}; P0[1] = cd{0, -kw / w}; P0[2] = cd{-1.0 / w, 0}; }
            else { P0[0] = cd{0, -lw * Cg / w}; P0[1] = cd{0, kw * Cg / w}; P0[2] = cd{-ff / w, 0}; }
            Pp[0] = cd{w * kw * sq / w, ff * lw * sq / w};  Pp[1] = cd{w * lw * sq / w, -ff * kw * sq / w};
            Pm[0] = cd{-w * kw * sq / w, ff * lw * sq / w}; Pm[1] = cd{-w * lw * sq / w, -ff * kw * sq / w};
            Pp[2] = Pm[2] = cd{(ty ? (w * w - 1.0) : Cg * K2) * sq / w, 0};
        } else {
            const double r = 0.70710678118654752440;   // 1/sqrt(2)
            P0[0] = cd{0, 0}; P0[1] = cd{0, 0}; P0[2] = cd{1, 0};
            Pp[0] = cd{0, r}; Pp[1] = cd{r, 0}; Pp[2] = cd{0, 0};
            if (ty) { Pm[0] = cd{0, r}; Pm[1] = cd{-r, 0}; } else { Pm[0] = cd{0, -r}; Pm[1] = cd{r, 0}; }
            Pm[2] = cd{0, 0};
        }
        const cd x[3] = {cd{s0.x, s0.y}, cd{s1.x, s1.y}, cd{(ty ? 1.0 : Cg) * s2.x, (ty ? 1.0 : Cg) * s2.y}};
        cd c0{0, 0}, cp{0, 0}, cm{0, 0};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const cd a = cmulc(x[c], P0[c]), b = cmulc(x[c], Pp[c]), d = cmulc(x[c], Pm[c]);
            c0.x += a.x; c0.y += a.y; cp.x += b.x; cp.y += b.y; cm.x += d.x; cm.y += d.y;
        }
        if (mode == DEC_RSW_WTS) {
            A[off] = make_double2(c0.x, c0.y); A[L.vs + off] = make_double2(cp.x, cp.y); A[2 * L.vs + off] = make_double2(cm.x, cm.y);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const cd g = cmul(c0, P0[c]), wp = cmul(cp, Pp[c]), wm = cmul(cm, Pm[c]);
                A[c * L.vs + off] = make_double2(g.x, g.y);
                B[c * L.vs + off] = make_double2(wp.x + wm.x, wp.y + wm.y);
            }
        }
    }
}

// ---------------------------------------------------------------- k-omega accumulator
// One frame of the spectral series at a single kr index, all l, appended to device-resident time series buf[series][l][frame]:
//   SERIES_TY  (thomasyamada/TY_k_omega.jl:72-86): ut = -i l zeta_t, vt = i k zeta_t, (ug, vg), (uw, vw) of the Phi projections
//   SERIES_RSW (rsw/fourier-analysis/mrsw/FourierRSW.jl:118-137): u, v, eta; balanced u, v, eta; wave u, v, eta; c0, c+, c-
// A, B hold the projections written by decompose_kernel (A = balanced, B = wave; for SERIES_RSW C = weights).
enum { SERIES_TY = 0, SERIES_RSW = 1 };
__global__ void __launch_bounds__(128) series_append_kernel(const double2* __restrict__ sol, const double2* __restrict__ A,
                                                            const double2* __restrict__ B, const double2* __restrict__ C, SpecLayout L,
                                                            int kind, int kr, long long frame, long long maxf, double2* __restrict__ buf) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L.ny) return;
    const int nser = kind == SERIES_TY ? 6 : 12;
    const bool live = kr < L.kr_keep && l_retained(L, l);
    const long long off = (long long)l * L.kr_pad + kr;
    const double kw = (L.kr_off + kr) * L.dk, lw = wave_l(L, l);
    for (int s = 0; s < nser; ++s) {
        double2 v = make_double2(0.0, 0.0);
        if (live) {
            if (kind == SERIES_TY) {
                if (s < 2) {
                    const double2 z = sol[off];
                    v = s == 0 ? make_double2(lw * z.y, -lw * z.x) : make_double2(-kw * z.y, kw * z.x);
                } else v = (s < 4 ? A : B)[(s & 1) * L.vs + off];
            } else {
                const int j = s % 3;
                v = (s < 3 ? sol : s < 6 ? A : s < 9 ? B : C)[j * L.vs + off];
            }
        }
        buf[((long long)s * L.ny + l) * maxf + frame] = v;
    }
}
// e^{-2 pi i j / T}, j = 0..T-1
__global__ void twiddle_table_kernel(double2* __restrict__ tw, long long T) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= T) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)T, &s, &c);
    tw[j] = make_double2(c, s);
}
// Windowed (periodic Hann), optionally detrended, transform in time of one series row per CTA:
//   y = x - m t - b (detrend of FourierRSW.jl:33-36: slope from the demeaned data, intercept -m sum(t)/N), out[w] = sum_t hann[t] y[t] e^{-2 pi i w t / T}.
// `combo` builds the input from up to three series pairs: x = sum_j a[j] + i sum_j b[j]  (U_balanced / U_wave / U_total of TY_k_omega.jl:104-106).
struct SeriesSel { int n; int a[3]; int b[3]; };
__global__ void __launch_bounds__(256) series_dft_kernel(const double2* __restrict__ buf, SeriesSel sel, int ny, long long T, long long maxf,
                                                         const double* __restrict__ tv, const double2* __restrict__ tw, int detrend,
                                                         double2* __restrict__ work, double2* __restrict__ out) {
    const int l = blockIdx.x;
    __shared__ double red[256][5];
    double2* y = work + (long long)l * T;
    // gather (and combine) the row
    double sx = 0, sy = 0, st = 0, st2 = 0;
    for (long long t = threadIdx.x; t < T; t += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        for (int j = 0; j < sel.n; ++j) {
            const double2 a = buf[((long long)sel.a[j] * ny + l) * maxf + t];
            v.x += a.x; v.y += a.y;
            if (sel.b[j] >= 0) {
                const double2 b = buf[((long long)sel.b[j] * ny + l) * maxf + t];
                v.x -= b.y; v.y += b.x;
            }
        }
        y[t] = v;
        sx += v.x; sy += v.y; st += tv[t]; st2 += tv[t] * tv[t];
    }
    auto reduce = [&](double a0, double a1, double a2, double a3, double a4, double* r) {
        red[threadIdx.x][0] = a0; red[threadIdx.x][1] = a1; red[threadIdx.x][2] = a2; red[threadIdx.x][3] = a3; red[threadIdx.x][4] = a4;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) for (int q = 0; q < 5; ++q) red[threadIdx.x][q] += red[threadIdx.x + s][q];
            __syncthreads();
        }
        for (int q = 0; q < 5; ++q) r[q] = red[0][q];
        __syncthreads();
    };
    double mx = 0, my = 0, bx = 0, by = 0;
    if (detrend) {
        double r[5];
        reduce(sx, sy, st, st2, 0.0, r);
        const double N = (double)T, meanx = r[0] / N, meany = r[1] / N, tsum = r[2], t2sum = r[3];
        double tx = 0, ty = 0;
        for (long long t = threadIdx.x; t < T; t += blockDim.x) { tx += tv[t] * (y[t].x - meanx); ty += tv[t] * (y[t].y - meany); }
        reduce(tx, ty, 0.0, 0.0, 0.0, r);
        const double den = N * t2sum - tsum * tsum;
        mx = N * r[0] / den; my = N * r[1] / den;
        bx = -mx * tsum / N; by = -my * tsum / N;
    }
    for (long long t = threadIdx.x; t < T; t += blockDim.x) {
        const double w = 0.5 * (1.0 - cospi(2.0 * (double)t / (double)T));
        y[t] = make_double2(w * (y[t].x - mx * tv[t] - bx), w * (y[t].y - my * tv[t] - by));
    }
    __syncthreads();
    for (long long w = threadIdx.x; w < T; w += blockDim.x) {
        double ax = 0, ay = 0;
        long long idx = 0;   // (w t) mod T, advanced incrementally
        for (long long t = 0; t < T; ++t) {
            const double2 e = tw[idx], v = y[t];
            ax = fma(v.x, e.x, fma(-v.y, e.y, ax));
            ay = fma(v.x, e.y, fma(v.y, e.x, ay));
            idx += w;
            if (idx >= T) idx -= T;
        }
        out[(long long)l * T + w] = make_double2(ax, ay);
    }
}

// ---------------------------------------------------------------- spectral diagnostics (parseval-weighted sums)
// value(kr,l) per `which`, summed with weights 1 (kr = 0, Nyquist) / 2 (parsevalsum / parsevalsum2 of FourierFlows)
enum { DIAG_ABS2_VAR = 0, DIAG_QG_K2PSI2 = 1, DIAG_QG_PSI2 = 2, DIAG_QG_DPSI2 = 3, DIAG_INVK2_ABS2_VAR = 4 };
__global__ void __launch_bounds__(256) spectral_diag_kernel(const double2* __restrict__ sol, SpecLayout L, int which, int arg, int nlayers,
                                                            double P, double* __restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    const long long total = (long long)L.ny * L.kr_keep;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i / L.kr_keep), kr = (int)(i - (long long)l * L.kr_keep);
        if (!l_retained(L, l)) continue;
        const long long off = (long long)l * L.kr_pad + kr;
        const double kw = (L.kr_off + kr) * L.dk, lw = wave_l(L, l), K2 = kw * kw + lw * lw;
        double val;
        if (which == DIAG_ABS2_VAR || which == DIAG_INVK2_ABS2_VAR) {
            const double2 v = sol[arg * L.vs + off];
            val = v.x * v.x + v.y * v.y;
            if (which == DIAG_INVK2_ABS2_VAR) val = K2 > 0.0 ? val / K2 : 0.0;
        } else if (which == DIAG_QG_DPSI2) {
            const double2 a = qg_streamfunction(sol, L.vs, 2, 0, K2, P, off), b = qg_streamfunction(sol, L.vs, 2, 1, K2, P, off);
            val = (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y);
        } else {
            const double2 p = qg_streamfunction(sol, L.vs, nlayers, arg, K2, P, off);
            val = (which == DIAG_QG_K2PSI2 ? K2 : 1.0) * (p.x * p.x + p.y * p.y);
        }
        acc += ((L.kr_off + kr == 0 || L.kr_off + kr == L.nx / 2) ? 1.0 : 2.0) * val;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

}  // namespace swrt
