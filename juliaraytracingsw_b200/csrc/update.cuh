// Fused IFMAB3 update: one coalesced pass over the retained spectral state.
// Reference: utils/IFMAB3.jl:129-169 (IFMAB3update! + filter + history), closed-form exp(L dt)
// per SURVEY App. A.4.  Included by api.cu only (non-template kernels live in one TU).
#pragma once
#include "passes.cuh"

namespace swrt {

// ---------------------------------------------------------------- IFMAB3 update (RSW family)
// coef[l][kr] = { e^{D dt}, sin(w dt)/w, (1-cos(w dt))/w^2, filter }
// exp(L dt) x = e^{D dt} [ x + s L0 x + c L0 L0 x ],  L0 = L - D I  
struct RswLin {
    double f, c2;  // c2 = Cg^2 for RSW/Lindborg, 0 for Modified (no -i k c2 coupling)
    double w2c;    // w^2 = f^2 + w2c K^2
    __device__ __forceinline__ void L0(const double2 (&x)[3], double k, double l, double2 (&y)[3]) const {
        // [0, f, -i k c2; -f, 0, -i l c2; -i k, -i l, 0]
        y[0] = make_double2(f * x[1].x + k * c2 * x[2].y, f * x[1].y - k * c2 * x[2].x);
        y[1] = make_double2(-f * x[0].x + l * c2 * x[2].y, -f * x[0].y - l * c2 * x[2].x);
        y[2] = make_double2(k * x[0].y + l * x[1].y, -(k * x[0].x + l * x[1].x));
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
        const double kw = kr * L.dk, lw = wave_l(L, l);
        const double4 cf = a.coef[off];
        const double eD = cf.x, s = cf.y, c = cf.z;
        double2 x[3], n[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            x[v] = a.sol[v * L.vs + off];
            n[v] = a.N[v * L.vs + off];
        }
        if (a.euler) {
#pragma unroll
            for (int v = 0; v < 3; ++v) x[v] = make_double2(x[v].x + a.dt * n[v].x, x[v].y + a.dt * n[v].y);
        } else {
            double2 n1[3], n2[3], A[3], B[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                n1[v] = a.Nm1[v * L.vs + off];
                n2[v] = a.Nm2[v * L.vs + off];
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

}  // namespace swrt
