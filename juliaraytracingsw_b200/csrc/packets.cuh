// Wave-packet ray tracer kernels (sm_100a, fp64): one thread per packet, state in registers,
// `nsub` classical RK4 steps per launch, bilinear gathers from two time levels of the
// interleaved background field F[y][x][5] = (u, v, ux, uy, vx), vy = -ux.
//
// Reference semantics: raytracing/GPURaytracing.jl:18-65 (dxkdt, texture-coordinate bilinear
// sampling with wrap addressing, dispersion relation with frequency sign), :67-109
// (interpolate_velocity!/gradients!), raytracing/GPUTwoLayerRaytracing.jl:136-138 (k-cutoff).
#pragma once
#include <cuda_runtime.h>

namespace swrt {

struct PacketGrid {
    int nx, ny;
    double x0, y0, dx, dy;
};

struct Sample5 {
    double u, v, ux, uy, vx;
};

// s = (x - x0)/dx ; i = floor(s) mod n ; a = s - floor(s)   (division kept: index parity with the oracle)
__device__ __forceinline__ void cell(double x, double x0, double dx, int n, int& i0, int& i1, double& a) {
    const double s = (x - x0) / dx;
    const double fl = floor(s);
    a = s - fl;
    long long ii = (long long)fl % n;
    if (ii < 0) ii += n;
    i0 = (int)ii;
    i1 = i0 + 1 == n ? 0 : i0 + 1;
}

__device__ __forceinline__ void bilinear5(const double* __restrict__ F, const PacketGrid& g, int i0, int i1, int j0, int j1,
                                          double a, double b, double (&out)[5]) {
    const double* p00 = F + ((long long)j0 * g.nx + i0) * 5;
    const double* p10 = F + ((long long)j0 * g.nx + i1) * 5;
    const double* p01 = F + ((long long)j1 * g.nx + i0) * 5;
    const double* p11 = F + ((long long)j1 * g.nx + i1) * 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double bottom = (1.0 - a) * __ldg(p00 + c) + a * __ldg(p10 + c);
        const double top = (1.0 - a) * __ldg(p01 + c) + a * __ldg(p11 + c);
        out[c] = (1.0 - b) * bottom + b * top;
    }
}

struct RayParams {
    double f, Cg, t0, t1;
    int nsub, lerp;  // lerp: 0 physical ((1-a) old + a new), 1 reference GPU (a old + (1-a) new)
};

__device__ __forceinline__ void ray_rhs(const double (&s)[4], double sign, double alpha, const double* __restrict__ Fo,
                                        const double* __restrict__ Fn, const PacketGrid& g, const RayParams& p,
                                        double (&d)[4]) {
    int i0, i1, j0, j1;
    double a, b;
    cell(s[0], g.x0, g.dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.dy, g.ny, j0, j1, b);
    double So[5], Sn[5];
    bilinear5(Fo, g, i0, i1, j0, j1, a, b, So);
    bilinear5(Fn, g, i0, i1, j0, j1, a, b, Sn);
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
    double W[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) W[c] = wo * So[c] + wn * Sn[c];
    const double k = s[2], l = s[3];
    const double w = sign * sqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));
    d[0] = W[0] + p.Cg * p.Cg * k / w;
    d[1] = W[1] + p.Cg * p.Cg * l / w;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

// xk: (N,4) column-major = 4 arrays of N
__global__ void __launch_bounds__(128) raytrace_rk4_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                           const double* __restrict__ Fo, const double* __restrict__ Fn,
                                                           PacketGrid g, RayParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[4] = {xk[i], xk[n + i], xk[2 * n + i], xk[3 * n + i]};
    const double sg = sign[i];
    const double h = (p.t1 - p.t0) / p.nsub, span = p.t1 - p.t0;
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k1[4], k2[4], k3[4], k4[4], y[4];
        ray_rhs(s, sg, (t - p.t0) / span, Fo, Fn, g, p, k1);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k1[c];
        ray_rhs(y, sg, (t + 0.5 * h - p.t0) / span, Fo, Fn, g, p, k2);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k2[c];
        ray_rhs(y, sg, (t + 0.5 * h - p.t0) / span, Fo, Fn, g, p, k3);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + h * k3[c];
        ray_rhs(y, sg, (t + h - p.t0) / span, Fo, Fn, g, p, k4);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
    }
    xk[i] = s[0];
    xk[n + i] = s[1];
    xk[2 * n + i] = s[2];
    xk[3 * n + i] = s[3];
}

// interpolate_velocity!/gradients!: U (N,2), Gd (N,4) = ux, uy, vx, vy
__global__ void __launch_bounds__(128) sample_kernel(const double* __restrict__ xk, long long n, const double* __restrict__ F,
                                                     PacketGrid g, double* __restrict__ U, double* __restrict__ Gd) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int i0, i1, j0, j1;
    double a, b, S[5];
    cell(xk[i], g.x0, g.dx, g.nx, i0, i1, a);
    cell(xk[n + i], g.y0, g.dy, g.ny, j0, j1, b);
    bilinear5(F, g, i0, i1, j0, j1, a, b, S);
    U[i] = S[0];
    U[n + i] = S[1];
    if (Gd) {
        Gd[i] = S[2];
        Gd[n + i] = S[3];
        Gd[2 * n + i] = S[4];
        Gd[3 * n + i] = -S[2];
    }
}

__global__ void kcutoff_kernel(double* __restrict__ xk, long long n, double kc2, double k0, unsigned long long* count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double k = xk[2 * n + i], l = xk[3 * n + i];
    if (k * k + l * l >= kc2) {
        xk[2 * n + i] = k0;
        xk[3 * n + i] = 0.0;
        atomicAdd(count, 1ULL);
    }
}

// generate_initial_wavepackets (raytracing/RaytracingDriver.jl:27-47); p = global 1-based packet index
__global__ void generate_packets_kernel(double* __restrict__ xk, double* __restrict__ sign, long long n, long long first,
                                        long long sqrtN, double L, double k0) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long p0 = first + i;  // 0-based global
    const long long ntot = sqrtN * sqrtN;
    const double offset = L / (double)sqrtN / 2.0;
    const long long ix = p0 % sqrtN + 1, iy = p0 / sqrtN + 1;
    xk[i] = (double)ix * L / (double)sqrtN - L / 2.0 - offset;
    xk[n + i] = (double)iy * L / (double)sqrtN - L / 2.0 - offset;
    const double phase = 2.0 * 3.141592653589793 * (double)(p0 + 1) / (double)ntot;
    xk[2 * n + i] = k0 * cos(phase);
    xk[3 * n + i] = k0 * sin(phase);
    sign[i] = (p0 % 2 == 0) ? -1.0 : 1.0;
}

}  // namespace swrt
