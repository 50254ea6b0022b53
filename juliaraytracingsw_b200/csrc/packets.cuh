// Wave-packet ray tracer kernels (sm_100a, fp64): one thread per packet, state in registers,
// `nsub` classical RK4 steps per launch, bilinear gathers from the two time levels of the
// background field.
//
// Field layout (snapshot_layout.cuh): one array per time level, S[y][x][6] doubles = (u, v, ux, uy, vx, pad), 48 B =
// three 16-byte vectors per point, so the two x-taps of a bilinear stencil are 96 contiguous bytes (six LDG.128).
// Which array is "old" is a launch parameter (the snapshot kernel overwrites the other one each flow step).
//
// Locality: packets are kept sorted by an 8x8-cell-tiled cell key (counting sort, re-run every
// `sort_every` raytrace calls); `idx` remembers each packet's original row so that every
// host-visible array is returned in the caller's order, bit for bit (per-packet arithmetic does
// not depend on the storage order).
//
// Reference semantics: raytracing/GPURaytracing.jl:18-65 (dxkdt, texture-coordinate bilinear
// sampling with wrap addressing, dispersion relation with frequency sign), :67-109
// (interpolate_velocity!/gradients!), raytracing/GPUTwoLayerRaytracing.jl:136-138 (k-cutoff).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "snapshot_layout.cuh"

namespace swrt {

// Sort tiles: TILE x TILE cells; the packets of one tile are contiguous after the cell sort (cell_key), which is what lets a
// CTA of the tile kernel own one tile and stage its node records in shared memory.
constexpr int TILE_SHIFT = 4, TILE = 1 << TILE_SHIFT;
constexpr int TILE_MARGIN = 3;                           // cells of slack around the tile (packets drift between two sorts)
constexpr int PATCH = TILE + 2 * TILE_MARGIN + 1;        // nodes per side of the staged patch (23)
// Row pitch of the patch in shared memory, in doubles.  A stencil fill is LDS.128s whose lanes sit in a small neighbourhood of
// cells (the sorted order is up to 16 steps stale): cell (r, c) starts at 16-byte chunk r * pitch + 3 c, and the eight bank
// groups of a wavefront are chunk mod 8.  Pitch = 69 chunks (the dense 23 nodes, = 5 mod 8) collides (r, c) with (r + 1, c + 1),
// pitch = 72 (= 0) collides every row with the next (ncu: 29 % / 27 % excess shared wavefronts, profiles/r02_b, r02_d); 73 chunks
// = 1 (mod 8) keeps a 3 x 3 neighbourhood of cells on distinct bank groups except one pair.
constexpr int PATCH_ROW = 146;                           // doubles per patch row (73 chunks = 1168 B >= 23 nodes); the TMA box is PATCH_ROW x PATCH
static_assert(PATCH_ROW >= PATCH * SNAP_STRIDE && PATCH_ROW % 2 == 0 && PATCH_ROW <= 256, "patch row must hold the nodes, stay 16-byte granular and fit a TMA box");
constexpr int PATCH_BYTES = (PATCH * PATCH_ROW * 8 + 127) / 128 * 128;   // one level, padded to the TMA destination alignment
constexpr int TILE_THREADS = 128;
constexpr int TILE_STAGE_BYTES = 2 * 5 * TILE_THREADS * 8;               // double-buffered packet state of the CTA's next round
constexpr int TILE_SMEM_BYTES = 2 * PATCH_BYTES + TILE_STAGE_BYTES;

struct PacketGrid {
    int nx, ny;
    double x0, y0, dx, dy, inv_dx, inv_dy;
    long long ld;          // column stride of the packet arrays ((N,4) column-major: n on one GPU, the capacity in band mode)
    // Band mode (team of GPUs, y-band-sharded packets): the snapshot arrays hold only rows [jb, jb + jrows) of the grid (this
    // rank's band plus halo rows, jb may be negative = wraps), so a grid row j lives at local row (j - jb) mod ny.
    int band, jb, jrows;
    int tile_row0;         // band mode: first tile row the tile kernel's grid covers (wraps)
    int nw;                // type-2 NUFFT mode: kernel width in nodes of the (2x oversampled) snapshot grid
    double nbeta;          //                    shape parameter of the "exponential of semicircle" kernel
};
// grid rows (j0, j0 + 1) of a bilinear stencil -> rows of the snapshot array
__device__ __forceinline__ void stencil_rows(const PacketGrid& g, int& j0, int& j1) {
    if (g.band) {
        int jl = (j0 - g.jb) & (g.ny - 1);
        if (jl > g.jrows - 2) jl = g.jrows - 2;      // outside band + halo: clamped here, counted at the next sort (sort_hist_kernel)
        j0 = jl;
        j1 = jl + 1;
    }
}

// s = (x - x0)/dx ; i = floor(s) mod n ; a = s - floor(s).  The division is a multiplication by 1/dx (fp64 division
// costs ~15 issue slots): s can differ from the oracle's by one ulp, which moves the interpolated value by O(1e-16)
// (bilinear interpolation is continuous across cell edges).
__device__ __forceinline__ void cell(double x, double x0, double inv_dx, int n, int& i0, int& i1, double& a) {
    const double s = (x - x0) * inv_dx;
    const double fl = floor(s);
    a = s - fl;
    // n is a power of two (api.cu: supported_n): the periodic wrap is a mask, also for negative cells
    i0 = (int)((long long)fl & (long long)(n - 1));
    i1 = (i0 + 1) & (n - 1);
}

// one x-row of the stencil of one level: (1-a) S[j][i0] + a S[j][i1] for the five fields
__device__ __forceinline__ void lerp_row(const double* __restrict__ S, long long p0, long long p1, double a, double (&r)[5]) {
    const double2* q0 = reinterpret_cast<const double2*>(S + p0 * SNAP_STRIDE);
    const double2* q1 = reinterpret_cast<const double2*>(S + p1 * SNAP_STRIDE);
    double2 u[3], v[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) u[q] = __ldg(q0 + q);
#pragma unroll
    for (int q = 0; q < 3; ++q) v[q] = __ldg(q1 + q);
    r[0] = (1.0 - a) * u[0].x + a * v[0].x;
    r[1] = (1.0 - a) * u[0].y + a * v[0].y;
    r[2] = (1.0 - a) * u[1].x + a * v[1].x;
    r[3] = (1.0 - a) * u[1].y + a * v[1].y;
    r[4] = (1.0 - a) * u[2].x + a * v[2].x;
}

__device__ __forceinline__ void bilinear5(const double* __restrict__ S, const PacketGrid& g, int i0, int i1, int j0, int j1,
                                          double a, double b, double (&out)[5]) {
    double bottom[5], top[5];
    lerp_row(S, (long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, a, bottom);
    lerp_row(S, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1, a, top);
#pragma unroll
    for (int c = 0; c < 5; ++c) out[c] = (1.0 - b) * bottom[c] + b * top[c];
}

struct RayParams {
    double f, Cg, t0, t1;
    int nsub, lerp;  // lerp: 0 physical ((1-a) old + a new), 1 reference GPU (a old + (1-a) new)
};

__device__ __forceinline__ void ray_rhs(const double (&s)[4], double sign, double alpha, const double* __restrict__ So,
                                        const double* __restrict__ Sn, const PacketGrid& g, const RayParams& p, double (&d)[4]) {
    int i0, i1, j0, j1;
    double a, b;
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    stencil_rows(g, j0, j1);
    double o[5], nw[5], W[5];
    bilinear5(So, g, i0, i1, j0, j1, a, b, o);
    bilinear5(Sn, g, i0, i1, j0, j1, a, b, nw);
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
#pragma unroll
    for (int c = 0; c < 5; ++c) W[c] = wo * o[c] + wn * nw[c];
    const double k = s[2], l = s[3];
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));   // Cg^2 / omega
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

// xk: (N,4) column-major = 4 arrays of N
template <int MINB>
__global__ void __launch_bounds__(128, MINB) raytrace_rk4_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                                 const double* __restrict__ So, const double* __restrict__ Sn,
                                                                 PacketGrid g, RayParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[4] = {__ldcs(xk + i), __ldcs(xk + g.ld + i), __ldcs(xk + 2 * g.ld + i), __ldcs(xk + 3 * g.ld + i)};   // streaming: read once
    const double sg = sign[i];
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k1[4], k2[4], k3[4], k4[4], y[4];
        ray_rhs(s, sg, (t - p.t0) * inv_span, So, Sn, g, p, k1);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k1[c];
        ray_rhs(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k2);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k2[c];
        ray_rhs(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k3);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = s[c] + h * k3[c];
        ray_rhs(y, sg, (t + h - p.t0) * inv_span, So, Sn, g, p, k4);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
    }
    __stcs(xk + i, s[0]);
    __stcs(xk + g.ld + i, s[1]);
    __stcs(xk + 2 * g.ld + i, s[2]);
    __stcs(xk + 3 * g.ld + i, s[3]);
}

// ---------------------------------------------------------------- stencil-cached variant
// The four RK4 stages of a packet almost always fall into the same grid cell (a step moves a packet by
// <= ~0.2 cells), so the 2x2x(2 levels x 5 fields) stencil is kept in registers and only re-gathered when
// the cell changes: ~4x fewer L1 wavefronts, which is what bounds the plain kernel (ncu: l1tex data-pipe
// 76 %).  A time level whose lerp weight is exactly 0 (alpha = 0 at stage 1, alpha = 1 at stage 4) is
// skipped; 0*x + 1*y == y, so the result equals the full formula bit for bit for finite fields.
struct Stencil {   // [level][corner 00,10,01,11][3 vectors]
    double2 c[2][4][3];
    int ci, cj;
};

// weighted sum of the four corners of one level: out[f] (+)= sum_c w[c] * field f at corner c   (1 DMUL + 3 DFMA per field)
template <bool ACC>
__device__ __forceinline__ void corners5(const double2 (&c)[4][3], const double (&w)[4], double (&out)[5]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        const int q = f >> 1;
        const double v00 = (f & 1) ? c[0][q].y : c[0][q].x, v10 = (f & 1) ? c[1][q].y : c[1][q].x;
        const double v01 = (f & 1) ? c[2][q].y : c[2][q].x, v11 = (f & 1) ? c[3][q].y : c[3][q].x;
        double r = ACC ? fma(w[0], v00, out[f]) : w[0] * v00;
        r = fma(w[1], v10, r);
        r = fma(w[2], v01, r);
        out[f] = fma(w[3], v11, r);
    }
}

__device__ __forceinline__ void ray_rhs_cached(const double (&s)[4], double sign, double alpha, const double* __restrict__ So,
                                               const double* __restrict__ Sn, const PacketGrid& g, const RayParams& p, Stencil& st,
                                               double (&d)[4]) {
    int i0, i1, j0, j1;
    double a, b;
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    if (i0 != st.ci || j0 != st.cj) {
        st.ci = i0;
        st.cj = j0;
        stencil_rows(g, j0, j1);
        const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
#pragma unroll
        for (int lev = 0; lev < 2; ++lev) {
            const double* S = lev == 0 ? So : Sn;
#pragma unroll
            for (int cr = 0; cr < 4; ++cr) {
                const double2* q = reinterpret_cast<const double2*>(S + pt[cr] * SNAP_STRIDE);
                st.c[lev][cr][0] = __ldg(q);
                st.c[lev][cr][1] = __ldg(q + 1);
                st.c[lev][cr][2].x = __ldg(reinterpret_cast<const double*>(q + 2));
            }
        }
    }
    // W = wo * bilinear(old) + wn * bilinear(new) evaluated as one weighted sum over the (up to) eight stencil values per
    // field: the same polynomial as the oracle's nested lerps, associated differently (agreement ~1e-16 relative)
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
    const double a1 = 1.0 - a, b1 = 1.0 - b;
    const double wb[4] = {a1 * b1, a * b1, a1 * b, a * b};
    double W[5];
    if (wn == 0.0) {
        const double w[4] = {wo * wb[0], wo * wb[1], wo * wb[2], wo * wb[3]};
        corners5<false>(st.c[0], w, W);
    } else if (wo == 0.0) {
        const double w[4] = {wn * wb[0], wn * wb[1], wn * wb[2], wn * wb[3]};
        corners5<false>(st.c[1], w, W);
    } else {
        const double w0[4] = {wo * wb[0], wo * wb[1], wo * wb[2], wo * wb[3]};
        const double w1[4] = {wn * wb[0], wn * wb[1], wn * wb[2], wn * wb[3]};
        corners5<false>(st.c[0], w0, W);
        corners5<true>(st.c[1], w1, W);
    }
    const double k = s[2], l = s[3];
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));   // Cg^2 / omega
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

__device__ __forceinline__ void raytrace_rk4_cached_body(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                         const double* __restrict__ So, const double* __restrict__ Sn,
                                                         const PacketGrid& g, const RayParams& p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the packet state is read and written exactly once per launch: streaming accesses leave L2 to the node records
    double s[4] = {__ldcs(xk + i), __ldcs(xk + g.ld + i), __ldcs(xk + 2 * g.ld + i), __ldcs(xk + 3 * g.ld + i)};
    const double sg = __ldcs(sign + i);
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    Stencil st;
    st.ci = -1;
    st.cj = -1;
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k[4], acc[4], y[4];
        ray_rhs_cached(s, sg, (t - p.t0) * inv_span, So, Sn, g, p, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_cached(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_cached(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
        // the last stage of the last sub-step sits at t1: alpha = 1 exactly (the oracle's (t + h - t0)/(t1 - t0) can be 1 - ulp)
        const double a4 = it == p.nsub - 1 ? 1.0 : (t + h - p.t0) * inv_span;
        ray_rhs_cached(y, sg, a4, So, Sn, g, p, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
    }
    __stcs(xk + i, s[0]);
    __stcs(xk + g.ld + i, s[1]);
    __stcs(xk + 2 * g.ld + i, s[2]);
    __stcs(xk + 3 * g.ld + i, s[3]);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) raytrace_rk4_cached_kernel(double* __restrict__ xk, const double* __restrict__ sign,
                                                                        long long n, const double* __restrict__ So,
                                                                        const double* __restrict__ Sn, PacketGrid g, RayParams p) {
    raytrace_rk4_cached_body(xk, sign, n, So, Sn, g, p);
}
// ---------------------------------------------------------------- TMA-staged tile variant
// After the cell sort the packets of one TILE x TILE block of cells are contiguous (`tile_end` = the sort's per-key end
// offsets).  One CTA owns one tile: a single thread issues two bulk-tensor copies (cp.async.bulk.tensor.2d, one per time
// level) that land the (TILE + 2 MARGIN + 1)^2 patch of 48-byte node records in shared memory and complete on an mbarrier;
// while they fly the threads already fetch their first packet's state.  The 2x2x2 stencil of a packet is then filled from
// shared memory (LDS.128, ~30 cycles) instead of L1/L2 (LDG.128, 200-600 cycles): the plain stencil-cached kernel is bound by
// exactly that latency at 16 warps/SM (ncu long_scoreboard, profiles/r01_i).  A packet that has drifted out of the patch (more
// than MARGIN cells since the last sort) and the tiles on the domain boundary (their patch would wrap) use the global path.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

struct TilePatch {
    const double* lev[2];   // shared-memory patches of the old / new level
    int pi, pj;             // node (pi, pj) of the grid is patch node (0, 0)
    bool staged;            // false: boundary tile, everything from global memory
};

__device__ __forceinline__ void ray_rhs_tile(const double (&s)[4], double sign, double alpha, const double* __restrict__ So,
                                             const double* __restrict__ Sn, const PacketGrid& g, const RayParams& p, const TilePatch& tp,
                                             Stencil& st, double (&d)[4]) {
    int i0, i1, j0, j1;
    double a, b;
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    if (i0 != st.ci || j0 != st.cj) {
        st.ci = i0;
        st.cj = j0;
        // patch-relative node: masked differences, so a patch that wraps in y (band mode: halo rows beyond the domain edge) works;
        // in x staged tiles are interior and the mask is a no-op
        const unsigned ri = (unsigned)((i0 - tp.pi) & (g.nx - 1)), rj = (unsigned)((j0 - tp.pj) & (g.ny - 1));
        if (tp.staged && ri < (unsigned)(PATCH - 1) && rj < (unsigned)(PATCH - 1)) {
            const int o = rj * PATCH_ROW + ri * SNAP_STRIDE;
#pragma unroll
            for (int lev = 0; lev < 2; ++lev) {
                const double2* q = reinterpret_cast<const double2*>(tp.lev[lev] + o);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    st.c[lev][0][k] = q[k];
                    st.c[lev][1][k] = q[3 + k];
                    st.c[lev][2][k] = q[PATCH_ROW / 2 + k];
                    st.c[lev][3][k] = q[PATCH_ROW / 2 + 3 + k];
                }
                // fifth field (vx): an 8-byte load, the record's pad is never read (two wavefronts instead of four)
                st.c[lev][0][2].x = reinterpret_cast<const double*>(q + 2)[0];
                st.c[lev][1][2].x = reinterpret_cast<const double*>(q + 5)[0];
                st.c[lev][2][2].x = reinterpret_cast<const double*>(q + PATCH_ROW / 2 + 2)[0];
                st.c[lev][3][2].x = reinterpret_cast<const double*>(q + PATCH_ROW / 2 + 5)[0];
            }
        } else {
            stencil_rows(g, j0, j1);
            const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
#pragma unroll
            for (int lev = 0; lev < 2; ++lev) {
                const double* S = lev == 0 ? So : Sn;
#pragma unroll
                for (int cr = 0; cr < 4; ++cr) {
                    const double2* q = reinterpret_cast<const double2*>(S + pt[cr] * SNAP_STRIDE);
                    st.c[lev][cr][0] = __ldg(q);
                    st.c[lev][cr][1] = __ldg(q + 1);
                    st.c[lev][cr][2].x = __ldg(reinterpret_cast<const double*>(q + 2));
                }
            }
        }
    }
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
    const double a1 = 1.0 - a, b1 = 1.0 - b;
    const double wb[4] = {a1 * b1, a * b1, a1 * b, a * b};
    double W[5];
    if (wn == 0.0) {
        const double w[4] = {wo * wb[0], wo * wb[1], wo * wb[2], wo * wb[3]};
        corners5<false>(st.c[0], w, W);
    } else if (wo == 0.0) {
        const double w[4] = {wn * wb[0], wn * wb[1], wn * wb[2], wn * wb[3]};
        corners5<false>(st.c[1], w, W);
    } else {
        const double w0[4] = {wo * wb[0], wo * wb[1], wo * wb[2], wo * wb[3]};
        const double w1[4] = {wn * wb[0], wn * wb[1], wn * wb[2], wn * wb[3]};
        corners5<false>(st.c[0], w0, W);
        corners5<true>(st.c[1], w1, W);
    }
    const double k = s[2], l = s[3];
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));   // Cg^2 / omega
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

// 8-byte asynchronous copy global -> shared (LDGSTS), evict-first in L2: the packet state is touched once per launch
__device__ __forceinline__ void cp_async8_stream(void* smem_dst, const void* gsrc, unsigned long long policy) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "l"(policy));
}
template <int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB)
    raytrace_rk4_tile_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n, const double* __restrict__ So,
                             const double* __restrict__ Sn, const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapN,
                             const unsigned* __restrict__ tile_end, PacketGrid g, RayParams p) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tiles_x = g.nx >> TILE_SHIFT, tiles_y = g.ny >> TILE_SHIFT;
    const int tjl = blockIdx.x / tiles_x, ti = blockIdx.x - tjl * tiles_x;
    const int tj = (g.tile_row0 + tjl) % tiles_y;                // band mode: the grid covers the tile rows of band + halo only
    const int tile = tj * tiles_x + ti;
    const long long key0 = (long long)tile << (2 * TILE_SHIFT);
    const long long start = tile == 0 ? 0 : (long long)tile_end[key0 - 1], end = (long long)tile_end[key0 + (TILE * TILE - 1)];
    if (start >= end) return;                                     // empty tile (uniform over the CTA)
    TilePatch tp;
    tp.lev[0] = reinterpret_cast<const double*>(tile_smem);
    tp.lev[1] = reinterpret_cast<const double*>(tile_smem + PATCH_BYTES);
    tp.pi = ti * TILE - TILE_MARGIN;
    tp.pj = tj * TILE - TILE_MARGIN;
    int prow = tp.pj;                                            // row of the snapshot array where the patch starts
    if (g.band) {
        prow = (tp.pj - g.jb) & (g.ny - 1);
        if (prow >= g.ny / 2) prow -= g.ny;                      // signed distance from the first resident row
    }
    tp.staged = tp.pi >= 0 && prow >= 0 && tp.pi + PATCH <= g.nx && prow + PATCH <= (g.band ? g.jrows : g.ny);
    if (tp.staged) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, 2u * PATCH * PATCH_ROW * 8);   // (a box that sticks out of the tensor still delivers its full byte count, zero filled)
            tma_load_2d(tile_smem, &mapO, tp.pi * SNAP_STRIDE, prow, &bar);
            tma_load_2d(tile_smem + PATCH_BYTES, &mapN, tp.pi * SNAP_STRIDE, prow, &bar);
        }
    }
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    // The state of a thread's NEXT packet streams into shared memory (cp.async, no registers) while it integrates the current
    // one: the first use of a freshly loaded state was the largest single stall of the kernel (ncu: 24 % of all samples,
    // long_scoreboard).  A thread only ever reads the slots it copied itself, so no barrier is needed.
    double* stage = reinterpret_cast<double*>(tile_smem + 2 * PATCH_BYTES);     // [2][5][TILE_THREADS]
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    auto prefetch = [&](long long i, int buf) {
        if (i < end) {
            double* d = stage + buf * 5 * TILE_THREADS + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async8_stream(d + c * TILE_THREADS, xk + c * g.ld + i, pol);
            cp_async8_stream(d + 4 * TILE_THREADS, sign + i, pol);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    prefetch(start + threadIdx.x, 0);
    bool waited = !tp.staged;
    int buf = 0;
    for (long long i = start + threadIdx.x; i < end; i += TILE_THREADS, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        const double* sv = stage + buf * 5 * TILE_THREADS + threadIdx.x;
        double s[4] = {sv[0], sv[TILE_THREADS], sv[2 * TILE_THREADS], sv[3 * TILE_THREADS]};
        const double sg = sv[4 * TILE_THREADS];
        prefetch(i + TILE_THREADS, buf ^ 1);
        if (!waited) { mbar_wait(&bar, 0); waited = true; }      // the patch has landed (the first state copy overlapped it)
        Stencil st;
        st.ci = -1;
        st.cj = -1;
        for (int it = 0; it < p.nsub; ++it) {
            const double t = p.t0 + it * h;
            double k[4], acc[4], y[4];
            ray_rhs_tile(s, sg, (t - p.t0) * inv_span, So, Sn, g, p, tp, st, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
            ray_rhs_tile(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, tp, st, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
            ray_rhs_tile(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, tp, st, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
            const double a4 = it == p.nsub - 1 ? 1.0 : (t + h - p.t0) * inv_span;
            ray_rhs_tile(y, sg, a4, So, Sn, g, p, tp, st, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
        }
        __stcs(xk + i, s[0]);
        __stcs(xk + g.ld + i, s[1]);
        __stcs(xk + 2 * g.ld + i, s[2]);
        __stcs(xk + 3 * g.ld + i, s[3]);
    }
    // (thread 0 always owns packet `start` and waits for the copy above, so the CTA never retires with a copy in flight)
}

// ---------------------------------------------------------------- three-level tile variant (nsub == 1)
// With one RK4 step per flow step the four stages sample the background at exactly three times: the first level (lerp weight
// 1, 0), the midpoint (1/2, 1/2) twice, and the last level (0, 1).  The CTA therefore keeps THREE patches in shared memory --
// the two levels staged by TMA and their mean, computed once per tile by the CTA -- and every stage reads the 2x2 stencil of
// ONE patch: 20 doubles per thread instead of the 40 of the two-level stencil cache (128 registers, 16 warps per SM instead of
// 12 at 160), no lerp weights in time (64 fewer fp64 instructions per packet-step of ~306), no "has the cell changed" refill
// logic except between the two midpoint stages.  Bilinear interpolation is linear in the node values, so
// bilinear(mean of levels) = mean of bilinear(levels): the same polynomial as the oracle's, associated differently.
// Patches that wrap around the domain edge are filled by the CTA's threads (masked node indices) instead of TMA, so every tile
// whose rows are resident is staged; a packet that has left the patch gathers from global memory with the same arithmetic.
constexpr int TILE3_THREADS = 256;
constexpr int TILE3_STAGE_BYTES = 2 * 5 * TILE3_THREADS * 8;
constexpr int TILE3_SMEM_BYTES = 3 * PATCH_BYTES + TILE3_STAGE_BYTES;

struct Stencil1 {   // [corner 00,10,01,11][3 vectors] of one time level
    double2 c[4][3];
};
struct TilePatch3 {
    const double* lev[3];   // shared-memory patches: first level, mean, last level
    int pi, pj;
    bool staged;
};
__device__ __forceinline__ double2 mean2(double2 a, double2 b) { return make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y)); }

// LEV: 0 first level (S1), 1 mean of the two, 2 last level (S4)
template <int LEV>
__device__ __forceinline__ void fill_stencil1(Stencil1& st, int i0, int j0, const double* __restrict__ S1, const double* __restrict__ S4,
                                              const PacketGrid& g, const TilePatch3& tp) {
    const unsigned ri = (unsigned)((i0 - tp.pi) & (g.nx - 1)), rj = (unsigned)((j0 - tp.pj) & (g.ny - 1));
    if (tp.staged && ri < (unsigned)(PATCH - 1) && rj < (unsigned)(PATCH - 1)) {
        const double2* q = reinterpret_cast<const double2*>(tp.lev[LEV] + rj * PATCH_ROW + ri * SNAP_STRIDE);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            st.c[0][k] = q[k];
            st.c[1][k] = q[3 + k];
            st.c[2][k] = q[PATCH_ROW / 2 + k];
            st.c[3][k] = q[PATCH_ROW / 2 + 3 + k];
        }
        st.c[0][2].x = reinterpret_cast<const double*>(q + 2)[0];
        st.c[1][2].x = reinterpret_cast<const double*>(q + 5)[0];
        st.c[2][2].x = reinterpret_cast<const double*>(q + PATCH_ROW / 2 + 2)[0];
        st.c[3][2].x = reinterpret_cast<const double*>(q + PATCH_ROW / 2 + 5)[0];
    } else {
        const int i1 = (i0 + 1) & (g.nx - 1);
        int j1 = (j0 + 1) & (g.ny - 1);
        stencil_rows(g, j0, j1);
        const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
#pragma unroll
        for (int cr = 0; cr < 4; ++cr) {
            const double2* qa = reinterpret_cast<const double2*>((LEV == 2 ? S4 : S1) + pt[cr] * SNAP_STRIDE);
            st.c[cr][0] = __ldg(qa);
            st.c[cr][1] = __ldg(qa + 1);
            st.c[cr][2].x = __ldg(reinterpret_cast<const double*>(qa + 2));
            if (LEV == 1) {
                const double2* qb = reinterpret_cast<const double2*>(S4 + pt[cr] * SNAP_STRIDE);
                st.c[cr][0] = mean2(st.c[cr][0], __ldg(qb));
                st.c[cr][1] = mean2(st.c[cr][1], __ldg(qb + 1));
                st.c[cr][2].x = 0.5 * (st.c[cr][2].x + __ldg(reinterpret_cast<const double*>(qb + 2)));
            }
        }
    }
}
__device__ __forceinline__ void ray_rhs1(const Stencil1& st, double a, double b, double k, double l, double sign, const RayParams& p,
                                         double (&d)[4]) {
    const double a1 = 1.0 - a, b1 = 1.0 - b;
    const double w[4] = {a1 * b1, a * b1, a1 * b, a * b};
    double W[5];
    corners5<false>(st.c, w, W);
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));   // Cg^2 / omega
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

// one classical RK4 step of one packet over (t0, t0 + h) against the three patches (first level, mean, last level)
__device__ __forceinline__ void rk4_three_level(double (&s)[4], double sg, double h, const double* __restrict__ S1, const double* __restrict__ S4,
                                                const PacketGrid& g, const RayParams& p, const TilePatch3& tp) {
    Stencil1 st;
    int i0, i1, j0, j1, ci, cj;
    double a, b, k[4], acc[4], y[4];
    // stage 1: first level at t0
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1<0>(st, i0, j0, S1, S4, g, tp);
    ray_rhs1(st, a, b, s[2], s[3], sg, p, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
    // stages 2 and 3: mean of the levels at t0 + h/2
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1<1>(st, i0, j0, S1, S4, g, tp);
    ci = i0;
    cj = j0;
    ray_rhs1(st, a, b, y[2], y[3], sg, p, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    if (i0 != ci || j0 != cj) fill_stencil1<1>(st, i0, j0, S1, S4, g, tp);
    ray_rhs1(st, a, b, y[2], y[3], sg, p, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
    // stage 4: last level at t1
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1<2>(st, i0, j0, S1, S4, g, tp);
    ray_rhs1(st, a, b, y[2], y[3], sg, p, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
}

// S1 / map1: the level whose lerp weight is 1 at t0 (the old one for the physical convention, the new one for the reference's
// GPU convention, RayParams::lerp -- resolved by the caller); S4 / map4: the other one.
__global__ void __launch_bounds__(TILE3_THREADS, 2)
    raytrace_rk4_tile3_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n, const double* __restrict__ S1,
                              const double* __restrict__ S4, const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map4,
                              const unsigned* __restrict__ tile_end, PacketGrid g, RayParams p) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tiles_x = g.nx >> TILE_SHIFT, tiles_y = g.ny >> TILE_SHIFT;
    const int tjl = blockIdx.x / tiles_x, ti = blockIdx.x - tjl * tiles_x;
    const int tj = (g.tile_row0 + tjl) % tiles_y;
    const int tile = tj * tiles_x + ti;
    const long long key0 = (long long)tile << (2 * TILE_SHIFT);
    // (the tile record is read first but not used until the bulk copies are on their way: the two latencies overlap)
    const long long start = tile == 0 ? 0 : (long long)__ldg(&tile_end[key0 - 1]), end = (long long)__ldg(&tile_end[key0 + (TILE * TILE - 1)]);
    double* const patch1 = reinterpret_cast<double*>(tile_smem);
    double* const patchm = reinterpret_cast<double*>(tile_smem + PATCH_BYTES);
    double* const patch4 = reinterpret_cast<double*>(tile_smem + 2 * PATCH_BYTES);
    TilePatch3 tp;
    tp.lev[0] = patch1;
    tp.lev[1] = patchm;
    tp.lev[2] = patch4;
    tp.pi = ti * TILE - TILE_MARGIN;
    tp.pj = tj * TILE - TILE_MARGIN;
    int prow = tp.pj;                                            // row of the snapshot array where the patch starts
    bool rows_ok = true;                                         // all PATCH rows resident (band mode: inside band + halo)
    if (g.band) {
        prow = (tp.pj - g.jb) & (g.ny - 1);
        if (prow >= g.ny / 2) prow -= g.ny;                      // signed distance from the first resident row
        rows_ok = prow >= 0 && prow + PATCH <= g.jrows;
    }
    const bool by_tma = rows_ok && tp.pi >= 0 && tp.pi + PATCH <= g.nx && prow >= 0 && prow + PATCH <= (g.band ? g.jrows : g.ny);
    tp.staged = rows_ok;
    if (by_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, 2u * PATCH * PATCH_ROW * 8);
            tma_load_2d(patch1, &map1, tp.pi * SNAP_STRIDE, prow, &bar);
            tma_load_2d(patch4, &map4, tp.pi * SNAP_STRIDE, prow, &bar);
        }
    }
    if (start >= end) {                                           // empty tile (uniform over the CTA): no copy may be left in flight
        if (by_tma) mbar_wait(&bar, 0);
        return;
    }
    const double h = p.t1 - p.t0;
    double* stage = reinterpret_cast<double*>(tile_smem + 3 * PATCH_BYTES);     // [2][5][TILE3_THREADS]
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    auto prefetch = [&](long long i, int buf) {
        if (i < end) {
            double* d = stage + buf * 5 * TILE3_THREADS + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async8_stream(d + c * TILE3_THREADS, xk + c * g.ld + i, pol);
            cp_async8_stream(d + 4 * TILE3_THREADS, sign + i, pol);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    prefetch(start + threadIdx.x, 0);
    if (by_tma) {
        mbar_wait(&bar, 0);                                      // every thread observes the completed phase before reading the patches
        double2* m = reinterpret_cast<double2*>(patchm);
        const double2 *a = reinterpret_cast<const double2*>(patch1), *b = reinterpret_cast<const double2*>(patch4);
        constexpr int NCH = PATCH * PATCH_ROW / 2, NIT = (NCH + TILE3_THREADS - 1) / TILE3_THREADS;
        double2 va[NIT], vb[NIT];
#pragma unroll
        for (int u = 0; u < NIT; ++u) {
            const int c = threadIdx.x + u * TILE3_THREADS;
            if (c < NCH) { va[u] = a[c]; vb[u] = b[c]; }
        }
#pragma unroll
        for (int u = 0; u < NIT; ++u) {
            const int c = threadIdx.x + u * TILE3_THREADS;
            if (c < NCH) m[c] = mean2(va[u], vb[u]);
        }
        __syncthreads();
    } else if (rows_ok) {
        // the patch wraps around the domain edge (or starts left of column 0): filled node by node with masked indices
        for (int e = threadIdx.x; e < PATCH * PATCH * 3; e += TILE3_THREADS) {
            const int node = e / 3, q = e - node * 3, r = node / PATCH, c = node - r * PATCH;
            const int col = (tp.pi + c) & (g.nx - 1), row = g.band ? prow + r : ((tp.pj + r) & (g.ny - 1));
            const long long src = ((long long)row * g.nx + col) * SNAP_STRIDE + 2 * q;
            const double2 va = __ldg(reinterpret_cast<const double2*>(S1 + src)), vb = __ldg(reinterpret_cast<const double2*>(S4 + src));
            const int dst = r * PATCH_ROW + c * SNAP_STRIDE + 2 * q;
            *reinterpret_cast<double2*>(patch1 + dst) = va;
            *reinterpret_cast<double2*>(patch4 + dst) = vb;
            *reinterpret_cast<double2*>(patchm + dst) = mean2(va, vb);
        }
        __syncthreads();
    }
    int buf = 0;
    for (long long i = start + threadIdx.x; i < end; i += TILE3_THREADS, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        const double* sv = stage + buf * 5 * TILE3_THREADS + threadIdx.x;
        double s[4] = {sv[0], sv[TILE3_THREADS], sv[2 * TILE3_THREADS], sv[3 * TILE3_THREADS]};
        const double sg = sv[4 * TILE3_THREADS];
        prefetch(i + TILE3_THREADS, buf ^ 1);
        rk4_three_level(s, sg, h, S1, S4, g, p, tp);
        __stcs(xk + i, s[0]);
        __stcs(xk + g.ld + i, s[1]);
        __stcs(xk + 2 * g.ld + i, s[2]);
        __stcs(xk + 3 * g.ld + i, s[3]);
    }
}

// ---------------------------------------------------------------- persistent, warp-specialised three-level variant
// One CTA per SM walks through the tiles (dynamic hand-out, ascending: the ~148 tiles in flight are neighbours, their halos hit
// L2).  Warp 15 is the PRODUCER: for the next tile it issues the two bulk-tensor copies into the other buffer set, waits for them,
// writes the mean patch and publishes (tile, packet range) -- all while the 15 CONSUMER warps integrate the packets of the current
// tile.  Three mbarriers per buffer set: `full` (TMA bytes landed), `ready` (producer: mean patch + tile record written),
// `empty` (one arrival per consumer warp: nobody reads this set any more).  No CTA-wide barrier anywhere: a consumer warp that
// has finished its share of a tile moves on to the next, and the packets of a tile are dealt round-robin CONTINUING where the
// previous tile stopped (`rot`), so every consumer thread gets the same number of packets (+-1) over the launch.  Against the
// one-CTA-per-tile kernel this removes the start-up chain of every tile (tile record -> TMA -> mean patch: ~2 us of a ~9 us CTA
// life at half occupancy, ncu: 25 % of the stall samples) and the last, nearly empty round of every tile.
// MEASURED (profiles/r02_n, r02_o): 0.62-0.68 ms against 0.52 ms for the one-CTA-per-tile kernel at 16.8 M packets.  Shared memory
// only holds TWO buffer sets of three patches, a tile is ~2.1 packets per consumer thread (~4 us) and the producer's chain
// empty -> TMA -> mean patch -> ready takes ~2 us, starting only when the slowest warp has left the previous tile: the consumer
// warps wait for each other at every tile.  Kept selectable (SWRT_RAYKERNEL_PIPE), not the default.
constexpr int PIPE_THREADS = 512, PIPE_CONSUMERS = PIPE_THREADS - 32, PIPE_CWARPS = PIPE_CONSUMERS / 32;
constexpr int PIPE_STAGE_BYTES = 2 * 5 * PIPE_CONSUMERS * 8;
constexpr int PIPE_SMEM_BYTES = 6 * PATCH_BYTES + PIPE_STAGE_BYTES;
struct PipeTile {   // published by the producer with `ready`
    int tile;       // -1: no more work
    int staged;
    int pi, pj;     // grid node of patch node (0, 0)
    long long start, end;
};
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// sched = {next tile, finished CTAs}; the last CTA re-arms it for the next launch
__global__ void __launch_bounds__(PIPE_THREADS, 1)
    raytrace_rk4_pipe_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n, const double* __restrict__ S1,
                             const double* __restrict__ S4, const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map4,
                             const unsigned* __restrict__ tile_end, int ntiles, unsigned* __restrict__ sched, PacketGrid g, RayParams p) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long full[2], ready[2], empty[2];
    __shared__ PipeTile rec[2];
    const int tiles_x = g.nx >> TILE_SHIFT, tiles_y = g.ny >> TILE_SHIFT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&ready[b], 1);
            mbar_init(&empty[b], PIPE_CWARPS);
        }
    }
    __syncthreads();
    auto patch = [&](int b, int lev) { return reinterpret_cast<double*>(tile_smem + (size_t)(3 * b + lev) * PATCH_BYTES); };
    // patch origin of a tile and whether / how it can be staged
    auto tile_origin = [&](int tile, int& pi, int& pj, int& prow, bool& rows_ok, bool& by_tma) {
        const int tjl = tile / tiles_x, ti = tile - tjl * tiles_x;
        const int tj = (g.tile_row0 + tjl) % tiles_y;
        pi = ti * TILE - TILE_MARGIN;
        pj = tj * TILE - TILE_MARGIN;
        prow = pj;
        rows_ok = true;
        if (g.band) {
            prow = (pj - g.jb) & (g.ny - 1);
            if (prow >= g.ny / 2) prow -= g.ny;
            rows_ok = prow >= 0 && prow + PATCH <= g.jrows;
        }
        by_tma = rows_ok && pi >= 0 && pi + PATCH <= g.nx && prow >= 0 && prow + PATCH <= (g.band ? g.jrows : g.ny);
        return tj * tiles_x + ti;                                   // global tile number (band mode: the grid covers a window of tile rows)
    };
    if (warp == PIPE_CWARPS) {
        // ------------------------------------------------------------ producer warp
        // Tile records are fetched eight at a time (one atomic per batch, lanes 0-7 load the packet ranges), one batch ahead of
        // their use, so neither the atomic nor the loads sit on the per-tile chain  empty -> TMA -> mean patch -> ready.
        int q_slot, n_slot;
        long long q_start, q_end, n_start, n_end;
        auto fetch = [&](int& slot, long long& st, long long& en) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&sched[0], 8u);
            base = __shfl_sync(0xffffffffu, base, 0);
            slot = lane < 8 && base + lane < (unsigned)ntiles ? (int)(base + lane) : ntiles;
            st = en = 0;
            if (slot < ntiles) {
                int pi, pj, prow; bool r_ok, tma;
                const int tile = tile_origin(slot, pi, pj, prow, r_ok, tma);
                const long long key0 = (long long)tile << (2 * TILE_SHIFT);
                st = tile == 0 ? 0 : (long long)__ldg(&tile_end[key0 - 1]);
                en = (long long)__ldg(&tile_end[key0 + (TILE * TILE - 1)]);
            }
        };
        fetch(q_slot, q_start, q_end);
        fetch(n_slot, n_start, n_end);
        int qpos = 0;
        for (int it = 0;; ++it) {
            const int b = it & 1, use = it >> 1;
            int tslot = 0;
            long long start = 0, end = 0;
            for (;;) {                                              // next non-empty tile
                if (qpos == 8) {
                    q_slot = n_slot; q_start = n_start; q_end = n_end;
                    fetch(n_slot, n_start, n_end);
                    qpos = 0;
                }
                tslot = __shfl_sync(0xffffffffu, q_slot, qpos);
                start = __shfl_sync(0xffffffffu, q_start, qpos);
                end = __shfl_sync(0xffffffffu, q_end, qpos);
                ++qpos;
                if (tslot >= ntiles || start < end) break;
            }
            if (use >= 1) mbar_wait(&empty[b], (unsigned)((use - 1) & 1));   // every consumer warp has left this buffer set
            if (tslot >= ntiles) {
                if (lane == 0) {
                    rec[b].tile = -1;
                    mbar_arrive(&ready[b]);
                }
                break;
            }
            int pi, pj, prow; bool rows_ok, by_tma;
            tile_origin(tslot, pi, pj, prow, rows_ok, by_tma);
            double *p1 = patch(b, 0), *pm = patch(b, 1), *p4 = patch(b, 2);
            if (by_tma) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // earlier generic-proxy accesses of this set before the async-proxy writes
                    mbar_expect_tx(&full[b], 2u * PATCH * PATCH_ROW * 8);
                    tma_load_2d(p1, &map1, pi * SNAP_STRIDE, prow, &full[b]);
                    tma_load_2d(p4, &map4, pi * SNAP_STRIDE, prow, &full[b]);
                }
                mbar_wait(&full[b], (unsigned)(use & 1));
                double2* m = reinterpret_cast<double2*>(pm);
                const double2 *a = reinterpret_cast<const double2*>(p1), *c = reinterpret_cast<const double2*>(p4);
                constexpr int NCH = PATCH * PATCH_ROW / 2;
#pragma unroll 1
                for (int c0 = lane; c0 < NCH; c0 += 32 * 4) {
                    double2 va[4], vc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (c0 + 32 * u < NCH) { va[u] = a[c0 + 32 * u]; vc[u] = c[c0 + 32 * u]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (c0 + 32 * u < NCH) m[c0 + 32 * u] = mean2(va[u], vc[u]);
                }
            } else if (rows_ok) {
                // the patch wraps around the domain edge: filled node by node with masked indices
                if (lane == 0) mbar_arrive(&full[b]);                // keeps `full` in step with the buffer's use count (no bytes expected)
#pragma unroll 1
                for (int e0 = lane; e0 < PATCH * PATCH * 3; e0 += 32 * 4) {
                    double2 va[4], vb[4];
                    int dst[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + 32 * u;
                        if (e < PATCH * PATCH * 3) {
                            const int node = e / 3, q = e - node * 3, r = node / PATCH, c = node - r * PATCH;
                            const int col = (pi + c) & (g.nx - 1), row = g.band ? prow + r : ((pj + r) & (g.ny - 1));
                            const long long src = ((long long)row * g.nx + col) * SNAP_STRIDE + 2 * q;
                            va[u] = __ldg(reinterpret_cast<const double2*>(S1 + src));
                            vb[u] = __ldg(reinterpret_cast<const double2*>(S4 + src));
                            dst[u] = r * PATCH_ROW + c * SNAP_STRIDE + 2 * q;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e0 + 32 * u < PATCH * PATCH * 3) {
                            *reinterpret_cast<double2*>(p1 + dst[u]) = va[u];
                            *reinterpret_cast<double2*>(p4 + dst[u]) = vb[u];
                            *reinterpret_cast<double2*>(pm + dst[u]) = mean2(va[u], vb[u]);
                        }
                    }
                }
            } else {
                if (lane == 0) mbar_arrive(&full[b]);                // not staged (band mode, rows not resident): global path
            }
            __syncwarp();
            if (lane == 0) {
                rec[b].tile = tslot;
                rec[b].staged = rows_ok ? 1 : 0;
                rec[b].pi = pi;
                rec[b].pj = pj;
                rec[b].start = start;
                rec[b].end = end;
                mbar_arrive(&ready[b]);                               // release: the patches and the record are visible to whoever observes the phase
            }
        }
    } else {
        // ------------------------------------------------------------ consumer warps
        const int ctid = threadIdx.x;                                 // 0 .. PIPE_CONSUMERS-1
        const double h = p.t1 - p.t0;
        double* stage = reinterpret_cast<double*>(tile_smem + 6 * PATCH_BYTES);     // [2][5][PIPE_CONSUMERS]
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
        auto prefetch = [&](long long i, int buf) {
            double* d = stage + buf * 5 * PIPE_CONSUMERS + ctid;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async8_stream(d + c * PIPE_CONSUMERS, xk + c * g.ld + i, pol);
            cp_async8_stream(d + 4 * PIPE_CONSUMERS, sign + i, pol);
            asm volatile("cp.async.commit_group;\n" ::);
        };
        int rot = 0, buf = 0;
        long long pf = -1;                                            // packet whose state sits (or is landing) in stage[buf]
        for (int it = 0;; ++it) {
            const int b = it & 1;
            mbar_wait(&ready[b], (unsigned)((it >> 1) & 1));
            const int tslot = rec[b].tile;
            if (tslot < 0) break;
            const long long start = rec[b].start, end = rec[b].end;
            TilePatch3 tp;
            tp.lev[0] = patch(b, 0);
            tp.lev[1] = patch(b, 1);
            tp.lev[2] = patch(b, 2);
            tp.staged = rec[b].staged != 0;
            tp.pi = rec[b].pi;
            tp.pj = rec[b].pj;
            int r = ctid - rot;
            if (r < 0) r += PIPE_CONSUMERS;
            const int ntile = (int)((end - start) % PIPE_CONSUMERS);
            int rot_next = rot + ntile;
            if (rot_next >= PIPE_CONSUMERS) rot_next -= PIPE_CONSUMERS;
            for (long long i = start + r; i < end; i += PIPE_CONSUMERS) {
                if (pf != i) prefetch(i, buf);                        // (first packet of the launch, or the look-ahead below was not possible)
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                const double* sv = stage + buf * 5 * PIPE_CONSUMERS + ctid;
                double s[4] = {sv[0], sv[PIPE_CONSUMERS], sv[2 * PIPE_CONSUMERS], sv[3 * PIPE_CONSUMERS]};
                const double sg = sv[4 * PIPE_CONSUMERS];
                // look ahead: this thread's next packet, in this tile or -- if the producer has already published it -- the next one
                long long nxt = i + PIPE_CONSUMERS;
                if (nxt >= end) {
                    nxt = -1;
                    const int bn = b ^ 1;
                    if (mbar_test(&ready[bn], (unsigned)(((it + 1) >> 1) & 1)) && rec[bn].tile >= 0) {
                        int rn = ctid - rot_next;
                        if (rn < 0) rn += PIPE_CONSUMERS;
                        const long long cand = rec[bn].start + rn;
                        if (cand < rec[bn].end) nxt = cand;
                    }
                }
                buf ^= 1;
                pf = nxt;
                if (nxt >= 0) prefetch(nxt, buf);
                rk4_three_level(s, sg, h, S1, S4, g, p, tp);
                __stcs(xk + i, s[0]);
                __stcs(xk + g.ld + i, s[1]);
                __stcs(xk + 2 * g.ld + i, s[2]);
                __stcs(xk + 3 * g.ld + i, s[3]);
            }
            rot = rot_next;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);                    // this warp reads the buffer set no more
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
        }
    }
}

// ---------------------------------------------------------------- Hermite-bicubic mode
// u, v interpolated from (f, f_x, f_y, f_xy) node data (utils/CUDAInterpolations.jl:39-53,71-108); the gradient that enters
// dk/dt is the analytic gradient of that interpolant.  Node record (snapshot_layout.cuh): u, v, ux, uy, vx, uxy, vxy, pad.
__device__ __forceinline__ double hcubic(double al, double f0, double f1, double m0, double m1) {
    return f0 + m0 * al + (-3.0 * f0 + 3.0 * f1 - 2.0 * m0 - m1) * (al * al) + (2.0 * f0 - 2.0 * f1 + m0 + m1) * (al * al * al);
}
__device__ __forceinline__ double hdcubic(double al, double f0, double f1, double m0, double m1) {
    return m0 + 2.0 * (-3.0 * f0 + 3.0 * f1 - 2.0 * m0 - m1) * al + 3.0 * (2.0 * f0 - 2.0 * f1 + m0 + m1) * (al * al);
}
// value and gradient of the bicubic of one scalar from its four corner records (f, fx, fy, fxy)
__device__ __forceinline__ void hermite2d(const double (&f)[4], const double (&fx)[4], const double (&fy)[4], const double (&fxy)[4],
                                          double a, double b, double dx, double dy, double& val, double& ddx, double& ddy) {
    // corner order 00, 10, 01, 11
    const double f0 = hcubic(a, f[0], f[1], fx[0] * dx, fx[1] * dx), f1 = hcubic(a, f[2], f[3], fx[2] * dx, fx[3] * dx);
    const double g0 = hcubic(a, fy[0] * dy, fy[1] * dy, fxy[0] * (dx * dy), fxy[1] * (dx * dy));
    const double g1 = hcubic(a, fy[2] * dy, fy[3] * dy, fxy[2] * (dx * dy), fxy[3] * (dx * dy));
    val = hcubic(b, f0, f1, g0, g1);
    const double d0 = hdcubic(a, f[0], f[1], fx[0] * dx, fx[1] * dx), d1 = hdcubic(a, f[2], f[3], fx[2] * dx, fx[3] * dx);
    const double e0 = hdcubic(a, fy[0] * dy, fy[1] * dy, fxy[0] * (dx * dy), fxy[1] * (dx * dy));
    const double e1 = hdcubic(a, fy[2] * dy, fy[3] * dy, fxy[2] * (dx * dy), fxy[3] * (dx * dy));
    ddx = hcubic(b, d0, d1, e0, e1) / dx;
    ddy = hdcubic(b, f0, f1, g0, g1) / dy;
}
// out = u, v, ux, uy, vx of one level at (i0 + a, j0 + b)
__device__ __forceinline__ void sample_hermite5(const double* __restrict__ S, const PacketGrid& g, int i0, int i1, int j0, int j1,
                                                double a, double b, double (&out)[5]) {
    const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
    double u[4], v[4], ux[4], uy[4], vx[4], uxy[4], vxy[4], vy[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double2* q = reinterpret_cast<const double2*>(S + pt[c] * SNAP3_STRIDE);
        const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
        u[c] = q0.x; v[c] = q0.y; ux[c] = q1.x; uy[c] = q1.y; vx[c] = q2.x; uxy[c] = q2.y; vxy[c] = q3.x;
        vy[c] = -q1.x;
    }
    double dvy;
    hermite2d(u, ux, uy, uxy, a, b, g.dx, g.dy, out[0], out[2], out[3]);
    hermite2d(v, vx, vy, vxy, a, b, g.dx, g.dy, out[1], out[4], dvy);
}

__device__ __forceinline__ void ray_rhs_cubic(const double (&s)[4], double sign, double alpha, const double* __restrict__ So,
                                              const double* __restrict__ Sn, const PacketGrid& g, const RayParams& p, double (&d)[4]) {
    int i0, i1, j0, j1;
    double a, b;
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
    double W[5];
    if (wn == 0.0) {
        double o[5];
        sample_hermite5(So, g, i0, i1, j0, j1, a, b, o);
#pragma unroll
        for (int c = 0; c < 5; ++c) W[c] = wo * o[c];
    } else if (wo == 0.0) {
        double nw[5];
        sample_hermite5(Sn, g, i0, i1, j0, j1, a, b, nw);
#pragma unroll
        for (int c = 0; c < 5; ++c) W[c] = wn * nw[c];
    } else {
        double o[5], nw[5];
        sample_hermite5(So, g, i0, i1, j0, j1, a, b, o);
        sample_hermite5(Sn, g, i0, i1, j0, j1, a, b, nw);
#pragma unroll
        for (int c = 0; c < 5; ++c) W[c] = wo * o[c] + wn * nw[c];
    }
    const double k = s[2], l = s[3];
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

__global__ void __launch_bounds__(128, 4) raytrace_rk4_cubic_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                                    const double* __restrict__ So, const double* __restrict__ Sn,
                                                                    PacketGrid g, RayParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[4] = {__ldcs(xk + i), __ldcs(xk + g.ld + i), __ldcs(xk + 2 * g.ld + i), __ldcs(xk + 3 * g.ld + i)};   // streaming: read once
    const double sg = sign[i];
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k[4], acc[4], y[4];
        ray_rhs_cubic(s, sg, (t - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_cubic(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_cubic(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
        const double a4 = it == p.nsub - 1 ? 1.0 : (t + h - p.t0) * inv_span;
        ray_rhs_cubic(y, sg, a4, So, Sn, g, p, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
    }
    __stcs(xk + i, s[0]);
    __stcs(xk + g.ld + i, s[1]);
    __stcs(xk + 2 * g.ld + i, s[2]);
    __stcs(xk + 3 * g.ld + i, s[3]);
}

__global__ void __launch_bounds__(128) sample_cubic_kernel(const double* __restrict__ xk, const unsigned* __restrict__ idx, long long n,
                                                           const double* __restrict__ S, PacketGrid g, double* __restrict__ U,
                                                           double* __restrict__ Gd, long long ldo) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int i0, i1, j0, j1;
    double a, b, Sv[5];
    cell(xk[i], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(xk[g.ld + i], g.y0, g.inv_dy, g.ny, j0, j1, b);
    sample_hermite5(S, g, i0, i1, j0, j1, a, b, Sv);
    const long long o = idx ? (long long)idx[i] : i;
    U[o] = Sv[0];
    U[ldo + o] = Sv[1];
    if (Gd) {
        Gd[o] = Sv[2];
        Gd[ldo + o] = Sv[3];
        Gd[2 * ldo + o] = Sv[4];
        Gd[3 * ldo + o] = -Sv[2];
    }
}

// ---------------------------------------------------------------- generic sampler / integrator combinations
// INTERP: 0 bilinear, 1 Hermite bicubic, 2 quadratic B-spline (coefficients prefiltered in the snapshot, raytracing/Raytracing.jl:161-170)
__device__ __forceinline__ void sample_bspline2_5(const double* __restrict__ S, const PacketGrid& g, double x, double y, double (&out)[5]) {
    const double sx = (x - g.x0) * g.inv_dx, sy = (y - g.y0) * g.inv_dy;
    const double rx = floor(sx + 0.5), ry = floor(sy + 0.5);
    const double dx = sx - rx, dy = sy - ry;
    const int ic = (int)((long long)rx & (long long)(g.nx - 1)), jc = (int)((long long)ry & (long long)(g.ny - 1));
    const double wx[3] = {0.5 * (dx - 0.5) * (dx - 0.5), 0.75 - dx * dx, 0.5 * (dx + 0.5) * (dx + 0.5)};
    const double wy[3] = {0.5 * (dy - 0.5) * (dy - 0.5), 0.75 - dy * dy, 0.5 * (dy + 0.5) * (dy + 0.5)};
#pragma unroll
    for (int c = 0; c < 5; ++c) out[c] = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int i = (ic + a - 1) & (g.nx - 1);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int j = (jc + b - 1) & (g.ny - 1);
            const double2* q = reinterpret_cast<const double2*>(S + ((long long)j * g.nx + i) * SNAP_STRIDE);
            const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const double w = wx[a] * wy[b];
            out[0] += w * q0.x; out[1] += w * q0.y; out[2] += w * q1.x; out[3] += w * q1.y; out[4] += w * q2.x;
        }
    }
}
// cubic B-spline (raytracing/Raytracing.jl:152-159): nodes floor(s)-1 .. floor(s)+2
__device__ __forceinline__ void sample_bspline3_5(const double* __restrict__ S, const PacketGrid& g, double x, double y, double (&out)[5]) {
    const double sx = (x - g.x0) * g.inv_dx, sy = (y - g.y0) * g.inv_dy;
    const double fx = floor(sx), fy = floor(sy);
    const double dx = sx - fx, dy = sy - fy;
    const int ic = (int)((long long)fx & (long long)(g.nx - 1)), jc = (int)((long long)fy & (long long)(g.ny - 1));
    const double s6 = 1.0 / 6.0;
    const double wx[4] = {(1 - dx) * (1 - dx) * (1 - dx) * s6, (3 * dx * dx * dx - 6 * dx * dx + 4) * s6,
                          (-3 * dx * dx * dx + 3 * dx * dx + 3 * dx + 1) * s6, dx * dx * dx * s6};
    const double wy[4] = {(1 - dy) * (1 - dy) * (1 - dy) * s6, (3 * dy * dy * dy - 6 * dy * dy + 4) * s6,
                          (-3 * dy * dy * dy + 3 * dy * dy + 3 * dy + 1) * s6, dy * dy * dy * s6};
#pragma unroll
    for (int c = 0; c < 5; ++c) out[c] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int j = (jc + b - 1) & (g.ny - 1);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = (ic + a - 1) & (g.nx - 1);
            const double2* q = reinterpret_cast<const double2*>(S + ((long long)j * g.nx + i) * SNAP_STRIDE);
            const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const double w = wx[a] * wy[b];
            out[0] += w * q0.x; out[1] += w * q0.y; out[2] += w * q1.x; out[3] += w * q1.y; out[4] += w * q2.x;
        }
    }
}
// Type-2 NUFFT evaluation (the intent of raytracing/NUFFTRaytracing.jl:68-84, nufft2d2 of every spectral field at the packet
// positions): the snapshot grid is the 2x oversampled grid of the deconvolved spectrum (the snapshot builder divides psih by the
// kernel's Fourier transform), and the value at (x, y) is the sum over nw x nw nodes weighted by the separable
// "exponential of semicircle" kernel phi(z) = exp(beta (sqrt(1 - z^2) - 1)), |z| <= 1 (FINUFFT's kernel; error ~ 10^(1 - nw)).
constexpr int NUFFT_MAXW = 16;
__device__ __forceinline__ void sample_nufft_5(const double* __restrict__ S, const PacketGrid& g, double x, double y, double (&out)[5]) {
    const double sx = (x - g.x0) * g.inv_dx, sy = (y - g.y0) * g.inv_dy, half = 0.5 * g.nw, ih = 1.0 / half;
    const double fx = ceil(sx - half), fy = ceil(sy - half);
    double wx[NUFFT_MAXW], wy[NUFFT_MAXW];
    for (int a = 0; a < g.nw; ++a) {
        const double zx = (sx - (fx + a)) * ih, zy = (sy - (fy + a)) * ih;
        const double qx = 1.0 - zx * zx, qy = 1.0 - zy * zy;
        wx[a] = qx > 0.0 ? exp(g.nbeta * (sqrt(qx) - 1.0)) : 0.0;
        wy[a] = qy > 0.0 ? exp(g.nbeta * (sqrt(qy) - 1.0)) : 0.0;
    }
    const long long ix0 = (long long)fx, iy0 = (long long)fy;
#pragma unroll
    for (int c = 0; c < 5; ++c) out[c] = 0.0;
    for (int b = 0; b < g.nw; ++b) {
        const long long row = (long long)(int)((iy0 + b) & (long long)(g.ny - 1)) * g.nx;
        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int a = 0; a < g.nw; ++a) {
            const int i = (int)((ix0 + a) & (long long)(g.nx - 1));
            const double2* q = reinterpret_cast<const double2*>(S + (row + i) * SNAP_STRIDE);
            const double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            r[0] += wx[a] * q0.x; r[1] += wx[a] * q0.y; r[2] += wx[a] * q1.x; r[3] += wx[a] * q1.y; r[4] += wx[a] * q2.x;
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) out[c] += wy[b] * r[c];
    }
}
template <int INTERP>
__device__ __forceinline__ void sample_level5(const double* __restrict__ S, const PacketGrid& g, double x, double y, double (&out)[5]) {
    if (INTERP == 5) {
        sample_nufft_5(S, g, x, y, out);
    } else if (INTERP == 4) {
        sample_bspline3_5(S, g, x, y, out);
    } else if (INTERP == 2) {
        sample_bspline2_5(S, g, x, y, out);
    } else {
        int i0, i1, j0, j1;
        double a, b;
        cell(x, g.x0, g.inv_dx, g.nx, i0, i1, a);
        cell(y, g.y0, g.inv_dy, g.ny, j0, j1, b);
        if (INTERP == 1) sample_hermite5(S, g, i0, i1, j0, j1, a, b, out);
        else bilinear5(S, g, i0, i1, j0, j1, a, b, out);
    }
}
template <int INTERP>
__device__ __forceinline__ void ray_rhs_generic(const double (&s)[4], double sign, double alpha, const double* __restrict__ So,
                                                const double* __restrict__ Sn, const PacketGrid& g, const RayParams& p, double (&d)[4]) {
    const double wo = p.lerp == 0 ? 1.0 - alpha : alpha, wn = p.lerp == 0 ? alpha : 1.0 - alpha;
    double o[5], nw[5], W[5];
    sample_level5<INTERP>(So, g, s[0], s[1], o);
    sample_level5<INTERP>(Sn, g, s[0], s[1], nw);
#pragma unroll
    for (int c = 0; c < 5; ++c) W[c] = wo * o[c] + wn * nw[c];
    const double k = s[2], l = s[3];
    const double cg = p.Cg * p.Cg * sign * rsqrt(p.f * p.f + p.Cg * p.Cg * (k * k + l * l));
    d[0] = W[0] + cg * k;
    d[1] = W[1] + cg * l;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}
// INTEG: 0 classical RK4, 1 implicit midpoint (raytracing/Raytracing.jl:106-109) by 12 fixed-point sweeps on the midpoint
template <int INTERP, int INTEG>
__global__ void __launch_bounds__(128, 3) raytrace_generic_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                                  const double* __restrict__ So, const double* __restrict__ Sn,
                                                                  PacketGrid g, RayParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[4] = {__ldcs(xk + i), __ldcs(xk + g.ld + i), __ldcs(xk + 2 * g.ld + i), __ldcs(xk + 3 * g.ld + i)};   // streaming: read once
    const double sg = sign[i];
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k[4], y[4];
        if (INTEG == 0) {
            double acc[4];
            ray_rhs_generic<INTERP>(s, sg, (t - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
            ray_rhs_generic<INTERP>(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
            ray_rhs_generic<INTERP>(y, sg, (t + 0.5 * h - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
            ray_rhs_generic<INTERP>(y, sg, (t + h - p.t0) * inv_span, So, Sn, g, p, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
        } else {
            const double alpha = (p.t0 + (it + 0.5) * h - p.t0) * inv_span;
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = s[c];
            for (int sweep = 0; sweep < 12; ++sweep) {
                ray_rhs_generic<INTERP>(y, sg, alpha, So, Sn, g, p, k);
#pragma unroll
                for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k[c];
            }
            ray_rhs_generic<INTERP>(y, sg, alpha, So, Sn, g, p, k);
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += h * k[c];
        }
    }
    __stcs(xk + i, s[0]);
    __stcs(xk + g.ld + i, s[1]);
    __stcs(xk + 2 * g.ld + i, s[2]);
    __stcs(xk + 3 * g.ld + i, s[3]);
}
template <int INTERP>
__global__ void __launch_bounds__(128) sample_generic_kernel(const double* __restrict__ xk, const unsigned* __restrict__ idx, long long n,
                                                             const double* __restrict__ S, PacketGrid g, double* __restrict__ U,
                                                             double* __restrict__ Gd, long long ldo) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double Sv[5];
    sample_level5<INTERP>(S, g, xk[i], xk[g.ld + i], Sv);
    const long long o = idx ? (long long)idx[i] : i;
    U[o] = Sv[0];
    U[ldo + o] = Sv[1];
    if (Gd) {
        Gd[o] = Sv[2];
        Gd[ldo + o] = Sv[3];
        Gd[2 * ldo + o] = Sv[4];
        Gd[3 * ldo + o] = -Sv[2];
    }
}

// ---------------------------------------------------------------- fp32 packet mode (north star: "optional fp32 packet mode")
// The reference's GPU tracer samples Float32 textures (raytracing/GPURaytracing.jl:118-127).  Here the node records are
// eight floats, the stencil (40 floats) and the whole right-hand side are evaluated in fp32, and the packet state, the cell
// coordinate and the RK4 combination stay in fp64, so positions do not lose increments to fp32 rounding.
struct StencilF {
    float4 c[2][4][2];
    int ci, cj;
};
__device__ __forceinline__ void ray_rhs_f32(const double (&s)[4], float sign, float alpha, const float4* __restrict__ So,
                                            const float4* __restrict__ Sn, const PacketGrid& g, float f2, float cg2, int lerp,
                                            StencilF& st, double (&d)[4]) {
    int i0, i1, j0, j1;
    double ad, bd;
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, ad);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, bd);
    if (i0 != st.ci || j0 != st.cj) {
        st.ci = i0;
        st.cj = j0;
        const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
#pragma unroll
        for (int lev = 0; lev < 2; ++lev) {
            const float4* S = lev == 0 ? So : Sn;
#pragma unroll
            for (int cr = 0; cr < 4; ++cr) {
                st.c[lev][cr][0] = __ldg(S + 2 * pt[cr]);
                st.c[lev][cr][1] = __ldg(S + 2 * pt[cr] + 1);
            }
        }
    }
    const float a = (float)ad, b = (float)bd, a1 = 1.f - a, b1 = 1.f - b;
    const float wo = lerp == 0 ? 1.f - alpha : alpha, wn = lerp == 0 ? alpha : 1.f - alpha;
    const float wb[4] = {a1 * b1, a * b1, a1 * b, a * b};
    float W[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int lev = 0; lev < 2; ++lev) {
        const float wl = lev == 0 ? wo : wn;
        if (wl == 0.f) continue;
#pragma unroll
        for (int cr = 0; cr < 4; ++cr) {
            const float w = wl * wb[cr];
            const float4 q0 = st.c[lev][cr][0];
            W[0] = fmaf(w, q0.x, W[0]); W[1] = fmaf(w, q0.y, W[1]); W[2] = fmaf(w, q0.z, W[2]); W[3] = fmaf(w, q0.w, W[3]);
            W[4] = fmaf(w, st.c[lev][cr][1].x, W[4]);
        }
    }
    const float k = (float)s[2], l = (float)s[3];
    const float cg = cg2 * sign * rsqrtf(fmaf(cg2, fmaf(k, k, l * l), f2));
    d[0] = (double)fmaf(cg, k, W[0]);
    d[1] = (double)fmaf(cg, l, W[1]);
    d[2] = (double)(-(W[2] * k + W[4] * l));
    d[3] = (double)(-(W[3] * k - W[2] * l));
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) raytrace_rk4_f32_kernel(double* __restrict__ xk, const double* __restrict__ sign, long long n,
                                                                  const float4* __restrict__ So, const float4* __restrict__ Sn, PacketGrid g,
                                                                  RayParams p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[4] = {__ldcs(xk + i), __ldcs(xk + g.ld + i), __ldcs(xk + 2 * g.ld + i), __ldcs(xk + 3 * g.ld + i)};   // streaming: read once
    const float sg = (float)sign[i], f2 = (float)(p.f * p.f), cg2 = (float)(p.Cg * p.Cg);
    const double h = (p.t1 - p.t0) / p.nsub, inv_span = 1.0 / (p.t1 - p.t0);
    StencilF st;
    st.ci = -1;
    st.cj = -1;
    for (int it = 0; it < p.nsub; ++it) {
        const double t = p.t0 + it * h;
        double k[4], acc[4], y[4];
        ray_rhs_f32(s, sg, (float)((t - p.t0) * inv_span), So, Sn, g, f2, cg2, p.lerp, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_f32(y, sg, (float)((t + 0.5 * h - p.t0) * inv_span), So, Sn, g, f2, cg2, p.lerp, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
        ray_rhs_f32(y, sg, (float)((t + 0.5 * h - p.t0) * inv_span), So, Sn, g, f2, cg2, p.lerp, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
        const float a4 = it == p.nsub - 1 ? 1.f : (float)((t + h - p.t0) * inv_span);
        ray_rhs_f32(y, sg, a4, So, Sn, g, f2, cg2, p.lerp, st, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
    }
    __stcs(xk + i, s[0]);
    __stcs(xk + g.ld + i, s[1]);
    __stcs(xk + 2 * g.ld + i, s[2]);
    __stcs(xk + 3 * g.ld + i, s[3]);
}
__global__ void __launch_bounds__(128) sample_f32_kernel(const double* __restrict__ xk, const unsigned* __restrict__ idx, long long n,
                                                         const float4* __restrict__ S, PacketGrid g, double* __restrict__ U,
                                                         double* __restrict__ Gd, long long ldo) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int i0, i1, j0, j1;
    double ad, bd;
    cell(xk[i], g.x0, g.inv_dx, g.nx, i0, i1, ad);
    cell(xk[g.ld + i], g.y0, g.inv_dy, g.ny, j0, j1, bd);
    const float a = (float)ad, b = (float)bd;
    const float wb[4] = {(1.f - a) * (1.f - b), a * (1.f - b), (1.f - a) * b, a * b};
    const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
    float W[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cr = 0; cr < 4; ++cr) {
        const float4 q0 = __ldg(S + 2 * pt[cr]), q1 = __ldg(S + 2 * pt[cr] + 1);
        W[0] = fmaf(wb[cr], q0.x, W[0]); W[1] = fmaf(wb[cr], q0.y, W[1]); W[2] = fmaf(wb[cr], q0.z, W[2]); W[3] = fmaf(wb[cr], q0.w, W[3]);
        W[4] = fmaf(wb[cr], q1.x, W[4]);
    }
    const long long o = idx ? (long long)idx[i] : i;
    U[o] = W[0];
    U[ldo + o] = W[1];
    if (Gd) {
        Gd[o] = W[2];
        Gd[ldo + o] = W[3];
        Gd[2 * ldo + o] = W[4];
        Gd[3 * ldo + o] = -W[2];
    }
}
// fp32 node records -> planar (nx, ny, 5) doubles for swrt_flow_get_snapshot
__global__ void snapf_to_planar_kernel(const float* __restrict__ S, long long npts, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= npts) return;
#pragma unroll
    for (int c = 0; c < 5; ++c) out[c * npts + i] = (double)S[i * SNAPF_STRIDE + c];
}

// ---------------------------------------------------------------- fp32 packet mode, three staged levels (nsub == 1)
// The three-level tile kernel for Float32 node records (8 floats = 32 B per node: u, v, ux, uy | vx, 0, 0, 0): the patches are
// half the bytes of the fp64 ones, the stencil of a stage is 20 floats, the right-hand side is evaluated in fp32 like
// ray_rhs_f32; packet state, cell coordinate and the RK4 combination stay fp64.
constexpr int PATCHF_ROW = 188;                          // floats per patch row: 47 sixteen-byte chunks (>= 23 nodes x 2, odd bank-group stride)
static_assert(PATCHF_ROW >= PATCH * SNAPF_STRIDE && PATCHF_ROW % 4 == 0 && PATCHF_ROW <= 256, "fp32 patch row");
constexpr int PATCHF_BYTES = (PATCH * PATCHF_ROW * 4 + 127) / 128 * 128;
constexpr int TILE3F_SMEM_BYTES = 3 * PATCHF_BYTES + TILE3_STAGE_BYTES;
struct Stencil1F {
    float4 c[4];   // u, v, ux, uy at the corners 00, 10, 01, 11
    float vx[4];
};
struct TilePatch3F {
    const float* lev[3];
    int pi, pj;
    bool staged;
};
__device__ __forceinline__ float4 mean4(float4 a, float4 b) { return make_float4(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z), 0.5f * (a.w + b.w)); }
template <int LEV>
__device__ __forceinline__ void fill_stencil1f(Stencil1F& st, int i0, int j0, const float4* __restrict__ F1, const float4* __restrict__ F4,
                                               const PacketGrid& g, const TilePatch3F& tp) {
    const unsigned ri = (unsigned)((i0 - tp.pi) & (g.nx - 1)), rj = (unsigned)((j0 - tp.pj) & (g.ny - 1));
    if (tp.staged && ri < (unsigned)(PATCH - 1) && rj < (unsigned)(PATCH - 1)) {
        const float4* q = reinterpret_cast<const float4*>(tp.lev[LEV] + rj * PATCHF_ROW + ri * SNAPF_STRIDE);
        st.c[0] = q[0];
        st.c[1] = q[2];
        st.c[2] = q[PATCHF_ROW / 4];
        st.c[3] = q[PATCHF_ROW / 4 + 2];
        st.vx[0] = reinterpret_cast<const float*>(q + 1)[0];
        st.vx[1] = reinterpret_cast<const float*>(q + 3)[0];
        st.vx[2] = reinterpret_cast<const float*>(q + PATCHF_ROW / 4 + 1)[0];
        st.vx[3] = reinterpret_cast<const float*>(q + PATCHF_ROW / 4 + 3)[0];
    } else {
        const int i1 = (i0 + 1) & (g.nx - 1);
        int j1 = (j0 + 1) & (g.ny - 1);
        stencil_rows(g, j0, j1);
        const long long pt[4] = {(long long)j0 * g.nx + i0, (long long)j0 * g.nx + i1, (long long)j1 * g.nx + i0, (long long)j1 * g.nx + i1};
#pragma unroll
        for (int cr = 0; cr < 4; ++cr) {
            const float4* qa = (LEV == 2 ? F4 : F1) + 2 * pt[cr];
            st.c[cr] = __ldg(qa);
            st.vx[cr] = __ldg(reinterpret_cast<const float*>(qa + 1));
            if (LEV == 1) {
                const float4* qb = F4 + 2 * pt[cr];
                st.c[cr] = mean4(st.c[cr], __ldg(qb));
                st.vx[cr] = 0.5f * (st.vx[cr] + __ldg(reinterpret_cast<const float*>(qb + 1)));
            }
        }
    }
}
__device__ __forceinline__ void ray_rhs1f(const Stencil1F& st, double ad, double bd, double kd, double ld, float sign, float f2, float cg2,
                                          double (&d)[4]) {
    const float a = (float)ad, b = (float)bd, a1 = 1.f - a, b1 = 1.f - b;
    const float w[4] = {a1 * b1, a * b1, a1 * b, a * b};
    float W[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cr = 0; cr < 4; ++cr) {
        W[0] = fmaf(w[cr], st.c[cr].x, W[0]); W[1] = fmaf(w[cr], st.c[cr].y, W[1]); W[2] = fmaf(w[cr], st.c[cr].z, W[2]);
        W[3] = fmaf(w[cr], st.c[cr].w, W[3]); W[4] = fmaf(w[cr], st.vx[cr], W[4]);
    }
    const float k = (float)kd, l = (float)ld;
    const float cg = cg2 * sign * rsqrtf(fmaf(cg2, fmaf(k, k, l * l), f2));
    d[0] = (double)fmaf(cg, k, W[0]);
    d[1] = (double)fmaf(cg, l, W[1]);
    d[2] = (double)(-(W[2] * k + W[4] * l));
    d[3] = (double)(-(W[3] * k - W[2] * l));
}
__device__ __forceinline__ void rk4_three_level_f32(double (&s)[4], float sg, double h, const float4* __restrict__ F1, const float4* __restrict__ F4,
                                                    const PacketGrid& g, float f2, float cg2, const TilePatch3F& tp) {
    Stencil1F st;
    int i0, i1, j0, j1, ci, cj;
    double a, b, k[4], acc[4], y[4];
    cell(s[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(s[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1f<0>(st, i0, j0, F1, F4, g, tp);
    ray_rhs1f(st, a, b, s[2], s[3], sg, f2, cg2, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] = k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1f<1>(st, i0, j0, F1, F4, g, tp);
    ci = i0;
    cj = j0;
    ray_rhs1f(st, a, b, y[2], y[3], sg, f2, cg2, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + 0.5 * h * k[c]; }
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    if (i0 != ci || j0 != cj) fill_stencil1f<1>(st, i0, j0, F1, F4, g, tp);
    ray_rhs1f(st, a, b, y[2], y[3], sg, f2, cg2, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[c] += 2.0 * k[c]; y[c] = s[c] + h * k[c]; }
    cell(y[0], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y[1], g.y0, g.inv_dy, g.ny, j0, j1, b);
    fill_stencil1f<2>(st, i0, j0, F1, F4, g, tp);
    ray_rhs1f(st, a, b, y[2], y[3], sg, f2, cg2, k);
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] += (h / 6.0) * (acc[c] + k[c]);
}
// F1 / map1: the level whose lerp weight is 1 at t0 (resolved by the caller), viewed as [ny][nx * 8] floats
template <int MINB>
__global__ void __launch_bounds__(TILE3_THREADS, MINB)
    raytrace_rk4_tile3_f32_kernel(double* __restrict__ xk, const double* __restrict__ sign, const float4* __restrict__ F1, const float4* __restrict__ F4,
                                  const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map4,
                                  const unsigned* __restrict__ tile_end, PacketGrid g, RayParams p) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int tiles_x = g.nx >> TILE_SHIFT, tiles_y = g.ny >> TILE_SHIFT;
    const int tjl = blockIdx.x / tiles_x, ti = blockIdx.x - tjl * tiles_x;
    const int tj = (g.tile_row0 + tjl) % tiles_y;
    const int tile = tj * tiles_x + ti;
    const long long key0 = (long long)tile << (2 * TILE_SHIFT);
    const long long start = tile == 0 ? 0 : (long long)__ldg(&tile_end[key0 - 1]), end = (long long)__ldg(&tile_end[key0 + (TILE * TILE - 1)]);
    float* const patch1 = reinterpret_cast<float*>(tile_smem);
    float* const patchm = reinterpret_cast<float*>(tile_smem + PATCHF_BYTES);
    float* const patch4 = reinterpret_cast<float*>(tile_smem + 2 * PATCHF_BYTES);
    TilePatch3F tp;
    tp.lev[0] = patch1;
    tp.lev[1] = patchm;
    tp.lev[2] = patch4;
    tp.pi = ti * TILE - TILE_MARGIN;
    tp.pj = tj * TILE - TILE_MARGIN;
    int prow = tp.pj;
    bool rows_ok = true;
    if (g.band) {
        prow = (tp.pj - g.jb) & (g.ny - 1);
        if (prow >= g.ny / 2) prow -= g.ny;
        rows_ok = prow >= 0 && prow + PATCH <= g.jrows;
    }
    const bool by_tma = rows_ok && tp.pi >= 0 && tp.pi + PATCH <= g.nx && prow >= 0 && prow + PATCH <= (g.band ? g.jrows : g.ny);
    tp.staged = rows_ok;
    if (by_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, 2u * PATCH * PATCHF_ROW * 4);
            tma_load_2d(patch1, &map1, tp.pi * SNAPF_STRIDE, prow, &bar);
            tma_load_2d(patch4, &map4, tp.pi * SNAPF_STRIDE, prow, &bar);
        }
    }
    if (start >= end) {
        if (by_tma) mbar_wait(&bar, 0);
        return;
    }
    const double h = p.t1 - p.t0;
    const float f2 = (float)(p.f * p.f), cg2 = (float)(p.Cg * p.Cg);
    double* stage = reinterpret_cast<double*>(tile_smem + 3 * PATCHF_BYTES);     // [2][5][TILE3_THREADS]
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    auto prefetch = [&](long long i, int buf) {
        if (i < end) {
            double* d = stage + buf * 5 * TILE3_THREADS + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async8_stream(d + c * TILE3_THREADS, xk + c * g.ld + i, pol);
            cp_async8_stream(d + 4 * TILE3_THREADS, sign + i, pol);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    prefetch(start + threadIdx.x, 0);
    if (by_tma) {
        mbar_wait(&bar, 0);
        float4* m = reinterpret_cast<float4*>(patchm);
        const float4 *a = reinterpret_cast<const float4*>(patch1), *b = reinterpret_cast<const float4*>(patch4);
        for (int c = threadIdx.x; c < PATCH * PATCHF_ROW / 4; c += TILE3_THREADS) m[c] = mean4(a[c], b[c]);
        __syncthreads();
    } else if (rows_ok) {
        for (int e = threadIdx.x; e < PATCH * PATCH * 2; e += TILE3_THREADS) {
            const int node = e >> 1, q = e & 1, r = node / PATCH, c = node - r * PATCH;
            const int col = (tp.pi + c) & (g.nx - 1), row = g.band ? prow + r : ((tp.pj + r) & (g.ny - 1));
            const long long src = ((long long)row * g.nx + col) * 2 + q;
            const float4 va = __ldg(F1 + src), vb = __ldg(F4 + src);
            const int dst = r * PATCHF_ROW + c * SNAPF_STRIDE + 4 * q;
            *reinterpret_cast<float4*>(patch1 + dst) = va;
            *reinterpret_cast<float4*>(patch4 + dst) = vb;
            *reinterpret_cast<float4*>(patchm + dst) = mean4(va, vb);
        }
        __syncthreads();
    }
    int buf = 0;
    for (long long i = start + threadIdx.x; i < end; i += TILE3_THREADS, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        const double* sv = stage + buf * 5 * TILE3_THREADS + threadIdx.x;
        double s[4] = {sv[0], sv[TILE3_THREADS], sv[2 * TILE3_THREADS], sv[3 * TILE3_THREADS]};
        const float sg = (float)sv[4 * TILE3_THREADS];
        prefetch(i + TILE3_THREADS, buf ^ 1);
        rk4_three_level_f32(s, sg, h, F1, F4, g, f2, cg2, tp);
        __stcs(xk + i, s[0]);
        __stcs(xk + g.ld + i, s[1]);
        __stcs(xk + 2 * g.ld + i, s[2]);
        __stcs(xk + 3 * g.ld + i, s[3]);
    }
}

// interpolate_velocity!/gradients!: U (N,2), Gd (N,4) = ux, uy, vx, vy, written at the packets' ORIGINAL rows
__global__ void __launch_bounds__(128) sample_kernel(const double* __restrict__ xk, const unsigned* __restrict__ idx, long long n,
                                                     const double* __restrict__ S, PacketGrid g, double* __restrict__ U,
                                                     double* __restrict__ Gd, long long ldo) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int i0, i1, j0, j1;
    double a, b, Sv[5];
    cell(xk[i], g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(xk[g.ld + i], g.y0, g.inv_dy, g.ny, j0, j1, b);
    stencil_rows(g, j0, j1);
    bilinear5(S, g, i0, i1, j0, j1, a, b, Sv);
    const long long o = idx ? (long long)idx[i] : i;
    U[o] = Sv[0];
    U[ldo + o] = Sv[1];
    if (Gd) {
        Gd[o] = Sv[2];
        Gd[ldo + o] = Sv[3];
        Gd[2 * ldo + o] = Sv[4];
        Gd[3 * ldo + o] = -Sv[2];
    }
}

__global__ void kcutoff_kernel(double* __restrict__ xk, long long n, long long ld, double kc2, double k0, unsigned long long* count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double k = xk[2 * ld + i], l = xk[3 * ld + i];
    if (k * k + l * l >= kc2) {
        xk[2 * ld + i] = k0;
        xk[3 * ld + i] = 0.0;
        atomicAdd(count, 1ULL);
    }
}

// generate_initial_wavepackets (raytracing/RaytracingDriver.jl:27-47); p = global 1-based packet index
__global__ void generate_packets_kernel(double* __restrict__ xk, double* __restrict__ sign, unsigned* __restrict__ idx, long long n,
                                        long long ld, long long first, long long sqrtN, double L, double k0) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long p0 = first + i;  // 0-based global
    const long long ntot = sqrtN * sqrtN;
    const double offset = L / (double)sqrtN / 2.0;
    const long long ix = p0 % sqrtN + 1, iy = p0 / sqrtN + 1;
    xk[i] = (double)ix * L / (double)sqrtN - L / 2.0 - offset;
    xk[ld + i] = (double)iy * L / (double)sqrtN - L / 2.0 - offset;
    const double phase = 2.0 * 3.141592653589793 * (double)(p0 + 1) / (double)ntot;
    xk[2 * ld + i] = k0 * cos(phase);
    xk[3 * ld + i] = k0 * sin(phase);
    sign[i] = (p0 % 2 == 0) ? -1.0 : 1.0;
    idx[i] = (unsigned)i;
}

__global__ void iota_kernel(unsigned* __restrict__ idx, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (unsigned)i;
}

// ---------------------------------------------------------------- counting sort by tiled cell key
__device__ __forceinline__ unsigned cell_key(double x, double y, const PacketGrid& g) {
    int i0, i1, j0, j1;
    double a;
    cell(x, g.x0, g.inv_dx, g.nx, i0, i1, a);
    cell(y, g.y0, g.inv_dy, g.ny, j0, j1, a);
    return (unsigned)((((j0 >> TILE_SHIFT) * (g.nx >> TILE_SHIFT) + (i0 >> TILE_SHIFT)) << (2 * TILE_SHIFT)) + ((j0 & (TILE - 1)) << TILE_SHIFT) + (i0 & (TILE - 1)));
}

__global__ void sort_hist_kernel(const double* __restrict__ xk, long long n, PacketGrid g, unsigned* __restrict__ keys,
                                 unsigned* __restrict__ hist, unsigned long long* __restrict__ violations) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = cell_key(xk[i], xk[g.ld + i], g);
    if (g.band) {   // a packet that left band + halo since the last migration sampled clamped rows: report it
        int i0, i1, j0, j1;
        double a;
        cell(xk[g.ld + i], g.y0, g.inv_dy, g.ny, j0, j1, a);
        (void)i0; (void)i1;
        if (((j0 - g.jb) & (g.ny - 1)) > g.jrows - 2) atomicAdd(violations, 1ULL);
    }
    keys[i] = k;
    atomicAdd(&hist[k], 1u);
}

// exclusive scan of `hist` (nb elements) in three launches: per-block scan, scan of block sums (one block), add
constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_kernel(unsigned* __restrict__ a, long long nb, unsigned* __restrict__ sums) {
    // warp-shuffle scan: inclusive scan inside every warp (5 shuffle steps), the 32 warp totals scanned by the first warp,
    // two barriers per block (the shared-memory Hillis-Steele version it replaces needed twenty)
    __shared__ unsigned wsum[SCAN_BLOCK / 32];
    const long long i = blockIdx.x * (long long)SCAN_BLOCK + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned v = i < nb ? a[i] : 0u;
    unsigned x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += t;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned w = wsum[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        wsum[lane] = w;                                   // inclusive totals of warps 0..lane
    }
    __syncthreads();
    const unsigned base = wid ? wsum[wid - 1] : 0u;
    if (i < nb) a[i] = base + x - v;                      // exclusive
    if (threadIdx.x == SCAN_BLOCK - 1 && sums) sums[blockIdx.x] = base + x;
}
__global__ void scan_add_kernel(unsigned* __restrict__ a, long long nb, const unsigned* __restrict__ sums) {
    const long long i = blockIdx.x * (long long)SCAN_BLOCK + threadIdx.x;
    if (i < nb) a[i] += sums[blockIdx.x];
}

// scatter packet i to its sorted slot and move its state (the order inside a cell is arbitrary)
__global__ void sort_scatter_kernel(const double* __restrict__ xk, const double* __restrict__ sign, const unsigned* __restrict__ idx,
                                    const unsigned* __restrict__ keys, unsigned* __restrict__ offsets, long long n, long long ld,
                                    double* __restrict__ xk2, double* __restrict__ sign2, unsigned* __restrict__ idx2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned pos = atomicAdd(&offsets[keys[i]], 1u);
#pragma unroll
    for (int c = 0; c < 4; ++c) xk2[c * ld + pos] = xk[c * ld + i];
    sign2[pos] = sign[i];
    idx2[pos] = idx[i];
}

// host-visible order: out[c][idx[i]] = xk[c][i]
__global__ void unpermute_kernel(const double* __restrict__ xk, const unsigned* __restrict__ idx, long long n, long long ld, int ncols,
                                 double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long o = idx ? (long long)idx[i] : i;
    for (int c = 0; c < ncols; ++c) out[c * ld + o] = xk[c * ld + i];
}

// snapshot half <-> planar (nx, ny, 5) host layout staging
__global__ void snap_to_planar_kernel(const double* __restrict__ S, long long npts, int nc, int stride, double* __restrict__ planar) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= npts) return;
    for (int c = 0; c < nc; ++c) planar[c * npts + i] = S[i * stride + c];
}
__global__ void planar_to_snap_kernel(const double* __restrict__ planar, long long npts, int nc, int stride, double* __restrict__ S) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= npts) return;
    for (int c = 0; c < stride; ++c) S[i * stride + c] = c < nc ? planar[c * npts + i] : 0.0;
}

}  // namespace swrt
