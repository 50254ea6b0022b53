// Model functors for the generic passes + the fused IFMAB3 update + per-size launchers.
//
// Reference arithmetic being reproduced (file:line in the reference checkout):
//   rsw/RotatingShallowWater.jl:140-230   calcN!  (RSW)      rsw/ModifiedShallowWater.jl:209-226 (extra term)
//   utils/IFMAB3.jl:129-169               IFMAB3update! / stepforward!
//   rsw/RSWRaytracingDriver.jl:63-67      get_streamfunction!
//   raytracing/RaytracingDriver.jl:132-154 get_velocity_info
#pragma once
#include <cstdlib>

#include "passes.cuh"
#include "snapshot_layout.cuh"

namespace swrt {

enum { MODEL_RSW = 0, MODEL_RSW_MODIFIED = 1, MODEL_RSW_LINDBORG = 2, MODEL_RSW_QUADHEIGHT = 3, MODEL_SWQG = 4, MODEL_TWOLAYERQG = 5, MODEL_THOMASYAMADA = 6, MODEL_MULTILAYERQG2 = 7 };

// per-model sizes: state variables, y-transformed intermediates (stage A jobs), x-transformed products (stage B outputs)
__host__ __device__ constexpr bool model_two_layer(int m) { return m == MODEL_TWOLAYERQG || m == MODEL_MULTILAYERQG2; }
__host__ __device__ constexpr int model_nvar(int m) { return m == MODEL_SWQG ? 1 : (model_two_layer(m) ? 2 : (m == MODEL_THOMASYAMADA ? 4 : 3)); }
__host__ __device__ constexpr int model_njobs_a(int m) { return m == MODEL_THOMASYAMADA ? 9 : m == MODEL_SWQG ? 3 : (model_two_layer(m) ? 6 : (m == MODEL_RSW_LINDBORG ? 8 : 5)); }
__host__ __device__ constexpr int model_njobs_b(int m) { return m == MODEL_THOMASYAMADA ? 9 : m == MODEL_SWQG ? 2 : (model_two_layer(m) ? 4 : (m == MODEL_RSW_LINDBORG ? 3 : ((m == MODEL_RSW_MODIFIED || m == MODEL_RSW_QUADHEIGHT) ? 5 : 4))); }

// ---------------------------------------------------------------- RSW family, stage A
// jobs: 0 uh, 1 vh, 2 etah, 3 i l uh, 4 i l vh     (ux, vx are derived in the x-pass as i k G)
struct RswLoaderA {
    const double2* sol;
    long long vs;
    __device__ __forceinline__ double2 operator()(int job, int, int, double, double lw, long long off) const {
        switch (job) {
            case 0: return sol[off];
            case 1: return sol[vs + off];
            case 2: return sol[2 * vs + off];
            case 3: { const double2 a = sol[off]; return make_double2(-lw * a.y, lw * a.x); }
            default: { const double2 a = sol[vs + off]; return make_double2(-lw * a.y, lw * a.x); }
        }
    }
};

// ---------------------------------------------------------------- RSW family, stage B
// p1 = u ux + v uy, p2 = u vx + v vy, p3 = u eta, p4 = v eta [, p5 = 1.5 - 0.5/(1+eta)^2]
// MODIFIED: 0 plain RSW, 1 Modified (G = 1.5 - 0.5/(1+eta)^2), 2 QuadHeight (third variable m, G = 1.5 - 0.5 m^2)
template <int N, int MODIFIED, bool SLAB = false>
struct RswXOp {
    static constexpr int NBUF = 2;
    const double2* G;  // [5][ny][kr_pad]
    double2* H;        // [4 or 5][ny][kr_pad]
    double sc;         // (1/(nx ny))^2 / 2
    double s1;         // 1/(nx ny)
    OutPeers peers;    // slab mode: destinations of the output column segments
    OutPeers gin;      // slab mode: sources of the input column segments
    int nj = 5;        // jobs the y-pass put into G (8 when the snapshot's three psi jobs ride along, team mode)
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        const auto Gu = row_in<SLAB>(L, G, gin, nj, 0, y), Gv = row_in<SLAB>(L, G, gin, nj, 1, y), Ge = row_in<SLAB>(L, G, gin, nj, 2, y),
                   Guy = row_in<SLAB>(L, G, gin, nj, 3, y), Gvy = row_in<SLAB>(L, G, gin, nj, 4, y);
        const RowPlain none{};
        constexpr int NH = MODIFIED ? 5 : 4;
        // Thread g owns x = g + m N/16 (m = 0..15) of the physical row.  Inverse transforms are loaded through shared
        // memory (Hermitian extension needs k and N-k) but deliver their result in registers (last FFT stage); products
        // are formed there and enter the forward transform's first stage directly.  Buffer 0 only parks u, v at the
        // thread's own x (no other thread reads them); buffer 1 is the FFT work space.
        double *ur = cx.re(0), *vr = cx.im(0);
        double2 v[16];
        double p1[16];
        cx.template load_pair<MUL_ONE, MUL_ONE>(1, Gu, Gv);
        cx.ifft_regs_out(1, v);                          // u + i v
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            ur[x] = v[m].x;
            vr[x] = v[m].y;
        }
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gu, Guy);
        cx.ifft_regs_out(1, v);                          // ux + i uy
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            p1[m] = sc * (ur[x] * v[m].x + vr[x] * v[m].y);
        }
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gv, Gvy);
        cx.ifft_regs_out(1, v);                          // vx + i vy
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            v[m] = make_double2(p1[m], sc * (ur[x] * v[m].x + vr[x] * v[m].y));
        }
        cx.fft_regs_in(v, 1);
        cx.template store_pair<MUL_ONE, MUL_ONE>(1, row_out<SLAB>(L, H, peers, NH, 0, y), row_out<SLAB>(L, H, peers, NH, 1, y));
        cx.template load_pair<MUL_ONE, MUL_ZERO>(1, Ge, none);
        cx.ifft_regs_out(1, v);                          // eta
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            const double e = v[m].x;
            if (MODIFIED == 1) {
                const double e1 = 1.0 + s1 * e;
                p1[m] = 0.5 * (1.5 - 0.5 / (e1 * e1));
            } else if (MODIFIED == 2) {
                const double mm = s1 * e;
                p1[m] = 0.5 * (1.5 - 0.5 * (mm * mm));
            }
            v[m] = make_double2(sc * (ur[x] * e), sc * (vr[x] * e));
        }
        cx.fft_regs_in(v, 1);
        cx.template store_pair<MUL_ONE, MUL_ONE>(1, row_out<SLAB>(L, H, peers, NH, 2, y), row_out<SLAB>(L, H, peers, NH, 3, y));
        if (MODIFIED) {
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = make_double2(p1[m], 0.0);
            cx.fft_regs_in(v, 1);
            cx.template store_pair<MUL_ONE, MUL_ZERO>(1, row_out<SLAB>(L, H, peers, NH, 4, y), none);
        }
    }
};

// ---------------------------------------------------------------- RSW family, stage C
// N_u = -P1 [- i c2 k P5], N_v = -P2 [- i c2 l P5], N_eta = -i k P3 - i l P4
struct RswCombiner {
    int modified;
    double c2;
    __device__ __forceinline__ int var_of(int slot) const { return slot == 0 ? 2 : slot - 1; }   // eta (two inputs) first
    __device__ __forceinline__ double2 init(int, double, double, long long) const { return make_double2(0.0, 0.0); }
    __device__ __forceinline__ int nin(int var) const { return var == 2 ? 2 : (modified ? 2 : 1); }
    __device__ __forceinline__ int src(int var, int i) const { return var == 2 ? 2 + i : (i == 0 ? var : 4); }
    __device__ __forceinline__ double2 apply(int var, int i, double2 v, double kw, double lw) const {
        if (var == 2) {
            const double w = i == 0 ? kw : lw;
            return make_double2(w * v.y, -w * v.x);          // -i w v
        }
        if (i == 0) return make_double2(-v.x, -v.y);
        const double w = c2 * (var == 0 ? kw : lw);
        return make_double2(w * v.y, -w * v.x);              // -i c2 w v
    }
};

// ---------------------------------------------------------------- Lindborg RSW (rsw/LinborgShallowWater.jl:144-241)
// advecting velocity = rotational part: urh = -l (k vh - l uh)/K^2, vrh = k (k vh - l uh)/K^2
// jobs: 0 urh, 1 vrh, 2 uh (ux = i k G), 3 i l uh, 4 vh (vx = i k G), 5 i l vh, 6 etah (eta_x = i k G), 7 i l etah
struct LindborgLoaderA {
    const double2* sol;
    long long vs;
    __device__ __forceinline__ double2 operator()(int job, int, int, double kw, double lw, long long off) const {
        if (job < 2) {
            const double2 u = sol[off], v = sol[vs + off];
            const double K2 = kw * kw + lw * lw, inv = K2 > 0.0 ? 1.0 / K2 : 0.0;
            const double rx = (kw * v.x - lw * u.x) * inv, ry = (kw * v.y - lw * u.y) * inv;
            return job == 0 ? make_double2(-lw * rx, -lw * ry) : make_double2(kw * rx, kw * ry);
        }
        const int var = (job - 2) >> 1;
        const double2 a = sol[var * vs + off];
        return (job & 1) ? make_double2(-lw * a.y, lw * a.x) : a;
    }
};

template <int N, bool SLAB = false>
struct LindborgXOp {
    static constexpr int NBUF = 2;
    const double2* G;  // [8][ny][kr_pad]
    double2* H;        // [3][ny][kr_pad]
    double sc;
    OutPeers peers{};  // slab mode: destinations of the output column segments
    OutPeers gin{};    // slab mode: sources of the input column segments
    int nj = 8;        // jobs the y-pass put into G (11 when the snapshot's three psi jobs ride along, team mode)
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        auto Gp = [&](int j) { return row_in<SLAB>(L, G, gin, nj, j, y); };
        auto Hp = [&](int j) { return row_out<SLAB>(L, H, peers, 3, j, y); };
        const RowPlain none{};
        double *ur = cx.re(0), *vr = cx.im(0);
        double2 v[16];
        double p1[16];
        cx.template load_pair<MUL_ONE, MUL_ONE>(1, Gp(0), Gp(1));
        cx.ifft_regs_out(1, v);                          // ur + i vr
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            ur[x] = v[m].x;
            vr[x] = v[m].y;
        }
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gp(2), Gp(3));
        cx.ifft_regs_out(1, v);                          // ux + i uy
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            p1[m] = sc * (ur[x] * v[m].x + vr[x] * v[m].y);
        }
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gp(4), Gp(5));
        cx.ifft_regs_out(1, v);                          // vx + i vy
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            v[m] = make_double2(p1[m], sc * (ur[x] * v[m].x + vr[x] * v[m].y));
        }
        cx.fft_regs_in(v, 1);
        cx.template store_pair<MUL_ONE, MUL_ONE>(1, Hp(0), Hp(1));
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gp(6), Gp(7));
        cx.ifft_regs_out(1, v);                          // eta_x + i eta_y
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            v[m] = make_double2(sc * (ur[x] * v[m].x + vr[x] * v[m].y), 0.0);
        }
        cx.fft_regs_in(v, 1);
        cx.template store_pair<MUL_ONE, MUL_ZERO>(1, Hp(2), none);
    }
};

struct NegateCombiner {  // N_var = -P_var
    __device__ __forceinline__ int var_of(int slot) const { return slot; }
    __device__ __forceinline__ double2 init(int, double, double, long long) const { return make_double2(0.0, 0.0); }
    __device__ __forceinline__ int nin(int) const { return 1; }
    __device__ __forceinline__ int src(int var, int) const { return var; }
    __device__ __forceinline__ double2 apply(int, int, double2 v, double, double) const { return make_double2(-v.x, -v.y); }
};

// ---------------------------------------------------------------- QG models (swqg/SWQG.jl:140-170, swqg/TwoLayerQG.jl:152-182)
// psi inversion on the fly; per layer: q, psi (psi_x = i k G in the x-pass), i l psi
__device__ __forceinline__ double2 qg_streamfunction(const double2* sol, long long vs, int nlayers, int layer, double K2, double P,
                                                     long long off) {
    if (nlayers == 1) {                                  // psih = -qh / (K^2 + Kd2)
        const double2 q = sol[off];
        const double inv = -1.0 / (K2 + P);
        return make_double2(inv * q.x, inv * q.y);
    }
    const double2 q1 = sol[off], q2 = sol[vs + off];     // psi_j = -(K^2 q_j + F (q1 + q2)) / ((K^2 + 2F) K^2)
    const double2 qj = layer == 0 ? q1 : q2;
    const double inv = K2 > 0.0 ? -1.0 / ((K2 + 2.0 * P) * K2) : 0.0;
    return make_double2(inv * (K2 * qj.x + P * (q1.x + q2.x)), inv * (K2 * qj.y + P * (q1.y + q2.y)));
}

struct QgLoaderA {  // jobs: layer*3 + {0 q, 1 psi, 2 i l psi}
    const double2* sol;
    long long vs;
    int nlayers;
    double P;  // Kd2 (SWQG) or F (two-layer)
    __device__ __forceinline__ double2 operator()(int job, int, int, double kw, double lw, long long off) const {
        const int layer = job / 3, which = job - 3 * layer;
        if (which == 0) return sol[layer * vs + off];
        const double2 psi = qg_streamfunction(sol, vs, nlayers, layer, kw * kw + lw * lw, P, off);
        return which == 1 ? psi : make_double2(-lw * psi.y, lw * psi.x);
    }
};

template <int N, int NL, bool SLAB = false>
struct QgXOp {  // per layer: a = psi_x q, b = psi_y q  ->  H[2 layer], H[2 layer + 1]
    static constexpr int NBUF = 2;
    const double2* G;  // [3 NL][ny][kr_pad]
    double2* H;        // [2 NL][ny][kr_pad]
    double sc;
    OutPeers peers;
    OutPeers gin;      // slab mode: sources of the input column segments
    int nj = 3 * NL;   // jobs the y-pass put into G (+ 3 when the snapshot's psi jobs ride along)
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        auto Gp = [&](int j) { return row_in<SLAB>(L, G, gin, nj, j, y); };
        auto Hp = [&](int j) { return row_out<SLAB>(L, H, peers, 2 * NL, j, y); };
        double *q1 = cx.re(0), *q2 = cx.im(0);
        double2 v[16];
        if (NL == 2) cx.template load_pair<MUL_ONE, MUL_ONE>(1, Gp(0), Gp(3));
        else cx.template load_pair<MUL_ONE, MUL_ZERO>(1, Gp(0), RowPlain{});
        cx.ifft_regs_out(1, v);                          // q1 + i q2
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            q1[x] = v[m].x;
            q2[x] = v[m].y;
        }
#pragma unroll
        for (int layer = 0; layer < NL; ++layer) {
            const double* q = layer == 0 ? q1 : q2;
            cx.template load_pair<MUL_IK, MUL_ONE>(1, Gp(3 * layer + 1), Gp(3 * layer + 2));
            cx.ifft_regs_out(1, v);                      // psi_x + i psi_y
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const double qq = sc * q[pad_index(cx.g + m * Gt)];
                v[m] = make_double2(qq * v[m].x, qq * v[m].y);
            }
            cx.fft_regs_in(v, 1);
            cx.template store_pair<MUL_ONE, MUL_ONE>(1, Hp(2 * layer), Hp(2 * layer + 1));
        }
    }
};

struct QgCombiner {  // N_layer = -i l F[psi_x q] + i k F[psi_y q]
    __device__ __forceinline__ int var_of(int slot) const { return slot; }
    __device__ __forceinline__ double2 init(int, double, double, long long) const { return make_double2(0.0, 0.0); }
    __device__ __forceinline__ int nin(int) const { return 2; }
    __device__ __forceinline__ int src(int var, int i) const { return 2 * var + i; }
    __device__ __forceinline__ double2 apply(int, int i, double2 v, double kw, double lw) const {
        return i == 0 ? make_double2(lw * v.y, -lw * v.x) : make_double2(-kw * v.y, kw * v.x);
    }
};

// GeophysicalFlows MultiLayerQG with two equal layers (raytracing/TwoLayerRaytracing.jl:174): L is the diagonal
// hyperviscosity and calcN! carries, besides the Jacobian, the mean-flow and background-PV-gradient terms and the bottom drag:
//   N_j = -i k F[(u_j + U_j) q_j] - i l F[v_j q_j] - F[v_j Qy_j]  (+ mu K^2 psih_2 for j = 2),  Qy = beta +- F (U_1 - U_2).
// A constant times a field transforms exactly, so the U_j and Qy_j terms are added in spectral space.
struct MlqgCombiner {
    const double2* sol;
    long long vs;
    double F, U1, U2, beta, mu;
    __device__ __forceinline__ int var_of(int slot) const { return slot; }
    __device__ __forceinline__ int nin(int) const { return 2; }
    __device__ __forceinline__ int src(int var, int i) const { return 2 * var + i; }
    __device__ __forceinline__ double2 apply(int, int i, double2 v, double kw, double lw) const {
        return i == 0 ? make_double2(lw * v.y, -lw * v.x) : make_double2(-kw * v.y, kw * v.x);
    }
    __device__ __forceinline__ double2 init(int var, double kw, double lw, long long off) const {
        const double K2 = kw * kw + lw * lw;
        const double2 q = sol[var * vs + off], psi = qg_streamfunction(sol, vs, 2, var, K2, F, off);
        const double Uj = var == 0 ? U1 : U2, Qy = var == 0 ? beta + F * (U1 - U2) : beta - F * (U1 - U2);
        const double drag = var == 1 ? mu * K2 : 0.0;
        // -i k (U_j q + Qy psi) + drag psi
        return make_double2(kw * (Uj * q.y + Qy * psi.y) + drag * psi.x, -kw * (Uj * q.x + Qy * psi.x) + drag * psi.y);
    }
};

// ---------------------------------------------------------------- Thomas-Yamada (thomasyamada/ThomasYamada.jl:129-274)
// state (zeta_t, u_c, v_c, p_c); psi_t = -zeta_t/K^2, u_t = -i l psi_t, v_t = i k psi_t
// y-jobs: 0 zeta_t, 1 psi_t, 2 u_t, 3 d_y u_t = l^2 psi_t, 4 u_c, 5 d_y u_c, 6 v_c, 7 p_c, 8 d_y p_c
struct TyLoaderA {
    const double2* sol;
    long long vs;
    __device__ __forceinline__ double2 operator()(int job, int, int, double kw, double lw, long long off) const {
        if (job < 4) {
            const double2 z = sol[off];
            if (job == 0) return z;
            const double K2 = kw * kw + lw * lw, inv = K2 > 0.0 ? -1.0 / K2 : 0.0;
            const double2 psi = make_double2(inv * z.x, inv * z.y);
            if (job == 1) return psi;
            if (job == 2) return make_double2(lw * psi.y, -lw * psi.x);
            return make_double2(lw * lw * psi.x, lw * lw * psi.y);
        }
        if (job == 4) return sol[vs + off];
        if (job == 5) { const double2 a = sol[vs + off]; return make_double2(-lw * a.y, lw * a.x); }
        if (job == 6) return sol[2 * vs + off];
        const double2 p = sol[3 * vs + off];
        return job == 7 ? p : make_double2(-lw * p.y, lw * p.x);
    }
};

// products (linear terms merged): p1 = vt zt, p2 = ut zt, p3 = uc vc, p4 = uc^2 - vc^2, p5 = ut uc, p6 = vt uc_y + vc ut_y,
// p7 = vt vc, p8 = ut vc_x + uc vt_x, p9 = ut pc_x + vt pc_y   ->  H[0..8]
template <int N, bool SLAB = false>
struct TyXOp {
    static constexpr int NBUF = 3;
    const double2* G;  // [9][ny][kr_pad]
    double2* H;        // [9][ny][kr_pad]
    double sc;
    OutPeers peers{};  // slab mode: destinations of the output column segments
    OutPeers gin{};    // slab mode: sources of the input column segments
    int nj = 9;        // jobs the y-pass put into G
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        auto Gp = [&](int j) { return row_in<SLAB>(L, G, gin, nj, j, y); };
        auto Hp = [&](int j) { return row_out<SLAB>(L, H, peers, 9, j, y); };
        const RowPlain none{};
        double *ut = cx.re(0), *vt = cx.im(0), *uc = cx.re(1), *vc = cx.im(1);
        double2 v[16];
        double ta[16], tb[16];
        cx.template load_pair<MUL_ONE, MUL_IK>(2, Gp(2), Gp(1));
        cx.ifft_regs_out(2, v);                          // ut + i vt
#pragma unroll
        for (int m = 0; m < 16; ++m) { const int x = pad_index(cx.g + m * Gt); ut[x] = v[m].x; vt[x] = v[m].y; }
        cx.template load_pair<MUL_ONE, MUL_ONE>(2, Gp(4), Gp(6));
        cx.ifft_regs_out(2, v);                          // uc + i vc
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            uc[x] = v[m].x; vc[x] = v[m].y;
            ta[m] = sc * (ut[x] * v[m].x);               // p5
            tb[m] = sc * (vt[x] * v[m].y);               // p7
            v[m] = make_double2(sc * (v[m].x * v[m].y), sc * (v[m].x * v[m].x - v[m].y * v[m].y));   // p3, p4
        }
        cx.fft_regs_in(v, 2);
        cx.template store_pair<MUL_ONE, MUL_ONE>(2, Hp(2), Hp(3));
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = make_double2(ta[m], tb[m]);
        cx.fft_regs_in(v, 2);
        cx.template store_pair<MUL_ONE, MUL_ONE>(2, Hp(4), Hp(6));
        cx.template load_pair<MUL_ONE, MUL_MK2>(2, Gp(0), Gp(1));
        cx.ifft_regs_out(2, v);                          // zeta_t + i d_x v_t
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            tb[m] = sc * (uc[x] * v[m].y);               // uc vt_x  (half of p8)
            v[m] = make_double2(sc * (vt[x] * v[m].x), sc * (ut[x] * v[m].x));   // p1, p2
        }
        cx.fft_regs_in(v, 2);
        cx.template store_pair<MUL_ONE, MUL_ONE>(2, Hp(0), Hp(1));
        cx.template load_pair<MUL_ONE, MUL_ONE>(2, Gp(5), Gp(3));
        cx.ifft_regs_out(2, v);                          // d_y u_c + i d_y u_t
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            ta[m] = sc * (vt[x] * v[m].x + vc[x] * v[m].y);   // p6
        }
        cx.template load_pair<MUL_IK, MUL_IK>(2, Gp(6), Gp(7));
        cx.ifft_regs_out(2, v);                          // d_x v_c + i d_x p_c
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            const double p9a = sc * (ut[x] * v[m].y);
            v[m] = make_double2(ta[m], tb[m] + sc * (ut[x] * v[m].x));   // p6, p8
            ta[m] = p9a;
        }
        cx.fft_regs_in(v, 2);
        cx.template store_pair<MUL_ONE, MUL_ONE>(2, Hp(5), Hp(7));
        cx.template load_pair<MUL_ONE, MUL_ZERO>(2, Gp(8), none);
        cx.ifft_regs_out(2, v);                          // d_y p_c
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = pad_index(cx.g + m * Gt);
            v[m] = make_double2(ta[m] + sc * (vt[x] * v[m].x), 0.0);    // p9
        }
        cx.fft_regs_in(v, 2);
        cx.template store_pair<MUL_ONE, MUL_ZERO>(2, Hp(8), none);
    }
};

struct TyCombiner {
    const double2* sol;
    long long vs;
    double Ro;
    __device__ __forceinline__ int var_of(int slot) const { return slot; }
    __device__ __forceinline__ int nin(int var) const { return var == 0 ? 4 : (var == 3 ? 1 : 2); }
    __device__ __forceinline__ int src(int var, int i) const { return var == 0 ? i : (var == 1 ? 4 + i : (var == 2 ? 6 + i : 8)); }
    __device__ __forceinline__ double2 apply(int var, int i, double2 v, double kw, double lw) const {
        if (var == 0) {
            if (i == 0) return make_double2(Ro * lw * v.y, -Ro * lw * v.x);                  // -Ro i l
            if (i == 1) return make_double2(Ro * kw * v.y, -Ro * kw * v.x);                  // -Ro i k
            const double w = i == 2 ? -Ro * (lw * lw - kw * kw) : -Ro * kw * lw;
            return make_double2(w * v.x, w * v.y);
        }
        if (var != 3 && i == 0) {
            const double w = Ro * (var == 1 ? kw : lw);
            return make_double2(w * v.y, -w * v.x);                                          // -Ro i k | -Ro i l
        }
        return make_double2(-Ro * v.x, -Ro * v.y);
    }
    // linear f-plane wave terms (:143-146)
    __device__ __forceinline__ double2 init(int var, double kw, double lw, long long off) const {
        if (var == 0) return make_double2(0.0, 0.0);
        const double2 u = sol[vs + off], v = sol[2 * vs + off], p = sol[3 * vs + off];
        if (var == 1) return make_double2(v.x + kw * p.y, v.y - kw * p.x);                   // vc - i k pc
        if (var == 2) return make_double2(-u.x + lw * p.y, -u.y - lw * p.x);                 // -uc - i l pc
        return make_double2(kw * u.y + lw * v.y, -(kw * u.x + lw * v.x));                    // -i k uc - i l vc
    }
};

// ---------------------------------------------------------------- plain spectral -> physical
// one job: an expression of the state selected by `which` (see swrt.h SWRT_FIELD_*)
struct FieldLoader {
    const double2* sol;
    long long vs;
    int which, nlayers;
    double f, P;
    __device__ __forceinline__ double2 operator()(int, int, int, double kw, double lw, long long off) const {
        if (which < 16) return sol[(long long)which * vs + off];
        if (which == 16) {   // RSW linear PV  zeta = i k vh - i l uh - f etah   (rsw/RotatingShallowWater.jl:108)
            const double2 u = sol[off], v = sol[vs + off], e = sol[2 * vs + off];
            return make_double2(-(kw * v.y - lw * u.y) - f * e.x, (kw * v.x - lw * u.x) - f * e.y);
        }
        // QG diagnostics (swqg/SWQG.jl:109-125, swqg/TwoLayerQG.jl:113-129): 32+j psi_j, 40+j u_j = -i l psi, 48+j v_j = i k psi,
        // 56+j zeta_j = -K^2 psi
        const int kind = (which - 32) >> 3, layer = (which - 32) & 7;
        const double K2 = kw * kw + lw * lw;
        const double2 psi = qg_streamfunction(sol, vs, nlayers, layer, K2, P, off);
        if (kind == 0) return psi;
        if (kind == 1) return make_double2(lw * psi.y, -lw * psi.x);
        if (kind == 2) return make_double2(-kw * psi.y, kw * psi.x);
        return make_double2(-K2 * psi.x, -K2 * psi.y);
    }
};

template <int N, bool SLAB = false>
struct C2ROp {
    static constexpr int NBUF = 2;
    const double2* G;  // [1][ny][kr_pad]
    double* out;       // [ny][nx]
    double s1;
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        double2 v[16];
        cx.template load_pair_regs<MUL_ONE, MUL_ZERO>(v, row_ref<SLAB>(L, G, 1, 0, y), RowPlain{});
        cx.ifft_regs(v, 1);
#pragma unroll
        for (int m = 0; m < 16; ++m) out[(long long)y * N + cx.g + m * Gt] = s1 * v[m].x;
    }
};

// physical -> spectral: forward transform of one real field (set a state variable from physical space,
// e.g. m = 1/(1+eta) of the QuadHeight model, or initial conditions given on the grid)
template <int N, bool SLAB = false>
struct R2COp {
    static constexpr int NBUF = 2;
    const double* in;  // [ny][nx]
    double2* H;        // [1][ny][kr_pad]
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = make_double2(0.5 * in[(long long)y * N + cx.g + m * Gt], 0.0);   // store_pair returns 2 P
        cx.fft_regs_in(v, 1);
        cx.template store_pair<MUL_ONE, MUL_ZERO>(1, row_ref<SLAB>(L, H, 1, 0, y), RowPlain{});
    }
};
struct IdentityCombiner {
    __device__ __forceinline__ int var_of(int slot) const { return slot; }
    __device__ __forceinline__ int nin(int) const { return 1; }
    __device__ __forceinline__ int src(int, int) const { return 0; }
    __device__ __forceinline__ double2 apply(int, int, double2 v, double, double) const { return v; }
    __device__ __forceinline__ double2 init(int, double, double, long long) const { return make_double2(0.0, 0.0); }
};

// ---------------------------------------------------------------- velocity snapshot for packets
// psi kinds
enum { PSI_RSW_BALANCED = 0, PSI_SWQG = 1, PSI_TWOLAYER_BAROCLINIC = 2, PSI_TWOLAYER_MEAN = 3 };
// jobs: 0 psih, 1 -i l psih (u), 2 l^2 psih (uy);   v = i k G0, ux = i k G1, vx = -k^2 G0
struct PsiLoader {
    const double2* sol;
    long long vs;
    int kind;
    double f, P;   // P: Kd2 = f^2/Cg^2 (RSW balanced, SWQG) or F (two-layer)
    double pdx, pdy;   // > 0: B-spline prefilter (Interpolations.jl, raytracing/Raytracing.jl:152-170) folded into psi: every
                       // sampled field is linear in psih, and the prefilter is 1/((pc0 + pc1 cos(k dx))(pc0 + pc1 cos(l dy)))
    double pc0 = 0.75, pc1 = 0.25;   // quadratic 3/4, 1/4; cubic 2/3, 1/3
    // type-2 NUFFT mode: deconvolution by the sampling kernel's Fourier transform, tabulated per kr index and per l index
    // (1 / phihat; api.cu nufft_tables); kr_pad / kr_off recover the indices from the element offset
    const double *ptab_x = nullptr, *ptab_y = nullptr;
    int tab_kr_pad = 1, tab_kr_off = 0;
    __device__ __forceinline__ double2 psi_of(double kw, double lw, long long off) const {
        double2 r = psi_raw(kw, lw, off);
        if (ptab_x) {
            const double pf = ptab_x[tab_kr_off + (int)(off % tab_kr_pad)] * ptab_y[(int)(off / tab_kr_pad)];
            r.x *= pf;
            r.y *= pf;
        } else if (pdx > 0.0) {
            const double pf = 1.0 / ((pc0 + pc1 * cos(kw * pdx)) * (pc0 + pc1 * cos(lw * pdy)));
            r.x *= pf;
            r.y *= pf;
        }
        return r;
    }
    __device__ __forceinline__ double2 psi_raw(double kw, double lw, long long off) const {
        const double K2 = kw * kw + lw * lw;
        if (kind == PSI_RSW_BALANCED) {   // -(i k vh - i l uh - f etah)/(K^2 + f^2/Cg^2)   rsw/RSWRaytracingDriver.jl:63-67
            const double2 u = sol[off], v = sol[vs + off], e = sol[2 * vs + off];
            const double inv = -1.0 / (K2 + P);
            return make_double2(inv * (-(kw * v.y - lw * u.y) - f * e.x), inv * ((kw * v.x - lw * u.x) - f * e.y));
        }
        if (kind == PSI_SWQG) return qg_streamfunction(sol, vs, 1, 0, K2, P, off);
        const double2 a = qg_streamfunction(sol, vs, 2, 0, K2, P, off), b = qg_streamfunction(sol, vs, 2, 1, K2, P, off);
        // baroclinic 0.5 (psi1 - psi2): swqg/TwoLayerRaytracingDriver.jl:232 ; mean (psi1 + psi2)/2: raytracing/TwoLayerRaytracing.jl:122
        return kind == PSI_TWOLAYER_BAROCLINIC ? make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y)) : make_double2((a.x + b.x) / 2, (a.y + b.y) / 2);
    }
    __device__ __forceinline__ double2 operator()(int job, int, int, double kw, double lw, long long off) const {
        const double2 psi = psi_of(kw, lw, off);
        if (job == 0) return psi;
        if (job == 1) return make_double2(lw * psi.y, -lw * psi.x);
        return make_double2(lw * lw * psi.x, lw * lw * psi.y);
    }
};

// Team mode: the snapshot's three y-jobs (psih, -i l psih, l^2 psih from the materialised streamfunction of the SAME state) appended
// to the model's own stage-A jobs, so that one y-pass, one transpose and one barrier serve both the flow step and the snapshot.
template <class A>
struct FusedLoaderA {
    A a;
    const double2* psih;
    int na;
    __device__ __forceinline__ double2 operator()(int job, int kr, int l, double kw, double lw, long long off) const {
        if (job < na) return a(job, kr, l, kw, lw, off);
        const double2 p = psih[off];
        return job == na ? p : job == na + 1 ? make_double2(lw * p.y, -lw * p.x) : make_double2(lw * lw * p.x, lw * lw * p.y);
    }
};

template <int N, bool SLAB = false, bool F32 = false>
struct SnapshotXOp {
    static constexpr int NBUF = 2;
    const double2* G;  // [3][ny][kr_pad]
    double* out;       // [ny][nx][6] of the level being written ([ny][nx][8] floats in the fp32 packet mode)
    double s1;
    OutPeers gin;      // slab mode: sources of the input column segments
    int nj = 3, j0 = 0;   // the three psi jobs are jobs j0 .. j0 + 2 of nj (they ride behind the model's jobs in team mode)
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G, EPT = XCtx<N>::EPT;
        static_assert(EPT == 16, "one register per owned point");
        const auto Gp = row_in<SLAB>(L, G, gin, nj, j0, y), Gu = row_in<SLAB>(L, G, gin, nj, j0 + 1, y), Guy = row_in<SLAB>(L, G, gin, nj, j0 + 2, y);
        double2 vx[16];
        cx.template load_pair<MUL_MK2, MUL_ZERO>(1, Gp, RowPlain{});
        cx.ifft_regs_out(1, vx);  // vx
        cx.template load_pair<MUL_ONE, MUL_IK>(0, Gu, Gp);
        cx.ifft(0);  // u + i v
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gu, Guy);
        cx.ifft(1);  // ux + i uy
        if (F32) {
            float4* o = reinterpret_cast<float4*>(out) + (long long)y * N * 2;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int x = cx.g + i * Gt, p = pad_index(x);
                o[2 * x] = make_float4((float)(s1 * cx.re(0)[p]), (float)(s1 * cx.im(0)[p]), (float)(s1 * cx.re(1)[p]), (float)(s1 * cx.im(1)[p]));
                o[2 * x + 1] = make_float4((float)(s1 * vx[i].x), 0.f, 0.f, 0.f);
            }
            return;
        }
        double* o = out + (long long)y * N * SNAP_STRIDE;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int x = cx.g + i * Gt, p = pad_index(x);
            double2* q = reinterpret_cast<double2*>(o + (long long)x * SNAP_STRIDE);   // 48-byte record, three 16-byte stores
            q[0] = make_double2(s1 * cx.re(0)[p], s1 * cx.im(0)[p]);
            q[1] = make_double2(s1 * cx.re(1)[p], s1 * cx.im(1)[p]);
            q[2] = make_double2(s1 * vx[i].x, 0.0);
            // Point after point (the warp barrier is what keeps the machine code in this order).  Left to the scheduler, all 64
            // shared-memory reads are hoisted and the 48 stores of a thread leave as one burst at the end of the row (each a
            // 16-byte piece per lane at a 48-byte stride, 12 lines per warp instruction), which stalls the other warps' traffic
            // behind it: 0.116 ms per launch against 0.104 ms paced; storing each transform's piece as soon as it is ready (three
            // passes of 16 stores, no buffering) writes partial sectors and costs 0.151 ms (profiles/r02_README.md).
            __syncwarp();
        }
    }
};

// Hermite-bicubic node data from the same three y-jobs (G0 = psi, G1 = -i l psi, G2 = l^2 psi):
// u = G1, v = i k G0, ux = i k G1, uy = G2, vx = -k^2 G0, uxy = i k G2, vxy = psi_xxy = k^2 G1
template <int N, bool SLAB = false>
struct SnapshotCubicXOp {
    static constexpr int NBUF = 2;
    const double2* G;  // [3][ny][kr_pad]
    double* out;       // [ny][nx][8]
    double s1;
    __device__ __forceinline__ void row(const XCtx<N>& cx, const SpecLayout& L, int y) const {
        constexpr int Gt = XCtx<N>::G;
        const auto Gp = row_ref<SLAB>(L, G, 3, 0, y), Gu = row_ref<SLAB>(L, G, 3, 1, y), Guy = row_ref<SLAB>(L, G, 3, 2, y);
        double2* o = reinterpret_cast<double2*>(out + (long long)y * N * SNAP3_STRIDE);
        double2 v[16];
        auto put = [&](int slot) {
#pragma unroll
            for (int m = 0; m < 16; ++m) o[(long long)(cx.g + m * Gt) * (SNAP3_STRIDE / 2) + slot] = make_double2(s1 * v[m].x, s1 * v[m].y);
        };
        cx.template load_pair<MUL_ONE, MUL_IK>(1, Gu, Gp);
        cx.ifft_regs_out(1, v);
        put(0);                                          // u, v
        cx.template load_pair<MUL_IK, MUL_ONE>(1, Gu, Guy);
        cx.ifft_regs_out(1, v);
        put(1);                                          // ux, uy
        cx.template load_pair<MUL_MK2, MUL_IK>(1, Gp, Guy);
        cx.ifft_regs_out(1, v);
        put(2);                                          // vx, uxy
        cx.template load_pair<MUL_K2, MUL_ZERO>(1, Gu, RowPlain{});
        cx.ifft_regs_out(1, v);
        put(3);                                          // vxy, 0
    }
};

// ---------------------------------------------------------------- per-size launchers
#ifndef SWRT_TK_LARGE
#define SWRT_TK_LARGE 4   // columns per y-pass CTA for N >= 2048 (tuning knob)
#endif
__host__ __device__ constexpr int tile_k(int N) { return N >= 4096 ? 2 : (N >= 2048 ? SWRT_TK_LARGE : (N >= 256 ? 4096 / N : 16)); }
template <int N>
struct Launch {
    static constexpr int TK = tile_k(N);
    static constexpr int G = group_size(N);
    static constexpr size_t ysmem = (size_t)ypass_smem(N, TK);

    // one-time per kernel: opt in to the dynamic shared memory and size a persistent grid
    template <class K>
    static cudaError_t prep(K kernel, size_t smem, int threads, int* ctas) {
        static int cached = 0;  // one static per (N, K) instantiation
        static size_t cached_smem = 0;
        if (!cached || smem != cached_smem) {   // the staged y-passes size their buffer by the retained rows: re-derive when it changes
            cached_smem = smem;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0, occ = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
            if (e != cudaSuccess) return e;
            cached = sms * (occ > 0 ? occ : 1);
        }
        *ctas = cached;
        return cudaSuccess;
    }
    template <class Loader>
    static cudaError_t ypass_inv(const Loader& ld, const SpecLayout& L, int njobs, const OutPeers& out, const double2* tw,
                                 cudaStream_t st) {
        auto k = ypass_inv_kernel<N, TK, Loader>;
        int mc = 1;
        cudaError_t e = prep(k, ysmem, TK * G, &mc);
        if (e != cudaSuccess) return e;
        const int work = ((L.kr_keep + TK - 1) / TK) * njobs;
        if (work == 0) return cudaSuccess;   // a slab rank beyond the retained columns owns nothing
        k<<<work < mc ? work : mc, TK * G, ysmem, st>>>(ld, L, njobs, out, tw);
        return cudaGetLastError();
    }
    // prefetching variant for simple jobs; falls back to the plain kernel when the staging buffer does not fit
    static constexpr bool kPrefetchFits = ysmem + (size_t)ypass_stage_smem(N, TK) * 2 / 3 + 4096 <= (size_t)kSmemPerSM;
    template <class Loader>
    static cudaError_t ypass_inv_simple(const SimpleJobs& jobs, const Loader& fallback, const SpecLayout& L, int njobs, const OutPeers& out,
                                        const double2* tw, cudaStream_t st) {
        const size_t stage = (size_t)(L.ny - (L.lz1 - L.lz0)) * TK * 16;
        static const int enabled = [] { const char* e = getenv("SWRT_YPASS_PREFETCH"); return e ? atoi(e) : 1; }();
        if constexpr (kPrefetchFits) {
            if (enabled && ysmem + stage + 1024 <= (size_t)kSmemPerSM) {
                auto k = ypass_inv_prefetch_kernel<N, TK>;
                int mc = 1;
                cudaError_t e = prep(k, ysmem + stage, TK * G, &mc);
                if (e != cudaSuccess) return e;
                const int work = ((L.kr_keep + TK - 1) / TK) * njobs;
                if (work == 0) return cudaSuccess;
                k<<<work < mc ? work : mc, TK * G, ysmem + stage, st>>>(jobs, L, njobs, out, tw);
                return cudaGetLastError();
            }
        }
        return ypass_inv(fallback, L, njobs, out, tw, st);
    }
    template <class Combiner>
    static cudaError_t ypass_fwd(const Combiner& cb, const SpecLayout& L, int nvars, int nh, const double2* H, double2* out,
                                 const double2* tw, cudaStream_t st) {
        static const int enabled = [] { const char* e = getenv("SWRT_YPASS_PREFETCH"); return e ? atoi(e) : 1; }();
        if constexpr (kPrefetchFits) {
            if (enabled) {
                // as many staged rows as fit beside the work buffers; two CTAs per SM when each can still stage half of the rows
                // (narrow tiles: the phases of one CTA -- wait, first stage, barriers, stores -- then overlap the other's)
                const size_t per_sm = 228 * 1024;
                const bool two = 2 * (ysmem + 1024 + (size_t)(L.ny / 2) * TK * 16) <= per_sm;
                int rows_s = (int)(((two ? per_sm / 2 : (size_t)kSmemPerSM) - ysmem - 1024) / ((size_t)TK * 16));
                if (rows_s > L.ny) rows_s = L.ny;
                const size_t smem = ysmem + (size_t)rows_s * TK * 16;
                auto kp = ypass_fwd_prefetch_kernel<N, TK, Combiner>;
                int mcp = 1;
                cudaError_t ep = prep(kp, smem, TK * G, &mcp);
                if (ep != cudaSuccess) return ep;
                const int workp = ((L.kr_keep + TK - 1) / TK) * nvars;
                if (workp == 0) return cudaSuccess;
                kp<<<workp < mcp ? workp : mcp, TK * G, smem, st>>>(cb, L, nvars, nh, rows_s, H, out, tw);
                return cudaGetLastError();
            }
        }
        auto k = ypass_fwd_kernel<N, TK, Combiner>;
        int mc = 1;
        cudaError_t e = prep(k, ysmem, TK * G, &mc);
        if (e != cudaSuccess) return e;
        const int work = ((L.kr_keep + TK - 1) / TK) * nvars;
        if (work == 0) return cudaSuccess;
        k<<<work < mc ? work : mc, TK * G, ysmem, st>>>(cb, L, nvars, nh, H, out, tw);
        return cudaGetLastError();
    }
    template <class Op>
    static cudaError_t xpass(const Op& op, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st) {
        auto k = xpass_kernel<N, Op>;
        constexpr size_t smem = (size_t)xpass_smem(N, Op::NBUF);
        int mc = 1;
        cudaError_t e = prep(k, smem, G, &mc);
        if (e != cudaSuccess) return e;
        k<<<L.yrows < mc ? L.yrows : mc, G, smem, st>>>(op, L, tw, sched);
        return cudaGetLastError();
    }

    // concrete entry points (explicitly specialised per size in inst.cu); `model` = SWRT_* model id
    static cudaError_t stage_a(int model, const double2* sol, const OutPeers& G_, const SpecLayout& L, const double2* tw, cudaStream_t st);
    static cudaError_t stage_b(int model, const double2* G_, double2* H, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st);
    // slab mode (P > 1): segmented rows; built for the models of the >= 4096^2 configurations (RSW, SWQG, two-layer QG)
    static cudaError_t stage_b_slab(int model, const OutPeers& Gin, const OutPeers& H, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st,
                                    int nj_total = 0);
    static cudaError_t snap_stage_b_slab(const OutPeers& Gin, double* out, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st,
                                         int nj_total = 3, int j0 = 0);
    // team mode: stage A of the model + the three psi jobs of the snapshot of the same state (psih materialised by psi_kernel)
    static cudaError_t stage_a_fused(int model, const double2* sol, const double2* psih, const OutPeers& G_, const SpecLayout& L, const double2* tw, cudaStream_t st);
    static cudaError_t stage_c(int model, const double2* sol, const double2* H, double2* Nout, const SpecLayout& L, const double2* tw, cudaStream_t st);
    static cudaError_t field_stage_a(const FieldLoader& ld, const SpecLayout& L, const OutPeers& G_, const double2* tw, cudaStream_t st);
    static cudaError_t field_stage_b(const double2* G_, double* out, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st);
    // physical field -> spectral field `out` ([l][kr_pad]) through H scratch
    static cudaError_t forward_field(const double* in, double2* H, double2* out, const SpecLayout& L, const double2* tw_x, unsigned* sched, cudaStream_t st);
    static cudaError_t forward_field_y(const double2* H, double2* out, const SpecLayout& L, const double2* tw_y, cudaStream_t st);
    static cudaError_t psi_stage_a(const PsiLoader& ld, const double2* psih, const SpecLayout& L, const OutPeers& G_, const double2* tw, cudaStream_t st);
    static constexpr bool psi_prefetch = kPrefetchFits;   // psih must have been materialised (update.cuh psi_kernel) when true
    static cudaError_t snap_stage_b(const double2* G_, double* out, int cubic, const SpecLayout& L, const double2* tw, unsigned* sched, cudaStream_t st,
                                    double s1 = 0.0);
    // spectrally refined snapshot (zero-padded transform onto a finer node grid): psih is already laid out for L
    static cudaError_t psi_stage_a_refined(const double2* psih, const SpecLayout& L, const OutPeers& G_, const double2* tw, cudaStream_t st);
};

}  // namespace swrt
