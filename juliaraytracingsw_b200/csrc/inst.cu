// Per-size explicit instantiation of the pass kernels (compiled once per SWRT_N so that
// `make -j` builds the sizes in parallel).
#include "models.cuh"

#ifndef SWRT_N
#error "compile with -DSWRT_N=<transform length>"
#endif

namespace swrt {

template <>
cudaError_t Launch<SWRT_N>::stage_a(int model, const double2* sol, const OutPeers& G_, const SpecLayout& L, const double2* tw, cudaStream_t st) {
    switch (model) {
        case MODEL_RSW:
        case MODEL_RSW_QUADHEIGHT:
        case MODEL_RSW_MODIFIED: {
            SimpleJobs sj{};
            const int fld[5] = {0, 1, 2, 0, 1}, mul[5] = {YMUL_ONE, YMUL_ONE, YMUL_ONE, YMUL_IL, YMUL_IL};
            const int last[5] = {0, 0, 1, 1, 1};   // eta is read once; the i l u, i l v jobs are the second readers of u, v
            for (int j = 0; j < 5; ++j) { sj.src[j] = sol + fld[j] * L.vs; sj.mul[j] = mul[j]; sj.last[j] = last[j]; }
            return ypass_inv_simple(sj, RswLoaderA{sol, L.vs}, L, 5, G_, tw, st);
        }
        case MODEL_RSW_LINDBORG: return ypass_inv(LindborgLoaderA{sol, L.vs}, L, 8, G_, tw, st);
        case MODEL_SWQG: return ypass_inv(QgLoaderA{sol, L.vs, 1, L.aux0}, L, 3, G_, tw, st);
        case MODEL_MULTILAYERQG2:
        case MODEL_TWOLAYERQG: return ypass_inv(QgLoaderA{sol, L.vs, 2, L.aux0}, L, 6, G_, tw, st);
        case MODEL_THOMASYAMADA: return ypass_inv(TyLoaderA{sol, L.vs}, L, 9, G_, tw, st);
    }
    return cudaErrorInvalidValue;
}
template <>
cudaError_t Launch<SWRT_N>::stage_b(int model, const double2* G_, double2* H, const SpecLayout& L, const double2* tw, unsigned* sched,
                                    cudaStream_t st) {
    const double s1 = 1.0 / ((double)L.nx * (double)L.ny), sc = 0.5 * s1 * s1;
    switch (model) {
        case MODEL_RSW: return xpass(RswXOp<SWRT_N, 0>{G_, H, sc, s1, OutPeers{}}, L, tw, sched, st);
        case MODEL_RSW_MODIFIED: return xpass(RswXOp<SWRT_N, 1>{G_, H, sc, s1, OutPeers{}}, L, tw, sched, st);
        case MODEL_RSW_QUADHEIGHT: return xpass(RswXOp<SWRT_N, 2>{G_, H, sc, s1, OutPeers{}}, L, tw, sched, st);
        case MODEL_RSW_LINDBORG: return xpass(LindborgXOp<SWRT_N>{G_, H, sc}, L, tw, sched, st);
        case MODEL_SWQG: return xpass(QgXOp<SWRT_N, 1>{G_, H, sc, OutPeers{}}, L, tw, sched, st);
        case MODEL_MULTILAYERQG2:
        case MODEL_TWOLAYERQG: return xpass(QgXOp<SWRT_N, 2>{G_, H, sc, OutPeers{}}, L, tw, sched, st);
        case MODEL_THOMASYAMADA: return xpass(TyXOp<SWRT_N>{G_, H, sc}, L, tw, sched, st);
    }
    return cudaErrorInvalidValue;
}
template <>
cudaError_t Launch<SWRT_N>::stage_b_slab(int model, const OutPeers& Gin, const OutPeers& H, const SpecLayout& L, const double2* tw,
                                         unsigned* sched, cudaStream_t st, int nj_total) {
    const double s1 = 1.0 / ((double)L.nx * (double)L.ny), sc = 0.5 * s1 * s1;
    const int nj = nj_total > 0 ? nj_total : model_njobs_a(model);
    switch (model) {
        case MODEL_RSW: return xpass(RswXOp<SWRT_N, 0, true>{nullptr, nullptr, sc, s1, H, Gin, nj}, L, tw, sched, st);
        case MODEL_RSW_MODIFIED: return xpass(RswXOp<SWRT_N, 1, true>{nullptr, nullptr, sc, s1, H, Gin, nj}, L, tw, sched, st);
        case MODEL_RSW_QUADHEIGHT: return xpass(RswXOp<SWRT_N, 2, true>{nullptr, nullptr, sc, s1, H, Gin, nj}, L, tw, sched, st);
        case MODEL_RSW_LINDBORG: return xpass(LindborgXOp<SWRT_N, true>{nullptr, nullptr, sc, H, Gin, nj}, L, tw, sched, st);
        case MODEL_THOMASYAMADA: return xpass(TyXOp<SWRT_N, true>{nullptr, nullptr, sc, H, Gin, nj}, L, tw, sched, st);
        case MODEL_SWQG: return xpass(QgXOp<SWRT_N, 1, true>{nullptr, nullptr, sc, H, Gin, nj}, L, tw, sched, st);
        case MODEL_MULTILAYERQG2:
        case MODEL_TWOLAYERQG: return xpass(QgXOp<SWRT_N, 2, true>{nullptr, nullptr, sc, H, Gin, nj}, L, tw, sched, st);
    }
    return cudaErrorInvalidValue;
}
template <>
cudaError_t Launch<SWRT_N>::snap_stage_b_slab(const OutPeers& Gin, double* out, const SpecLayout& L, const double2* tw, unsigned* sched,
                                              cudaStream_t st, int nj_total, int j0) {
    return xpass(SnapshotXOp<SWRT_N, true>{nullptr, out, 1.0 / ((double)L.nx * (double)L.ny), Gin, nj_total, j0}, L, tw, sched, st);
}
template <>
cudaError_t Launch<SWRT_N>::stage_a_fused(int model, const double2* sol, const double2* psih, const OutPeers& G_, const SpecLayout& L, const double2* tw,
                                          cudaStream_t st) {
    switch (model) {
        case MODEL_RSW_MODIFIED:
        case MODEL_RSW_QUADHEIGHT:
        case MODEL_RSW: {   // 5 + 3 simple jobs: the prefetching y-pass
            SimpleJobs sj{};
            const int fld[5] = {0, 1, 2, 0, 1}, mul[8] = {YMUL_ONE, YMUL_ONE, YMUL_ONE, YMUL_IL, YMUL_IL, YMUL_ONE, YMUL_NEG_IL, YMUL_L2};
            const int last[8] = {0, 0, 1, 1, 1, 0, 0, 1};
            for (int j = 0; j < 5; ++j) sj.src[j] = sol + fld[j] * L.vs;
            for (int j = 5; j < 8; ++j) sj.src[j] = psih;
            for (int j = 0; j < 8; ++j) { sj.mul[j] = mul[j]; sj.last[j] = last[j]; }
            return ypass_inv_simple(sj, FusedLoaderA<RswLoaderA>{RswLoaderA{sol, L.vs}, psih, 5}, L, 8, G_, tw, st);
        }
        case MODEL_RSW_LINDBORG: return ypass_inv(FusedLoaderA<LindborgLoaderA>{LindborgLoaderA{sol, L.vs}, psih, 8}, L, 11, G_, tw, st);
        case MODEL_SWQG: return ypass_inv(FusedLoaderA<QgLoaderA>{QgLoaderA{sol, L.vs, 1, L.aux0}, psih, 3}, L, 6, G_, tw, st);
        case MODEL_MULTILAYERQG2:
        case MODEL_TWOLAYERQG: return ypass_inv(FusedLoaderA<QgLoaderA>{QgLoaderA{sol, L.vs, 2, L.aux0}, psih, 6}, L, 9, G_, tw, st);
    }
    return cudaErrorInvalidValue;
}
template <>
cudaError_t Launch<SWRT_N>::stage_c(int model, const double2* sol, const double2* H, double2* Nout, const SpecLayout& L, const double2* tw, cudaStream_t st) {
    switch (model) {
        case MODEL_RSW: return ypass_fwd(RswCombiner{0, L.Cg2}, L, 3, 4, H, Nout, tw, st);
        case MODEL_RSW_QUADHEIGHT:
        case MODEL_RSW_MODIFIED: return ypass_fwd(RswCombiner{1, L.Cg2}, L, 3, 5, H, Nout, tw, st);
        case MODEL_RSW_LINDBORG: return ypass_fwd(NegateCombiner{}, L, 3, 3, H, Nout, tw, st);
        case MODEL_SWQG: return ypass_fwd(QgCombiner{}, L, 1, 2, H, Nout, tw, st);
        case MODEL_TWOLAYERQG: return ypass_fwd(QgCombiner{}, L, 2, 4, H, Nout, tw, st);
        case MODEL_MULTILAYERQG2: return ypass_fwd(MlqgCombiner{sol, L.vs, L.aux0, L.aux2, L.aux3, L.aux4, L.aux5}, L, 2, 4, H, Nout, tw, st);
        case MODEL_THOMASYAMADA: return ypass_fwd(TyCombiner{sol, L.vs, L.aux1}, L, 4, 9, H, Nout, tw, st);
    }
    return cudaErrorInvalidValue;
}
template <>
cudaError_t Launch<SWRT_N>::field_stage_a(const FieldLoader& ld, const SpecLayout& L, const OutPeers& G_, const double2* tw, cudaStream_t st) {
    return ypass_inv(ld, L, 1, G_, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::field_stage_b(const double2* G_, double* out, const SpecLayout& L, const double2* tw, unsigned* sched,
                                          cudaStream_t st) {
    C2ROp<SWRT_N> op{G_, out, 1.0 / ((double)L.nx * (double)L.ny)};
    return xpass(op, L, tw, sched, st);
}
template <>
cudaError_t Launch<SWRT_N>::forward_field(const double* in, double2* H, double2*, const SpecLayout& L, const double2* tw_x, unsigned* sched,
                                          cudaStream_t st) {
    return xpass(R2COp<SWRT_N>{in, H}, L, tw_x, sched, st);
}
template <>
cudaError_t Launch<SWRT_N>::forward_field_y(const double2* H, double2* out, const SpecLayout& L, const double2* tw_y, cudaStream_t st) {
    return ypass_fwd(IdentityCombiner{}, L, 1, 1, H, out, tw_y, st);
}
template <>
cudaError_t Launch<SWRT_N>::psi_stage_a(const PsiLoader& ld, const double2* psih, const SpecLayout& L, const OutPeers& G_, const double2* tw,
                                        cudaStream_t st) {
    if (psih) {   // jobs: psih, -i l psih, l^2 psih from the materialised field
        SimpleJobs sj{};
        const int mul[3] = {YMUL_ONE, YMUL_NEG_IL, YMUL_L2};
        for (int j = 0; j < 3; ++j) { sj.src[j] = psih; sj.mul[j] = mul[j]; sj.last[j] = j == 2; }
        return ypass_inv_simple(sj, ld, L, 3, G_, tw, st);
    }
    return ypass_inv(ld, L, 3, G_, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::psi_stage_a_refined(const double2* psih, const SpecLayout& L, const OutPeers& G_, const double2* tw, cudaStream_t st) {
    if constexpr (kPrefetchFits) {
        const size_t stage = (size_t)(L.ny - (L.lz1 - L.lz0)) * TK * 16;
        if (ysmem + stage + 1024 > (size_t)kSmemPerSM) return cudaErrorNotSupported;
        SimpleJobs sj{};
        const int mul[3] = {YMUL_ONE, YMUL_NEG_IL, YMUL_L2};
        for (int j = 0; j < 3; ++j) { sj.src[j] = psih; sj.mul[j] = mul[j]; sj.last[j] = j == 2; }
        auto k = ypass_inv_prefetch_kernel<SWRT_N, TK>;
        int mc = 1;
        cudaError_t e = prep(k, ysmem + stage, TK * G, &mc);
        if (e != cudaSuccess) return e;
        const int work = ((L.kr_keep + TK - 1) / TK) * 3;
        if (work == 0) return cudaSuccess;
        k<<<work < mc ? work : mc, TK * G, ysmem + stage, st>>>(sj, L, 3, G_, tw);
        return cudaGetLastError();
    } else {
        return cudaErrorNotSupported;
    }
}
template <>
cudaError_t Launch<SWRT_N>::snap_stage_b(const double2* G_, double* out, int cubic, const SpecLayout& L, const double2* tw, unsigned* sched,
                                         cudaStream_t st, double s1_in) {
    const double s1 = s1_in > 0.0 ? s1_in : 1.0 / ((double)L.nx * (double)L.ny);
    if (cubic == 2) return xpass(SnapshotXOp<SWRT_N, false, true>{G_, out, s1}, L, tw, sched, st);   // fp32 node records
    if (cubic) return xpass(SnapshotCubicXOp<SWRT_N>{G_, out, s1}, L, tw, sched, st);
    return xpass(SnapshotXOp<SWRT_N>{G_, out, s1}, L, tw, sched, st);
}

}  // namespace swrt
