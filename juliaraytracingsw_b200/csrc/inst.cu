// Per-size explicit instantiation of the pass kernels (compiled once per SWRT_N so that
// `make -j` builds the sizes in parallel).
#include "models.cuh"

#ifndef SWRT_N
#error "compile with -DSWRT_N=<transform length>"
#endif

namespace swrt {

template <>
cudaError_t Launch<SWRT_N>::rsw_stage_a(const RswLoaderA& ld, const SpecLayout& L, double2* G_, const double2* tw, cudaStream_t st) {
    return ypass_inv(ld, L, 5, G_, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::rsw_stage_b(int modified, const double2* G_, double2* H, const SpecLayout& L, const double2* tw,
                                        cudaStream_t st) {
    const double s1 = 1.0 / ((double)L.nx * (double)L.ny);
    if (modified) {
        RswXOp<SWRT_N, true> op{G_, H, 0.5 * s1 * s1, s1};
        return xpass(op, L, tw, st);
    }
    RswXOp<SWRT_N, false> op{G_, H, 0.5 * s1 * s1, s1};
    return xpass(op, L, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::rsw_stage_c(const RswCombiner& cb, const SpecLayout& L, const double2* H, double2* Nout, const double2* tw,
                                        cudaStream_t st) {
    return ypass_fwd(cb, L, 3, H, Nout, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::field_stage_a(const FieldLoader& ld, const SpecLayout& L, double2* G_, const double2* tw, cudaStream_t st) {
    return ypass_inv(ld, L, 1, G_, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::field_stage_b(const double2* G_, double* out, const SpecLayout& L, const double2* tw, cudaStream_t st) {
    C2ROp<SWRT_N> op{G_, out, 1.0 / ((double)L.nx * (double)L.ny)};
    return xpass(op, L, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::psi_stage_a(const PsiLoader& ld, const SpecLayout& L, double2* G_, const double2* tw, cudaStream_t st) {
    return ypass_inv(ld, L, 3, G_, tw, st);
}
template <>
cudaError_t Launch<SWRT_N>::snap_stage_b(const double2* G_, double* out, const SpecLayout& L, const double2* tw, cudaStream_t st) {
    SnapshotXOp<SWRT_N> op{G_, out, 1.0 / ((double)L.nx * (double)L.ny)};
    return xpass(op, L, tw, st);
}

}  // namespace swrt
