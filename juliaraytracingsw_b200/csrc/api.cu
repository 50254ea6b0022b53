// libswrt C ABI: handle management, host-side tables, kernel dispatch.  See include/swrt.h.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/swrt.h"
#include "models.cuh"
#include "packets.cuh"
#include "team.cuh"
#include "update.cuh"

using namespace swrt;

namespace swrt {
extern template struct Launch<32>;
extern template struct Launch<64>;
extern template struct Launch<128>;
extern template struct Launch<256>;
extern template struct Launch<512>;
extern template struct Launch<1024>;
extern template struct Launch<2048>;
extern template struct Launch<4096>;
}  // namespace swrt

#define SWRT_DISPATCH(n, RES, ...)                                   \
    switch (n) {                                                     \
        case 32: { using LN = Launch<32>; RES = __VA_ARGS__; } break;   \
        case 64: { using LN = Launch<64>; RES = __VA_ARGS__; } break;   \
        case 128: { using LN = Launch<128>; RES = __VA_ARGS__; } break; \
        case 256: { using LN = Launch<256>; RES = __VA_ARGS__; } break; \
        case 512: { using LN = Launch<512>; RES = __VA_ARGS__; } break; \
        case 1024: { using LN = Launch<1024>; RES = __VA_ARGS__; } break; \
        case 2048: { using LN = Launch<2048>; RES = __VA_ARGS__; } break; \
        case 4096: { using LN = Launch<4096>; RES = __VA_ARGS__; } break; \
        default: RES = cudaErrorInvalidValue;                        \
    }

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail(SWRT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

struct swrt_flow {
    swrt_flow_desc d{};
    SpecLayout L{};
    int nkr = 0, nvar = 3, njobs_a = 5, njobs_b = 4;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr;
    cudaStream_t io_up = nullptr, io_down = nullptr;   // host transfers of own-stream packet handles, one stream per direction (io_route)
    double2 *sol = nullptr, *Nb[3] = {nullptr, nullptr, nullptr}, *G = nullptr, *H = nullptr, *stage = nullptr;
    double2* forcing = nullptr;             // vars.Fh of the forcing hook (swrt_flow_set_forcing), [l][kr_pad]; nullptr: none
    double2 *tw_x = nullptr, *tw_y = nullptr;
    double4* coef = nullptr;
    double2* psih = nullptr;                    // materialised streamfunction for the packet snapshot
    // spectrally refined snapshots ("FFT interpolation"): node grid refine x finer than the flow's, by zero padding
    int refine = 1;
    SpecLayout Ls{};
    double2 *psih_s = nullptr, *Gs = nullptr, *tw_xs = nullptr, *tw_ys = nullptr;
    double4* coef2 = nullptr;                   // ETDRK4 coefficients {zeta, alpha, beta, Gamma}
    double2 *S1 = nullptr, *S2 = nullptr, *N4 = nullptr;   // stage states and 4th N buffer of the multi-stage steppers
    double2 *Etab = nullptr, *E2tab = nullptr;   // tabulated exp(L dt), exp(2 L dt) for general NV x NV blocks (two-layer QG)
    double* snap[2] = {nullptr, nullptr};   // S[ny][nx][6] per time level (snapshot_layout.cuh)
    int slot_map[2] = {0, 1};               // slot (0 = old, 1 = new) -> array
    int interp = 0;                         // snapshot node data: 0 bilinear (5 fields / 48 B), 1 Hermite bicubic (7 fields / 64 B)
    int nufft_w = 8;                        // SWRT_INTERP_NUFFT: kernel width; ptab = 1 / phihat per kr index (nkr) then per l index (ny)
    double* ptab = nullptr;
    int ptab_w = 0, ptab_refine = 0;
    CUtensorMap tmapf[2];                   // the same arrays viewed as [ny][nx * 8] floats (fp32 packet mode), box = one fp32 patch
    CUtensorMap tmap[2];                    // TMA descriptors of the two levels viewed as [ny][nx * 6] doubles, box = one tile patch
    bool tmap_ok = false;
    double* phys = nullptr;
    double* red = nullptr;  // reduction scratch (device)
    double2 *G2 = nullptr, *H2 = nullptr;   // slab mode: A_RECV / B_SEND (G = A_SEND, H = B_RECV)
    int P = 1, rank = 0;
    bool own_stream = true;
    // peer receive buffers (cudaIpcOpenMemHandle) of the two transposes: [0] = A_RECV (G2), [1] = B_RECV (H) of every rank
    double2* peer[3][kMaxPeers] = {};   // [2] = A_SEND (G) of every rank, mapped for the pull variant of the first transpose
    bool p2p = false, pull = false;
    int slab_mode = 0;   // first transpose: 0 = the y-pass stores into the peers (32-64 B pieces), 1 = the x-pass pulls, 2 = local stores + block copy kernel
    int slab_b_copy = 0; // second transpose: 0 = the x-pass stores into the peers, 1 = local stores + block copy kernel
    unsigned* sched = nullptr;   // {next row, finished CTAs} of the dynamically scheduled x-pass (self re-arming)
    int ring = 0;
    double t = 0.0;
    long long step = 0, launches = 0;
    // optional per-kernel timing (CUDA events around every launch on the handle's stream)
    bool prof = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    const char* ray_name = "raytrace_rk4_cached_kernel<4>";   // the ray kernel the last raytrace call launched (profile label)
    double prof_ms[16] = {0};
    long long prof_n[16] = {0};
    std::vector<struct swrt_packets*> readers;   // packet handles with their own stream (snapshot writers wait for their reads)
    int npackets = 0;                            // every packet handle attached to this flow (own stream or not)
    // team mode (P > 1): band snapshots [2 levels][halo + yrows + halo][nx][6] in one IPC-shared allocation (snap[] point into it),
    // barrier flags, and the peers' mappings of both
    double* band = nullptr;
    int halo = 0;
    TeamFlags* flags = nullptr;
    TeamFlags* peerflags[kMaxPeers] = {};
    double* peerband[kMaxPeers] = {};
    unsigned long long epoch[2] = {0, 0};       // channel 0: the flow's stream, channel 1: packet collectives (may run on the packets' own stream)
    int barrier_mode = 0;                        // 0 = device flags over NVLink, 1 = host callback after a stream synchronise
    void (*barrier_cb)(void*) = nullptr;
    void* barrier_arg = nullptr;
    // CUDA graphs of the step for launch-bound grid sizes: one per ring phase (3 steps each; 1 step for the multi-stage steppers)
    cudaGraphExec_t gexec[3] = {nullptr, nullptr, nullptr};
    long long glaunches[3] = {0, 0, 0};
    bool no_graph = false;   // set while an outer capture (the coupled loop) records the step
};

enum { K_STAGE_A = 0, K_STAGE_B, K_STAGE_C, K_UPDATE, K_PSI_A, K_SNAP_B, K_RAYTRACE, K_SAMPLE, K_FIELD_A, K_FIELD_B, K_SORT, K_PSI, K_OTHER, K_COUNT };
static const char* kKernelNames[K_COUNT] = {"ypass_inv_kernel<RswLoaderA>", "xpass_kernel<RswXOp>", "ypass_fwd_kernel<RswCombiner>",
                                            "ifmab3_update_rsw_kernel", "ypass_inv_kernel<PsiLoader>", "xpass_kernel<SnapshotXOp>",
                                            "raytrace_rk4_kernel", "sample_kernel", "ypass_inv_kernel<FieldLoader>", "xpass_kernel<C2ROp>",
                                            "packet_sort_kernels", "psi_kernel", "other"};

static inline bool is_etd(int stepper) { return stepper == SWRT_ETDRK4 || stepper == SWRT_FILTEREDETDRK4; }

struct ProfScope {
    swrt_flow* h;
    int id;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    ProfScope(swrt_flow* h_, int id_, cudaStream_t st_ = nullptr) : h(h_), id(id_), st(st_ ? st_ : h_->st) {
        h->launches++;
        if (!h->prof) return;
        auto get = [&]() { cudaEvent_t e; if (h->pool.empty()) cudaEventCreate(&e); else { e = h->pool.back(); h->pool.pop_back(); } return e; };
        a = get(); b = get();
        cudaEventRecord(a, st);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, st);
        h->recs.push_back({id, a, b});
    }
};
static void prof_collect(swrt_flow* h) {
    for (auto& r : h->recs) {
        cudaEventSynchronize(r.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        h->prof_ms[r.id] += ms;
        h->prof_n[r.id] += 1;
        h->pool.push_back(r.a);
        h->pool.push_back(r.b);
    }
    h->recs.clear();
}

struct swrt_packets {
    swrt_packets_desc d{};
    swrt_flow* flow = nullptr;
    // state (sorted order) + alternates for the out-of-place sort; idx = original row of each packet
    double *xk = nullptr, *sign = nullptr, *xk2 = nullptr, *sign2 = nullptr, *U = nullptr, *Gd = nullptr;
    unsigned *idx = nullptr, *idx2 = nullptr, *keys = nullptr, *hist = nullptr, *sums = nullptr;
    unsigned* rsched = nullptr;             // {next tile, finished CTAs} of the persistent ray kernel (re-armed by its last CTA)
    unsigned long long* count = nullptr;
    long long nbins = 0;
    int kernel_sel = SWRT_RAYKERNEL_AUTO;   // swrt_packets_set_kernel
    // band mode (the flow is slab-decomposed over a team): one IPC-shared arena, capacity `cap` rows per column, `ncur` resident
    // packets; idx holds GLOBAL original rows; `first` = global row of this rank's caller-order block of d.n rows
    bool band = false;
    long long cap = 0, ncur = 0, first = 0;
    char* arena = nullptr;
    double* out6 = nullptr;                  // [6][cap]: sampler output in resident order / staging block of a scatter
    unsigned long long* tab = nullptr;       // team.cuh PacketArena::tab
    unsigned long long* tab_host = nullptr;  // pinned mirror (resident count, overflow flag, violations)
    char* peer_arena[kMaxPeers] = {};
    int cur = 0;                             // which half of the double buffers `xk` currently is (0: xk = A)
    bool tiles_valid = false;   // `hist` holds the per-key end offsets of the CURRENT packet order (set by the sort)
    int since_sort = 1 << 30;   // raytrace calls since the last sort
    bool permuted = false;
    cudaStream_t st = nullptr;      // the handle's own stream (swrt_packets_use_own_stream); otherwise pst() resolves the flow's
    bool own = false;
    cudaEvent_t ev_done = nullptr;  // last read of the flow's snapshots by this handle
    cudaEvent_t ev_io = nullptr;    // hand-over between the handle's stream and the flow's transfer streams
    // CUDA graphs of six coupled steps (ring period 3 x snapshot-slot period 2), one per parity of the sort's double buffer
    struct Cycle { cudaGraphExec_t exec = nullptr; long long launches = 0; const void *xk = nullptr, *snap0 = nullptr; int psi_kind = 0, interp = 0; double kcut = 0, k0 = 0; };
    Cycle cycle[2];
};
static inline cudaEvent_t packets_done_event(const swrt_packets* p) { return p->ev_done; }
// The stream a packet handle launches on: its own (swrt_packets_use_own_stream) or -- resolved at every use, because
// swrt_flow_set_stream may replace it after the handle was created -- the flow's.
static inline cudaStream_t pst(const swrt_packets* p) { return p->own ? p->st : p->flow->st; }

struct swrt_series {
    swrt_flow* flow = nullptr;
    int kind = 0, kr = 0, nser = 0;
    long long maxf = 0;
    double2 *buf = nullptr, *wts = nullptr;   // [series][l][maxf]; RSW: copy of the projection weights
    std::vector<double> t;
};

// ------------------------------------------------------------------ small kernels (api TU only)
// host (nkr, nl, nvar) column-major  <->  device [var][l][kr_pad], dealiased
__global__ void pack_sol_kernel(const double2* __restrict__ host_layout, double2* __restrict__ sol, SpecLayout L, int nkr, int nvar) {
    const long long total = (long long)nvar * L.ny * L.kr_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kr = (int)(i % L.kr_pad);
        const long long r = i / L.kr_pad;
        const int l = (int)(r % L.ny), v = (int)(r / L.ny);
        double2 val = make_double2(0.0, 0.0);
        if (kr < L.kr_keep && l_retained(L, l)) val = host_layout[((long long)v * L.ny + l) * nkr + L.kr_off + kr];
        sol[i] = val;
    }
}
__global__ void unpack_sol_kernel(const double2* __restrict__ sol, double2* __restrict__ host_layout, SpecLayout L, int nkr, int nvar) {
    const long long total = (long long)nvar * L.ny * nkr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kg = (int)(i % nkr), kr = kg - L.kr_off;    // slab mode: only this rank's columns are filled, the rest stay zero
        const long long r = i / nkr;
        const int l = (int)(r % L.ny), v = (int)(r / L.ny);
        double2 val = make_double2(0.0, 0.0);
        if (kr >= 0 && kr < L.kr_keep && l_retained(L, l)) val = sol[((long long)v * L.ny + l) * L.kr_pad + kr];
        host_layout[i] = val;
    }
}

// block partial reductions: mode 0 = parsevalsum2 weights of |a|^2, mode 1 = max |x| of a real array, mode 2 = NaN count
__global__ void __launch_bounds__(256) reduce_kernel(const double* __restrict__ a, long long n, int mode, SpecLayout L, double* __restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) {
            const int kr = (int)(i % L.kr_pad);
            const double2 v = reinterpret_cast<const double2*>(a)[i];
            const double w = (kr == 0 || kr == L.nx / 2) ? 1.0 : 2.0;
            acc += w * (v.x * v.x + v.y * v.y);
        } else if (mode == 1) {
            const double x = fabs(a[i]);
            acc = (x > acc || x != x) ? x : acc;      // NaN sticks (maximum(abs.(u)) of a blown-up field is NaN in the reference)
        } else {
            acc += isnan(a[i]) ? 1.0 : 0.0;
        }
    }
    // warp-shuffle tree, then the 8 warp results through shared memory (NaN sticks in the max: a0 + a1 is NaN if either is)
    auto comb = [mode](double a0, double a1) { return mode == 1 ? ((a0 != a0 || a1 != a1) ? a0 + a1 : (a1 > a0 ? a1 : a0)) : a0 + a1; };
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = comb(acc, __shfl_down_sync(0xffffffffu, acc, off));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) r = comb(r, __shfl_down_sync(0xffffffffu, r, off));
        if (threadIdx.x == 0) sh[0] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// destination table of a pass output: local array (single GPU), local send buffer split per destination (slab, NCCL exchange),
// or the peers' receive buffers (slab, direct NVLink stores)
static OutPeers out_local(double2* arr) {
    OutPeers o{};
    o.p[0] = arr;
    o.self = 0;
    return o;
}
// sources of the x-pass input segments: this rank's receive buffer [src][job][row][chunk], or (pull) every rank's own send
// buffer, where this rank's block sits at index `rank`
// block copy of the first transpose: the send buffer [dest][job][row][chunk] goes to every peer's receive buffer in full lines
__global__ void __launch_bounds__(256) slab_block_copy_kernel(const double2* __restrict__ send, OutPeers dst, long long blk, int P) {
    const long long total = blk * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i / blk);
        dst.p[d][(long long)dst.self * blk + (i - (long long)d * blk)] = __ldcs(send + i);
    }
}
static OutPeers in_slab(const swrt_flow* h, int njobs) {
    OutPeers o{};
    if (h->pull && h->slab_mode == 1) {
        for (int s = 0; s < h->P; ++s) o.p[s] = h->peer[2][s];
        o.self = h->rank;
    } else {
        for (int s = 0; s < h->P; ++s) o.p[s] = h->G2 + (long long)s * njobs * h->L.yrows * h->L.kr_pad;
        o.self = 0;
    }
    return o;
}
static OutPeers out_slab(const swrt_flow* h, int which /*0: A, 1: B*/, double2* send, int njobs) {
    OutPeers o{};
    if (h->p2p && !(which == 0 && h->slab_mode != 0) && !(which == 1 && h->slab_b_copy)) {
        for (int d = 0; d < h->P; ++d) o.p[d] = h->peer[which][d];
        o.self = h->rank;
    } else {   // send buffer laid out [dest][job][row][chunk]: block d starts at d * njobs * yrows * chunk
        for (int d = 0; d < h->P; ++d) o.p[d] = send + (long long)d * njobs * h->L.yrows * h->L.kr_pad;
        o.self = 0;
    }
    return o;
}

// ------------------------------------------------------------------ helpers
static bool supported_n(int n) { return n >= 32 && n <= 4096 && (n & (n - 1)) == 0; }

static void alias_ranges(int n, int nkr, double a, int* lo, int* hi, int* krlo) {
    // FourierFlows getaliasedwavenumbers; same float arithmetic as Julia
    if (a > 0) {
        const double Lf = (1 - a) / 2, Rf = (1 + a) / 2;
        const int iL = (int)std::floor(Lf * n) + 1, iR = (int)std::ceil(Rf * n);
        *lo = iL - 1;
        *hi = iR;
        *krlo = iL - 1;
    } else {
        *lo = n / 2;
        *hi = n / 2 + 1;
        *krlo = nkr - 1;
    }
}

static cudaError_t upload_twiddles(int n, double2** out) {
    std::vector<double2> tw(n);
    for (int m = 0; m < n; ++m) {
        // exact octant symmetries are not needed at 1e-16; use sincos of the reduced angle
        const double ang = -2.0 * M_PI * (double)m / (double)n;
        tw[m] = make_double2(std::cos(ang), std::sin(ang));
    }
    // exact values on the axes
    tw[0] = make_double2(1.0, 0.0);
    if (n >= 4) { tw[n / 4] = make_double2(0.0, -1.0); tw[n / 2] = make_double2(-1.0, 0.0); tw[3 * n / 4] = make_double2(0.0, 1.0); }
    cudaError_t e = cudaMalloc(out, sizeof(double2) * n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*out, tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice);
}

static double reduce_host(swrt_flow* h, const double* a, long long n, int mode, cudaError_t* err) {
    const int blocks = 296;
    { ProfScope ps(h, K_OTHER); reduce_kernel<<<blocks, 256, 0, h->st>>>(a, n, mode, h->L, h->red); }
    std::vector<double> part(blocks);
    *err = cudaMemcpyAsync(part.data(), h->red, sizeof(double) * blocks, cudaMemcpyDeviceToHost, h->st);
    if (*err != cudaSuccess) return 0;
    *err = cudaStreamSynchronize(h->st);
    double acc = 0;
    for (double p : part) acc = mode == 1 ? ((acc != acc || p != p) ? acc + p : std::max(acc, p)) : acc + p;   // NaN propagates
    return acc;
}

static int spectral_to_physical(swrt_flow* h, int which, double* dev_out) {
    if (h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "physical-space fields of a slab-decomposed flow are not gathered by the library");
    FieldLoader ld{h->sol, h->L.vs, which, h->nvar, h->d.f, h->L.aux0};
    cudaError_t e;
    { ProfScope ps(h, K_FIELD_A); SWRT_DISPATCH(h->L.ny, e, LN::field_stage_a(ld, h->L, out_local(h->G), h->tw_y, h->st)); }
    CK(e);
    { ProfScope ps(h, K_FIELD_B); SWRT_DISPATCH(h->L.nx, e, LN::field_stage_b(h->G, dev_out, h->L, h->tw_x, h->sched, h->st)); }
    CK(e);
    return SWRT_OK;
}

// TMA descriptors for the tile kernel of the ray tracer (packets.cuh): each level of the 5-field snapshot is a 2-D tensor of
// doubles [ny][nx * SNAP_STRIDE]; one box = the PATCH x PATCH node records around a sort tile.  cuTensorMapEncodeTiled is
// fetched through the runtime (no link-time dependency on libcuda).
typedef CUresult (*swrt_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void build_tmaps(swrt_flow* h) {
    h->tmap_ok = false;
    const long long nx = (long long)h->refine * h->d.nx;
    const long long ny = h->P > 1 ? (long long)(h->L.yrows + 2 * h->halo) : (long long)h->refine * h->d.ny;   // rows the arrays hold
    if (nx < PATCH || ny < PATCH) return;
    static swrt_encode_tiled_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
        return (swrt_encode_tiled_fn)fn;
    }();
    if (!encode) return;
    for (int lev = 0; lev < 2; ++lev) {
        const cuuint64_t dims[2] = {(cuuint64_t)(nx * SNAP_STRIDE), (cuuint64_t)ny};
        const cuuint64_t strides[1] = {(cuuint64_t)(nx * SNAP_STRIDE * sizeof(double))};
        const cuuint32_t box[2] = {(cuuint32_t)PATCH_ROW, (cuuint32_t)PATCH}, estr[2] = {1, 1};
        if (encode(&h->tmap[lev], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->snap[lev], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
        const cuuint64_t dimsf[2] = {(cuuint64_t)(nx * SNAPF_STRIDE), (cuuint64_t)ny};
        const cuuint64_t stridesf[1] = {(cuuint64_t)(nx * SNAPF_STRIDE * sizeof(float))};
        const cuuint32_t boxf[2] = {(cuuint32_t)PATCHF_ROW, (cuuint32_t)PATCH};
        if (encode(&h->tmapf[lev], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->snap[lev], dimsf, stridesf, boxf, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
    }
    h->tmap_ok = true;
}

// NUFFT mode: 1 / phihat(xi) of the sampling kernel phi(z) = exp(beta (sqrt(1 - z^2) - 1)) on |t| <= w/2 (t in nodes of the
// oversampled grid, z = 2 t / w):  phihat(xi) = int phi cos(xi t) dt = (w/2) int_{-pi/2}^{pi/2} exp(beta (cos th - 1)) cos(xi w sin(th) / 2) cos th dth
// (smooth after z = sin th; composite Simpson with 4096 panels is exact to rounding).
static double nufft_beta(int w) { return 2.30 * w; }
static double nufft_phihat(double xi, int w) {
    const int n = 4096;
    const double beta = nufft_beta(w), a = -0.5 * M_PI, hh = M_PI / n;
    auto f = [&](double th) { return std::exp(beta * (std::cos(th) - 1.0)) * std::cos(0.5 * xi * w * std::sin(th)) * std::cos(th); };
    double acc = f(a) + f(a + M_PI);
    for (int i = 1; i < n; ++i) acc += (i & 1 ? 4.0 : 2.0) * f(a + i * hh);
    return 0.5 * w * acc * hh / 3.0;
}
static int nufft_tables(swrt_flow* h) {
    if (h->ptab && h->ptab_w == h->nufft_w && h->ptab_refine == h->refine) return SWRT_OK;
    const int nkr = h->nkr, ny = h->d.ny, w = h->nufft_w;
    std::vector<double> t((size_t)nkr + ny);
    for (int k = 0; k < nkr; ++k) t[k] = 1.0 / nufft_phihat(2.0 * M_PI * k / ((double)h->refine * h->d.nx), w);
    for (int l = 0; l < ny; ++l) t[nkr + l] = 1.0 / nufft_phihat(2.0 * M_PI * (l < ny / 2 ? l : l - ny) / ((double)h->refine * ny), w);
    if (!h->ptab) CK(cudaMalloc(&h->ptab, sizeof(double) * t.size()));
    CK(cudaMemcpy(h->ptab, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice));
    h->ptab_w = w;
    h->ptab_refine = h->refine;
    return SWRT_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" {
#pragma GCC visibility push(default)

const char* swrt_last_error(void) { return g_err.c_str(); }
int swrt_version(void) { return SWRT_VERSION; }
int swrt_device_count(int* n) {
    if (!n) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaGetDeviceCount(n));
    return SWRT_OK;
}

int swrt_flow_destroy(swrt_flow* h) {
    if (!h) return SWRT_OK;
    cudaSetDevice(h->d.device);
    if (h->st) cudaStreamSynchronize(h->st);
    cudaFree(h->sol);
    cudaFree(h->forcing);
    for (auto p : h->Nb) cudaFree(p);
    cudaFree(h->G); cudaFree(h->H); cudaFree(h->stage); cudaFree(h->tw_x); cudaFree(h->tw_y); cudaFree(h->coef); cudaFree(h->Etab); cudaFree(h->E2tab); cudaFree(h->coef2); cudaFree(h->psih); cudaFree(h->psih_s); cudaFree(h->Gs); cudaFree(h->tw_xs); cudaFree(h->tw_ys); cudaFree(h->S1); cudaFree(h->S2); cudaFree(h->N4);
    if (h->band) cudaFree(h->band); else { cudaFree(h->snap[0]); cudaFree(h->snap[1]); }
    cudaFree(h->flags); cudaFree(h->ptab);
    cudaFree(h->phys); cudaFree(h->red); cudaFree(h->sched); cudaFree(h->G2); cudaFree(h->H2);
    for (int w = 0; w < 3; ++w)
        for (int r = 0; r < h->P; ++r)
            if (h->peer[w][r] && r != h->rank) cudaIpcCloseMemHandle(h->peer[w][r]);
    for (int r = 0; r < h->P; ++r) {
        if (h->peerflags[r] && r != h->rank) cudaIpcCloseMemHandle(h->peerflags[r]);
        if (h->peerband[r] && r != h->rank) cudaIpcCloseMemHandle(h->peerband[r]);
    }
    prof_collect(h);
    for (auto e : h->pool) cudaEventDestroy(e);
    for (auto& g : h->gexec) if (g) cudaGraphExecDestroy(g);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev_sync) cudaEventDestroy(h->ev_sync);
    if (h->io_up) cudaStreamDestroy(h->io_up);
    if (h->io_down) cudaStreamDestroy(h->io_down);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->st && h->own_stream) cudaStreamDestroy(h->st);
    delete h;
    return SWRT_OK;
}

int swrt_flow_create(const swrt_flow_desc* desc, swrt_flow** out) {
    if (!desc || !out) return fail(SWRT_ERR_ARG, "null pointer");
    *out = nullptr;
    const swrt_flow_desc& d = *desc;
    if (!supported_n(d.nx) || !supported_n(d.ny)) return fail(SWRT_ERR_UNSUPPORTED, "nx, ny must be powers of two in [32, 4096] (got %d x %d)", d.nx, d.ny);
    const bool rsw_family = d.model == SWRT_RSW || d.model == SWRT_RSW_MODIFIED || d.model == SWRT_RSW_LINDBORG || d.model == SWRT_RSW_QUADHEIGHT;
    const bool diag_L = d.model == SWRT_SWQG || d.model == SWRT_THOMASYAMADA || d.model == SWRT_MULTILAYERQG2;
    if (!rsw_family && !diag_L && d.model != SWRT_TWOLAYERQG) return fail(SWRT_ERR_UNSUPPORTED, "model %d not implemented", d.model);
    if (d.stepper < SWRT_IFMAB3 || d.stepper > SWRT_FILTEREDETDRK4) return fail(SWRT_ERR_ARG, "unknown stepper %d", d.stepper);
    if (d.stepper != SWRT_IFMAB3 && !diag_L)
        return fail(SWRT_ERR_UNSUPPORTED, "stepper %d needs a diagonal L (FourierFlows applies L .* sol); model %d has matrix blocks", d.stepper, d.model);
    if (!(d.Lx > 0 && d.Ly > 0 && d.dt > 0)) return fail(SWRT_ERR_ARG, "Lx, Ly, dt must be positive");
    if (!(d.aliased_fraction >= 0 && d.aliased_fraction < 1)) return fail(SWRT_ERR_ARG, "aliased_fraction must be in [0,1)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SWRT_ERR_CUDA, "no CUDA device available (libswrt has no CPU fallback)");
    if (d.device < 0 || d.device >= ndev) return fail(SWRT_ERR_ARG, "device %d out of range (%d devices)", d.device, ndev);
    CK(cudaSetDevice(d.device));

    swrt_flow* h = new swrt_flow;
    h->d = d;
    h->nkr = d.nx / 2 + 1;
    h->nvar = model_nvar(d.model);
    h->njobs_a = model_njobs_a(d.model) > 3 ? model_njobs_a(d.model) : 3;   // the snapshot pass needs 3 slots of G
    if (d.slab_size > 1) h->njobs_a = model_njobs_a(d.model) + 3;           // team mode: the snapshot's psi jobs can ride behind the model's
    h->njobs_b = model_njobs_b(d.model);
    SpecLayout& L = h->L;
    L.nx = d.nx; L.ny = d.ny;
    int klo, dummy0, dummy1;
    alias_ranges(d.ny, h->nkr, d.aliased_fraction, &L.lz0, &L.lz1, &dummy0);
    alias_ranges(d.nx, h->nkr, d.aliased_fraction, &dummy0, &dummy1, &klo);
    L.kr_keep = klo;
    L.kr_pad = (L.kr_keep + 15) / 16 * 16;
    L.vs = (long long)L.ny * L.kr_pad;
    L.kr_off = 0; L.kr_keep_g = L.kr_keep; L.yrows = L.ny;
    L.yshift = 0; while ((1 << L.yshift) < L.ny) ++L.yshift;
    if (d.slab_size > 1) {
        const int P = d.slab_size;
        if ((P & (P - 1)) || P > 16 || d.slab_rank < 0 || d.slab_rank >= P || d.ny % P || d.ny / P < 16) { delete h; return fail(SWRT_ERR_ARG, "slab_size must be a power of two <= 16 dividing ny (>= 16 rows per rank), 0 <= slab_rank < slab_size"); }
        // (every model; with IFMAB3 / FilteredAB3 the slab step is stage A -> B -> C + update, with the multi-stage steppers every
        //  calcN! of a stage runs the three slab passes: slab_compute_N)
        h->P = P; h->rank = d.slab_rank;
        const int chunk = ((L.kr_keep_g + P - 1) / P + 15) / 16 * 16;
        L.kr_off = d.slab_rank * chunk;
        int mine = L.kr_keep_g - L.kr_off;
        L.kr_keep = mine < 0 ? 0 : (mine > chunk ? chunk : mine);
        L.kr_pad = chunk;
        L.vs = (long long)L.ny * chunk;
        L.yrows = d.ny / P;
        L.yshift = 0; while ((1 << L.yshift) < L.yrows) ++L.yshift;
    }
    L.dk = 2.0 * M_PI / d.Lx; L.dl = 2.0 * M_PI / d.Ly;
    L.f = d.f; L.Cg2 = d.Cg * d.Cg;
    // model constant used by the loaders: Kd2 = f^2/Cg^2 (SWQG, swqg/SWQG.jl:85; RSW balanced psi) or F (two-layer, swqg/TwoLayerQG.jl:79)
    L.aux0 = (d.model == SWRT_TWOLAYERQG || d.model == SWRT_MULTILAYERQG2) ? d.F : (d.Kd2 > 0 ? d.Kd2 : d.f * d.f / L.Cg2);
    L.aux1 = d.Ro;   // Thomas-Yamada Rossby number
    L.aux2 = d.U; L.aux3 = d.U2; L.aux4 = d.beta; L.aux5 = d.mu;   // MultiLayerQG-2 mean flow, beta, bottom drag

    auto bail = [&](int code) { swrt_flow_destroy(h); return code; };
#define CKB(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { fail(SWRT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); return bail(SWRT_ERR_CUDA); } } while (0)
    {   // the flow's stream gets the greatest priority: its short, latency-bound kernels then take SM slots ahead of the CTAs of a
        // long ray-tracing kernel that runs beside it on a packet handle's own stream (team mode, PacketPipeline)
        int lo = 0, hi = 0;
        CKB(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CKB(cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, hi));
    }
    CKB(cudaEventCreate(&h->ev0));
    CKB(cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming));
    CKB(cudaEventCreate(&h->ev1));
    const size_t fb = sizeof(double2) * (size_t)L.vs;
    CKB(cudaMalloc(&h->sol, fb * h->nvar));
    CKB(cudaMemset(h->sol, 0, fb * h->nvar));
    for (int i = 0; i < 3; ++i) { CKB(cudaMalloc(&h->Nb[i], fb * h->nvar)); CKB(cudaMemset(h->Nb[i], 0, fb * h->nvar)); }
    if (is_etd(d.stepper) || d.stepper == SWRT_FILTEREDRK4) {
        CKB(cudaMalloc(&h->S1, fb * h->nvar)); CKB(cudaMemset(h->S1, 0, fb * h->nvar));
        CKB(cudaMalloc(&h->S2, fb * h->nvar)); CKB(cudaMemset(h->S2, 0, fb * h->nvar));
        CKB(cudaMalloc(&h->N4, fb * h->nvar)); CKB(cudaMemset(h->N4, 0, fb * h->nvar));
    }
    CKB(cudaMalloc(&h->G, fb * h->njobs_a)); CKB(cudaMemset(h->G, 0, fb * h->njobs_a));
    CKB(cudaMalloc(&h->H, fb * h->njobs_b)); CKB(cudaMemset(h->H, 0, fb * h->njobs_b));
    if (h->P > 1) {
        CKB(cudaMalloc(&h->G2, fb * h->njobs_a)); CKB(cudaMemset(h->G2, 0, fb * h->njobs_a));
        CKB(cudaMalloc(&h->H2, fb * h->njobs_b)); CKB(cudaMemset(h->H2, 0, fb * h->njobs_b));
    }
    CKB(cudaMalloc(&h->psih, fb)); CKB(cudaMemset(h->psih, 0, fb));
    CKB(cudaMalloc(&h->stage, sizeof(double2) * (size_t)h->nkr * d.ny * h->nvar));
    CKB(cudaMalloc(&h->phys, sizeof(double) * (size_t)d.nx * d.ny));
    CKB(cudaMalloc(&h->red, sizeof(double) * 1024));
    CKB(cudaMalloc(&h->sched, sizeof(unsigned) * 4)); CKB(cudaMemset(h->sched, 0, sizeof(unsigned) * 4));
    if (h->P > 1) {
        // team mode: a rank only ever samples its own band of rows (packets are sharded by y-band) plus `halo` rows of each neighbour
        h->halo = L.yrows < 8 ? L.yrows : 8;
        const size_t lev_doubles = (size_t)(L.yrows + 2 * h->halo) * d.nx * SNAP_STRIDE;
        CKB(cudaMalloc(&h->band, sizeof(double) * 2 * lev_doubles));
        CKB(cudaMemset(h->band, 0, sizeof(double) * 2 * lev_doubles));
        h->snap[0] = h->band;
        h->snap[1] = h->band + lev_doubles;
        CKB(cudaMalloc(&h->flags, sizeof(TeamFlags)));
        CKB(cudaMemset(h->flags, 0, sizeof(TeamFlags)));
    } else
    for (int lev = 0; lev < 2; ++lev) {
        CKB(cudaMalloc(&h->snap[lev], sizeof(double) * (size_t)d.nx * d.ny * SNAP3_STRIDE));   // sized for either node record
        CKB(cudaMemset(h->snap[lev], 0, sizeof(double) * (size_t)d.nx * d.ny * SNAP3_STRIDE));
    }
    build_tmaps(h);
    CKB(upload_twiddles(d.nx, &h->tw_x));
    CKB(upload_twiddles(d.ny, &h->tw_y));

    // coefficient table {e^{D dt}, sin(w dt)/w | D, (1-cos(w dt))/w^2, filter}; two-layer QG: tabulated 2x2 exponentials
    {
        std::vector<double4> cf((size_t)L.vs, make_double4(1.0, 0.0, 0.0, 1.0));
        std::vector<double2> E, E2;
        std::vector<double4> cf2;
        if (is_etd(d.stepper)) cf2.assign((size_t)L.vs, make_double4(0, 0, 0, 0));
        if (d.model == SWRT_TWOLAYERQG) { E.assign((size_t)4 * L.vs, make_double2(0, 0)); E2 = E; }
        const double w2c = (d.model == SWRT_RSW_MODIFIED || d.model == SWRT_RSW_QUADHEIGHT) ? 0.0 : L.Cg2;
        const double innerK = d.filter_innerK > 0 ? d.filter_innerK : 2.0 / 3.0, outerK = d.filter_outerK > 0 ? d.filter_outerK : 1.0;
        const double tol = d.filter_tol > 0 ? d.filter_tol : 1e-15;
        const int order = d.filter_order > 0 ? d.filter_order : 4;
        const double decay = -std::log(tol) / std::pow(outerK - innerK, order);
        const double dx = d.Lx / d.nx, dy = d.Ly / d.ny;
        using cplx = std::complex<double>;
        auto expm2 = [](cplx a, cplx b, cplx c, cplx dd, cplx* o) {   // e^s [cosh q I + sinh(q)/q (A - s I)]  
            const cplx s = 0.5 * (a + dd), q = std::sqrt(0.25 * (a - dd) * (a - dd) + b * c);
            const cplx sh = std::abs(q) < 1e-8 ? 1.0 + q * q / 6.0 : std::sinh(q) / q, ch = std::cosh(q), es = std::exp(s);
            o[0] = es * (ch + sh * (a - s)); o[1] = es * sh * b; o[2] = es * sh * c; o[3] = es * (ch + sh * (dd - s));
        };
        for (int l = 0; l < d.ny; ++l) {
            const double lw = (double)(l < d.ny / 2 ? l : l - d.ny) * L.dl;
            for (int kr = 0; kr < L.kr_keep; ++kr) {
                const double kw = (L.kr_off + kr) * L.dk, K2 = kw * kw + lw * lw;
                const double D = -d.nu * std::pow(K2, (double)d.nnu);
                const size_t off = (size_t)l * L.kr_pad + kr;
                double filt = 1.0;
                if (d.use_filter || d.stepper == SWRT_FILTEREDAB3 || d.stepper == SWRT_FILTEREDRK4 || d.stepper == SWRT_FILTEREDETDRK4) {
                    const double Kn = std::sqrt((kw * dx / M_PI) * (kw * dx / M_PI) + (lw * dy / M_PI) * (lw * dy / M_PI));
                    if (Kn >= innerK) filt = std::exp(-decay * std::pow(Kn - innerK, order));
                }
                if (rsw_family) {
                    const double w2 = d.f * d.f + w2c * K2, w = std::sqrt(w2), th = w * d.dt;
                    double s, c;
                    if (w > 0) { s = std::sin(th) / w; const double sh = std::sin(0.5 * th); c = 2.0 * sh * sh / w2; }
                    else { s = d.dt; c = 0.5 * d.dt * d.dt; }
                    cf[off] = make_double4(std::exp(D * d.dt), s, c, filt);
                } else if (diag_L) {
                    cf[off] = make_double4(std::exp(D * d.dt), D, std::exp(0.5 * D * d.dt), filt);
                    if (is_etd(d.stepper)) {   // FourierFlows getetdcoeffs: 32-point contour mean around dt L
                        cplx z(0), a(0), b(0), g(0);
                        for (int j = 0; j < 32; ++j) {
                            const cplx zc = D * d.dt + std::exp(cplx(0.0, 2.0 * M_PI / 32 * (j + 0.5)));
                            const cplx ez = std::exp(zc), z3 = zc * zc * zc;
                            z += (std::exp(0.5 * zc) - 1.0) / zc;
                            a += (-4.0 - zc + ez * (4.0 - 3.0 * zc + zc * zc)) / z3;
                            b += (2.0 + zc + ez * (-2.0 + zc)) / z3;
                            g += (-4.0 - 3.0 * zc - zc * zc + ez * (4.0 - zc)) / z3;
                        }
                        cf2[off] = make_double4(d.dt * z.real() / 32, d.dt * a.real() / 32, d.dt * b.real() / 32, d.dt * g.real() / 32);
                    }
                } else {   // two-layer QG, swqg/TwoLayerQG.jl:184-198 (evaluated in double; the reference's Float32 temporaries are a bug)
                    const double F = d.F, U = d.U, K2inv = K2 > 0 ? 1.0 / K2 : 0.0;
                    const cplx p0(0.0, -2.0 * kw * F * U), p1 = cplx(0.0, 2.0 * kw * F * U) + d.mu * K2;
                    const double sc = 1.0 / (K2 + 2 * F) * K2inv;
                    const double S00 = (-K2 - F) * sc, S01 = -F * sc;
                    cplx a = p0 * S00 + cplx(D, -kw * U), b = p0 * S01, c = p1 * S01, dd = p1 * S00 + cplx(D, kw * U);
                    cplx o[4];
                    expm2(a * d.dt, b * d.dt, c * d.dt, dd * d.dt, o);
                    for (int q = 0; q < 4; ++q) E[(size_t)q * L.vs + off] = make_double2(o[q].real(), o[q].imag());
                    expm2(a * (2 * d.dt), b * (2 * d.dt), c * (2 * d.dt), dd * (2 * d.dt), o);
                    for (int q = 0; q < 4; ++q) E2[(size_t)q * L.vs + off] = make_double2(o[q].real(), o[q].imag());
                    cf[off] = make_double4(1.0, 0.0, 0.0, filt);
                }
            }
        }
        CKB(cudaMalloc(&h->coef, sizeof(double4) * cf.size()));
        CKB(cudaMemcpy(h->coef, cf.data(), sizeof(double4) * cf.size(), cudaMemcpyHostToDevice));
        if (!cf2.empty()) {
            CKB(cudaMalloc(&h->coef2, sizeof(double4) * cf2.size()));
            CKB(cudaMemcpy(h->coef2, cf2.data(), sizeof(double4) * cf2.size(), cudaMemcpyHostToDevice));
        }
        if (!E.empty()) {
            CKB(cudaMalloc(&h->Etab, sizeof(double2) * E.size()));
            CKB(cudaMemcpy(h->Etab, E.data(), sizeof(double2) * E.size(), cudaMemcpyHostToDevice));
            CKB(cudaMalloc(&h->E2tab, sizeof(double2) * E2.size()));
            CKB(cudaMemcpy(h->E2tab, E2.data(), sizeof(double2) * E2.size(), cudaMemcpyHostToDevice));
        }
    }
#undef CKB
    *out = h;
    return SWRT_OK;
}

int swrt_flow_set_solution(swrt_flow* h, const void* sol_host) {
    if (!h || !sol_host) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    const size_t bytes = sizeof(double2) * (size_t)h->nkr * h->d.ny * h->nvar;
    CK(cudaMemcpyAsync(h->stage, sol_host, bytes, cudaMemcpyHostToDevice, h->st));
    { ProfScope ps(h, K_OTHER); pack_sol_kernel<<<592, 256, 0, h->st>>>(h->stage, h->sol, h->L, h->nkr, h->nvar); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

// addforcing! / calcF! (rsw/RotatingShallowWater.jl:228-240 and the same lines of the Modified / QuadHeight / Lindborg variants):
// the user's calcF! fills vars.Fh (nkr, nl) and calcN! ends with `@. N += vars.Fh`.  Here the host hands over Fh whenever its calcF!
// has produced a new one (between steps); the field stays on the device and is added at every calcN! evaluation until it is
// replaced or cleared (NULL).  A flow with forcing steps un-captured (the replayed CUDA graphs of the small grids bake their
// arguments in).
int swrt_flow_set_forcing(swrt_flow* h, const void* Fh_host) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    const int m = h->d.model;
    if (!(m == SWRT_RSW || m == SWRT_RSW_MODIFIED || m == SWRT_RSW_LINDBORG || m == SWRT_RSW_QUADHEIGHT))
        return fail(SWRT_ERR_UNSUPPORTED, "the forcing hook exists in the (u, v, eta) models only (model %d: its calcN! never calls addforcing!)", m);
    CK(cudaSetDevice(h->d.device));
    CK(cudaStreamSynchronize(h->st));
    for (auto& g : h->gexec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (!Fh_host) {
        cudaFree(h->forcing);
        h->forcing = nullptr;
        return SWRT_OK;
    }
    if (!h->forcing) CK(cudaMalloc(&h->forcing, sizeof(double2) * (size_t)h->L.vs));
    CK(cudaMemcpyAsync(h->stage, Fh_host, sizeof(double2) * (size_t)h->nkr * h->d.ny, cudaMemcpyHostToDevice, h->st));
    { ProfScope ps(h, K_OTHER); pack_sol_kernel<<<592, 256, 0, h->st>>>(h->stage, h->forcing, h->L, h->nkr, 1); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

int swrt_flow_get_solution(swrt_flow* h, void* sol_host) {
    if (!h || !sol_host) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    { ProfScope ps(h, K_OTHER); unpack_sol_kernel<<<592, 256, 0, h->st>>>(h->sol, h->stage, h->L, h->nkr, h->nvar); }
    CK(cudaGetLastError());
    const size_t bytes = sizeof(double2) * (size_t)h->nkr * h->d.ny * h->nvar;
    CK(cudaMemcpyAsync(sol_host, h->stage, bytes, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

int swrt_flow_set_rsw_initial_condition(swrt_flow* h, const double* phase_host, const double* sgn_host, double Kg0, double Kg1, double ag,
                                        double Kw0, double Kw1, double aw, double* scales_out) {
    // set_initial_condition! (rsw/RSWRaytracingDriver.jl:15-54) with the mode arithmetic, both normalising inverse transforms
    // and both maxima on the device; only the random numbers come from the host (Julia's / NumPy's stream)
    if (!h || !phase_host || !sgn_host) return fail(SWRT_ERR_ARG, "null pointer");
    const bool rsw = h->d.model == SWRT_RSW || h->d.model == SWRT_RSW_MODIFIED || h->d.model == SWRT_RSW_LINDBORG;
    if (!rsw) return fail(SWRT_ERR_UNSUPPORTED, "the RSW initial condition needs an (u, v, eta) model (model %d)", h->d.model);
    if (h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "not available for a slab-decomposed flow (build it on one GPU and hand the state over)");
    if (!(ag > 0 && aw >= 0 && Kg1 >= Kg0 && Kw1 >= Kw0)) return fail(SWRT_ERR_ARG, "bad band / amplitude");
    const SpecLayout& L = h->L;
    const double kmax_keep = (L.kr_keep - 1) * L.dk;
    if (Kg1 > kmax_keep || Kw1 > kmax_keep) return fail(SWRT_ERR_ARG, "the bands must lie inside the retained (dealiased) wavenumbers");
    CK(cudaSetDevice(h->d.device));
    const size_t nm = (size_t)h->nkr * h->d.ny;
    std::vector<double2> rnd(nm);
    for (size_t i = 0; i < nm; ++i) rnd[i] = make_double2(phase_host[i], sgn_host[i]);
    CK(cudaMemcpyAsync(h->stage, rnd.data(), sizeof(double2) * nm, cudaMemcpyHostToDevice, h->st));
    const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
    const unsigned blocks = (unsigned)((nmodes + 255) / 256);
    double scale[2] = {1.0, 1.0};
    const double amp[2] = {ag, aw};
    for (int part = 0; part < 2; ++part) {
        if (amp[part] == 0.0) { scale[part] = 0.0; continue; }
        { ProfScope ps(h, K_OTHER); rsw_ic_kernel<<<blocks, 256, 0, h->st>>>(h->stage, h->nkr, L, part, h->d.f, L.Cg2, Kg0, Kg1, Kw0, Kw1, 1.0, 1.0, h->sol); }
        CK(cudaGetLastError());
        int rc = spectral_to_physical(h, 0, h->phys);
        if (rc) return rc;
        cudaError_t e = cudaSuccess;
        const double m = reduce_host(h, h->phys, (long long)h->d.nx * h->d.ny, 1, &e);
        CK(e);
        if (!(m > 0)) return fail(SWRT_ERR_ARG, "band %d holds no mode (max|u| = %g)", part, m);
        scale[part] = amp[part] / m;
    }
    { ProfScope ps(h, K_OTHER); rsw_ic_kernel<<<blocks, 256, 0, h->st>>>(h->stage, h->nkr, L, 2, h->d.f, L.Cg2, Kg0, Kg1, Kw0, Kw1, scale[0], scale[1], h->sol); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->st));
    if (scales_out) { scales_out[0] = scale[0]; scales_out[1] = scale[1]; }
    return SWRT_OK;
}

int swrt_flow_enforce_reality(swrt_flow* h) {
    // rsw/RotatingShallowWater.jl:118-133: dealias!(sol) (already an invariant here); the round-tripped
    // fields go to vars.*h, which calcN! overwrites from sol on the next step -- sol is otherwise untouched.
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    return SWRT_OK;
}

// IFMAB3 / FilteredAB3 update of the state from the freshly computed N (history = the two other ring buffers)
static int ifmab3_update_launch(swrt_flow* h, double2* Ncur) {
    const SpecLayout& L = h->L;
    const int model = h->d.model, stepper = h->d.stepper;
    const bool modified = model == SWRT_RSW_MODIFIED || model == SWRT_RSW_QUADHEIGHT;
    RswLin lin{h->d.f, modified ? 0.0 : L.Cg2, modified ? 0.0 : L.Cg2, model == SWRT_RSW_QUADHEIGHT ? 0.0 : 1.0};
    const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
    if (nmodes == 0) return SWRT_OK;
    const int ublocks = (int)((nmodes + 255) / 256);
    const double2* Nm1 = h->Nb[(h->ring + 2) % 3];
    const double2* Nm2 = h->Nb[(h->ring + 1) % 3];
    UpdateArgs ua{h->sol, Ncur, Nm1, Nm2, h->coef, h->d.dt, h->step < 3 ? 1 : 0};
    {
        ProfScope ps(h, K_UPDATE);
        if (model == SWRT_SWQG) {
            if (stepper == SWRT_FILTEREDAB3) update_diag_kernel<1, true><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
            else update_diag_kernel<1, false><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
        } else if (model == SWRT_MULTILAYERQG2) {
            if (stepper == SWRT_FILTEREDAB3) update_diag_kernel<2, true><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
            else update_diag_kernel<2, false><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
        } else if (model == SWRT_THOMASYAMADA) {
            if (stepper == SWRT_FILTEREDAB3) update_diag_kernel<4, true><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
            else update_diag_kernel<4, false><<<ublocks, 256, 0, h->st>>>(ua, Ncur, L);
        } else if (model == SWRT_TWOLAYERQG) {
            ifmab3_update_table_kernel<2><<<ublocks, 256, 0, h->st>>>(ua, h->Etab, h->E2tab, L);
        } else {
            ifmab3_update_rsw_kernel<<<ublocks, 256, 0, h->st>>>(ua, lin, L);
        }
    }
    CK(cudaGetLastError());
    return SWRT_OK;
}

static int slab_compute_N(swrt_flow* h, const double2* state, double2* Nout);
// N = calcN!(state): the three transform passes of the model
static int compute_N(swrt_flow* h, const double2* state, double2* Nout) {
    if (h->P > 1) return slab_compute_N(h, state, Nout);
    const SpecLayout& L = h->L;
    const int model = h->d.model;
    cudaError_t e;
    { ProfScope ps(h, K_STAGE_A); SWRT_DISPATCH(L.ny, e, LN::stage_a(model, state, out_local(h->G), L, h->tw_y, h->st)); }
    CK(e);
    { ProfScope ps(h, K_STAGE_B); SWRT_DISPATCH(L.nx, e, LN::stage_b(model, h->G, h->H, L, h->tw_x, h->sched, h->st)); }
    CK(e);
    SpecLayout Lc = L;
    Lc.forcing = h->forcing;                 // addforcing!: the last thing calcN! does
    { ProfScope ps(h, K_STAGE_C); SWRT_DISPATCH(L.ny, e, LN::stage_c(model, state, h->H, Nout, Lc, h->tw_y, h->st)); }
    CK(e);
    return SWRT_OK;
}

static int flow_step_impl(swrt_flow* h, int nsteps);
int swrt_flow_step(swrt_flow* h, int nsteps) {
    if (!h || nsteps < 0) return fail(SWRT_ERR_ARG, "bad argument");
    if (h->P > 1) return fail(SWRT_ERR_STATE, "slab-decomposed flow: step it with swrt_slab_step (or drive swrt_slab_stage_a/b/c and the two all-to-alls)");
    return flow_step_impl(h, nsteps);
}
// (also the step of a slab-decomposed flow with a multi-stage stepper: compute_N then runs the slab passes and their barriers)
static int flow_step_impl(swrt_flow* h, int nsteps) {
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    const int stepper = h->d.stepper;
    const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
    const int ublocks = (int)((nmodes + 255) / 256);
    const double dt = h->d.dt;
    auto stage = [&](int mode, double2* out, const double2* x, double2* n1, const double2* n2, const double2* n3, const double2* n4,
                     const double2* xs, double c) {
        if (ublocks == 0) return cudaSuccess;          // a slab rank beyond the retained columns owns no modes
        StageArgs sa{out, x, n1, n2, n3, n4, xs, h->coef, h->coef2, c, mode, h->nvar};
        ProfScope ps(h, K_UPDATE);
        diag_stage_kernel<<<ublocks, 256, 0, h->st>>>(sa, L);
        return cudaGetLastError();
    };
    auto step_body = [&]() -> int {
        int rc;
        if (is_etd(stepper)) {            // FourierFlows ETDRK4 / FilteredETDRK4 stepforward! (the filter rides in the update stage)
            double2 *N1 = h->Nb[0], *N2 = h->Nb[1], *N3 = h->Nb[2], *N4 = h->N4;
            if ((rc = compute_N(h, h->sol, N1))) return rc;
            CK(stage(ST_ETD_SUB12, h->S1, h->sol, N1, nullptr, nullptr, nullptr, nullptr, 0));
            if ((rc = compute_N(h, h->S1, N2))) return rc;
            CK(stage(ST_ETD_SUB12, h->S2, h->sol, N2, nullptr, nullptr, nullptr, nullptr, 0));
            if ((rc = compute_N(h, h->S2, N3))) return rc;
            CK(stage(ST_ETD_SUB3, h->S2, h->S1, N1, nullptr, N3, nullptr, nullptr, 0));
            if ((rc = compute_N(h, h->S2, N4))) return rc;
            CK(stage(ST_ETD_UPDATE, h->sol, h->sol, N1, N2, N3, N4, nullptr, 0));
        } else if (stepper == SWRT_FILTEREDRK4) {   // FourierFlows (Filtered)RK4: RHS = N + L .* state
            double2 *R1 = h->Nb[0], *R2 = h->Nb[1], *R3 = h->Nb[2], *R4 = h->N4;
            if ((rc = compute_N(h, h->sol, R1))) return rc;
            CK(stage(ST_RK4_STAGE, h->S1, h->sol, R1, nullptr, nullptr, nullptr, h->sol, 0.5 * dt));
            if ((rc = compute_N(h, h->S1, R2))) return rc;
            CK(stage(ST_RK4_STAGE, h->S2, h->sol, R2, nullptr, nullptr, nullptr, h->S1, 0.5 * dt));
            if ((rc = compute_N(h, h->S2, R3))) return rc;
            CK(stage(ST_RK4_STAGE, h->S1, h->sol, R3, nullptr, nullptr, nullptr, h->S2, dt));
            if ((rc = compute_N(h, h->S1, R4))) return rc;
            CK(stage(ST_RK4_FINAL, h->sol, h->sol, R1, R2, R3, R4, h->S1, dt));
        } else {
            double2* Ncur = h->Nb[h->ring];
            if ((rc = compute_N(h, h->sol, Ncur))) return rc;
            if ((rc = ifmab3_update_launch(h, Ncur))) return rc;
            h->ring = (h->ring + 1) % 3;
        }
        h->t += dt;
        h->step += 1;
        return SWRT_OK;
    };
    // Small grids are launch bound (a 512^2 step is four ~10 us kernels): replay the step from a CUDA graph.  A graph holds one
    // period of the history ring (three steps; one for the multi-stage steppers), whose kernel arguments then repeat exactly.
    static const int graph_mode = [] { const char* e = getenv("SWRT_GRAPH"); return e ? atoi(e) : 1; }();
    const bool ring_stepper = stepper == SWRT_IFMAB3 || stepper == SWRT_FILTEREDAB3;
    const int period = ring_stepper ? 3 : 1;
    const bool use_graph = graph_mode > 0 && h->P == 1 && !h->prof && !h->no_graph && !h->forcing && (graph_mode > 1 || (long long)h->d.nx * h->d.ny <= 1024LL * 1024LL);
    for (int s = 0; s < nsteps;) {
        if (use_graph && h->step >= 3 && nsteps - s >= period) {
            const int phase = ring_stepper ? h->ring : 0;
            if (!h->gexec[phase]) {
                const long long l0 = h->launches;
                CK(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeRelaxed));
                int rc = SWRT_OK;
                for (int p = 0; p < period && rc == SWRT_OK; ++p) rc = step_body();
                cudaGraph_t g = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(h->st, &g);
                if (rc != SWRT_OK) { if (g) cudaGraphDestroy(g); return rc; }
                CK(ce);
                const cudaError_t ie = cudaGraphInstantiate(&h->gexec[phase], g, 0);
                cudaGraphDestroy(g);
                CK(ie);
                h->glaunches[phase] = h->launches - l0;
            } else {
                for (int p = 0; p < period; ++p) { h->t += dt; h->step += 1; }   // the ring is back at `phase` after one period
                h->launches += h->glaunches[phase];
            }
            CK(cudaGraphLaunch(h->gexec[phase], h->st));
            s += period;
        } else {
            const int rc = step_body();
            if (rc) return rc;
            s += 1;
        }
    }
    return SWRT_OK;
}

int swrt_flow_clock(swrt_flow* h, double* t, long long* step) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (t) *t = h->t;
    if (step) *step = h->step;
    return SWRT_OK;
}
int swrt_flow_set_clock(swrt_flow* h, double t, long long step) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    h->t = t;
    h->step = step;
    return SWRT_OK;
}

int swrt_flow_get_field(swrt_flow* h, int which, double* real_host) {
    if (!h || !real_host) return fail(SWRT_ERR_ARG, "null pointer");
    const bool qg = h->d.model == SWRT_SWQG || h->d.model == SWRT_TWOLAYERQG || h->d.model == SWRT_MULTILAYERQG2;
    const bool ok = (which >= 0 && which < h->nvar) || (!qg && which == SWRT_FIELD_ZETA) ||
                    (qg && which >= SWRT_FIELD_QG_PSI && which < SWRT_FIELD_QG_PSI + 32 && ((which - 32) & 7) < h->nvar);
    if (!ok) return fail(SWRT_ERR_ARG, "unknown field %d for model %d", which, h->d.model);
    CK(cudaSetDevice(h->d.device));
    int rc = spectral_to_physical(h, which, h->phys);
    if (rc) return rc;
    CK(cudaMemcpyAsync(real_host, h->phys, sizeof(double) * (size_t)h->d.nx * h->d.ny, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

static double spectral_diag(swrt_flow* h, int which, int arg, cudaError_t* err, const double2* src = nullptr) {
    const int blocks = 296;
    { ProfScope ps(h, K_OTHER); spectral_diag_kernel<<<blocks, 256, 0, h->st>>>(src ? src : h->sol, h->L, which, arg, h->nvar, h->L.aux0, h->red); }
    std::vector<double> part(blocks);
    *err = cudaMemcpyAsync(part.data(), h->red, sizeof(double) * blocks, cudaMemcpyDeviceToHost, h->st);
    if (*err != cudaSuccess) return 0;
    *err = cudaStreamSynchronize(h->st);
    double acc = 0;
    for (double p : part) acc += p;
    const SpecLayout& L = h->L;
    return acc * h->d.Lx * h->d.Ly / ((double)L.nx * L.nx * (double)L.ny * L.ny);   // parsevalsum normalisation
}

int swrt_flow_set_field_physical(swrt_flow* h, int var, const double* real_host) {
    if (!h || !real_host) return fail(SWRT_ERR_ARG, "null pointer");
    if (var < 0 || var >= h->nvar) return fail(SWRT_ERR_ARG, "state variable %d out of range", var);
    if (h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "not available for a slab-decomposed flow");
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    CK(cudaMemcpyAsync(h->phys, real_host, sizeof(double) * (size_t)h->d.nx * h->d.ny, cudaMemcpyHostToDevice, h->st));
    cudaError_t e;
    { ProfScope ps(h, K_FIELD_B); SWRT_DISPATCH(L.nx, e, LN::forward_field(h->phys, h->H, nullptr, L, h->tw_x, h->sched, h->st)); }
    CK(e);
    { ProfScope ps(h, K_FIELD_A); SWRT_DISPATCH(L.ny, e, LN::forward_field_y(h->H, h->sol + (long long)var * L.vs, L, h->tw_y, h->st)); }
    CK(e);
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

int swrt_flow_energies(swrt_flow* h, double* ke, double* pe) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    const double A = h->d.Lx * h->d.Ly;
    cudaError_t e = cudaSuccess;
    double k = 0, p = 0;
    if (h->d.model == SWRT_SWQG) {            // swqg/SWQG.jl:205-222
        k = spectral_diag(h, DIAG_QG_K2PSI2, 0, &e) / (2 * A); CK(e);
        p = L.aux0 * spectral_diag(h, DIAG_QG_PSI2, 0, &e) / (2 * A); CK(e);
    } else if (h->d.model == SWRT_TWOLAYERQG || h->d.model == SWRT_MULTILAYERQG2) {   // swqg/TwoLayerQG.jl:221-250 (KE_1 + KE_2)
        k = spectral_diag(h, DIAG_QG_K2PSI2, 0, &e) / A; CK(e);
        k += spectral_diag(h, DIAG_QG_K2PSI2, 1, &e) / A; CK(e);
        p = L.aux0 * spectral_diag(h, DIAG_QG_DPSI2, 0, &e) / (2 * A); CK(e);
    } else if (h->d.model == SWRT_THOMASYAMADA) {  // thomasyamada/ThomasYamada.jl:333-338 baroclinic_energy(prob) = (P2(uc)+P2(vc), P2(pc))
        k = spectral_diag(h, DIAG_ABS2_VAR, 1, &e); CK(e);      // baroclinic_energy: raw parsevalsum2 values, no 1/(2A)
        k += spectral_diag(h, DIAG_ABS2_VAR, 2, &e); CK(e);
        p = spectral_diag(h, DIAG_ABS2_VAR, 3, &e); CK(e);
    } else {                                   // rsw/RotatingShallowWater.jl:323-336
        k = spectral_diag(h, DIAG_ABS2_VAR, 0, &e) / (2 * A); CK(e);
        k += spectral_diag(h, DIAG_ABS2_VAR, 1, &e) / (2 * A); CK(e);
        if (h->d.model == SWRT_RSW_QUADHEIGHT) {   // rsw/QuadHeightModifiedShallowWater.jl:355-358: 0.5 Cg2 Re(mh[1,1]) / (Lx Ly)
            double2 m00 = make_double2(0, 0);
            if (L.kr_off == 0) {
                CK(cudaMemcpyAsync(&m00, h->sol + 2 * L.vs, sizeof m00, cudaMemcpyDeviceToHost, h->st));
                CK(cudaStreamSynchronize(h->st));
            }
            p = 0.5 * L.Cg2 * m00.x / A;
        } else {
            p = 0.5 * L.Cg2 * spectral_diag(h, DIAG_ABS2_VAR, 2, &e) / A; CK(e);
        }
    }
    if (ke) *ke = k;
    if (pe) *pe = p;
    return SWRT_OK;
}

// ---- wave / balanced projections on the device
static int decompose_launch(swrt_flow* h, int mode) {
    const bool rsw = h->d.model == SWRT_RSW || h->d.model == SWRT_RSW_MODIFIED || h->d.model == SWRT_RSW_LINDBORG;
    if (mode == DEC_TY ? h->d.model != SWRT_THOMASYAMADA : !rsw)
        return fail(SWRT_ERR_UNSUPPORTED, "the wave/balanced projection is defined for the eta-based RSW models and Thomas-Yamada (model %d)", h->d.model);
    if (h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "not available for a slab-decomposed flow");
    const SpecLayout& L = h->L;
    const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
    if (nmodes == 0) return SWRT_OK;
    { ProfScope ps(h, K_OTHER); decompose_kernel<<<(int)((nmodes + 255) / 256), 256, 0, h->st>>>(h->sol, L, mode, h->d.f, L.Cg2, h->G, h->H); }
    CK(cudaGetLastError());
    return SWRT_OK;
}
static int fetch3(swrt_flow* h, const double2* src, void* host) {
    { ProfScope ps(h, K_OTHER); unpack_sol_kernel<<<592, 256, 0, h->st>>>(src, h->stage, h->L, h->nkr, 3); }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, h->stage, sizeof(double2) * (size_t)h->nkr * h->d.ny * 3, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

int swrt_flow_wave_balanced_decomposition(swrt_flow* h, void* balanced_host, void* wave_host) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    int rc = decompose_launch(h, h->d.model == SWRT_THOMASYAMADA ? DEC_TY : DEC_RSW);
    if (rc) return rc;
    if (balanced_host && (rc = fetch3(h, h->G, balanced_host))) return rc;
    if (wave_host && (rc = fetch3(h, h->H, wave_host))) return rc;
    return SWRT_OK;
}

int swrt_flow_wave_balanced_weights(swrt_flow* h, void* c0_host, void* cp_host, void* cm_host) {
    if (!h || !c0_host || !cp_host || !cm_host) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    int rc = decompose_launch(h, DEC_RSW_WTS);
    if (rc) return rc;
    { ProfScope ps(h, K_OTHER); unpack_sol_kernel<<<592, 256, 0, h->st>>>(h->G, h->stage, h->L, h->nkr, 3); }
    CK(cudaGetLastError());
    const size_t one = sizeof(double2) * (size_t)h->nkr * h->d.ny;
    void* outs[3] = {c0_host, cp_host, cm_host};
    for (int c = 0; c < 3; ++c) CK(cudaMemcpyAsync(outs[c], reinterpret_cast<char*>(h->stage) + c * one, one, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}

int swrt_flow_wave_balanced_energies(swrt_flow* h, double* out) {
    if (!h || !out) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    const bool ty = h->d.model == SWRT_THOMASYAMADA;
    int rc = decompose_launch(h, ty ? DEC_TY : DEC_RSW);
    if (rc) return rc;
    // TY: raw parsevalsum2 values (thomasyamada/ThomasYamada.jl:355-367); RSW: kinetic_energy / potential_energy of each part (:323-336)
    const double A2 = 2 * h->d.Lx * h->d.Ly, ks = ty ? 1.0 : 1.0 / A2, ps = ty ? 1.0 : h->L.Cg2 / A2;
    cudaError_t e = cudaSuccess;
    const double2* parts[2] = {h->H, h->G};   // wave first, like the reference's return value
    for (int p = 0; p < 2; ++p) {
        double k = spectral_diag(h, DIAG_ABS2_VAR, 0, &e, parts[p]); CK(e);
        k += spectral_diag(h, DIAG_ABS2_VAR, 1, &e, parts[p]); CK(e);
        const double pe = spectral_diag(h, DIAG_ABS2_VAR, 2, &e, parts[p]); CK(e);
        out[2 * p] = ks * k;
        out[2 * p + 1] = ps * pe;
    }
    return SWRT_OK;
}

int swrt_flow_barotropic_energy(swrt_flow* h, double* e_out) {
    if (!h || !e_out) return fail(SWRT_ERR_ARG, "null pointer");
    if (h->d.model != SWRT_THOMASYAMADA) return fail(SWRT_ERR_UNSUPPORTED, "barotropic_energy is defined for Thomas-Yamada");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e = cudaSuccess;
    *e_out = spectral_diag(h, DIAG_INVK2_ABS2_VAR, 0, &e); CK(e);
    return SWRT_OK;
}

int swrt_flow_layer_kinetic_energy(swrt_flow* h, int layer, double* ke) {
    if (!h || !ke) return fail(SWRT_ERR_ARG, "null pointer");
    if ((h->d.model != SWRT_TWOLAYERQG && h->d.model != SWRT_MULTILAYERQG2) || layer < 0 || layer > 1) return fail(SWRT_ERR_ARG, "layer kinetic energy is defined for the two-layer model");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e = cudaSuccess;
    *ke = spectral_diag(h, DIAG_QG_K2PSI2, layer, &e) / (h->d.Lx * h->d.Ly);
    CK(e);
    return SWRT_OK;
}

int swrt_flow_max_abs_uv(swrt_flow* h, double* umax, double* vmax) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e = cudaSuccess;
    double* outs[2] = {umax, vmax};
    const bool qg = h->d.model == SWRT_SWQG || h->d.model == SWRT_TWOLAYERQG || h->d.model == SWRT_MULTILAYERQG2;
    const int uv0 = h->d.model == SWRT_THOMASYAMADA ? 1 : 0;   // TY: baroclinic (u_c, v_c) are state variables 1, 2
    for (int v = 0; v < 2; ++v) {
        if (!outs[v]) continue;
        *outs[v] = 0.0;
        for (int layer = 0; layer < (qg ? h->nvar : 1); ++layer) {   // QG: u = -psi_y, v = psi_x of every layer
            int rc = spectral_to_physical(h, qg ? (v == 0 ? SWRT_FIELD_QG_U : SWRT_FIELD_QG_V) + layer : v + uv0, h->phys);
            if (rc) return rc;
            const double m = reduce_host(h, h->phys, (long long)h->d.nx * h->d.ny, 1, &e);
            *outs[v] = (m != m || *outs[v] != *outs[v]) ? m + *outs[v] : std::max(*outs[v], m);
            CK(e);
        }
    }
    return SWRT_OK;
}

int swrt_flow_has_nan(swrt_flow* h, int* flag) {
    if (!h || !flag) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e = cudaSuccess;
    const double c = reduce_host(h, reinterpret_cast<const double*>(h->sol), 2 * h->L.vs, 2, &e);  // vars.uh / vars.qh[:,:,1]
    CK(e);
    *flag = c > 0;
    return SWRT_OK;
}

static int check_psi_kind(swrt_flow* h, int psi_kind) {
    const bool rsw_family = h->d.model == SWRT_RSW || h->d.model == SWRT_RSW_MODIFIED || h->d.model == SWRT_RSW_LINDBORG;   // (QuadHeight carries m, not eta)
    const bool ok = (psi_kind == SWRT_PSI_RSW_BALANCED && rsw_family) || (psi_kind == SWRT_PSI_SWQG && h->d.model == SWRT_SWQG) ||
                    ((psi_kind == SWRT_PSI_TWOLAYER_BAROCLINIC || psi_kind == SWRT_PSI_TWOLAYER_MEAN) &&
                     (h->d.model == SWRT_TWOLAYERQG || h->d.model == SWRT_MULTILAYERQG2));
    return ok ? SWRT_OK : fail(SWRT_ERR_ARG, "psi kind %d does not apply to model %d", psi_kind, h->d.model);
}

// ------------------------------------------------------------------ slab-decomposed step (phases between the caller's all-to-alls)
int swrt_flow_set_stream(swrt_flow* h, void* cuda_stream) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    CK(cudaStreamSynchronize(h->st));
    if (h->own_stream && h->st) cudaStreamDestroy(h->st);
    h->st = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return SWRT_OK;
}
int swrt_slab_info(swrt_flow* h, int* yrows, int* chunk, int* njobs_a, int* njobs_b) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (yrows) *yrows = h->L.yrows;
    if (chunk) *chunk = h->L.kr_pad;
    if (njobs_a) *njobs_a = model_njobs_a(h->d.model);
    if (njobs_b) *njobs_b = model_njobs_b(h->d.model);
    return SWRT_OK;
}
int swrt_slab_buffer(swrt_flow* h, int which, void** device_ptr, long long* nbytes) {
    if (!h || !device_ptr) return fail(SWRT_ERR_ARG, "null pointer");
    if (h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    const long long fb = (long long)sizeof(double2) * h->L.vs;
    void* p = nullptr;
    long long nb = 0;
    switch (which) {
        case SWRT_SLAB_A_SEND: p = h->G; nb = fb * h->njobs_a; break;
        case SWRT_SLAB_A_RECV: p = h->G2; nb = fb * h->njobs_a; break;
        case SWRT_SLAB_B_SEND: p = h->H2; nb = fb * h->njobs_b; break;
        case SWRT_SLAB_B_RECV: p = h->H; nb = fb * h->njobs_b; break;
        case SWRT_SLAB_SNAP0: case SWRT_SLAB_SNAP1:
            p = h->snap[h->slot_map[which - SWRT_SLAB_SNAP0]];
            nb = (long long)sizeof(double) * h->d.nx * (h->L.yrows + 2 * h->halo) * SNAP_STRIDE;   // band + halo rows of the 5-field records
            break;
        default: return fail(SWRT_ERR_ARG, "unknown slab buffer %d", which);
    }
    *device_ptr = p;
    if (nbytes) *nbytes = nb;
    return SWRT_OK;
}
int swrt_slab_ipc_handle(swrt_flow* h, int which, void* handle64) {
    if (!h || !handle64 || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    void* ptr = which == SWRT_SLAB_A_RECV ? (void*)h->G2 : which == SWRT_SLAB_A_SEND ? (void*)h->G : which == SWRT_SLAB_B_RECV ? (void*)h->H
              : which == SWRT_SLAB_FLAGS ? (void*)h->flags : which == SWRT_SLAB_BAND ? (void*)h->band : nullptr;
    if (!ptr) return fail(SWRT_ERR_ARG, "shared buffers: the two receive buffers, (pull variant) the first send buffer, the barrier flags and the band snapshots");
    CK(cudaSetDevice(h->d.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, ptr));
    memcpy(handle64, &mh, 64);
    return SWRT_OK;
}
int swrt_slab_ipc_open(swrt_flow* h, int which, int peer_rank, const void* handle64) {
    if (!h || !handle64 || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    const bool team = which == SWRT_SLAB_FLAGS || which == SWRT_SLAB_BAND;
    if ((!team && which != SWRT_SLAB_A_RECV && which != SWRT_SLAB_B_RECV && which != SWRT_SLAB_A_SEND) || peer_rank < 0 || peer_rank >= h->P)
        return fail(SWRT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(h->d.device));
    void* p = nullptr;
    if (peer_rank != h->rank) {
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handle64, 64);
        CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    }
    if (which == SWRT_SLAB_FLAGS) { h->peerflags[peer_rank] = p ? (TeamFlags*)p : h->flags; return SWRT_OK; }
    if (which == SWRT_SLAB_BAND) { h->peerband[peer_rank] = p ? (double*)p : h->band; return SWRT_OK; }
    const int w = which == SWRT_SLAB_A_RECV ? 0 : which == SWRT_SLAB_B_RECV ? 1 : 2;
    h->peer[w][peer_rank] = p ? (double2*)p : (w == 0 ? h->G2 : w == 1 ? h->H : h->G);
    bool all = true;
    for (int w2 = 0; w2 < 2; ++w2)
        for (int r = 0; r < h->P; ++r) all = all && h->peer[w2][r] != nullptr;
    h->p2p = all;     // both transposes become direct peer stores once every receive buffer is mapped
    bool allg = all;
    for (int r = 0; r < h->P; ++r) allg = allg && h->peer[2][r] != nullptr;
    h->pull = allg;   // with every rank's first send buffer mapped too, the x-pass pulls its input segments instead
    return SWRT_OK;
}
int swrt_slab_set_mode(swrt_flow* h, int mode) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    if (mode >= 16) { h->slab_b_copy = 1; mode -= 16; }     // + 16: the second transpose ships through a block-copy kernel too
    else h->slab_b_copy = 0;
    if (mode < 0 || mode > 2) return fail(SWRT_ERR_ARG, "mode must be 0 (push), 1 (pull) or 2 (block copy) [+ 16: block copy for the second transpose]");
    if (mode == 1 && !h->pull) return fail(SWRT_ERR_STATE, "the pull variant needs every rank's first send buffer mapped");
    if (mode == 2 && !h->p2p) return fail(SWRT_ERR_STATE, "the block-copy variant needs the receive buffers mapped");
    h->slab_mode = mode;
    return SWRT_OK;
}
// after a y-pass that stored into the local send buffer: ship the blocks (mode 2)
static int slab_ship_a(swrt_flow* h, int njobs) {
    if (h->slab_mode != 2 || !h->p2p) return SWRT_OK;
    OutPeers dst{};
    for (int d = 0; d < h->P; ++d) dst.p[d] = h->peer[0][d];
    dst.self = h->rank;
    const long long blk = (long long)njobs * h->L.yrows * h->L.kr_pad;
    { ProfScope ps(h, K_OTHER); slab_block_copy_kernel<<<148 * 8, 256, 0, h->st>>>(h->G, dst, blk, h->P); }
    CK(cudaGetLastError());
    return SWRT_OK;
}
int swrt_slab_p2p(swrt_flow* h, int* enabled) {
    if (!h || !enabled) return fail(SWRT_ERR_ARG, "null pointer");
    *enabled = h->p2p ? 1 : 0;
    return SWRT_OK;
}
int swrt_slab_stage_a(swrt_flow* h) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e;
    { ProfScope ps(h, K_STAGE_A); SWRT_DISPATCH(h->L.ny, e, LN::stage_a(h->d.model, h->sol, out_slab(h, 0, h->G, model_njobs_a(h->d.model)), h->L, h->tw_y, h->st)); }
    CK(e);
    return slab_ship_a(h, model_njobs_a(h->d.model));
}
static int slab_stage_b(swrt_flow* h, int nj_total) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    CK(cudaSetDevice(h->d.device));
    cudaError_t e;
    { ProfScope ps(h, K_STAGE_B); SWRT_DISPATCH(h->L.nx, e, LN::stage_b_slab(h->d.model, in_slab(h, nj_total), out_slab(h, 1, h->H2, model_njobs_b(h->d.model)), h->L, h->tw_x, h->sched, h->st, nj_total)); }
    CK(e);
    if (h->slab_b_copy && h->p2p) {   // ship the product blocks in full lines
        OutPeers dst{};
        for (int d = 0; d < h->P; ++d) dst.p[d] = h->peer[1][d];
        dst.self = h->rank;
        const long long blk = (long long)model_njobs_b(h->d.model) * h->L.yrows * h->L.kr_pad;
        { ProfScope ps(h, K_OTHER); slab_block_copy_kernel<<<148 * 8, 256, 0, h->st>>>(h->H2, dst, blk, h->P); }
        CK(cudaGetLastError());
    }
    return SWRT_OK;
}
int swrt_slab_stage_b(swrt_flow* h) { return slab_stage_b(h, h ? model_njobs_a(h->d.model) : 0); }
static int ifmab3_update_launch(swrt_flow* h, double2* Ncur);
int swrt_slab_stage_c(swrt_flow* h) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    if (is_etd(h->d.stepper) || h->d.stepper == SWRT_FILTEREDRK4)
        return fail(SWRT_ERR_UNSUPPORTED, "the phase-by-phase slab step (stage_a / all-to-all / stage_b / all-to-all / stage_c) is the IFMAB3 / FilteredAB3 step; "
                                          "a multi-stage stepper runs through swrt_slab_step with the peers mapped");
    CK(cudaSetDevice(h->d.device));
    double2* Ncur = h->Nb[h->ring];
    cudaError_t e;
    SpecLayout Lc = h->L;
    Lc.forcing = h->forcing;
    { ProfScope ps(h, K_STAGE_C); SWRT_DISPATCH(h->L.ny, e, LN::stage_c(h->d.model, h->sol, h->H, Ncur, Lc, h->tw_y, h->st)); }
    CK(e);
    int rc = ifmab3_update_launch(h, Ncur);
    if (rc) return rc;
    h->ring = (h->ring + 1) % 3;
    h->t += h->d.dt;
    h->step += 1;
    return SWRT_OK;
}
// snapshot writers wait for the reads of packet handles that run on their own streams
static cudaError_t wait_readers(swrt_flow* h) {
    for (swrt_packets* r : h->readers) {
        cudaError_t e = cudaStreamWaitEvent(h->st, packets_done_event(r), 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
int swrt_slab_psi_a(swrt_flow* h, int psi_kind) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    int rc = check_psi_kind(h, psi_kind);
    if (rc) return rc;
    CK(cudaSetDevice(h->d.device));
    const bool pf = h->interp == SWRT_INTERP_BSPLINE2 || h->interp == SWRT_INTERP_BSPLINE3;
    PsiLoader ld{h->sol, h->L.vs, psi_kind, h->d.f, h->L.aux0, pf ? h->d.Lx / h->d.nx : 0.0, pf ? h->d.Ly / h->d.ny : 0.0};
    if (h->interp == SWRT_INTERP_BSPLINE3) { ld.pc0 = 2.0 / 3.0; ld.pc1 = 1.0 / 3.0; }
    cudaError_t e;
    // like the single-GPU snapshot: psih materialised once (this rank's columns), so that the three y-jobs run through the
    // prefetching y-pass instead of re-deriving the streamfunction three times
    bool materialise = false;
    SWRT_DISPATCH(h->L.ny, e, (materialise = LN::psi_prefetch, cudaSuccess));
    CK(e);
    if (materialise && h->L.kr_keep > 0) {
        const long long nmodes = (long long)(h->L.ny - (h->L.lz1 - h->L.lz0)) * h->L.kr_keep;
        ProfScope ps(h, K_PSI);
        psi_kernel<<<(unsigned)((nmodes + 255) / 256), 256, 0, h->st>>>(ld, h->L, h->psih);
        CK(cudaGetLastError());
    }
    { ProfScope ps(h, K_PSI_A); SWRT_DISPATCH(h->L.ny, e, LN::psi_stage_a(ld, materialise ? h->psih : nullptr, h->L, out_slab(h, 0, h->G, 3), h->tw_y, h->st)); }
    CK(e);
    return slab_ship_a(h, 3);
}
static PsiLoader slab_psi_loader(swrt_flow* h, int psi_kind) {
    const bool pf = h->interp == SWRT_INTERP_BSPLINE2 || h->interp == SWRT_INTERP_BSPLINE3;
    PsiLoader ld{h->sol, h->L.vs, psi_kind, h->d.f, h->L.aux0, pf ? h->d.Lx / h->d.nx : 0.0, pf ? h->d.Ly / h->d.ny : 0.0};
    if (h->interp == SWRT_INTERP_BSPLINE3) { ld.pc0 = 2.0 / 3.0; ld.pc1 = 1.0 / 3.0; }
    return ld;
}
static int slab_snap_b(swrt_flow* h, int slot, int nj_total, int j0) {
    if (!h || h->P <= 1 || slot < 0 || slot > 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow / bad slot");
    if (h->interp == SWRT_INTERP_HERMITE_BICUBIC) return fail(SWRT_ERR_UNSUPPORTED, "slab snapshots are built for the 5-field node data");
    CK(cudaSetDevice(h->d.device));
    double* rows = h->snap[h->slot_map[slot]] + (long long)h->halo * h->d.nx * SNAP_STRIDE;   // this rank's band: the owned rows follow the lower halo
    CK(wait_readers(h));
    cudaError_t e;
    { ProfScope ps(h, K_SNAP_B); SWRT_DISPATCH(h->L.nx, e, LN::snap_stage_b_slab(in_slab(h, nj_total), rows, h->L, h->tw_x, h->sched, h->st, nj_total, j0)); }
    CK(e);
    return SWRT_OK;
}
int swrt_slab_snap_b(swrt_flow* h, int slot) { return slab_snap_b(h, slot, 3, 0); }
// team mode, fused: stage A of the flow step AND the snapshot's y-jobs of the same state in one y-pass / one transpose
static int slab_stage_a_fused(swrt_flow* h, int psi_kind) {
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    const PsiLoader ld = slab_psi_loader(h, psi_kind);
    if (L.kr_keep > 0) {
        const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
        ProfScope ps(h, K_PSI);
        psi_kernel<<<(unsigned)((nmodes + 255) / 256), 256, 0, h->st>>>(ld, L, h->psih);
        CK(cudaGetLastError());
    }
    const int nj = model_njobs_a(h->d.model) + 3;
    cudaError_t e;
    { ProfScope ps(h, K_STAGE_A); SWRT_DISPATCH(L.ny, e, LN::stage_a_fused(h->d.model, h->sol, h->psih, out_slab(h, 0, h->G, nj), L, h->tw_y, h->st)); }
    CK(e);
    return slab_ship_a(h, nj);
}

// ---- team mode: barrier, native step, band snapshot (no communication library on the data path)
int swrt_slab_set_barrier(swrt_flow* h, int mode, void (*callback)(void*), void* arg) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    if (mode != 0 && mode != 1) return fail(SWRT_ERR_ARG, "barrier mode must be 0 (device flags) or 1 (host callback)");
    if (mode == 1 && !callback) return fail(SWRT_ERR_ARG, "the host barrier needs a callback");
    h->barrier_mode = mode;
    h->barrier_cb = callback;
    h->barrier_arg = arg;
    return SWRT_OK;
}
static int team_barrier(swrt_flow* h, int channel = 0, cudaStream_t st = nullptr) {
    // Two independent channels (flag words + epochs): every rank issues the same sequence of barriers PER CHANNEL, but the flow's
    // stream and a packet handle's own stream advance independently, so their barriers must not share a counter.
    if (!st) st = h->st;
    if (h->barrier_mode == 1) {          // every rank's stream drained, then the caller's host barrier (processes sharing one GPU)
        CK(cudaStreamSynchronize(st));
        h->barrier_cb(h->barrier_arg);
        return SWRT_OK;
    }
    TeamPeers tp{};
    for (int r = 0; r < h->P; ++r) {
        if (!h->peerflags[r]) return fail(SWRT_ERR_STATE, "team barrier: the flags of rank %d are not mapped (swrt_slab_ipc_open SWRT_SLAB_FLAGS)", r);
        tp.f[r] = h->peerflags[r];
    }
    h->epoch[channel] += 1;
    { ProfScope ps(h, K_OTHER, st); team_barrier_kernel<<<1, 32, 0, st>>>(tp, h->P, h->rank, h->epoch[channel], channel); }
    CK(cudaGetLastError());
    return SWRT_OK;
}
int swrt_slab_barrier(swrt_flow* h) {
    if (!h || h->P <= 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow");
    CK(cudaSetDevice(h->d.device));
    return team_barrier(h);
}
// calcN!(state) of a slab-decomposed flow for an arbitrary state (the stages of ETDRK4 / FilteredRK4): stage A of `state`, barrier,
// stage B, barrier, stage C into Nout.  COLLECTIVE: every rank calls it in the same sequence.
static int slab_compute_N(swrt_flow* h, const double2* state, double2* Nout) {
    if (!h->p2p) return fail(SWRT_ERR_STATE, "a slab-decomposed flow with a multi-stage stepper needs the peers' receive buffers mapped (swrt_slab_ipc_open)");
    const int nj = model_njobs_a(h->d.model);
    cudaError_t e;
    int rc;
    { ProfScope ps(h, K_STAGE_A); SWRT_DISPATCH(h->L.ny, e, LN::stage_a(h->d.model, state, out_slab(h, 0, h->G, nj), h->L, h->tw_y, h->st)); }
    CK(e);
    if ((rc = slab_ship_a(h, nj))) return rc;
    if ((rc = team_barrier(h))) return rc;
    if ((rc = slab_stage_b(h, nj))) return rc;
    if ((rc = team_barrier(h))) return rc;
    SpecLayout Lc = h->L;
    Lc.forcing = h->forcing;
    { ProfScope ps(h, K_STAGE_C); SWRT_DISPATCH(h->L.ny, e, LN::stage_c(h->d.model, state, h->H, Nout, Lc, h->tw_y, h->st)); }
    CK(e);
    return SWRT_OK;
}
int swrt_slab_step(swrt_flow* h, int nsteps) {
    if (!h || h->P <= 1 || nsteps < 0) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow / bad step count");
    if (!h->p2p) return fail(SWRT_ERR_STATE, "swrt_slab_step needs the peers' receive buffers mapped (swrt_slab_ipc_open); without them drive the phases and the all-to-alls from the host");
    if (is_etd(h->d.stepper) || h->d.stepper == SWRT_FILTEREDRK4) return flow_step_impl(h, nsteps);   // four calcN! per step
    int rc;
    for (int s = 0; s < nsteps; ++s) {
        if ((rc = swrt_slab_stage_a(h))) return rc;
        if ((rc = team_barrier(h))) return rc;
        if ((rc = swrt_slab_stage_b(h))) return rc;
        if ((rc = team_barrier(h))) return rc;
        if ((rc = swrt_slab_stage_c(h))) return rc;
    }
    return SWRT_OK;
}
// halo rows of one level from the two neighbours' band buffers (the caller has placed a barrier after the writes)
static int band_halo_pull(swrt_flow* h, int slot) {
    const int lev = h->slot_map[slot];
    const long long row_doubles = (long long)h->d.nx * SNAP_STRIDE, lev_doubles = (long long)(h->L.yrows + 2 * h->halo) * row_doubles;
    const int below = (h->rank + h->P - 1) % h->P, above = (h->rank + 1) % h->P;
    if (!h->peerband[below] || !h->peerband[above]) return fail(SWRT_ERR_STATE, "band snapshots of the neighbours are not mapped (swrt_slab_ipc_open SWRT_SLAB_BAND)");
    { ProfScope ps(h, K_OTHER);
      team_halo_pull_kernel<<<64, 256, 0, h->st>>>(h->snap[lev], h->peerband[below] + lev * lev_doubles, h->peerband[above] + lev * lev_doubles, h->halo, h->L.yrows, row_doubles); }
    CK(cudaGetLastError());
    return SWRT_OK;
}
int swrt_slab_band_snapshot(swrt_flow* h, int psi_kind, int slot) {
    if (!h || h->P <= 1 || slot < 0 || slot > 1) return fail(SWRT_ERR_STATE, "not a slab-decomposed flow / bad slot");
    if (!h->p2p) return fail(SWRT_ERR_STATE, "swrt_slab_band_snapshot needs the peers' receive buffers mapped");
    int rc;
    if ((rc = swrt_slab_psi_a(h, psi_kind))) return rc;
    if ((rc = team_barrier(h))) return rc;
    if ((rc = swrt_slab_snap_b(h, slot))) return rc;
    if ((rc = team_barrier(h))) return rc;          // every rank's band rows are written
    return band_halo_pull(h, slot);                 // (the neighbours overwrite this level two steps later, behind >= 4 more barriers)
}
int swrt_slab_band_info(swrt_flow* h, int* row0, int* rows, int* halo) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (row0) *row0 = h->rank * h->L.yrows;
    if (rows) *rows = h->L.yrows;
    if (halo) *halo = h->halo;
    return SWRT_OK;
}

int swrt_flow_velocity_snapshot(swrt_flow* h, int psi_kind, int slot) {
    if (!h || slot < 0 || slot > 1) return fail(SWRT_ERR_ARG, "bad argument");
    if (h->P > 1) return fail(SWRT_ERR_STATE, "slab-decomposed flow: use swrt_slab_psi_a / swrt_slab_snap_b around the all-to-all");
    { int rc = check_psi_kind(h, psi_kind); if (rc) return rc; }
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    const bool pf = h->interp == SWRT_INTERP_BSPLINE2 || h->interp == SWRT_INTERP_BSPLINE3;
    PsiLoader ld{h->sol, L.vs, psi_kind, h->d.f, L.aux0, pf ? h->d.Lx / h->d.nx : 0.0, pf ? h->d.Ly / h->d.ny : 0.0};
    if (h->interp == SWRT_INTERP_BSPLINE3) { ld.pc0 = 2.0 / 3.0; ld.pc1 = 1.0 / 3.0; }
    if (h->interp == SWRT_INTERP_NUFFT) {
        if (h->refine != 2) return fail(SWRT_ERR_STATE, "the NUFFT mode samples a 2x oversampled node grid: call swrt_flow_set_snapshot_refinement(h, 2) first");
        int rc = nufft_tables(h);
        if (rc) return rc;
        ld.ptab_x = h->ptab; ld.ptab_y = h->ptab + h->nkr; ld.tab_kr_pad = L.kr_pad; ld.tab_kr_off = L.kr_off;
    }
    CK(wait_readers(h));
    cudaError_t e;
    if (h->refine > 1) {
        const SpecLayout& Ls = h->Ls;
        ld.pdx = pf ? h->d.Lx / Ls.nx : 0.0;
        ld.pdy = pf ? h->d.Ly / Ls.ny : 0.0;
        if (L.kr_keep > 0) {
            const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
            ProfScope ps(h, K_PSI);
            psi_kernel<<<(unsigned)((nmodes + 255) / 256), 256, 0, h->st>>>(ld, L, h->psih_s, Ls.ny - L.ny);
            CK(cudaGetLastError());
        }
        { ProfScope ps(h, K_PSI_A); SWRT_DISPATCH(Ls.ny, e, LN::psi_stage_a_refined(h->psih_s, Ls, out_local(h->Gs), h->tw_ys, h->st)); }
        CK(e);
        const int mode = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? 1 : (h->interp == SWRT_INTERP_BILINEAR_F32 ? 2 : 0);
        { ProfScope ps(h, K_SNAP_B); SWRT_DISPATCH(Ls.nx, e, LN::snap_stage_b(h->Gs, h->snap[h->slot_map[slot]], mode, Ls, h->tw_xs, h->sched, h->st, 1.0 / ((double)L.nx * (double)L.ny))); }
        CK(e);
        return SWRT_OK;
    }
    bool materialise = false;
    SWRT_DISPATCH(L.ny, e, (materialise = LN::psi_prefetch, cudaSuccess));
    CK(e);
    if (materialise && L.kr_keep > 0) {
        const long long nmodes = (long long)(L.ny - (L.lz1 - L.lz0)) * L.kr_keep;
        ProfScope ps(h, K_PSI);
        psi_kernel<<<(unsigned)((nmodes + 255) / 256), 256, 0, h->st>>>(ld, L, h->psih);
        CK(cudaGetLastError());
    }
    { ProfScope ps(h, K_PSI_A); SWRT_DISPATCH(L.ny, e, LN::psi_stage_a(ld, materialise ? h->psih : nullptr, L, out_local(h->G), h->tw_y, h->st)); }
    CK(e);
    { ProfScope ps(h, K_SNAP_B); SWRT_DISPATCH(L.nx, e, LN::snap_stage_b(h->G, h->snap[h->slot_map[slot]], h->interp == SWRT_INTERP_HERMITE_BICUBIC ? 1 : (h->interp == SWRT_INTERP_BILINEAR_F32 ? 2 : 0), L, h->tw_x, h->sched, h->st)); }
    CK(e);
    return SWRT_OK;
}

int swrt_flow_set_snapshot_refinement(swrt_flow* h, int refine) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (refine != 1 && refine != 2) return fail(SWRT_ERR_ARG, "refinement must be 1 or 2");
    if (h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "not available for a slab-decomposed flow");
    if (!supported_n(refine * h->d.nx) || !supported_n(refine * h->d.ny)) return fail(SWRT_ERR_UNSUPPORTED, "refined grid %d x %d exceeds the supported sizes", refine * h->d.nx, refine * h->d.ny);
    if (h->npackets > 0) return fail(SWRT_ERR_STATE, "set the refinement before creating packet handles (%d attached)", h->npackets);
    CK(cudaSetDevice(h->d.device));
    CK(cudaStreamSynchronize(h->st));
    // allocate everything first: a failing allocation leaves the handle exactly as it was
    const size_t nodes = (size_t)refine * h->d.nx * (size_t)refine * h->d.ny;
    double* nsnap[2] = {nullptr, nullptr};
    double2 *npsi = nullptr, *nG = nullptr, *ntx = nullptr, *nty = nullptr;
    SpecLayout Ls = h->L;
    cudaError_t e = cudaSuccess;
    for (int lev = 0; lev < 2 && e == cudaSuccess; ++lev) {
        e = cudaMalloc(&nsnap[lev], sizeof(double) * nodes * SNAP3_STRIDE);
        if (e == cudaSuccess) e = cudaMemset(nsnap[lev], 0, sizeof(double) * nodes * SNAP3_STRIDE);
    }
    if (refine > 1 && e == cudaSuccess) {
        Ls.nx = refine * h->d.nx; Ls.ny = refine * h->d.ny;
        Ls.lz1 = Ls.ny - (h->L.ny - h->L.lz1);      // the negative-l rows keep their distance from the end
        Ls.vs = (long long)Ls.ny * Ls.kr_pad;
        Ls.yrows = Ls.ny;
        Ls.yshift = 0; while ((1 << Ls.yshift) < Ls.ny) ++Ls.yshift;
        const size_t fb = sizeof(double2) * (size_t)Ls.vs;
        e = cudaMalloc(&npsi, fb);
        if (e == cudaSuccess) e = cudaMemset(npsi, 0, fb);
        if (e == cudaSuccess) e = cudaMalloc(&nG, 3 * fb);
        if (e == cudaSuccess) e = cudaMemset(nG, 0, 3 * fb);
        if (e == cudaSuccess) e = upload_twiddles(Ls.nx, &ntx);
        if (e == cudaSuccess) e = upload_twiddles(Ls.ny, &nty);
    }
    if (e != cudaSuccess) {
        cudaFree(nsnap[0]); cudaFree(nsnap[1]); cudaFree(npsi); cudaFree(nG); cudaFree(ntx); cudaFree(nty);
        return fail(SWRT_ERR_CUDA, "snapshot refinement: %s", cudaGetErrorString(e));
    }
    cudaFree(h->psih_s); cudaFree(h->Gs); cudaFree(h->tw_xs); cudaFree(h->tw_ys);
    cudaFree(h->snap[0]); cudaFree(h->snap[1]);
    h->snap[0] = nsnap[0]; h->snap[1] = nsnap[1];
    h->psih_s = npsi; h->Gs = nG; h->tw_xs = ntx; h->tw_ys = nty;
    h->refine = refine;
    if (refine > 1) h->Ls = Ls;
    build_tmaps(h);
    return SWRT_OK;
}
int swrt_flow_snapshot_dims(swrt_flow* h, int* nx, int* ny) {
    if (!h || !nx || !ny) return fail(SWRT_ERR_ARG, "null pointer");
    *nx = h->refine * h->d.nx;
    *ny = h->refine * h->d.ny;
    return SWRT_OK;
}

int swrt_flow_set_interp(swrt_flow* h, int interp) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (interp < SWRT_INTERP_BILINEAR || interp > SWRT_INTERP_NUFFT) return fail(SWRT_ERR_UNSUPPORTED, "interpolant %d not implemented", interp);
    if (interp != SWRT_INTERP_BILINEAR && h->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "a slab-decomposed flow holds band snapshots of the 5-field bilinear node records only");
    h->interp = interp;
    return SWRT_OK;
}
int swrt_flow_set_nufft_width(swrt_flow* h, int nw) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (nw < 4 || nw > NUFFT_MAXW) return fail(SWRT_ERR_ARG, "NUFFT kernel width must be in [4, %d]", NUFFT_MAXW);
    h->nufft_w = nw;
    return SWRT_OK;
}
int swrt_flow_snapshot_fields(swrt_flow* h, int* nfields) {
    if (!h || !nfields) return fail(SWRT_ERR_ARG, "null pointer");
    *nfields = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? SNAP3_NC : SNAP_NC;
    return SWRT_OK;
}

int swrt_flow_swap_snapshots(swrt_flow* h, int alias) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (alias) h->slot_map[0] = h->slot_map[1];
    else { const int t = h->slot_map[0]; h->slot_map[0] = h->slot_map[1]; h->slot_map[1] = t; }
    return SWRT_OK;
}

int swrt_flow_get_snapshot(swrt_flow* h, int slot, double* out_host) {
    if (!h || !out_host || slot < 0 || slot > 1) return fail(SWRT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(h->d.device));
    // (a slab-decomposed flow returns its own band: rows [rank ny/P, (rank+1) ny/P) as an (nx, ny/P, 5) array)
    const long long n = h->P > 1 ? (long long)h->d.nx * h->L.yrows : (long long)h->refine * h->d.nx * h->refine * h->d.ny;
    const int nc = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? SNAP3_NC : SNAP_NC, stride = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? SNAP3_STRIDE : SNAP_STRIDE;
    const double* src0 = h->snap[h->slot_map[slot]] + (h->P > 1 ? (long long)h->halo * h->d.nx * SNAP_STRIDE : 0);
    double* tmp = nullptr;
    CK(cudaMalloc(&tmp, sizeof(double) * n * nc));
    if (h->interp == SWRT_INTERP_BILINEAR_F32) {
        ProfScope ps(h, K_OTHER);
        snapf_to_planar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(reinterpret_cast<const float*>(src0), n, tmp);
    } else
    { ProfScope ps(h, K_OTHER); snap_to_planar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(src0, n, nc, stride, tmp); }
    cudaError_t e = cudaMemcpyAsync(out_host, tmp, sizeof(double) * n * nc, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cudaFree(tmp);
    CK(e);
    return SWRT_OK;
}

int swrt_flow_set_snapshot(swrt_flow* h, int slot, const double* in_host) {
    if (!h || !in_host || slot < 0 || slot > 1) return fail(SWRT_ERR_ARG, "bad argument");
    if (h->interp == SWRT_INTERP_BILINEAR_F32) return fail(SWRT_ERR_UNSUPPORTED, "snapshots of the fp32 packet mode are built by swrt_flow_velocity_snapshot only");
    CK(cudaSetDevice(h->d.device));
    // (a slab-decomposed flow takes its own band of rows, (nx, ny/P, 5); COLLECTIVE: the halo rows come from the neighbours)
    const long long n = h->P > 1 ? (long long)h->d.nx * h->L.yrows : (long long)h->refine * h->d.nx * h->refine * h->d.ny;
    const int nc = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? SNAP3_NC : SNAP_NC, stride = h->interp == SWRT_INTERP_HERMITE_BICUBIC ? SNAP3_STRIDE : SNAP_STRIDE;
    double* dst0 = h->snap[h->slot_map[slot]] + (h->P > 1 ? (long long)h->halo * h->d.nx * SNAP_STRIDE : 0);
    double* tmp = nullptr;
    CK(cudaMalloc(&tmp, sizeof(double) * n * nc));
    CK(wait_readers(h));
    cudaError_t e = cudaMemcpyAsync(tmp, in_host, sizeof(double) * n * nc, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) {
        ProfScope ps(h, K_OTHER);
        planar_to_snap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(tmp, n, nc, stride, dst0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cudaFree(tmp);
    CK(e);
    if (h->P > 1) {
        int rc;
        if ((rc = team_barrier(h))) return rc;
        if ((rc = band_halo_pull(h, slot))) return rc;
        if ((rc = team_barrier(h))) return rc;
        CK(cudaStreamSynchronize(h->st));
    }
    return SWRT_OK;
}

int swrt_flow_timer_start(swrt_flow* h) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    CK(cudaEventRecord(h->ev0, h->st));
    return SWRT_OK;
}
int swrt_flow_timer_stop(swrt_flow* h, float* ms) {
    if (!h || !ms) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    CK(cudaEventRecord(h->ev1, h->st));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return SWRT_OK;
}
int swrt_flow_sync(swrt_flow* h) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    CK(cudaStreamSynchronize(h->st));
    return SWRT_OK;
}
int swrt_flow_profile(swrt_flow* h, int enable) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(h->d.device));
    prof_collect(h);
    if (enable == 2) { for (int i = 0; i < 16; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; } enable = 1; }
    h->prof = enable != 0;
    return SWRT_OK;
}
int swrt_flow_profile_get(swrt_flow* h, int id, double* ms_total, long long* count, const char** name) {
    if (!h) return fail(SWRT_ERR_ARG, "null pointer");
    if (id < 0 || id >= K_COUNT) return fail(SWRT_ERR_ARG, "kernel id out of range");
    CK(cudaSetDevice(h->d.device));
    prof_collect(h);
    if (ms_total) *ms_total = h->prof_ms[id];
    if (count) *count = h->prof_n[id];
    if (name) *name = id == K_RAYTRACE ? h->ray_name : kKernelNames[id];
    return SWRT_OK;
}
int swrt_flow_launch_count(swrt_flow* h, long long* n) {
    if (!h || !n) return fail(SWRT_ERR_ARG, "null pointer");
    *n = h->launches;
    return SWRT_OK;
}

// ------------------------------------------------------------------ packets
static int packets_free(swrt_packets* p) {
    if (p->band) {
        for (int r = 0; r < (p->flow ? p->flow->P : 0); ++r)
            if (p->peer_arena[r] && p->peer_arena[r] != p->arena) cudaIpcCloseMemHandle(p->peer_arena[r]);
        cudaFree(p->arena);
        if (p->tab_host) cudaFreeHost(p->tab_host);
    } else {
        cudaFree(p->xk); cudaFree(p->sign); cudaFree(p->xk2); cudaFree(p->sign2); cudaFree(p->U); cudaFree(p->Gd);
        cudaFree(p->idx); cudaFree(p->idx2);
    }
    cudaFree(p->keys); cudaFree(p->hist); cudaFree(p->sums); cudaFree(p->count); cudaFree(p->rsched);
    return SWRT_OK;
}
int swrt_packets_destroy(swrt_packets* p) {
    if (!p) return SWRT_OK;
    if (p->flow) {
        cudaSetDevice(p->flow->d.device);
        cudaStreamSynchronize(pst(p));
        auto& rd = p->flow->readers;
        rd.erase(std::remove(rd.begin(), rd.end(), p), rd.end());
        p->flow->npackets--;
        for (auto& c : p->cycle) if (c.exec) cudaGraphExecDestroy(c.exec);
        if (p->own) cudaStreamDestroy(p->st);
        if (p->ev_done) cudaEventDestroy(p->ev_done);
        if (p->ev_io) cudaEventDestroy(p->ev_io);
    }
    packets_free(p);
    delete p;
    return SWRT_OK;
}

// band mode: layout of the IPC-shared arena (identical on every rank: same capacity), team.cuh PacketArena
static size_t arena_layout(long long cap, int P, PacketArena* a, char* base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return q; };
    for (int b = 0; b < 2; ++b) a->xk[b] = (double*)take(sizeof(double) * 4 * (size_t)cap);
    for (int b = 0; b < 2; ++b) a->sign[b] = (double*)take(sizeof(double) * (size_t)cap);
    for (int b = 0; b < 2; ++b) a->idx[b] = (unsigned*)take(sizeof(unsigned) * (size_t)cap);
    a->out = (double*)take(sizeof(double) * 6 * (size_t)cap);
    a->tab = (unsigned long long*)take(sizeof(unsigned long long) * (2 * (size_t)P + 8));
    return off;
}
static ArenaPeers arena_peers(const swrt_packets* p) {
    ArenaPeers ap{};
    for (int r = 0; r < p->flow->P; ++r) arena_layout(p->cap, p->flow->P, &ap.a[r], p->peer_arena[r]);
    return ap;
}
static int band_peers_ready(const swrt_packets* p) {
    for (int r = 0; r < p->flow->P; ++r)
        if (!p->peer_arena[r]) return fail(SWRT_ERR_STATE, "band packets: the arena of rank %d is not mapped (swrt_packets_ipc_open)", r);
    return SWRT_OK;
}

int swrt_packets_create(const swrt_packets_desc* desc, swrt_flow* flow, swrt_packets** out) {
    if (!desc || !flow || !out) return fail(SWRT_ERR_ARG, "null pointer");
    *out = nullptr;
    if (desc->n <= 0 || desc->n >= (1LL << 32)) return fail(SWRT_ERR_ARG, "n must be in [1, 2^32)");
    if (desc->interp < SWRT_INTERP_BILINEAR || desc->interp > SWRT_INTERP_NUFFT)
        return fail(SWRT_ERR_UNSUPPORTED, "interpolant %d not implemented", desc->interp);
    if (desc->interp == SWRT_INTERP_BILINEAR_F32 && desc->integrator != SWRT_INTEG_RK4)
        return fail(SWRT_ERR_UNSUPPORTED, "the fp32 packet mode integrates with RK4");
    if (desc->integrator != SWRT_INTEG_RK4 && desc->integrator != SWRT_INTEG_IMPLICIT_MIDPOINT)
        return fail(SWRT_ERR_UNSUPPORTED, "integrator %d not implemented", desc->integrator);
    if (desc->nsub < 1) return fail(SWRT_ERR_ARG, "nsub must be >= 1");
    if (desc->sort_every < 0) return fail(SWRT_ERR_ARG, "sort_every must be >= 0");
    const bool band = flow->P > 1;
    if (band && (desc->interp != SWRT_INTERP_BILINEAR || desc->integrator != SWRT_INTEG_RK4))
        return fail(SWRT_ERR_UNSUPPORTED, "packets of a slab-decomposed flow (y-band sharded) are built for the fp64 bilinear RK4 mode");
    if (band && (desc->sort_every < 1 || desc->band_capacity < desc->n || desc->band_first < 0 || desc->band_first + desc->n >= (1LL << 32)))
        return fail(SWRT_ERR_ARG, "band packets need sort_every >= 1 (the hand-over between bands rides on the sort), band_capacity >= n (the same on every rank) and a global row range below 2^32");
    CK(cudaSetDevice(flow->d.device));
    swrt_packets* p = new swrt_packets;
    p->d = *desc;
    p->flow = flow;
    flow->npackets++;
    p->nbins = (long long)flow->refine * flow->d.nx * flow->refine * flow->d.ny;   // cells of the snapshots' node grid
    p->band = band;
    p->cap = band ? desc->band_capacity : desc->n;
    p->ncur = band ? 0 : desc->n;
    p->first = band ? desc->band_first : 0;
    const size_t n = (size_t)p->cap;
    const size_t nsums = (size_t)(p->nbins / SCAN_BLOCK + 2) + (size_t)(p->nbins / SCAN_BLOCK / SCAN_BLOCK + 2) + 8;
    cudaError_t e = cudaSuccess;
    if (band) {
        PacketArena a{};
        const size_t bytes = arena_layout(p->cap, flow->P, &a, nullptr);
        e = cudaMalloc(&p->arena, bytes);
        if (e == cudaSuccess) e = cudaMemset(p->arena, 0, bytes);
        if (e == cudaSuccess) e = cudaHostAlloc(&p->tab_host, sizeof(unsigned long long) * 8, cudaHostAllocDefault);
        if (e == cudaSuccess) {
            arena_layout(p->cap, flow->P, &a, p->arena);
            p->xk = a.xk[0]; p->xk2 = a.xk[1]; p->sign = a.sign[0]; p->sign2 = a.sign[1]; p->idx = a.idx[0]; p->idx2 = a.idx[1];
            p->out6 = a.out; p->U = a.out; p->Gd = a.out + 2 * p->cap; p->tab = a.tab;
            p->peer_arena[flow->rank] = p->arena;
        }
    } else if ((e = cudaMalloc(&p->xk, sizeof(double) * 4 * n)) != cudaSuccess || (e = cudaMalloc(&p->sign, sizeof(double) * n)) != cudaSuccess ||
        (e = cudaMalloc(&p->xk2, sizeof(double) * 4 * n)) != cudaSuccess || (e = cudaMalloc(&p->sign2, sizeof(double) * n)) != cudaSuccess ||
        (e = cudaMalloc(&p->U, sizeof(double) * 2 * n)) != cudaSuccess || (e = cudaMalloc(&p->Gd, sizeof(double) * 4 * n)) != cudaSuccess ||
        (e = cudaMalloc(&p->idx, sizeof(unsigned) * n)) != cudaSuccess || (e = cudaMalloc(&p->idx2, sizeof(unsigned) * n)) != cudaSuccess) {
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->keys, sizeof(unsigned) * n);
    if (e == cudaSuccess) e = cudaMalloc(&p->hist, sizeof(unsigned) * (size_t)p->nbins);
    if (e == cudaSuccess) e = cudaMalloc(&p->sums, sizeof(unsigned) * nsums);
    if (e == cudaSuccess) e = cudaMalloc(&p->count, sizeof(unsigned long long) * 2);
    if (e == cudaSuccess) e = cudaMalloc(&p->rsched, sizeof(unsigned) * 2);
    if (e != cudaSuccess) {
        swrt_packets_destroy(p);
        return fail(SWRT_ERR_CUDA, "cudaMalloc(packets): %s", cudaGetErrorString(e));
    }
    CK(cudaMemsetAsync(p->count, 0, sizeof(unsigned long long) * 2, pst(p)));
    CK(cudaMemsetAsync(p->rsched, 0, sizeof(unsigned) * 2, pst(p)));
    if (!band) {
        CK(cudaMemsetAsync(p->xk, 0, sizeof(double) * 4 * n, pst(p)));
        CK(cudaMemsetAsync(p->sign, 0, sizeof(double) * n, pst(p)));
        { ProfScope ps(flow, K_OTHER, pst(p)); iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->idx, (long long)n); }
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(pst(p)));
    *out = p;
    return SWRT_OK;
}

// band mode: every rank exports the IPC handle of its arena and opens the peers' (like swrt_slab_ipc_handle / _open)
int swrt_packets_ipc_handle(swrt_packets* p, void* handle64) {
    if (!p || !handle64 || !p->band) return fail(SWRT_ERR_STATE, "not band-sharded packets (the flow is not slab-decomposed)");
    CK(cudaSetDevice(p->flow->d.device));
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, p->arena));
    memcpy(handle64, &mh, 64);
    return SWRT_OK;
}
int swrt_packets_ipc_open(swrt_packets* p, int peer_rank, const void* handle64) {
    if (!p || !handle64 || !p->band) return fail(SWRT_ERR_STATE, "not band-sharded packets (the flow is not slab-decomposed)");
    if (peer_rank < 0 || peer_rank >= p->flow->P) return fail(SWRT_ERR_ARG, "peer rank out of range");
    CK(cudaSetDevice(p->flow->d.device));
    if (peer_rank == p->flow->rank) { p->peer_arena[peer_rank] = p->arena; return SWRT_OK; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle64, 64);
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, mh, cudaIpcMemLazyEnablePeerAccess));
    p->peer_arena[peer_rank] = (char*)q;
    return SWRT_OK;
}
// resident packets on this rank (band mode; = n otherwise)
int swrt_packets_resident(swrt_packets* p, long long* n) {
    if (!p || !n) return fail(SWRT_ERR_ARG, "null pointer");
    *n = p->ncur;
    return SWRT_OK;
}

// host (n, ncol) column-major with leading dimension ld  <->  device [ncol][ldd]
static cudaError_t copy_cols(void* dst, const void* src, long long n, int ncol, long long ld, cudaMemcpyKind kind, cudaStream_t st, long long ldd = -1) {
    if (ldd < 0) ldd = n;
    if (ld == n && ldd == n) return cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)ncol * (size_t)n, kind, st);
    const size_t hp = sizeof(double) * (size_t)ld, dp = sizeof(double) * (size_t)ldd, w = sizeof(double) * (size_t)n;
    return kind == cudaMemcpyHostToDevice ? cudaMemcpy2DAsync(dst, dp, src, hp, w, (size_t)ncol, kind, st)
                                          : cudaMemcpy2DAsync(dst, hp, src, dp, w, (size_t)ncol, kind, st);
}
// Packets with their own stream read snapshots the flow's stream writes: order the two streams with events.
static cudaError_t wait_flow(swrt_packets* p) {
    if (!p->own) return cudaSuccess;
    swrt_flow* f = p->flow;
    cudaError_t e = cudaEventRecord(f->ev_sync, f->st);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(p->st, f->ev_sync, 0);
}
static cudaError_t mark_read(swrt_packets* p) { return p->own ? cudaEventRecord(p->ev_done, p->st) : cudaSuccess; }
// Asynchronous host transfers of handles with their own stream run on ONE upload and ONE download stream per flow, in call order.
// Copies issued on many streams are spread over the copy engines and share the PCIe link: the uploads of all row blocks of a
// PacketPipeline step then finish together at the end and no download overlaps them (measured: 12 blocks 15.1 ms per step
// against 11.3 ms for the two directions side by side, profiles/r02_r).  io_begin orders the transfer stream behind the handle's
// stream and returns it; io_end makes the handle's stream (and so swrt_packets_sync) wait for the transfer.
static cudaError_t io_begin(swrt_packets* p, bool up, cudaStream_t* out) {
    swrt_flow* f = p->flow;
    cudaStream_t& s = up ? f->io_up : f->io_down;
    cudaError_t e = cudaSuccess;
    if (!s) e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventRecord(p->ev_io, p->st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, p->ev_io, 0);
    *out = s;
    return e;
}
static cudaError_t io_end(swrt_packets* p, cudaStream_t s) {
    cudaError_t e = cudaEventRecord(p->ev_io, s);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(p->st, p->ev_io, 0);
}

int swrt_packets_use_own_stream(swrt_packets* p) {
    if (!p) return fail(SWRT_ERR_ARG, "null pointer");
    if (p->own) return SWRT_OK;
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    CK(cudaStreamSynchronize(f->st));
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&p->st, cudaStreamNonBlocking, lo));      // least priority: fills what the flow's stream leaves
    }
    CK(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->ev_io, cudaEventDisableTiming));
    p->own = true;
    f->readers.push_back(p);
    return SWRT_OK;
}

int swrt_packets_set_kernel(swrt_packets* p, int kernel) {
    if (!p) return fail(SWRT_ERR_ARG, "null pointer");
    if (kernel < SWRT_RAYKERNEL_AUTO || kernel > SWRT_RAYKERNEL_PIPE) return fail(SWRT_ERR_ARG, "unknown ray kernel %d", kernel);
    p->kernel_sel = kernel;
    for (auto& c : p->cycle) if (c.exec) { cudaGraphExecDestroy(c.exec); c.exec = nullptr; }   // captured launches name the old kernel
    return SWRT_OK;
}

int swrt_packets_sync(swrt_packets* p) {
    if (!p) return fail(SWRT_ERR_ARG, "null pointer");
    CK(cudaSetDevice(p->flow->d.device));
    CK(cudaStreamSynchronize(pst(p)));
    return SWRT_OK;
}

static PacketGrid packet_grid(const swrt_flow* f, const swrt_packets* p = nullptr) {
    PacketGrid g{};
    g.nx = f->refine * f->d.nx; g.ny = f->refine * f->d.ny;    // the node grid of the snapshots
    g.dx = f->d.Lx / g.nx; g.dy = f->d.Ly / g.ny;
    g.x0 = -f->d.Lx / 2; g.y0 = -f->d.Ly / 2;
    g.inv_dx = 1.0 / g.dx; g.inv_dy = 1.0 / g.dy;
    g.ld = p ? p->cap : 0;
    g.nw = f->nufft_w;
    g.nbeta = nufft_beta(f->nufft_w);
    if (f->P > 1) {
        g.band = 1;
        g.jb = f->rank * f->L.yrows - f->halo;
        g.jrows = f->L.yrows + 2 * f->halo;
        const int tiles_y = g.ny >> TILE_SHIFT;
        int t_lo = g.jb >= 0 ? g.jb >> TILE_SHIFT : -((-g.jb + TILE - 1) >> TILE_SHIFT);
        g.tile_row0 = ((t_lo % tiles_y) + tiles_y) % tiles_y;
    }
    return g;
}
// tile rows the tile kernel's grid covers
static int tile_rows(const swrt_flow* f, const PacketGrid& g) {
    const int tiles_y = g.ny >> TILE_SHIFT;
    if (!g.band) return tiles_y;
    const int t_lo = g.jb >= 0 ? g.jb >> TILE_SHIFT : -((-g.jb + TILE - 1) >> TILE_SHIFT);
    const int t_hi = (g.jb + g.jrows - 1) >> TILE_SHIFT;
    const int n = t_hi - t_lo + 1;
    return n < tiles_y ? n : tiles_y;
}

// ---- band mode: scatter the caller-order staging block (5 columns in `out`: x, y, k, l, sign) to the band owners.  COLLECTIVE.
static int band_scatter(swrt_packets* p, long long nrows) {
    swrt_flow* f = p->flow;
    int rc = band_peers_ready(p);
    if (rc) return rc;
    const int P = f->P;
    unsigned long long hdr[4] = {0ULL, (unsigned long long)nrows, (unsigned long long)p->first, 0ULL};   // resident count, staging rows, first row, overflow
    CK(cudaMemcpyAsync(p->tab + 2 * P, hdr, sizeof hdr, cudaMemcpyHostToDevice, pst(p)));
    if ((rc = team_barrier(f, 1, pst(p)))) return rc;                                  // every staging block and header is in place
    int bshift = 0; while ((1 << bshift) < f->L.yrows) ++bshift;
    { ProfScope ps(f, K_OTHER, pst(p)); team_scatter_scan_kernel<<<148 * 4, 256, 0, pst(p)>>>(arena_peers(p), P, f->rank, p->cur, p->cap, packet_grid(f, p), bshift); }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p->tab_host, p->tab + 2 * P, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, pst(p)));
    if ((rc = team_barrier(f, 1, pst(p)))) return rc;                                  // the staging blocks may be overwritten again
    CK(cudaStreamSynchronize(pst(p)));
    if (p->tab_host[3]) return fail(SWRT_ERR_STATE, "band packets: more than band_capacity = %lld packets fall into the band of rank %d", p->cap, f->rank);
    p->ncur = (long long)p->tab_host[0];
    p->permuted = true;
    p->since_sort = 1 << 30;
    p->tiles_valid = false;
    return SWRT_OK;
}
// ---- band mode: collect this rank's caller-order rows from wherever the packets live.  which = 0: packet state, 1: `out`
// columns [c0, c0 + ncols) in resident order.  The result lands in dst (device, [ncols][ldd]).  COLLECTIVE; the caller
// places the barriers.
static int band_gather_launch(swrt_packets* p, int which, int c0, int ncols, double* dst, long long ldd) {
    swrt_flow* f = p->flow;
    ArenaPeers ap = arena_peers(p);
    if (which == 1) for (int r = 0; r < f->P; ++r) ap.a[r].out += (long long)c0 * p->cap;
    { ProfScope ps(f, K_OTHER, pst(p)); team_gather_scan_kernel<<<148 * 4, 256, 0, pst(p)>>>(ap, f->P, p->cur, which, ncols, p->cap, p->first, p->d.n, dst, ldd); }
    CK(cudaGetLastError());
    return SWRT_OK;
}

static int packets_set_impl(swrt_packets* p, const double* xk_host, long long ld, const double* sign_host, bool sync) {
    if (!p || !xk_host) return fail(SWRT_ERR_ARG, "null pointer");
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    const long long n = p->d.n;
    if (ld < n) return fail(SWRT_ERR_ARG, "leading dimension %lld < n", ld);
    if (p->band) {
        // caller-order block -> staging (x, y, k, l, sign); the signs of the previous ensemble are fetched first when none are given
        if (!sign_host) {
            int rc = band_peers_ready(p);
            if (rc) return rc;
            if ((rc = team_barrier(f, 1, pst(p)))) return rc;
            ArenaPeers ap = arena_peers(p);
            for (int r = 0; r < f->P; ++r) ap.a[r].out = ap.a[r].sign[p->cur];          // gather column: the resident signs
            { ProfScope ps(f, K_OTHER, pst(p)); team_gather_scan_kernel<<<148 * 4, 256, 0, pst(p)>>>(ap, f->P, p->cur, 1, 1, p->cap, p->first, n, p->xk2, p->cap); }
            CK(cudaGetLastError());
            if ((rc = team_barrier(f, 1, pst(p)))) return rc;
            CK(cudaMemcpyAsync(p->out6 + 4 * p->cap, p->xk2, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, pst(p)));
        } else {
            CK(cudaMemcpyAsync(p->out6 + 4 * p->cap, sign_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, pst(p)));
        }
        CK(copy_cols(p->out6, xk_host, n, 4, ld, cudaMemcpyHostToDevice, pst(p), p->cap));
        return band_scatter(p, n);
    }
    if (!sign_host && p->permuted) {   // keep the frequency signs: bring them back to the caller's order first
        { ProfScope ps(f, K_OTHER, pst(p)); unpermute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->sign, p->idx, n, n, 1, p->sign2); }
        CK(cudaGetLastError());
        std::swap(p->sign, p->sign2);
    }
    cudaStream_t cs = pst(p);
    const bool via_io = p->own && !sync;
    if (via_io) CK(io_begin(p, true, &cs));
    CK(copy_cols(p->xk, xk_host, n, 4, ld, cudaMemcpyHostToDevice, cs));
    if (sign_host) CK(cudaMemcpyAsync(p->sign, sign_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, cs));
    if (via_io) CK(io_end(p, cs));
    { ProfScope ps(f, K_OTHER, pst(p)); iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->idx, n); }
    CK(cudaGetLastError());
    p->permuted = false;
    p->since_sort = 1 << 30;
    p->tiles_valid = false;
    if (sync) CK(cudaStreamSynchronize(pst(p)));
    return SWRT_OK;
}
int swrt_packets_set(swrt_packets* p, const double* xk_host, const double* sign_host) {
    return packets_set_impl(p, xk_host, p ? p->d.n : 0, sign_host, true);
}
int swrt_packets_set_async(swrt_packets* p, const double* xk_host, long long ld, const double* sign_host) {
    return packets_set_impl(p, xk_host, ld, sign_host, false);
}

static int packets_get_impl(swrt_packets* p, double* xk_host, long long ld, bool sync) {
    if (!p || !xk_host) return fail(SWRT_ERR_ARG, "null pointer");
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    const long long n = p->d.n;
    if (ld < n) return fail(SWRT_ERR_ARG, "leading dimension %lld < n", ld);
    if (p->band) {   // COLLECTIVE: every rank pulls its caller-order rows out of all ranks' resident arrays
        int rc = band_peers_ready(p);
        if (rc) return rc;
        if ((rc = team_barrier(f, 1, pst(p)))) return rc;
        if ((rc = band_gather_launch(p, 0, 0, 4, p->out6, p->cap))) return rc;
        if ((rc = team_barrier(f, 1, pst(p)))) return rc;
        CK(copy_cols(xk_host, p->out6, n, 4, ld, cudaMemcpyDeviceToHost, pst(p), p->cap));
        CK(cudaStreamSynchronize(pst(p)));
        return SWRT_OK;
    }
    const double* src = p->xk;
    if (p->permuted) {
        { ProfScope ps(f, K_OTHER, pst(p)); unpermute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->xk, p->idx, n, n, 4, p->xk2); }
        CK(cudaGetLastError());
        src = p->xk2;
    }
    cudaStream_t cs = pst(p);
    const bool via_io = p->own && !sync;
    if (via_io) CK(io_begin(p, false, &cs));
    CK(copy_cols(xk_host, src, n, 4, ld, cudaMemcpyDeviceToHost, cs));
    if (via_io) CK(io_end(p, cs));
    if (sync) CK(cudaStreamSynchronize(pst(p)));
    return SWRT_OK;
}
int swrt_packets_get(swrt_packets* p, double* xk_host) { return packets_get_impl(p, xk_host, p ? p->d.n : 0, true); }
int swrt_packets_get_async(swrt_packets* p, double* xk_host, long long ld) { return packets_get_impl(p, xk_host, ld, false); }

int swrt_packets_generate(swrt_packets* p, double L, double k0, long long sqrtN, long long first) {
    if (!p || sqrtN <= 0 || first < 0 || first + p->d.n > sqrtN * sqrtN) return fail(SWRT_ERR_ARG, "bad argument");
    CK(cudaSetDevice(p->flow->d.device));
    const long long n = p->d.n;
    if (p->band) {   // the caller-order block of the lattice goes to the staging columns, then to the band owners.  COLLECTIVE.
        if (first != p->first) return fail(SWRT_ERR_ARG, "band packets: `first` (%lld) must equal the band_first the handle was created with (%lld)", first, p->first);
        { ProfScope ps(p->flow, K_OTHER); generate_packets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, p->flow->st>>>(p->out6, p->out6 + 4 * p->cap, p->idx2, n, p->cap, first, sqrtN, L, k0); }
        CK(cudaGetLastError());
        return band_scatter(p, n);
    }
    { ProfScope ps(p->flow, K_OTHER, pst(p)); generate_packets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->xk, p->sign, p->idx, n, n, first, sqrtN, L, k0); }
    CK(cudaGetLastError());
    p->permuted = false;
    p->since_sort = 1 << 30;
    p->tiles_valid = false;
    return SWRT_OK;
}

static cudaError_t exclusive_scan(swrt_packets* p, unsigned* a, long long nb, unsigned* scratch) {
    swrt_flow* f = p->flow;
    const unsigned blocks = (unsigned)((nb + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (blocks <= 1) {
        ProfScope ps(f, K_SORT, pst(p));
        scan_block_kernel<<<1, SCAN_BLOCK, 0, pst(p)>>>(a, nb, nullptr);
        return cudaGetLastError();
    }
    { ProfScope ps(f, K_SORT, pst(p)); scan_block_kernel<<<blocks, SCAN_BLOCK, 0, pst(p)>>>(a, nb, scratch); }
    cudaError_t e = exclusive_scan(p, scratch, blocks, scratch + blocks);
    if (e != cudaSuccess) return e;
    { ProfScope ps(f, K_SORT, pst(p)); scan_add_kernel<<<blocks, SCAN_BLOCK, 0, pst(p)>>>(a, nb, scratch); }
    return cudaGetLastError();
}

// counting sort of the packets by tiled cell key (packets.cuh); out of place, then swap the buffers
static int sort_local(swrt_packets* p) {
    swrt_flow* f = p->flow;
    const long long n = p->ncur;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    CK(cudaMemsetAsync(p->hist, 0, sizeof(unsigned) * (size_t)p->nbins, pst(p)));
    if (n > 0) {
        { ProfScope ps(f, K_SORT, pst(p)); sort_hist_kernel<<<blocks, 256, 0, pst(p)>>>(p->xk, n, packet_grid(f, p), p->keys, p->hist, p->count + 1); }
        CK(cudaGetLastError());
    }
    CK(exclusive_scan(p, p->hist, p->nbins, p->sums));
    if (n > 0) {
        { ProfScope ps(f, K_SORT, pst(p)); sort_scatter_kernel<<<blocks, 256, 0, pst(p)>>>(p->xk, p->sign, p->idx, p->keys, p->hist, n, p->cap, p->xk2, p->sign2, p->idx2); }
        CK(cudaGetLastError());
    }
    std::swap(p->xk, p->xk2);
    std::swap(p->sign, p->sign2);
    std::swap(p->idx, p->idx2);
    p->cur ^= 1;
    p->permuted = true;
    p->since_sort = 0;
    p->tiles_valid = true;
    return SWRT_OK;
}
static int sort_packets(swrt_packets* p) {
    int rc = sort_local(p);
    if (rc || !p->band) return rc;
    // ---- hand-over between bands (COLLECTIVE): the sorted order groups the packets by owner rank (the key is tile-row major)
    swrt_flow* f = p->flow;
    const int P = f->P;
    if ((rc = band_peers_ready(p))) return rc;
    const ArenaPeers ap = arena_peers(p);
    const long long keys_per_rank = (long long)f->L.yrows * f->d.nx;
    { ProfScope ps(f, K_SORT, pst(p)); team_publish_segments_kernel<<<1, 32, 0, pst(p)>>>(p->hist, keys_per_rank, ap, P, f->rank); }
    CK(cudaGetLastError());
    if ((rc = team_barrier(f, 1, pst(p)))) return rc;                                  // every rank's sorted array and segment table are complete
    { ProfScope ps(f, K_SORT, pst(p)); team_pull_segments_kernel<<<148 * 4, 256, 0, pst(p)>>>(ap, P, f->rank, p->cur, p->cur ^ 1, p->cap); }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(p->tab_host, p->tab + 2 * P, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, pst(p)));
    CK(cudaMemcpyAsync(p->tab_host + 4, p->count + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, pst(p)));
    if ((rc = team_barrier(f, 1, pst(p)))) return rc;                                  // the peers have read my sorted array: it may be reused
    CK(cudaStreamSynchronize(pst(p)));                                       // the new resident count sizes the next launches
    if (p->tab_host[3]) return fail(SWRT_ERR_STATE, "band packets: more than band_capacity = %lld packets moved into the band of rank %d", p->cap, f->rank);
    if (p->tab_host[4]) return fail(SWRT_ERR_STATE, "band packets: %llu packets left band + halo (%d rows) between two hand-overs; lower sort_every", p->tab_host[4], f->halo);
    std::swap(p->xk, p->xk2);
    std::swap(p->sign, p->sign2);
    std::swap(p->idx, p->idx2);
    p->cur ^= 1;
    p->ncur = (long long)p->tab_host[0];
    return sort_local(p);                                                   // merge the P sorted runs; rebuilds the tile offsets
}

int swrt_packets_raytrace(swrt_packets* p, double t0, double t1) {
    if (!p) return fail(SWRT_ERR_ARG, "null pointer");
    if (!(t1 != t0)) return fail(SWRT_ERR_ARG, "t1 must differ from t0");
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    if (p->nbins != (long long)f->refine * f->d.nx * f->refine * f->d.ny)
        return fail(SWRT_ERR_STATE, "the snapshot refinement changed after these packets were created");
    if (p->d.sort_every > 0 && p->since_sort >= p->d.sort_every) {
        int rc = sort_packets(p);
        if (rc) return rc;
    }
    if (p->d.interp != f->interp) return fail(SWRT_ERR_STATE, "packets use interpolant %d but the flow's snapshots hold node data for %d (swrt_flow_set_interp)", p->d.interp, f->interp);
    CK(wait_flow(p));
    RayParams rp{p->d.f, p->d.Cg, t0, t1, p->d.nsub, p->d.time_lerp};
    const double *So = f->snap[f->slot_map[0]], *Sn = f->snap[f->slot_map[1]];
    const long long n = p->ncur;
    p->since_sort++;
    if (n == 0) return SWRT_OK;
    static const int minb = [] { const char* e = getenv("SWRT_RAYTRACE_MINB"); return e ? atoi(e) : 5; }();  // tuning knobs
    static const int cached = [] { const char* e = getenv("SWRT_RAYTRACE_CACHE"); return e ? atoi(e) : 4; }();   // 0 = plain kernel, 3 / 4 = stencil-cached kernel with that many CTAs per SM
    const unsigned grid = (unsigned)((n + 127) / 128);
    // TMA-staged tile kernel (packets.cuh): fp64 bilinear RK4, packets in sorted order, enough packets per tile to pay for the patch
    static const int tile_mode = [] { const char* e = getenv("SWRT_RAYTRACE_TILE"); return e ? atoi(e) : 1; }();
    static const int tile_minb = [] { const char* e = getenv("SWRT_RAYTRACE_TILE_MINB"); return e ? atoi(e) : 3; }();   // 3 CTAs x 128 threads: 162 registers, no spills (0.81 ms vs 1.04 ms with 4 CTAs at 128 registers + spills, profiles/r02_a)
    static const int tile_min_pk = [] { const char* e = getenv("SWRT_RAYTRACE_TILE_MINPK"); return e ? atoi(e) : 192; }();
    const PacketGrid pg = packet_grid(f, p);
    const long long ntiles = (long long)(pg.nx >> TILE_SHIFT) * tile_rows(f, pg);
    const bool want_tile = p->kernel_sel == SWRT_RAYKERNEL_TILE || p->kernel_sel == SWRT_RAYKERNEL_TILE3 || p->kernel_sel == SWRT_RAYKERNEL_PIPE ||
                           (p->kernel_sel == SWRT_RAYKERNEL_AUTO && tile_mode > 0 && n >= ntiles * (long long)tile_min_pk);
    const bool use_tile = want_tile && p->d.interp == SWRT_INTERP_BILINEAR && p->d.integrator == SWRT_INTEG_RK4 && p->tiles_valid &&
                          f->tmap_ok && ntiles > 0;
    // one RK4 step per call: the three-level kernel (first level, mean, last level in shared memory), one CTA per tile (default);
    // SWRT_RAYTRACE_TILE=2 keeps the two-level kernel, =4 selects the persistent warp-specialised variant (measured slower:
    // 0.62-0.68 ms against 0.52 ms, profiles/r02_o_ray_kernel_ab.log)
    const bool use_pipe = use_tile && p->d.nsub == 1 && (p->kernel_sel == SWRT_RAYKERNEL_PIPE || (p->kernel_sel == SWRT_RAYKERNEL_AUTO && tile_mode == 4));
    const bool use_tile3 = use_tile && p->d.nsub == 1 && (p->kernel_sel == SWRT_RAYKERNEL_TILE3 || (p->kernel_sel == SWRT_RAYKERNEL_AUTO && (tile_mode == 1 || tile_mode == 3)));
    // fp32 packet mode, one RK4 step per call: its own three-level staged kernel (SWRT_RAYTRACE_F32_TILE = its CTAs per SM, 2 or 3;
    // 0: the stencil-cached fp32 kernel)
    static const int f32_tile = [] { const char* e = getenv("SWRT_RAYTRACE_F32_TILE"); return e ? atoi(e) : 3; }();   // CTAs per SM: 3 (80 registers) 0.412 ms, 2 (114) 0.442 ms, cached fp32 kernel 0.58 ms
    const bool use_tile3f = want_tile && f32_tile > 0 && p->d.interp == SWRT_INTERP_BILINEAR_F32 && p->d.integrator == SWRT_INTEG_RK4 && p->d.nsub == 1 &&
                            p->tiles_valid && f->tmap_ok && ntiles > 0;
    if (use_tile3f) {
        static bool attr_f_done = false;
        if (!attr_f_done) {
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_f32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE3F_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_f32_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_f32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE3F_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_f32_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_f_done = true;
        }
    }
    if (use_tile) {
        static bool attr_done = false;
        if (!attr_done) {
            CK(cudaFuncSetAttribute(raytrace_rk4_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE3_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile3_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CK(cudaFuncSetAttribute(raytrace_rk4_tile_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            attr_done = true;
        }
    }
    f->ray_name = use_tile3f ? "raytrace_rk4_tile3_f32_kernel" : p->d.interp == SWRT_INTERP_BILINEAR_F32 ? "raytrace_rk4_f32_kernel"
                : (p->d.integrator == SWRT_INTEG_IMPLICIT_MIDPOINT || p->d.interp == SWRT_INTERP_BSPLINE2 || p->d.interp == SWRT_INTERP_BSPLINE3 || p->d.interp == SWRT_INTERP_NUFFT) ? "raytrace_generic_kernel"
                : p->d.interp == SWRT_INTERP_HERMITE_BICUBIC ? "raytrace_rk4_cubic_kernel"
                : use_pipe ? "raytrace_rk4_pipe_kernel"
                : use_tile3 ? "raytrace_rk4_tile3_kernel"
                : use_tile ? (tile_minb >= 4 ? "raytrace_rk4_tile_kernel<4>" : "raytrace_rk4_tile_kernel<3>")
                : (cached || p->kernel_sel == SWRT_RAYKERNEL_CACHED) ? "raytrace_rk4_cached_kernel<4>" : "raytrace_rk4_kernel";
    { ProfScope ps(f, K_RAYTRACE, pst(p));
#define SWRT_GEN(I, G) raytrace_generic_kernel<I, G><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp)
      if (use_tile3f) {
          const int first = p->d.time_lerp == 0 ? 0 : 1;          // the level whose weight is 1 at t0
          const float4 *F1 = reinterpret_cast<const float4*>(first == 0 ? So : Sn), *F4 = reinterpret_cast<const float4*>(first == 0 ? Sn : So);
          if (f32_tile >= 3) raytrace_rk4_tile3_f32_kernel<3><<<(unsigned)ntiles, TILE3_THREADS, (size_t)TILE3F_SMEM_BYTES, pst(p)>>>(
              p->xk, p->sign, F1, F4, f->tmapf[f->slot_map[first]], f->tmapf[f->slot_map[first ^ 1]], p->hist, pg, rp);
          else raytrace_rk4_tile3_f32_kernel<2><<<(unsigned)ntiles, TILE3_THREADS, (size_t)TILE3F_SMEM_BYTES, pst(p)>>>(
              p->xk, p->sign, F1, F4, f->tmapf[f->slot_map[first]], f->tmapf[f->slot_map[first ^ 1]], p->hist, pg, rp);
      }
      else if (p->d.interp == SWRT_INTERP_BILINEAR_F32) {
          static const int fminb = [] { const char* e = getenv("SWRT_RAYTRACE_F32_MINB"); return e ? atoi(e) : 6; }();
          const float4 *Fo = reinterpret_cast<const float4*>(So), *Fn = reinterpret_cast<const float4*>(Sn);
          if (fminb <= 4) raytrace_rk4_f32_kernel<4><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, Fo, Fn, pg, rp);
          else if (fminb == 5) raytrace_rk4_f32_kernel<5><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, Fo, Fn, pg, rp);
          else if (fminb == 6) raytrace_rk4_f32_kernel<6><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, Fo, Fn, pg, rp);
          else raytrace_rk4_f32_kernel<8><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, Fo, Fn, pg, rp);
      }
      else if (p->d.integrator == SWRT_INTEG_IMPLICIT_MIDPOINT) {
          if (p->d.interp == 0) SWRT_GEN(0, 1); else if (p->d.interp == 1) SWRT_GEN(1, 1); else if (p->d.interp == 2) SWRT_GEN(2, 1); else if (p->d.interp == 5) SWRT_GEN(5, 1); else SWRT_GEN(4, 1);
      }
      else if (p->d.interp == SWRT_INTERP_NUFFT) SWRT_GEN(5, 0);
      else if (p->d.interp == SWRT_INTERP_BSPLINE2) SWRT_GEN(2, 0);
      else if (p->d.interp == SWRT_INTERP_BSPLINE3) SWRT_GEN(4, 0);
#undef SWRT_GEN
      else if (p->d.interp == SWRT_INTERP_HERMITE_BICUBIC) raytrace_rk4_cubic_kernel<<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp);
      else if (use_pipe) {
          const int first = p->d.time_lerp == 0 ? 0 : 1;          // the level whose weight is 1 at t0
          static const int sms = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
          const unsigned ctas = (unsigned)std::min<long long>(ntiles, sms);
          raytrace_rk4_pipe_kernel<<<ctas, PIPE_THREADS, (size_t)PIPE_SMEM_BYTES, pst(p)>>>(
              p->xk, p->sign, n, first == 0 ? So : Sn, first == 0 ? Sn : So, f->tmap[f->slot_map[first]], f->tmap[f->slot_map[first ^ 1]], p->hist,
              (int)ntiles, p->rsched, pg, rp);
      }
      else if (use_tile3) {
          const int first = p->d.time_lerp == 0 ? 0 : 1;          // the level whose weight is 1 at t0
          raytrace_rk4_tile3_kernel<<<(unsigned)ntiles, TILE3_THREADS, (size_t)TILE3_SMEM_BYTES, pst(p)>>>(
              p->xk, p->sign, n, first == 0 ? So : Sn, first == 0 ? Sn : So, f->tmap[f->slot_map[first]], f->tmap[f->slot_map[first ^ 1]], p->hist, pg, rp);
      }
      else if (use_tile) {
          const size_t smem = (size_t)TILE_SMEM_BYTES;
          if (tile_minb >= 4) raytrace_rk4_tile_kernel<4><<<(unsigned)ntiles, TILE_THREADS, smem, pst(p)>>>(p->xk, p->sign, n, So, Sn, f->tmap[f->slot_map[0]], f->tmap[f->slot_map[1]], p->hist, pg, rp);
          else raytrace_rk4_tile_kernel<3><<<(unsigned)ntiles, TILE_THREADS, smem, pst(p)>>>(p->xk, p->sign, n, So, Sn, f->tmap[f->slot_map[0]], f->tmap[f->slot_map[1]], p->hist, pg, rp);
      }
      else if (cached == 3 && p->kernel_sel == SWRT_RAYKERNEL_AUTO) raytrace_rk4_cached_kernel<3><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp);
      else if (cached || p->kernel_sel == SWRT_RAYKERNEL_CACHED) raytrace_rk4_cached_kernel<4><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp);
      else if (minb <= 4) raytrace_rk4_kernel<4><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp);
      else if (minb == 5) raytrace_rk4_kernel<5><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp);
      else raytrace_rk4_kernel<6><<<grid, 128, 0, pst(p)>>>(p->xk, p->sign, n, So, Sn, pg, rp); }
    CK(cudaGetLastError());
    CK(mark_read(p));
    return SWRT_OK;
}

static int packets_sample_impl(swrt_packets* p, int slot, double* u_host, double* g_host, long long ld, bool sync) {
    if (!p || !u_host || slot < 0 || slot > 1) return fail(SWRT_ERR_ARG, "bad argument");
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    const long long n = p->d.n, nres = p->ncur;
    if (ld < n) return fail(SWRT_ERR_ARG, "leading dimension %lld < n", ld);
    CK(wait_flow(p));
    if (p->d.interp != f->interp) return fail(SWRT_ERR_STATE, "packets use interpolant %d but the flow's snapshots hold node data for %d", p->d.interp, f->interp);
    const PacketGrid pg = packet_grid(f, p);
    const unsigned* idx = p->band ? nullptr : p->idx;          // band mode: resident order, brought to the caller's order by the gather
    const unsigned grid = (unsigned)((nres + 127) / 128);
    const double* S = f->snap[f->slot_map[slot]];
    if (nres > 0) {
        ProfScope ps(f, K_SAMPLE, pst(p));
        if (p->d.interp == SWRT_INTERP_BILINEAR_F32) sample_f32_kernel<<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, reinterpret_cast<const float4*>(S), pg, p->U, g_host ? p->Gd : nullptr, p->cap);
        else if (p->d.interp == SWRT_INTERP_NUFFT) sample_generic_kernel<5><<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, S, pg, p->U, g_host ? p->Gd : nullptr, p->cap);
        else if (p->d.interp == SWRT_INTERP_BSPLINE3) sample_generic_kernel<4><<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, S, pg, p->U, g_host ? p->Gd : nullptr, p->cap);
        else if (p->d.interp == SWRT_INTERP_BSPLINE2) sample_generic_kernel<2><<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, S, pg, p->U, g_host ? p->Gd : nullptr, p->cap);
        else if (p->d.interp == SWRT_INTERP_HERMITE_BICUBIC) sample_cubic_kernel<<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, S, pg, p->U, g_host ? p->Gd : nullptr, p->cap);
        else sample_kernel<<<grid, 128, 0, pst(p)>>>(p->xk, idx, nres, S, pg, p->U, g_host ? p->Gd : nullptr, p->cap);
    }
    CK(cudaGetLastError());
    CK(mark_read(p));
    if (p->band) {   // COLLECTIVE: resident-order samples -> caller-order rows (through the alternate state buffer, 4 columns at a time)
        int rc = band_peers_ready(p);
        if (rc) return rc;
        if ((rc = team_barrier(f, 1, pst(p)))) return rc;
        if ((rc = band_gather_launch(p, 1, 0, 2, p->xk2, p->cap))) return rc;
        CK(copy_cols(u_host, p->xk2, n, 2, ld, cudaMemcpyDeviceToHost, pst(p), p->cap));
        if (g_host) {
            if ((rc = band_gather_launch(p, 1, 2, 4, p->xk2, p->cap))) return rc;
            CK(copy_cols(g_host, p->xk2, n, 4, ld, cudaMemcpyDeviceToHost, pst(p), p->cap));
        }
        if ((rc = team_barrier(f, 1, pst(p)))) return rc;
        CK(cudaStreamSynchronize(pst(p)));
        return SWRT_OK;
    }
    cudaStream_t cs = pst(p);
    const bool via_io = p->own && !sync;
    if (via_io) CK(io_begin(p, false, &cs));
    CK(copy_cols(u_host, p->U, n, 2, ld, cudaMemcpyDeviceToHost, cs));
    if (g_host) CK(copy_cols(g_host, p->Gd, n, 4, ld, cudaMemcpyDeviceToHost, cs));
    if (via_io) CK(io_end(p, cs));
    if (sync) CK(cudaStreamSynchronize(pst(p)));
    return SWRT_OK;
}
int swrt_packets_sample(swrt_packets* p, int slot, double* u_host, double* g_host) {
    return packets_sample_impl(p, slot, u_host, g_host, p ? p->d.n : 0, true);
}
int swrt_packets_sample_async(swrt_packets* p, int slot, double* u_host, double* g_host, long long ld) {
    return packets_sample_impl(p, slot, u_host, g_host, ld, false);
}

int swrt_packets_kcutoff_reset(swrt_packets* p, double kcut, double k0, long long* nreset) {
    if (!p) return fail(SWRT_ERR_ARG, "null pointer");
    swrt_flow* f = p->flow;
    CK(cudaSetDevice(f->d.device));
    const long long n = p->ncur;     // (band mode: the resident packets; the count is this rank's share)
    CK(cudaMemsetAsync(p->count, 0, sizeof(unsigned long long), pst(p)));
    if (n > 0) {
        ProfScope ps(f, K_OTHER, pst(p));
        kcutoff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pst(p)>>>(p->xk, n, p->cap, kcut * kcut, k0, p->count);
    }
    CK(cudaGetLastError());
    if (nreset) {   // the count is only fetched (and the stream only synchronised) when the caller asks for it
        unsigned long long c = 0;
        CK(cudaMemcpyAsync(&c, p->count, sizeof c, cudaMemcpyDeviceToHost, pst(p)));
        CK(cudaStreamSynchronize(pst(p)));
        *nreset = (long long)c;
    }
    return SWRT_OK;
}

int swrt_packets_coupled_steps(swrt_packets* p, int psi_kind, int nsteps, double kcut, double k0) {
    // the hot loop of start_raytracing! (raytracing/RaytracingDriver.jl:256-270; with the k-cutoff of
    // raytracing/TwoLayerRaytracing.jl:136-141 when kcut > 0), nsteps times, without returning to the host language
    if (!p || nsteps < 0) return fail(SWRT_ERR_ARG, "bad argument");
    swrt_flow* f = p->flow;
    // everything a step can reject is checked before any step (a failure inside a graph capture would leave the clock advanced)
    if (p->d.interp != f->interp) return fail(SWRT_ERR_STATE, "packets use interpolant %d but the flow's snapshots hold node data for %d (swrt_flow_set_interp)", p->d.interp, f->interp);
    { int rc = check_psi_kind(f, psi_kind); if (rc) return rc; }
    if (f->P > 1 && !f->p2p) return fail(SWRT_ERR_STATE, "slab-decomposed flow without mapped peers: drive the loop from the host (slab.py)");
    // Team mode, fused loop.  The snapshot of state n and stage A of the flow step n -> n+1 transform the SAME state, so their
    // y-jobs share one y-pass, one transpose and one barrier (two barriers per coupled step instead of four):
    //   step 1 (plain) | step k >= 2: [psi; stage A + psi jobs] B [x-pass snapshot(k-1); x-pass products] B [halo; trace(k-2 -> k-1)] stage C
    //   after the last step: the plain band snapshot of the final state and the last trace.
    // Same kernels on the same data as the plain loop below (only the job layout of the exchange buffer differs).
    static const int fused_mode = [] { const char* e = getenv("SWRT_TEAM_FUSED"); return e ? atoi(e) : 1; }();
    const bool fusable = (f->d.stepper == SWRT_IFMAB3 || f->d.stepper == SWRT_FILTEREDAB3) && f->d.model != SWRT_THOMASYAMADA;
    if (f->P > 1 && fused_mode && fusable && nsteps > 0) {
        const int nj = model_njobs_a(f->d.model) + 3, j0 = model_njobs_a(f->d.model);
        auto trace = [&]() -> int {
            int rc = swrt_packets_raytrace(p, 0.0, f->d.dt);
            if (rc) return rc;
            if (kcut > 0.0 && (rc = swrt_packets_kcutoff_reset(p, kcut, k0, nullptr))) return rc;
            return swrt_flow_swap_snapshots(f, 0);
        };
        int rc;
        for (int s = 0; s < nsteps; ++s) {
            if (s == 0) {
                if ((rc = swrt_slab_step(f, 1))) return rc;
                continue;
            }
            if ((rc = slab_stage_a_fused(f, psi_kind))) return rc;
            if ((rc = team_barrier(f))) return rc;
            if ((rc = slab_snap_b(f, 1, nj, j0))) return rc;               // snapshot of the state the previous step produced
            if ((rc = slab_stage_b(f, nj))) return rc;
            if ((rc = team_barrier(f))) return rc;                         // product rows are with their owners, band rows are written
            if ((rc = band_halo_pull(f, 1))) return rc;
            if ((rc = trace())) return rc;                                 // (on the packets' own stream when they have one)
            if ((rc = swrt_slab_stage_c(f))) return rc;
        }
        if ((rc = swrt_slab_band_snapshot(f, psi_kind, 1))) return rc;
        if ((rc = trace())) return rc;
        if (p->own) CK(cudaStreamWaitEvent(f->st, p->ev_done, 0));
        return SWRT_OK;
    }
    auto body = [&]() -> int {
        int rc = f->P > 1 ? swrt_slab_step(f, 1) : swrt_flow_step(f, 1);
        if (rc) return rc;
        if ((rc = f->P > 1 ? swrt_slab_band_snapshot(f, psi_kind, 1) : swrt_flow_velocity_snapshot(f, psi_kind, 1))) return rc;
        // the tracer only sees t - t0 and t1 - t0: every step is traced over (0, dt), which makes the kernel arguments of
        // consecutive steps identical (the absolute (old_t, new_t) of the per-call loop differ from this by rounding only)
        if ((rc = swrt_packets_raytrace(p, 0.0, f->d.dt))) return rc;
        if (kcut > 0.0 && (rc = swrt_packets_kcutoff_reset(p, kcut, k0, nullptr))) return rc;
        return swrt_flow_swap_snapshots(f, 0);
    };
    // Launch-bound grid sizes: replay six coupled steps from one CUDA graph.  After six steps the history ring (period 3) and
    // the snapshot slots (period 2) are back where they started, so every kernel argument repeats -- except the packet
    // buffers, which a sort swaps (one graph per buffer parity; steps that contain a sort run un-captured).
    static const int graph_mode = [] { const char* e = getenv("SWRT_GRAPH"); return e ? atoi(e) : 1; }();
    const int period = 6;
    const bool ring_stepper = f->d.stepper == SWRT_IFMAB3 || f->d.stepper == SWRT_FILTEREDAB3;
    // (other handles reading the snapshots on their own streams need the event waits a replay would skip: no graph then)
    const bool eligible = graph_mode > 0 && !f->prof && !p->own && f->readers.empty() && f->P == 1 && f->refine == 1 && !f->forcing &&
                          (graph_mode > 1 || (long long)f->d.nx * f->d.ny <= 1024LL * 1024LL);
    for (int s = 0; s < nsteps;) {
        const bool aligned = f->step >= 3 && f->slot_map[0] == 0 && f->slot_map[1] == 1 && (!ring_stepper || f->ring == 0);   // (aliased slots {0,0} run un-captured)
        const bool sort_inside = p->d.sort_every > 0 && p->since_sort + period > p->d.sort_every;
        if (eligible && aligned && !sort_inside && nsteps - s >= period && p->nbins == (long long)f->d.nx * f->d.ny) {
            swrt_packets::Cycle& c = p->cycle[p->xk < p->xk2 ? 0 : 1];
            if (c.exec && (c.xk != p->xk || c.snap0 != f->snap[0] || c.psi_kind != psi_kind || c.interp != f->interp || c.kcut != kcut || c.k0 != k0)) {
                cudaGraphExecDestroy(c.exec);
                c.exec = nullptr;
            }
            if (!c.exec) {
                CK(cudaSetDevice(f->d.device));
                const long long l0 = f->launches;
                f->no_graph = true;
                cudaError_t be = cudaStreamBeginCapture(f->st, cudaStreamCaptureModeRelaxed);
                if (be != cudaSuccess) { f->no_graph = false; CK(be); }
                int rc = SWRT_OK;
                for (int q = 0; q < period && rc == SWRT_OK; ++q) rc = body();
                cudaGraph_t g = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(f->st, &g);
                f->no_graph = false;
                if (rc != SWRT_OK) { if (g) cudaGraphDestroy(g); return rc; }
                CK(ce);
                const cudaError_t ie = cudaGraphInstantiate(&c.exec, g, 0);
                cudaGraphDestroy(g);
                CK(ie);
                c.launches = f->launches - l0;
                c.xk = p->xk; c.snap0 = f->snap[0]; c.psi_kind = psi_kind; c.interp = f->interp; c.kcut = kcut; c.k0 = k0;
            } else {
                for (int q = 0; q < period; ++q) { f->t += f->d.dt; f->step += 1; }
                p->since_sort += period;
                f->launches += c.launches;
            }
            CK(cudaGraphLaunch(c.exec, f->st));
            s += period;
        } else {
            const int rc = body();
            if (rc) return rc;
            s += 1;
        }
    }
    // packets on their own stream: whatever the caller times or reads on the flow's stream after this call includes their last trace
    if (p->own && nsteps > 0) CK(cudaStreamWaitEvent(f->st, p->ev_done, 0));
    return SWRT_OK;
}

// ------------------------------------------------------------------ k-omega accumulator
int swrt_series_create(swrt_flow* flow, int kind, int kr_index, long long max_frames, swrt_series** out) {
    if (!flow || !out) return fail(SWRT_ERR_ARG, "null pointer");
    *out = nullptr;
    if (kind != SWRT_SERIES_TY && kind != SWRT_SERIES_RSW) return fail(SWRT_ERR_ARG, "unknown series kind %d", kind);
    const bool rsw = flow->d.model == SWRT_RSW || flow->d.model == SWRT_RSW_MODIFIED || flow->d.model == SWRT_RSW_LINDBORG;
    if (kind == SWRT_SERIES_TY ? flow->d.model != SWRT_THOMASYAMADA : !rsw) return fail(SWRT_ERR_ARG, "series kind %d does not apply to model %d", kind, flow->d.model);
    if (kr_index < 0 || kr_index >= flow->nkr || max_frames < 1) return fail(SWRT_ERR_ARG, "bad kr index / max_frames");
    if (flow->P > 1) return fail(SWRT_ERR_UNSUPPORTED, "not available for a slab-decomposed flow");
    CK(cudaSetDevice(flow->d.device));
    swrt_series* s = new swrt_series;
    s->flow = flow; s->kind = kind; s->kr = kr_index; s->maxf = max_frames;
    s->nser = kind == SWRT_SERIES_TY ? 6 : 12;
    cudaError_t e = cudaMalloc(&s->buf, sizeof(double2) * (size_t)s->nser * flow->d.ny * (size_t)max_frames);
    if (e == cudaSuccess) e = cudaMalloc(&s->wts, sizeof(double2) * 3 * (size_t)flow->L.vs);
    if (e != cudaSuccess) { swrt_series_destroy(s); return fail(SWRT_ERR_CUDA, "cudaMalloc(series): %s", cudaGetErrorString(e)); }
    *out = s;
    return SWRT_OK;
}
int swrt_series_destroy(swrt_series* s) {
    if (!s) return SWRT_OK;
    cudaSetDevice(s->flow->d.device);
    cudaStreamSynchronize(s->flow->st);
    cudaFree(s->buf); cudaFree(s->wts);
    delete s;
    return SWRT_OK;
}
int swrt_series_append(swrt_series* s) {
    if (!s) return fail(SWRT_ERR_ARG, "null pointer");
    swrt_flow* h = s->flow;
    if ((long long)s->t.size() >= s->maxf) return fail(SWRT_ERR_STATE, "series is full (%lld frames)", s->maxf);
    CK(cudaSetDevice(h->d.device));
    const SpecLayout& L = h->L;
    int rc;
    if (s->kind == SWRT_SERIES_RSW) {   // weights first: decompose_kernel writes them to G, which the decomposition then reuses
        if ((rc = decompose_launch(h, DEC_RSW_WTS))) return rc;
        CK(cudaMemcpyAsync(s->wts, h->G, sizeof(double2) * 3 * (size_t)L.vs, cudaMemcpyDeviceToDevice, h->st));
    }
    if ((rc = decompose_launch(h, s->kind == SWRT_SERIES_TY ? DEC_TY : DEC_RSW))) return rc;
    { ProfScope ps(h, K_OTHER);
      series_append_kernel<<<(L.ny + 127) / 128, 128, 0, h->st>>>(h->sol, h->G, h->H, s->wts, L, s->kind, s->kr - L.kr_off, (long long)s->t.size(), s->maxf, s->buf); }
    CK(cudaGetLastError());
    s->t.push_back(h->t);
    return SWRT_OK;
}
int swrt_series_frames(swrt_series* s, long long* n) {
    if (!s || !n) return fail(SWRT_ERR_ARG, "null pointer");
    *n = (long long)s->t.size();
    return SWRT_OK;
}
int swrt_series_times(swrt_series* s, double* t_host) {
    if (!s || !t_host) return fail(SWRT_ERR_ARG, "null pointer");
    std::copy(s->t.begin(), s->t.end(), t_host);
    return SWRT_OK;
}
// out(T, nl) <- rows of `which` (series) or its windowed transform (spectrum)
static int series_fetch(swrt_series* s, int which, bool spectrum, void* host) {
    if (!s || !host) return fail(SWRT_ERR_ARG, "null pointer");
    swrt_flow* h = s->flow;
    const long long T = (long long)s->t.size();
    const int ny = h->d.ny, nspec = s->kind == SWRT_SERIES_TY ? 9 : 12;
    if (T < 1) return fail(SWRT_ERR_STATE, "no frames recorded");
    if (which < 0 || which >= (spectrum ? nspec : s->nser)) return fail(SWRT_ERR_ARG, "series / spectrum %d out of range", which);
    CK(cudaSetDevice(h->d.device));
    double2 *work = nullptr, *out = nullptr, *tw = nullptr;
    double* tv = nullptr;
    cudaError_t e = cudaMalloc(&work, sizeof(double2) * (size_t)ny * T);
    if (e == cudaSuccess) e = cudaMalloc(&out, sizeof(double2) * (size_t)ny * T);
    if (e == cudaSuccess) e = cudaMalloc(&tw, sizeof(double2) * (size_t)T);
    if (e == cudaSuccess) e = cudaMalloc(&tv, sizeof(double) * (size_t)T);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tv, s->t.data(), sizeof(double) * (size_t)T, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) {
        SeriesSel sel{1, {which, 0, 0}, {-1, -1, -1}};
        if (s->kind == SWRT_SERIES_TY && which >= 6) {   // TY_k_omega.jl:104-106
            if (which == 6) sel = SeriesSel{2, {0, 2, 0}, {1, 3, -1}};        // (ut + ug) + i (vt + vg)
            else if (which == 7) sel = SeriesSel{1, {4, 0, 0}, {5, -1, -1}};  // uw + i vw
            else sel = SeriesSel{3, {4, 2, 0}, {5, 3, 1}};                    // (uw + ug + ut) + i (vw + vg + vt)
        }
        if (spectrum) {
            { ProfScope ps(h, K_OTHER); twiddle_table_kernel<<<(unsigned)((T + 255) / 256), 256, 0, h->st>>>(tw, T); }
            { ProfScope ps(h, K_OTHER); series_dft_kernel<<<ny, 256, 0, h->st>>>(s->buf, sel, ny, T, s->maxf, tv, tw, s->kind == SWRT_SERIES_RSW ? 1 : 0, work, out); }
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(host, out, sizeof(double2) * (size_t)ny * T, cudaMemcpyDeviceToHost, h->st);
        } else {
            e = cudaMemcpy2DAsync(host, sizeof(double2) * (size_t)T, s->buf + (size_t)which * ny * s->maxf, sizeof(double2) * (size_t)s->maxf,
                                  sizeof(double2) * (size_t)T, (size_t)ny, cudaMemcpyDeviceToHost, h->st);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cudaFree(work); cudaFree(out); cudaFree(tw); cudaFree(tv);
    CK(e);
    return SWRT_OK;
}
int swrt_series_get(swrt_series* s, int which, void* series_host) { return series_fetch(s, which, false, series_host); }
int swrt_series_spectrum(swrt_series* s, int which, void* spectrum_host) { return series_fetch(s, which, true, spectrum_host); }

// ------------------------------------------------------------------ output roll-over arithmetic
int swrt_seqout_init(swrt_seqout* s, long long max_writes) {
    if (!s || max_writes <= 0) return fail(SWRT_ERR_ARG, "bad argument");
    s->max_writes = max_writes;
    s->current_writes = 0;
    s->file_index = 0;
    return SWRT_OK;
}
int swrt_seqout_write(swrt_seqout* s, long long nwrites, long long* file_index) {
    // utils/SequencedOutputs.jl:37-44,58-63: the key lands in the current file, THEN the counter is checked
    if (!s) return fail(SWRT_ERR_ARG, "null pointer");
    if (file_index) *file_index = s->file_index;
    s->current_writes += nwrites;
    if (s->current_writes >= s->max_writes) {
        s->current_writes = 0;
        s->file_index += 1;
    }
    return SWRT_OK;
}
int swrt_seqout_filename(const char* base, long long idx, char* buf, int buflen) {
    if (!base || !buf || buflen <= 0) return fail(SWRT_ERR_ARG, "bad argument");
    snprintf(buf, (size_t)buflen, "%s.%06lld.jld2", base, idx);   // raytracing/RaytracingDriver.jl:173-174
    return SWRT_OK;
}
int swrt_collated_filename(const char* base, long long idx, char* buf, int buflen) {
    if (!base || !buf || buflen <= 0) return fail(SWRT_ERR_ARG, "bad argument");
    snprintf(buf, (size_t)buflen, "%s_%08lld.out", base, idx);    // utils/Collated.jl:58-60
    return SWRT_OK;
}

#pragma GCC visibility pop
}  // extern "C"
