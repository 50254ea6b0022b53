// Generic y-pass / x-pass kernels of the pseudo-spectral step (sm_100a, fp64).
//
// Spectral arrays are stored dealiased and padded:  a[var][l][kr_pad] (complex128), only
// kr < kr_keep and l outside [lz0, lz1) are ever non-zero ("sol is dealiased after every
// step": rsw/RotatingShallowWater.jl:141 applies dealias!(sol) at the top of every calcN!).
// A 2-D transform is two passes with the physical-space row living only on chip:
//
//   ypass_inv   columns (fixed kr, all l)  -> G[job][y][kr]     complex FFT along l
//   xpass       rows    (fixed y,  all kr) -> c2r, pointwise products, r2c -> H[job][y][kr]
//   ypass_fwd   columns of H               -> spectral N[var][l][kr]
//
// Model-specific arithmetic (which spectral expression feeds which transform, which
// products are formed, how transforms combine into N) is supplied by small functors.
#pragma once
#include <type_traits>
#include "fft.cuh"

namespace swrt {

struct SpecLayout {
    int nx, ny;        // physical grid
    int kr_keep;       // retained kr columns [0, kr_keep)
    int kr_pad;        // leading dimension (multiple of 16)
    int lz0, lz1;      // zeroed l index band [lz0, lz1)
    long long vs;      // var stride = ny * kr_pad (complex elements)
    double dk, dl;     // 2 pi / Lx, 2 pi / Ly
    double f, Cg2;     // model constants used by loaders
    double aux0, aux1; // model specific (Kd2, ...)
    double aux2, aux3, aux4, aux5;   // MultiLayerQG-2: U1, U2, beta, mu
    // Slab decomposition over P ranks (all equal to the single-GPU values when P = 1): a rank owns `kr_keep` columns starting at
    // global column kr_off (kr_pad = columns per rank, the same on every rank) and `yrows` = ny / P physical rows.  The
    // y-transformed / x-transformed intermediates are stored [y block][job][row in block][kr_pad]: exactly the send / receive
    // layout of the all-to-all transposes, and the plain [job][y][kr_pad] array when P = 1.
    int kr_off, kr_keep_g, yshift, yrows;
    // addforcing! (rsw/RotatingShallowWater.jl:234-240): a stored spectral field [l][kr_pad] that the forward y-pass adds to EVERY
    // variable of N (`@. N += vars.Fh` broadcasts the 2-D Fh over the three components); nullptr = no forcing
    const double2* forcing;
};

// element offset of (job, y, column 0) in an intermediate array holding `njobs` jobs
__device__ __forceinline__ long long inter_off(const SpecLayout& L, int njobs, int job, int y) {
    const int blk = y >> L.yshift, yl = y & ((1 << L.yshift) - 1);
    return ((((long long)blk * njobs + job) << L.yshift) + yl) * L.kr_pad;
}
// Where a transform pass writes its output.  Single GPU: one local array.  Slab mode: the block of rows (y-pass) or of
// columns (x-pass) destined for rank d is stored straight into rank d's receive buffer over NVLink (peer memory mapped
// with cudaIpcOpenMemHandle) at the position reserved for this rank -- the all-to-all transpose IS the store of the FFT
// pass; no send buffer, no separate exchange kernel.  Without peer mapping p[d] points into the local send buffer.
constexpr int kMaxPeers = 16;
struct OutPeers {
    double2* p[kMaxPeers];
    int self;    // index of this rank's block inside a peer's receive buffer
};
__device__ __forceinline__ double2* out_at(const OutPeers& o, const SpecLayout& L, int njobs, int job, int y) {
    const int blk = y >> L.yshift, yl = y & ((1 << L.yshift) - 1);
    return o.p[blk] + ((((long long)o.self * njobs + job) << L.yshift) + yl) * L.kr_pad;
}

// One x-pass row (fixed job and local row).  Single GPU: kr_pad contiguous columns (RowPlain).  Slab mode: P segments of
// kr_pad columns (one per source/destination rank), `skip + chunk` elements apart (RowSeg).  The row type is a template
// parameter of the x-pass ops so that the single-GPU kernels carry no segment arithmetic at all.
struct RowPlain {
    double2* base;
    __device__ __forceinline__ double2* at(int k) const { return base + k; }
};
// (segment of column k = k / chunk as one multiply-high with magic = ceil(2^32 / chunk), exact for k < 8192 and every chunk that is a
// multiple of 16 -- checked exhaustively; a compare-and-add loop over the segments made the slab x-pass kernels twice as long as the
// single-GPU ones in instructions: 12 312 against 6 712 for the RSW op at N = 2048, profiles/r02_sass_summary.txt)
__host__ __device__ inline unsigned seg_magic(int chunk) { return (unsigned)(((1ULL << 32) + (unsigned)chunk - 1) / (unsigned)chunk); }
struct RowSeg {
    double2* base;
    long long skip;   // segment stride - chunk: added once per segment boundary crossed
    unsigned magic;
    int chunk, nseg;
    __device__ __forceinline__ double2* at(int k) const {
        const int sgm = (int)__umulhi((unsigned)k, magic);
        return base + k + (long long)sgm * skip;
    }
};
struct RowSegOut {   // x-pass output row in slab mode: column segment s goes to peer s
    const OutPeers* peers;
    long long off;    // offset of (this rank's block, job, local row) inside a peer's buffer
    unsigned magic;
    int chunk, nseg;
    __device__ __forceinline__ double2* at(int k) const {
        const int sgm = (int)__umulhi((unsigned)k, magic);
        return peers->p[sgm] + off + (k - sgm * chunk);
    }
};
template <bool SLAB>
struct RowOf { using type = RowPlain; };
template <>
struct RowOf<true> { using type = RowSeg; };
template <bool SLAB>
__device__ __forceinline__ typename RowOf<SLAB>::type row_ref(const SpecLayout& L, const double2* arr, int njobs, int job, int yl) {
    typename RowOf<SLAB>::type r;
    r.base = const_cast<double2*>(arr) + (((long long)job << L.yshift) + yl) * L.kr_pad;
    if constexpr (SLAB) {
        r.skip = ((long long)njobs << L.yshift) * L.kr_pad - L.kr_pad;
        r.chunk = L.kr_pad;
        r.magic = seg_magic(L.kr_pad);
        r.nseg = L.ny >> L.yshift;
    }
    return r;
}
// Input row of the x-pass in slab mode, addressed through a table of source buffers: entry s is where the column segment of
// source rank s lives -- this rank's receive buffer (the y-pass pushed it there) or, in pull mode, rank s's own send buffer
// mapped over NVLink (the y-pass stored locally in full lines and this pass fetches whole 1-3 KB column segments).
template <bool SLAB>
__device__ __forceinline__ typename std::conditional<SLAB, RowSegOut, RowPlain>::type row_in(const SpecLayout& L, const double2* arr,
                                                                                           const OutPeers& src, int njobs, int job, int yl) {
    typename std::conditional<SLAB, RowSegOut, RowPlain>::type r;
    if constexpr (SLAB) {
        r.peers = &src;
        r.off = ((((long long)src.self * njobs + job) << L.yshift) + yl) * L.kr_pad;
        r.chunk = L.kr_pad;
        r.magic = seg_magic(L.kr_pad);
        r.nseg = L.ny >> L.yshift;
    } else {
        r.base = const_cast<double2*>(arr) + (((long long)job << L.yshift) + yl) * L.kr_pad;
    }
    return r;
}
template <bool SLAB>
struct RowOutOf { using type = RowPlain; };
template <>
struct RowOutOf<true> { using type = RowSegOut; };
// output row of the x-pass: `arr` is the local array on a single GPU, `peers` the destination table in slab mode
template <bool SLAB>
__device__ __forceinline__ typename RowOutOf<SLAB>::type row_out(const SpecLayout& L, double2* arr, const OutPeers& peers, int njobs, int job,
                                                                 int yl) {
    typename RowOutOf<SLAB>::type r;
    if constexpr (SLAB) {
        r.peers = &peers;
        r.off = ((((long long)peers.self * njobs + job) << L.yshift) + yl) * L.kr_pad;
        r.chunk = L.kr_pad;
        r.magic = seg_magic(L.kr_pad);
        r.nseg = L.ny >> L.yshift;
    } else {
        r.base = arr + (((long long)job << L.yshift) + yl) * L.kr_pad;
    }
    return r;
}

__device__ __forceinline__ double wave_l(const SpecLayout& L, int j) {
    return (double)(j < L.ny / 2 ? j : j - L.ny) * L.dl;
}
__device__ __forceinline__ bool l_retained(const SpecLayout& L, int j) { return j < L.lz0 || j >= L.lz1; }

__host__ __device__ constexpr int group_size(int N) { return N >= 16 ? N / 16 : 1; }
// column stride (in doubles) of the y-pass tile: >= padded_len(N) and == 16/TK (mod 16) so that the TK adjacent
// columns touched by one half-warp fall into disjoint 8-byte banks
__host__ __device__ constexpr int col_stride(int N, int TK) {
    int cs = padded_len(N);
    const int want = (16 / TK) % 16;
    while (cs % 16 != want) ++cs;
    return cs;
}
constexpr int kSmemPerSM = 227 * 1024;
__host__ __device__ constexpr int clamp_blocks(long long smem_bytes, int threads) {
    int b = (int)(kSmemPerSM / (smem_bytes + 1024));
    const int by_threads = 1536 / threads;          // keep >= ~42 registers per thread available
    if (b > by_threads) b = by_threads;
    if (b > 8) b = 8;
    return b < 1 ? 1 : b;
}
__host__ __device__ constexpr long long ypass_smem(int N, int TK) { return 2LL * TK * col_stride(N, TK) * 8; }
__host__ __device__ constexpr long long xpass_smem(int N, int nbuf) { return 2LL * nbuf * padded_len(N) * 8; }

// ------------------------------------------------------------------------------------
// y-pass, inverse direction: out[job][y][kr] = sum_l src_job(kr, l) exp(+2 pi i l y / ny)
// Loader: __device__ double2 operator()(int job, int kr, int l, double kw, double lw, long long off)
//
// Thread (g, c) = (tid / TK, tid % TK) owns rows g + m N/16 of column kr0 + c: the values go from global
// memory straight into the first FFT stage and from the last stage straight back to global memory; the TK
// adjacent lanes of a row access TK adjacent complex numbers (16 TK contiguous bytes).
// ------------------------------------------------------------------------------------
template <int N, int TK, class Loader>
__global__ void __launch_bounds__(TK* group_size(N), clamp_blocks(ypass_smem(N, TK), TK* group_size(N)))
    ypass_inv_kernel(Loader ld, SpecLayout L, int njobs, OutPeers out, const double2* __restrict__ tw) {
    extern __shared__ double smem[];
    constexpr int G = group_size(N), NP = col_stride(N, TK);
    const int tid = threadIdx.x;
    const int c = tid % TK, g = tid / TK;
    double* re = smem + c * NP;
    double* im = smem + (TK + c) * NP;
    const int ntiles = (L.kr_keep + TK - 1) / TK;
    for (int w = blockIdx.x; w < ntiles * njobs; w += gridDim.x) {
        const int job = w / ntiles, kr = (w % ntiles) * TK + c;
        const double kw = (L.kr_off + kr) * L.dk;
        const bool col_ok = kr < L.kr_keep;
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int l = g + m * G;
            v[m] = make_double2(0.0, 0.0);
            if (col_ok && l_retained(L, l)) v[m] = ld(job, kr, l, kw, wave_l(L, l), (long long)l * L.kr_pad + kr);
        }
        block_fft_regs<N, +1>(v, re, im, g, tw);
        if (col_ok) {
#pragma unroll
            for (int m = 0; m < 16; ++m) out_at(out, L, njobs, job, g + m * G)[kr] = v[m];
        }
    }
}

// ------------------------------------------------------------------------------------
// y-pass, inverse direction, with an asynchronous input prefetch (cp.async -> shared-memory staging).
// For "simple" jobs whose source is one stored spectral field times a multiplier that depends on l only
// (1, i l, -i l, l^2): while the CTA transforms tile w, the retained rows of tile w+1 stream into the staging
// buffer without passing through registers, so the global-load latency (the dominant stall of the plain kernel:
// ncu long_scoreboard) is hidden behind the FFT.  One CTA per SM: 4 columns x (work 139.5 KB + staging 87.3 KB) at N = 2048.
// ------------------------------------------------------------------------------------
enum { YMUL_ONE = 0, YMUL_IL = 1, YMUL_NEG_IL = 2, YMUL_L2 = 3 };
struct SimpleJobs {
    const double2* src[8];  // source spectral field of each job ([l][kr_pad])
    int mul[8];
    int last[8];            // this job is the last reader of its source in this pass: load it evict-first
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
// Same copy for data that is dead once read (the x-transformed products): evict-first in L2, so the lines it frees go to
// the arrays the step still needs instead of to these.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_stream(void* smem_dst, const void* gsrc, unsigned long long policy) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "l"(policy));
}
__host__ __device__ constexpr long long ypass_stage_smem(int N, int TK) { return (long long)N * TK * 16; }  // upper bound (all rows)

template <int N, int TK>
__global__ void __launch_bounds__(TK* group_size(N), (TK * group_size(N) <= 256 ? 2 : 1))
    ypass_inv_prefetch_kernel(SimpleJobs jobs, SpecLayout L, int njobs, OutPeers out, const double2* __restrict__ tw) {
    extern __shared__ double smem[];
    constexpr int G = group_size(N), NP = col_stride(N, TK), NT = TK * G;
    const int tid = threadIdx.x;
    const int c = tid % TK, g = tid / TK;
    double* re = smem + c * NP;
    double* im = smem + (TK + c) * NP;
    double2* stg = reinterpret_cast<double2*>(smem + 2 * TK * NP);   // [retained row][TK]
    const int ntiles = (L.kr_keep + TK - 1) / TK;
    const int nz = L.lz1 - L.lz0, rows = L.ny - nz;
    const int nwork = ntiles * njobs;
    const unsigned long long pol = l2_evict_first_policy();
    auto prefetch = [&](int w) {
        const int job = w / ntiles, kr0 = (w % ntiles) * TK;
        const double2* src = jobs.src[job];
        const bool last = jobs.last[job] != 0;
        for (int ch = tid; ch < rows * TK; ch += NT) {
            const int r = ch / TK, cc = ch - r * TK;
            const int l = r < L.lz0 ? r : r + nz;
            // columns beyond kr_keep inside the padded row are zero in every stored field: safe to copy
            if (last) cp_async16_stream(&stg[ch], &src[(long long)l * L.kr_pad + kr0 + cc], pol);
            else cp_async16(&stg[ch], &src[(long long)l * L.kr_pad + kr0 + cc]);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    int w = blockIdx.x;
    if (w < nwork) prefetch(w);
    for (; w < nwork; w += gridDim.x) {
        const int job = w / ntiles, kr = (w % ntiles) * TK + c;
        const int mul = jobs.mul[job];
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int l = g + m * G;
            v[m] = make_double2(0.0, 0.0);
            if (l_retained(L, l)) {
                const int r = l < L.lz0 ? l : l - nz;
                const double2 a = stg[r * TK + c];
                const double lw = wave_l(L, l);
                v[m] = mul == YMUL_ONE ? a : mul == YMUL_IL ? make_double2(-lw * a.y, lw * a.x)
                     : mul == YMUL_NEG_IL ? make_double2(lw * a.y, -lw * a.x) : make_double2(lw * lw * a.x, lw * lw * a.y);
            }
        }
        __syncthreads();                       // staging consumed: the next tile may stream in
        if (w + (int)gridDim.x < nwork) prefetch(w + gridDim.x);
        block_fft_regs<N, +1>(v, re, im, g, tw);
        if (kr < L.kr_keep) {
#pragma unroll
            for (int m = 0; m < 16; ++m) out_at(out, L, njobs, job, g + m * G)[kr] = v[m];
        }
    }
}

// ------------------------------------------------------------------------------------
// y-pass, forward direction, with combination into output variables.
// Combiner:  int nin(int var); int src(int var, int i);
//            double2 apply(int var, int i, double2 v, double kw, double lw)
//            double2 init(int var, double kw, double lw, long long off)      (terms of N that are linear in the state)
//            int var_of(int slot)      (processing order: variables with more inputs first)
// out[var][l][kr] = init + sum_i apply(var, i, FFT_y(H[src(var,i)])[kr, l])
// ------------------------------------------------------------------------------------
template <int N, int TK, class Combiner>
__global__ void __launch_bounds__(TK* group_size(N), clamp_blocks(ypass_smem(N, TK), TK* group_size(N)))
    ypass_fwd_kernel(Combiner cb, SpecLayout L, int nvars, int nh, const double2* __restrict__ H, double2* __restrict__ out,
                     const double2* __restrict__ tw) {
    extern __shared__ double smem[];
    constexpr int G = group_size(N), NP = col_stride(N, TK);
    const int tid = threadIdx.x;
    const int c = tid % TK, g = tid / TK;
    double* re = smem + c * NP;
    double* im = smem + (TK + c) * NP;
    const int ntiles = (L.kr_keep + TK - 1) / TK;
    for (int w = blockIdx.x; w < ntiles * nvars; w += gridDim.x) {
        const int var = cb.var_of(w / ntiles), kr = (w % ntiles) * TK + c;   // heaviest variables first (static balance)
        const double kw = (L.kr_off + kr) * L.dk;
        const bool col_ok = kr < L.kr_keep;
        double2* o = out + (long long)var * L.vs + kr;
        const int nin = cb.nin(var);
        for (int i_in = 0; i_in < nin; ++i_in) {
            const int srcj = cb.src(var, i_in);
            double2 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m] = make_double2(0.0, 0.0);
                if (col_ok) v[m] = H[inter_off(L, nh, srcj, g + m * G) + kr];
            }
            block_fft_regs<N, -1>(v, re, im, g, tw);
            if (col_ok) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int l = g + m * G;
                    if (!l_retained(L, l)) continue;
                    double2 r = cb.apply(var, i_in, v[m], kw, wave_l(L, l));
                    double2 prev = i_in > 0 ? o[(long long)l * L.kr_pad]
                                            : cb.init(var, kw, wave_l(L, l), (long long)l * L.kr_pad + kr);
                    if (i_in == 0 && L.forcing) {
                        const double2 fh = L.forcing[(long long)l * L.kr_pad + kr];
                        prev.x += fh.x;
                        prev.y += fh.y;
                    }
                    r.x += prev.x;
                    r.y += prev.y;
                    o[(long long)l * L.kr_pad] = r;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// y-pass, forward direction, with the same asynchronous prefetch.  All N rows of a product column are non-zero and the
// staging buffer only holds `rows_s` of them (1436 of 2048 at N = 2048, TK = 4): rows >= rows_s are loaded directly, issued
// before the wait on the staged part so that both latencies overlap.
// ------------------------------------------------------------------------------------
template <int N, int TK, class Combiner>
__global__ void __launch_bounds__(TK* group_size(N), (TK * group_size(N) <= 256 ? 2 : 1))
    ypass_fwd_prefetch_kernel(Combiner cb, SpecLayout L, int nvars, int nh, int rows_s, const double2* __restrict__ H,
                              double2* __restrict__ out, const double2* __restrict__ tw) {
    extern __shared__ double smem[];
    constexpr int G = group_size(N), NP = col_stride(N, TK), NT = TK * G;
    const int tid = threadIdx.x;
    const int c = tid % TK, g = tid / TK;
    double* re = smem + c * NP;
    double* im = smem + (TK + c) * NP;
    double2* stg = reinterpret_cast<double2*>(smem + 2 * TK * NP);   // [row < rows_s][TK]
    const int ntiles = (L.kr_keep + TK - 1) / TK;
    const int nwork = ntiles * nvars;
    const unsigned long long pol = l2_evict_first_policy();
    auto prefetch = [&](int w, int i_in) {
        const int var = cb.var_of(w / ntiles), kr0 = (w % ntiles) * TK;
        const int srcj = cb.src(var, i_in);
        for (int ch = tid; ch < rows_s * TK; ch += NT) {
            const int r = ch / TK, cc = ch - r * TK;
            cp_async16_stream(&stg[ch], &H[inter_off(L, nh, srcj, r) + kr0 + cc], pol);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    int w = blockIdx.x, i_in = 0;
    if (w < nwork) prefetch(w, 0);
    while (w < nwork) {
        const int var = cb.var_of(w / ntiles), kr = (w % ntiles) * TK + c;
        const double kw = (L.kr_off + kr) * L.dk;
        const bool col_ok = kr < L.kr_keep;
        const int nin = cb.nin(var);
        const int srcj = cb.src(var, i_in);
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {           // rows outside the staging buffer: direct, issued first
            const int y = g + m * G;
            v[m] = make_double2(0.0, 0.0);
            if (y >= rows_s && col_ok) v[m] = __ldcs(&H[inter_off(L, nh, srcj, y) + kr]);
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int y = g + m * G;
            if (y < rows_s) v[m] = stg[y * TK + c];
        }
        __syncthreads();                         // staging consumed
        int wn = w, in = i_in + 1;               // next FFT task of this CTA
        if (in >= nin) { wn = w + gridDim.x; in = 0; }
        if (wn < nwork) prefetch(wn, in);
        block_fft_regs<N, -1>(v, re, im, g, tw);
        if (col_ok) {
            double2* o = out + (long long)var * L.vs + kr;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int l = g + m * G;
                if (!l_retained(L, l)) continue;
                double2 r = cb.apply(var, i_in, v[m], kw, wave_l(L, l));
                double2 prev = i_in > 0 ? o[(long long)l * L.kr_pad]
                                        : cb.init(var, kw, wave_l(L, l), (long long)l * L.kr_pad + kr);
                if (i_in == 0 && L.forcing) {
                    const double2 fh = L.forcing[(long long)l * L.kr_pad + kr];
                    prev.x += fh.x;
                    prev.y += fh.y;
                }
                r.x += prev.x;
                r.y += prev.y;
                o[(long long)l * L.kr_pad] = r;
            }
        }
        w = wn;
        i_in = in;
    }
}

// ------------------------------------------------------------------------------------
// x-pass helpers: one CTA of G = N/16 threads owns NB complex work buffers of length N.
// Two real fields travel through one complex transform: z = a + i b.
// ------------------------------------------------------------------------------------
enum { MUL_ONE = 0, MUL_IK = 1, MUL_MK2 = 2, MUL_ZERO = 3, MUL_K2 = 4 };

template <int M>
__device__ __forceinline__ double2 apply_mul(double2 v, double kw) {
    if (M == MUL_IK) return make_double2(-kw * v.y, kw * v.x);
    if (M == MUL_MK2) return make_double2(-kw * kw * v.x, -kw * kw * v.y);
    if (M == MUL_K2) return make_double2(kw * kw * v.x, kw * kw * v.y);
    return v;
}

template <int N>
struct XCtx {
    static constexpr int G = group_size(N), NP = padded_len(N), EPT = N / G;
    double* smem;
    const double2* tw;
    int g, kr_keep;   // kr_keep: GLOBAL number of retained columns (rows are whole in the x-pass)
    double dk;
    __device__ __forceinline__ double* re(int b) const { return smem + (2 * b) * NP; }
    __device__ __forceinline__ double* im(int b) const { return smem + (2 * b + 1) * NP; }

    // Build the Hermitian-extended spectrum of z = a + i b from the half spectra A, B (rows of
    // kr_pad complex).  Only the real parts of A[0], B[0] enter, like a c2r transform.
    template <int MA, int MB, class RA, class RB>
    __device__ __forceinline__ void load_pair(int b, const RA& A, const RB& B) const {
        double *r = re(b), *m = im(b);
        // The global loads are issued in batches before the shared-memory stores of the batch (the compiler cannot move a load across
        // a store through unrelated generic pointers, so a load-store loop paid one L2 round trip per iteration: ncu long_scoreboard
        // 17 % of the x-pass).
        constexpr int IT = (N / 2 + G - 1) / G, BATCH = IT < 4 ? IT : 4;   // (all eight iterations at once spill: 64 registers of loads)
        __syncthreads();  // earlier pointwise readers of this buffer are done
#pragma unroll 1
        for (int i0 = 0; i0 < IT; i0 += BATCH) {
            double2 za[BATCH], zb[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int k = g + (i0 + i) * G;
                za[i] = make_double2(0.0, 0.0);
                zb[i] = make_double2(0.0, 0.0);
                if (k < N / 2 && k < kr_keep) {
                    za[i] = __ldcs(A.at(k));
                    if (MB != MUL_ZERO) zb[i] = __ldcs(B.at(k));
                }
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int k = g + (i0 + i) * G;
                if (k >= N / 2) continue;
                const double kw = k * dk;
                const double2 a = apply_mul<MA>(za[i], kw), bb = MB != MUL_ZERO ? apply_mul<MB>(zb[i], kw) : make_double2(0.0, 0.0);
                if (k == 0) {
                    r[0] = a.x;
                    m[0] = bb.x;
                    r[pad_index(N / 2)] = 0.0;
                    m[pad_index(N / 2)] = 0.0;
                } else {
                    r[pad_index(k)] = a.x - bb.y;
                    m[pad_index(k)] = a.y + bb.x;
                    r[pad_index(N - k)] = a.x + bb.y;
                    m[pad_index(N - k)] = bb.x - a.y;
                }
            }
        }
    }
    __device__ __forceinline__ void ifft(int b) const { block_fft<N, +1>(re(b), im(b), g, tw); }
    __device__ __forceinline__ void fft(int b) const { block_fft<N, -1>(re(b), im(b), g, tw); }

    // Register-interface versions (fft.cuh): v[m] <-> x (or k) = g + m N/16.
    // Hermitian-extended spectrum of z = a + i b straight from the half spectra in global memory.
    template <int MA, int MB, class RA, class RB>
    __device__ __forceinline__ void load_pair_regs(double2 (&v)[16], const RA& A, const RB& B) const {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int x = g + m * G;
            const int k = m < 8 ? x : N - x;            // m < 8  <=>  x < N/2
            double2 za = make_double2(0.0, 0.0), zb = make_double2(0.0, 0.0);
            if (k < kr_keep) {                           // also excludes k = N/2 (kr_keep <= N/2)
                const double kw = k * dk;
                za = apply_mul<MA>(__ldcs(A.at(k)), kw);
                if (MB != MUL_ZERO) zb = apply_mul<MB>(__ldcs(B.at(k)), kw);
            }
            if (m < 8) v[m] = (m == 0 && x == 0) ? make_double2(za.x, zb.x) : make_double2(za.x - zb.y, za.y + zb.x);
            else v[m] = make_double2(za.x + zb.y, zb.x - za.y);
        }
    }
    // inverse transform of buffer b (filled by load_pair) -> registers out
    __device__ __forceinline__ void ifft_regs_out(int b, double2 (&v)[16]) const { block_fft_regs_out<N, +1>(re(b), im(b), g, tw, v); }
    // inverse transform, registers in -> registers out, work buffer b
    __device__ __forceinline__ void ifft_regs(double2 (&v)[16], int b) const { block_fft_regs<N, +1>(v, re(b), im(b), g, tw); }
    // forward transform, registers in -> shared memory buffer b (natural order; then store_pair)
    __device__ __forceinline__ void fft_regs_in(double2 (&v)[16], int b) const { block_fft_regs_in<N, -1>(v, re(b), im(b), g, tw); }

    // After a forward transform of z = p + i q:  2 P[k] = Z[k] + conj Z[N-k],  2i Q[k] = Z[k] - conj Z[N-k].
    // Writes 2P and 2Q (callers fold the 1/2 into their scaling) for k < kr_keep.
    template <int MP, int MQ, class RA, class RB>
    __device__ __forceinline__ void store_pair(int b, const RA& P, const RB& Q) const {
        const double *r = re(b), *m = im(b);
        for (int k = g; k < kr_keep; k += G) {
            const int kn = (N - k) & (N - 1);
            const double a = r[pad_index(k)], bb = m[pad_index(k)], c = r[pad_index(kn)], d = m[pad_index(kn)];
            const double kw = k * dk;
            *P.at(k) = apply_mul<MP>(make_double2(a + c, bb - d), kw);
            if (MQ != MUL_ZERO) *Q.at(k) = apply_mul<MQ>(make_double2(bb + d, c - a), kw);
        }
        __syncthreads();
    }
};

// Rows are handed out dynamically (one atomic per row): 2048 rows over 444 resident CTAs would otherwise leave an
// 8 % tail (5 rounds for 4.6 rounds of work).  `sched` = {next row, finished CTAs}; the last CTA re-arms it for the next launch.
template <int N, class Op>
__global__ void __launch_bounds__(group_size(N), clamp_blocks(xpass_smem(N, Op::NBUF), group_size(N)))
    xpass_kernel(Op op, SpecLayout L, const double2* __restrict__ tw, unsigned* __restrict__ sched) {
    extern __shared__ double smem[];
    __shared__ int next_row;
    XCtx<N> cx;
    cx.smem = smem;
    cx.tw = tw;
    cx.g = threadIdx.x;
    cx.kr_keep = L.kr_keep_g;
    cx.dk = L.dk;
    int y = blockIdx.x;                      // first row: static
    while (y < L.yrows) {
        if (threadIdx.x == 0) next_row = (int)(gridDim.x + atomicAdd(&sched[0], 1u));
        op.row(cx, L, y);                    // (contains barriers: next_row is visible afterwards)
        __syncthreads();
        y = next_row;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
        }
    }
}

}  // namespace swrt
