// Block-level fp64 complex FFT in shared memory for sm_100a.
//
// One "group" of G = N/16 threads transforms one length-N sequence held in shared
// memory as split re/im arrays (8-byte elements: a half-warp access is conflict free when
// the 16 lanes hit 16 different 8-byte banks).  Every thread owns E = 16 complex values in
// registers per stage: Stockham autosort stages of radix 16 (4x4 register butterflies) and
// one remainder stage of radix 2/4/8, natural order in, natural order out, in place
// (load all -> barrier -> butterfly + store -> barrier).  Indices are padded by one slot per
// 16 (PAD) so that the stride-R stores of the first stage spread over all banks.
//
// Twiddles: one table lookup per butterfly (w = exp(-2 pi i m / N), read-only path) and
// powers by squaring/multiplying in registers -- LSU bandwidth, not FP64 issue, is the
// scarcer resource in these kernels.
#pragma once
#include <cuda_runtime.h>

namespace swrt {

__host__ __device__ constexpr int pad_index(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4) + 1; }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by SIGN * i   (SIGN = -1: forward transform, +1: inverse)
template <int SIGN>
__device__ __forceinline__ double2 mul_si(double2 a) {
    return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// cos/sin of 2 pi m / 16
__device__ constexpr double kC16[16] = {
    1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
    0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848,
    -1.0, -0.92387953251128673848, -0.70710678118654752440, -0.38268343236508977173,
    0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848};
__device__ constexpr double kS16[16] = {
    0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848,
    1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
    0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848,
    -1.0, -0.92387953251128673848, -0.70710678118654752440, -0.38268343236508977173};

// v *= exp(SIGN * 2 pi i m / 16), m compile-time after unrolling
template <int SIGN>
__device__ __forceinline__ double2 mul_w16(double2 a, int m) {
    m &= 15;
    if (m == 0) return a;
    if (m == 4) return mul_si<SIGN>(a);
    if (m == 8) return make_double2(-a.x, -a.y);
    if (m == 12) return mul_si<-SIGN>(a);
    const double c = kC16[m], s = SIGN * kS16[m];
    return make_double2(fma(a.x, c, -a.y * s), fma(a.x, s, a.y * c));
}

template <int SIGN>
__device__ __forceinline__ void dft2(double2& a, double2& b) {
    double2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <int SIGN>
__device__ __forceinline__ void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
    double2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = mul_si<SIGN>(csub(x1, x3));
    x0 = cadd(a, c);
    x1 = cadd(b, d);
    x2 = csub(a, c);
    x3 = csub(b, d);
}

// In-register DFT of length R (2,4,8,16), natural order in and out.
// R = R1*R2 with n = R2*a + b, k = c + R1*d:  w_R^{nk} = w_R1^{ac} w_R^{bc} w_R2^{bd}.
template <int R, int SIGN>
struct Dft;
template <int SIGN>
struct Dft<1, SIGN> {
    static __device__ __forceinline__ void run(double2 (&)[1]) {}
};
template <int SIGN>
struct Dft<2, SIGN> {
    static __device__ __forceinline__ void run(double2 (&v)[2]) { dft2<SIGN>(v[0], v[1]); }
};
template <int SIGN>
struct Dft<4, SIGN> {
    static __device__ __forceinline__ void run(double2 (&v)[4]) { dft4<SIGN>(v[0], v[1], v[2], v[3]); }
};
template <int SIGN>
struct Dft<8, SIGN> {
    static __device__ __forceinline__ void run(double2 (&v)[8]) {
        // R1 = 4 (over a), R2 = 2 (over b): n = 2a + b, k = c + 4d
#pragma unroll
        for (int b = 0; b < 2; ++b) dft4<SIGN>(v[b], v[2 + b], v[4 + b], v[6 + b]);  // Y_b[c] at v[2c+b]
#pragma unroll
        for (int c = 1; c < 4; ++c) v[2 * c + 1] = mul_w16<SIGN>(v[2 * c + 1], 2 * c);  // w8^{c}
        double2 o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double2 p = v[2 * c], q = v[2 * c + 1];
            dft2<SIGN>(p, q);
            o[c] = p;
            o[c + 4] = q;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[i];
    }
};
template <int SIGN>
struct Dft<16, SIGN> {
    static __device__ __forceinline__ void run(double2 (&v)[16]) {
        // n = 4a + b, k = c + 4d
#pragma unroll
        for (int b = 0; b < 4; ++b) dft4<SIGN>(v[b], v[4 + b], v[8 + b], v[12 + b]);  // Y_b[c] at v[4c+b]
#pragma unroll
        for (int c = 1; c < 4; ++c)
#pragma unroll
            for (int b = 1; b < 4; ++b) v[4 * c + b] = mul_w16<SIGN>(v[4 * c + b], b * c);
        double2 o[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double2 p0 = v[4 * c], p1 = v[4 * c + 1], p2 = v[4 * c + 2], p3 = v[4 * c + 3];
            dft4<SIGN>(p0, p1, p2, p3);
            o[c] = p0;
            o[c + 4] = p1;
            o[c + 8] = p2;
            o[c + 12] = p3;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = o[i];
    }
};

// One Stockham stage: radix R, P = product of the radices of the previous stages.
// `g` = thread index inside the group (0 .. N/16-1).  Must be called by every thread of the
// CTA (uses __syncthreads); threads with active == false only take part in the barriers.
template <int N, int R, int P, int SIGN>
__device__ __forceinline__ void fft_stage(double* __restrict__ re, double* __restrict__ im, int g,
                                          const double2* __restrict__ tw, bool active) {
    constexpr int E = (N >= 16) ? 16 : N;
    constexpr int G = N / E;
    constexpr int B = E / R;  // butterflies per thread
    constexpr int T = N / R;  // butterflies per transform
    double2 v[B][R];
    if (active) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int j = g + b * G;
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int idx = pad_index(j + t * T);
                v[b][t] = make_double2(re[idx], im[idx]);
            }
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int j = g + b * G;
            const int k = j & (P - 1);
            if (P > 1) {
                double2 w1 = __ldg(&tw[k * (N / (P * R))]);
                if (SIGN > 0) w1.y = -w1.y;
                double2 w[R];
                w[1 % R] = w1;
#pragma unroll
                for (int t = 2; t < R; ++t) w[t] = (t & 1) ? cmul(w[t - 1], w1) : cmul(w[t / 2], w[t / 2]);
#pragma unroll
                for (int t = 1; t < R; ++t) v[b][t] = cmul(v[b][t], w[t]);
            }
            Dft<R, SIGN>::run(v[b]);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int idx = pad_index(j0 + t * P);
                re[idx] = v[b][t].x;
                im[idx] = v[b][t].y;
            }
        }
    }
    __syncthreads();
}

template <int N, int P, int SIGN>
struct FftStages {
    static __device__ __forceinline__ void run(double* re, double* im, int g, const double2* tw, bool active) {
        constexpr int rem = N / P;
        constexpr int R = rem >= 16 ? 16 : rem;
        fft_stage<N, R, P, SIGN>(re, im, g, tw, active);
        if constexpr (P * R < N) FftStages<N, P * R, SIGN>::run(re, im, g, tw, active);
    }
};

// Unnormalised in-place transform  X[k] = sum_n x[n] exp(SIGN 2 pi i n k / N).
// Starts with a barrier so that preceding shared-memory writes by other threads are visible.
template <int N, int SIGN>
__device__ __forceinline__ void block_fft(double* re, double* im, int g, const double2* tw, bool active = true) {
    __syncthreads();
    FftStages<N, 1, SIGN>::run(re, im, g, tw, active);
}

// ---------------------------------------------------------------------------------------------
// Register-interface variants (N >= 32).  "Strided register layout": thread g owns the 16 elements
// n = g + m N/16, m = 0..15.  That is exactly what the first Stockham stage consumes (radix 16, P = 1)
// and -- because the last stage of radix R writes j + t N/R with j = g + b N/16 -- also what the last
// stage produces (m = b + t 16/R).  So data can enter the transform from registers (straight from
// global memory) and leave it in registers (straight to global memory, or into the next transform's
// first stage after a pointwise product) without the two extra shared-memory passes.
// ---------------------------------------------------------------------------------------------
template <int N, int SIGN>
__device__ __forceinline__ void fft_first_stage_regs(double2 (&v)[16], double* __restrict__ re, double* __restrict__ im, int g) {
    static_assert(N >= 32, "register interface needs N >= 32");
    Dft<16, SIGN>::run(v);
    __syncthreads();  // earlier readers of this buffer are done
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int idx = pad_index(g * 16 + t);
        re[idx] = v[t].x;
        im[idx] = v[t].y;
    }
    __syncthreads();
}

template <int N, int R, int P, int SIGN>
__device__ __forceinline__ void fft_last_stage_regs(const double* __restrict__ re, const double* __restrict__ im, int g,
                                                    const double2* __restrict__ tw, double2 (&out)[16]) {
    constexpr int G = N / 16, B = 16 / R, T = N / R;
    static_assert(P * R == N, "last stage");
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = g + b * G;
        double2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const int idx = pad_index(j + t * T);
            v[t] = make_double2(re[idx], im[idx]);
        }
        double2 w1 = __ldg(&tw[j]);  // k = j (j < T = P), table stride N / (P R) = 1
        if (SIGN > 0) w1.y = -w1.y;
        double2 w[R];
        w[1 % R] = w1;
#pragma unroll
        for (int t = 2; t < R; ++t) w[t] = (t & 1) ? cmul(w[t - 1], w1) : cmul(w[t / 2], w[t / 2]);
#pragma unroll
        for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
        Dft<R, SIGN>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) out[b + t * B] = v[t];
    }
}

__host__ __device__ constexpr int last_stage_P(int N) {
    int P = 16;
    while (N / P > 16) P *= 16;
    return P;
}
template <int N, int P, int SIGN>
__device__ __forceinline__ void fft_middle_stages(double* re, double* im, int g, const double2* tw) {
    if constexpr (N / P > 16) {  // radix-16 stages strictly between the first and the last
        fft_stage<N, 16, P, SIGN>(re, im, g, tw, true);
        fft_middle_stages<N, P * 16, SIGN>(re, im, g, tw);
    }
}

// registers in -> registers out (both in the strided register layout); `re/im` is the work buffer
template <int N, int SIGN>
__device__ __forceinline__ void block_fft_regs(double2 (&v)[16], double* re, double* im, int g, const double2* tw) {
    fft_first_stage_regs<N, SIGN>(v, re, im, g);
    fft_middle_stages<N, 16, SIGN>(re, im, g, tw);
    constexpr int LP = last_stage_P(N);
    fft_last_stage_regs<N, N / LP, LP, SIGN>(re, im, g, tw, v);
}

// shared memory in (natural order, written by other threads) -> registers out (strided register layout)
template <int N, int SIGN>
__device__ __forceinline__ void block_fft_regs_out(double* re, double* im, int g, const double2* tw, double2 (&v)[16]) {
    __syncthreads();
    fft_stage<N, 16, 1, SIGN>(re, im, g, tw, true);
    fft_middle_stages<N, 16, SIGN>(re, im, g, tw);
    constexpr int LP = last_stage_P(N);
    fft_last_stage_regs<N, N / LP, LP, SIGN>(re, im, g, tw, v);
}

// registers in -> shared memory out (natural order), ends with a barrier
template <int N, int SIGN>
__device__ __forceinline__ void block_fft_regs_in(double2 (&v)[16], double* re, double* im, int g, const double2* tw) {
    fft_first_stage_regs<N, SIGN>(v, re, im, g);
    FftStages<N, 16, SIGN>::run(re, im, g, tw, true);
}

}  // namespace swrt
