/* C restatement of oracle/raytrace.py (`cell_index`, `sample_bilinear`, `rhs`, `raytrace`) -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Same arithmetic as the NumPy oracle, operation for operation: dxkdt of raytracing/GPURaytracing.jl:32-65 (group velocity
 * Cg^2 k / omega with the frequency sign, refraction -(grad U)^T k, vy = -ux), texture-style bilinear sampling with wrap
 * addressing (:18-30), linear interpolation in time between the two snapshots, classical RK4 with `nsub` sub-steps.
 * It exists so that the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) is compiled, threaded code -- like the
 * Julia reference -- rather than vectorised NumPy.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * Arrays are NumPy C-order: xk (n, 4), F_old / F_new (nx, ny, 5) = u, v, ux, uy, vx.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline void cell_index(double x, double x0, double dx, long long n, long long* i, double* a) {
    const double s = (x - x0) / dx, fl = floor(s);
    long long m = (long long)fmod(fl, (double)n);
    if (m < 0) m += n;
    *i = m;
    *a = s - fl;
}

static inline void sample_bilinear(const double* F, long long nx, long long ny, double x0, double y0, double dx, double dy, double x,
                                   double y, double* out /*[5]*/) {
    long long i, j;
    double a, b;
    cell_index(x, x0, dx, nx, &i, &a);
    cell_index(y, y0, dy, ny, &j, &b);
    const long long i1 = (i + 1) % nx, j1 = (j + 1) % ny;
    const double *f00 = F + (i * ny + j) * 5, *f10 = F + (i1 * ny + j) * 5, *f01 = F + (i * ny + j1) * 5, *f11 = F + (i1 * ny + j1) * 5;
    for (int c = 0; c < 5; ++c) {
        const double bottom = (1 - a) * f00[c] + a * f10[c], top = (1 - a) * f01[c] + a * f11[c];
        out[c] = (1 - b) * bottom + b * top;
    }
}

static inline void rhs(const double* s, double sign, double alpha, const double* Fo, const double* Fn, long long nx, long long ny, double x0,
                       double y0, double dx, double dy, double f, double Cg, int lerp, double* d) {
    const double k = s[2], l = s[3];
    const double w = sign * sqrt(f * f + Cg * Cg * (k * k + l * l));
    double So[5], Sn[5], W[5];
    sample_bilinear(Fo, nx, ny, x0, y0, dx, dy, s[0], s[1], So);
    sample_bilinear(Fn, nx, ny, x0, y0, dx, dy, s[0], s[1], Sn);
    for (int c = 0; c < 5; ++c) W[c] = lerp == 0 ? (1 - alpha) * So[c] + alpha * Sn[c] : alpha * So[c] + (1 - alpha) * Sn[c];
    d[0] = W[0] + Cg * Cg * k / w;
    d[1] = W[1] + Cg * Cg * l / w;
    d[2] = -(W[2] * k + W[4] * l);
    d[3] = -(W[3] * k - W[2] * l);
}

void oracle_raytrace_rk4(double* xk, const double* sign, long long n, double t0, double t1, const double* Fo, const double* Fn, long long nx,
                         long long ny, double x0, double y0, double dx, double dy, double f, double Cg, int nsub, int lerp, int threads) {
    const double h = (t1 - t0) / nsub;
    /* `threads` > 0 is passed explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which must not decide the baseline */
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < n; ++p) {
        double* s = xk + 4 * p;
        for (int it = 0; it < nsub; ++it) {
            const double t = t0 + it * h;
            double k1[4], k2[4], k3[4], k4[4], y[4];
            rhs(s, sign[p], (t - t0) / (t1 - t0), Fo, Fn, nx, ny, x0, y0, dx, dy, f, Cg, lerp, k1);
            for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k1[c];
            rhs(y, sign[p], (t + 0.5 * h - t0) / (t1 - t0), Fo, Fn, nx, ny, x0, y0, dx, dy, f, Cg, lerp, k2);
            for (int c = 0; c < 4; ++c) y[c] = s[c] + 0.5 * h * k2[c];
            rhs(y, sign[p], (t + 0.5 * h - t0) / (t1 - t0), Fo, Fn, nx, ny, x0, y0, dx, dy, f, Cg, lerp, k3);
            for (int c = 0; c < 4; ++c) y[c] = s[c] + h * k3[c];
            rhs(y, sign[p], (t + h - t0) / (t1 - t0), Fo, Fn, nx, ny, x0, y0, dx, dy, f, Cg, lerp, k4);
            for (int c = 0; c < 4; ++c) s[c] += (h / 6) * (k1[c] + 2 * k2[c] + 2 * k3[c] + k4[c]);
        }
    }
}

/* number of threads the next parallel region will use (what bench.py records in cpu_baseline.cores) */
int oracle_raytrace_threads(int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    int n = 1;
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
    return n;
#else
    (void)threads;
    return 1;
#endif
}
