// Team mode (one process per GPU, peers' memory mapped with CUDA IPC): the device-side pieces that replace a communication
// library on the data path.
//
//   * team_barrier_kernel   stream-ordered barrier over NVLink: every rank stores its epoch into each peer's flag word and
//                           spins until all peers' epochs have arrived in its own (what a one-element all-reduce was used for).
//   * packets sharded by y-band (SURVEY 8e + DESIGN 5): a rank traces the packets whose grid row lies in its band of ny / P
//     rows, so the slab-decomposed flow step hands it exactly the snapshot rows it needs (plus a few halo rows from the two
//     neighbours) and no all-gather of the velocity field exists.  Packets that drift into another band are handed over at
//     every re-sort: the sorted order already groups them by owner (the sort key is tile-row major), so a rank publishes P
//     (start, count) pairs, and after a barrier every rank PULLS its segments from the peers' sorted arrays.
//   * scatter / gather between the caller's row order (contiguous index blocks per rank, raytracing/RaytracingDriver.jl:27-47
//     row order is what every host-visible array keeps) and the band owners are "pull scans": a rank reads the peers' staging
//     blocks (or resident arrays) over NVLink and keeps the rows that are its own.
#pragma once
#include <cuda_runtime.h>

#include "packets.cuh"
#include "passes.cuh"

namespace swrt {

struct TeamFlags {                      // lives in every rank's shared arena; [channel][source rank]
    unsigned long long arrive[2][kMaxPeers];
};
struct TeamPeers {
    TeamFlags* f[kMaxPeers];
};

__global__ void team_barrier_kernel(TeamPeers peers, int P, int self, unsigned long long epoch, int channel) {
    const int d = threadIdx.x;
    if (d >= P) return;
    __threadfence_system();                                            // everything this stream did before is visible to the peers
    volatile unsigned long long* theirs = &peers.f[d]->arrive[channel][self];
    *theirs = epoch;
    __threadfence_system();
    volatile unsigned long long* mine = &peers.f[self]->arrive[channel][d];
    while (*mine < epoch) __nanosleep(64);
    __threadfence_system();
}

// ---------------------------------------------------------------- packet hand-over between bands
struct PacketArena {       // pointers into one rank's packet arena (capacity `cap` per column)
    double* xk[2];         // [4][cap] double-buffered
    double* sign[2];
    unsigned* idx[2];      // GLOBAL original row of each packet
    double* out;           // [6][cap]: sampler output in resident order / staging block of a scatter
    unsigned long long* tab;   // [0..P): segment starts, [P..2P): segment counts (written by the peers), [2P]: resident count,
                               // [2P+1]: staging rows, [2P+2]: first global row of the staging block, [2P+3]: overflow flag
};
struct ArenaPeers {
    PacketArena a[kMaxPeers];
};

// owner of grid row j: bands of `yrows` rows
__device__ __forceinline__ int band_owner(int j, int yshift_rows) { return j >> yshift_rows; }

// after the local cell sort: tell every destination where its segment of my sorted array starts and how long it is
__global__ void team_publish_segments_kernel(const unsigned* __restrict__ hist_end, long long keys_per_rank, ArenaPeers peers, int P, int self) {
    const int d = threadIdx.x;
    if (d >= P) return;
    const unsigned long long start = d == 0 ? 0ULL : (unsigned long long)hist_end[(long long)d * keys_per_rank - 1];
    const unsigned long long end = (unsigned long long)hist_end[(long long)(d + 1) * keys_per_rank - 1];
    peers.a[d].tab[self] = start;
    peers.a[d].tab[P + self] = end - start;
}

// every rank pulls its segments from the peers' sorted arrays (buffer `src_buf` on every rank) into its buffer `dst_buf`
__global__ void __launch_bounds__(256) team_pull_segments_kernel(ArenaPeers peers, int P, int self, int src_buf, int dst_buf, long long cap) {
    const PacketArena me = peers.a[self];
    __shared__ unsigned long long off[kMaxPeers + 1];
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int s = 0; s < P; ++s) { off[s] = acc; acc += me.tab[P + s]; }
        off[P] = acc;
        if (blockIdx.x == 0) {
            me.tab[2 * P] = acc > (unsigned long long)cap ? (unsigned long long)cap : acc;
            if (acc > (unsigned long long)cap) me.tab[2 * P + 3] = 1ULL;      // overflow: the caller reports it
        }
    }
    __syncthreads();
    const long long total = (long long)(off[P] > (unsigned long long)cap ? (unsigned long long)cap : off[P]);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < P && (unsigned long long)i >= off[s + 1]) ++s;
        const long long r = (long long)me.tab[s] + (i - (long long)off[s]);
        const PacketArena src = peers.a[s];
#pragma unroll
        for (int c = 0; c < 4; ++c) me.xk[dst_buf][c * cap + i] = __ldcs(src.xk[src_buf] + c * cap + r);
        me.sign[dst_buf][i] = __ldcs(src.sign[src_buf] + r);
        me.idx[dst_buf][i] = __ldcs(src.idx[src_buf] + r);
    }
}

// scatter: every rank scans all staging blocks (5 columns x, y, k, l, sign in `out`, rows in caller order) and appends the
// packets of its own band to its resident arrays (buffer `dst_buf`); tab[2P] must have been zeroed
__global__ void __launch_bounds__(256) team_scatter_scan_kernel(ArenaPeers peers, int P, int self, int dst_buf, long long cap, PacketGrid g,
                                                                int band_shift) {
    const PacketArena me = peers.a[self];
    const unsigned lane = threadIdx.x & 31;
    for (int s = 0; s < P; ++s) {
        const PacketArena src = peers.a[s];
        const long long rows = (long long)src.tab[2 * P + 1], first = (long long)src.tab[2 * P + 2];
        // uniform trip count per warp: the append position comes from ONE atomic per warp (ballot + popc), not one per packet
        for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < rows; i0 += (long long)gridDim.x * blockDim.x) {
            const long long i = i0 + threadIdx.x;
            double y = 0.0;
            bool mine = false;
            if (i < rows) {
                y = __ldcs(src.out + cap + i);
                int j0, j1;
                double b;
                cell(y, g.y0, g.inv_dy, g.ny, j0, j1, b);
                mine = band_owner(j0, band_shift) == self;
            }
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (m == 0u) continue;
            unsigned long long base = 0ULL;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(&me.tab[2 * P], (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!mine) continue;
            const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            if (pos >= (unsigned long long)cap) { me.tab[2 * P + 3] = 1ULL; continue; }
            me.xk[dst_buf][pos] = __ldcs(src.out + i);
            me.xk[dst_buf][cap + pos] = y;
            me.xk[dst_buf][2 * cap + pos] = __ldcs(src.out + 2 * cap + i);
            me.xk[dst_buf][3 * cap + pos] = __ldcs(src.out + 3 * cap + i);
            me.sign[dst_buf][pos] = __ldcs(src.out + 4 * cap + i);
            me.idx[dst_buf][pos] = (unsigned)(first + i);
        }
    }
}

// gather: every rank scans all resident packets of all ranks and copies the rows that belong to its caller-order block
// [first, first + n) into dst[c][idx - first].  which = 0: the packet state (4 columns of xk[src_buf]), 1: `out` (ncols columns,
// resident order: the sampler's output).
__global__ void __launch_bounds__(256) team_gather_scan_kernel(ArenaPeers peers, int P, int src_buf, int which, int ncols, long long cap,
                                                               long long first, long long n, double* __restrict__ dst, long long ldd) {
    for (int s = 0; s < P; ++s) {
        const PacketArena src = peers.a[s];
        const long long cnt = (long long)src.tab[2 * P];
        const double* cols = which == 0 ? src.xk[src_buf] : src.out;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) {
            const long long o = (long long)__ldcs(src.idx[src_buf] + i) - first;
            if (o < 0 || o >= n) continue;
            for (int c = 0; c < ncols; ++c) dst[c * ldd + o] = __ldcs(cols + c * cap + i);
        }
    }
}

// halo rows of the band snapshot: copy `halo` rows from each neighbour's band buffer (their first / last owned rows) into
// this rank's halo rows.  Buffer layout per level: [halo + yrows + halo][nx][SNAP_STRIDE].
__global__ void __launch_bounds__(256) team_halo_pull_kernel(double* __restrict__ mine, const double* __restrict__ below, const double* __restrict__ above,
                                                             int halo, int yrows, long long row_doubles) {
    const long long n2 = (long long)halo * row_doubles / 2;       // double2 elements per halo block
    double2* lo = reinterpret_cast<double2*>(mine);                                            // rows [0, halo)            <- below's last owned rows
    double2* hi = reinterpret_cast<double2*>(mine + (long long)(halo + yrows) * row_doubles);  // rows [halo+yrows, +halo)  <- above's first owned rows
    const double2* sb = reinterpret_cast<const double2*>(below + (long long)yrows * row_doubles);   // below: owned rows start at `halo`; last `halo` of them
    const double2* sa = reinterpret_cast<const double2*>(above + (long long)halo * row_doubles);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * n2; i += (long long)gridDim.x * blockDim.x) {
        if (i < n2) lo[i] = __ldcs(sb + i);
        else hi[i - n2] = __ldcs(sa + (i - n2));
    }
}

}  // namespace swrt
