// Layout of the device-resident background-velocity snapshots read by the ray tracer: one array per time level,
// S[y][x][6] doubles = (u, v, ux, uy, vx, pad), 48 B = three 16-byte vectors per grid point (vy = -ux).
// Separate arrays per level: the snapshot kernel then writes whole 32-byte sectors (an interleaved two-level record made
// every store a partial-sector read-modify-write: ncu showed 663 MB of DRAM traffic for 168 MB of payload).
#pragma once
namespace swrt {
constexpr int SNAP_NC = 5;      // u, v, ux, uy, vx   (vy = -ux)
constexpr int SNAP_STRIDE = 6;  // doubles per grid point of one level
// Hermite-bicubic mode (utils/CUDAInterpolations.jl:71-108): node data (u, v, ux, uy, vx, uxy, vxy, pad), 64 B per point
constexpr int SNAP3_NC = 7;
constexpr int SNAP3_STRIDE = 8;
// fp32 packet mode: node data (u, v, ux, uy | vx, 0, 0, 0) as eight floats, 32 B = two 16-byte vectors = one sector per point
constexpr int SNAPF_STRIDE = 8;   // floats per grid point of one level
}  // namespace swrt
