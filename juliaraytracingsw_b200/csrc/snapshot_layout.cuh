// Layout of the device-resident background-velocity snapshots read by the ray tracer:
// S[y][x][2][5] doubles, the two time levels (halves) of (u, v, ux, uy, vx) interleaved per grid point.
#pragma once
namespace swrt {
constexpr int SNAP_NC = 5;       // u, v, ux, uy, vx   (vy = -ux)
constexpr int SNAP_STRIDE = 10;  // doubles per grid point (two time levels)
}  // namespace swrt
