// walls.cuh -- bounce-back streaming of the reference as a PULL rule (shared by K1 and the layout kernels).
//
// streaming::StreamingBounceBack / ThermalStreamingBounceBack (/root/reference/src/streaming.cpp:66-112, 150-196)
// push every population to its neighbour, reflecting at the walls, in one serial loop over (x outer, y, i):
// several sources can target one slot (the last one wins) and some slots are targeted by nobody and keep what the
// destination array held before.  Per DESTINATION slot (X, Y, j) the candidates are
//   (1) plain streaming from (X - cx_j, Y - cy_j), direction j, when that cell exists;
//   (2) a "y out only" reflection of direction o = opp(j) at cell (X + cx_j, Y): needs Y - cy_j outside;
//   (3) an "x out only" reflection of direction o at cell (X, Y + cy_j): needs X - cx_j outside;
//   (4) a corner reflection of direction o at (X, Y) itself: needs both outside.
// The destination array is the reference's shared temp_* array (src/plasma.cpp:478-504): when f is streamed it
// holds the PRE-collision f of the same step (left there by the swap in Collisions), when g is streamed it holds
// the POST-collision f of the same species (left by the swap after streaming f).  A slot nobody targets therefore
// reads: f -> its own pre-collision value of the previous step, g -> the post-collision f at the same (cell, j).
#pragma once
#include "lbm_consts.h"

namespace plbm {

__device__ constexpr int WALL_CX[NQ] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 };
__device__ constexpr int WALL_CY[NQ] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
__device__ constexpr int WALL_OPP[NQ] = { 0, 3, 4, 1, 2, 7, 8, 5, 6 };

struct WallSource { int x, y, dir; };       // dir < 0: nobody writes the slot

__device__ __forceinline__ WallSource wall_source(int X, int Y, int j, int NX, int NY)
{
    const int cx = WALL_CX[j], cy = WALL_CY[j], o = WALL_OPP[j];
    const bool xin = (X - cx >= 0 && X - cx < NX), yin = (Y - cy >= 0 && Y - cy < NY);
    WallSource b = { -1, -1, -1 };           // winning source: the last in (x, y, i) order
    auto offer = [&](int sx, int sy, int si) {
        if (sx > b.x || (sx == b.x && (sy > b.y || (sy == b.y && si > b.dir)))) { b.x = sx; b.y = sy; b.dir = si; }
    };
    if (xin && yin) offer(X - cx, Y - cy, j);                                   // (1)
    if (!yin && X + cx >= 0 && X + cx < NX) offer(X + cx, Y, o);                // (2)
    if (!xin && Y + cy >= 0 && Y + cy < NY) offer(X, Y + cy, o);                // (3)
    if (!xin && !yin) offer(X, Y, o);                                           // (4)
    return b;
}

// wall cells numbered along the four edges (corners belong to the horizontal edges)
__host__ __device__ __forceinline__ int wall_rim_index(int x, int y, int NX, int NY)
{
    if (y == 0) return x;
    if (y == NY - 1) return NX + x;
    if (x == 0) return 2 * NX + y;
    return 2 * NX + NY + y;
}
__host__ __device__ __forceinline__ int wall_rim_count(int NX, int NY) { return 2 * NX + 2 * NY; }

// rim: [species][direction][wall cell] pre-collision f of the last step (read); rim_next: the same for this step (written)
struct WallArgs {
    const double* rim;
    double* rim_next;
    int nrim;
    int identity;      // 1: `src` holds the state at the top of the loop itself (after Initialize / an upload): no streaming
};

} // namespace plbm
