// Row-wise force evaluation over a world held in shared memory (one thread = one body), shared by
// the resident kernel (K3) and the ensemble kernel (K4).  Per-row view of World::set_forces
// (src/World.cpp:42-57) + Object::update_forces_against (src/Object.cpp:64-95).
#pragma once

#include "common.cuh"

namespace essa {

template<bool EXACT, bool ARGMAX>
__device__ __forceinline__ void row_forces(int k, int n, int npad, double const* sx, double const* sy, double const* sz, double const* sg,
    double x, double y, double z, double gfk, double eps2, double& ax, double& ay, double& az, double& maxattr, int& most) {
    if (EXACT) {
        double fx = 0, fy = 0, fz = 0, best = 0;
        int bp = -1;
        for (int p = 0; p < n; p++) {
            double gp = sg[p];
            if (p == k || gp == 0.0)
                continue;
            exact_interact(x, y, z, gfk, sx[p], sy[p], sz[p], gp, p, fx, fy, fz, best, bp);
        }
        ax = fx;
        ay = fy;
        az = fz;
        maxattr = best;
        if (bp >= 0)
            most = bp;
    } else {
        double fx[4] = { 0, 0, 0, 0 }, fy[4] = { 0, 0, 0, 0 }, fz[4] = { 0, 0, 0, 0 }, bm[4] = { 0, 0, 0, 0 };
        int bj[4] = { -1, -1, -1, -1 };
        for (int p = 0; p < npad; p += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double4 q = make_double4(sx[p + u], sy[p + u], sz[p + u], sg[p + u]);
                if (ARGMAX)
                    fast_interact_argmax(x, y, z, gfk, q, p + u, eps2, fx[u], fy[u], fz[u], bm[u], bj[u]);
                else
                    fast_interact(x, y, z, q, eps2, fx[u], fy[u], fz[u]);
            }
        }
        ax = (fx[0] + fx[1]) + (fx[2] + fx[3]);
        ay = (fy[0] + fy[1]) + (fy[2] + fy[3]);
        az = (fz[0] + fz[1]) + (fz[2] + fz[3]);
        if (ARGMAX) {
            double best = 0;
            int bp = -1;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                // lowest index wins ties, as a sequential strict '>' scan would have it
                if (bm[u] > best || (bm[u] == best && bj[u] >= 0 && bp >= 0 && bj[u] < bp)) {
                    best = bm[u];
                    bp = bj[u];
                }
            }
            maxattr = best;
            if (bp >= 0)
                most = bp;
        }
    }
}

} // namespace essa
