// Shared by the kernels for worlds of up to 32 bodies: K3w (resident_warp.cu, exact_order on one warp) and K3q
// (resident_quad.cu, fast arithmetic on a two-CTA cluster, optionally persistent across World::update calls).
#pragma once

#include "common.cuh"
#include "ctx.hpp"

namespace essa {

constexpr int kRing = 16; // snapshots in flight

struct WarpArgs {
    WorldPtrs w;
    RecordPtrs r;
    long long const* cdate;
    long long const* ddate;
    uint8_t const* delflag;
    uint8_t* active_out;
    int n, steps, L, iters;
    double hm, dtm; // (dt / 2) * mul, dt * mul
    long long date, date_step;
    double eps2;
    int record, argmax, offset_trails, forward_sim, two_pass, acc_valid;
};

// Persistent mode (K3q only): the kernel outlives the World::update call that launched it.  The host posts the next
// tick in mapped pinned memory (PersistMailbox, persist.hpp), the kernel publishes the packed state there and acknowledges.
struct PersistArgs {
    unsigned long long const* cmd; // (seq << 32) | steps, steps == 0: quit
    unsigned* ack;                 // [1] exit reason (1: retired on request, 2: idle timeout)
    unsigned long long* out;       // published state, one 32-bit payload per 64-bit word (tag in the high half)
    unsigned seq0;                 // seq of the tick passed as WarpArgs::steps
    long long idle_clk;            // SM cycles without a doorbell after which the kernel retires by itself
};

struct Snapshot {
    double x[32], y[32], z[32], vx[32], vy[32], vz[32];
    int most[32], old_most[32];
    unsigned active_mask;
    int pad;
};


__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// gf_p / r^4 of partner p as seen from body k (the reference's |ta|^2 / gf_p), to ~1e-15.
__device__ __forceinline__ double attraction_metric(double xk, double yk, double zk, double xp, double yp, double zp, double gp) {
    double dx = xp - xk, dy = yp - yk, dz = zp - zk;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double y0 = rsqrt_seed(r2);
    double t = r2 * y0;
    double e = fma(-t, y0, 1.0);
    double y = fma(y0 * e, fma(0.375, e, 0.5), y0); // 1 / r
    double y2 = y * y;
    return gp * y2 * y2;
}

cudaError_t launch_resident_quad(essa_ctx* c, WarpArgs const& A, PersistArgs const* pa, cudaStream_t st);

} // namespace essa
