// K3w -- resident kernel for worlds of up to 32 bodies (the Solar System template has 10): the whole
// World::update loop (src/World.cpp:86-139) with ONE WARP on the critical path and three more
// consuming per-step snapshots from a shared-memory ring (mbarrier hand-off), so that nothing but
// the kick-drift-kick chain itself delays the next step:
//
//   warp 0  physics      forces -> kick -> drift -> forces -> kick, no block barrier per step
//   warp 3  attraction   Object::update_forces_against's most-attracting bookkeeping
//                        (src/Object.cpp:84-94): it never feeds back into the trajectory, and its
//                        result after a step depends only on the step's final positions
//   warp 1  trail        recalculate_trails_with_offset + Trail::push_back (record.cuh)
//   warp 2  history      History::move_forward + ap / pe
//
// Both arithmetic flavours (exact_order keeps the most-attracting comparison in the physics warp: it
// needs the exact terms).  set_forces() alone and closest-approach tracking go through resident.cu.
//
// A single trajectory is a dependent chain (each step needs the previous one), so this kernel is
// bound by dependent-issue latency (measured on B200: DFMA/DADD/DMUL 8.3 cycles, MUFU.RSQ64H seed
// 23, LDS 29, shuffle of a double 30), not by any throughput roofline.  What shortens the chain:
//   * lane splitting: a world of n bodies uses L = 32 / n lanes per body; lane (k, q) evaluates the
//     partners p = q, q + L, ... of body k (P of them, a compile-time bound, all independent), and
//     every lane of a body then forms the same ordered sum of the L partials, so the replicated
//     state stays bit-identical across lanes without a broadcast;
//   * kick and drift use the exact identities (a h) mul == a (h mul) and (v dt) mul == v (dt mul)
//     (multiplying by +-1 is exact);
//   * positions are exchanged through shared memory with __syncwarp only.
#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"

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

struct Snapshot {
    double x[32], y[32], z[32], vx[32], vy[32], vz[32];
    int most[32], old_most[32];
    unsigned active_mask;
    int pad;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Partial force of lane (k, q) over its partners, then the ordered sum over the L lanes of body k.
template<int P>
__device__ __forceinline__ void lane_forces(int k, int q, int L, int iters, int n, double const* sx, double const* sy, double const* sz,
    double const* sg, double x, double y, double z, double eps2, double& ax, double& ay, double& az) {
    double fx[P], fy[P], fz[P];
#pragma unroll
    for (int u = 0; u < P; u++)
        fx[u] = fy[u] = fz[u] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < P; u++) {
            int p = q + (it * P + u) * L;
            bool in = p < n;
            int pc = in ? p : k; // out of range -> the self term, which vanishes
            double4 qv = make_double4(sx[pc], sy[pc], sz[pc], in ? sg[pc] : 0.0);
            fast_interact(x, y, z, qv, eps2, fx[u], fy[u], fz[u]);
        }
    }
    double px = fx[0], py = fy[0], pz = fz[0];
#pragma unroll
    for (int u = 1; u < P; u++) {
        px += fx[u];
        py += fy[u];
        pz += fz[u];
    }
    double sxm = 0, sym = 0, szm = 0;
    for (int qq = 0; qq < L; qq++) {
        int src = k + qq * n;
        sxm += __shfl_sync(0xffffffffu, px, src);
        sym += __shfl_sync(0xffffffffu, py, src);
        szm += __shfl_sync(0xffffffffu, pz, src);
    }
    ax = sxm;
    ay = sym;
    az = szm;
}

// exact_order flavour: the expensive part of a term -- three IEEE divisions and a square root -- does not
// depend on the summation order, so the L lanes of a body evaluate their partners' terms in parallel
// (common.cuh exact arithmetic); every lane of the body then adds ALL terms in ascending partner order
// (p = qq + slot * L is ascending in (slot, qq) order), which is the reference's order bit for bit, and
// runs the most-attracting comparison of src/Object.cpp:84-94 in the same order.
template<int P>
__device__ __forceinline__ void lane_forces_exact(int k, int q, int L, int iters, int n, double const* sx, double const* sy,
    double const* sz, double const* sg, double x, double y, double z, double gfk, double& ax, double& ay, double& az, double& maxattr,
    int& most) {
    double fx = 0, fy = 0, fz = 0, best = 0;
    int bp = -1;
    for (int it = 0; it < iters; it++) {
        double tx[P], ty[P], tz[P], mg[P];
        int ok[P];
#pragma unroll
        for (int u = 0; u < P; u++) {
            int p = q + (it * P + u) * L;
            bool in = p < n && p != k;
            int pc = in ? p : (k + 1 < n ? k + 1 : 0); // any other body: keeps the arithmetic finite, result unused
            double gp = sg[pc];
            in = in && gp != 0.0;
            double dx = __dsub_rn(sx[pc], x), dy = __dsub_rn(sy[pc], y), dz = __dsub_rn(sz[pc], z);
            double den = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            den = __dmul_rn(den, __dsqrt_rn(den));
            tx[u] = __dmul_rn(__ddiv_rn(dx, den), gp);
            ty[u] = __dmul_rn(__ddiv_rn(dy, den), gp);
            tz[u] = __dmul_rn(__ddiv_rn(dz, den), gp);
            mg[u] = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(tx[u], tx[u]), __dmul_rn(ty[u], ty[u])), __dmul_rn(tz[u], tz[u])), gp);
            ok[u] = in ? 1 : 0;
        }
#pragma unroll
        for (int u = 0; u < P; u++) {
            for (int qq = 0; qq < L; qq++) {
                int src = k + qq * n;
                double vx = __shfl_sync(0xffffffffu, tx[u], src);
                double vy = __shfl_sync(0xffffffffu, ty[u], src);
                double vz = __shfl_sync(0xffffffffu, tz[u], src);
                double vm = __shfl_sync(0xffffffffu, mg[u], src);
                int vok = __shfl_sync(0xffffffffu, ok[u], src);
                int p = qq + (it * P + u) * L;
                if (vok) {
                    fx = __dadd_rn(fx, vx);
                    fy = __dadd_rn(fy, vy);
                    fz = __dadd_rn(fz, vz);
                    if (vm > best && sg[p] > gfk) {
                        best = vm;
                        bp = p;
                    }
                }
            }
        }
    }
    ax = fx;
    ay = fy;
    az = fz;
    maxattr = best;
    if (bp >= 0)
        most = bp;
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

template<int P, bool EXACT>
__global__ void __launch_bounds__(128, 1) k_resident_warp(WarpArgs const A) {
    __shared__ double sx[32], sy[32], sz[32], sg[32];
    __shared__ double sgf[32]; // attraction warp: every body's gf (who is heavier than whom)
    __shared__ Snapshot ring[kRing];
    // full: physics -> attraction;  ready: attraction -> recorders;  empty: consumers -> physics
    __shared__ __align__(8) unsigned long long full[kRing], ready[kRing], empty[kRing];

    int const n = A.n, L = A.L;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const k = lane % n, q = lane / n; // (body, split) of a lane
    bool const worker = lane < n * L;
    bool const leader = lane < n;
    bool const rec_on = A.record != 0;
    // exact_order keeps the most-attracting comparison in the physics warp (it needs the exact terms), so
    // the attraction warp is idle and snapshots exist for the recorders only
    bool const snap_on = EXACT ? rec_on : (rec_on || A.argmax != 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRing; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&ready[i], 1);
            mbar_init(&empty[i], rec_on ? 2 : 1); // the recorder warps, or the attraction warp alone
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 3) {
        // ---------------- attraction warp ----------------
        if (!snap_on || EXACT)
            return;
        double gfk = 0, maxattr = 0;
        int most = -1, old_most = -1;
        if (worker) {
            gfk = A.w.gf[k];
            maxattr = A.w.maxattr[k];
            most = A.w.most[k];
            old_most = A.w.old_most[k];
        }
        if (leader)
            sgf[k] = gfk;
        __syncwarp();
        int const per = (n + L - 1) / L; // partners per lane
        for (int s = 0; s < A.steps; s++) {
            int const slot = s % kRing;
            mbar_wait(&full[slot], (unsigned)(s / kRing) & 1u);
#ifdef ESSA_STAGE_CLOCKS
            long long t0 = clock64();
#endif
            Snapshot& sn = ring[slot];
            unsigned const amask = sn.active_mask;
            bool const active = worker && ((amask >> k) & 1u);
            old_most = most; // Object::before_update
            double best = 0;
            int bp = -1;
            double const xk = sn.x[k], yk = sn.y[k], zk = sn.z[k];
            for (int u = 0; u < per; u++) {
                int p = q + u * L;
                if (p < n && p != k && ((amask >> p) & 1u)) {
                    double gp = sgf[p];
                    double m = attraction_metric(xk, yk, zk, sn.x[p], sn.y[p], sn.z[p], gp);
                    if (m > best && gp > gfk) { // ascending p within a lane: the lowest index wins ties
                        best = m;
                        bp = p;
                    }
                }
            }
            double tb = 0;
            int tp = -1;
            for (int qq = 0; qq < L; qq++) {
                int src = k + qq * n;
                double m = __shfl_sync(0xffffffffu, best, src);
                int j = __shfl_sync(0xffffffffu, bp, src);
                if (m > tb || (m == tb && j >= 0 && tp >= 0 && j < tp)) {
                    tb = m;
                    tp = j;
                }
            }
            if (active) { // masked bodies keep max_attraction and the link (clear_forces skips them)
                maxattr = tb;
                if (tp >= 0)
                    most = tp;
            }
            if (leader) {
                sn.most[k] = most;
                sn.old_most[k] = old_most;
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(rec_on ? &ready[slot] : &empty[slot]);
        }
        if (leader) {
            A.w.maxattr[k] = maxattr;
            A.w.most[k] = most;
            A.w.old_most[k] = old_most;
        }
        return;
    }

    if (warp >= 1) {
        // ---------------- recorder warps: 1 = Trail, 2 = History + ap / pe ----------------
        if (!rec_on)
            return;
        bool const mine = lane < n && lane < A.r.tracked;
        BodyRec b;
        if (mine)
            rec_load(A.r, lane, b);
        for (int s = 0; s < A.steps; s++) {
            int const slot = s % kRing;
            mbar_wait(&ready[slot], (unsigned)(s / kRing) & 1u);
            Snapshot const& sn = ring[slot];
            if (mine && ((sn.active_mask >> lane) & 1u)) {
                int const most = sn.most[lane];
                int const mi = most >= 0 ? most : 0;
                double const mx = sn.x[mi], my = sn.y[mi], mz = sn.z[mi];
                if (warp == 1) {
                    record_trail(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane], mx, my, mz, most,
                        sn.old_most[lane], A.offset_trails != 0, A.forward_sim != 0, false);
                } else {
                    if (!A.forward_sim)
                        history_push(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane]);
                    record_appe(b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane], mx, my, mz, most);
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[slot]);
        }
        if (mine) {
            if (warp == 1) {
                rec_store_trail(A.r, lane, b, false); // the History ring's cursor belongs to warp 2
            } else {
                rec_store_appe(A.r, lane, b);
                A.r.hist_start[lane] = b.hist_start;
                A.r.hist_len[lane] = b.hist_len;
            }
        }
        return;
    }

    // ---------------- physics warp ----------------
    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0, maxattr = 0;
    int most = -1, old_most = -1; // exact_order only
    bool active = false, delf = false;
    long long cd = 0, dd = 0;
    if (worker) {
        double4 p = A.w.pos_cur[k];
        x = p.x;
        y = p.y;
        z = p.z;
        gfk = A.w.gf[k];
        vx = A.w.v[0][k];
        vy = A.w.v[1][k];
        vz = A.w.v[2][k];
        ax = A.w.a[0][k];
        ay = A.w.a[1][k];
        az = A.w.a[2][k];
        active = A.w.active[k] != 0;
        delf = A.delflag[k] != 0;
        cd = A.cdate[k];
        dd = A.ddate[k];
        if (EXACT) {
            maxattr = A.w.maxattr[k];
            most = A.w.most[k];
            old_most = A.w.old_most[k];
        }
    }
    if (leader) {
        sx[k] = x;
        sy[k] = y;
        sz[k] = z;
        sg[k] = active ? gfk : 0.0;
    }
    __syncwarp();

    bool acc_valid = A.acc_valid != 0;
    long long date = A.date;
    double const hm = A.hm, dtm = A.dtm;

    auto forces = [&]() {
        double fx, fy, fz, fm = maxattr;
        int fmost = most;
        if (EXACT)
            lane_forces_exact<P>(k, q, L, A.iters, n, sx, sy, sz, sg, x, y, z, gfk, fx, fy, fz, fm, fmost);
        else
            lane_forces<P>(k, q, L, A.iters, n, sx, sy, sz, sg, x, y, z, A.eps2, fx, fy, fz);
        if (active) {
            ax = fx;
            ay = fy;
            az = fz;
            maxattr = fm;
            most = fmost;
        }
    };

    for (int s = 0; s < A.steps; s++) {
        if (A.date_step != 0) { // update_history_and_date: the mask may flip with the date
            date += A.date_step;
            bool act = worker && body_active(delf, dd, cd, date);
            bool changed = act != active;
            active = act;
            if (__any_sync(0xffffffffu, changed)) {
                if (leader)
                    sg[k] = active ? gfk : 0.0;
                acc_valid = false;
                __syncwarp();
            }
        }
        old_most = most; // Object::before_update (exact_order; the attraction warp does it otherwise)
        if (!acc_valid || A.two_pass)
            forces();
        if (active) { // first half kick + drift (src/World.cpp:112,115)
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
            x = __dadd_rn(x, __dmul_rn(vx, dtm));
            y = __dadd_rn(y, __dmul_rn(vy, dtm));
            z = __dadd_rn(z, __dmul_rn(vz, dtm));
        }
        __syncwarp();
        if (leader) {
            sx[k] = x;
            sy[k] = y;
            sz[k] = z;
        }
        __syncwarp();
        forces();
        if (active) { // second half kick (src/World.cpp:127)
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
        }
        acc_valid = true;
        if (snap_on) {
            int const slot = s % kRing;
            if (s >= kRing)
                mbar_wait(&empty[slot], (unsigned)(s / kRing - 1) & 1u);
            Snapshot& sn = ring[slot];
            unsigned amask = __ballot_sync(0xffffffffu, leader && active);
            if (leader) {
                sn.x[k] = x;
                sn.y[k] = y;
                sn.z[k] = z;
                sn.vx[k] = vx;
                sn.vy[k] = vy;
                sn.vz[k] = vz;
                if (EXACT) {
                    sn.most[k] = most;
                    sn.old_most[k] = old_most;
                }
            }
            if (lane == 0)
                sn.active_mask = amask;
            __syncwarp();
            if (lane == 0)
                mbar_arrive(EXACT ? &ready[slot] : &full[slot]);
        }
    }

    if (leader) {
        A.w.pos_cur[k] = make_double4(x, y, z, active ? gfk : 0.0);
        A.w.v[0][k] = vx;
        A.w.v[1][k] = vy;
        A.w.v[2][k] = vz;
        A.w.a[0][k] = ax;
        A.w.a[1][k] = ay;
        A.w.a[2][k] = az;
        A.active_out[k] = active ? 1 : 0;
        if (EXACT) {
            A.w.maxattr[k] = maxattr;
            A.w.most[k] = most;
            A.w.old_most[k] = old_most;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// K3q -- the fast flavour as a CLUSTER OF TWO CTAs (two SMs) talking through distributed shared memory.
//
// Every FP64 instruction of a warp occupies its SM sub-partition's FP64 pipe for two cycles (measured), and
// one step of a 10-body world needs several hundred of them for the physics alone, so on a single SM the
// attraction / recorder warps steal issue slots from the chain that limits the run.  Hence:
//
//   CTA 0  warps 0-3  physics.  Every warp carries the whole replicated state (positions, velocities,
//                     accelerations: identical bits in all four) but evaluates only every fourth partner
//                     slot; the four partial accelerations meet in shared memory behind ONE named barrier
//                     per force pass and are added in a fixed order by every lane, so the replicas stay
//                     identical and no position is ever broadcast.  Warps 1-3 publish the step's snapshot
//                     (positions / velocities / mask) straight into CTA 1's ring with st.async: the store
//                     itself completes the ring slot's mbarrier there (transaction bytes), so publishing
//                     costs the physics warps seven stores and no handshake.
//   CTA 1  warps 0-2  attraction: most-attracting links, steps dealt round-robin (a step's link depends
//                     only on that step's final positions)
//          warps 3, 5 Trail (even / odd bodies)
//          warp 4     link bookkeeping + History
//          warp 6     ap / pe
//                     The consumers report progress with a release store into CTA 0's shared memory;
//                     the physics warps look at it only when their cached copy says the ring could be full.
// ------------------------------------------------------------------------------------------------
constexpr int kQuadWarps = 4;
constexpr int kAttrWarps = 3;
constexpr int kHistWarp = 4, kConsumerEnd = 7; // CTA 1: warps 3 .. 6 consume
constexpr int kQuadThreads = 32 * kConsumerEnd;

template<int T, int LOG2L>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQuadThreads, 1) k_resident_quad(WarpArgs const A) {
    __shared__ double spos[kQuadWarps][4][32];            // CTA 0: per physics warp sx, sy, sz, sg (private copies)
    __shared__ double part[2][kQuadWarps][3][32];         // CTA 0: partial accelerations, double-buffered
    __shared__ double sgf[32];                            // CTA 1
    __shared__ __align__(16) Snapshot ring[kRing];        // CTA 1
    __shared__ __align__(8) unsigned long long full[kRing], ready[kRing]; // CTA 1
    __shared__ unsigned done[4];                          // CTA 0: steps consumed by CTA 1's warps 3 .. 6

    int const n = A.n, L = A.L; // attraction / recorder warps: L = 32 / n lanes per body, lane = k + q n
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned const cta = cluster_ctarank();
    bool const rec_on = A.record != 0;
    bool const argmax = A.argmax != 0;
    bool const snap_on = rec_on || argmax;
    unsigned const snap_bytes = (unsigned)n * 48u + 8u;

    // full: st.async from CTA 0 (transaction bytes) -> the attraction warp that owns the step;
    // ready: that warp -> warps 3, 4
    if (threadIdx.x == 0) {
        if (cta == 1) {
            for (int i = 0; i < kRing; i++) {
                mbar_init(&full[i], 1);
                mbar_init(&ready[i], 1);
            }
            mbar_fence_init();
            for (int i = 0; i < kRing; i++)
                mbar_expect_tx(&full[i], snap_bytes);
        } else {
            done[0] = done[1] = done[2] = done[3] = 0;
        }
    }
    __syncthreads();
    cluster_sync_all();

    if (cta == 1 && snap_on && warp < kAttrWarps) {
        // ---------------- attraction warps ----------------
        // Which heavier body pulls hardest on body k (src/Object.cpp:84-94) depends only on a step's final
        // positions, so different steps are evaluated by different warps; the one sequential piece -- "keep
        // the previous link when there is no candidate" -- is a register in the consumer warps.
        // Candidates are ranked by gf_p y0^4 with y0 the rsqrt seed (relative error < 4e-6: enough unless two
        // heavier bodies attract within 4 ppm of each other); max_attraction is formed for the final state only.
        int const k = lane % n, q = lane / n;
        bool const worker = lane < n * L;
        bool const leader = lane < n;
        double gfk = worker ? A.w.gf[k] : 0.0;
        if (warp == 0 && leader)
            sgf[k] = gfk;
        asm volatile("bar.sync 2, %0;" ::"r"(32 * kAttrWarps) : "memory");
        constexpr int kCand = 4;
        int const per = (n + L - 1) / L;
#ifdef ESSA_STAGE_CLOCKS
        long long abusy = 0;
#endif
        for (int s = warp; s < A.steps; s += kAttrWarps) {
            int const slot = s % kRing;
            mbar_wait(&full[slot], (unsigned)(s / kRing) & 1u);
            if (lane == 0)
                mbar_expect_tx(&full[slot], snap_bytes); // armed for the slot's next use
#ifdef ESSA_STAGE_CLOCKS
            long long t0 = clock64();
#endif
            Snapshot& sn = ring[slot];
            unsigned const amask = sn.active_mask;
            double const xk = sn.x[k], yk = sn.y[k], zk = sn.z[k];
            double best = 0;
            int bp = -1;
            for (int u0 = 0; u0 < per; u0 += kCand) {
                double m[kCand];
                int pi[kCand];
#pragma unroll
                for (int u = 0; u < kCand; u++) {
                    int p = q + (u0 + u) * L;
                    bool ok = (u0 + u) < per && p < n && p != k && ((amask >> p) & 1u);
                    int pc = ok ? p : k;
                    double gp = sgf[pc];
                    ok = ok && gp > gfk;
                    double dx = sn.x[pc] - xk, dy = sn.y[pc] - yk, dz = sn.z[pc] - zk;
                    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    double y0 = rsqrt_seed(r2);
                    double y2 = y0 * y0;
                    m[u] = ok ? gp * y2 * y2 : 0.0;
                    pi[u] = ok ? p : -1;
                }
#pragma unroll
                for (int u = 0; u < kCand; u++) // ascending p: strict '>' keeps the lowest index on ties
                    if (m[u] > best) {
                        best = m[u];
                        bp = pi[u];
                    }
            }
            double tb = 0;
            int tp = -1;
            for (int qq = 0; qq < L; qq++) {
                int src = k + qq * n;
                double mm = __shfl_sync(0xffffffffu, best, src);
                int j = __shfl_sync(0xffffffffu, bp, src);
                if (mm > tb || (mm == tb && j >= 0 && tp >= 0 && j < tp)) {
                    tb = mm;
                    tp = j;
                }
            }
            if (leader)
                sn.most[k] = ((amask >> k) & 1u) ? tp : -1; // the step's winner, -1: none (or body masked)
            __syncwarp();
#ifdef ESSA_STAGE_CLOCKS
            abusy += clock64() - t0;
#endif
            if (lane == 0)
                mbar_arrive(&ready[slot]);
        }
#ifdef ESSA_STAGE_CLOCKS
        if (lane == 0)
            printf("attraction warp %d: %lld busy clk per step of its own\n", warp, abusy / (A.steps / kAttrWarps + 1));
#endif
    } else if (cta == 1 && snap_on && warp < kConsumerEnd && (warp == kHistWarp || rec_on)) {
        // ---------------- consumer warps ----------------
        //   3, 5  Trail (bodies with even / odd index: an append is ~4x the cost of a decimated step and stalls
        //         the whole warp, so two warps halve how often that happens to each)
        //   4     link bookkeeping + History
        //   6     ap / pe
        // Every consumer follows the most-attracting links itself (a few integer instructions per step).
        bool const body = lane < n;
        bool const is_trail = warp == 3 || warp == 5;
        bool const mine = rec_on && body && lane < A.r.tracked && (!is_trail || (lane & 1) == (warp == 5 ? 1 : 0));
        uint32_t const done_remote = mapa_u32(&done[warp - 3], 0);
        BodyRec b;
        if (mine)
            rec_load(A.r, lane, b);
        int most = body ? A.w.most[lane] : -1, old_most = body ? A.w.old_most[lane] : -1;
        bool const offset_trails = A.offset_trails != 0, forward_sim = A.forward_sim != 0;
#ifdef ESSA_STAGE_CLOCKS
        long long busy = 0;
#endif
        for (int s = 0; s < A.steps; s++) {
            int const slot = s % kRing;
            mbar_wait(&ready[slot], (unsigned)(s / kRing) & 1u);
#ifdef ESSA_STAGE_CLOCKS
            long long t0 = clock64();
#endif
            Snapshot const& sn = ring[slot];
            bool const act = body && ((sn.active_mask >> lane) & 1u);
            int const winner = body ? sn.most[lane] : -1;
            old_most = most; // Object::before_update (every object)
            if (act && winner >= 0)
                most = winner;
            if (mine && act) {
                int const mi = most >= 0 ? most : 0;
                double const mx = sn.x[mi], my = sn.y[mi], mz = sn.z[mi];
                if (is_trail)
                    record_trail(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane], mx, my, mz, most,
                        old_most, offset_trails, forward_sim, false);
                else if (warp == kHistWarp) {
                    if (!forward_sim)
                        history_push(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane]);
                } else
                    record_appe(b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane], mx, my, mz, most);
            }
            if (warp == kHistWarp && s == A.steps - 1 && act) {
                // max_attraction as the last force pass left it: the winner's gf / r^4 to full precision, or 0
                A.w.maxattr[lane] = winner >= 0 ? attraction_metric(sn.x[lane], sn.y[lane], sn.z[lane], sn.x[winner], sn.y[winner],
                                                      sn.z[winner], A.w.gf[winner])
                                                : 0.0;
            }
            __syncwarp();
#ifdef ESSA_STAGE_CLOCKS
            busy += clock64() - t0;
#endif
            if (lane == 0)
                st_relaxed_cluster_u32(done_remote, (unsigned)(s + 1));
        }
#ifdef ESSA_STAGE_CLOCKS
        if (lane == 0)
            printf("consumer warp %d: %lld busy clk/step\n", warp, busy / A.steps);
#endif
        if (warp == kHistWarp && body) {
            A.w.most[lane] = most;
            A.w.old_most[lane] = old_most;
        }
        if (mine) {
            if (is_trail) {
                rec_store_trail(A.r, lane, b, false);
            } else if (warp == kHistWarp) {
                A.r.hist_start[lane] = b.hist_start;
                A.r.hist_len[lane] = b.hist_len;
            } else {
                rec_store_appe(A.r, lane, b);
            }
        }
    } else if (cta == 0 && warp < kQuadWarps) {
        // ---------------- physics warps 0 .. 3 ----------------
        // Lane layout here: PL = 2^LOG2L lanes per body in one contiguous group (lane = kb PL + qb), so that the
        // group sum is a shuffle butterfly: every lane ends with the same bits (a + b == b + a exactly).
        constexpr int PL = 1 << LOG2L;
        int const kb = lane >> LOG2L, qb = lane & (PL - 1);
        bool const pworker = kb < n;
        bool const pleader = pworker && qb == 0;
        int const kc = pworker ? kb : 0;
        double* const sx = spos[warp][0];
        double* const sy = spos[warp][1];
        double* const sz = spos[warp][2];
        double* const sg = spos[warp][3];
        double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0;
        bool active = false, delf = false;
        long long cd = 0, dd = 0;
        if (pworker) {
            double4 p = A.w.pos_cur[kb];
            x = p.x;
            y = p.y;
            z = p.z;
            gfk = A.w.gf[kb];
            vx = A.w.v[0][kb];
            vy = A.w.v[1][kb];
            vz = A.w.v[2][kb];
            ax = A.w.a[0][kb];
            ay = A.w.a[1][kb];
            az = A.w.a[2][kb];
            active = A.w.active[kb] != 0;
            delf = A.delflag[kb] != 0;
            cd = A.cdate[kb];
            dd = A.ddate[kb];
        }
        if (pleader) {
            sx[kb] = x;
            sy[kb] = y;
            sz[kb] = z;
            sg[kb] = active ? gfk : 0.0;
        }
        __syncwarp();

        // this warp's partner slots of this lane never change: indices and weights are loop invariants
        int pidx[T];
        double pgf[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            int p = qb + (warp + kQuadWarps * t) * PL;
            bool in = p < n;
            pidx[t] = in ? p : kc; // out of range -> the self term, which vanishes
            pgf[t] = in ? sg[pidx[t]] : 0.0;
        }
        double* const my_part0 = &part[0][warp][0][kc];
        double* const my_part1 = &part[1][warp][0][kc];
        double const* const all_part0 = &part[0][0][0][kc];
        double const* const all_part1 = &part[1][0][0][kc];

        bool acc_valid = A.acc_valid != 0;
        long long date = A.date;
        double const hm = A.hm, dtm = A.dtm;
        int par = 0;
        // bit kb of the body mask <- bit (kb PL) of the leaders' ballot; it only changes with the date
        auto body_mask = [&]() {
            unsigned lead = __ballot_sync(0xffffffffu, pleader && active);
            unsigned m = 0;
            for (int b2 = 0; b2 < n; b2++)
                m |= ((lead >> (b2 << LOG2L)) & 1u) << b2;
            return m;
        };
        unsigned amask = body_mask();

        // where this lane's part of a snapshot goes in CTA 1 (slot 0), and slot 0's mbarrier there
        uint32_t const full_remote = mapa_u32(&full[0], 1);
        uint32_t snap_remote = mapa_u32(&ring[0], 1);
        if (warp == 1)
            snap_remote += (uint32_t)offsetof(Snapshot, x) + 8u * kc;
        else if (warp == 2)
            snap_remote += (uint32_t)offsetof(Snapshot, vx) + 8u * kc;
        else
            snap_remote += (uint32_t)offsetof(Snapshot, active_mask);
        int consumed = 0; // cached lower bound of the steps CTA 1 has finished with

        auto forces = [&]() {
            double fx[T], fy[T], fz[T];
#pragma unroll
            for (int t = 0; t < T; t++) {
                fx[t] = fy[t] = fz[t] = 0.0;
                int const pc = pidx[t];
                double4 qv = make_double4(sx[pc], sy[pc], sz[pc], pgf[t]);
                fast_interact(x, y, z, qv, A.eps2, fx[t], fy[t], fz[t]);
            }
            double px = fx[0], py = fy[0], pz = fz[0];
#pragma unroll
            for (int t = 1; t < T; t++) {
                px += fx[t];
                py += fy[t];
                pz += fz[t];
            }
#pragma unroll
            for (int off = PL >> 1; off > 0; off >>= 1) {
                px += __shfl_xor_sync(0xffffffffu, px, off);
                py += __shfl_xor_sync(0xffffffffu, py, off);
                pz += __shfl_xor_sync(0xffffffffu, pz, off);
            }
            double* const mine = par ? my_part1 : my_part0;
            double const* const all = par ? all_part1 : all_part0;
            if (pleader) {
                mine[0] = px;
                mine[32] = py;
                mine[64] = pz;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // part[par][w][c][k]: w stride 96 doubles, c stride 32
            double tx = (all[0] + all[96]) + (all[192] + all[288]);
            double ty = (all[32] + all[128]) + (all[224] + all[320]);
            double tz = (all[64] + all[160]) + (all[256] + all[352]);
            par ^= 1;
            if (active) {
                ax = tx;
                ay = ty;
                az = tz;
            }
        };

        for (int s = 0; s < A.steps; s++) {
            if (A.date_step != 0) { // only when the active mask can flip during this call
                date += A.date_step;
                bool act = pworker && body_active(delf, dd, cd, date);
                bool changed = act != active;
                active = act;
                if (__any_sync(0xffffffffu, changed)) { // every physics warp sees the same change
                    if (pleader)
                        sg[kb] = active ? gfk : 0.0;
                    acc_valid = false;
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < T; t++)
                        pgf[t] = (qb + (warp + kQuadWarps * t) * PL) < n ? sg[pidx[t]] : 0.0;
                    amask = body_mask();
                }
            }
            if (!acc_valid || A.two_pass)
                forces();
            if (active) {
                vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                vz = __dadd_rn(vz, __dmul_rn(az, hm));
                x = __dadd_rn(x, __dmul_rn(vx, dtm));
                y = __dadd_rn(y, __dmul_rn(vy, dtm));
                z = __dadd_rn(z, __dmul_rn(vz, dtm));
            }
            __syncwarp();
            if (pleader) {
                sx[kb] = x;
                sy[kb] = y;
                sz[kb] = z;
            }
            __syncwarp();
            forces();
            if (active) {
                vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                vz = __dadd_rn(vz, __dmul_rn(az, hm));
            }
            acc_valid = true;
            if (snap_on && warp != 0) {
                int const slot = s % kRing;
                if (s - consumed >= kRing) { // the slot may still be in use: look at CTA 1's progress
                    do {
                        consumed = (int)ld_acquire_cluster_u32(&done[kHistWarp - 3]);
                        if (rec_on) {
                            int const d0 = (int)ld_acquire_cluster_u32(&done[0]), d2 = (int)ld_acquire_cluster_u32(&done[2]),
                                      d3 = (int)ld_acquire_cluster_u32(&done[3]);
                            consumed = min(min(consumed, d0), min(d2, d3));
                        }
                    } while (s - consumed >= kRing);
                }
                uint32_t const dst = snap_remote + (uint32_t)slot * (uint32_t)sizeof(Snapshot);
                uint32_t const bar = full_remote + 8u * (uint32_t)slot;
                if (warp == 1) {
                    if (pleader) {
                        st_async_f64(dst, x, bar);
                        st_async_f64(dst + 256u, y, bar);
                        st_async_f64(dst + 512u, z, bar);
                    }
                } else if (warp == 2) {
                    if (pleader) {
                        st_async_f64(dst, vx, bar);
                        st_async_f64(dst + 256u, vy, bar);
                        st_async_f64(dst + 512u, vz, bar);
                    }
                } else if (lane == 0) {
                    st_async_u64(dst, (unsigned long long)amask, bar);
                }
            }
        }

        if (warp == 0 && pleader) {
            A.w.pos_cur[kb] = make_double4(x, y, z, active ? gfk : 0.0);
            A.w.v[0][kb] = vx;
            A.w.v[1][kb] = vy;
            A.w.v[2][kb] = vz;
            A.w.a[0][kb] = ax;
            A.w.a[1][kb] = ay;
            A.w.a[2][kb] = az;
            A.active_out[kb] = active ? 1 : 0;
        }
    }
    // nobody leaves while a peer may still store into (or signal) its shared memory
    cluster_sync_all();
}

template<int T, int LOG2L>
static void launch_quad(WarpArgs const& A, cudaStream_t st) {
    k_resident_quad<T, LOG2L><<<2, kQuadThreads, 0, st>>>(A); // one cluster of two CTAs
}

template<int P>
static void launch_warp_P(bool exact, WarpArgs const& A, cudaStream_t st) {
    if (exact)
        k_resident_warp<P, true><<<1, 128, 0, st>>>(A);
    else
        k_resident_warp<P, false><<<1, 128, 0, st>>>(A);
}

cudaError_t launch_resident_warp(essa_ctx* c, int steps, essa_step_opts const& o, bool mask_may_change, cudaStream_t st) {
    WarpArgs A;
    A.w = world_ptrs(c);
    A.r = record_ptrs(c);
    A.cdate = (long long const*)c->cdate;
    A.ddate = (long long const*)c->ddate;
    A.delflag = c->delflag;
    A.active_out = c->active;
    A.n = c->n;
    A.steps = steps < 0 ? -steps : steps;
    double const mul = steps >= 0 ? 1.0 : -1.0;
    A.dtm = (double)o.dt_seconds * mul;
    A.hm = (double)o.dt_seconds / 2.0 * mul;
    A.date = (long long)c->date;
    // the kernels need the date only to re-evaluate Object::deleted(); the host moves the clock itself
    A.date_step = (o.forward_simulated || !mask_may_change) ? 0 : (long long)(steps >= 0 ? 1 : -1) * o.dt_seconds;
    A.eps2 = o.eps2;
    A.record = o.record && c->tracked > 0;
    A.argmax = o.track_most_attracting || o.record;
    A.offset_trails = o.offset_trails;
    A.forward_sim = o.forward_simulated;
    A.two_pass = o.two_pass;
    // the acceleration left by an earlier call is reusable whatever most-attracting tracking came with it
    bool const exact = o.exact_order != 0;
    A.acc_valid = exact ? (c->acc_valid && c->acc_exact && c->acc_has_most) : (c->acc_valid && !c->acc_exact);
    // lanes per body and partners per lane
    A.L = 32 / c->n;
    int per_lane = (c->n + A.L - 1) / A.L;
    int P = per_lane <= 1 ? 1 : (per_lane <= 2 ? 2 : (per_lane <= 4 ? 4 : 8));
    A.iters = (per_lane + P - 1) / P;
    if (!exact) {
        // fast flavour: four physics warps, each with ceil(per_lane / 4) partner slots per lane
        // physics lanes per body: the largest power of two that fits 32 lanes
        int log2l = 0;
        while ((2 << log2l) * c->n <= 32)
            log2l++;
        int const pl = 1 << log2l;
        int const slots = (c->n + pl - 1) / pl;               // partner slots per lane
        int const T = (slots + kQuadWarps - 1) / kQuadWarps;   // ... per physics warp
        if (log2l == 4)
            launch_quad<1, 4>(A, st);
        else if (log2l == 3)
            launch_quad<1, 3>(A, st);
        else if (log2l == 2)
            launch_quad<1, 2>(A, st);
        else if (log2l == 1 && T <= 1)
            launch_quad<1, 1>(A, st);
        else if (log2l == 1)
            launch_quad<2, 1>(A, st);
        else if (T <= 1)
            launch_quad<1, 0>(A, st);
        else if (T <= 2)
            launch_quad<2, 0>(A, st);
        else if (T <= 4)
            launch_quad<4, 0>(A, st);
        else
            launch_quad<8, 0>(A, st);
    } else if (P == 1)
        launch_warp_P<1>(exact, A, st);
    else if (P == 2)
        launch_warp_P<2>(exact, A, st);
    else if (P == 4)
        launch_warp_P<4>(exact, A, st);
    else
        launch_warp_P<8>(exact, A, st);
    c->launches++;
    int passes = o.two_pass ? 2 * A.steps : A.steps + (A.acc_valid ? 0 : 1);
    c->passes += passes;
    c->interactions += (double)passes * (double)c->n * (double)(c->n - 1);
    c->last_force_passes = passes;
    return cudaGetLastError();
}

} // namespace essa
