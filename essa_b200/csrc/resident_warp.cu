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
#include "resident_small.cuh"

namespace essa {

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


template<int P>
static void launch_warp_P(bool, WarpArgs const& A, cudaStream_t st) {
    k_resident_warp<P, true><<<1, 128, 0, st>>>(A); // the fast flavour runs on the two-CTA cluster (resident_quad.cu)
}

// Worlds the persistent flavour of K3q serves (capi.cu decides per call whether a live kernel can take the tick).
bool resident_persist_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    return !forces_only && !o.exact_order && c->n >= 1 && c->n <= 32 && c->world == 1
        && !(o.forward_simulated && c->closest && c->closest_n == c->n);
}

cudaError_t launch_resident_warp(essa_ctx* c, int steps, essa_step_opts const& o, bool mask_may_change, cudaStream_t st, bool persist) {
    PersistArgs PA {};
    PersistArgs const* pa = nullptr;
    if (persist && !o.exact_order) {
        PersistMailbox* mb = c->persist.mailbox;
        PA.cmd = &mb->cmd;
        PA.ack = mb->ack;
        PA.out = mb->ll;
        PA.seq0 = c->persist.seq;
        PA.idle_clk = (long long)c->persist.idle_us * 2000; // ~ SM cycles (1.965 GHz at full clock)
        pa = &PA;
    }
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
        cudaError_t e = launch_resident_quad(c, A, pa, st); // resident_quad.cu
        if (e != cudaSuccess)
            return e;
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
