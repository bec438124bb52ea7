// K3 -- resident single-CTA kernel: the whole World::update loop (src/World.cpp:86-139) for worlds
// of up to ESSA_RESIDENT_MAX bodies, all |steps| iterations inside ONE launch with positions held
// in shared memory, one thread per body, and Object::update's recording (History ring, Trail
// ring, ap/pe -- record.cuh) streamed to global memory as the steps go.
//
// Two arithmetic modes:
//   exact_order -- per-row form of World::set_forces in ascending partner order with IEEE
//                  add/mul/div/sqrt (common.cuh exact_interact): bit-identical to the reference
//                  restatement, including the most-attracting argmax.
//   fast        -- 17-instruction interaction, four independent partial sums per row.
//
// The reference evaluates set_forces() twice per step; the second result of step k is what the
// first call of step k+1 recomputes, so it is reused unless the active mask changed
// (Object::deleted() depends on the date, src/Object.cpp:60-62) or `two_pass` is set.
//
// K6 -- the same recording as a kernel of its own for the tiled path (one thread per tracked
// body), plus the constructor-time pushes (src/Object.cpp:32-43).
#include <algorithm>

#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"
#include "rows.cuh"

namespace essa {

struct ResidentArgs {
    WorldPtrs w;
    RecordPtrs r;
    long long const* cdate;
    long long const* ddate;
    uint8_t const* delflag;
    uint8_t* active_out;
    int n;
    int steps;          // |steps|, 0 = set_forces only
    double mul, dt, h;
    long long date;     // date before the first step
    long long date_step; // signed seconds per step, 0 when forward-simulated
    double eps2;
    int exact, record, offset_trails, forward_sim, argmax, two_pass, acc_valid;
    double* closest;    // [n][n][7] {distance, this xyz, other xyz} or null (forward-simulated worlds)
};

template<int MAXT, bool EXACT, bool ARGMAX>
__global__ void __launch_bounds__(MAXT, 1) k_resident(ResidentArgs const A) {
    __shared__ double sx[ESSA_RESIDENT_MAX], sy[ESSA_RESIDENT_MAX], sz[ESSA_RESIDENT_MAX], sg[ESSA_RESIDENT_MAX];
    int const k = threadIdx.x;
    int const n = A.n;
    int const npad = (n + 3) & ~3;
    bool const mine = k < n;

    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0, maxattr = 0;
    int most = -1, old_most = -1;
    bool active = false, delf = false;
    long long cd = 0, dd = 0;
    if (mine) {
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
        maxattr = A.w.maxattr[k];
        most = A.w.most[k];
        old_most = A.w.old_most[k];
        active = A.w.active[k] != 0;
        delf = A.delflag[k] != 0;
        cd = A.cdate[k];
        dd = A.ddate[k];
    }
    bool const rec = A.record && mine && k < A.r.tracked;
    BodyRec b;
    if (rec)
        rec_load(A.r, k, b);
    if (k < npad) {
        sx[k] = x;
        sy[k] = y;
        sz[k] = z;
        sg[k] = active ? gfk : 0.0;
    }
    __syncthreads();

    bool acc_valid = A.acc_valid != 0;
    long long date = A.date;

    if (A.steps == 0) { // World::set_forces() alone
        if (active)
            row_forces<EXACT, ARGMAX>(k, n, npad, sx, sy, sz, sg, x, y, z, gfk, A.eps2, ax, ay, az, maxattr, most);
    }

    for (int s = 0; s < A.steps; s++) {
        // update_history_and_date: the date moves first; deleted() may flip with it
        if (A.date_step != 0) {
            date += A.date_step;
            bool act = mine && body_active(delf, dd, cd, date);
            int changed = act != active;
            active = act;
            if (__syncthreads_or(changed)) {
                if (k < npad)
                    sg[k] = active ? gfk : 0.0;
                acc_valid = false;
                __syncthreads();
            }
        }
        if (A.closest && mine) { // Object::update_closest_approaches (src/Object.cpp:405-417), positions at step start
            for (int j = 0; j < n; j++) {
                if (j == k)
                    continue;
                double* e = A.closest + ((size_t)k * n + j) * 7;
                double d = length3(__dsub_rn(sx[j], x), __dsub_rn(sy[j], y), __dsub_rn(sz[j], z));
                double cur = e[0];
                if (cur == 0.0 || d < cur) {
                    e[0] = d;
                    e[1] = x;
                    e[2] = y;
                    e[3] = z;
                    e[4] = sx[j];
                    e[5] = sy[j];
                    e[6] = sz[j];
                }
            }
        }
        old_most = most; // Object::before_update, every object
        if (!acc_valid || A.two_pass) {
            if (active)
                row_forces<EXACT, ARGMAX>(k, n, npad, sx, sy, sz, sg, x, y, z, gfk, A.eps2, ax, ay, az, maxattr, most);
        }
        if (active) {
            vx = kick1(vx, ax, A.h, A.mul);
            vy = kick1(vy, ay, A.h, A.mul);
            vz = kick1(vz, az, A.h, A.mul);
            x = __dadd_rn(x, __dmul_rn(__dmul_rn(vx, A.dt), A.mul));
            y = __dadd_rn(y, __dmul_rn(__dmul_rn(vy, A.dt), A.mul));
            z = __dadd_rn(z, __dmul_rn(__dmul_rn(vz, A.dt), A.mul));
        }
        __syncthreads(); // every row has read the old positions (forces above, recording of the previous step)
        if (mine) {
            sx[k] = x;
            sy[k] = y;
            sz[k] = z;
        }
        __syncthreads();
        if (active) {
            row_forces<EXACT, ARGMAX>(k, n, npad, sx, sy, sz, sg, x, y, z, gfk, A.eps2, ax, ay, az, maxattr, most);
            vx = kick1(vx, ax, A.h, A.mul);
            vy = kick1(vy, ay, A.h, A.mul);
            vz = kick1(vz, az, A.h, A.mul);
        }
        acc_valid = true;
        if (rec && active) {
            record_body(A.r, k, b, x, y, z, vx, vy, vz, most, old_most, A.offset_trails != 0, A.forward_sim != 0,
                [&](int j, double& ox, double& oy, double& oz) {
                    ox = sx[j];
                    oy = sy[j];
                    oz = sz[j];
                });
        }
    }

    if (mine) {
        A.w.pos_cur[k] = make_double4(x, y, z, active ? gfk : 0.0);
        A.w.v[0][k] = vx;
        A.w.v[1][k] = vy;
        A.w.v[2][k] = vz;
        A.w.a[0][k] = ax;
        A.w.a[1][k] = ay;
        A.w.a[2][k] = az;
        A.w.maxattr[k] = maxattr;
        A.w.most[k] = most;
        A.w.old_most[k] = old_most;
        A.active_out[k] = active ? 1 : 0;
    }
    if (rec)
        rec_store(A.r, k, b);
}

// ---- K6: recording for the tiled path ------------------------------------------------------------
// Sharded contexts record the tracked bodies they OWN ([i0, i1), essa_shard_range): only their velocity, links and ap / pe are
// current on this rank (positions of every body are: the most-attracting object may live anywhere).
__global__ void k_record(WorldPtrs const W, RecordPtrs const R, int offset_trails, int forward_sim, int i0, int i1) {
    int k = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= i1 || k >= R.tracked || !W.active[k])
        return;
    BodyRec b;
    rec_load(R, k, b);
    double4 p = W.pos_cur[k];
    record_body(R, k, b, p.x, p.y, p.z, W.v[0][k], W.v[1][k], W.v[2][k], W.most[k], W.old_most[k], offset_trails != 0, forward_sim != 0,
        [&](int j, double& ox, double& oy, double& oz) {
            double4 q = W.pos_cur[j];
            ox = q.x;
            oy = q.y;
            oz = q.z;
        });
    rec_store(R, k, b);
}

// Constructor-time state (src/Object.cpp:32-43): History seeded with {pos, vel}, Trail of length 1
// that immediately receives the initial position.
__global__ void k_record_init(WorldPtrs const W, RecordPtrs const R, double* hist_first, int first) {
    int k = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= R.tracked)
        return;
    BodyRec b;
    rec_load(R, k, b);
    double4 p = W.pos_cur[k];
    double e[6] = { p.x, p.y, p.z, W.v[0][k], W.v[1][k], W.v[2][k] };
    for (int c = 0; c < 6; c++) {
        R.hist[(size_t)k * ESSA_HISTORY_SEGMENTS * 6 + c] = e[c];
        hist_first[(size_t)k * 6 + c] = e[c];
    }
    b.hist_start = 0;
    b.hist_len = 1;
    b.append_off = 1;
    b.length = 1;
    b.offx = b.offy = b.offz = 0;
    trail_set_angle(b, 0.0);
    trail_push(R, b, p.x, p.y, p.z);
    rec_store(R, k, b);
}

RecordPtrs record_ptrs(essa_ctx* c) {
    RecordPtrs r;
    r.tracked = c->tracked;
    r.hist = c->hist;
    r.hist_start = c->hist_start;
    r.hist_len = c->hist_len;
    r.verts = c->verts;
    r.vert_off = c->vert_off;
    r.cap = c->tcap;
    r.append_off = c->append_off;
    r.length = c->tlength;
    r.toff = c->toff;
    r.last_angle = c->last_angle;
    r.ap = c->appe;
    r.pe = c->appe + c->cap;
    r.ap_vel = c->appe + 2 * (size_t)c->cap;
    r.pe_vel = c->appe + 3 * (size_t)c->cap;
    return r;
}

cudaError_t launch_record(essa_ctx* c, int offset_trails, int forward_simulated, cudaStream_t st) {
    int const i0 = (int)((long long)c->n * c->rank / c->world);
    int const i1 = std::min(c->tracked, (int)((long long)c->n * (c->rank + 1) / c->world));
    if (i1 <= i0)
        return cudaSuccess;
    k_record<<<(i1 - i0 + 127) / 128, 128, 0, st>>>(world_ptrs(c), record_ptrs(c), offset_trails, forward_simulated, i0, i1);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_record_init(essa_ctx* c, int first, cudaStream_t st) {
    int cnt = c->tracked - first;
    if (cnt <= 0)
        return cudaSuccess;
    k_record_init<<<(cnt + 127) / 128, 128, 0, st>>>(world_ptrs(c), record_ptrs(c), c->hist_first, first);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_resident(essa_ctx* c, int steps, essa_step_opts const& o, bool forces_only, cudaStream_t st) {
    ResidentArgs A;
    A.w = world_ptrs(c);
    A.r = record_ptrs(c);
    A.cdate = (long long const*)c->cdate;
    A.ddate = (long long const*)c->ddate;
    A.delflag = c->delflag;
    A.active_out = c->active;
    A.n = c->n;
    A.steps = forces_only ? 0 : (steps < 0 ? -steps : steps);
    A.mul = steps >= 0 ? 1.0 : -1.0;
    A.dt = (double)o.dt_seconds;
    A.h = A.dt / 2.0;
    A.date = (long long)c->date;
    A.date_step = o.forward_simulated ? 0 : (long long)(steps >= 0 ? 1 : -1) * o.dt_seconds;
    A.eps2 = o.eps2;
    A.exact = o.exact_order;
    A.record = o.record && !forces_only;
    A.offset_trails = o.offset_trails;
    A.forward_sim = o.forward_simulated;
    bool argmax = o.track_most_attracting || o.record || o.exact_order;
    A.argmax = argmax;
    A.two_pass = o.two_pass;
    A.closest = (o.forward_simulated && c->closest && c->closest_n == c->n) ? c->closest : nullptr;
    A.acc_valid = c->acc_valid && (!argmax || c->acc_has_most) && (c->acc_exact == (o.exact_order != 0));
    int threads = (c->n + 31) / 32 * 32;
    // two register budgets: worlds of up to 256 bodies get 255 registers per thread (no spills)
    if (threads <= 256) {
        if (o.exact_order)
            k_resident<256, true, true><<<1, threads, 0, st>>>(A);
        else if (argmax)
            k_resident<256, false, true><<<1, threads, 0, st>>>(A);
        else
            k_resident<256, false, false><<<1, threads, 0, st>>>(A);
    } else {
        if (o.exact_order)
            k_resident<ESSA_RESIDENT_MAX, true, true><<<1, threads, 0, st>>>(A);
        else if (argmax)
            k_resident<ESSA_RESIDENT_MAX, false, true><<<1, threads, 0, st>>>(A);
        else
            k_resident<ESSA_RESIDENT_MAX, false, false><<<1, threads, 0, st>>>(A);
    }
    c->launches++;
    int P = A.steps == 0 ? 1 : (o.two_pass ? 2 * A.steps : A.steps + (A.acc_valid ? 0 : 1));
    c->passes += P;
    c->interactions += (double)P * (double)c->n * (double)(c->n - 1);
    c->last_force_passes = P;
    return cudaGetLastError();
}

} // namespace essa
