// K3c -- resident single-CTA kernel for worlds of 33 .. 128 bodies in fast arithmetic
// (jupiter_system.essa: Jupiter + 76 moons): the whole World::update loop (src/World.cpp:86-139)
// inside one launch, like K3, but laid out for throughput on ONE SM:
//
//   physics warps   L lanes per body (L = 2..8, a power of two, as many as fit 768 threads): lane (k, q) sums the
//                   partners p = q, q + L, ... of body k, a shuffle tree over the L-lane group hands
//                   the total to the group's leader, which alone kicks, drifts and publishes the new
//                   position in shared memory.  A 77-body world keeps 616 lanes (20 warps, all four
//                   FP64 sub-partitions) busy instead of 77 lanes of 3 warps.  Two named barriers
//                   (bar.sync 1) per step, for the physics warps only.
//   recorder warps  one lane per body, in two groups (Trail / History + ap-pe, record.cuh), fed by
//                   per-step snapshots in a shared-memory ring (mbarrier full / empty), off the
//                   physics critical path.
//
// The most-attracting bookkeeping (src/Object.cpp:84-94) rides in the physics lanes at ONE extra
// FP64 multiply per interaction: the candidates are ranked by gf_j y0^4 with y0 the rsqrt seed
// (relative error < 4e-6, enough to rank unless two heavier bodies attract within 4 ppm of each
// other), and max_attraction is recomputed to full precision for the winner only.
//
// Bound: FP64 pipe of one SM -- n (n-1) x 20 instructions / 64 lanes per pass (77 bodies: 1,830
// cycles); measured numbers in profiles/README.md.  Worlds whose active mask changes during the
// call, two_pass, exact_order and closest-approach tracking go through resident.cu.
#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"

namespace essa {

constexpr int kCtaMaxBodies = 128; // partners per lane <= 64: the `heavier` bit mask is one 64-bit word
constexpr int kCtaRing = 4;
constexpr int kCtaPad = 32; // zero-weight records behind the last body: the 2-way unrolled partner loop may overrun

struct CtaArgs {
    WorldPtrs w;
    RecordPtrs r;
    int n, steps, L, per_lane, phys_threads, phys_warps, rec_warps;
    double hm, dtm;
    double eps2;
    int record, argmax, offset_trails, forward_sim, acc_valid;
};

struct CtaSnapshot {
    double x[kCtaMaxBodies], y[kCtaMaxBodies], z[kCtaMaxBodies];
    double vx[kCtaMaxBodies], vy[kCtaMaxBodies], vz[kCtaMaxBodies];
    int most[kCtaMaxBodies], old_most[kCtaMaxBodies];
};

struct CtaShared {
    // SoA on purpose: the 8 lanes of a body read 8 consecutive partners -- conflict-free as 8-byte words,
    // a 2-way bank conflict as 32-byte records (measured: 4,560 vs 3,950 cycles per step)
    double sx[kCtaMaxBodies + kCtaPad], sy[kCtaMaxBodies + kCtaPad], sz[kCtaMaxBodies + kCtaPad], sg[kCtaMaxBodies + kCtaPad];
    unsigned long long full[kCtaRing], empty[kCtaRing];
    CtaSnapshot ring[kCtaRing];
};

__device__ __forceinline__ void bar_phys(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

__device__ __forceinline__ void mbar_arrive_cta(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// gf_p / r^4 to full precision (the reference's |ta|^2 / gf_p), for the winner only.
__device__ __forceinline__ double exact_metric(double xk, double yk, double zk, double xp, double yp, double zp, double gp) {
    double dx = xp - xk, dy = yp - yk, dz = zp - zk;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double y0 = rsqrt_seed(r2);
    double t = r2 * y0;
    double e = fma(-t, y0, 1.0);
    double y = fma(y0 * e, fma(0.375, e, 0.5), y0);
    double y2 = y * y;
    return gp * y2 * y2;
}

// `heavier` holds, for this lane's partner slots u = 0, 1, ..., one bit each: gf of partner q + u L
// exceeds gf of body k.  The partner set of a lane never changes, so the test is done once per
// kernel instead of with a DSETP per interaction.
template<bool ARGMAX>
__device__ __forceinline__ void group_forces(CtaShared const& S, int k, int q, int L, int per_lane, unsigned long long heavier,
    double eps2, double& ax, double& ay, double& az, double& maxattr, int& most) {
    double const x = S.sx[k], y = S.sy[k], z = S.sz[k];
    double fx[2] = { 0, 0 }, fy[2] = { 0, 0 }, fz[2] = { 0, 0 }, bm[2] = { 0, 0 };
    int bj[2] = { -1, -1 };
    for (int u0 = 0; u0 < per_lane; u0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int p = q + (u0 + u) * L; // < n + 2 L: the array is padded with zero-weight records
            double4 qv = make_double4(S.sx[p], S.sy[p], S.sz[p], S.sg[p]);
            double dx = qv.x - x, dy = qv.y - y, dz = qv.z - z;
            double r2 = fma(dx, dx, eps2);
            r2 = fma(dy, dy, r2);
            r2 = fma(dz, dz, r2);
            double y0 = rsqrt_seed(r2);
            double y2 = y0 * y0; // exact: the seed has 21 significant bits
            double e = fma(-r2, y2, 1.0);
            double gy = qv.w * y0;
            double c0 = gy * y2;
            double w = fma(1.875, e, 1.5);
            double we = w * e;
            double s = fma(c0, we, c0);
            fx[u] = fma(s, dx, fx[u]);
            fy[u] = fma(s, dy, fy[u]);
            fz[u] = fma(s, dz, fz[u]);
            if (ARGMAX) {
                double m = c0 * y0; // gf_p y0^4 ~ gf_p / r^4 (ranking only)
                if (((heavier >> (u0 + u)) & 1ull) && m > bm[u]) {
                    bm[u] = m;
                    bj[u] = p;
                }
            }
        }
    }
    double px = fx[0] + fx[1], py = fy[0] + fy[1], pz = fz[0] + fz[1];
    double pm = bm[0];
    int pj = bj[0];
    if (ARGMAX && (bm[1] > pm || (bm[1] == pm && bj[1] >= 0 && pj >= 0 && bj[1] < pj))) {
        pm = bm[1];
        pj = bj[1];
    }
    // shuffle tree over the L-lane group (L consecutive lanes of one warp); the leader (q == 0) ends with the total
    for (int off = L >> 1; off > 0; off >>= 1) {
        px += __shfl_down_sync(0xffffffffu, px, off);
        py += __shfl_down_sync(0xffffffffu, py, off);
        pz += __shfl_down_sync(0xffffffffu, pz, off);
        if (ARGMAX) {
            double m = __shfl_down_sync(0xffffffffu, pm, off);
            int j = __shfl_down_sync(0xffffffffu, pj, off);
            if (m > pm || (m == pm && j >= 0 && pj >= 0 && j < pj)) {
                pm = m;
                pj = j;
            }
        }
    }
    ax = px;
    ay = py;
    az = pz;
    if (ARGMAX && q == 0) {
        if (pj >= 0) {
            most = pj;
            maxattr = exact_metric(x, y, z, S.sx[pj], S.sy[pj], S.sz[pj], S.sg[pj]);
        } else {
            maxattr = 0.0;
        }
    }
}

template<bool ARGMAX>
__global__ void __launch_bounds__(768, 1) k_resident_cta(CtaArgs const A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared& S = *reinterpret_cast<CtaShared*>(smem_raw);
    int const n = A.n, L = A.L;
    int const tid = threadIdx.x, lane = tid & 31;
    bool const phys = tid < A.phys_threads;
    bool const rec_on = A.record != 0;

    if (tid == 0) {
        for (int i = 0; i < kCtaRing; i++) {
            mbar_init(&S.full[i], A.phys_warps);
            mbar_init(&S.empty[i], 2 * A.rec_warps);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < kCtaMaxBodies + kCtaPad; i += blockDim.x) {
        bool in = i < n;
        double4 p = in ? A.w.pos_cur[i] : make_double4(0, 0, 0, 0);
        S.sx[i] = p.x;
        S.sy[i] = p.y;
        S.sz[i] = p.z;
        S.sg[i] = p.w; // gf_eff: masked bodies weigh nothing; padding is {0, 0, 0, 0}
    }
    __syncthreads();

    if (!phys) {
        // ---------------- recorder warps: one lane per body, two groups ----------------
        // group 0: recalculate_trails_with_offset + Trail::push_back;  group 1: History + ap / pe
        if (!rec_on)
            return;
        int const r = tid - A.phys_threads;
        int const group = r / (32 * A.rec_warps);
        int const k = r - group * 32 * A.rec_warps;
        bool const mine = k < n && k < A.r.tracked;
        bool const active = mine && A.w.active[k] != 0;
        BodyRec b;
        if (mine)
            rec_load(A.r, k, b);
        for (int s = 0; s < A.steps; s++) {
            int const slot = s % kCtaRing;
            mbar_wait(&S.full[slot], (unsigned)(s / kCtaRing) & 1u);
            CtaSnapshot const& sn = S.ring[slot];
            if (active) {
                int const most = sn.most[k];
                int const mi = most >= 0 ? most : 0;
                double const mx = sn.x[mi], my = sn.y[mi], mz = sn.z[mi];
                if (group == 0) {
                    record_trail(A.r, k, b, sn.x[k], sn.y[k], sn.z[k], sn.vx[k], sn.vy[k], sn.vz[k], mx, my, mz, most, sn.old_most[k],
                        A.offset_trails != 0, A.forward_sim != 0, false);
                } else {
                    if (!A.forward_sim)
                        history_push(A.r, k, b, sn.x[k], sn.y[k], sn.z[k], sn.vx[k], sn.vy[k], sn.vz[k]);
                    record_appe(b, sn.x[k], sn.y[k], sn.z[k], sn.vx[k], sn.vy[k], sn.vz[k], mx, my, mz, most);
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive_cta(&S.empty[slot]);
        }
        if (mine) {
            if (group == 0) {
                rec_store_trail(A.r, k, b, false);
            } else {
                rec_store_appe(A.r, k, b);
                A.r.hist_start[k] = b.hist_start;
                A.r.hist_len[k] = b.hist_len;
            }
        }
        return;
    }

    // ---------------- physics warps ----------------
    int const k = tid / L, q = tid - k * L; // L consecutive lanes per body
    bool const live = k < n;
    bool const leader = live && q == 0;
    int const kc = live ? k : 0;
    double vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0, maxattr = 0;
    double x = 0, y = 0, z = 0;
    int most = -1, old_most = -1;
    bool active = false;
    if (live) {
        gfk = A.w.gf[k];
        active = A.w.active[k] != 0;
    }
    unsigned long long heavier = 0;
    if (live)
        for (int u = 0; u < A.per_lane && u < 64; u++)
            if (S.sg[q + u * L] > gfk)
                heavier |= 1ull << u;
    if (leader) {
        x = S.sx[k];
        y = S.sy[k];
        z = S.sz[k];
        vx = A.w.v[0][k];
        vy = A.w.v[1][k];
        vz = A.w.v[2][k];
        ax = A.w.a[0][k];
        ay = A.w.a[1][k];
        az = A.w.a[2][k];
        maxattr = A.w.maxattr[k];
        most = A.w.most[k];
        old_most = A.w.old_most[k];
    }
    double const hm = A.hm, dtm = A.dtm;

    auto forces = [&]() {
        double fx, fy, fz, fm = maxattr;
        int fmost = most;
        group_forces<ARGMAX>(S, kc, q, L, A.per_lane, heavier, A.eps2, fx, fy, fz, fm, fmost);
        if (leader && active) {
            ax = fx;
            ay = fy;
            az = fz;
            maxattr = fm;
            most = fmost;
        }
    };

    if (!A.acc_valid) { // the first set_forces() of the first step
        old_most = most;
        forces();
    }

    for (int s = 0; s < A.steps; s++) {
        if (s > 0 || A.acc_valid)
            old_most = most; // Object::before_update
        if (leader && active) { // first half kick + drift (src/World.cpp:112,115)
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
            x = __dadd_rn(x, __dmul_rn(vx, dtm));
            y = __dadd_rn(y, __dmul_rn(vy, dtm));
            z = __dadd_rn(z, __dmul_rn(vz, dtm));
        }
        bar_phys(A.phys_threads); // every lane has read the old positions
        if (leader) {
            S.sx[k] = x;
            S.sy[k] = y;
            S.sz[k] = z;
        }
        bar_phys(A.phys_threads);
        forces();
        if (leader && active) { // second half kick (src/World.cpp:127)
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
        }
        if (rec_on) {
            int const slot = s % kCtaRing;
            if (s >= kCtaRing)
                mbar_wait(&S.empty[slot], (unsigned)(s / kCtaRing - 1) & 1u);
            CtaSnapshot& sn = S.ring[slot];
            if (leader) {
                sn.x[k] = x;
                sn.y[k] = y;
                sn.z[k] = z;
                sn.vx[k] = vx;
                sn.vy[k] = vy;
                sn.vz[k] = vz;
                sn.most[k] = most;
                sn.old_most[k] = old_most;
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive_cta(&S.full[slot]);
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
        A.w.maxattr[k] = maxattr;
        A.w.most[k] = most;
        A.w.old_most[k] = old_most;
    }
}

bool resident_cta_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    return !forces_only && !o.exact_order && !o.two_pass && c->n > 32 && c->n <= kCtaMaxBodies
        && !(o.forward_simulated && c->closest && c->closest_n == c->n);
}

cudaError_t launch_resident_cta(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st) {
    CtaArgs A;
    A.w = world_ptrs(c);
    A.r = record_ptrs(c);
    A.n = c->n;
    A.steps = steps < 0 ? -steps : steps;
    double const mul = steps >= 0 ? 1.0 : -1.0;
    A.dtm = (double)o.dt_seconds * mul;
    A.hm = (double)o.dt_seconds / 2.0 * mul;
    A.eps2 = o.eps2;
    A.record = o.record && c->tracked > 0;
    A.argmax = o.track_most_attracting || o.record;
    A.offset_trails = o.offset_trails;
    A.forward_sim = o.forward_simulated;
    A.acc_valid = c->acc_valid && (!A.argmax || c->acc_has_most) && !c->acc_exact;
    int const rec_threads = 2 * ((c->n + 31) / 32 * 32); // two recorder groups
    int L = 8;
    while (L > 1 && (c->n * L + 31) / 32 * 32 + rec_threads > 768)
        L >>= 1;
    A.L = L;
    int per = (c->n + L - 1) / L;
    A.per_lane = (per + 1) & ~1; // the inner loop takes partners two at a time; the array is zero-padded
    A.phys_threads = (c->n * L + 31) / 32 * 32;
    A.phys_warps = A.phys_threads / 32;
    A.rec_warps = (c->n + 31) / 32;
    int threads = A.phys_threads + 2 * 32 * A.rec_warps;
    size_t smem = sizeof(CtaShared);
    bool& attr_set = c->attr_cta; // per context: the attribute is per device
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_resident_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_resident_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        attr_set = true;
    }
    if (A.argmax)
        k_resident_cta<true><<<1, threads, smem, st>>>(A);
    else
        k_resident_cta<false><<<1, threads, smem, st>>>(A);
    c->launches++;
    int passes = A.steps + (A.acc_valid ? 0 : 1);
    c->passes += passes;
    c->interactions += (double)passes * (double)c->n * (double)(c->n - 1);
    c->last_force_passes = passes;
    return cudaGetLastError();
}

} // namespace essa
