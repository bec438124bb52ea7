// K3x -- resident kernel for worlds of 33 .. 128 bodies (jupiter_system.essa: Jupiter + 76 moons) in fast
// arithmetic, spread over a THREAD-BLOCK CLUSTER of 16 SMs that exchange positions through distributed
// shared memory.  Same World::update loop (src/World.cpp:86-139) in one launch as K3 / K3c, but a step of a
// single trajectory is a dependent chain, and on one SM that chain is n (n - 1) x 16 FP64 instructions deep
// for the 64 FP64 lanes of the SM (K3c: 3,600 cycles per step at 77 bodies).  Here:
//
//   * every CTA of the cluster owns up to 8 bodies; ONE WARP PER BODY evaluates that body's partners
//     (lane l takes p = l, l + 32, ...), a five-level shuffle butterfly leaves the same total in every lane, and
//     every lane kicks and drifts the replicated state, so nothing inside a warp waits for a leader;
//   * the new position (one 32-byte record {x, y, z, gf}) goes out with st.async: lane j stores it into CTA j's
//     shared memory as two 16-byte halves, each store completing the destination's per-step mbarrier by its
//     byte count; a warp starts the next force pass when ITS CTA's barrier has seen all n records.  No block
//     barrier, no cluster barrier in the loop.  (One cp.async.bulk per peer for the CTA's whole block -- 15
//     instead of 160 remote transactions per CTA and step -- was measured SLOWER, 2,311 vs 1,716 cycles per
//     step: the exchange is latency bound and the bulk-copy engine adds latency.)
//     Positions live in a ring of 8 step slots per CTA;
//   * the most-attracting link (src/Object.cpp:84-94) rides in the physics lanes at one extra FP64 multiply per
//     interaction -- candidates are ranked by gf_p y0^4 with y0 the rsqrt seed (relative error < 4e-6: enough
//     unless two heavier bodies attract within 4 ppm of each other), the warp-wide winner comes from two integer
//     warp reductions over sortable keys, and max_attraction is formed to full precision for the final state only.  (Separate attraction
//     warps, as in K3q, cost the physics warps of the same SM more in FP64-pipe contention than the inline
//     multiply costs in latency: measured.)
//   * the recording (record.cuh) runs on four consumer warps per CTA for the CTA's own bodies, off the critical
//     path, reading the same ring: Trail even / odd bodies, link bookkeeping + History, ap / pe, as in K3q.
//     Consumers publish their progress to every CTA; a physics warp looks at it only when its cached copy says
//     the slot it is about to overwrite could still be in use.
//
// Worlds whose active mask changes during the call, two_pass, exact_order and closest-approach tracking go
// through resident.cu; if the cluster cannot be scheduled the single-SM K3c takes the call.
#include <cstdlib>

#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"
#include "resident_small.cuh"

namespace essa {

constexpr int kXMaxBodies = 128;
constexpr int kXLocalMax = 8; // bodies per CTA
constexpr int kXRing = 8;     // step slots
constexpr int kXRoles = 4;    // consumer warps per CTA
constexpr int kXCtas = 16;    // cluster size (non-portable: needs the opt-in attribute)
constexpr int kXThreads = 32 * (kXLocalMax + kXRoles + 1); // + the poller warp of the persistent flavour

struct XArgs {
    WorldPtrs w;
    RecordPtrs r;
    int n, steps, nbl, ctas; // nbl bodies per CTA, `ctas` CTAs own at least one body
    double hm, dtm, eps2;
    int record, argmax, offset_trails, forward_sim, acc_valid;
};

struct XShared {
    double4 pos[kXRing][kXMaxBodies];  // {x, y, z, gf_eff} of every body, by step slot
    double vel[kXRing][3][kXLocalMax]; // this CTA's bodies: velocity after the step's second kick
    int most[kXRing][kXLocalMax];      // this CTA's bodies: the step's most-attracting winner (-1: none)
    unsigned long long full[kXRing];   // all n records of the slot have arrived (transaction bytes)
    unsigned long long vready[kXRing]; // local velocities / winners written
    unsigned done[kXRoles * kXCtas];   // steps consumed by role r of CTA c, at [r * kXCtas + c]
    unsigned long long grant;          // persistent flavour: (tick number << 32) | steps granted so far, written by the poller
    unsigned long long tick_t0;        // ... globaltimer at the doorbell
    unsigned quit;                     // ... the order to retire
};

__device__ __forceinline__ void x_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// version v (positions after step v - 1; version 0 = the state at launch) lives in slot v % kXRing; slot 0's
// barrier is first used by version kXRing
__device__ __forceinline__ unsigned x_parity(int v) { return (unsigned)((v / kXRing) - ((v % kXRing) == 0 ? 1 : 0)) & 1u; }

// gf_p / r^4 of partner p as seen from body k (the reference's |ta|^2 / gf_p), to ~1e-15
__device__ __forceinline__ double x_metric(double4 const& pk, double4 const& pp, double gp) {
    double dx = pp.x - pk.x, dy = pp.y - pk.y, dz = pp.z - pk.z;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double y0 = rsqrt_seed(r2);
    double t = r2 * y0;
    double e = fma(-t, y0, 1.0);
    double y = fma(y0 * e, fma(0.375, e, 0.5), y0); // 1 / r
    double y2 = y * y;
    return gp * y2 * y2;
}

// persistent flavour: wait until steps beyond `s` are granted (the doorbell, forwarded by CTA 0's poller); false: retire
__device__ __forceinline__ bool x_more_steps(unsigned long long const* grant, unsigned const* quit, int s, int& end, unsigned& seq) {
    for (;;) {
        unsigned long long g;
        asm volatile("ld.acquire.cluster.shared::cta.u64 %0, [%1];" : "=l"(g) : "r"(smem_u32(grant)) : "memory");
        end = (int)(unsigned)g;
        seq = (unsigned)(g >> 32);
        if (s < end)
            return true;
        if (ld_acquire_cluster_u32(quit) != 0)
            return false;
        __nanosleep(32);
    }
}

// PER = partners per lane = ceil(n / 32): a compile-time bound keeps the partner loop one straight-line block, so
// that the PER dependent chains interleave (with a run-time bound each partner was its own basic block: 760
// instead of 440 cycles for three partners, measured)
// PERSIST (essa_persist_configure): the kernel outlives the call.  An extra warp of CTA 0 polls the host's command word in
// mapped pinned memory and forwards every doorbell to the 16 CTAs; at a tick's end every physics warp writes its body's
// x v a, and the link / ap-pe consumer warps their results, straight to the host as self-validating words {tick number,
// 32 bits of payload} (resident_quad.cu has the protocol): the host has the tick when every word carries its number.
template<int PER, bool ARGMAX, bool PERSIST>
__global__ void __launch_bounds__(kXThreads, 1) k_resident_cluster(XArgs const A, PersistArgs const P) {
    __shared__ __align__(16) XShared S;
    int const n = A.n, nbl = A.nbl;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const cta = (int)cluster_ctarank();
    int const first = cta * nbl;
    int const nl = max(0, min(nbl, n - first)); // bodies this CTA owns
    bool const rec_on = A.record != 0, snap_on = rec_on || ARGMAX;
    unsigned const pos_bytes = (unsigned)n * 32u; // what arrives per step

    if (threadIdx.x == 0) {
        for (int i = 0; i < kXRing; i++) {
            mbar_init(&S.full[i], 1);
            mbar_init(&S.vready[i], nl > 0 ? nl : 1);
        }
        mbar_fence_init();
        for (int i = 0; i < kXRing; i++)
            mbar_expect_tx(&S.full[i], pos_bytes);
    }
    for (int i = threadIdx.x; i < kXRoles * kXCtas; i += blockDim.x)
        S.done[i] = 0;
    if (threadIdx.x == 0) {
        S.grant = ((unsigned long long)P.seq0 << 32) | (unsigned)A.steps;
        S.quit = 0;
        S.tick_t0 = PERSIST ? globaltimer_ns() : 0ull;
    }
    for (int i = threadIdx.x; i < kXRing * kXMaxBodies; i += blockDim.x) {
        // slot 0 = the state at launch; everywhere the padding behind body n - 1 is read with weight 0
        int const b = i % kXMaxBodies;
        (&S.pos[0][0])[i] = (i < kXMaxBodies && b < n) ? A.w.pos_cur[b] : make_double4(0, 0, 0, 0);
    }
    __syncthreads();
    cluster_sync_all();

    if (warp < nbl) {
        // ---------------- physics: one warp per body ----------------
        int const kl = warp, k = first + kl;
        if (kl < nl) {
            double4 const p0 = S.pos[0][k];
            double x = p0.x, y = p0.y, z = p0.z;
            double vx = A.w.v[0][k], vy = A.w.v[1][k], vz = A.w.v[2][k];
            double ax = A.w.a[0][k], ay = A.w.a[1][k], az = A.w.a[2][k];
            double const gfk = A.w.gf[k];
            bool const active = A.w.active[k] != 0;
            double const gfe = active ? gfk : 0.0;
            double const hm = A.hm, dtm = A.dtm;
            bool heavier[PER]; // the partner set of a lane never changes: one test per kernel, not per interaction
#pragma unroll
            for (int u = 0; u < PER; u++) {
                int const p = lane + 32 * u;
                heavier[u] = p < n && p != k && S.pos[0][p].w > gfk; // masked partners weigh 0: never heavier
            }
            int winner = -1; // the last force pass's most-attracting candidate

            auto forces = [&](int slot) {
                double4 const* pp = S.pos[slot];
                double fx[PER], fy[PER], fz[PER], m[PER];
#pragma unroll
                for (int u = 0; u < PER; u++) {
                    double4 const q = pp[lane + 32 * u];
                    double dx = q.x - x, dy = q.y - y, dz = q.z - z;
                    double r2 = fma(dx, dx, A.eps2);
                    r2 = fma(dy, dy, r2);
                    r2 = fma(dz, dz, r2);
                    double y0 = rsqrt_seed(r2);
                    double y2 = y0 * y0; // exact: the seed has 21 significant bits
                    double e = fma(-r2, y2, 1.0);
                    double gy = q.w * y0;
                    double c0 = gy * y2;
                    double w = fma(1.875, e, 1.5);
                    double we = w * e;
                    double s = fma(c0, we, c0);
                    fx[u] = s * dx;
                    fy[u] = s * dy;
                    fz[u] = s * dz;
                    if (ARGMAX)
                        m[u] = heavier[u] ? c0 * y0 : 0.0; // gf_p y0^4 ~ gf_p / r^4 (ranking only)
                }
                double sx = fx[0], sy = fy[0], sz = fz[0];
#pragma unroll
                for (int u = 1; u < PER; u++) {
                    sx += fx[u];
                    sy += fy[u];
                    sz += fz[u];
                }
                // Candidates as sortable 64-bit keys: the metric's bit pattern (positive doubles order like unsigned
                // integers) with its 7 lowest mantissa bits replaced by 127 - p, so that the largest key is the largest
                // metric and, among equal ones, the lowest index -- then two integer warp reductions instead of a
                // five-level compare-and-select butterfly on the FP64 pipe.
                unsigned long long key = 0;
                if (ARGMAX) {
#pragma unroll
                    for (int u = 0; u < PER; u++) {
                        unsigned long long ku = ((unsigned long long)__double_as_longlong(m[u]) & ~127ull) | (unsigned)(127 - (lane + 32 * u));
                        ku = m[u] > 0.0 ? ku : 0ull;
                        key = ku > key ? ku : key;
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { // a + b == b + a: every lane ends with the same bits
                    sx += __shfl_xor_sync(0xffffffffu, sx, off);
                    sy += __shfl_xor_sync(0xffffffffu, sy, off);
                    sz += __shfl_xor_sync(0xffffffffu, sz, off);
                }
                int bp = -1;
                if (ARGMAX) {
                    unsigned const hi = (unsigned)(key >> 32), lo = (unsigned)key;
                    unsigned const mh = __reduce_max_sync(0xffffffffu, hi);
                    unsigned const ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                    bp = mh != 0u ? 127 - (int)(ml & 127u) : -1;
                }
                if (active) {
                    ax = sx;
                    ay = sy;
                    az = sz;
                    winner = bp;
                }
            };

            if (!A.acc_valid)
                forces(0); // the first set_forces() of the first step

            // lane j < ctas sends this body's record to CTA j (its own CTA included)
            bool const sender = lane < A.ctas;
            uint32_t const rec_remote = mapa_u32(&S.pos[0][k], sender ? lane : 0);
            uint32_t const full_remote = mapa_u32(&S.full[0], sender ? lane : 0);
            int consumed = 0; // cached lower bound of the versions every consumer of the cluster has finished with

            int s = 0, end = A.steps;
            unsigned seq = P.seq0;
            for (;;) {
            for (; s < end; s++) {
                int const v = s + 1, slot = v % kXRing;
                if (active) { // first half kick + drift (src/World.cpp:112,115)
                    vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                    vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                    vz = __dadd_rn(vz, __dmul_rn(az, hm));
                    x = __dadd_rn(x, __dmul_rn(vx, dtm));
                    y = __dadd_rn(y, __dmul_rn(vy, dtm));
                    z = __dadd_rn(z, __dmul_rn(vz, dtm));
                }
                if (snap_on && v - consumed > kXRing) { // the slot holds version v - kXRing: is everybody done with it?
                    do {
                        unsigned m = 0xffffffffu;
                        for (int i = lane; i < kXRoles * kXCtas; i += 32) {
                            int const role = i / kXCtas, c = i % kXCtas;
                            if (c < A.ctas && (rec_on || role == 1))
                                m = min(m, ld_acquire_cluster_u32(&S.done[i]));
                        }
                        consumed = (int)__reduce_min_sync(0xffffffffu, m);
                    } while (v - consumed > kXRing);
                }
                if (sender) {
                    uint32_t const dst = rec_remote + (uint32_t)slot * (uint32_t)(kXMaxBodies * sizeof(double4));
                    uint32_t const bar = full_remote + 8u * (uint32_t)slot;
                    st_async_v2_f64(dst, x, y, bar);
                    st_async_v2_f64(dst + 16u, z, gfe, bar);
                }
                mbar_wait(&S.full[slot], x_parity(v));
                if (kl == 0 && lane == 0)
                    mbar_expect_tx(&S.full[slot], pos_bytes); // armed for the slot's next version
                forces(slot);
                if (active) { // second half kick (src/World.cpp:127)
                    vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                    vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                    vz = __dadd_rn(vz, __dmul_rn(az, hm));
                }
                if (snap_on && lane == 0) {
                    if (rec_on) {
                        S.vel[slot][0][kl] = vx;
                        S.vel[slot][1][kl] = vy;
                        S.vel[slot][2][kl] = vz;
                    }
                    S.most[slot][kl] = active ? winner : -1;
                    x_arrive(&S.vready[slot]);
                }
            }
            if (!PERSIST)
                break;
            // ---- end of a tick: this body's x v a to the host (words 0 .. 17 of its record, ctx.hpp: one coalesced store), then the next doorbell ----
            {
                unsigned long long const tag = (unsigned long long)seq << 32;
                int const f = lane >> 1;
                double const d = f == 0 ? x : f == 1 ? y : f == 2 ? z : f == 3 ? vx : f == 4 ? vy : f == 5 ? vz : f == 6 ? ax : f == 7 ? ay : az;
                unsigned const w = (lane & 1) ? (unsigned)__double2hiint(d) : (unsigned)__double2loint(d);
                if (lane < 18)
                    st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * k + lane, tag | w);
            }
            if (!x_more_steps(&S.grant, &S.quit, s, end, seq))
                break; // retire
            }
            if (lane == 0) {
                A.w.pos_cur[k] = make_double4(x, y, z, gfe);
                A.w.v[0][k] = vx;
                A.w.v[1][k] = vy;
                A.w.v[2][k] = vz;
                A.w.a[0][k] = ax;
                A.w.a[1][k] = ay;
                A.w.a[2][k] = az;
            }
        }
    } else if (warp < nbl + kXRoles) {
        // ---------------- consumers: 0, 2 = Trail (even / odd local bodies), 1 = links + History, 3 = ap / pe ----------------
        int const role = warp - nbl;
        if (snap_on && nl > 0 && (role == 1 || rec_on)) {
            bool const is_trail = role == 0 || role == 2;
            int const k = first + lane;
            bool const body = lane < nl;
            bool const mine = rec_on && body && k < A.r.tracked && (!is_trail || (lane & 1) == (role == 2 ? 1 : 0));
            bool const act = body && A.w.active[k] != 0; // the mask does not change during the call
            bool const offset_trails = A.offset_trails != 0, forward_sim = A.forward_sim != 0;
            uint32_t const done_remote = mapa_u32(&S.done[role * kXCtas + cta], lane < A.ctas ? lane : 0);
            BodyRec b;
            if (mine)
                rec_load(A.r, k, b);
            int most = body ? A.w.most[k] : -1, old_most = body ? A.w.old_most[k] : -1;
            double maxattr = (role == 1 && body) ? A.w.maxattr[k] : 0.0;
            int end = A.steps;
            unsigned seq = P.seq0;
            for (int v = 1;; v++) {
                if (v > end) {
                    if (!PERSIST || !x_more_steps(&S.grant, &S.quit, v - 1, end, seq))
                        break;
                }
                bool const last = v == end; // (the host extends `end` only after it has this tick's result)
                int const slot = v % kXRing;
                unsigned const par = x_parity(v);
                mbar_wait(&S.full[slot], par);
                mbar_wait(&S.vready[slot], par);
                double4 const* pp = S.pos[slot];
                int const winner = body ? S.most[slot][lane] : -1;
                old_most = most; // Object::before_update (every object)
                if (act && winner >= 0)
                    most = winner;
                if (mine && act) {
                    double4 const pm = pp[most >= 0 ? most : 0], pk = pp[k];
                    double const vx = S.vel[slot][0][lane], vy = S.vel[slot][1][lane], vz = S.vel[slot][2][lane];
                    if (is_trail)
                        record_trail(A.r, k, b, pk.x, pk.y, pk.z, vx, vy, vz, pm.x, pm.y, pm.z, most, old_most, offset_trails, forward_sim,
                            false);
                    else if (role == 1) {
                        if (!forward_sim)
                            history_push(A.r, k, b, pk.x, pk.y, pk.z, vx, vy, vz);
                    } else
                        record_appe(b, pk.x, pk.y, pk.z, vx, vy, vz, pm.x, pm.y, pm.z, most);
                }
                if (role == 1 && last && act) {
                    // max_attraction as the last force pass left it: the winner's gf / r^4 to full precision, or 0
                    maxattr = winner >= 0 ? x_metric(pp[k], pp[winner], A.w.gf[winner]) : 0.0;
                }
                if (PERSIST && last && body) { // this warp's share of what the host reads back (fields 9 .. 13, then the links)
                    unsigned long long const tag = (unsigned long long)seq << 32;
                    auto put = [&](int f, double d) { // field f of body k: words 2 f, 2 f + 1 of its record
                        st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * k + 2 * f, tag | (unsigned)__double2loint(d));
                        st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * k + 2 * f + 1, tag | (unsigned)__double2hiint(d));
                    };
                    if (role == 1) {
                        put(9, maxattr);
                        st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * k + 28, tag | (unsigned)most);
                        if (k == 0) {
                            unsigned long long const dtn = globaltimer_ns() - S.tick_t0;
                            st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * n, tag | (unsigned)dtn);
                            st_relaxed_sys_u64(P.out + (size_t)kPersistRecWords * n + 1, tag | (unsigned)(dtn >> 32));
                        }
                    } else if (role == 3) {
                        if (mine) {
                            ap_flush(b);
                            pe_flush(b);
                        }
                        put(10, mine ? b.ap : A.r.ap[k]);
                        put(11, mine ? b.pe : A.r.pe[k]);
                        put(12, mine ? b.apv : A.r.ap_vel[k]);
                        put(13, mine ? b.pev : A.r.pe_vel[k]);
                    }
                }
                __syncwarp();
                if (lane < A.ctas)
                    st_relaxed_cluster_u32(done_remote, (unsigned)v);
            }
            if (role == 1 && body) {
                A.w.most[k] = most;
                A.w.old_most[k] = old_most;
                A.w.maxattr[k] = maxattr;
            }
            if (mine) {
                if (is_trail) {
                    rec_store_trail(A.r, k, b, false);
                } else if (role == 1) {
                    A.r.hist_start[k] = b.hist_start;
                    A.r.hist_len[k] = b.hist_len;
                } else {
                    rec_store_appe(A.r, k, b);
                }
            }
        }
    }
    else if (PERSIST && cta == 0 && warp == nbl + kXRoles) {
        // ---------------- poller: the host's doorbell (one lane reads the command word over PCIe, one read in flight) ----------------
        unsigned seq = P.seq0;
        unsigned end = (unsigned)A.steps;
        uint32_t const grant_remote = mapa_u32(&S.grant, lane < kXCtas ? lane : 0), quit_remote = mapa_u32(&S.quit, lane < kXCtas ? lane : 0),
                       t0_remote = mapa_u32(&S.tick_t0, 0);
        for (;;) {
            unsigned long long c = 0ull;
            bool idle = false;
            if (lane == 0) {
                long long const i0 = clock64(); // (a tick longer than the idle time retires the kernel after it: harmless)
                for (;;) {
                    c = ld_relaxed_sys_u64(P.cmd);
                    if ((unsigned)(c >> 32) != seq)
                        break;
                    if (clock64() - i0 > P.idle_clk) {
                        idle = true;
                        break;
                    }
                }
            }
            c = __shfl_sync(0xffffffffu, c, 0);
            idle = __shfl_sync(0xffffffffu, idle ? 1 : 0, 0) != 0;
            unsigned const add = idle ? 0u : (unsigned)c; // steps of the new tick; 0: retire
            if (add == 0) {
                if (lane == 0)
                    st_relaxed_sys_u32(P.ack + 1, idle ? 2u : 1u);
                if (lane < kXCtas)
                    st_relaxed_cluster_u32(quit_remote, 1u);
                break;
            }
            end += add;
            seq = (unsigned)(c >> 32);
            if (lane == 0) {
                unsigned long long const now = globaltimer_ns();
                asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(t0_remote), "l"(now) : "memory");
            }
            __syncwarp();
            if (lane < kXCtas) {
                unsigned long long const g = ((unsigned long long)seq << 32) | end;
                asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(grant_remote), "l"(g) : "memory");
            }
        }
    }
    // nobody leaves while a peer may still store into (or signal) its shared memory
    cluster_sync_all();
}

bool resident_cluster_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    static int const enabled = [] {
        char const* e = getenv("ESSA_RESIDENT_CLUSTER"); // diagnostics: 0 keeps such worlds on the single-SM kernel
        return e ? atoi(e) : 1;
    }();
    return enabled && !forces_only && !o.exact_order && !o.two_pass && c->n > 32 && c->n <= kXMaxBodies
        && !(o.forward_simulated && c->closest && c->closest_n == c->n);
}

template<int PER>
static cudaError_t launch_x(essa_ctx* c, cudaLaunchConfig_t const& cfg, XArgs const& A, PersistArgs const* pa) {
    if (!c->attr_cluster16[PER]) {
        cudaError_t e = cudaFuncSetAttribute(k_resident_cluster<PER, true, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_resident_cluster<PER, false, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_resident_cluster<PER, true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess)
            return e;
        c->attr_cluster16[PER] = true;
    }
    if (pa)
        return cudaLaunchKernelEx(&cfg, k_resident_cluster<PER, true, true>, A, *pa);
    PersistArgs const none {};
    return A.argmax ? cudaLaunchKernelEx(&cfg, k_resident_cluster<PER, true, false>, A, none)
                    : cudaLaunchKernelEx(&cfg, k_resident_cluster<PER, false, false>, A, none);
}

// Worlds the persistent flavour of K3x serves: recording on for every body (what the World facade runs).
bool resident_cluster_persist_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    return resident_cluster_applicable(c, o, forces_only) && o.record && c->tracked >= c->n && c->world == 1;
}

cudaError_t launch_resident_cluster(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st, bool persist) {
    PersistArgs PA {};
    PersistArgs const* pa = nullptr;
    if (persist) {
        PersistMailbox* mb = c->persist.mailbox;
        PA.cmd = &mb->cmd;
        PA.ack = mb->ack;
        PA.out = mb->ll;
        PA.seq0 = c->persist.seq;
        PA.idle_clk = (long long)c->persist.idle_us * 2000; // ~ SM cycles (1.965 GHz at full clock)
        pa = &PA;
    }
    XArgs A;
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
    A.nbl = (c->n + kXCtas - 1) / kXCtas;
    A.ctas = (c->n + A.nbl - 1) / A.nbl;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kXCtas, 1, 1);
    cfg.blockDim = dim3(32 * (A.nbl + kXRoles + 1), 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kXCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int const per = (c->n + 31) / 32; // 2 .. 4
    cudaError_t e = per == 2 ? launch_x<2>(c, cfg, A, pa) : per == 3 ? launch_x<3>(c, cfg, A, pa) : launch_x<4>(c, cfg, A, pa);
    if (e != cudaSuccess)
        return e;
    c->launches++;
    int passes = A.steps + (A.acc_valid ? 0 : 1);
    c->passes += passes;
    c->interactions += (double)passes * (double)c->n * (double)(c->n - 1);
    c->last_force_passes = passes;
    return cudaGetLastError();
}

} // namespace essa
