// K3q -- resident kernel for worlds of up to 32 bodies in fast arithmetic (the Solar System template has 10): the
// whole World::update loop (src/World.cpp:86-139) on a CLUSTER OF TWO CTAs (two SMs) joined by distributed shared
// memory, optionally PERSISTENT across World::update calls.
//
// A single trajectory is a dependent chain (each step needs the previous one), bound by dependent-issue latency
// (measured on B200: DFMA/DADD/DMUL 8.3 cycles, MUFU.RSQ64H seed 23, LDS 29, shuffle of a double 30), and every
// warp-wide FP64 instruction occupies its SM sub-partition's pipe for two cycles, so whatever is not the chain
// itself runs on the other SM:
//
//   CTA 0  warps 0-3  physics, one warp per SM sub-partition.  Every warp carries the whole replicated state
//                     (identical bits in all four) but evaluates only every fourth partner slot; the four partial
//                     accelerations meet in shared memory behind ONE named barrier per force pass and are added in
//                     a fixed order by every lane, so the replicas stay identical and no position is ever
//                     broadcast.  After a step each warp stores two components of the state into a LOCAL staging
//                     buffer (two predicated STS + one mbarrier arrive; remote stores from these warps cost more).
//          warp 4     courier: forwards a staged step to CTA 1's ring with st.async -- the store itself completes
//                     the ring slot's mbarrier there (transaction bytes) -- and, at a tick's end, the final x v a.
//          warp 5     poller (persistent flavour): reads the host's command word, grants ticks to both CTAs.
//   CTA 1  warps 0-2  attraction: the most-attracting candidate of every body of ONE step (src/Object.cpp:84-94): a lane
//                     is a partner, four bodies in flight, the winner from two full-warp redux.sync over sortable keys
//                     (gf_p / r^4 with the refined reciprocal, the 5 lowest mantissa bits replaced by 31 - p).  Ranking
//                     inside the physics warps was measured at +170 cycles per step; dealing whole steps round-robin to
//                     these warps (round 1) at 800 cycles of latency that every World::update call paid at its end.
//          warps 3-6  Trail (bodies by index mod 4: an append stalls the whole warp)
//          warp 7     link bookkeeping + History
//          warps 8, 9 apoapsis / periapsis (the two halves of src/Object.cpp:111-123 are independent chains)
//          warp 10    publisher (persistent flavour)
//                     The consumers report progress with a relaxed store into CTA 0's shared memory; the courier looks
//                     at it only when its cached copy says the ring could be full.
//
// Every warp has ONE role and its own copy of the step loop (generic lambdas over std::integral_constant): in these
// single-warp dependent chains a taken branch costs ~20 cycles and a divergent `if (lane == 0) asm...` region ~90.
//
// Persistent flavour (essa_persist_configure, capi.cu): the poller reads the host's command word in mapped pinned
// memory (ONE read in flight, see ESSA_POLL_DEPTH), forwards each grant to CTA 1 and releases the physics warps from
// a named barrier.  At a tick's end the courier sends the final state to CTA 1; the publisher waits for it and for the
// links / ap / pe warps' last step, then writes the packed state to the host as self-validating 64-bit words
// {tick number, 32 bits of payload} (what NCCL's LL protocol does over PCIe): no system-wide fence, no separate
// acknowledgement -- the host has a tick's result when every word carries the tick's number.  Trail and History may
// lag by up to the ring's depth; their rings are only read after the kernel has retired.
#include <type_traits>

#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"
#include "resident_small.cuh"

namespace essa {

#ifndef ESSA_POLL_DEPTH
#define ESSA_POLL_DEPTH 1 // reads of the host's command word in flight.  Measured on B200 (PCIe Gen5 host): ONE read at a time sees
#endif                    // a doorbell ~1.2 us after it was rung; four staggered reads of the same word took ~5 us (they do not
                          // pipeline: each waits for the one before it at the host bridge), so the deeper variant is a dead end kept
                          // for the record
// How a tick's result reaches the host: self-validating 64-bit words (below).  Measured alternative: plain stores of the packed
// doubles + __threadfence_system() + an acknowledgement word cost 2 us more per tick (the system-scope fence).

constexpr int kQuadWarps = 4;
// CTA 0: warps 0-3 physics, 4 courier, 5 poller.  CTA 1: warps 0-2 attraction, 3-6 Trail, 7 links + History, 8 apoapsis,
// 9 periapsis, 10 publisher.
constexpr int kCourierWarp = 4, kPollWarp = 5;
constexpr int kAttrWarps = 3, kTrailWarp0 = 3, kTrailWarps = 4, kHistWarp = 7, kApWarp = 8, kPeWarp = 9, kPubWarp = 10;
constexpr int kConsumers = kTrailWarps + 3;
constexpr int kQuadThreads = 32 * 11;
constexpr int kStage = 4; // physics -> courier staging buffers

struct QSnap {
    double x[32], y[32], z[32], vx[32], vy[32], vz[32];
    int win[32];             // the step's most-attracting winner (-1: none, or body masked)
    unsigned long long mask; // active bodies
};

__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000) // parked for up to ~2 us per try
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire_cta_u32(unsigned const* p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// The same on 32-bit shared-window addresses formed ONCE outside the step loop: a generic pointer into the shared memory of a
// cluster kernel is converted with a read of SR_CgaCtaId (tens of cycles) wherever it is used, and the step loop of the physics
// warps is one dependent chain.
// (... and the compiler re-materialises such an address inside the loop instead of keeping it in a register unless it cannot
// see where the value comes from)
__device__ __forceinline__ uint32_t keep_in_register(uint32_t a) {
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared::cta.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, unsigned v) { asm volatile("st.shared::cta.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned lds_acquire_u32(uint32_t a) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }

struct Q0 { // CTA 0
    double spos[kQuadWarps][4][32];            // per physics warp sx, sy, sz, sg (private copies)
    double part[2][kQuadWarps][3][32];         // partial accelerations, double-buffered
    double stage[kStage][9][32];               // x y z vx vy vz ax ay az of a step, for the courier
    unsigned long long stagebar[kStage];       // physics warps 1-3 have written a staging buffer
    unsigned long long cmdt0;                  // globaltimer at the doorbell
    unsigned stage_mask[kStage];               // active bodies of that step
    unsigned courier_done;                     // steps the courier has forwarded
    unsigned fin_sent;                         // ticks whose final bundle has left
    unsigned done[kConsumers];                 // steps consumed by CTA 1's warps 3 .. 9
    unsigned cmdw[2];                          // poller -> physics / courier: steps granted so far, tick number
};
struct Q1 { // CTA 1
    QSnap ring[kRing];
    unsigned long long full[kRing], ready[kRing]; // snapshot arrived (transaction bytes) / winners known
    unsigned long long finbar, outbar;         // bundle arrived / consumers finished the last step
    unsigned long long fin_meta[2];            // globaltimer at the doorbell, tick number
    double sgf[32];                            // gravity factors
    double fin[9][32];                         // x y z vx vy vz ax ay az as the tick left them
    double outd[5][32];                        // max_attraction, ap, pe, ap_vel, pe_vel at the tick's end
    int outmost[32];                           // most-attracting links at the tick's end
    unsigned ctl[2];                           // steps granted so far, retire flag
};

template<int T, int LOG2L, bool ARG, bool PERSIST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQuadThreads, 1) k_resident_quad(WarpArgs const A, PersistArgs const P) {
#ifdef ESSA_FORCE_STAGE
    constexpr bool STAGE = true; // (diagnostics: what the staging costs the physics warps, in the flavour that needs none)
#else
    constexpr bool STAGE = ARG || PERSIST; // the physics warps hand every step's state to the courier
#endif
    // the two CTAs use disjoint sets of shared variables: one union, same offsets in both (mapa needs that)
    __shared__ __align__(16) union QShared {
        Q0 a;
        Q1 b;
    } S;
    auto& spos = S.a.spos;
    auto& part = S.a.part;
    auto& stage = S.a.stage;
    auto& stage_mask = S.a.stage_mask;
    auto& stagebar = S.a.stagebar;
    auto& courier_done = S.a.courier_done;
    auto& fin_sent = S.a.fin_sent;
    auto& done = S.a.done;
    auto& cmdw = S.a.cmdw;
    auto& cmdt0 = S.a.cmdt0;
    auto& ring = S.b.ring;
    auto& full = S.b.full;
    auto& ready = S.b.ready;
    auto& sgf = S.b.sgf;
    auto& ctl = S.b.ctl;
    auto& fin = S.b.fin;
    auto& fin_meta = S.b.fin_meta;
    auto& finbar = S.b.finbar;
    auto& outbar = S.b.outbar;
    auto& outd = S.b.outd;
    auto& outmost = S.b.outmost;

    int const n = A.n;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned const cta = cluster_ctarank();
    bool const rec_on = A.record != 0;
    constexpr bool snap_on = ARG; // recording implies tracking
    unsigned const snap_bytes = (unsigned)n * 48u + 8u;
    unsigned const fin_bytes = (unsigned)n * 72u + 16u;
    unsigned const nout = (snap_on ? 1u : 0u) + (rec_on ? 2u : 0u); // consumer warps that report a tick's end

    if (threadIdx.x == 0) {
        if (cta == 1) {
            for (int i = 0; i < kRing; i++) {
                mbar_init(&full[i], 1);
                mbar_init(&ready[i], kAttrWarps);
            }
            mbar_init(&finbar, 1);
            mbar_init(&outbar, nout ? nout : 1u);
            mbar_fence_init();
            for (int i = 0; i < kRing; i++)
                mbar_expect_tx(&full[i], snap_bytes);
            mbar_expect_tx(&finbar, fin_bytes);
            ctl[0] = (unsigned)A.steps;
            ctl[1] = 0;
        } else {
            for (int i = 0; i < kStage; i++)
                mbar_init(&stagebar[i], kQuadWarps);
            mbar_fence_init();
            for (int i = 0; i < kConsumers; i++)
                done[i] = 0;
            courier_done = 0;
            fin_sent = 0;
        }
    }
    if (cta == 1 && threadIdx.x < 32) {
        bool const in = (int)threadIdx.x < n;
        int const i = in ? (int)threadIdx.x : 0;
        sgf[threadIdx.x] = in ? A.w.gf[i] : 0.0;
        if (PERSIST) { // what the consumers overwrite at every tick's end for the bodies they handle; the rest stays
            outd[0][threadIdx.x] = in ? A.w.maxattr[i] : 0.0;
            outd[1][threadIdx.x] = in ? A.r.ap[i] : 0.0;
            outd[2][threadIdx.x] = in ? A.r.pe[i] : 0.0;
            outd[3][threadIdx.x] = in ? A.r.ap_vel[i] : 0.0;
            outd[4][threadIdx.x] = in ? A.r.pe_vel[i] : 0.0;
            outmost[threadIdx.x] = in ? A.w.most[i] : -1;
        }
    }
    __syncthreads();
    cluster_sync_all();

    // persistent mode, CTA 1: wait until steps beyond `end` are granted (the doorbell, forwarded by CTA 0); false: retire
    auto more_steps = [&](int s, int& end) {
        if constexpr (PERSIST) {
            bool quit;
            for (;;) {
                end = (int)ld_acquire_cluster_u32(&ctl[0]);
                quit = ld_acquire_cluster_u32(&ctl[1]) != 0;
                if (s < end || quit)
                    break;
                __nanosleep(64); // (a spinning warp would take issue slots from the warps that still work)
            }
        }
        return s < end;
    };

    if (cta == 1 && warp < kAttrWarps) {
        // ---------------- attraction warps: the step's most-attracting candidates (src/Object.cpp:84-94) ----------------
        // Which heavier body pulls hardest on body k depends only on the step's final positions and never feeds back into
        // the trajectory.  The three warps share the bodies of ONE step (a third each; a lane is a partner, four bodies in
        // flight at a time), so a step's winners are known a few hundred cycles after its snapshot landed; the one
        // sequential piece -- "keep the previous link when there is no candidate" -- is a register in the consumer warps.
        // Candidates are ranked by gf_p y0^4 (1 + 2 e) ~ gf_p / r^4 (y0: rsqrt seed, e = 1 - r^2 y0^2; relative error
        // ~1e-11) as sortable keys -- bit pattern with the 5 lowest mantissa bits replaced by 31 - p -- and the winner of a
        // body comes from two full-warp integer reductions (redux.sync; with a partial mask the intrinsic is a loop);
        // max_attraction is formed to full precision for a tick's final state only.
        if (snap_on) {
            int const per_w = (n + kAttrWarps - 1) / kAttrWarps; // bodies per warp
            int const k0 = warp * per_w, k1 = min(n, k0 + per_w);
            int const p = lane;                                   // a lane is a partner
            int const pc = p < n ? p : 0;
            double const gp = sgf[pc];
            int end = A.steps;
            for (int s = 0;; s++) {
                if (s >= end && !more_steps(s, end))
                    break;
                int const slot = s % kRing;
                mbar_wait(&full[slot], (unsigned)(s / kRing) & 1u);
                if (warp == 0 && lane == 0)
                    mbar_expect_tx(&full[slot], snap_bytes); // armed for the slot's next use (16 steps from now)
                QSnap& sn = ring[slot];
                unsigned const amask = (unsigned)sn.mask;
                bool const inp = p < n && ((amask >> p) & 1u);
                double const xp = sn.x[pc], yp = sn.y[pc], zp = sn.z[pc];
                for (int kb = k0; kb < k1; kb += 4) {
                    unsigned hi[4], lo[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { // four bodies at a time: independent chains
                        int const k = min(kb + u, n - 1);
                        double const gfk = sgf[k];
                        bool const ok = inp && p != k && gp > gfk;
                        double const dx = xp - sn.x[k], dy = yp - sn.y[k], dz = zp - sn.z[k];
                        double const r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        double const y0 = rsqrt_seed(r2);
                        double const y2 = y0 * y0;
                        double const e = fma(-r2, y2, 1.0);
                        double const t4 = gp * (y2 * y2);
                        double const m = fma(t4, e + e, t4);
                        unsigned const mh = (unsigned)__double2hiint(m), ml = (unsigned)__double2loint(m);
                        bool const cand = ok && (mh | ml) != 0u;
                        hi[u] = cand ? mh : 0u;
                        lo[u] = cand ? ((ml & ~31u) | (unsigned)(31 - p)) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        unsigned const mh = __reduce_max_sync(0xffffffffu, hi[u]);
                        unsigned const ml = __reduce_max_sync(0xffffffffu, hi[u] == mh ? lo[u] : 0u);
                        int const k = kb + u;
                        if (lane == 0 && k < k1)
                            sn.win[k] = ((mh | ml) != 0u && ((amask >> k) & 1u)) ? 31 - (int)(ml & 31u) : -1;
                    }
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&ready[slot]);
            }
        }
    } else if (cta == 1 && warp <= kPeWarp) {
        // ---------------- consumer warps: 3 .. 6 Trail (bodies by index mod 4); 7 links + History; 8 apoapsis; 9 periapsis ----------------
        // Every consumer follows the most-attracting links itself (a few integer instructions per step).  One copy of the loop
        // per role (ROLE is a compile-time constant in each): a step costs a warp ~5 cycles per instruction and ~20 per taken
        // branch, and only the links + ap / pe warps are waited for at a tick's end (the Trail warps may lag up to the depth of
        // the ring and catch up while the host reads the tick).
        auto consumer = [&](auto role) {
            constexpr int ROLE = decltype(role)::value; // 0: Trail, 1: links + History, 2: apoapsis, 3: periapsis (the two halves of
                                                        // Object.cpp:111-123 are independent chains of ~30 dependent instructions each)
            bool const body = lane < n;
            bool const mine = rec_on && body && lane < A.r.tracked && (ROLE != 0 || (lane & (kTrailWarps - 1)) == (warp - kTrailWarp0));
            uint32_t const done_remote = mapa_u32(&done[warp - kTrailWarp0], 0);
            BodyRec b;
            if (mine)
                rec_load(A.r, lane, b);
            int most = body ? A.w.most[lane] : -1, old_most = body ? A.w.old_most[lane] : -1;
            double maxattr = (ROLE == 1 && body) ? A.w.maxattr[lane] : 0.0;
            bool const offset_trails = A.offset_trails != 0, forward_sim = A.forward_sim != 0;
#ifdef ESSA_STAGE_CLOCKS
            long long busy = 0;
#endif
            int end = A.steps; // steps granted so far (persistent mode: grows with every doorbell)
            for (int s = 0;; s++) {
                if (s >= end && !more_steps(s, end))
                    break;
                // the host extends `end` only after it has this tick's result, i.e. after this warp reported its last step
                bool const last = s == end - 1;
                int const slot = s % kRing;
                mbar_wait(&ready[slot], (unsigned)(s / kRing) & 1u);
#ifdef ESSA_STAGE_CLOCKS
                long long t0 = clock64();
#endif
                QSnap const& sn = ring[slot];
                bool const act = body && ((sn.mask >> lane) & 1ull);
                int const winner = body ? sn.win[lane] : -1;
                ESSA_CHECK(slot >= 0 && slot < kRing && winner >= -1 && winner < n && most >= -1 && most < n);
                old_most = most; // Object::before_update (every object)
                if (act && winner >= 0)
                    most = winner;
                if (mine && act) {
                    int const mi = most >= 0 ? most : 0;
                    if (ROLE == 0)
                        record_trail(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane], sn.x[mi], sn.y[mi],
                            sn.z[mi], most, old_most, offset_trails, forward_sim, false);
                    else if (ROLE == 1) {
                        if (!forward_sim)
                            history_push(A.r, lane, b, sn.x[lane], sn.y[lane], sn.z[lane], sn.vx[lane], sn.vy[lane], sn.vz[lane]);
                    } else if (most >= 0) {
                        double const dx = __dsub_rn(sn.x[lane], sn.x[mi]), dy = __dsub_rn(sn.y[lane], sn.y[mi]), dz = __dsub_rn(sn.z[lane], sn.z[mi]);
                        double const d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (ROLE == 2)
                            apoapsis_step(b, d2, sn.vx[lane], sn.vy[lane], sn.vz[lane]);
                        else
                            periapsis_step(b, d2, sn.vx[lane], sn.vy[lane], sn.vz[lane]);
                    }
                }
                if (ROLE == 1 && last && act) {
                    // max_attraction as the last force pass left it: the winner's gf / r^4 to full precision, or 0
                    maxattr = winner >= 0 ? attraction_metric(sn.x[lane], sn.y[lane], sn.z[lane], sn.x[winner], sn.y[winner], sn.z[winner],
                                                sgf[winner])
                                          : 0.0;
                }
                if (PERSIST && ROLE != 0 && last) { // what the host reads back after the tick: handed to the publisher warp
                    if (ROLE == 1 && body) {
                        outd[0][lane] = maxattr;
                        outmost[lane] = most;
                    }
                    if (ROLE == 2 && mine) {
                        ap_flush(b);
                        outd[1][lane] = b.ap;
                        outd[3][lane] = b.apv;
                    }
                    if (ROLE == 3 && mine) {
                        pe_flush(b);
                        outd[2][lane] = b.pe;
                        outd[4][lane] = b.pev;
                    }
                }
                __syncwarp();
#ifdef ESSA_STAGE_CLOCKS
                busy += clock64() - t0;
#endif
                if (lane == 0) {
                    if (PERSIST && ROLE != 0 && last)
                        mbar_arrive(&outbar);
                    st_relaxed_cluster_u32(done_remote, (unsigned)(s + 1));
                }
            }
#ifdef ESSA_STAGE_CLOCKS
            if (lane == 0)
                printf("consumer warp %d: %lld busy clk/step\n", warp, busy / max(end, 1));
#endif
            if (ROLE == 1 && body) {
                A.w.most[lane] = most;
                A.w.old_most[lane] = old_most;
                A.w.maxattr[lane] = maxattr;
            }
            if (mine) {
                if (ROLE == 0) {
                    rec_store_trail(A.r, lane, b, false);
                } else if (ROLE == 1) {
                    A.r.hist_start[lane] = b.hist_start;
                    A.r.hist_len[lane] = b.hist_len;
                } else if (ROLE == 2) {
                    ap_flush(b);
                    A.r.ap[lane] = b.ap;
                    A.r.ap_vel[lane] = b.apv;
                } else {
                    pe_flush(b);
                    A.r.pe[lane] = b.pe;
                    A.r.pe_vel[lane] = b.pev;
                }
            }
        };
        if (snap_on && (warp == kHistWarp || rec_on)) {
            if (warp == kHistWarp)
                consumer(std::integral_constant<int, 1> {});
            else if (warp == kApWarp)
                consumer(std::integral_constant<int, 2> {});
            else if (warp == kPeWarp)
                consumer(std::integral_constant<int, 3> {});
            else
                consumer(std::integral_constant<int, 0> {});
        }
    } else if (PERSIST && cta == 1 && warp == kPubWarp) {
        // ---------------- publisher ----------------
        // Word w of the host buffer carries 32 bits of payload under the tick's number: fields x y z vx vy vz ax ay az
        // max_attraction ap pe ap_vel pe_vel as (low, high) halves of body 0, 1, ..., then the links, then the device
        // nanoseconds between doorbell and publication.
        for (unsigned tick = 0;; tick++) {
            bool quit = false;
            while (!mbar_try(&finbar, tick & 1u)) {
                if (ld_acquire_cluster_u32(&ctl[1]) != 0) {
                    quit = true;
                    break;
                }
            }
            if (quit)
                break;
            if (lane == 0)
                mbar_expect_tx(&finbar, fin_bytes); // armed for the next tick's bundle
            if (nout)
                mbar_wait(&outbar, tick & 1u);
            unsigned const seq = (unsigned)fin_meta[1];
            unsigned long long const tag = (unsigned long long)seq << 32;
            unsigned long long const t0 = fin_meta[0];
            // field-major words (ctx.hpp): a lane is (body, half) of a field: no division, 14 coalesced stores of 2 n words
            int const words_d = 28 * n;
            for (int h = 0; h < 2; h++) {
                int const r = lane + 32 * h;
                if (r < 2 * n) {
                    int const k = r >> 1;
#pragma unroll
                    for (int f = 0; f < 14; f++) {
                        double const d = f < 9 ? fin[f][k] : outd[f - 9][k];
                        unsigned const v = (r & 1) ? (unsigned)__double2hiint(d) : (unsigned)__double2loint(d);
                        st_relaxed_sys_u64(P.out + f * 2 * n + r, tag | v);
                    }
                }
            }
            if (lane < n)
                st_relaxed_sys_u64(P.out + words_d + lane, tag | (unsigned)outmost[lane]);
            if (lane == 0) {
                unsigned long long const dtn = globaltimer_ns() - t0;
                st_relaxed_sys_u64(P.out + words_d + n, tag | (unsigned)dtn);
                st_relaxed_sys_u64(P.out + words_d + n + 1, tag | (unsigned)(dtn >> 32));
            }
        }
    } else if (STAGE && cta == 0 && warp == kCourierWarp) {
        // ---------------- courier: forwards what the physics warps staged, so that they never issue a remote store ----------------
        uint32_t const full_remote = mapa_u32(&full[0], 1);
        uint32_t const ring_remote = mapa_u32(&ring[0], 1);
        uint32_t const fr = mapa_u32(&fin[0][0], 1) + 8u * (uint32_t)lane, fb = mapa_u32(&finbar, 1), fm = mapa_u32(&fin_meta[0], 1);
        int consumed = 0; // cached lower bound of the steps CTA 1 has finished with
        int s = 0, end = A.steps;
        unsigned seq = P.seq0;
        unsigned long long tick_t0 = PERSIST ? globaltimer_ns() : 0ull;
        for (;;) {
            for (; s < end; s++) {
                int const buf = s % kStage;
                mbar_wait(&stagebar[buf], (unsigned)(s / kStage) & 1u);
                bool const tick_end = PERSIST && s == end - 1;
                double v[9];
#pragma unroll
                for (int c = 0; c < 6; c++)
                    v[c] = stage[buf][c][lane];
#pragma unroll
                for (int c = 6; c < 9; c++)
                    v[c] = tick_end ? stage[buf][c][lane] : 0.0;
                unsigned const amask = stage_mask[buf];
                __syncwarp();
                if (lane == 0)
                    st_release_cta_u32(&courier_done, (unsigned)(s + 1)); // the buffer may be rewritten
                if (snap_on) {
                    int const slot = s % kRing;
                    if (s - consumed >= kRing) { // the slot may still be in use: look at CTA 1's progress
                        do {
                            consumed = (int)ld_acquire_cluster_u32(&done[kHistWarp - kTrailWarp0]);
                            if (rec_on) {
#pragma unroll
                                for (int i = 0; i < kConsumers; i++)
                                    if (i != kHistWarp - kTrailWarp0)
                                        consumed = min(consumed, (int)ld_acquire_cluster_u32(&done[i]));
                            }
                        } while (s - consumed >= kRing);
                    }
                    uint32_t const dst = ring_remote + (uint32_t)slot * (uint32_t)sizeof(QSnap) + 8u * (uint32_t)lane;
                    uint32_t const bar = full_remote + 8u * (uint32_t)slot;
                    if (lane < n) {
#pragma unroll
                        for (int c = 0; c < 6; c++)
                            st_async_f64(dst + 256u * c, v[c], bar);
                    }
                    if (lane == 0)
                        st_async_u64(ring_remote + (uint32_t)slot * (uint32_t)sizeof(QSnap) + (uint32_t)offsetof(QSnap, mask),
                            (unsigned long long)amask, bar);
                }
                if (tick_end) { // the tick's final state, for the publisher
                    if (lane < n) {
#pragma unroll
                        for (int c = 0; c < 9; c++)
                            st_async_f64(fr + 256u * c, v[c], fb);
                    }
                    if (lane == 0) {
                        st_async_u64(fm, tick_t0, fb);
                        st_async_u64(fm + 8u, (unsigned long long)seq, fb);
                    }
                }
            }
            if (!PERSIST)
                break;
            __syncwarp();
            if (lane == 0)
                st_release_cta_u32(&fin_sent, ld_acquire_cta_u32(&fin_sent) + 1u); // the poller's idle clock may start
            asm volatile("bar.sync 3, 192;" ::: "memory");
            if (cmdw[0] == (unsigned)end)
                break; // retire
            end = (int)cmdw[0];
            seq = cmdw[1];
            tick_t0 = cmdt0;
        }
    } else if (PERSIST && cta == 0 && warp == kPollWarp) {
        // ---------------- poller: the host's doorbell ----------------
        // One lane reads the command word over PCIe, one read in flight (see ESSA_POLL_DEPTH); on its own warp, so that the
        // round trip never stalls a physics warp.
        unsigned seq = P.seq0;
        unsigned ticks = 1; // granted so far
        int end = A.steps;
#if ESSA_POLL_DEPTH != 1
        long long q = 0;
#endif
        for (;;) {
            unsigned long long c = 0ull;
            bool idle = false;
            if (lane == 0) { // one lane reads: 32 lanes reading the same word of system memory are 32 requests
#if ESSA_POLL_DEPTH == 1
                long long idle0 = 0;
                for (;;) {
                    c = ld_relaxed_sys_u64(P.cmd);
                    if ((unsigned)(c >> 32) != seq)
                        break;
                    if (idle0 == 0) {
                        if (ld_acquire_cta_u32(&fin_sent) == ticks)
                            idle0 = clock64();
                    } else if (clock64() - idle0 > P.idle_clk) {
                        idle = true;
                        break;
                    }
                }
#else
                unsigned long long c0, c1, c2, c3;
                long long idle0 = 0;
                long long t = clock64();
                c0 = ld_relaxed_sys_u64(P.cmd);
                if (q == 0 && (unsigned)(c0 >> 32) == seq) {
                    q = (clock64() - t) >> 2; // (a returned value was consumed: this is one round trip)
                    q = q < 700 ? q : 700;    // (the very first read of the page can take much longer than the rest)
                    t = clock64();
                    c0 = ld_relaxed_sys_u64(P.cmd);
                }
                while (clock64() - t < q) { }
                c1 = ld_relaxed_sys_u64(P.cmd);
                while (clock64() - t < 2 * q) { }
                c2 = ld_relaxed_sys_u64(P.cmd);
                while (clock64() - t < 3 * q) { }
                c3 = ld_relaxed_sys_u64(P.cmd);
                for (;;) {
                    if ((unsigned)((c = c0) >> 32) != seq)
                        break;
                    c0 = ld_relaxed_sys_u64(P.cmd);
                    if ((unsigned)((c = c1) >> 32) != seq)
                        break;
                    c1 = ld_relaxed_sys_u64(P.cmd);
                    if ((unsigned)((c = c2) >> 32) != seq)
                        break;
                    c2 = ld_relaxed_sys_u64(P.cmd);
                    if ((unsigned)((c = c3) >> 32) != seq)
                        break;
                    c3 = ld_relaxed_sys_u64(P.cmd);
                    // the idle clock runs from the moment the tick in flight has left for the host
                    if (idle0 == 0) {
                        if (ld_acquire_cta_u32(&fin_sent) == ticks)
                            idle0 = clock64();
                    } else if (clock64() - idle0 > P.idle_clk) {
                        idle = true;
                        break;
                    }
                }
#endif
            }
            c = __shfl_sync(0xffffffffu, c, 0);
            idle = __shfl_sync(0xffffffffu, idle ? 1 : 0, 0) != 0;
            unsigned const add = idle ? 0u : (unsigned)c; // steps of the new tick; 0: retire
            if (lane == 0) {
                cmdt0 = globaltimer_ns();
                if (add == 0) { // (2: nobody rang for idle_clk cycles; the host starts another kernel when it comes back)
                    st_relaxed_sys_u32(P.ack + 1, idle ? 2u : 1u);
                    st_relaxed_cluster_u32(mapa_u32(&ctl[1], 1), 1u);
                    cmdw[0] = (unsigned)end;
                } else { // `add` more steps, for everybody
                    cmdw[0] = (unsigned)end + add;
                    cmdw[1] = (unsigned)(c >> 32);
                    st_relaxed_cluster_u32(mapa_u32(&ctl[0], 1), (unsigned)end + add);
                }
            }
            __syncwarp();
            asm volatile("bar.arrive 3, 192;" ::: "memory");
            if (add == 0)
                break;
            end += (int)add;
            seq = (unsigned)(c >> 32);
            ticks++;
            // nothing to listen for until this tick has left for the host; asleep, this warp does not compete with physics
            // warp 1 for their sub-partition's issue slots (measured: a spinning neighbour costs the chain ~300 cycles per step)
            while (ld_acquire_cta_u32(&fin_sent) != ticks)
                __nanosleep(200);
        }
    } else if (cta == 0 && warp < kQuadWarps) {
        // ---------------- physics warps 0 .. 3 ----------------
        // One copy of the loop per warp (W is a compile-time constant in each): the step is one dependent chain issued by
        // one warp, where a taken branch costs ~20 cycles -- selecting "what this warp stages" at run time cost 170 cycles
        // per step (measured), three predicated stores in straight-line code cost ~20.
        auto physics = [&](auto role) {
        constexpr int W = decltype(role)::value;
        // Lane layout: PL = 2^LOG2L lanes per body in one contiguous group (lane = kb PL + qb), so that the group sum
        // is a shuffle butterfly: every lane ends with the same bits (a + b == b + a exactly).
        constexpr int PL = 1 << LOG2L;
        int const kb = lane >> LOG2L, qb = lane & (PL - 1);
        bool const pworker = kb < n;
        bool const pleader = pworker && qb == 0;
        int const kc = pworker ? kb : 0;
        double* const sx = spos[W][0];
        double* const sy = spos[W][1];
        double* const sz = spos[W][2];
        double* const sg = spos[W][3];
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
            int p = qb + (W + kQuadWarps * t) * PL;
            bool in = p < n;
            pidx[t] = in ? p : kc; // out of range -> the self term, which vanishes
            pgf[t] = in ? sg[pidx[t]] : 0.0;
        }
        double* const my_part0 = &part[0][W][0][kc];
        double* const my_part1 = &part[1][W][0][kc];
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
        int cdone = 0; // cached lower bound of the steps the courier has forwarded
        // staging addresses of this lane (buffer 0): this warp's rows of stage[][9][32], column kb.  Warp 0 stages x y,
        // warp 1 z vx, warp 2 vy vz, warp 3 the mask (and the accelerations at a tick's end)
        uint32_t const stage_a = keep_in_register(smem_u32(&stage[0][W == 3 ? 6 : 2 * W][kc]));
        uint32_t const mask_a = keep_in_register(smem_u32(&stage_mask[0])), sbar_a = keep_in_register(smem_u32(&stagebar[0])),
                       cdone_a = keep_in_register(smem_u32(&courier_done));

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

        int s = 0, end = A.steps;
        for (;;) {
            for (; s < end; s++) {
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
                            pgf[t] = (qb + (W + kQuadWarps * t) * PL) < n ? sg[pidx[t]] : 0.0;
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
                if (STAGE) {
                    // the step's state goes to a local staging buffer (two shared-memory stores per warp); the courier warp
                    // forwards it to CTA 1
                    unsigned const buf = (unsigned)s % kStage;
                    if (s - cdone >= kStage) {
                        do
                            cdone = (int)lds_acquire_u32(cdone_a);
                        while (s - cdone >= kStage);
                    }
                    uint32_t const a = stage_a + buf * (uint32_t)(9 * 32 * sizeof(double));
                    if (W == 0) {
                        if (pleader) {
                            sts_f64(a, x);
                            sts_f64(a + 256u, y);
                        }
                    } else if (W == 1) {
                        if (pleader) {
                            sts_f64(a, z);
                            sts_f64(a + 256u, vx);
                        }
                    } else if (W == 2) {
                        if (pleader) {
                            sts_f64(a, vy);
                            sts_f64(a + 256u, vz);
                        }
                    } else {
                        if (lane == 0)
                            sts_u32(mask_a + 4u * buf, amask);
                        if (PERSIST && s == end - 1 && pleader) { // the accelerations leave with a tick's final state only
                            sts_f64(a, ax);
                            sts_f64(a + 256u, ay);
                            sts_f64(a + 512u, az);
                        }
                    }
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive_a(sbar_a + 8u * buf);
                }
            }
            if (!PERSIST)
                break;
            // ---------------- end of a tick: wait for the next doorbell (poller warp) ----------------
            // (the next write to cmdw is a tick away: every step in between has a 128-thread barrier of its own)
            asm volatile("bar.sync 3, 192;" ::: "memory");
            if (cmdw[0] == (unsigned)end)
                break; // retire
            end = (int)cmdw[0];
        }

        if (W == 0 && pleader) {
            A.w.pos_cur[kb] = make_double4(x, y, z, active ? gfk : 0.0);
            A.w.v[0][kb] = vx;
            A.w.v[1][kb] = vy;
            A.w.v[2][kb] = vz;
            A.w.a[0][kb] = ax;
            A.w.a[1][kb] = ay;
            A.w.a[2][kb] = az;
            A.active_out[kb] = active ? 1 : 0;
        }
        };
        switch (warp) {
        case 0: physics(std::integral_constant<int, 0> {}); break;
        case 1: physics(std::integral_constant<int, 1> {}); break;
        case 2: physics(std::integral_constant<int, 2> {}); break;
        default: physics(std::integral_constant<int, 3> {}); break;
        }
    }
    // nobody leaves while a peer may still store into (or signal) its shared memory
    cluster_sync_all();
}

template<int T, int LOG2L>
static void launch_quad(WarpArgs const& A, PersistArgs const* pa, cudaStream_t st) {
    // one cluster of two CTAs
    if (pa) {
        if (A.argmax)
            k_resident_quad<T, LOG2L, true, true><<<2, kQuadThreads, 0, st>>>(A, *pa);
        else
            k_resident_quad<T, LOG2L, false, true><<<2, kQuadThreads, 0, st>>>(A, *pa);
    } else {
        if (A.argmax)
            k_resident_quad<T, LOG2L, true, false><<<2, kQuadThreads, 0, st>>>(A, PersistArgs {});
        else
            k_resident_quad<T, LOG2L, false, false><<<2, kQuadThreads, 0, st>>>(A, PersistArgs {});
    }
}

cudaError_t launch_resident_quad(essa_ctx* c, WarpArgs const& A, PersistArgs const* pa, cudaStream_t st) {
    // physics lanes per body: the largest power of two that fits 32 lanes
    int log2l = 0;
    while ((2 << log2l) * c->n <= 32)
        log2l++;
    int const pl = 1 << log2l;
    int const slots = (c->n + pl - 1) / pl;               // partner slots per lane
    int const T = (slots + kQuadWarps - 1) / kQuadWarps;   // ... per physics warp
    if (log2l == 4)
        launch_quad<1, 4>(A, pa, st);
    else if (log2l == 3)
        launch_quad<1, 3>(A, pa, st);
    else if (log2l == 2)
        launch_quad<1, 2>(A, pa, st);
    else if (log2l == 1 && T <= 1)
        launch_quad<1, 1>(A, pa, st);
    else if (log2l == 1)
        launch_quad<2, 1>(A, pa, st);
    else if (T <= 1)
        launch_quad<1, 0>(A, pa, st);
    else if (T <= 2)
        launch_quad<2, 0>(A, pa, st);
    else if (T <= 4)
        launch_quad<4, 0>(A, pa, st);
    else
        launch_quad<8, 0>(A, pa, st);
    return cudaGetLastError();
}

} // namespace essa
