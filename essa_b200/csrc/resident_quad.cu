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
//                     broadcast.  The most-attracting candidate (src/Object.cpp:84-94) rides along: a sortable
//                     64-bit key per partner -- bit pattern of gf_p / r^4 (to ~1e-12: the refined reciprocal) with
//                     the 5 lowest mantissa bits replaced by 31 - p -- maximised with integer instructions next to
//                     the floating-point sums (round 1 ran this on three more warps of CTA 1: 800 cycles of latency
//                     per step that every World::update call paid at its end).  Warps 1-3 publish the step's
//                     snapshot (positions / velocities / mask + winners) into CTA 1's ring with st.async: the store
//                     itself completes the ring slot's mbarrier there (transaction bytes).
//   CTA 1  warps 0, 1 Trail (even / odd bodies: an append stalls the whole warp)
//          warp 2     link bookkeeping + History
//          warp 3     ap / pe
//          warp 4     publisher (persistent mode)
//                     The consumers report progress with a relaxed store into CTA 0's shared memory; the physics
//                     warps look at it only when their cached copy says the ring could be full.
//
// Persistent mode (essa_persist_configure, capi.cu): after the steps of a call the physics warps send their final
// state to CTA 1 (st.async again), and warp 0 starts polling the host's command word in mapped pinned memory, four
// reads in flight a quarter of a PCIe round trip apart.  CTA 1's publisher waits for that bundle and for the
// consumers' last step, then writes the packed state to the host as self-validating 64-bit words {sequence number,
// 32 bits of payload} (what NCCL's LL protocol does over PCIe): no system-wide fence, no separate
// acknowledgement -- the host has a tick's result when every word carries the tick's number.
#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"
#include "resident_small.cuh"

namespace essa {

constexpr int kQuadWarps = 4;
constexpr int kHistWarp = 2, kAppeWarp = 3, kPubWarp = 4; // CTA 1
constexpr int kQuadThreads = 32 * 5;

struct QSnap {
    double x[32], y[32], z[32], vx[32], vy[32], vz[32];
    int win[32];             // the step's most-attracting winner (-1: none, or body masked)
    unsigned long long mask; // active bodies
};

__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, unsigned v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template<int T, int LOG2L, bool ARG, bool PERSIST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQuadThreads, 1) k_resident_quad(WarpArgs const A, PersistArgs const P) {
    __shared__ double spos[kQuadWarps][4][32];            // CTA 0: per physics warp sx, sy, sz, sg (private copies)
    __shared__ double part[2][kQuadWarps][4][32];         // CTA 0: partial accelerations + candidate keys, double-buffered
    __shared__ unsigned done[4];                          // CTA 0: steps consumed by CTA 1's warps 0 .. 3
    __shared__ unsigned cmdw[2];                          // CTA 0: poller -> physics warps: steps granted so far, tick number
    __shared__ unsigned long long cmdt0;                  // CTA 0: globaltimer at the doorbell
    __shared__ __align__(16) QSnap ring[kRing];           // CTA 1
    __shared__ __align__(8) unsigned long long full[kRing]; // CTA 1: a slot's snapshot has arrived (transaction bytes)
    __shared__ double sgf[32];                            // CTA 1: gravity factors
    __shared__ unsigned ctl[2];                           // CTA 1: steps granted so far, retire flag
    __shared__ double fin[9][32];                         // CTA 1: x y z vx vy vz ax ay az as the tick left them
    __shared__ unsigned long long fin_meta[2];            // CTA 1: globaltimer at the doorbell, tick number
    __shared__ __align__(8) unsigned long long finbar, outbar; // CTA 1: bundle arrived / consumers finished the last step
    __shared__ double outd[5][32];                        // CTA 1: max_attraction, ap, pe, ap_vel, pe_vel at the tick's end
    __shared__ int outmost[32];                           // CTA 1: most-attracting links at the tick's end

    int const n = A.n;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned const cta = cluster_ctarank();
    bool const rec_on = A.record != 0;
    bool const snap_on = ARG; // recording implies tracking
    unsigned const snap_bytes = (unsigned)n * 48u + 8u + (ARG ? (unsigned)n * 4u : 0u);
    unsigned const fin_bytes = (unsigned)n * 72u + 16u;
    unsigned const nout = (snap_on ? 1u : 0u) + (rec_on ? 1u : 0u); // consumer warps that report a tick's end

    if (threadIdx.x == 0) {
        if (cta == 1) {
            for (int i = 0; i < kRing; i++)
                mbar_init(&full[i], 1);
            mbar_init(&finbar, 1);
            mbar_init(&outbar, nout ? nout : 1u);
            mbar_fence_init();
            for (int i = 0; i < kRing; i++)
                mbar_expect_tx(&full[i], snap_bytes);
            mbar_expect_tx(&finbar, fin_bytes);
            ctl[0] = (unsigned)A.steps;
            ctl[1] = 0;
        } else {
            done[0] = done[1] = done[2] = done[3] = 0;
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

    if (cta == 1 && warp <= kAppeWarp) {
        // ---------------- consumer warps: 0, 1 Trail; 2 links + History; 3 ap / pe ----------------
        // Every consumer follows the most-attracting links itself (a few integer instructions per step).
        if (snap_on && (warp == kHistWarp || rec_on)) {
            bool const body = lane < n;
            bool const is_trail = warp < kHistWarp;
            bool const mine = rec_on && body && lane < A.r.tracked && (!is_trail || (lane & 1) == warp);
            uint32_t const done_remote = mapa_u32(&done[warp], 0);
            BodyRec b;
            if (mine)
                rec_load(A.r, lane, b);
            int most = body ? A.w.most[lane] : -1, old_most = body ? A.w.old_most[lane] : -1;
            double maxattr = (warp == kHistWarp && body) ? A.w.maxattr[lane] : 0.0;
            bool const offset_trails = A.offset_trails != 0, forward_sim = A.forward_sim != 0;
#ifdef ESSA_STAGE_CLOCKS
            long long busy = 0;
#endif
            int end = A.steps; // steps granted so far (persistent mode: grows with every doorbell)
            for (int s = 0;; s++) {
                if (s >= end) {
                    if constexpr (PERSIST) { // wait for the next doorbell (forwarded by CTA 0), or for the order to retire
                        bool quit;
                        do {
                            end = (int)ld_acquire_cluster_u32(&ctl[0]);
                            quit = ld_acquire_cluster_u32(&ctl[1]) != 0;
                        } while (s >= end && !quit);
                    }
                    if (s >= end)
                        break;
                }
                // the host extends `end` only after it has this tick's result, i.e. after this warp reported its last step
                bool const last = s == end - 1;
                int const slot = s % kRing;
                mbar_wait(&full[slot], (unsigned)(s / kRing) & 1u);
                if (warp == kHistWarp && lane == 0)
                    mbar_expect_tx(&full[slot], snap_bytes); // armed for the slot's next use (16 steps from now)
#ifdef ESSA_STAGE_CLOCKS
                long long t0 = clock64();
#endif
                QSnap const& sn = ring[slot];
                bool const act = body && ((sn.mask >> lane) & 1ull);
                int const winner = body ? sn.win[lane] : -1;
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
                if (warp == kHistWarp && last && act) {
                    // max_attraction as the last force pass left it: the winner's gf / r^4 to full precision, or 0
                    maxattr = winner >= 0 ? attraction_metric(sn.x[lane], sn.y[lane], sn.z[lane], sn.x[winner], sn.y[winner], sn.z[winner],
                                                sgf[winner])
                                          : 0.0;
                }
                if (PERSIST && last) { // what the host reads back after the tick: handed to the publisher warp
                    if (warp == kHistWarp && body) {
                        outd[0][lane] = maxattr;
                        outmost[lane] = most;
                    }
                    if (warp == kAppeWarp && mine) {
                        ap_flush(b);
                        pe_flush(b);
                        outd[1][lane] = b.ap;
                        outd[2][lane] = b.pe;
                        outd[3][lane] = b.apv;
                        outd[4][lane] = b.pev;
                    }
                }
                __syncwarp();
#ifdef ESSA_STAGE_CLOCKS
                busy += clock64() - t0;
#endif
                if (lane == 0) {
                    if (PERSIST && last && warp >= kHistWarp)
                        mbar_arrive(&outbar);
                    st_relaxed_cluster_u32(done_remote, (unsigned)(s + 1));
                }
            }
#ifdef ESSA_STAGE_CLOCKS
            if (lane == 0)
                printf("consumer warp %d: %lld busy clk/step\n", warp, busy / max(end, 1));
#endif
            if (warp == kHistWarp && body) {
                A.w.most[lane] = most;
                A.w.old_most[lane] = old_most;
                A.w.maxattr[lane] = maxattr;
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
        }
    } else if (PERSIST && cta == 1 && warp == kPubWarp) {
        // ---------------- publisher ----------------
        // Word w of the host buffer carries 32 bits of payload under the tick's number: fields x y z vx vy vz ax ay az
        // max_attraction ap pe ap_vel pe_vel as (low, high) halves of body 0, 1, ..., then the links, then the device
        // nanoseconds between doorbell and publication.
        int const words_d = 28 * n, words = words_d + n + 2;
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
            unsigned long long const tag = (unsigned long long)(unsigned)fin_meta[1] << 32;
            unsigned long long const t0 = fin_meta[0];
            for (int w = lane; w < words; w += 32) {
                unsigned v;
                if (w < words_d) {
                    int const f = w / (2 * n), r = w - f * 2 * n, k = r >> 1;
                    double const d = f < 9 ? fin[f][k] : outd[f - 9][k];
                    v = (r & 1) ? (unsigned)__double2hiint(d) : (unsigned)__double2loint(d);
                } else if (w < words_d + n) {
                    v = (unsigned)outmost[w - words_d];
                } else {
                    unsigned long long const dtn = globaltimer_ns() - t0;
                    v = w == words_d + n ? (unsigned)dtn : (unsigned)(dtn >> 32);
                }
                st_relaxed_sys_u64(P.out + w, tag | v);
            }
        }
    } else if (cta == 0 && warp < kQuadWarps) {
        // ---------------- physics warps 0 .. 3 ----------------
        // Lane layout: PL = 2^LOG2L lanes per body in one contiguous group (lane = kb PL + qb), so that the group sum
        // is a shuffle butterfly: every lane ends with the same bits (a + b == b + a exactly).
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

        // this warp's partner slots of this lane never change: indices, weights and who is heavier are loop invariants
        int pidx[T];
        double pgf[T];
        bool heav[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            int p = qb + (warp + kQuadWarps * t) * PL;
            bool in = p < n;
            pidx[t] = in ? p : kc; // out of range -> the self term, which vanishes
            pgf[t] = in ? sg[pidx[t]] : 0.0;
            heav[t] = pgf[t] > gfk; // masked partners weigh 0: never heavier
        }
        double* const my_part0 = &part[0][warp][0][kc];
        double* const my_part1 = &part[1][warp][0][kc];
        double const* const all_part0 = &part[0][0][0][kc];
        double const* const all_part1 = &part[1][0][0][kc];

        bool acc_valid = A.acc_valid != 0;
        long long date = A.date;
        double const hm = A.hm, dtm = A.dtm;
        int par = 0;
        int winner = -1; // the last force pass's most-attracting candidate of body kb
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
            snap_remote += (uint32_t)offsetof(QSnap, x) + 8u * kc;
        else if (warp == 2)
            snap_remote += (uint32_t)offsetof(QSnap, vx) + 8u * kc;
        else
            snap_remote += (uint32_t)offsetof(QSnap, win) + 4u * kc;
        uint32_t const mask_remote = mapa_u32(&ring[0].mask, 1);
        int consumed = 0; // cached lower bound of the steps CTA 1 has finished with

        auto forces = [&]() {
            double fx[T], fy[T], fz[T];
            unsigned long long key = 0ull;
#pragma unroll
            for (int t = 0; t < T; t++) {
                int const pc = pidx[t];
                double const dx = sx[pc] - x, dy = sy[pc] - y, dz = sz[pc] - z;
                double r2 = fma(dx, dx, A.eps2);
                r2 = fma(dy, dy, r2);
                r2 = fma(dz, dz, r2);
                double const y0 = rsqrt_seed(r2);
                double const y2 = y0 * y0; // exact: the seed has 21 significant bits
                double const e = fma(-r2, y2, 1.0);
                double const gy = pgf[t] * y0;
                double const c0 = gy * y2;
                double const w = fma(1.875, e, 1.5);
                double const we = w * e;
                double const sc = fma(c0, we, c0);
                fx[t] = sc * dx;
                fy[t] = sc * dy;
                fz[t] = sc * dz;
                if (ARG) {
                    // |ta|^2 / gf_p = gf_p / r^4 = s (1/r), 1/r ~ y0 (1 + e/2); strictly heavier partners only; the
                    // largest key is the largest metric and, among equal ones, the lowest index
                    double const m = sc * fma(y0 * 0.5, e, y0);
                    unsigned long long const mb = (unsigned long long)__double_as_longlong(m);
                    unsigned long long const ku = (heav[t] && mb != 0ull) ? ((mb & ~31ull) | (unsigned long long)(31 - pc)) : 0ull;
                    key = ku > key ? ku : key;
                }
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
                if (ARG) {
                    unsigned long long const o = __shfl_xor_sync(0xffffffffu, key, off);
                    key = o > key ? o : key;
                }
            }
            double* const mine = par ? my_part1 : my_part0;
            double const* const all = par ? all_part1 : all_part0;
            if (pleader) {
                mine[0] = px;
                mine[32] = py;
                mine[64] = pz;
                if (ARG)
                    mine[96] = __longlong_as_double((long long)key);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // part[par][w][c][k]: w stride 128 doubles, c stride 32
            double tx = (all[0] + all[128]) + (all[256] + all[384]);
            double ty = (all[32] + all[160]) + (all[288] + all[416]);
            double tz = (all[64] + all[192]) + (all[320] + all[448]);
            int win = -1;
            if (ARG) {
                unsigned long long k0 = (unsigned long long)__double_as_longlong(all[96]), k1 = (unsigned long long)__double_as_longlong(all[224]),
                                   k2 = (unsigned long long)__double_as_longlong(all[352]), k3 = (unsigned long long)__double_as_longlong(all[480]);
                k0 = k1 > k0 ? k1 : k0;
                k2 = k3 > k2 ? k3 : k2;
                k0 = k2 > k0 ? k2 : k0;
                win = k0 != 0ull ? 31 - (int)(k0 & 31ull) : -1;
            }
            par ^= 1;
            if (active) {
                ax = tx;
                ay = ty;
                az = tz;
                winner = win;
            }
        };

        // four reads of the host's command word in flight, a quarter of the measured round trip apart: a doorbell is seen
        // about half a one-way trip + an eighth of a round trip after it was rung
        auto doorbell = [&](unsigned seq, bool& idle) -> unsigned long long {
            long long t = clock64();
            unsigned long long c0 = ld_relaxed_sys_u64(P.cmd);
            if ((unsigned)(c0 >> 32) != seq)
                return c0;
            long long const q = (clock64() - t) >> 2; // (a returned value was consumed: this is one round trip)
            unsigned long long c1, c2, c3;
            t = clock64();
            c0 = ld_relaxed_sys_u64(P.cmd);
            while (clock64() - t < q) { }
            c1 = ld_relaxed_sys_u64(P.cmd);
            while (clock64() - t < 2 * q) { }
            c2 = ld_relaxed_sys_u64(P.cmd);
            while (clock64() - t < 3 * q) { }
            c3 = ld_relaxed_sys_u64(P.cmd);
            for (;;) {
                if ((unsigned)(c0 >> 32) != seq)
                    return c0;
                c0 = ld_relaxed_sys_u64(P.cmd);
                if ((unsigned)(c1 >> 32) != seq)
                    return c1;
                c1 = ld_relaxed_sys_u64(P.cmd);
                if ((unsigned)(c2 >> 32) != seq)
                    return c2;
                c2 = ld_relaxed_sys_u64(P.cmd);
                if ((unsigned)(c3 >> 32) != seq)
                    return c3;
                c3 = ld_relaxed_sys_u64(P.cmd);
                if (clock64() - t > P.idle_clk) {
                    idle = true;
                    return 0ull;
                }
            }
        };

        int s = 0, end = A.steps;
        unsigned seq = P.seq0;
        unsigned long long tick_t0 = PERSIST ? globaltimer_ns() : 0ull;
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
                        for (int t = 0; t < T; t++) {
                            pgf[t] = (qb + (warp + kQuadWarps * t) * PL) < n ? sg[pidx[t]] : 0.0;
                            heav[t] = pgf[t] > gfk;
                        }
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
                            consumed = (int)ld_acquire_cluster_u32(&done[kHistWarp]);
                            if (rec_on) {
                                int const d0 = (int)ld_acquire_cluster_u32(&done[0]), d1 = (int)ld_acquire_cluster_u32(&done[1]),
                                          d3 = (int)ld_acquire_cluster_u32(&done[3]);
                                consumed = min(min(consumed, d0), min(d1, d3));
                            }
                        } while (s - consumed >= kRing);
                    }
                    uint32_t const dst = snap_remote + (uint32_t)slot * (uint32_t)sizeof(QSnap);
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
                    } else {
                        if (pleader)
                            st_async_u32(dst, (unsigned)(active ? winner : -1), bar);
                        if (lane == 0)
                            st_async_u64(mask_remote + (uint32_t)slot * (uint32_t)sizeof(QSnap), (unsigned long long)amask, bar);
                    }
                }
            }
            if (!PERSIST)
                break;
            // ---------------- end of a tick: final state to the publisher, then wait for the next doorbell ----------------
            {
                uint32_t const fr = mapa_u32(&fin[0][0], 1) + 8u * (uint32_t)kc, fb = mapa_u32(&finbar, 1);
                if (warp == 1 && pleader) {
                    st_async_f64(fr, x, fb);
                    st_async_f64(fr + 256u, y, fb);
                    st_async_f64(fr + 512u, z, fb);
                } else if (warp == 2 && pleader) {
                    st_async_f64(fr + 768u, vx, fb);
                    st_async_f64(fr + 1024u, vy, fb);
                    st_async_f64(fr + 1280u, vz, fb);
                } else if (warp == 3) {
                    if (pleader) {
                        st_async_f64(fr + 1536u, ax, fb);
                        st_async_f64(fr + 1792u, ay, fb);
                        st_async_f64(fr + 2048u, az, fb);
                    }
                    if (lane == 0) {
                        uint32_t const fm = mapa_u32(&fin_meta[0], 1);
                        st_async_u64(fm, tick_t0, fb);
                        st_async_u64(fm + 8u, (unsigned long long)seq, fb);
                    }
                }
            }
            if (warp == 0 && lane == 0) {
                bool idle = false;
                unsigned long long const c = doorbell(seq, idle);
                unsigned const add = (unsigned)c; // steps of the new tick; 0: retire
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
            // (the next write to cmdw is a tick away: every step in between has a 128-thread barrier of its own)
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (cmdw[0] == (unsigned)end)
                break; // retire
            end = (int)cmdw[0];
            seq = cmdw[1];
            tick_t0 = cmdt0;
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
