// K4 -- ensemble batch: many perturbed copies of one template world advanced in parallel.
//
// Models the reference's forward-simulation use (src/World.cpp:230-238 clone_for_forward_simulation,
// src/essagui/EssaCreateObject.cpp:166-189 update(ticks) on the clone): forward-simulated worlds
// keep the date frozen and record no History; every object tracks its closest approaches
// (src/Object.cpp:405-417).  Instead of one clone stepped inline on the GUI thread, `copies`
// clones are resident in shared memory, several per CTA, for all steps of a launch.
//
// Layout in HBM: state[copy][6][n] (x y z vx vy vz), copy-major, so a CTA's copies are contiguous.
// Fast arithmetic: two bodies per thread (k_ensemble_duo); exact_order: one thread per (copy, body) in the reference's
// summation order (k_ensemble_exact).  Besides the final state every copy can return, for a designated body (the candidate
// of EssaCreateObject), its closest approach to EVERY other body -- distance and both positions, as
// Object::ClosestApproachEntry holds them -- and its trail (positions after every k-th step).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ctx.hpp"
#include "rows.cuh"

namespace essa {

constexpr int kEnsThreadsMax = kEnsembleMaxBodies;

__global__ void k_ensemble_init(double const* __restrict__ tx, double const* __restrict__ tv, double const* __restrict__ tgf,
    double4 const* __restrict__ tpos, uint8_t const* __restrict__ tactive, int cap, int n, int copies, long long first_copy,
    unsigned long long seed, double rel_sigma, double* state, double* gf_out, uint8_t* active_out, double* mind) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        gf_out[idx] = tpos[idx].w; // gf_eff: bodies masked by Object::deleted() are not cloned -> weight 0
        // ... and are neither attracted nor moved; a live body of mass 0 (a test particle) is both
        active_out[idx] = tactive[idx];
    }
    if (idx >= (long long)copies * n)
        return;
    int c = (int)(idx / n), k = (int)(idx - (long long)c * n);
    long long gc = first_copy + c;
    double4 p = tpos[k];
    double comp[6] = { p.x, p.y, p.z, tv[k], tv[cap + k], tv[2 * (size_t)cap + k] };
    double* s = state + (size_t)c * 6 * n;
    for (int q = 0; q < 6; q++)
        s[(size_t)q * n + k] = __dmul_rn(comp[q], ensemble_factor(seed, rel_sigma, gc, k, q));
    mind[(size_t)c * n + k] = 0.0;
}

struct EnsembleArgs {
    double* state;
    double const* gf;
    uint8_t const* active; // Object::deleted() of the template, per body
    double* mind;
    double* apos;          // [copies][n][6] {this xyz, other xyz} where mind was reached (ClosestApproachEntry), or null
    double* trail;         // [copies][trail_cap][3] positions of body `approach_to` after every trail_every-th step, or null
    int* trail_len;        // [copies]
    int n, copies, per_cta, steps;
    double mul, dt, h;
    int approaches, approach_to;
    int trail_every, trail_cap;
    long long steps_done;  // steps of earlier launches (the sampling continues across launches)
};

// Object::update_closest_approaches for one (body, designated body) pair at a step's start: the running minimum of d2 with the
// two positions it was seen at (src/Object.cpp:405-417 stores {this position, other position, distance}).
struct Approach {
    double m2 = -1.0; // < 0: nothing seen since the last restart
    double p[6];
    __device__ __forceinline__ void see(double x, double y, double z, double ox, double oy, double oz, double& md) {
        double dx = __dsub_rn(ox, x), dy = __dsub_rn(oy, y), dz = __dsub_rn(oz, z);
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d2 == 0.0) { // a distance of exactly zero stores 0 == "unset" in the reference: the search restarts
            md = 0.0;
            m2 = -1.0;
        } else if (m2 < 0.0 || d2 < m2) {
            m2 = d2;
            p[0] = x; p[1] = y; p[2] = z; p[3] = ox; p[4] = oy; p[5] = oz;
        }
    }
    // RN o sqrt is monotone, so the minimum of d2 is tracked and the root taken once, at the end of a launch
    __device__ __forceinline__ void settle(double& md, double* apos) const {
        if (m2 < 0.0)
            return;
        double d = __dsqrt_rn(m2);
        if (md == 0.0 || d < md) {
            md = d;
            if (apos)
                for (int c = 0; c < 6; c++)
                    apos[c] = p[c];
        }
    }
};

// The designated body's trail: its position after every trail_every-th step (what EssaCreateObject draws for the candidate).
__device__ __forceinline__ void ens_trail(EnsembleArgs const& A, long long copy, int st, double x, double y, double z) {
    long long const g = A.steps_done + st + 1;
    if (g % A.trail_every != 0)
        return;
    long long const idx = g / A.trail_every - 1;
    if (idx >= A.trail_cap)
        return;
    double* t = A.trail + ((size_t)copy * A.trail_cap + (size_t)idx) * 3;
    t[0] = x; t[1] = y; t[2] = z;
    A.trail_len[copy] = (int)idx + 1;
}

// exact_order flavour: one thread per (copy, body), the reference's summation order (rows.cuh), positions in shared memory,
// copy-major (copy lc owns [lc n, lc n + n) of sx / sy / sz / sg); all copies of a CTA step in lock step.
__global__ void __launch_bounds__(kEnsThreadsMax, 2) k_ensemble_exact(EnsembleArgs const A) {
    extern __shared__ double smem[];
    int const n = A.n;
    int const t = threadIdx.x;
    int const lc = t / n, k = t - lc * n;
    int const copy = blockIdx.x * A.per_cta + lc;
    bool const live = lc < A.per_cta && copy < A.copies;
    size_t const plane = (size_t)A.per_cta * n;
    double* sx = smem + (size_t)(lc < A.per_cta ? lc : 0) * n;
    double* sy = sx + plane;
    double* sz = sy + plane;
    double* sg = sz + plane;

    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0, md = 0, dummy_m = 0;
    int dummy_most = -1;
    double* s = A.state + (size_t)(live ? copy : 0) * 6 * n;
    bool on = false;
    if (live) {
        x = s[k];
        y = s[(size_t)n + k];
        z = s[2 * (size_t)n + k];
        vx = s[3 * (size_t)n + k];
        vy = s[4 * (size_t)n + k];
        vz = s[5 * (size_t)n + k];
        gfk = A.gf[k];
        md = A.mind[(size_t)copy * n + k];
        on = A.active[k] != 0; // a body that deleted() masks is neither attracted nor moved
        sx[k] = x;
        sy[k] = y;
        sz[k] = z;
        sg[k] = gfk;
    }
    __syncthreads();
    if (on)
        row_forces<true, false>(k, n, n, sx, sy, sz, sg, x, y, z, gfk, 0.0, ax, ay, az, dummy_m, dummy_most);
    // (a h) mul == a (h mul) and (v dt) mul == v (dt mul) exactly: mul is +-1
    double const hm = A.h * A.mul, dtm = A.dt * A.mul;
    Approach ap;
    bool const approaches = A.approaches && live && k != A.approach_to;
    int const to = A.approach_to;

    for (int st = 0; st < A.steps; st++) {
        if (approaches) // at the step's start
            ap.see(x, y, z, sx[to], sy[to], sz[to], md);
        if (on) {
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
            x = __dadd_rn(x, __dmul_rn(vx, dtm));
            y = __dadd_rn(y, __dmul_rn(vy, dtm));
            z = __dadd_rn(z, __dmul_rn(vz, dtm));
        }
        __syncthreads();
        if (live) {
            sx[k] = x;
            sy[k] = y;
            sz[k] = z;
        }
        __syncthreads();
        if (on) {
            row_forces<true, false>(k, n, n, sx, sy, sz, sg, x, y, z, gfk, 0.0, ax, ay, az, dummy_m, dummy_most);
            vx = __dadd_rn(vx, __dmul_rn(ax, hm));
            vy = __dadd_rn(vy, __dmul_rn(ay, hm));
            vz = __dadd_rn(vz, __dmul_rn(az, hm));
        }
        if (A.trail && live && k == to)
            ens_trail(A, copy, st, x, y, z);
    }
    if (approaches)
        ap.settle(md, A.apos ? A.apos + ((size_t)copy * n + k) * 6 : nullptr);
    if (live) {
        s[k] = x;
        s[(size_t)n + k] = y;
        s[2 * (size_t)n + k] = z;
        s[3 * (size_t)n + k] = vx;
        s[4 * (size_t)n + k] = vy;
        s[5 * (size_t)n + k] = vz;
        A.mind[(size_t)copy * n + k] = md;
    }
}

// Fast-arithmetic ensembles, two bodies per thread.  With one body per thread every interaction costs four
// 8-byte shared-memory loads per lane, and shared-memory bandwidth (63 % of its peak, ncu) holds the FP64 pipe
// at 67 %.  Here a thread owns bodies h and h + ceil(n / 2) of its copy and walks ALL n partners (the self term
// vanishes: the rsqrt seed of 0 is 0), two at a time: one 32-byte record {x, y, z, gf} per partner, read at the
// same address by every thread of the copy (a broadcast), feeds two interactions, and the four chains of an
// iteration are independent.  WARP = true: templates of up to 64 bodies, a warp owns 32 / ceil(n / 2) whole
// copies and steps them with __syncwarp only; otherwise 256 / ceil(n / 2) copies per CTA behind __syncthreads.
constexpr int kDuoWarps = 4;
constexpr int kDuoCtaThreads = 256;

template<bool WARP>
__global__ void __launch_bounds__(WARP ? 32 * kDuoWarps : kDuoCtaThreads, WARP ? 4 : 2) k_ensemble_duo(EnsembleArgs const A) {
    extern __shared__ __align__(16) unsigned char duo_smem[];
    int const n = A.n, half = (n + 1) >> 1, rs = n + 1; // records per copy: n bodies + one zero-weight pad
    int lc, h, group_copies;
    long long copy;
    double4* rec;
    if (WARP) {
        int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        group_copies = 32 / half;
        lc = lane / half;
        h = lane - lc * half;
        copy = ((long long)blockIdx.x * kDuoWarps + warp) * group_copies + lc;
        rec = reinterpret_cast<double4*>(duo_smem) + ((size_t)warp * group_copies + (lc < group_copies ? lc : 0)) * rs;
    } else {
        group_copies = A.per_cta;
        lc = threadIdx.x / half;
        h = threadIdx.x - lc * half;
        copy = (long long)blockIdx.x * group_copies + lc;
        rec = reinterpret_cast<double4*>(duo_smem) + (size_t)(lc < group_copies ? lc : 0) * rs;
    }
    bool const live = lc < group_copies && copy < A.copies;
    auto sync = [&]() {
        if (WARP)
            __syncwarp();
        else
            __syncthreads();
    };

    int const kb[2] = { h, h + half };
    bool own[2] = { live, live && h + half < n };
    double x[2], y[2], z[2], vx[2], vy[2], vz[2], ax[2], ay[2], az[2], gf[2], md[2];
    Approach ap[2];
    bool on[2] = { false, false };
    double* s = A.state + (size_t)(live ? copy : 0) * 6 * n;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        x[b] = y[b] = z[b] = vx[b] = vy[b] = vz[b] = ax[b] = ay[b] = az[b] = gf[b] = md[b] = 0.0;
        if (own[b]) {
            int const k = kb[b];
            x[b] = s[k];
            y[b] = s[(size_t)n + k];
            z[b] = s[2 * (size_t)n + k];
            vx[b] = s[3 * (size_t)n + k];
            vy[b] = s[4 * (size_t)n + k];
            vz[b] = s[5 * (size_t)n + k];
            gf[b] = A.gf[k];
            md[b] = A.mind[(size_t)copy * n + k];
            on[b] = A.active[k] != 0; // deleted() bodies are neither attracted nor moved; a live test particle (mass 0) is both
            rec[k] = make_double4(x[b], y[b], z[b], gf[b]);
        }
    }
    if (live && h == 0)
        rec[n] = make_double4(0, 0, 0, 0);
    sync();

    auto forces = [&]() {
        double fx[2][2] = { { 0, 0 }, { 0, 0 } }, fy[2][2] = { { 0, 0 }, { 0, 0 } }, fz[2][2] = { { 0, 0 }, { 0, 0 } };
        for (int p = 0; p < n; p += 2) { // p + 1 == n reads the pad
            double4 const q0 = rec[p], q1 = rec[p + 1];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                fast_interact(x[b], y[b], z[b], q0, 0.0, fx[b][0], fy[b][0], fz[b][0]);
                fast_interact(x[b], y[b], z[b], q1, 0.0, fx[b][1], fy[b][1], fz[b][1]);
            }
        }
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (on[b]) {
                ax[b] = fx[b][0] + fx[b][1];
                ay[b] = fy[b][0] + fy[b][1];
                az[b] = fz[b][0] + fz[b][1];
            }
    };
    forces();
    double const hm = A.h * A.mul, dtm = A.dt * A.mul;
    int const to = A.approach_to;

    for (int st = 0; st < A.steps; st++) {
        if (A.approaches) { // Object::update_closest_approaches, at the step's start
            double4 const t = rec[to];
#pragma unroll
            for (int b = 0; b < 2; b++)
                if (own[b] && kb[b] != to)
                    ap[b].see(x[b], y[b], z[b], t.x, t.y, t.z, md[b]);
        }
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (on[b]) {
                vx[b] = __dadd_rn(vx[b], __dmul_rn(ax[b], hm));
                vy[b] = __dadd_rn(vy[b], __dmul_rn(ay[b], hm));
                vz[b] = __dadd_rn(vz[b], __dmul_rn(az[b], hm));
                x[b] = __dadd_rn(x[b], __dmul_rn(vx[b], dtm));
                y[b] = __dadd_rn(y[b], __dmul_rn(vy[b], dtm));
                z[b] = __dadd_rn(z[b], __dmul_rn(vz[b], dtm));
            }
        sync();
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (on[b])
                rec[kb[b]] = make_double4(x[b], y[b], z[b], gf[b]);
        sync();
        forces();
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (on[b]) {
                vx[b] = __dadd_rn(vx[b], __dmul_rn(ax[b], hm));
                vy[b] = __dadd_rn(vy[b], __dmul_rn(ay[b], hm));
                vz[b] = __dadd_rn(vz[b], __dmul_rn(az[b], hm));
            }
        if (A.trail) {
#pragma unroll
            for (int b = 0; b < 2; b++)
                if (own[b] && kb[b] == to)
                    ens_trail(A, copy, st, x[b], y[b], z[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 2; b++)
        if (own[b]) {
            int const k = kb[b];
            if (A.approaches && k != to)
                ap[b].settle(md[b], A.apos ? A.apos + ((size_t)copy * n + k) * 6 : nullptr);
            s[k] = x[b];
            s[(size_t)n + k] = y[b];
            s[2 * (size_t)n + k] = z[b];
            s[3 * (size_t)n + k] = vx[b];
            s[4 * (size_t)n + k] = vy[b];
            s[5 * (size_t)n + k] = vz[b];
            A.mind[(size_t)copy * n + k] = md[b];
        }
}

// Fast-arithmetic ensembles of SMALL templates (up to kEnsembleSoloMax bodies; the Solar System has 10): ONE THREAD PER COPY.
// A thread holds the positions and accelerations of its whole copy in registers (N is a compile-time constant, every loop is
// unrolled), so a force pass evaluates each PAIR once -- Newton's third law, as the pair-shared kernel K1p does for large worlds:
// 3 DADD + 3 (r^2) + 6 (r^-3 from the MUFU seed) + 2 (the two weights) + 6 DFMA = 20 FP64 instructions per pair, 10 per
// directed interaction, where the two-bodies-per-thread kernel below spends 16 -- and no shared-memory traffic, no barrier, no
// shuffle in the loop at all: the N (N - 1) / 2 pairs of a step are independent chains for the FP64 pipe.  Weights gf_j are the
// same in every copy and arrive as kernel parameters (constant-bank operands of the multiplies).  Velocities are touched twice
// per step only and live in shared memory (component c of thread t at [c][t]: conflict free), as do the closest-approach
// records of the candidate service (running minimum of d^2 and the six coordinates it was seen at, per body).
constexpr int kSoloThreads = 32; // per CTA: small CTAs spread an ensemble of 65,536 copies evenly over 148 SMs in ONE wave

// Every pair (i, j) of N bodies once, expanded at compile time, in round-robin tournament order (circle method): the pairs of a
// round share no body, so neighbouring pairs in program order are independent chains (measured on B200, Solar System, 1 M copies:
// 1.42e10 copy-iterations/s against 1.34e10 in row-major order).  M = N rounded up to even; index M - 1 is the bye of an odd N.
template<int N, int R, int K, class F>
__device__ __forceinline__ void for_each_pair_rr(F&& f) {
    constexpr int M = N + (N & 1);
    if constexpr (R < M - 1) {
        constexpr int a = K == 0 ? M - 1 : (R + K) % (M - 1);
        constexpr int b = (R + M - 1 - K) % (M - 1);
        constexpr int lo = a < b ? a : b, hi = a < b ? b : a;
        if constexpr (hi < N)
            f(std::integral_constant<int, lo> {}, std::integral_constant<int, hi> {});
        if constexpr (K + 1 < M / 2)
            for_each_pair_rr<N, R, K + 1>(f);
        else
            for_each_pair_rr<N, R + 1, 0>(f);
    }
}

template<int N>
struct SoloConst {
    double gf[N];
    unsigned on; // bit k: body k is live (Object::deleted() masks the others: neither attracted nor moved)
};

template<int N, int MINB>
__global__ void __launch_bounds__(kSoloThreads, MINB) k_ensemble_solo(EnsembleArgs const A, SoloConst<N> const G) {
    extern __shared__ double solo_smem[];
    constexpr int T = kSoloThreads;
    int const tid = threadIdx.x;
    long long const mine = (long long)blockIdx.x * T + tid;
    bool const live = mine < A.copies;
    long long const copy = live ? mine : 0; // a surplus lane repeats copy 0 and stores nothing
    double* const sv = solo_smem + tid;     // velocity component c of body k at sv[(3 k + c) T]
    double* const sm2 = sv + 3 * N * T;     // closest approaches: minimum of d^2 of body k in this launch (< 0: none) at sm2[k T]
    double* const sp = sm2 + N * T;         // ... and the coordinates it was seen at, sp[(6 k + q) T]
    double* const s = A.state + (size_t)copy * 6 * N;
    double x[N], y[N], z[N], ax[N], ay[N], az[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        x[k] = s[k];
        y[k] = s[N + k];
        z[k] = s[2 * N + k];
        sv[(3 * k) * T] = s[3 * N + k];
        sv[(3 * k + 1) * T] = s[4 * N + k];
        sv[(3 * k + 2) * T] = s[5 * N + k];
    }
    bool const approaches = A.approaches != 0;
    bool const keep_pos = approaches && A.apos != nullptr;
    int const to = A.approach_to;
    if (approaches) {
#pragma unroll
        for (int k = 0; k < N; k++)
            sm2[k * T] = -1.0;
    }

    auto forces = [&]() {
#pragma unroll
        for (int k = 0; k < N; k++)
            ax[k] = ay[k] = az[k] = 0.0;
        // (template recursion, not `#pragma unroll`: a nest of this size is left partly rolled, and the state it indexes then
        // lives in local memory -- measured: 3.5e9 copy-iterations/s instead of 1.1e10)
        for_each_pair_rr<N, 0, 0>([&](auto I, auto J) {
            constexpr int i = decltype(I)::value, j = decltype(J)::value;
            double const dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
            double r2 = dx * dx;
            r2 = fma(dy, dy, r2);
            r2 = fma(dz, dz, r2);
            // |d|^-3 = y0^3 (1 + e (3/2 + 15/8 e)),  e = 1 - r2 y0^2  (common.cuh inv_r3_times); 0 for coincident bodies
            double const y0 = rsqrt_seed(r2);
            double const y2 = y0 * y0;
            double const e = fma(-r2, y2, 1.0);
            double const y3 = y0 * y2;
            double const we = fma(1.875, e, 1.5) * e;
            double const c = fma(y3, we, y3);
            double const si = G.gf[j] * c, sj = G.gf[i] * c;
            ax[i] = fma(si, dx, ax[i]);
            ay[i] = fma(si, dy, ay[i]);
            az[i] = fma(si, dz, az[i]);
            ax[j] = fma(-sj, dx, ax[j]);
            ay[j] = fma(-sj, dy, ay[j]);
            az[j] = fma(-sj, dz, az[j]);
        });
    };
    // (a h) mul == a (h mul) and (v dt) mul == v (dt mul) exactly: mul is +-1
    double const hm = A.h * A.mul, dtm = A.dt * A.mul;
    int until = (int)(A.trail_every - A.steps_done % A.trail_every); // steps until the designated body's next trail sample
    long long tidx = A.steps_done / A.trail_every;                   // ... which is sample number tidx

    // st == -1 is the force pass in front of the first step.  (ONE call site for the force pass: a second one makes the
    // compiler keep it out of line, and the state it works on would live in local memory instead of registers.)
    for (int st = -1; st < A.steps; st++) {
        if (st < 0)
            goto pass;
        if (approaches) { // Object::update_closest_approaches (src/Object.cpp:405-417), at the step's start
            double tx = 0, ty = 0, tz = 0;
#pragma unroll
            for (int k = 0; k < N; k++) {
                tx = k == to ? x[k] : tx;
                ty = k == to ? y[k] : ty;
                tz = k == to ? z[k] : tz;
            }
#pragma unroll
            for (int k = 0; k < N; k++) {
                if (k == to)
                    continue;
                double const dx = __dsub_rn(tx, x[k]), dy = __dsub_rn(ty, y[k]), dz = __dsub_rn(tz, z[k]);
                double const d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                double const m2 = sm2[k * T];
                // a distance of exactly zero stores 0 == "unset" in the reference: the search restarts
                bool const zero = d2 == 0.0;
                bool const better = !zero && (m2 < 0.0 || d2 < m2);
                sm2[k * T] = zero ? -1.0 : (better ? d2 : m2); // unconditional: the copies of a warp take one path
                if (zero && live)
                    A.mind[(size_t)copy * N + k] = 0.0;
                if (better && keep_pos) {
                    sp[(6 * k) * T] = x[k];
                    sp[(6 * k + 1) * T] = y[k];
                    sp[(6 * k + 2) * T] = z[k];
                    sp[(6 * k + 3) * T] = tx;
                    sp[(6 * k + 4) * T] = ty;
                    sp[(6 * k + 5) * T] = tz;
                }
            }
        }
        // the second kick of the previous step and the first of this one use the same acceleration: one pass over the velocities
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (!((G.on >> k) & 1u))
                continue;
            double const kx = __dmul_rn(ax[k], hm), ky = __dmul_rn(ay[k], hm), kz = __dmul_rn(az[k], hm);
            double vx = sv[(3 * k) * T], vy = sv[(3 * k + 1) * T], vz = sv[(3 * k + 2) * T];
            if (st > 0) {
                vx = __dadd_rn(vx, kx);
                vy = __dadd_rn(vy, ky);
                vz = __dadd_rn(vz, kz);
            }
            vx = __dadd_rn(vx, kx);
            vy = __dadd_rn(vy, ky);
            vz = __dadd_rn(vz, kz);
            x[k] = __dadd_rn(x[k], __dmul_rn(vx, dtm));
            y[k] = __dadd_rn(y[k], __dmul_rn(vy, dtm));
            z[k] = __dadd_rn(z[k], __dmul_rn(vz, dtm));
            sv[(3 * k) * T] = vx;
            sv[(3 * k + 1) * T] = vy;
            sv[(3 * k + 2) * T] = vz;
        }
    pass:
        forces();
        if (st >= 0 && A.trail && --until == 0) { // the designated body's position after every trail_every-th step
            until = A.trail_every;
            if (live && tidx < A.trail_cap) {
                double tx = 0, ty = 0, tz = 0;
#pragma unroll
                for (int k = 0; k < N; k++) {
                    tx = k == to ? x[k] : tx;
                    ty = k == to ? y[k] : ty;
                    tz = k == to ? z[k] : tz;
                }
                double* t = A.trail + ((size_t)copy * A.trail_cap + (size_t)tidx) * 3;
                t[0] = tx;
                t[1] = ty;
                t[2] = tz;
                A.trail_len[copy] = (int)tidx + 1;
            }
            tidx++;
        }
    }
    if (!live)
        return;
#pragma unroll
    for (int k = 0; k < N; k++) {
        double vx = sv[(3 * k) * T], vy = sv[(3 * k + 1) * T], vz = sv[(3 * k + 2) * T];
        if (A.steps > 0 && ((G.on >> k) & 1u)) { // the last step's second kick
            vx = __dadd_rn(vx, __dmul_rn(ax[k], hm));
            vy = __dadd_rn(vy, __dmul_rn(ay[k], hm));
            vz = __dadd_rn(vz, __dmul_rn(az[k], hm));
        }
        s[k] = x[k];
        s[N + k] = y[k];
        s[2 * N + k] = z[k];
        s[3 * N + k] = vx;
        s[4 * N + k] = vy;
        s[5 * N + k] = vz;
        if (approaches && k != to) { // RN o sqrt is monotone: the root of the launch's minimum, taken once (Approach::settle)
            double const m2 = sm2[k * T];
            if (m2 >= 0.0) {
                double const d = __dsqrt_rn(m2);
                double const md = A.mind[(size_t)copy * N + k];
                if (md == 0.0 || d < md) {
                    A.mind[(size_t)copy * N + k] = d;
                    if (keep_pos)
                        for (int q = 0; q < 6; q++)
                            A.apos[((size_t)copy * N + k) * 6 + q] = sp[(6 * k + q) * T];
                }
            }
        }
    }
}

template<int N>
static cudaError_t launch_solo(essa_ctx* c, EnsembleArgs const& A, cudaStream_t st) {
    SoloConst<N> G;
    for (int k = 0; k < N; k++)
        G.gf[k] = c->ens_gf_host[k];
    G.on = c->ens_on_mask;
    bool const approaches = A.approaches != 0;
    size_t const smem = (size_t)kSoloThreads * sizeof(double) * (3 * N + (approaches ? N + (A.apos ? 6 * N : 0) : 0));
    int const blocks = (int)(((long long)A.copies + kSoloThreads - 1) / kSoloThreads);
    // Registers: 6 N doubles of state + the pairs in flight: 8 warps per SM, two per scheduler, whose unrolled pairs keep the
    // FP64 pipe fed.  (A 65,536-copy ensemble is 2,048 warps on 592 schedulers: 3.46 each, so 4 on the busiest whichever way
    // they are packed -- two waves of 8 warps per SM cost what one wave of 14 would, without its register spills.  Also measured
    // (round 2) and rejected: a persistent grid of 1,184 CTAs walking (quarter of the steps, CTA) items part-major, so that the
    // tail of one part's last wave is filled with the next part's first items (8,192 items = 6.92 per place): 1.175e10 against
    // 1.189e10 copy-iterations/s -- the second wave's warps have the FP64 pipes more to themselves and finish sooner anyway.)
#ifndef SOLO_MINB
#define SOLO_MINB 8
#endif
    constexpr int MINB = SOLO_MINB;
    auto kern = k_ensemble_solo<N, MINB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    kern<<<blocks, kSoloThreads, smem, st>>>(A, G);
    return cudaGetLastError();
}

static cudaError_t launch_solo_n(essa_ctx* c, EnsembleArgs const& A, cudaStream_t st) {
    switch (A.n) {
#define ESSA_SOLO_CASE(N) \
    case N: \
        return launch_solo<N>(c, A, st);
        ESSA_SOLO_CASE(2) ESSA_SOLO_CASE(3) ESSA_SOLO_CASE(4) ESSA_SOLO_CASE(5) ESSA_SOLO_CASE(6) ESSA_SOLO_CASE(7) ESSA_SOLO_CASE(8)
        ESSA_SOLO_CASE(9) ESSA_SOLO_CASE(10) ESSA_SOLO_CASE(11) ESSA_SOLO_CASE(12) ESSA_SOLO_CASE(13) ESSA_SOLO_CASE(14)
        ESSA_SOLO_CASE(15) ESSA_SOLO_CASE(16)
#undef ESSA_SOLO_CASE
    default:
        return cudaErrorInvalidConfiguration;
    }
}

// Fast-arithmetic ensembles of templates of 17 .. 160 bodies (jupiter_system.essa has 77): a copy lives on P = 2 .. 32 lanes of
// ONE warp (P a power of two: 32 / P whole copies per warp, every lane busy), each lane the HOME of HB = ceil(n / P) <= 5 bodies
// whose positions, weights and accelerations stay in its registers.  Pairs are shared as in K1p: lane k evaluates its home
// bodies among themselves, then receives the home of lane k + d (d = 1 .. P / 2, the ring of a copy's lanes) as VISITORS from
// shared memory, evaluates all HB x HB pairs once -- both accelerations -- and adds the visitors' share to a reaction array in
// shared memory: in round d every lane's reaction slots have exactly one writer, so there is no atomic and no conflict, only a
// __syncwarp per round.  (Round P / 2 pairs lane k with lane k + P / 2 both ways round; the lower half of the ring does it.)
// 20 FP64 instructions per PAIR where the two-bodies-per-thread kernel spends 16 per directed interaction; padding costs
// (P HB / n)^2 -- 80 slots for 77 bodies -- and the half-idle last round 1 / (P - 1) of the visits.  Shared memory holds one
// double per (component, home slot, lane): consecutive lanes touch consecutive words, home and visitor accesses alike.
constexpr int kRingWarps = 4;
constexpr int kRingMaxHB = 5;
constexpr int kRingComp = 10;          // x y z gf | reaction x y z | vx vy vz
constexpr int kRingCompApproach = 7;   // + minimum d^2 of the launch and the six coordinates it was seen at

template<int HB>
__global__ void __launch_bounds__(32 * kRingWarps, 3) k_ensemble_ring(EnsembleArgs const A, int const P) {
    extern __shared__ double ring_smem[];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const k = lane & (P - 1), cbase = lane - k; // place in the copy's ring; first lane of the copy
    int const per_warp = 32 / P;
    long long const mine = ((long long)blockIdx.x * kRingWarps + warp) * per_warp + (lane / P);
    bool const live = mine < A.copies;
    long long const copy = live ? mine : 0; // a surplus ring repeats copy 0 and stores nothing
    int const n = A.n;
    bool const approaches = A.approaches != 0;
    bool const keep_pos = approaches && A.apos != nullptr;
    int const comps = kRingComp + (approaches ? kRingCompApproach : 0);
    double* const W = ring_smem + (size_t)warp * comps * HB * 32;
    auto at = [&](int comp, int b, int l) -> double& { return W[(comp * HB + b) * 32 + l]; };

    double hx[HB], hy[HB], hz[HB], hg[HB], ax[HB], ay[HB], az[HB];
    bool real[HB], on[HB];
    double* const s = A.state + (size_t)copy * 6 * n;
#pragma unroll
    for (int b = 0; b < HB; b++) {
        int const idx = k * HB + b;
        real[b] = idx < n;
        on[b] = real[b] && A.active[idx] != 0; // a body that deleted() masks is neither attracted nor moved
        // padding: weight 0, parked far away and apart from each other (no r = 0 among the dummies)
        hx[b] = real[b] ? s[idx] : 1e30 * (double)(idx + 1);
        hy[b] = real[b] ? s[(size_t)n + idx] : 0.0;
        hz[b] = real[b] ? s[2 * (size_t)n + idx] : 0.0;
        hg[b] = real[b] ? A.gf[idx] : 0.0;
        at(3, b, lane) = hg[b];
        at(7, b, lane) = real[b] ? s[3 * (size_t)n + idx] : 0.0;
        at(8, b, lane) = real[b] ? s[4 * (size_t)n + idx] : 0.0;
        at(9, b, lane) = real[b] ? s[5 * (size_t)n + idx] : 0.0;
        if (approaches)
            at(kRingComp, b, lane) = -1.0;
        ax[b] = ay[b] = az[b] = 0.0;
    }
    // one pair, both accelerations (the r = 0 guard of rsqrt_seed stays: two bodies of a template may coincide)
    auto pair = [&](double xi, double yi, double zi, double gi, double xj, double yj, double zj, double gj, double& aix, double& aiy,
                    double& aiz, double& ajx, double& ajy, double& ajz) {
        double const dx = xj - xi, dy = yj - yi, dz = zj - zi;
        double r2 = dx * dx;
        r2 = fma(dy, dy, r2);
        r2 = fma(dz, dz, r2);
        double const y0 = rsqrt_seed(r2);
        double const y2 = y0 * y0;
        double const e = fma(-r2, y2, 1.0);
        double const y3 = y0 * y2;
        double const we = fma(1.875, e, 1.5) * e;
        double const u = fma(y3, we, y3);
        double const si = gj * u, sj = gi * u;
        aix = fma(si, dx, aix);
        aiy = fma(si, dy, aiy);
        aiz = fma(si, dz, aiz);
        ajx = fma(-sj, dx, ajx);
        ajy = fma(-sj, dy, ajy);
        ajz = fma(-sj, dz, ajz);
    };
    int const half_ring = P >> 1;
    auto forces = [&]() {
#pragma unroll
        for (int b = 0; b < HB; b++) {
            at(0, b, lane) = hx[b];
            at(1, b, lane) = hy[b];
            at(2, b, lane) = hz[b];
            at(4, b, lane) = 0.0;
            at(5, b, lane) = 0.0;
            at(6, b, lane) = 0.0;
            ax[b] = ay[b] = az[b] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < HB; b++)
#pragma unroll
            for (int c = b + 1; c < HB; c++)
                pair(hx[b], hy[b], hz[b], hg[b], hx[c], hy[c], hz[c], hg[c], ax[b], ay[b], az[b], ax[c], ay[c], az[c]);
        for (int d = 1; d <= half_ring; d++) {
            int const m = cbase + ((k + d) & (P - 1));
            bool const act = d < half_ring || k < half_ring;
            if (act) {
                double vx[HB], vy[HB], vz[HB], vg[HB], rx[HB], ry[HB], rz[HB];
#pragma unroll
                for (int b = 0; b < HB; b++) {
                    vx[b] = at(0, b, m);
                    vy[b] = at(1, b, m);
                    vz[b] = at(2, b, m);
                    vg[b] = at(3, b, m);
                    rx[b] = ry[b] = rz[b] = 0.0;
                }
#pragma unroll
                for (int b = 0; b < HB; b++)
#pragma unroll
                    for (int c = 0; c < HB; c++)
                        pair(hx[b], hy[b], hz[b], hg[b], vx[c], vy[c], vz[c], vg[c], ax[b], ay[b], az[b], rx[c], ry[c], rz[c]);
#pragma unroll
                for (int b = 0; b < HB; b++) {
                    at(4, b, m) += rx[b];
                    at(5, b, m) += ry[b];
                    at(6, b, m) += rz[b];
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int b = 0; b < HB; b++) {
            ax[b] += at(4, b, lane);
            ay[b] += at(5, b, lane);
            az[b] += at(6, b, lane);
        }
    };
    double const hm = A.h * A.mul, dtm = A.dt * A.mul;
    int const to = A.approach_to;
    int const to_lane = cbase + to / HB, to_slot = to - (to / HB) * HB;
    int until = (int)(A.trail_every - A.steps_done % A.trail_every);
    long long tidx = A.steps_done / A.trail_every;

    // st == -1 is the force pass in front of the first step (one call site, as in k_ensemble_solo)
    for (int st = -1; st < A.steps; st++) {
        if (st < 0)
            goto pass;
        if (approaches) { // Object::update_closest_approaches, at the step's start: shared memory holds the current positions
            double tx = 0, ty = 0, tz = 0;
#pragma unroll
            for (int b = 0; b < HB; b++) {
                tx = b == to_slot ? at(0, b, to_lane) : tx;
                ty = b == to_slot ? at(1, b, to_lane) : ty;
                tz = b == to_slot ? at(2, b, to_lane) : tz;
            }
#pragma unroll
            for (int b = 0; b < HB; b++) {
                int const idx = k * HB + b;
                if (!real[b] || idx == to)
                    continue;
                double const dx = __dsub_rn(tx, hx[b]), dy = __dsub_rn(ty, hy[b]), dz = __dsub_rn(tz, hz[b]);
                double const d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                double const m2 = at(kRingComp, b, lane);
                bool const zero = d2 == 0.0; // stores 0 == "unset" in the reference: the search restarts
                bool const better = !zero && (m2 < 0.0 || d2 < m2);
                at(kRingComp, b, lane) = zero ? -1.0 : (better ? d2 : m2);
                if (zero && live)
                    A.mind[(size_t)copy * n + idx] = 0.0;
                if (better && keep_pos) {
                    at(kRingComp + 1, b, lane) = hx[b];
                    at(kRingComp + 2, b, lane) = hy[b];
                    at(kRingComp + 3, b, lane) = hz[b];
                    at(kRingComp + 4, b, lane) = tx;
                    at(kRingComp + 5, b, lane) = ty;
                    at(kRingComp + 6, b, lane) = tz;
                }
            }
            __syncwarp(); // the drift below must not overtake a neighbour's read of the designated body
        }
        // the second kick of the previous step and the first of this one use the same acceleration
#pragma unroll
        for (int b = 0; b < HB; b++) {
            if (!on[b])
                continue;
            double const kx = __dmul_rn(ax[b], hm), ky = __dmul_rn(ay[b], hm), kz = __dmul_rn(az[b], hm);
            double vx = at(7, b, lane), vy = at(8, b, lane), vz = at(9, b, lane);
            if (st > 0) {
                vx = __dadd_rn(vx, kx);
                vy = __dadd_rn(vy, ky);
                vz = __dadd_rn(vz, kz);
            }
            vx = __dadd_rn(vx, kx);
            vy = __dadd_rn(vy, ky);
            vz = __dadd_rn(vz, kz);
            hx[b] = __dadd_rn(hx[b], __dmul_rn(vx, dtm));
            hy[b] = __dadd_rn(hy[b], __dmul_rn(vy, dtm));
            hz[b] = __dadd_rn(hz[b], __dmul_rn(vz, dtm));
            at(7, b, lane) = vx;
            at(8, b, lane) = vy;
            at(9, b, lane) = vz;
        }
    pass:
        forces();
        if (st >= 0 && A.trail && --until == 0) {
            until = A.trail_every;
            if (live && lane == to_lane && tidx < A.trail_cap) {
                double tx = 0, ty = 0, tz = 0;
#pragma unroll
                for (int b = 0; b < HB; b++) {
                    tx = b == to_slot ? hx[b] : tx;
                    ty = b == to_slot ? hy[b] : ty;
                    tz = b == to_slot ? hz[b] : tz;
                }
                double* t = A.trail + ((size_t)copy * A.trail_cap + (size_t)tidx) * 3;
                t[0] = tx;
                t[1] = ty;
                t[2] = tz;
                A.trail_len[copy] = (int)tidx + 1;
            }
            tidx++;
        }
    }
    if (!live)
        return;
#pragma unroll
    for (int b = 0; b < HB; b++) {
        int const idx = k * HB + b;
        if (!real[b])
            continue;
        double vx = at(7, b, lane), vy = at(8, b, lane), vz = at(9, b, lane);
        if (A.steps > 0 && on[b]) { // the last step's second kick
            vx = __dadd_rn(vx, __dmul_rn(ax[b], hm));
            vy = __dadd_rn(vy, __dmul_rn(ay[b], hm));
            vz = __dadd_rn(vz, __dmul_rn(az[b], hm));
        }
        s[idx] = hx[b];
        s[(size_t)n + idx] = hy[b];
        s[2 * (size_t)n + idx] = hz[b];
        s[3 * (size_t)n + idx] = vx;
        s[4 * (size_t)n + idx] = vy;
        s[5 * (size_t)n + idx] = vz;
        if (approaches && idx != to) {
            double const m2 = at(kRingComp, b, lane);
            if (m2 >= 0.0) {
                double const d = __dsqrt_rn(m2);
                double const md = A.mind[(size_t)copy * n + idx];
                if (md == 0.0 || d < md) {
                    A.mind[(size_t)copy * n + idx] = d;
                    if (keep_pos)
                        for (int q = 0; q < 6; q++)
                            A.apos[((size_t)copy * n + idx) * 6 + q] = at(kRingComp + 1 + q, b, lane);
                }
            }
        }
    }
}

// lanes per copy and home bodies per lane for a template of n bodies: the smallest ring whose homes hold at most kRingMaxHB
static bool ring_shape(int n, int& P, int& HB) {
    for (P = 2; P <= 32; P <<= 1) {
        HB = (n + P - 1) / P;
        if (HB <= kRingMaxHB) {
            HB = HB < 3 ? 3 : HB;
            return true;
        }
    }
    return false;
}

template<int HB>
static cudaError_t launch_ring(EnsembleArgs const& A, int P, cudaStream_t st) {
    int const per_warp = 32 / P;
    long long const warps = ((long long)A.copies + per_warp - 1) / per_warp;
    int const blocks = (int)((warps + kRingWarps - 1) / kRingWarps);
    int const comps = kRingComp + (A.approaches ? kRingCompApproach : 0);
    size_t const smem = (size_t)kRingWarps * comps * HB * 32 * sizeof(double);
    auto kern = k_ensemble_ring<HB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    kern<<<blocks, 32 * kRingWarps, smem, st>>>(A, P);
    return cudaGetLastError();
}

cudaError_t launch_ensemble_init(essa_ctx* c, cudaStream_t st) {
    long long total = (long long)c->ens_copies * c->ens_n;
    int blocks = (int)((total + 255) / 256);
    k_ensemble_init<<<blocks, 256, 0, st>>>(nullptr, c->vel, c->gf, c->pos[c->cur], c->active, c->cap, c->ens_n, c->ens_copies,
        (long long)c->ens.first_copy, (unsigned long long)c->ens.seed, c->ens.rel_sigma, c->ens_state, c->ens_gf, c->ens_active, c->ens_mind);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_ensemble_step(essa_ctx* c, int steps, cudaStream_t st) {
    int n = c->ens_n;
    EnsembleArgs A;
    A.state = c->ens_state;
    A.gf = c->ens_gf;
    A.active = c->ens_active;
    A.mind = c->ens_mind;
    A.apos = c->ens_apos;
    A.trail = c->ens_trail;
    A.trail_len = c->ens_trail_len;
    A.trail_every = c->ens.trail_every > 0 ? c->ens.trail_every : 1;
    A.trail_cap = c->ens.trail_capacity;
    A.steps_done = c->ens_steps_done;
    A.n = n;
    A.copies = c->ens_copies;
    A.steps = steps < 0 ? -steps : steps;
    A.mul = steps >= 0 ? 1.0 : -1.0;
    A.dt = (double)c->ens.dt_seconds;
    A.h = A.dt / 2.0;
    A.approaches = c->ens.closest_approaches;
    A.approach_to = c->ens.approach_to;
    cudaError_t e;
    if (c->ens.exact_order) {
        int per_cta = n <= kEnsThreadsMax ? kEnsThreadsMax / n : 1;
        if (per_cta > c->ens_copies)
            per_cta = c->ens_copies;
        A.per_cta = per_cta;
        int threads = (per_cta * n + 31) / 32 * 32;
        if (threads > kEnsThreadsMax) // a template wider than one ensemble CTA
            return cudaErrorInvalidConfiguration;
        int blocks = (c->ens_copies + per_cta - 1) / per_cta;
        k_ensemble_exact<<<blocks, threads, (size_t)per_cta * 4 * n * sizeof(double), st>>>(A);
    } else if (n >= 2 && n <= kEnsembleSoloMax && c->ens_solo) {
        A.per_cta = kSoloThreads;
        e = launch_solo_n(c, A, st);
        if (e != cudaSuccess)
            return e;
    } else if (int P = 0, HB = 0; c->ens_solo && n > kEnsembleSoloMax && ring_shape(n, P, HB)) {
        A.per_cta = kRingWarps * (32 / P);
        e = HB == 3 ? launch_ring<3>(A, P, st) : HB == 4 ? launch_ring<4>(A, P, st) : launch_ring<5>(A, P, st);
        if (e != cudaSuccess)
            return e;
    } else {
        int const half = (n + 1) / 2;
        size_t const rec_bytes = (size_t)(n + 1) * sizeof(double4);
        if (half <= 32) { // up to 64 bodies: whole copies per warp
            int const cw = 32 / half;
            long long const warps = ((long long)c->ens_copies + cw - 1) / cw;
            int const wblocks = (int)((warps + kDuoWarps - 1) / kDuoWarps);
            A.per_cta = cw;
            k_ensemble_duo<true><<<wblocks, 32 * kDuoWarps, (size_t)kDuoWarps * cw * rec_bytes, st>>>(A);
        } else {
            A.per_cta = kDuoCtaThreads / half;
            if (A.per_cta < 1)
                return cudaErrorInvalidConfiguration;
            int const dblocks = (c->ens_copies + A.per_cta - 1) / A.per_cta;
            k_ensemble_duo<false><<<dblocks, kDuoCtaThreads, (size_t)A.per_cta * rec_bytes, st>>>(A);
        }
    }
    e = cudaGetLastError();
    c->launches++;
    int P = A.steps + 1;
    c->passes += P;
    c->interactions += (double)P * (double)c->ens_copies * (double)n * (double)(n - 1);
    c->last_force_passes = P;
    c->ens_steps_done += A.steps;
    return e;
}

} // namespace essa
