// K1g -- mid-size worlds (1,025 .. 5,120 bodies) stepped by ALL SMs in ONE cooperative launch per call.
//
// Replaces, for these sizes, the tiled path's one launch (K1) per World::update step (src/World.cpp:86-139): at 2,048 bodies a
// step is 3.6 us of FP64 arithmetic on 148 SMs, but a K1 launch costs 8 us from launch to launch before it computes anything.
//
// Owner computes.  CTA c owns T = ceil(n / G) consecutive bodies for the whole call and keeps their state in registers; every
// CTA holds the positions of ALL bodies in shared memory (32 B per body: 64 KB at 2,048 bodies).  A step:
//   1. owners: first half kick + drift (src/World.cpp:106-116), new record {x, y, z, gf} to the OTHER global position buffer
//   2. ONE grid barrier (release add + acquire spin on a counter in global memory; co-residency by cooperative launch)
//   3. every CTA reloads all records (L2 only: ld.global.cg), then L = floor(512 / T) threads per owned body walk the sources
//      j = l, l + L, ... (fast arithmetic of common.cuh, 16 FP64 instructions per interaction), partial sums meet in shared
//      memory and are added in ascending l by the body's owner threads -- a fixed order: deterministic, no atomics
//   4. owners: second half kick (src/World.cpp:121-127)
// No partial-sum exchange between SMs, no second barrier: the double-buffered positions make one barrier per step enough
// (buffer B is read between barriers s-1 and s, and written again only after barrier s).
// The state of a body is spread over four owner threads: thread (t, c), c = 0..2, owns component c (x_c, v_c, a_c: kick and
// drift never mix components), thread (t, 3) the body's gravity factor, active flag and most-attracting link
// (src/Object.cpp:84-94: strictly heavier partner, largest gf / r^4, lowest index on ties -- as K1's ARGMAX variant).
//
// Measured and rejected: a barrier-free exchange in which the positions travel as self-validating 64-bit words {step tag, 32 bits
// of payload} that every CTA polls (what the persistent kernels do towards the host): bit-identical results, but 48 instead of
// 32 bytes per body in 8-byte loads and the re-polls of 148 x 512 threads cost more than the barrier saves -- 2,048 bodies 9.1
// vs 8.1 us per step, 4,096: 25.9 vs 23.1; only 400 bodies gained (2.9 vs 3.1).
//
// Not taken (falls to K1): recording (History / Trail of a tracked prefix), two_pass, a call during which the active mask
// flips, sharded worlds.
#include <cstdlib>

#include "common.cuh"
#include "ctx.hpp"

namespace essa {

constexpr int kGridThreads = 512;

struct GridArgs {
    WorldPtrs w;      // pos_cur: the records at the start of the call
    int n, T, L, R;   // bodies per CTA, lanes per body, lanes of the first reduction stage (4 <= R <= L)
    int steps;        // |steps| of the call
    int first_pass;   // accelerations at the starting positions are not valid: evaluate them first (the step's first set_forces)
    int argmax;
    double h, dt, mul, eps2;
    unsigned* bar;    // arrivals (monotone over the life of the context)
    unsigned* timed_out; // mapped host memory: set if a barrier gave up (the host checks it after the call, for free)
    unsigned bar_base; // value of bar[0] when this launch starts
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target, unsigned* timed_out) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        long long const t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int)(v - target) >= 0)
                break;
            if (clock64() - t0 > 4000000000ll) { // ~2 s: a CTA of the grid is gone; do not hang the device
                *(volatile unsigned*)timed_out = 1u;
                break;
            }
        }
    }
    __syncthreads();
}

template<bool ARGMAX>
__global__ void __launch_bounds__(kGridThreads, 1) k_force_grid(GridArgs const A) {
    extern __shared__ __align__(16) unsigned char grid_smem[];
    double4* spos = reinterpret_cast<double4*>(grid_smem);                     // [n]
    double* red = reinterpret_cast<double*>(grid_smem + (size_t)A.n * sizeof(double4)); // [3][kGridThreads]
    double* redm = red + 3 * kGridThreads;                                     // [kGridThreads] (ARGMAX)
    double* red2 = redm + kGridThreads;                                        // [3][kGridThreads] second stage of the sums
    double* redm2 = red2 + 3 * kGridThreads;                                   // [kGridThreads]
    int* redj = reinterpret_cast<int*>(redm2 + kGridThreads);                  // [kGridThreads] (ARGMAX)
    int* redj2 = redj + kGridThreads;                                          // [kGridThreads]

    WorldPtrs const& W = A.w;
    int const tid = threadIdx.x, T = A.T, L = A.L, n = A.n;
    int const l = tid / T, t = tid - l * T;
    int const g = blockIdx.x * T + t;
    bool const worker = l < L && g < n;       // evaluates partial sums of body g
    bool const owner = l < 4 && g < n;        // (t, 0..2): a component; (t, 3): the link (L >= 12 always)
    auto P = [&](int b) { return b ? W.pos_next : W.pos_cur; }; // the two position buffers (roles at the start of the call)

    // owner state
    bool act = false;
    double xc = 0, vc = 0, ac = 0, gfe = 0; // component threads: position, velocity (integer step / half step), acceleration
    int most = -1, old_most = -1;
    double maxattr = 0.0;
    if (owner) {
        act = W.active[g] != 0;
        double4 const p = W.pos_cur[g];
        if (l < 3) {
            xc = l == 0 ? p.x : (l == 1 ? p.y : p.z);
            vc = W.v[l][g];
            ac = W.a[l][g];
        } else {
            gfe = p.w;
            most = W.most[g];
            old_most = W.old_most[g];
            maxattr = W.maxattr[g];
        }
    }

    int cur = 0;
    unsigned bar_target = A.bar_base;
    bool save_old = true; // Object::before_update precedes the first force evaluation of a step

    // all records of buffer `cur` into shared memory (other SMs wrote them: L2, never L1)
    auto load_all = [&]() {
        double2 const* src = reinterpret_cast<double2 const*>(P(cur));
        double2* dst = reinterpret_cast<double2*>(spos);
        // (a plain loop: the compiler keeps four loads in flight; eight explicit ones measured 0.7 us slower per 64 KB)
        for (int i = tid; i < 2 * n; i += kGridThreads)
            dst[i] = __ldcg(src + i);
        ESSA_CHECK(T * L <= kGridThreads && L >= 4 && A.R >= 4 && A.R <= L && (int)gridDim.x * T >= n);
        __syncthreads();
    };
    // one set_forces(): partial sums by the workers, ordered sums by the owners
    auto force = [&]() {
        double ax = 0, ay = 0, az = 0, best = 0.0;
        int bj = -1;
        if (worker) {
            double4 const p = spos[g];
#pragma unroll 4
            for (int j = l; j < n; j += L) {
                ESSA_CHECK(j >= 0 && j < n && g < n);
                if (ARGMAX)
                    fast_interact_argmax(p.x, p.y, p.z, p.w, spos[j], j, A.eps2, ax, ay, az, best, bj);
                else
                    fast_interact(p.x, p.y, p.z, spos[j], A.eps2, ax, ay, az);
            }
        }
        red[tid] = ax;
        red[kGridThreads + tid] = ay;
        red[2 * kGridThreads + tid] = az;
        if (ARGMAX) {
            redm[tid] = best;
            redj[tid] = bj;
        }
        __syncthreads();
        // The L partials of a body are added in a fixed order in two stages (a serial chain of L additions would cost up to
        // 0.5 us of latency per pass): lane r < R takes partials r, r + R, ... in ascending order, the owners then add the R
        // results in ascending r.  Candidates: largest metric, among equal ones the lowest index (a thread's own sources are
        // ascending and use a strict '>').
        int const R = A.R;
        if (l < R && g < n) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, m = 0.0;
            int mj = -1;
            for (int k = l; k < L; k += R) {
                int const e = k * T + t;
                ESSA_CHECK(e >= 0 && e < kGridThreads);
                s0 = __dadd_rn(s0, red[e]);
                s1 = __dadd_rn(s1, red[kGridThreads + e]);
                s2 = __dadd_rn(s2, red[2 * kGridThreads + e]);
                if (ARGMAX) {
                    double const mk = redm[e];
                    int const jk = redj[e];
                    if (jk >= 0 && (mk > m || (mk == m && (mj < 0 || jk < mj)))) {
                        m = mk;
                        mj = jk;
                    }
                }
            }
            red2[tid] = s0;
            red2[kGridThreads + tid] = s1;
            red2[2 * kGridThreads + tid] = s2;
            if (ARGMAX) {
                redm2[tid] = m;
                redj2[tid] = mj;
            }
        }
        __syncthreads();
        if (owner && l < 3) {
            double s = 0.0;
            for (int k = 0; k < R; k++) {
                ESSA_CHECK(k * T + t < kGridThreads && l < 3);
                s = __dadd_rn(s, red2[l * kGridThreads + k * T + t]);
            }
            if (act)
                ac = s;
        }
        if (ARGMAX && owner && l == 3) {
            double m = 0.0;
            int mj = -1;
            for (int k = 0; k < R; k++) {
                double const mk = redm2[k * T + t];
                int const jk = redj2[k * T + t];
                if (jk >= 0 && (mk > m || (mk == m && (mj < 0 || jk < mj)))) {
                    m = mk;
                    mj = jk;
                }
            }
            if (save_old)
                old_most = most;
            if (act) {
                maxattr = m; // Object::clear_forces, then the best heavier partner
                if (mj >= 0)
                    most = mj;
            }
        }
        save_old = false;
        __syncthreads(); // red is reused by the next pass
    };

    load_all();
    if (A.first_pass)
        force();
    for (int s = 0; s < A.steps; s++) {
        // first half kick + drift
        if (owner) {
            if (l < 3) {
                if (act) {
                    vc = kick1(vc, ac, A.h, A.mul);
                    xc = __dadd_rn(xc, __dmul_rn(__dmul_rn(vc, A.dt), A.mul));
                }
                double* q = reinterpret_cast<double*>(P(cur ^ 1) + g) + l;
                __stcg(q, xc);
            } else {
                __stcg(reinterpret_cast<double*>(P(cur ^ 1) + g) + 3, gfe);
            }
        }
        bar_target += gridDim.x;
        grid_barrier(A.bar, bar_target, A.timed_out);
        cur ^= 1;
        load_all();
        if (s > 0 || !A.first_pass)
            save_old = true;
        force();
        if (owner && l < 3 && act)
            vc = kick1(vc, ac, A.h, A.mul); // second half kick
    }
    if (owner) {
        if (l < 3) {
            if (act) {
                W.a[l][g] = ac;
                W.v[l][g] = vc;
                W.vh[l][g] = vc; // K1 leaves the half-step velocity here; nothing reads it before the next first kick rewrites it
            }
        } else if (ARGMAX) {
            W.old_most[g] = old_most;
            if (act) {
                W.most[g] = most;
                W.maxattr[g] = maxattr;
            }
        }
    }
}

bool force_grid_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    static int allow = -1;
    if (allow < 0) {
        char const* e = getenv("ESSA_GRID");
        allow = e ? atoi(e) != 0 : 1;
    }
    return allow && c->world == 1 && !forces_only && c->n >= 64 && c->n <= kGridMaxBodies && !o.record && !o.two_pass
        && !o.exact_order;
}

// All |steps| steps of a call in one cooperative launch (the caller swaps the position buffers' roles when |steps| is odd).
cudaError_t launch_force_grid(essa_ctx* c, int steps, essa_step_opts const& o, bool first_pass, bool argmax, cudaStream_t st) {
    int const n = c->n, G = c->sm_count;
    GridArgs A;
    A.w = world_ptrs(c);
    A.n = n;
    A.T = (n + G - 1) / G;
    A.L = kGridThreads / A.T;
    A.R = 4;
    while (A.R * A.R < A.L)
        A.R++;
    A.steps = steps < 0 ? -steps : steps;
    A.first_pass = first_pass ? 1 : 0;
    A.argmax = argmax ? 1 : 0;
    A.dt = (double)o.dt_seconds;
    A.h = A.dt / 2.0;
    A.mul = steps >= 0 ? 1.0 : -1.0;
    A.eps2 = o.eps2;
    if (!c->grid_bar) {
        cudaError_t e = cudaMalloc(&c->grid_bar, sizeof(unsigned));
        if (e == cudaSuccess)
            e = cudaMemsetAsync(c->grid_bar, 0, sizeof(unsigned), st);
        if (e == cudaSuccess)
            e = cudaHostAlloc((void**)&c->grid_timed_out, sizeof(unsigned), cudaHostAllocMapped);
        if (e != cudaSuccess)
            return e;
        *c->grid_timed_out = 0u;
        c->grid_bar_count = 0;
    }
    A.timed_out = nullptr;
    {
        cudaError_t e = cudaHostGetDevicePointer((void**)&A.timed_out, c->grid_timed_out, 0);
        if (e != cudaSuccess)
            return e;
    }
    int const blocks = (n + A.T - 1) / A.T; // <= G
    A.bar = c->grid_bar;
    A.bar_base = c->grid_bar_count;
    size_t const smem = (size_t)n * sizeof(double4) + 2 * kGridThreads * (4 * sizeof(double) + sizeof(int));
    void const* fn = argmax ? (void const*)k_force_grid<true> : (void const*)k_force_grid<false>;
    bool& attr = c->attr_grid[argmax ? 1 : 0];
    if (!attr || smem > c->grid_smem[argmax ? 1 : 0]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        attr = true;
        c->grid_smem[argmax ? 1 : 0] = smem;
    }
    void* args[] = { (void*)&A };
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(kGridThreads), args, smem, st);
    if (e == cudaSuccess) {
        c->launches++;
        c->grid_bar_count += (unsigned)blocks * (unsigned)A.steps; // what the device counter will read when this launch is through
    }
    return e;
}

} // namespace essa
