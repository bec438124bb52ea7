// K4 -- ensemble batch: many perturbed copies of one template world advanced in parallel.
//
// Models the reference's forward-simulation use (src/World.cpp:230-238 clone_for_forward_simulation,
// src/essagui/EssaCreateObject.cpp:166-189 update(ticks) on the clone): forward-simulated worlds
// keep the date frozen and record no History; every object tracks its closest approaches
// (src/Object.cpp:405-417).  Instead of one clone stepped inline on the GUI thread, `copies`
// clones are resident in shared memory, several per CTA, for all steps of a launch.
//
// Layout: state[copy][6][n] (x y z vx vy vz), copy-major, so a CTA's copies are contiguous.
// Thread t of a CTA owns (copy = t / n, body = t % n); all copies of a CTA step in lock step
// behind the same two __syncthreads per KDK iteration.
#include "common.cuh"
#include "ctx.hpp"
#include "rows.cuh"

namespace essa {

constexpr int kEnsThreadsMax = kEnsembleMaxBodies;

__global__ void k_ensemble_init(double const* __restrict__ tx, double const* __restrict__ tv, double const* __restrict__ tgf,
    double4 const* __restrict__ tpos, int cap, int n, int copies, long long first_copy, unsigned long long seed, double rel_sigma,
    double* state, double* gf_out, double* mind) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n)
        gf_out[idx] = tpos[idx].w; // gf_eff: bodies masked by Object::deleted() are not cloned -> weight 0
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
    double* mind;
    int n, copies, per_cta, steps;
    double mul, dt, h;
    int approaches, approach_to;
};

template<bool EXACT>
__global__ void __launch_bounds__(kEnsThreadsMax, 2) k_ensemble(EnsembleArgs const A) {
    extern __shared__ double smem[]; // per copy: sx[npad] sy[npad] sz[npad] sg[npad]
    int const n = A.n, npad = (n + 3) & ~3;
    int const t = threadIdx.x;
    int const lc = t / n, k = t - lc * n;
    int const copy = blockIdx.x * A.per_cta + lc;
    bool const live = lc < A.per_cta && copy < A.copies;
    double* sx = smem + (size_t)(lc < A.per_cta ? lc : 0) * 4 * npad;
    double* sy = sx + npad;
    double* sz = sy + npad;
    double* sg = sz + npad;

    double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, gfk = 0, md = 0, dummy_m = 0;
    int dummy_most = -1;
    double* s = A.state + (size_t)(live ? copy : 0) * 6 * n;
    if (live) {
        x = s[k];
        y = s[(size_t)n + k];
        z = s[2 * (size_t)n + k];
        vx = s[3 * (size_t)n + k];
        vy = s[4 * (size_t)n + k];
        vz = s[5 * (size_t)n + k];
        gfk = A.gf[k];
        md = A.mind[(size_t)copy * n + k];
    }
    // zero-weight padding so that the 4-way unrolled row loop may run to npad
    for (int i = t; i < A.per_cta * 4 * npad; i += blockDim.x)
        smem[i] = 0.0;
    __syncthreads();
    if (live) {
        sx[k] = x;
        sy[k] = y;
        sz[k] = z;
        sg[k] = gfk;
    }
    __syncthreads();
    bool const on = live && gfk != 0.0; // a body that deleted() masks is neither attracted nor moved
    if (on)
        row_forces<EXACT, false>(k, n, npad, sx, sy, sz, sg, x, y, z, gfk, 0.0, ax, ay, az, dummy_m, dummy_most);

    for (int st = 0; st < A.steps; st++) {
        if (A.approaches && live && k != A.approach_to) { // Object::update_closest_approaches, at the step's start
            double dx = __dsub_rn(sx[A.approach_to], x), dy = __dsub_rn(sy[A.approach_to], y), dz = __dsub_rn(sz[A.approach_to], z);
            double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (md == 0.0 || d < md)
                md = d;
        }
        if (on) {
            vx = kick1(vx, ax, A.h, A.mul);
            vy = kick1(vy, ay, A.h, A.mul);
            vz = kick1(vz, az, A.h, A.mul);
            x = __dadd_rn(x, __dmul_rn(__dmul_rn(vx, A.dt), A.mul));
            y = __dadd_rn(y, __dmul_rn(__dmul_rn(vy, A.dt), A.mul));
            z = __dadd_rn(z, __dmul_rn(__dmul_rn(vz, A.dt), A.mul));
        }
        __syncthreads();
        if (live) {
            sx[k] = x;
            sy[k] = y;
            sz[k] = z;
        }
        __syncthreads();
        if (on) {
            row_forces<EXACT, false>(k, n, npad, sx, sy, sz, sg, x, y, z, gfk, 0.0, ax, ay, az, dummy_m, dummy_most);
            vx = kick1(vx, ax, A.h, A.mul);
            vy = kick1(vy, ay, A.h, A.mul);
            vz = kick1(vz, az, A.h, A.mul);
        }
    }
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

cudaError_t launch_ensemble_init(essa_ctx* c, cudaStream_t st) {
    long long total = (long long)c->ens_copies * c->ens_n;
    int blocks = (int)((total + 255) / 256);
    k_ensemble_init<<<blocks, 256, 0, st>>>(nullptr, c->vel, c->gf, c->pos[c->cur], c->cap, c->ens_n, c->ens_copies, (long long)c->ens.first_copy,
        (unsigned long long)c->ens.seed, c->ens.rel_sigma, c->ens_state, c->ens_gf, c->ens_mind);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_ensemble_step(essa_ctx* c, int steps, cudaStream_t st) {
    int n = c->ens_n, npad = (n + 3) & ~3;
    EnsembleArgs A;
    A.state = c->ens_state;
    A.gf = c->ens_gf;
    A.mind = c->ens_mind;
    A.n = n;
    A.copies = c->ens_copies;
    A.steps = steps < 0 ? -steps : steps;
    A.mul = steps >= 0 ? 1.0 : -1.0;
    A.dt = (double)c->ens.dt_seconds;
    A.h = A.dt / 2.0;
    A.approaches = c->ens.closest_approaches;
    A.approach_to = c->ens.approach_to;
    int per_cta = n <= kEnsThreadsMax ? kEnsThreadsMax / n : 1;
    if (per_cta > c->ens_copies)
        per_cta = c->ens_copies;
    A.per_cta = per_cta;
    int threads = (per_cta * n + 31) / 32 * 32;
    size_t smem = (size_t)per_cta * 4 * npad * sizeof(double);
    int blocks = (c->ens_copies + per_cta - 1) / per_cta;
    cudaError_t e;
    if (c->ens.exact_order) {
        if (threads > kEnsThreadsMax) // a template wider than one ensemble CTA: one copy per CTA at full width
            return cudaErrorInvalidConfiguration;
        k_ensemble<true><<<blocks, threads, smem, st>>>(A);
    } else {
        if (threads > kEnsThreadsMax)
            return cudaErrorInvalidConfiguration;
        k_ensemble<false><<<blocks, threads, smem, st>>>(A);
    }
    e = cudaGetLastError();
    c->launches++;
    int P = A.steps + 1;
    c->passes += P;
    c->interactions += (double)P * (double)c->ens_copies * (double)n * (double)(n - 1);
    c->last_force_passes = P;
    return e;
}

} // namespace essa
