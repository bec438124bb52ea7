// K1 -- j-tiled FP64 force pass with fused kick / drift.
//
// Replaces World::set_forces + Object::update_forces_against + the kick/drift loops of
// World::update (reference src/World.cpp:42-57,106-116,121-127; src/Object.cpp:64-95) for worlds
// that do not fit one CTA.  Per-body view of the reference's pair-shared loop (SURVEY.md 8a'):
//     a_k = sum_{j != k, active} gf_j (x_j - x_k) / |x_j - x_k|^3
//
// Layout: the j-array is one 32-byte record {x, y, z, gf_eff} per body (gf_eff = 0 for bodies
// that Object::deleted() masks out), double-buffered so that the drift fused into this kernel's
// epilogue can write step k+1's records while other CTAs still stream step k's.
//
// Grid: (i-blocks of 256*R targets) x (S j-chunks).  Each CTA keeps its targets in registers,
// streams its j-chunk through a 4-stage shared-memory ring filled by the TMA unit
// (cp.async.bulk + mbarrier), and writes one partial acceleration per target.  The CTA that
// completes an i-block last (atomic ticket) sums the S partials in slot order -- deterministic,
// no floating-point atomics -- and runs the kick / drift epilogue.
//
// Roofline: FP64 pipe.  16 FP64 instructions per directed interaction (common.cuh), i.e.
// 10/16 = 62.5 % of the 20-flop-per-interaction convention is the ceiling of this formulation.
#include "common.cuh"
#include "ctx.hpp"

namespace essa {

template<int R, bool ARGMAX>
__device__ __forceinline__ void finish_body(ForceArgs const& A, int g, double ax, double ay, double az, double bm, int bj) {
    WorldPtrs const& W = A.w;
    KickArgs const& K = A.k;
    bool act = W.active[g] != 0;
    double4 p = W.pos_cur[g];
    if (ARGMAX && A.save_old)
        W.old_most[g] = W.most[g];
    double vx = 0, vy = 0, vz = 0;
    if (act) {
        W.a[0][g] = ax;
        W.a[1][g] = ay;
        W.a[2][g] = az;
        if (ARGMAX) {
            W.maxattr[g] = bm;
            if (bj >= 0)
                W.most[g] = bj;
        }
        if (K.second_kick) {
            vx = kick1(W.vh[0][g], ax, K.h, K.mul);
            vy = kick1(W.vh[1][g], ay, K.h, K.mul);
            vz = kick1(W.vh[2][g], az, K.h, K.mul);
            W.v[0][g] = vx;
            W.v[1][g] = vy;
            W.v[2][g] = vz;
        } else if (K.advance) {
            vx = W.v[0][g];
            vy = W.v[1][g];
            vz = W.v[2][g];
        }
    }
    if (K.advance) {
        if (act) {
            vx = kick1(vx, ax, K.h, K.mul);
            vy = kick1(vy, ay, K.h, K.mul);
            vz = kick1(vz, az, K.h, K.mul);
            W.vh[0][g] = vx;
            W.vh[1][g] = vy;
            W.vh[2][g] = vz;
            p.x = __dadd_rn(p.x, __dmul_rn(__dmul_rn(vx, K.dt), K.mul));
            p.y = __dadd_rn(p.y, __dmul_rn(__dmul_rn(vy, K.dt), K.mul));
            p.z = __dadd_rn(p.z, __dmul_rn(__dmul_rn(vz, K.dt), K.mul));
        }
        W.pos_next[g] = p;
    }
}

template<int R, bool ARGMAX>
__global__ void __launch_bounds__(kForceThreads, 2) k_force_tiled(ForceArgs const A) {
    __shared__ __align__(128) double4 tile[kStages][kTileJ];
    __shared__ __align__(8) unsigned long long full[kStages];
    __shared__ int s_last;

    int const tid = threadIdx.x;
    int const ib = blockIdx.x / A.S;
    int const sj = blockIdx.x - ib * A.S;
    // this CTA's share of the sources, cut at multiples of kChunkQ records (a whole 256-record tile per CTA would leave a
    // 2,048-body world with 64 CTAs of 18 us each), then streamed in tiles of up to kTileJ
    int const nq = (A.j1 - A.j0 + kChunkQ - 1) / kChunkQ;
    int const jlo = A.j0 + (int)((long long)nq * sj / A.S) * kChunkQ;
    int const jhi_ = A.j0 + (int)((long long)nq * (sj + 1) / A.S) * kChunkQ;
    int const jhi = jhi_ < A.j1 ? jhi_ : A.j1;
    int const t0 = 0;
    int const t1 = jhi > jlo ? (jhi - jlo + kTileJ - 1) / kTileJ : 0;

    double xi[R], yi[R], zi[R], gi[R], ax[R], ay[R], az[R], bm[R];
    int bj[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        int g = A.i0 + ib * (kForceThreads * R) + r * kForceThreads + tid;
        g = g < A.i1 ? g : A.i1 - 1;
        double4 p = A.w.pos_cur[g];
        xi[r] = p.x;
        yi[r] = p.y;
        zi[r] = p.z;
        gi[r] = ARGMAX ? A.w.gf[g] : 0.0;
        ax[r] = ay[r] = az[r] = 0.0;
        bm[r] = 0.0;
        bj[r] = -1;
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++)
            mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t, int s) {
        int jb = jlo + t * kTileJ;
        int cnt = min(kTileJ, jhi - jb);
        unsigned bytes = (unsigned)cnt * (unsigned)sizeof(double4);
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(&tile[s][0], A.w.pos_cur + jb, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages; s++)
            if (t0 + s < t1)
                issue(t0 + s, s);
    }

    for (int t = t0; t < t1; t++) {
        int const k = t - t0;
        int const s = k % kStages;
        mbar_wait(&full[s], (unsigned)(k / kStages) & 1u);
        int const jb = jlo + t * kTileJ;
        int const cnt = min(kTileJ, jhi - jb);
        double4 const* tp = tile[s];
#pragma unroll 4
        for (int jj = 0; jj < cnt; jj++) {
            double4 q = tp[jj];
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (ARGMAX)
                    fast_interact_argmax(xi[r], yi[r], zi[r], gi[r], q, jb + jj, A.eps2, ax[r], ay[r], az[r], bm[r], bj[r]);
                else
                    fast_interact(xi[r], yi[r], zi[r], q, A.eps2, ax[r], ay[r], az[r]);
            }
        }
        __syncthreads();
        if (tid == 0 && t + kStages < t1)
            issue(t + kStages, s);
    }

    // ---- partial results, slot order = launch order x chunk order ---------------------------
    int const slot = A.slot0 + sj;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int li = ib * (kForceThreads * R) + r * kForceThreads + tid;
        size_t base = (size_t)slot * 3 * A.ipad + li;
        __stcg(A.part + base, ax[r]);
        __stcg(A.part + base + A.ipad, ay[r]);
        __stcg(A.part + base + 2 * (size_t)A.ipad, az[r]);
        if (ARGMAX) {
            __stcg(A.part_m + (size_t)slot * A.ipad + li, bm[r]);
            __stcg(A.part_j + (size_t)slot * A.ipad + li, bj[r]);
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned prev = atomicAdd(A.counters + ib, 1u);
        s_last = (prev == (unsigned)(A.slots_total - 1));
    }
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();

    // ---- last CTA of this i-block: ordered reduction + kick / drift ------------------------------
#pragma unroll
    for (int r = 0; r < R; r++) {
        int li = ib * (kForceThreads * R) + r * kForceThreads + tid;
        int g = A.i0 + li;
        if (g >= A.i1)
            continue;
        double sx = 0, sy = 0, sz = 0, m = 0;
        int mj = -1;
        for (int sl = 0; sl < A.slots_total; sl++) {
            size_t base = (size_t)sl * 3 * A.ipad + li;
            sx = __dadd_rn(sx, __ldcg(A.part + base));
            sy = __dadd_rn(sy, __ldcg(A.part + base + A.ipad));
            sz = __dadd_rn(sz, __ldcg(A.part + base + 2 * (size_t)A.ipad));
            if (ARGMAX) {
                double pm = __ldcg(A.part_m + (size_t)sl * A.ipad + li);
                int pj = __ldcg(A.part_j + (size_t)sl * A.ipad + li);
                if (pm > m) {
                    m = pm;
                    mj = pj;
                }
            }
        }
        finish_body<R, ARGMAX>(A, g, sx, sy, sz, m, mj);
    }
    if (tid == 0)
        A.counters[ib] = 0;
}

// First half kick + drift of the first step of a call (later steps get it from the force
// epilogue).  src/World.cpp:106-116.
__global__ void k_kick_drift(WorldPtrs const W, KickArgs const K, int i0, int i1) {
    int g = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= i1)
        return;
    double4 p = W.pos_cur[g];
    if (W.active[g]) {
        double vx = kick1(W.v[0][g], W.a[0][g], K.h, K.mul);
        double vy = kick1(W.v[1][g], W.a[1][g], K.h, K.mul);
        double vz = kick1(W.v[2][g], W.a[2][g], K.h, K.mul);
        W.vh[0][g] = vx;
        W.vh[1][g] = vy;
        W.vh[2][g] = vz;
        p.x = __dadd_rn(p.x, __dmul_rn(__dmul_rn(vx, K.dt), K.mul));
        p.y = __dadd_rn(p.y, __dmul_rn(__dmul_rn(vy, K.dt), K.mul));
        p.z = __dadd_rn(p.z, __dmul_rn(__dmul_rn(vz, K.dt), K.mul));
    }
    W.pos_next[g] = p;
}

// Object::deleted() for every body at the context's date; keeps gf_eff of the j-records in step.
__global__ void k_refresh_active(double4* pos, double const* gf, uint8_t* active, uint8_t const* delflag, long long const* ddate,
    long long const* cdate, long long date, int n) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    bool act = body_active(delflag[g] != 0, ddate[g], cdate[g], date);
    active[g] = act ? 1 : 0;
    pos[g].w = act ? gf[g] : 0.0;
}

// Builds the j-records of the current buffer and clears the padding [n, cap) of BOTH buffers
// (kernels that stream whole blocks rely on zero-weight records there).
__global__ void k_pack(double4* pos, double4* other, double const* x, double const* y, double const* z, double const* gf,
    uint8_t const* active, int n, int cap) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cap)
        return;
    if (g < n) {
        pos[g] = make_double4(x[g], y[g], z[g], active[g] ? gf[g] : 0.0);
    } else {
        // padding records attract nothing (gf = 0) and sit far from everything, away from the origin a real body
        // may occupy: the pair-shared kernel's off-diagonal tiles run without an r = 0 test
        pos[g] = make_double4(kPadCoord, kPadCoord, kPadCoord, 0);
        other[g] = make_double4(kPadCoord, kPadCoord, kPadCoord, 0);
    }
}

__global__ void k_unpack(double4 const* pos, double* x, double* y, double* z, int n) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    double4 p = pos[g];
    x[g] = p.x;
    y[g] = p.y;
    z[g] = p.z;
}

WorldPtrs world_ptrs(essa_ctx* c) {
    WorldPtrs w;
    w.pos_cur = c->pos[c->cur];
    w.pos_next = c->pos[c->cur ^ 1];
    for (int k = 0; k < 3; k++) {
        w.v[k] = c->vel + (size_t)k * c->cap;
        w.vh[k] = c->velh + (size_t)k * c->cap;
        w.a[k] = c->acc + (size_t)k * c->cap;
    }
    w.gf = c->gf;
    w.active = c->active;
    w.most = c->most;
    w.old_most = c->old_most;
    w.maxattr = c->maxattr;
    return w;
}

// Targets per thread: more amortises the shared-memory reads, fewer gives small worlds enough
// CTAs to fill 148 SMs.
int force_iblocks(essa_ctx const* c, int* R_out) {
    int ni = c->n;
    if (c->world > 1) {
        int i0 = (int)((long long)c->n * c->rank / c->world), i1 = (int)((long long)c->n * (c->rank + 1) / c->world);
        ni = i1 - i0;
    }
    int R = ni >= 4 * kForceThreads * 2 * c->sm_count ? 4 : (ni >= 2 * kForceThreads * c->sm_count ? 2 : 1);
    *R_out = R;
    return (ni + kForceThreads * R - 1) / (kForceThreads * R);
}

// j-chunks per launch: minimise (waves of 2 CTAs/SM) x (quanta of kChunkQ sources in the longest chunk).
void choose_split(essa_ctx const* c, int iblocks, int nquanta, int* S_out) {
    long long slots = 2LL * c->sm_count;
    int smax = nquanta < 64 ? nquanta : 64;
    if (smax < 1)
        smax = 1;
    long long best = -1;
    int bestS = 1;
    for (int S = 1; S <= smax; S++) {
        long long waves = ((long long)iblocks * S + slots - 1) / slots;
        long long cost = waves * ((nquanta + S - 1) / S) + 3 * waves; // + per-wave fixed cost (launch ramp, first tile, ticket, ordered sum)
        if (best < 0 || cost < best) {
            best = cost;
            bestS = S;
        }
    }
    *S_out = bestS;
}

static cudaError_t ensure_partials(essa_ctx* c, int iblocks, int ipad, int slots_total) {
    size_t need = (size_t)slots_total * 3 * ipad;
    if (need > c->part_doubles) {
        cudaFree(c->part);
        cudaFree(c->part_m);
        cudaFree(c->part_j);
        c->part = nullptr;
        c->part_m = nullptr;
        c->part_j = nullptr;
        cudaError_t e = cudaMalloc(&c->part, need * sizeof(double));
        if (e != cudaSuccess)
            return e;
        e = cudaMalloc(&c->part_m, (size_t)slots_total * ipad * sizeof(double));
        if (e != cudaSuccess)
            return e;
        e = cudaMalloc(&c->part_j, (size_t)slots_total * ipad * sizeof(int));
        if (e != cudaSuccess)
            return e;
        c->part_doubles = need;
    }
    if (iblocks > c->counters_cap) {
        cudaFree(c->counters);
        cudaError_t e = cudaMalloc(&c->counters, (size_t)iblocks * sizeof(unsigned));
        if (e != cudaSuccess)
            return e;
        e = cudaMemsetAsync(c->counters, 0, (size_t)iblocks * sizeof(unsigned), c->stream);
        if (e != cudaSuccess)
            return e;
        c->counters_cap = iblocks;
    }
    return cudaSuccess;
}

template<int R>
static void launch_R(bool argmax, int grid, ForceArgs const& A, cudaStream_t st) {
    if (argmax)
        k_force_tiled<R, true><<<grid, kForceThreads, 0, st>>>(A);
    else
        k_force_tiled<R, false><<<grid, kForceThreads, 0, st>>>(A);
}

cudaError_t launch_force_pass(essa_ctx* c, int j0, int j1, int S, int slot0, int slots_total, bool argmax, bool save_old,
    KickArgs const& k, double eps2, cudaStream_t st) {
    int R;
    int iblocks = force_iblocks(c, &R);
    int ipad = iblocks * kForceThreads * R;
    cudaError_t e = ensure_partials(c, iblocks, ipad, slots_total);
    if (e != cudaSuccess)
        return e;
    ForceArgs A;
    A.w = world_ptrs(c);
    A.n = c->n;
    A.i0 = (int)((long long)c->n * c->rank / c->world);
    A.i1 = (int)((long long)c->n * (c->rank + 1) / c->world);
    A.j0 = j0;
    A.j1 = j1;
    A.S = S;
    A.slot0 = slot0;
    A.slots_total = slots_total;
    A.ipad = ipad;
    A.part = c->part;
    A.part_m = c->part_m;
    A.part_j = c->part_j;
    A.counters = c->counters;
    A.eps2 = eps2;
    A.save_old = save_old ? 1 : 0;
    A.k = k;
    int grid = iblocks * S;
    if (R == 4)
        launch_R<4>(argmax, grid, A, st);
    else if (R == 2)
        launch_R<2>(argmax, grid, A, st);
    else
        launch_R<1>(argmax, grid, A, st);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_kick_drift(essa_ctx* c, KickArgs const& k, cudaStream_t st) {
    int i0 = (int)((long long)c->n * c->rank / c->world), i1 = (int)((long long)c->n * (c->rank + 1) / c->world);
    int ni = i1 - i0;
    if (ni <= 0)
        return cudaSuccess;
    k_kick_drift<<<(ni + 255) / 256, 256, 0, st>>>(world_ptrs(c), k, i0, i1);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_refresh_active(essa_ctx* c, cudaStream_t st) {
    if (c->n <= 0)
        return cudaSuccess;
    k_refresh_active<<<(c->n + 255) / 256, 256, 0, st>>>(c->pos[c->cur], c->gf, c->active, c->delflag, (long long const*)c->ddate,
        (long long const*)c->cdate, (long long)c->date, c->n);
    c->launches++;
    return cudaGetLastError();
}

// Values the Object constructor gives to members an upload leaves out (ap = 0, pe = DBL_MAX, creation date "always"): written
// here instead of being looped over on the host and copied.
template<class T>
__global__ void k_fill(T* dst, T v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = v;
}
cudaError_t launch_fill_f64(essa_ctx* c, double* dst, double v, size_t n, cudaStream_t st) {
    if (n == 0)
        return cudaSuccess;
    k_fill<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, v, n);
    c->launches++;
    return cudaGetLastError();
}
cudaError_t launch_fill_i64(essa_ctx* c, int64_t* dst, int64_t v, size_t n, cudaStream_t st) {
    if (n == 0)
        return cudaSuccess;
    k_fill<long long><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<long long*>(dst), (long long)v, n);
    c->launches++;
    return cudaGetLastError();
}

// essa_upload_shard: the j-records of bodies [i0, i1) of the current buffer from freshly uploaded coordinates.
__global__ void k_pack_range(double4* pos, double const* x, double const* y, double const* z, double const* gf, uint8_t const* active,
    int i0, int i1) {
    int g = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (g < i1)
        pos[g] = make_double4(x[g], y[g], z[g], active[g] ? gf[g] : 0.0);
}
cudaError_t launch_pack_range(essa_ctx* c, double const* x, double const* y, double const* z, int i0, int i1, cudaStream_t st) {
    if (i1 <= i0)
        return cudaSuccess;
    k_pack_range<<<(i1 - i0 + 255) / 256, 256, 0, st>>>(c->pos[c->cur], x, y, z, c->gf, c->active, i0, i1);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_pack(essa_ctx* c, double const* x, double const* y, double const* z, cudaStream_t st) {
    if (c->n <= 0)
        return cudaSuccess;
    k_pack<<<(c->cap + 255) / 256, 256, 0, st>>>(c->pos[c->cur], c->pos[c->cur ^ 1], x, y, z, c->gf, c->active, c->n, c->cap);
    c->launches++;
    return cudaGetLastError();
}

// Small worlds: every field essa_download can return, written by ONE kernel straight into pinned host
// memory (device-visible under unified addressing) in the staging layout of capi.cu -- instead of ~16
// device-to-host copies of a few hundred bytes each, which cost ~50 us of a 10-step GUI tick.
__global__ void k_download_small(WorldPtrs const W, double const* __restrict__ appe, size_t cap, long long const* __restrict__ cdate,
    long long const* __restrict__ ddate, uint8_t const* __restrict__ delflag, double* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    size_t const N = (size_t)n;
    double4 p = W.pos_cur[i];
    out[i] = p.x;
    out[N + i] = p.y;
    out[2 * N + i] = p.z;
    for (int k = 0; k < 3; k++) {
        out[(3 + k) * N + i] = W.v[k][i];
        out[(6 + k) * N + i] = W.a[k][i];
    }
    out[9 * N + i] = W.gf[i];
    out[10 * N + i] = W.maxattr[i];
    for (int k = 0; k < 4; k++)
        out[(11 + k) * N + i] = appe[(size_t)k * cap + i];
    long long* dates = reinterpret_cast<long long*>(out + 15 * N);
    dates[i] = cdate[i];
    dates[N + i] = ddate[i];
    int* most = reinterpret_cast<int*>(dates + 2 * N);
    most[i] = W.most[i];
    uint8_t* flags = reinterpret_cast<uint8_t*>(most + N);
    flags[i] = delflag[i];
}

cudaError_t launch_download_small(essa_ctx* c, void* pinned_out, cudaStream_t st) {
    if (c->n <= 0)
        return cudaSuccess;
    k_download_small<<<(c->n + 127) / 128, 128, 0, st>>>(world_ptrs(c), c->appe, (size_t)c->cap, (long long const*)c->cdate,
        (long long const*)c->ddate, c->delflag, static_cast<double*>(pinned_out), c->n);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_unpack(essa_ctx* c, double* x, double* y, double* z, cudaStream_t st) {
    if (c->n <= 0)
        return cudaSuccess;
    k_unpack<<<(c->n + 255) / 256, 256, 0, st>>>(c->pos[c->cur], x, y, z, c->n);
    c->launches++;
    return cudaGetLastError();
}

} // namespace essa
