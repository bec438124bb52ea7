// K3xw -- K3x (resident_cluster.cu) for worlds of 129 .. 1024 bodies: the same 16-SM cluster, the same
// st.async position exchange through distributed shared memory and the same consumer warps, but several
// bodies per warp:
//
//   bodies      lanes per body (LB)   bodies per CTA   physics warps per CTA   ring slots
//   129 .. 256        16                 <= 16               <= 8                  8
//   257 .. 512         8                 <= 32               <= 8                  8
//   513 .. 1024        4                 <= 64               <= 8                  4
//
// (eight physics warps at most, so that the kernel may use 128 registers per thread: with 24 warps the
// compiler serialised the partner chains to fit 80)
//
// Lane q of a body's group takes the partners p = q, q + LB, ... in a four-way unrolled loop over 32-byte
// records {x, y, z, gf} (the ring is padded with zero-weight records to a multiple of 4 LB), a log2(LB)-level
// shuffle butterfly leaves the same total in every lane of the group, and every lane kicks and drifts the
// replicated state.  The most-attracting candidate rides along at one multiply and one compare per interaction
// and is reduced over the group with two integer warp reductions (redux.sync with the group's member mask) over
// sortable keys: metric bits with the 10 lowest mantissa bits replaced by 1023 - p.
//
// Why: a single SM (K3) needs n (n - 1) x 17 / 64 cycles per step -- 176 us at 1024 bodies -- and the tiled path
// pays ~24 us of launches per step; sixteen SMs that never leave the kernel need 1 .. 10 us (measured numbers in
// profiles/README.md).  Worlds whose active mask changes during the call, two_pass, exact_order and
// closest-approach tracking stay on resident.cu.
#include <cstdlib>

#include "common.cuh"
#include "ctx.hpp"
#include "record.cuh"

namespace essa {

constexpr int kWMaxBodies = 1024;
constexpr int kWLocalMax = 64; // bodies per CTA
constexpr int kWRoles = 4;     // consumer roles, each on ceil(local bodies / 32) warps
constexpr int kWCtas = 16;
constexpr int kWMaxWarps = 8 + 2 * kWRoles; // 8 physics warps + consumers: 512 threads, 128 registers each
constexpr int kWDone = kWRoles * 2 * kWCtas;

struct WArgs {
    WorldPtrs w;
    RecordPtrs r;
    int n, npad, steps, nbl, ctas, ring, phys_warps, chunks;
    double hm, dtm, eps2;
    int record, argmax, offset_trails, forward_sim, acc_valid;
};

// dynamic shared memory: pos[ring][npad] | vel[ring][3][kWLocalMax] | most[ring][kWLocalMax] | full[ring] | vready[ring] | done[kWDone]
struct WView {
    double4* pos;
    double* vel;
    int* most;
    unsigned long long* full;
    unsigned long long* vready;
    unsigned* done;
};

__device__ __forceinline__ WView w_view(unsigned char* base, int ring, int npad) {
    WView v;
    v.pos = reinterpret_cast<double4*>(base);
    v.vel = reinterpret_cast<double*>(v.pos + (size_t)ring * npad);
    v.most = reinterpret_cast<int*>(v.vel + (size_t)ring * 3 * kWLocalMax);
    v.full = reinterpret_cast<unsigned long long*>(v.most + (size_t)ring * kWLocalMax);
    v.vready = v.full + ring;
    v.done = reinterpret_cast<unsigned*>(v.vready + ring);
    return v;
}

static size_t w_smem_bytes(int ring, int npad) {
    return (size_t)ring * npad * sizeof(double4) + (size_t)ring * 3 * kWLocalMax * sizeof(double) + (size_t)ring * kWLocalMax * sizeof(int)
        + 2 * (size_t)ring * sizeof(unsigned long long) + kWDone * sizeof(unsigned);
}

__device__ __forceinline__ void w_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// version v lives in slot v % ring; slot 0's barrier is first used by version `ring`
__device__ __forceinline__ unsigned w_parity(int v, int ring) { return (unsigned)((v / ring) - ((v % ring) == 0 ? 1 : 0)) & 1u; }

__device__ __forceinline__ double w_metric(double4 const& pk, double4 const& pp, double gp) {
    double dx = pp.x - pk.x, dy = pp.y - pk.y, dz = pp.z - pk.z;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double y0 = rsqrt_seed(r2);
    double t = r2 * y0;
    double e = fma(-t, y0, 1.0);
    double y = fma(y0 * e, fma(0.375, e, 0.5), y0); // 1 / r
    double y2 = y * y;
    return gp * y2 * y2;
}

template<int LB, bool ARGMAX>
__global__ void __launch_bounds__(32 * kWMaxWarps, 1) k_resident_cluster_wide(WArgs const A) {
    extern __shared__ __align__(16) unsigned char w_smem[];
    constexpr int BW = 32 / LB; // bodies per warp
    int const n = A.n, npad = A.npad, nbl = A.nbl, ring = A.ring;
    WView const S = w_view(w_smem, ring, npad);
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const cta = (int)cluster_ctarank();
    int const first = cta * nbl;
    int const nl = max(0, min(nbl, n - first)); // bodies this CTA owns
    bool const rec_on = A.record != 0, snap_on = rec_on || ARGMAX;
    unsigned const pos_bytes = (unsigned)n * 32u; // what arrives per step

    if (threadIdx.x == 0) {
        for (int i = 0; i < ring; i++) {
            mbar_init(&S.full[i], 1);
            mbar_init(&S.vready[i], nl > 0 ? nl : 1);
        }
        mbar_fence_init();
        for (int i = 0; i < ring; i++)
            mbar_expect_tx(&S.full[i], pos_bytes);
    }
    for (int i = threadIdx.x; i < kWDone; i += blockDim.x)
        S.done[i] = 0;
    for (int i = threadIdx.x; i < ring * npad; i += blockDim.x) {
        // slot 0 = the state at launch; everywhere the padding behind body n - 1 is read with weight 0
        int const b = i % npad;
        S.pos[i] = (i < npad && b < n) ? A.w.pos_cur[b] : make_double4(0, 0, 0, 0);
    }
    __syncthreads();
    cluster_sync_all();

    if (warp < A.phys_warps) {
      if (warp * BW < nl) { // warps without a body of their own stay out (and a CTA without bodies receives nothing)
        // ---------------- physics: BW bodies per warp, LB lanes each ----------------
        int const grp = lane / LB, q = lane - grp * LB;
        int const kl = warp * BW + grp, k = first + kl;
        bool const live = kl < nl;
        int const kc = live ? k : (n - 1);
        unsigned const gmask = LB == 32 ? 0xffffffffu : (((1u << LB) - 1u) << (grp * LB));
        double4 const p0 = S.pos[kc];
        double x = p0.x, y = p0.y, z = p0.z;
        double vx = A.w.v[0][kc], vy = A.w.v[1][kc], vz = A.w.v[2][kc];
        double ax = A.w.a[0][kc], ay = A.w.a[1][kc], az = A.w.a[2][kc];
        double const gfk = A.w.gf[kc];
        bool const active = live && A.w.active[kc] != 0;
        double const gfe = active ? gfk : 0.0;
        double const hm = A.hm, dtm = A.dtm;
        int const per = npad / LB; // a multiple of 4
        int winner = -1;

        auto forces = [&](int slot) {
            double4 const* pp = S.pos + (size_t)slot * npad + q;
            double fx[4] = { 0, 0, 0, 0 }, fy[4] = { 0, 0, 0, 0 }, fz[4] = { 0, 0, 0, 0 }, bm[4] = { 0, 0, 0, 0 };
            int bj[4] = { -1, -1, -1, -1 };
            for (int j = 0; j < per; j += 4) {
                // written stage by stage over the four partners so that the four dependent chains interleave
                // (as one block per partner the compiler ran them back to back to save registers: 47 % of the
                // FP64 pipe at 1024 bodies, measured)
                double4 qv[4];
                double dx[4], dy[4], dz[4], r2[4], y0[4], e[4], c0[4], sc[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    qv[u] = pp[(j + u) * LB];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    dx[u] = qv[u].x - x;
                    dy[u] = qv[u].y - y;
                    dz[u] = qv[u].z - z;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    r2[u] = fma(dz[u], dz[u], fma(dy[u], dy[u], fma(dx[u], dx[u], A.eps2)));
#pragma unroll
                for (int u = 0; u < 4; u++)
                    y0[u] = rsqrt_seed(r2[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    double const y2 = y0[u] * y0[u]; // exact: the seed has 21 significant bits
                    e[u] = fma(-r2[u], y2, 1.0);
                    c0[u] = (qv[u].w * y0[u]) * y2;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    sc[u] = fma(c0[u], fma(1.875, e[u], 1.5) * e[u], c0[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    fx[u] = fma(sc[u], dx[u], fx[u]);
                    fy[u] = fma(sc[u], dy[u], fy[u]);
                    fz[u] = fma(sc[u], dz[u], fz[u]);
                }
                if (ARGMAX) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        double m = c0[u] * y0[u]; // gf_p y0^4 ~ gf_p / r^4 (ranking only); 0 for the self term
                        if (qv[u].w > gfk && m > bm[u]) { // ascending p within a chain: the lowest index wins ties
                            bm[u] = m;
                            bj[u] = q + (j + u) * LB;
                        }
                    }
                }
            }
            double sx = (fx[0] + fx[1]) + (fx[2] + fx[3]);
            double sy = (fy[0] + fy[1]) + (fy[2] + fy[3]);
            double sz = (fz[0] + fz[1]) + (fz[2] + fz[3]);
            unsigned long long key = 0;
            if (ARGMAX) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    unsigned long long ku = ((unsigned long long)__double_as_longlong(bm[u]) & ~1023ull) | (unsigned)(1023 - (bj[u] & 1023));
                    ku = bj[u] >= 0 ? ku : 0ull;
                    key = ku > key ? ku : key;
                }
            }
#pragma unroll
            for (int off = LB / 2; off > 0; off >>= 1) { // a + b == b + a: every lane of the group ends with the same bits
                sx += __shfl_xor_sync(0xffffffffu, sx, off);
                sy += __shfl_xor_sync(0xffffffffu, sy, off);
                sz += __shfl_xor_sync(0xffffffffu, sz, off);
            }
            int bp = -1;
            if (ARGMAX) {
                unsigned const hi = (unsigned)(key >> 32), lo = (unsigned)key;
                unsigned const mh = __reduce_max_sync(gmask, hi);
                unsigned const ml = __reduce_max_sync(gmask, hi == mh ? lo : 0u);
                bp = mh != 0u ? 1023 - (int)(ml & 1023u) : -1;
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

        // lanes q, q + LB, ... of a body's group send its record to those CTAs
        constexpr int kDst = kWCtas / LB > 0 ? kWCtas / LB : 1; // destinations per lane
        uint32_t rec_remote[kDst], full_remote[kDst];
#pragma unroll
        for (int t = 0; t < kDst; t++) {
            int const dst = q + t * LB;
            rec_remote[t] = mapa_u32(&S.pos[kc], dst < A.ctas ? dst : 0);
            full_remote[t] = mapa_u32(&S.full[0], dst < A.ctas ? dst : 0);
        }
        int consumed = 0; // cached lower bound of the versions every consumer of the cluster has finished with

        for (int s = 0; s < A.steps; s++) {
            int const v = s + 1, slot = v % ring;
            if (active) { // first half kick + drift (src/World.cpp:112,115)
                vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                vz = __dadd_rn(vz, __dmul_rn(az, hm));
                x = __dadd_rn(x, __dmul_rn(vx, dtm));
                y = __dadd_rn(y, __dmul_rn(vy, dtm));
                z = __dadd_rn(z, __dmul_rn(vz, dtm));
            }
            if (snap_on && v - consumed > ring) { // the slot holds version v - ring: is everybody done with it?
                do {
                    unsigned m = 0xffffffffu;
                    for (int i = lane; i < kWDone; i += 32) {
                        int const role = i / (2 * kWCtas), chunk = (i / kWCtas) & 1, c = i % kWCtas;
                        int const c_nl = max(0, min(nbl, n - c * nbl));
                        if (c < A.ctas && chunk * 32 < c_nl && (rec_on || role == 1))
                            m = min(m, ld_acquire_cluster_u32(&S.done[i]));
                    }
                    consumed = (int)__reduce_min_sync(0xffffffffu, m);
                } while (v - consumed > ring);
            }
            if (live) {
#pragma unroll
                for (int t = 0; t < kDst; t++) {
                    if (q + t * LB < A.ctas) {
                        uint32_t const dst = rec_remote[t] + (uint32_t)slot * (uint32_t)npad * (uint32_t)sizeof(double4);
                        uint32_t const bar = full_remote[t] + 8u * (uint32_t)slot;
                        st_async_v2_f64(dst, x, y, bar);
                        st_async_v2_f64(dst + 16u, z, gfe, bar);
                    }
                }
            }
            mbar_wait(&S.full[slot], w_parity(v, ring));
            if (warp == 0 && lane == 0)
                mbar_expect_tx(&S.full[slot], pos_bytes); // armed for the slot's next version
            forces(slot);
            if (active) { // second half kick (src/World.cpp:127)
                vx = __dadd_rn(vx, __dmul_rn(ax, hm));
                vy = __dadd_rn(vy, __dmul_rn(ay, hm));
                vz = __dadd_rn(vz, __dmul_rn(az, hm));
            }
            if (snap_on && live && q == 0) {
                if (rec_on) {
                    double* vs = S.vel + (size_t)slot * 3 * kWLocalMax;
                    vs[kl] = vx;
                    vs[kWLocalMax + kl] = vy;
                    vs[2 * kWLocalMax + kl] = vz;
                }
                S.most[slot * kWLocalMax + kl] = active ? winner : -1;
                w_arrive(&S.vready[slot]);
            }
        }
        if (live && q == 0) {
            A.w.pos_cur[k] = make_double4(x, y, z, gfe);
            A.w.v[0][k] = vx;
            A.w.v[1][k] = vy;
            A.w.v[2][k] = vz;
            A.w.a[0][k] = ax;
            A.w.a[1][k] = ay;
            A.w.a[2][k] = az;
        }
      }
    } else if (warp < A.phys_warps + kWRoles * A.chunks) {
        // ---------------- consumers: role 0, 2 = Trail (even / odd local bodies), 1 = links + History, 3 = ap / pe ----------------
        int const cw = warp - A.phys_warps;
        int const role = cw % kWRoles, chunk = cw / kWRoles;
        int const kl = chunk * 32 + lane;
        if (snap_on && chunk * 32 < nl && (role == 1 || rec_on)) {
            bool const is_trail = role == 0 || role == 2;
            bool const body = kl < nl;
            int const k = first + (body ? kl : 0);
            bool const mine = rec_on && body && k < A.r.tracked && (!is_trail || (lane & 1) == (role == 2 ? 1 : 0));
            bool const act = body && A.w.active[k] != 0; // the mask does not change during the call
            bool const offset_trails = A.offset_trails != 0, forward_sim = A.forward_sim != 0;
            uint32_t const done_remote = mapa_u32(&S.done[(role * 2 + chunk) * kWCtas + cta], lane < A.ctas ? lane : 0);
            BodyRec b;
            if (mine)
                rec_load(A.r, k, b);
            int most = body ? A.w.most[k] : -1, old_most = body ? A.w.old_most[k] : -1;
            for (int v = 1; v <= A.steps; v++) {
                int const slot = v % ring;
                unsigned const par = w_parity(v, ring);
                mbar_wait(&S.full[slot], par);
                mbar_wait(&S.vready[slot], par);
                double4 const* pp = S.pos + (size_t)slot * npad;
                int const winner = body ? S.most[slot * kWLocalMax + kl] : -1;
                old_most = most; // Object::before_update (every object)
                if (act && winner >= 0)
                    most = winner;
                if (mine && act) {
                    double4 const pm = pp[most >= 0 ? most : 0], pk = pp[k];
                    double const* vs = S.vel + (size_t)slot * 3 * kWLocalMax;
                    double const vx = vs[kl], vy = vs[kWLocalMax + kl], vz = vs[2 * kWLocalMax + kl];
                    if (is_trail)
                        record_trail(A.r, k, b, pk.x, pk.y, pk.z, vx, vy, vz, pm.x, pm.y, pm.z, most, old_most, offset_trails, forward_sim,
                            false);
                    else if (role == 1) {
                        if (!forward_sim)
                            history_push(A.r, k, b, pk.x, pk.y, pk.z, vx, vy, vz);
                    } else
                        record_appe(b, pk.x, pk.y, pk.z, vx, vy, vz, pm.x, pm.y, pm.z, most);
                }
                if (role == 1 && v == A.steps && act) {
                    // max_attraction as the last force pass left it: the winner's gf / r^4 to full precision, or 0
                    A.w.maxattr[k] = winner >= 0 ? w_metric(pp[k], pp[winner], A.w.gf[winner]) : 0.0;
                }
                __syncwarp();
                if (lane < A.ctas)
                    st_relaxed_cluster_u32(done_remote, (unsigned)v);
            }
            if (role == 1 && body) {
                A.w.most[k] = most;
                A.w.old_most[k] = old_most;
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
    // nobody leaves while a peer may still store into (or signal) its shared memory
    cluster_sync_all();
}

template<int LB>
static cudaError_t launch_w(essa_ctx* c, cudaLaunchConfig_t const& cfg, WArgs const& A) {
    int const idx = LB == 16 ? 0 : (LB == 8 ? 1 : 2);
    if (!c->attr_cluster_wide[idx]) {
        cudaError_t e = cudaSuccess;
        void const* fns[2] = { (void const*)k_resident_cluster_wide<LB, true>, (void const*)k_resident_cluster_wide<LB, false> };
        for (int f = 0; f < 2 && e == cudaSuccess; f++) {
            e = cudaFuncSetAttribute(fns[f], cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(fns[f], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_smem_bytes(8, 512));
        }
        if (e != cudaSuccess)
            return e;
        c->attr_cluster_wide[idx] = true;
    }
    return A.argmax ? cudaLaunchKernelEx(&cfg, k_resident_cluster_wide<LB, true>, A)
                    : cudaLaunchKernelEx(&cfg, k_resident_cluster_wide<LB, false>, A);
}

bool resident_cluster_wide_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only) {
    static int const enabled = [] {
        char const* e = getenv("ESSA_RESIDENT_CLUSTER"); // diagnostics: 0 keeps such worlds on the single-SM kernel
        return e ? atoi(e) : 1;
    }();
    return enabled && !forces_only && !o.exact_order && !o.two_pass && c->n > 128 && c->n <= kWMaxBodies
        && !(o.forward_simulated && c->closest && c->closest_n == c->n);
}

cudaError_t launch_resident_cluster_wide(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st) {
    WArgs A;
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
    int const lb = c->n <= 256 ? 16 : (c->n <= 512 ? 8 : 4);
    A.ring = c->n <= 512 ? 8 : 4;
    A.npad = (c->n + 4 * lb - 1) / (4 * lb) * (4 * lb);
    A.nbl = (c->n + kWCtas - 1) / kWCtas;
    A.ctas = (c->n + A.nbl - 1) / A.nbl;
    A.phys_warps = (A.nbl + 32 / lb - 1) / (32 / lb);
    A.chunks = (A.nbl + 31) / 32;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kWCtas, 1, 1);
    cfg.blockDim = dim3(32 * (A.phys_warps + kWRoles * A.chunks), 1, 1);
    cfg.dynamicSmemBytes = w_smem_bytes(A.ring, A.npad);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kWCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = lb == 16 ? launch_w<16>(c, cfg, A) : (lb == 8 ? launch_w<8>(c, cfg, A) : launch_w<4>(c, cfg, A));
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
