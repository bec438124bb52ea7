// Device-side building blocks shared by every kernel of the World::update path.
//
// Two arithmetic flavours of the pair law of Object::update_forces_against
// (reference src/Object.cpp:64-95):
//   * exact_*  -- the reference's operation order with IEEE-754 round-to-nearest add/mul/div/sqrt
//                 and no FMA contraction: bit-identical to an x86-64 SSE2 build of the reference.
//   * fast_*   -- 16 FP64-pipe instructions per directed interaction: MUFU.RSQ64H seed, one
//                 third-order correction fused with the cube, FMA accumulation.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/essa_b200.h"

// Bounds / index assertions of our own (compute-sanitizer is not available on the GPU pool): compiled in with -DESSA_BOUNDS
// (tools/bounds_check.sh builds that variant and runs the kernels' GPU tests under it), absent from the product build.
#ifdef ESSA_BOUNDS
#include <cstdio>
#define ESSA_CHECK(cond) \
    do { \
        if (!(cond)) { \
            printf("ESSA_BOUNDS: %s:%d: %s (block %d thread %d)\n", __FILE__, __LINE__, #cond, (int)blockIdx.x, (int)threadIdx.x); \
            __trap(); \
        } \
    } while (0)
#else
#define ESSA_CHECK(cond) ((void)0)
#endif

namespace essa {

// ------------------------------------------------------------------------------------------
// fast flavour
// ------------------------------------------------------------------------------------------

// Coordinate of the padding records behind the last body of a world (2^400 m: its square is representable and
// the resulting |d|^-3 underflows to 0).
constexpr double kPadCoord = 2.5822498780869086e120;

// y0 ~ r2^-1/2 from the special-function unit (SASS: MUFU.RSQ64H; only the high word is
// produced, the low word is zero).  r2 == 0 (self term or coincident bodies) yields 0 instead of
// +inf so that the term vanishes; the test is on the high word, in the integer pipe, and also
// swallows denormal r2 (< 2^-1022 m^2).
__device__ __forceinline__ double rsqrt_seed(double r2) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
    int hi = __double2hiint(r2);
    int yh = __double2hiint(y0);
    yh = ((unsigned)hi < 0x00100000u) ? 0 : yh;
    return __hiloint2double(yh, 0);
}

// s = gf / r^3 from r2 = r^2 and the seed:  e = 1 - r2*y0^2,  r2^-3/2 = y0^3 (1-e)^-3/2
//   = y0^3 (1 + e (3/2 + 15/8 e) + O(e^3)).   7 FP64 instructions.
// The seed carries 21 significant bits (high word only), so y0*y0 is EXACT in binary64 and
// e = fma(-r2, y0^2, 1) is the correctly rounded residual: one instruction fewer and no rounding of
// r2*y0 in front of the cancellation.
__device__ __forceinline__ double inv_r3_times(double r2, double gf) {
    double y0 = rsqrt_seed(r2);
    double y2 = y0 * y0;
    double e = fma(-r2, y2, 1.0);
    double gy = gf * y0;
    double c0 = gy * y2;
    double w = fma(1.875, e, 1.5);
    double we = w * e;
    return fma(c0, we, c0);
}

// a_i += gf_j (x_j - x_i) / |x_j - x_i|^3      (3 DADD + 3 DFMA + 7 + 3 DFMA = 16)
__device__ __forceinline__ void fast_interact(double xi, double yi, double zi, double4 const& q, double eps2,
    double& ax, double& ay, double& az) {
    double dx = q.x - xi, dy = q.y - yi, dz = q.z - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double s = inv_r3_times(r2, q.w);
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
}

// Same plus the most-attracting metric gf_j / r^4 (Object.cpp:84-94), restricted to strictly
// heavier partners, strict '>' so that the lowest index wins ties.
__device__ __forceinline__ void fast_interact_argmax(double xi, double yi, double zi, double gfi, double4 const& q, int j,
    double eps2, double& ax, double& ay, double& az, double& best, int& best_j) {
    double dx = q.x - xi, dy = q.y - yi, dz = q.z - zi;
    double r2 = fma(dx, dx, eps2);
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
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
    // |ta|^2 / gf_j = gf_j / r^4 = s * (1/r);  1/r = y0 (1 + e/2 + 3/8 e^2 + O(e^3)), relative error ~1e-17
    double yr = fma(y0 * e, fma(0.375, e, 0.5), y0);
    double metric = s * yr;
    if (metric > best && q.w > gfi) {
        best = metric;
        best_j = j;
    }
}

// ------------------------------------------------------------------------------------------
// exact flavour (reference order).  Row k of World::set_forces's pair-shared loop receives its
// partners in ascending index order (pairs (0,k) .. (k-1,k) from earlier outer iterations, then
// (k,k+1) .. in its own), so a per-row ascending loop reproduces it exactly.
// ------------------------------------------------------------------------------------------

// Partner p acting on body k, p != k:   a_k += ((x_p - x_k) / (r2 * sqrt(r2))) * gf_p.
// For p > k the reference forms x_k - x_p and subtracts; negation is exact in IEEE arithmetic so
// both give the same bits.  Also evaluates Object.cpp:84-94 for row k.
__device__ __forceinline__ void exact_interact(double xk, double yk, double zk, double gfk, double xp, double yp, double zp,
    double gfp, int p, double& ax, double& ay, double& az, double& best, int& best_p) {
    double dx = __dsub_rn(xp, xk), dy = __dsub_rn(yp, yk), dz = __dsub_rn(zp, zk);
    double den = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    den = __dmul_rn(den, __dsqrt_rn(den));
    double tx = __dmul_rn(__ddiv_rn(dx, den), gfp);
    double ty = __dmul_rn(__ddiv_rn(dy, den), gfp);
    double tz = __dmul_rn(__ddiv_rn(dz, den), gfp);
    ax = __dadd_rn(ax, tx);
    ay = __dadd_rn(ay, ty);
    az = __dadd_rn(az, tz);
    double mag = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz)), gfp);
    if (mag > best && gfp > gfk) {
        best = mag;
        best_p = p;
    }
}

// v = v + (a * h) * mul ;  x = x + (v * dt) * mul     (src/World.cpp:112,115,127: separate
// multiply and add, left to right, no contraction)
__device__ __forceinline__ double kick1(double v, double a, double h, double mul) {
    return __dadd_rn(v, __dmul_rn(__dmul_rn(a, h), mul));
}

// Object::deleted() (src/Object.cpp:60-62)
__device__ __forceinline__ bool body_active(bool del_flag, long long ddate, long long cdate, long long date) {
    return !((del_flag && ddate <= date) || cdate > date);
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA unit; SASS UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(void const* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifndef ESSA_MBAR_HINT_NS
#define ESSA_MBAR_HINT_NS 0x989680
#endif
// (the suspend-time hint lets the hardware park the warp until the phase completes instead of returning after a short
// default interval: a warp that spins on try_wait takes issue slots from the warps of its SM sub-partition)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(ESSA_MBAR_HINT_NS)
        : "memory");
}
// ---- thread-block cluster / distributed shared memory ----
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(void const* local, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8-byte store into a peer CTA's shared memory that completes `bar` (a peer mbarrier) by 8 transaction bytes
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_v2_f64(uint32_t remote_addr, double a, double b, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "d"(a),
                 "d"(b), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, unsigned long long v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v),
                 "r"(remote_bar)
                 : "memory");
}
// progress counter in a peer CTA.  Relaxed on purpose: a release here would wait for the caller's
// outstanding GLOBAL stores (ring writes nobody in the cluster reads); what the peer must not overtake are
// the caller's shared-memory LOADS, which have returned their values before this store can issue.
__device__ __forceinline__ void st_relaxed_cluster_u32(uint32_t remote_addr, unsigned v) {
    asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
// the same with release semantics (everything the caller and, through __syncwarp, its warp wrote before is visible to
// a peer that acquires the counter): used where a peer reads DATA the caller stored remotely, once per tick
__device__ __forceinline__ void st_release_cluster_u32(uint32_t remote_addr, unsigned v) {
    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
// plain stores into a peer CTA's shared memory (ordered by a later st_release_cluster_u32)
__device__ __forceinline__ void st_cluster_f64(uint32_t remote_addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote_addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t remote_addr, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
// ---- host <-> device mailbox in mapped pinned memory (persistent resident kernels) ----
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(unsigned long long const* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_cluster_u32(void const* local) {
    unsigned v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(local)) : "memory");
    return v;
}

__device__ __forceinline__ void bulk_g2s(void* dst_smem, void const* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG; host+device so that a checker can rebuild any ensemble copy)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// 1 + rel_sigma * n with n = (sum of 12 uniform 32-bit draws) * 2^-32 - 6  (Irwin-Hall, variance 1).
// Integer sum, one exact conversion: identical bits on host and device.
__host__ __device__ inline double ensemble_factor(uint64_t seed, double rel_sigma, long long copy, int body, int component) {
    if (copy == 0)
        return 1.0;
    uint64_t sum = 0;
    for (uint32_t blk = 0; blk < 3; blk++) {
        uint32_t c[4] = { (uint32_t)copy, (uint32_t)((uint64_t)copy >> 32), (uint32_t)body, (uint32_t)component * 4u + blk };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        sum += (uint64_t)c[0] + c[1] + c[2] + c[3];
    }
    double n = (double)sum * (1.0 / 4294967296.0) - 6.0;
#ifdef __CUDA_ARCH__
    return __dadd_rn(1.0, __dmul_rn(rel_sigma, n));
#else
    return 1.0 + rel_sigma * n;
#endif
}

} // namespace essa
