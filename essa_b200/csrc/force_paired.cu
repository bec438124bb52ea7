// K1p -- pair-shared force pass: every unordered pair is evaluated ONCE and applied to both bodies,
// as the reference's own loop does (World::set_forces walks i < j and Object::update_forces_against
// updates both objects, src/World.cpp:48-56, src/Object.cpp:64-79).
//
//   d = x_j - x_i ;  u = |d|^-3 ;  a_i += gf_j u d ;  a_j -= gf_i u d
//
// 20 FP64-pipe instructions per pair = 10 per directed interaction (18 = 9 when every body has the same
// gravity factor; K1 needs 16), so the 20-flop-per-interaction convention of BASELINE.json tops out at
// 100 % (111 %) of the DFMA peak instead of K1's 62.5 %.
//
// Decomposition.  Bodies are cut into blocks of 2048.  Unordered block pairs are enumerated as
// (I, J = I + d mod nb), d = 0 .. nb/2, so that every block does the same amount of work.  One CTA
// (8 warps) owns block I for its whole life: warp w keeps targets [256 w, 256 w + 256) of the block
// in registers (8 per lane).  A J block arrives through the TMA unit (cp.async.bulk + mbarrier) in
// four 512-record parts; each warp walks a part in 32-body subsets, 1 body per lane, and ROTATES the
// subset -- positions, gf and the subset's own accumulators -- around the warp with shuffles, so
// that after 32 rotations all 256 x 32 pairs of (targets, subset) have been visited with nothing
// but registers (14 shuffles per 8 pairs).  Measured on B200 at N = 1,048,576 (then 21 instructions per pair): 8 x 1 blocking 755 ms,
// 4 x 2 blocking 804 ms, 4 x 2 at two CTAs per SM 824 ms per pass.  The subset's accumulated a_j goes to a per-warp shared-memory array; after a
// part, the 8 warps' arrays are summed in warp order and written to the slot (I, d) of the J block.
// No floating-point atomics anywhere: k_pair_finish adds, per body, the I-side partials and the
// slots in a fixed order, then runs the same kick / drift epilogue as K1.  Deterministic.
//
// d = 0 (the block against itself) is evaluated one-sided (a_i only), so ordered pairs are not
// double counted; it is 1 / (nb + 1) of the work.
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "ctx.hpp"

#include <type_traits>

namespace essa {

constexpr int kPThreads = 256;  // 8 warps
constexpr int kPRJ = 1;         // visiting bodies per lane
constexpr int kPHalf = 512;     // records per staged part of a J block
// Targets per lane RI is a template parameter: 8 (blocks of 2048 bodies) for large worlds, 4 (blocks
// of 1024) when 2048-body blocks would leave SMs idle.  Bodies per block = 256 * RI.
constexpr int kPSub = 32 * kPRJ;
constexpr int kPairMaxPeers = 8; // ranks whose slot buffers a kernel can address (one node)
constexpr size_t kPSmem = 2 * kPHalf * sizeof(double4) + 8 * 3 * kPHalf * sizeof(double) + 64;
constexpr size_t kPSmemKeys = 8 * kPHalf * sizeof(unsigned long long) + 2 * kPHalf * sizeof(unsigned) + 8 * kPThreads * sizeof(unsigned long long); // + per-warp J-side candidate keys and the staged prior thresholds of the visitors (ARGMAX)

// Most-attracting candidates (src/Object.cpp:84-94) travel as sortable 64-bit keys: the bit pattern of the ranking
// metric gf_p |d|^-3 y0 ~ gf_p / r^4 (positive doubles order like unsigned integers; y0 = rsqrt seed, relative
// error < 1e-6: enough unless two heavier bodies attract within 1 ppm of each other) with its kKeyIdxBits lowest
// mantissa bits replaced by (mask - index), so that max() picks the largest metric and, among equal ones, the
// lowest index -- integer-pipe work instead of FP64 compare / select chains.  0 = no candidate.
constexpr int kKeyIdxBits = 21; // worlds of up to 2,097,152 bodies
constexpr unsigned long long kKeyIdxMask = (1ull << kKeyIdxBits) - 1ull;
__device__ __forceinline__ unsigned long long pack_key(double m, int idx) {
    return ((unsigned long long)__double_as_longlong(m) & ~kKeyIdxMask) | (kKeyIdxMask - (unsigned long long)idx);
}
__device__ __forceinline__ unsigned long long key_max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ int key_index(unsigned long long k) { return (int)(kKeyIdxMask - (k & kKeyIdxMask)); }

// Two candidates whose keys agree in every metric bit a key keeps (the upper 31 mantissa bits: they attract within
// 5e-10 of each other) but name different bodies: the key cannot rank them, and "lowest index" would be a guess.
// The reference decides such a case in full binary64 (src/Object.cpp:84-94): mag = |ta|^2 / gf_p with
// ta = (d / (r2 sqrt r2)) gf_p, strict '>', partners in ascending index order.  Re-evaluated here with the same IEEE
// operations (no contraction) for exactly these two partners, so the link that comes out is the reference's -- also at an
// exact tie, where the lower index wins.  Cold: a few loads and divisions for the rare body that has such a pair.
__device__ __forceinline__ double reference_metric(double4 const t, double4 const p, double eps2) {
    double dx = __dsub_rn(p.x, t.x), dy = __dsub_rn(p.y, t.y), dz = __dsub_rn(p.z, t.z);
    double den = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    den = __dadd_rn(den, eps2); // eps2 == 0: the reference's expression
    den = __dmul_rn(den, __dsqrt_rn(den));
    double tx = __dmul_rn(__ddiv_rn(dx, den), p.w), ty = __dmul_rn(__ddiv_rn(dy, den), p.w), tz = __dmul_rn(__ddiv_rn(dz, den), p.w);
    return __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz)), p.w);
}
__device__ __noinline__ unsigned long long key_resolve(double4 const* __restrict__ pos, int target, unsigned long long a, unsigned long long b,
    double eps2) {
    int const ia = key_index(a), ib = key_index(b);
    double4 const t = pos[target];
    double const ma = reference_metric(t, pos[ia], eps2), mb = reference_metric(t, pos[ib], eps2);
    bool const take_a = ma > mb || (ma == mb && ia < ib);
    return take_a ? a : b;
}
// max of two candidate keys of body `target`, with the exact re-decision where the keys cannot tell
__device__ __forceinline__ unsigned long long key_merge(double4 const* __restrict__ pos, int target, unsigned long long a, unsigned long long b,
    double eps2) {
    if (a != b && a != 0ull && b != 0ull && ((a ^ b) >> kKeyIdxBits) == 0ull)
        return key_resolve(pos, target, a, b, eps2);
    return a > b ? a : b;
}

// Candidate filter in the integer pipe.  For a positive double v = 2^e (1 + f), its high word is
//     L(v) = 2^20 (e + 1023 + f) = 2^20 (log2 v + 1023 + delta),   delta = f - log2(1 + f)  in  [-0.0861, 0]
// (the mantissa is a piecewise-linear logarithm).  The metric of a pair is gf_p y0^4 up to 3e-6, so
//     L(gf_p) / 4 + L(y0) - L(best) / 4 - 1023 * 2^20  >=  2^20 [(log2 metric - log2 best) / 4 - 0.0861 (1/4 + 1)] - 3
// and a partner that beats the best candidate so far ALWAYS passes  lg(gf_p) + L(y0) >= thr(best)  with the margin
// below (as does one whose truncated gf is not smaller: a superset of "strictly heavier").  The pairs that pass --
// a few dozen per body and pass, the running maxima of a random sequence and their near misses within a factor
// 1.4 -- run the exact test (heavier partner, full key compare) in a cold block.  The hot path pays integer adds
// and compares per side instead of DMUL + DSETP + 64-bit select chains.
// Measured alternatives (N = 262,144, masses spread over a factor 4; untracked 1.56e12 interactions/s): the exact
// test for every pair (DMUL + DSETP + 64-bit selects) 1.01e12; this filter with thresholds that start from scratch
// in every CTA and visit 7.2e11 (the first rotations of each visit all pass); with priors 1.15e12; one test per
// rotation against the lane's weakest threshold 8.0e11 (thresholds differ by orders of magnitude between core and
// halo bodies, so the weakest of eight passes nearly everything).
constexpr unsigned kKeyMargin = 114688u; // > 0.0861 * 1.25 * 2^20 = 112,853, + slack for the truncations
// Thresholds survive from one force pass to the next ("priors"): k_pair_finish leaves, for every body, the threshold
// of its winning key lowered by kPriorSlack (a factor 4 in the metric, 41 % in distance: bodies move little in a step), and the next
// pass starts every filter there instead of at "anything goes" -- otherwise each CTA (I side) and each visit of a
// 32-body subset (J side) would spend its first rotations in the cold block re-discovering the running maximum.
// k_pair_finish accepts a winner only if  key_thr(winner) > prior  (then every better partner was above the prior
// too and has been seen); otherwise that body is re-ranked from scratch by the repair path.
// Measured at 1,048,576 bodies (ms per pass: pair kernel + k_pair_finish with its re-ranks): slack 2^18 (a factor 2) 740.9 + 21.1,
// 2^19 741.5 + 6.2, 2^20 744.2 + 4.8 -- a looser filter costs a few more re-walked rotations, a tighter one re-ranks (0.1 ms per
// body) every body whose partner came 19 % closer or farther within one step.  ESSA_PAIR_PRIOR_SLACK_LOG2 overrides (diagnostics).
constexpr unsigned kPriorSlack = 1u << 19;
constexpr unsigned kPriorNone = 0x3FF00000u - kKeyMargin + 1u; // = key_thr(0): every real candidate passes
constexpr unsigned kPriorNever = 0xFFFFFFFFu;              // padding and masked bodies: nothing passes
__device__ __forceinline__ unsigned key_lg(double gf) { return (unsigned)__double2hiint(gf) >> 2; }
// a partner passes when  lg(gf_p) + L(y0) >= key_thr(best)
__device__ __forceinline__ unsigned key_thr(unsigned long long key) { return ((unsigned)(key >> 32) >> 2) + (0x3FF00000u - kKeyMargin + 1u); }

struct PairArgs {
    double4 const* pos; // padded to a multiple of kPB with {0,0,0,0}
    int nb;             // blocks in the world
    int dmax;           // largest block offset evaluated: nb / 2
    // Offsets of THIS launch: d0 .. d0 + dn - 1.  One launch covers 0 .. dmax unless the J-side slots of all offsets would not
    // fit the slot budget (paired_plan): then the offsets go in batches of `dslots`, each batch folded into a running sum
    // (k_pair_fold) before the next one reuses the slot planes.  Offset d owns plane d - dslot0 of a block's dslots planes.
    int d0, dn, dslots, dslot0;
    int S;              // chunks of the offset range per I block (gridDim.x = owned blocks * S)
    int ib0, ib1;       // I blocks owned by this context
    int npad;           // nb * kPB
    double* part_i;     // [S][3][npad]       I-side partial accelerations
    // J-side partials of tile (I, d), d = 1 .. dmax, live with the OWNER of block J = I + d: slot (J, d) of
    // [blocks of the owner][dmax][3][kPB].  On one GPU that is this context's own buffer; in a sharded world the kernel
    // stores them straight into the owning rank's buffer over NVLink (peer-mapped with cudaIpc, capi.cu) -- the force pass
    // and its "reduce-scatter" are ONE kernel, and k_pair_finish adds a body's slots in the same fixed order (ascending d)
    // on any number of ranks (the I-side chunking S still follows the plan, so rank counts agree to rounding only).  owner_major == 0: the fallback when peer
    // access is not available: writer-major slots [ib1-ib0][dmax][3][kPB] in local memory, folded by k_pair_gather_slots
    // and reduce-scattered with NCCL.
    double* slot_peer[kPairMaxPeers];
    unsigned long long* slotk_peer[kPairMaxPeers]; // the same for the candidate keys (ARGMAX): [blocks of the owner][dmax][kPB]
    int bpr;            // blocks per rank
    int owner_major;
    double* slots;      // writer-major fallback
    unsigned long long* part_k;  // [S][npad]              I-side candidate keys (ARGMAX)
    unsigned long long* slots_k; // writer-major fallback
    unsigned const* prior;       // [npad] filter thresholds left by the previous pass (ARGMAX)
    // SORTED (mass-sorted layout, see k_force_paired): `pos`, `prior`, partials, slots and keys' homes are in the sorted domain;
    // slot s holds body perm[s] (-1: padding) of pos_orig, and candidate keys carry ORIGINAL indices
    int const* perm;
    double4 const* pos_orig;
    // sharded, peer-mapped position buffers (peer_pos != 0): the records of block J are read from the rank that owns it
    // (pos_peer[J / bpr]: its copy is final when the per-pass barrier has passed), not from this rank's gathered copy
    double4 const* pos_peer[kPairMaxPeers];
    int peer_pos;
    double eps2;
    // extents of the buffers above, in elements (ESSA_BOUNDS builds check every computed address against them)
    size_t n_part_i, n_slots, n_part_k, n_slots_k;
    int n_world; // bodies (what a key's index and perm[] entries must stay below)
};

// SORTED kernels keep MINUS the threshold of a target (what filter mode 2 adds to L(y0)); "nothing passes" becomes INT_MIN
__device__ __forceinline__ unsigned thr_neg(unsigned t) { return t == kPriorNever ? 0x80000000u : 0u - t; }
__device__ __forceinline__ unsigned thr_pos(unsigned x) { return x == 0x80000000u ? kPriorNever : 0u - x; }

__device__ __forceinline__ double rot1(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// UNIFORM: every body of the world is live and has the same gravity factor g; the kernel then accumulates
// sum |d|^-3 d and k_pair_finish multiplies by g once per body -- 18 instead of 20 FP64 instructions per pair.
// FILT: the candidate filter of most-attracting tracking.  0: none.  1: both sides, guarded by "partner not lighter" (any
// layout).  2 / 3: mass-sorted layout, off-diagonal tile whose J block lies above / below the I block in the sorted order, so
// that every visitor is at least / at most as heavy as every target: only the targets (2) / only the visitors (3) can gain a
// candidate, no guard, and the test of the eight pairs of a rotation is a running maximum -- ONE fused integer add-max per pair
// (SASS VIADDMNMX) and one compare per rotation:
//   2:  lg_j + L(y0) >= thr_i   for some target   <=>   max_r (L(y0_r) - thr_i[r]) >= -lg_j      (thi holds -thr_i here)
//   3:  lg_i + L(y0) >= thr_j   for some target   <=>   max_r (L(y0_r) + lg_i[r])  >=  thr_j
template<bool TWO_SIDED, int FILT, bool UNIFORM>
__device__ __forceinline__ void pair_interact(double xi, double yi, double zi, double gi, double xj, double yj, double zj, double gj,
    double eps2, double& aix, double& aiy, double& aiz, double& ajx, double& ajy, double& ajz,
    unsigned lgi, unsigned lgj, unsigned thi, unsigned thj, bool& anyp, unsigned& mrun, bool first) {
    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    // Off-diagonal tiles hold no self pair, so the r = 0 test of rsqrt_seed (ISETP + SEL per pair, in the dependent
    // chain between the MUFU seed and the first DMUL) is left out there: 753 -> 719 ms per pass at 1,048,576 bodies.
    // Two distinct bodies at one position then yield NaN for exactly those two bodies (0 * inf), which
    // k_pair_finish detects and repairs with a guarded one-sided sum -- the documented result (an r = 0 pair
    // contributes nothing) at no cost to the worlds that have no such pair.
    double y0;
    if (TWO_SIDED) {
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
        y0 = __hiloint2double(__double2hiint(y0), 0);
    } else {
        y0 = rsqrt_seed(r2);
    }
    double y2 = y0 * y0; // exact: the seed has 21 significant bits
    double e = fma(-r2, y2, 1.0);
    double c = y2 * y0;
    double w = fma(1.875, e, 1.5);
    double we = w * e;
    double u = fma(c, we, c); // |d|^-3
    double si = UNIFORM ? u : gj * u;
    aix = fma(si, dx, aix);
    aiy = fma(si, dy, aiy);
    aiz = fma(si, dz, aiz);
    double sj = 0.0;
    if (TWO_SIDED) {
        sj = UNIFORM ? u : gi * u;
        ajx = fma(-sj, dx, ajx);
        ajy = fma(-sj, dy, ajy);
        ajz = fma(-sj, dz, ajz);
    }
    if (FILT == 2) {
        int const ly = __double2hiint(y0);
        mrun = first ? (unsigned)(ly + (int)thi) : (unsigned)__viaddmax_s32(ly, (int)thi, (int)mrun);
    }
    if (FILT == 3) {
        unsigned const ly = (unsigned)__double2hiint(y0);
        mrun = first ? ly + lgi : __viaddmax_u32(ly, lgi, mrun);
    }
    if (FILT == 1) {
        // candidate filter; the exact tests run after the 32 rotations of a visit, behind ONE warp-uniform branch
        // (pair_exact below): a branch per pair would make every pair its own basic block and serialise the
        // eight independent chains of a rotation, a branch per rotation still drains the pipeline every 8 pairs
        // With A = lg_j + L(y0) and B = lg_i + L(y0):  "A >= thr_i and lg_j >= lg_i"  is  A >= max(thr_i, B)
        // (integer add-max, add, compare: six instructions per pair for both sides).
        unsigned const ly = (unsigned)__double2hiint(y0);
        unsigned const a = lgj + ly, b = lgi + ly;
        bool const pi = a >= max(b, thi);
        bool const pj = TWO_SIDED && b >= max(a, thj);
        anyp = anyp || pi || pj;
    }
}

// Cold path: some lane saw a pair pass the filter during this visit of a 32-body subset.  Re-evaluates the pair
// with the very same operations (so the seed, and with it the key, are the ones the hot path saw) and applies
// Object.cpp:84-94 -- strictly heavier partner, larger key -- to both bodies.
template<bool TWO_SIDED>
__device__ __forceinline__ void pair_exact(double4 const* __restrict__ pos, int const* __restrict__ perm, double xi, double yi, double zi,
    double gi, double xj, double yj, double zj, double gj, double eps2, int iidx, int jidx, unsigned long long& bi, unsigned long long& bj,
    unsigned lgi, unsigned lgj, unsigned& thi, unsigned& thj) {
    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double y0;
    if (TWO_SIDED) {
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
        y0 = __hiloint2double(__double2hiint(y0), 0);
    } else {
        y0 = rsqrt_seed(r2);
    }
    unsigned const ly = (unsigned)__double2hiint(y0);
    bool const pi = lgj + ly >= thi && gj > gi;
    bool const pj = TWO_SIDED && lgi + ly >= thj && gi > gj;
    if (pi || pj) {
        double y2 = y0 * y0;
        double e = fma(-r2, y2, 1.0);
        double c = y2 * y0;
        double w = fma(1.875, e, 1.5);
        double we = w * e;
        double u = fma(c, we, c);
        // ranked on gf_p |d|^-3 (1/r) with the REFINED reciprocal (relative error ~1e-15), not on the raw seed: two heavier
        // partners within 1e-6 of each other must not swap (the filter's margin covers the seed's error: it only decides
        // who gets here)
        double const yr = fma(y0 * e, fma(0.375, e, 0.5), y0);
        if (perm) { // mass-sorted layout: keys name bodies by their ORIGINAL index (lowest index wins a tie; most[] is an index)
            iidx = perm[iidx];
            jidx = perm[jidx];
            ESSA_CHECK(iidx >= 0 && jidx >= 0); // a padding slot (-1) has gf 0: it can be neither heavier nor lighter-with-a-candidate
        }
        if (pi) {
            double si = gj * u;
            unsigned long long const k = pack_key(si * yr, jidx);
            unsigned long long const nb_ = key_merge(pos, iidx, bi, k, eps2);
            if (nb_ != bi) {
                bi = nb_;
                thi = max(thi, key_thr(nb_));
            }
        } else {
            double sj = gi * u;
            unsigned long long const k = pack_key(sj * yr, iidx);
            unsigned long long const nb_ = key_merge(pos, jidx, bj, k, eps2);
            if (nb_ != bj) {
                bj = nb_;
                thj = max(thj, key_thr(nb_));
            }
        }
    }
}

// SORTED (only with ARGMAX; single GPU): the MASS-SORTED layout.  The pass runs on a copy of the records sorted by gravity
// factor (ascending, equal ones by index; padding first), so "is the partner heavier" (src/Object.cpp:86,91) stops being a
// question per pair and becomes a property of the TILE: in tile (I, J) with J above I every visitor is at least as heavy as
// every target -- only targets can gain a candidate -- and below I it is the other way round.  The filter of such a tile is
// pair_interact's mode 2 / 3: one VIADDMNMX per pair + one compare per rotation instead of 2 IADD + 2 VIMNMX + 2 ISETP per
// pair, and only one of {lg_j, thr_j} travels with the visitor.  A unit whose heaviest possible partner is not heavier than its
// lightest possible target (worlds with a few mass classes: whole tiles of one class) runs without any filter.  The diagonal
// tile keeps the guarded filter.  Keys carry ORIGINAL indices (perm), so ties and most[] are unaffected by the layout.
template<int kPRI, bool ARGMAX, bool UNIFORM, bool SORTED>
__global__ void __launch_bounds__(kPThreads, 1) k_force_paired(PairArgs const A) {
    static_assert(!SORTED || ARGMAX, "the sorted layout exists for the candidate filter");
    constexpr int kPB = kPThreads * kPRI;
    constexpr int kPParts = kPB / kPHalf;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double4* tile = reinterpret_cast<double4*>(smem_raw);                       // [2][kPHalf]
    double* accj = reinterpret_cast<double*>(smem_raw + 2 * kPHalf * sizeof(double4)); // [8][3][kPHalf]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(accj + 8 * 3 * kPHalf); // [2]
    unsigned long long* acck = full + 8;                                                        // [8][kPHalf] (ARGMAX)
    unsigned* tprior = reinterpret_cast<unsigned*>(acck + 8 * kPHalf);                          // [2][kPHalf] (ARGMAX)
    // best candidate key per target: touched by the cold path only, so it lives in shared memory ([r][thread]: conflict free)
    unsigned long long* bis = reinterpret_cast<unsigned long long*>(tprior + 2 * kPHalf);       // [kPRI][kPThreads] (ARGMAX)

    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const ibl = blockIdx.x / A.S, chunk = blockIdx.x - ibl * A.S;
    int const I = A.ib0 + ibl;
    // work units of block I: (offset d = 0 .. dmax) x (kPParts parts of the J block), split into S chunks
    int const nunits = A.dn * kPParts;
    int const ulo = (int)((long long)nunits * chunk / A.S), uhi = (int)((long long)nunits * (chunk + 1) / A.S);
    int const* const perm = SORTED ? A.perm : nullptr;
    double4 const* const kpos = SORTED ? A.pos_orig : A.pos; // what candidate keys index

    double xi[kPRI], yi[kPRI], zi[kPRI], gi[kPRI], aix[kPRI], aiy[kPRI], aiz[kPRI];
#pragma unroll
    for (int r = 0; r < kPRI; r++) {
        ESSA_CHECK(I >= 0 && I < A.nb && (size_t)I * kPB + warp * (32 * kPRI) + r * 32 + lane < (size_t)A.npad);
        double4 p = A.pos[(size_t)I * kPB + warp * (32 * kPRI) + r * 32 + lane];
        xi[r] = p.x;
        yi[r] = p.y;
        zi[r] = p.z;
        gi[r] = p.w;
        aix[r] = aiy[r] = aiz[r] = 0.0;
    }
    // the filter's view of the target's gf and of its best key; SORTED keeps MINUS the threshold (what mode 2 adds)
    unsigned lgi[kPRI], thx[kPRI];
#pragma unroll
    for (int r = 0; r < kPRI; r++) {
        lgi[r] = key_lg(gi[r]);
        unsigned const t = ARGMAX ? A.prior[(size_t)I * kPB + warp * (32 * kPRI) + r * 32 + lane] : 0u;
        thx[r] = SORTED ? thr_neg(t) : t;
        if (ARGMAX)
            bis[r * kPThreads + tid] = 0ull;
    }
    int const ibase = I * kPB + warp * (32 * kPRI) + lane; // + r * 32
    // gravity factors at the two ends of the I block (sorted: the lightest and the heaviest target)
    double gi_first = 0.0, gi_last = 0.0;
    if (SORTED) {
        gi_first = A.pos[(size_t)I * kPB].w;
        gi_last = A.pos[(size_t)I * kPB + kPB - 1].w;
    }

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tiles of this CTA in order: (d, half); skipped d's never appear
    auto valid = [&](int d) { return !((A.nb & 1) == 0 && d == A.nb / 2 && I >= A.nb / 2) && d <= A.dmax; };
    auto issue = [&](int d, int half, int buf) {
        int J = I + d;
        if (J >= A.nb)
            J -= A.nb;
        ESSA_CHECK(J >= 0 && J < A.nb && half >= 0 && half < kPParts && (size_t)J * kPB + (size_t)(half + 1) * kPHalf <= (size_t)A.npad);
        unsigned bytes = kPHalf * (unsigned)sizeof(double4);
        mbar_expect_tx(&full[buf], bytes + (ARGMAX ? kPHalf * (unsigned)sizeof(unsigned) : 0u));
        double4 const* const from = A.peer_pos ? A.pos_peer[J / A.bpr] : A.pos;
        bulk_g2s(tile + (size_t)buf * kPHalf, from + (size_t)J * kPB + half * kPHalf, bytes, &full[buf]);
        if (ARGMAX)
            bulk_g2s(tprior + (size_t)buf * kPHalf, A.prior + (size_t)J * kPB + half * kPHalf, kPHalf * (unsigned)sizeof(unsigned), &full[buf]);
    };

    int seq = 0; // tiles consumed so far (parity bookkeeping)
    // prime the pipeline with the first valid tile
    int nu = ulo;
    while (nu < uhi && !valid(A.d0 + nu / kPParts))
        nu++;
    if (tid == 0 && nu < uhi)
        issue(A.d0 + nu / kPParts, nu % kPParts, 0);

    for (int u = ulo; u < uhi; u++) {
        int const d = A.d0 + u / kPParts, half = u % kPParts;
        if (!valid(d))
            continue;
        int J = I + d;
        if (J >= A.nb)
            J -= A.nb;
        {
            int const buf = seq & 1;
            // prefetch the tile after this one into the other buffer
            {
                int pu = u + 1;
                while (pu < uhi && !valid(A.d0 + pu / kPParts))
                    pu++;
                if (tid == 0 && pu < uhi)
                    issue(A.d0 + pu / kPParts, pu % kPParts, buf ^ 1);
            }
            mbar_wait(&full[buf], (unsigned)(seq >> 1) & 1u);
            double4 const* tp = tile + (size_t)buf * kPHalf;
            double* myacc = accj + (size_t)warp * 3 * kPHalf;
            // which filter this unit runs (warp-uniform, one decision per 2048 x 512 pairs)
            int mode = ARGMAX ? 1 : 0;
            if (SORTED && d > 0) {
                bool const above = J > I;
                bool const cand = above ? tp[kPHalf - 1].w > gi_first : gi_last > tp[0].w;
                mode = !cand ? 0 : (above ? 2 : 3);
            }

            for (int s = 0; s < kPHalf / kPSub; s++) {
                int const sub = (s + warp) & (kPHalf / kPSub - 1); // stagger shared-memory traffic between warps
                double xj, yj, zj, gj, ajx = 0.0, ajy = 0.0, ajz = 0.0;
                {
                    double4 p = tp[sub * kPSub + lane];
                    xj = p.x;
                    yj = p.y;
                    zj = p.z;
                    gj = p.w;
                }
                unsigned long long bjk = 0ull; // the visitor's own candidate key (cold path only; travels with it there)
                unsigned lgj = key_lg(gj);
                unsigned thj = ARGMAX ? tprior[(size_t)buf * kPHalf + sub * kPSub + lane] : 0u;
                int const jbase = J * kPB + half * kPHalf + sub * kPSub; // + home lane of the visitor
                int const src = (lane + 31) & 31; // take from the previous lane = pass to the next
                // Rotations of this visit in which some pair of this lane passed the candidate filter (bit k = rotation k).  The
                // exact tests run after the 32 rotations, behind ONE warp-uniform branch, and only for those rotations (rewalk):
                // a pass is rare (a few per body and force pass), a branch per rotation would drain the pipeline every 8 pairs
                // (ncu: FP64 pipe 64 % busy, "wait" stalls 1.5 per issue), and walking all 32 rotations again for one pass cost
                // 12 % of the kernel (2.6 % of the visits at 4.5 x the instructions of a plain visit).
                unsigned rmask = 0u;
                // the exact tests of the flagged rotations.  The visitors are back in their home lanes; rotation k had lane L meet
                // the visitor of home lane (L - k) mod 32, whose key and threshold are fetched from there and sent back after.
                auto rewalk = [&](auto two_sided) {
                    constexpr bool TWO = decltype(two_sided)::value;
                    unsigned todo = __reduce_or_sync(0xffffffffu, rmask);
                    while (todo) {
                        int const rot = __ffs(todo) - 1;
                        todo &= todo - 1u;
                        int const from = (lane - rot) & 31, back = (lane + rot) & 31;
                        double const xr = rot1(xj, from), yr = rot1(yj, from), zr = rot1(zj, from), gr = rot1(gj, from);
                        unsigned long long br = __shfl_sync(0xffffffffu, bjk, from);
                        unsigned tr = __shfl_sync(0xffffffffu, thj, from);
                        unsigned const lr = key_lg(gr);
#pragma unroll
                        for (int r = 0; r < kPRI; r++) {
                            unsigned th = SORTED ? thr_pos(thx[r]) : thx[r];
                            pair_exact<TWO>(kpos, perm, xi[r], yi[r], zi[r], gi[r], xr, yr, zr, gr, A.eps2, ibase + r * 32, jbase + from,
                                bis[r * kPThreads + tid], br, lgi[r], lr, th, tr);
                            thx[r] = SORTED ? thr_neg(th) : th;
                        }
                        bjk = __shfl_sync(0xffffffffu, br, back);
                        thj = __shfl_sync(0xffffffffu, tr, back);
                    }
                };
                if (d == 0) {
                    for (int rot = 0; rot < 32; rot++) {
                        unsigned mrun = 0u;
                        bool prot = false;
#pragma unroll
                        for (int r = 0; r < kPRI; r++)
                            pair_interact<false, ARGMAX ? 1 : 0, UNIFORM>(xi[r], yi[r], zi[r], gi[r], xj, yj, zj, gj, A.eps2, aix[r], aiy[r], aiz[r],
                                ajx, ajy, ajz, lgi[r], lgj, SORTED ? thr_pos(thx[r]) : thx[r], thj, prot, mrun, r == 0);
                        if (ARGMAX)
                            rmask = (rmask >> 1) | (prot ? 0x80000000u : 0u);
                        xj = rot1(xj, src);
                        yj = rot1(yj, src);
                        zj = rot1(zj, src);
                        gj = rot1(gj, src);
                        if (ARGMAX)
                            lgj = __shfl_sync(0xffffffffu, lgj, src);
                    }
                    if (ARGMAX)
                        rewalk(std::false_type {});
                } else {
                    // the 32 rotations of a visit, for one filter mode (a compile-time constant inside the loop)
                    auto visit = [&](auto fc) {
                        constexpr int FILT = decltype(fc)::value;
                        for (int rot = 0; rot < 32; rot++) {
                            unsigned mrun = 0u;
                            bool prot = false;
#pragma unroll
                            for (int r = 0; r < kPRI; r++)
                                pair_interact<true, FILT, UNIFORM>(xi[r], yi[r], zi[r], gi[r], xj, yj, zj, gj, A.eps2, aix[r], aiy[r], aiz[r], ajx,
                                    ajy, ajz, lgi[r], lgj, thx[r], thj, prot, mrun, r == 0);
                            if (FILT == 2)
                                prot = (int)mrun >= -(int)lgj;
                            if (FILT == 3)
                                prot = mrun >= thj;
                            if (FILT != 0)
                                rmask = (rmask >> 1) | (prot ? 0x80000000u : 0u);
                            xj = rot1(xj, src);
                            yj = rot1(yj, src);
                            zj = rot1(zj, src);
                            gj = rot1(gj, src);
                            ajx = rot1(ajx, src);
                            ajy = rot1(ajy, src);
                            ajz = rot1(ajz, src);
                            // what the filter needs of the visitor; its key stays at home
                            if (FILT == 1 || FILT == 2)
                                lgj = __shfl_sync(0xffffffffu, lgj, src);
                            if (FILT == 1 || FILT == 3)
                                thj = __shfl_sync(0xffffffffu, thj, src);
                        }
                    };
                    if (!ARGMAX || (SORTED && mode == 0))
                        visit(std::integral_constant<int, 0> {});
                    else if (!SORTED)
                        visit(std::integral_constant<int, 1> {});
                    else if (mode == 2)
                        visit(std::integral_constant<int, SORTED ? 2 : 1> {});
                    else
                        visit(std::integral_constant<int, SORTED ? 3 : 1> {});
                    if (ARGMAX)
                        rewalk(std::true_type {});
                    {
                        int j = sub * kPSub + lane;
                        myacc[j] = ajx;
                        myacc[kPHalf + j] = ajy;
                        myacc[2 * kPHalf + j] = ajz;
                        if (ARGMAX)
                            acck[(size_t)warp * kPHalf + j] = bjk;
                    }
                }
            }
            __syncthreads();
            if (d > 0) {
                int const owner = A.owner_major ? J / A.bpr : 0;
                size_t const tile = A.owner_major ? (size_t)(J - owner * A.bpr) * A.dslots + (d - A.dslot0) : (size_t)ibl * A.dslots + (d - A.dslot0);
                double* slot = (A.owner_major ? A.slot_peer[owner] : A.slots) + tile * 3 * kPB;
                ESSA_CHECK(d >= A.dslot0 && d - A.dslot0 < A.dslots && owner >= 0 && owner < kPairMaxPeers && (tile + 1) * 3 * kPB <= A.n_slots);
                for (int e = tid; e < 3 * kPHalf; e += kPThreads) {
                    int c = e / kPHalf, j = e - c * kPHalf;
                    double sum = 0.0;
#pragma unroll
                    for (int w = 0; w < 8; w++)
                        sum = __dadd_rn(sum, accj[(size_t)w * 3 * kPHalf + e]);
                    __stcg(slot + (size_t)c * kPB + half * kPHalf + j, sum);
                }
                if (ARGMAX) {
                    unsigned long long* slotk = (A.owner_major ? A.slotk_peer[owner] : A.slots_k) + tile * kPB + half * kPHalf;
                    ESSA_CHECK((tile + 1) * kPB <= A.n_slots_k);
                    for (int j = tid; j < kPHalf; j += kPThreads) {
                        unsigned long long k = 0ull;
                        int const tgt = SORTED ? perm[J * kPB + half * kPHalf + j] : J * kPB + half * kPHalf + j;
                        ESSA_CHECK(tgt >= -1 && tgt < (SORTED ? A.n_world : A.npad));
#pragma unroll
                        for (int w = 0; w < 8; w++)
                            k = key_merge(kpos, tgt, k, acck[(size_t)w * kPHalf + j], A.eps2);
                        __stcg(slotk + j, k);
                    }
                }
            }
            __syncthreads();
            seq++;
        }
    }

#pragma unroll
    for (int r = 0; r < kPRI; r++) {
        size_t i = (size_t)I * kPB + warp * (32 * kPRI) + r * 32 + lane;
        double* p = A.part_i + (size_t)chunk * 3 * A.npad + i;
        ESSA_CHECK(chunk >= 0 && chunk < A.S && i < (size_t)A.npad && (size_t)(chunk + 1) * 3 * A.npad <= A.n_part_i);
        ESSA_CHECK(!ARGMAX || (size_t)(chunk + 1) * A.npad <= A.n_part_k);
        __stcg(p, aix[r]);
        __stcg(p + A.npad, aiy[r]);
        __stcg(p + 2 * (size_t)A.npad, aiz[r]);
        if (ARGMAX)
            __stcg(A.part_k + (size_t)chunk * A.npad + i, bis[r * kPThreads + tid]);
    }
}

// Ordered sum of the I-side partials and the J-side slots of one body, then the K1 epilogue.
struct PairFinishArgs {
    WorldPtrs w;
    KickArgs k;
    int n, i0, i1;
    int nb, dmax, S, ib0, ib1, npad;
    double const* part_i;
    double const* slots;
    double const* ajred; // sharded: [3][npad] J-side sums already reduced over ranks (slots are skipped)
    unsigned long long const* part_k;  // ARGMAX: I-side candidate keys [S][npad]
    unsigned long long const* slots_k; // ARGMAX: J-side candidate keys per tile
    unsigned long long const* akred;   // sharded ARGMAX: [world][i1 - i0] every rank's J-side key for this rank's bodies (an
                                       // ncclMax over the ranks could not re-decide two keys that differ in the index only)
    int world;
    int argmax;          // most-attracting tracking with real candidates (keys)
    int argmax_noop;     // most-attracting tracking requested on a world of equal gravity factors: clear_forces only
    int save_old;        // ... and Object::before_update's old_most = most
    unsigned* prior;     // ARGMAX: [npad] filter thresholds, read (what the pass used) and rewritten (for the next pass)
    unsigned prior_slack; // how far below the winner's threshold the next pass starts its filter (kPriorSlack)
    int const* perm;     // mass-sorted pass: thread s handles body perm[s] (-1: padding); partials, slots, priors are indexed by s
    double scale;        // UNIFORM worlds: the common gravity factor the sums still lack; otherwise 1
    double eps2;         // softening, for the repair of bodies that met an r = 0 pair in an unguarded tile
};

// J-side contributions to body g (block K = g / kPB, owned by this context) from every tile (I = K - d, d), in offset
// order, out of the owner-major slots of this context.
template<int kPB>
__device__ __forceinline__ void slot_sum_owner(int g, int nb, int dmax, int ib0, double const* slots, double& ax, double& ay, double& az) {
    int const K = g / kPB, jl = g - K * kPB;
    for (int d = 1; d <= dmax; d++) {
        int I = K - d;
        if (I < 0)
            I += nb;
        if ((nb & 1) == 0 && d == nb / 2 && I >= nb / 2)
            continue; // that tile belongs to the other half of the ring
        double const* s = slots + ((size_t)(K - ib0) * dmax + (d - 1)) * 3 * kPB + jl;
        ax = __dadd_rn(ax, __ldcg(s));
        ay = __dadd_rn(ay, __ldcg(s + kPB));
        az = __dadd_rn(az, __ldcg(s + 2 * kPB));
    }
}
template<int kPB>
__device__ __forceinline__ unsigned long long slot_key_owner(int s, int g, int nb, int dmax, int ib0, unsigned long long const* slots_k,
    double4 const* __restrict__ pos, double eps2) {
    int const K = s / kPB, jl = s - K * kPB; // s: where the body's slots are (its place in the pass's layout); g: its index
    unsigned long long k = 0ull;
    for (int d = 1; d <= dmax; d++) {
        int I = K - d;
        if (I < 0)
            I += nb;
        if ((nb & 1) == 0 && d == nb / 2 && I >= nb / 2)
            continue;
        k = key_merge(pos, g, k, __ldcg(slots_k + ((size_t)(K - ib0) * dmax + (d - 1)) * kPB + jl), eps2);
    }
    return k;
}

// (writer-major fallback) J-side contributions to body g from the tiles whose I block lies in [ib0, ib1), in offset order.
template<int kPB>
__device__ __forceinline__ void slot_sum(int g, int nb, int dmax, int ib0, int ib1, double const* slots, double& ax, double& ay,
    double& az) {
    int const K = g / kPB, jl = g - K * kPB;
    for (int d = 1; d <= dmax; d++) {
        int I = K - d;
        if (I < 0)
            I += nb;
        if ((nb & 1) == 0 && d == nb / 2 && I >= nb / 2)
            continue; // that tile belongs to the other half of the ring
        if (I < ib0 || I >= ib1)
            continue;
        double const* s = slots + ((size_t)(I - ib0) * dmax + (d - 1)) * 3 * kPB + jl;
        ax = __dadd_rn(ax, __ldcg(s));
        ay = __dadd_rn(ay, __ldcg(s + kPB));
        az = __dadd_rn(az, __ldcg(s + 2 * kPB));
    }
}

// J-side candidate keys of body g from the tiles whose I block lies in [ib0, ib1).
template<int kPB>
__device__ __forceinline__ unsigned long long slot_key(int g, int nb, int dmax, int ib0, int ib1, unsigned long long const* slots_k,
    double4 const* __restrict__ pos, double eps2) {
    int const K = g / kPB, jl = g - K * kPB;
    unsigned long long k = 0ull;
    for (int d = 1; d <= dmax; d++) {
        int I = K - d;
        if (I < 0)
            I += nb;
        if (I < ib0 || I >= ib1)
            continue;
        if ((nb & 1) == 0 && d == nb / 2 && I >= nb / 2)
            continue;
        k = key_merge(pos, g, k, __ldcg(slots_k + ((size_t)(I - ib0) * dmax + (d - 1)) * kPB + jl), eps2);
    }
    return k;
}

// Batched offsets (single GPU, slots over budget): the running sum of a body's I-side partials and J-side slots, batch after
// batch, every addition in a fixed order (chunks ascending, then offsets ascending); k_pair_finish then takes the sum as it is.
template<int kPB>
__global__ void __launch_bounds__(256) k_pair_fold(int nb, int npad, int S, int d_first, int d_last, int dslots, int dslot0, int first_batch,
    double const* part_i, double const* slots, double* jacc) {
    int const g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= npad)
        return;
    double ax = first_batch ? 0.0 : jacc[g], ay = first_batch ? 0.0 : jacc[(size_t)npad + g], az = first_batch ? 0.0 : jacc[2 * (size_t)npad + g];
    for (int c = 0; c < S; c++) {
        double const* q = part_i + (size_t)c * 3 * npad + g;
        ax = __dadd_rn(ax, __ldcg(q));
        ay = __dadd_rn(ay, __ldcg(q + npad));
        az = __dadd_rn(az, __ldcg(q + 2 * (size_t)npad));
    }
    int const K = g / kPB, jl = g - K * kPB;
    for (int d = d_first < 1 ? 1 : d_first; d <= d_last; d++) {
        int I = K - d;
        if (I < 0)
            I += nb;
        if ((nb & 1) == 0 && d == nb / 2 && I >= nb / 2)
            continue; // that tile belongs to the other half of the ring
        double const* s = slots + ((size_t)K * dslots + (d - dslot0)) * 3 * kPB + jl;
        ax = __dadd_rn(ax, __ldcg(s));
        ay = __dadd_rn(ay, __ldcg(s + kPB));
        az = __dadd_rn(az, __ldcg(s + 2 * kPB));
    }
    jacc[g] = ax;
    jacc[(size_t)npad + g] = ay;
    jacc[2 * (size_t)npad + g] = az;
}

// Sharded worlds: this rank's J-side sums for EVERY body, to be reduce-scattered over the ranks.
template<int kPB>
__global__ void __launch_bounds__(256) k_pair_gather_slots(int nb, int dmax, int ib0, int ib1, int npad, double const* slots,
    double* acc, unsigned long long const* slots_k, unsigned long long* kacc, double4 const* __restrict__ pos, double eps2) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= npad)
        return;
    double ax = 0, ay = 0, az = 0;
    slot_sum<kPB>(g, nb, dmax, ib0, ib1, slots, ax, ay, az);
    acc[g] = ax;
    acc[(size_t)npad + g] = ay;
    acc[2 * (size_t)npad + g] = az;
    if (kacc)
        kacc[g] = slot_key<kPB>(g, nb, dmax, ib0, ib1, slots_k, pos, eps2);
}

// Guarded one-sided sum over all n sources for the body of thread `t` of this block (its acceleration came out
// non-finite: it shares its position with another body in a different block, where the pair kernel runs without
// the r = 0 test).  All 256 threads walk the sources, the partial sums meet in shared memory and are added in a
// fixed order, so the result is deterministic; sources at r = 0 contribute nothing.  ~0.1 ms per repaired body at
// 1,048,576 bodies, and only the (two) coincident bodies ever need it.
__device__ __noinline__ void pair_repair_body(double4 const* __restrict__ pos, int n, double4 const p, double eps2, bool argmax,
    double (*red)[256], int* redj, double& ax, double& ay, double& az, unsigned long long& key) {
    int const tid = threadIdx.x;
    double sx = 0, sy = 0, sz = 0, best = 0.0;
    int best_j = -1;
    for (int j = tid; j < n; j += 256) {
        double4 const q = pos[j];
        if (argmax)
            fast_interact_argmax(p.x, p.y, p.z, p.w, q, j, eps2, sx, sy, sz, best, best_j);
        else
            fast_interact(p.x, p.y, p.z, q, eps2, sx, sy, sz);
    }
    red[0][tid] = sx;
    red[1][tid] = sy;
    red[2][tid] = sz;
    red[3][tid] = best;
    redj[tid] = best_j;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) {
            red[0][tid] = __dadd_rn(red[0][tid], red[0][tid + w]);
            red[1][tid] = __dadd_rn(red[1][tid], red[1][tid + w]);
            red[2][tid] = __dadd_rn(red[2][tid], red[2][tid + w]);
            double const m = red[3][tid + w];
            int const mj = redj[tid + w];
            // largest metric; the lowest index among equal ones (Object.cpp:86 is a strict '>')
            if (mj >= 0 && (m > red[3][tid] || (m == red[3][tid] && (redj[tid] < 0 || mj < redj[tid])))) {
                red[3][tid] = m;
                redj[tid] = mj;
            }
        }
        __syncthreads();
    }
    ax = red[0][0];
    ay = red[1][0];
    az = red[2][0];
    key = redj[0] >= 0 ? pack_key(red[3][0], redj[0]) : 0ull;
    __syncthreads();
}

template<int kPB>
__global__ void __launch_bounds__(256, 4) k_pair_finish(PairFinishArgs const A) {
    __shared__ double red[4][256];
    __shared__ int redj[256];
    __shared__ unsigned char sbad[256];
    __shared__ int sg[256];
    int const s = A.i0 + blockIdx.x * blockDim.x + threadIdx.x; // place in the layout of the pass
    int g = s;                                                  // the body
    bool in = s < A.i1;
    if (A.perm) {
        g = in ? A.perm[s] : -1;
        in = g >= 0;
    }
    sg[threadIdx.x] = g;
    WorldPtrs const& W = A.w;
    KickArgs const& Kk = A.k;
    double ax = 0, ay = 0, az = 0;
    unsigned long long key = 0ull;
    bool act = false;
    double4 p = make_double4(0, 0, 0, 0);
    if (in) {
        for (int c = 0; c < A.S; c++) {
            double const* q = A.part_i + (size_t)c * 3 * A.npad + s;
            ax = __dadd_rn(ax, __ldcg(q));
            ay = __dadd_rn(ay, __ldcg(q + A.npad));
            az = __dadd_rn(az, __ldcg(q + 2 * (size_t)A.npad));
        }
        if (A.ajred) {
            ax = __dadd_rn(ax, __ldcg(A.ajred + s));
            ay = __dadd_rn(ay, __ldcg(A.ajred + (size_t)A.npad + s));
            az = __dadd_rn(az, __ldcg(A.ajred + 2 * (size_t)A.npad + s));
        } else {
            slot_sum_owner<kPB>(s, A.nb, A.dmax, A.ib0, A.slots, ax, ay, az);
        }
        ax = __dmul_rn(ax, A.scale);
        ay = __dmul_rn(ay, A.scale);
        az = __dmul_rn(az, A.scale);
        act = W.active[g] != 0;
        p = W.pos_cur[g];
        if (act && A.argmax) {
            if (A.akred) {
                for (int r = 0; r < A.world; r++)
                    key = key_merge(W.pos_cur, g, key, __ldcg(A.akred + (size_t)r * (A.i1 - A.i0) + (g - A.i0)), A.eps2);
            } else {
                key = slot_key_owner<kPB>(s, g, A.nb, A.dmax, A.ib0, A.slots_k, W.pos_cur, A.eps2);
            }
            for (int c = 0; c < A.S; c++)
                key = key_merge(W.pos_cur, g, key, __ldcg(A.part_k + (size_t)c * A.npad + s), A.eps2);
        }
    }
    // ---- repair: bodies that met a coincident partner in an unguarded tile (see pair_interact), and bodies whose
    // winning candidate does not clear the prior threshold the pass filtered with (a better partner may have been
    // skipped: its most-attracting body weakened by more than kPriorSlack since the last pass) ----
    bool const bad_acc = in && act && !(isfinite(ax) && isfinite(ay) && isfinite(az));
    bool bad = bad_acc;
    if (in && act && A.argmax) {
        unsigned const P = A.prior[s];
        bad = bad || !(P == kPriorNone || (key != 0ull && key_thr(key) > P));
    }
    if (__syncthreads_or(bad)) {
        sbad[threadIdx.x] = bad ? 1 : 0;
        __syncthreads();
        for (int t = 0; t < 256; t++) {
            if (!sbad[t])
                continue; // uniform over the block
            int const gt = sg[t];
            double rx, ry, rz;
            unsigned long long rk;
            pair_repair_body(W.pos_cur, A.n, W.pos_cur[gt], A.eps2, A.argmax != 0, red, redj, rx, ry, rz, rk);
            if (t == (int)threadIdx.x) {
                if (bad_acc) {
                    ax = rx;
                    ay = ry;
                    az = rz;
                }
                key = rk;
            }
        }
    }
    if (!in)
        return;
    if (A.argmax) // where the next pass starts its filters
        A.prior[s] = !act ? kPriorNever : (key != 0ull ? key_thr(key) - A.prior_slack : kPriorNone);
    // ---- epilogue: identical to force_tiled.cu finish_body ----
    double vx = 0, vy = 0, vz = 0;
    if ((A.argmax_noop || A.argmax) && A.save_old)
        W.old_most[g] = W.most[g];
    if (act) {
        if (A.argmax_noop)
            W.maxattr[g] = 0.0; // Object::clear_forces; no pair can raise it (nobody is heavier than anybody)
        if (A.argmax) {
            double m = 0.0; // Object::clear_forces
            if (key != 0ull) {
                // the winner's gf / r^4 (the reference's |ta|^2 / gf) to full precision
                int const wj = (int)(kKeyIdxMask - (key & kKeyIdxMask));
                double4 const q = W.pos_cur[wj];
                double dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
                double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                double y0 = rsqrt_seed(r2);
                double t = r2 * y0;
                double e = fma(-t, y0, 1.0);
                double y = fma(y0 * e, fma(0.375, e, 0.5), y0); // 1 / r
                double y2 = y * y;
                m = q.w * y2 * y2;
                W.most[g] = wj;
            }
            W.maxattr[g] = m;
        }
        W.a[0][g] = ax;
        W.a[1][g] = ay;
        W.a[2][g] = az;
        if (Kk.second_kick) {
            vx = kick1(W.vh[0][g], ax, Kk.h, Kk.mul);
            vy = kick1(W.vh[1][g], ay, Kk.h, Kk.mul);
            vz = kick1(W.vh[2][g], az, Kk.h, Kk.mul);
            W.v[0][g] = vx;
            W.v[1][g] = vy;
            W.v[2][g] = vz;
        } else if (Kk.advance) {
            vx = W.v[0][g];
            vy = W.v[1][g];
            vz = W.v[2][g];
        }
    }
    if (Kk.advance) {
        if (act) {
            vx = kick1(vx, ax, Kk.h, Kk.mul);
            vy = kick1(vy, ay, Kk.h, Kk.mul);
            vz = kick1(vz, az, Kk.h, Kk.mul);
            W.vh[0][g] = vx;
            W.vh[1][g] = vy;
            W.vh[2][g] = vz;
            p.x = __dadd_rn(p.x, __dmul_rn(__dmul_rn(vx, Kk.dt), Kk.mul));
            p.y = __dadd_rn(p.y, __dmul_rn(__dmul_rn(vy, Kk.dt), Kk.mul));
            p.z = __dadd_rn(p.z, __dmul_rn(__dmul_rn(vz, Kk.dt), Kk.mul));
        }
        W.pos_next[g] = p;
    }
}

// First tracked pass after an upload or a change of the active mask: no priors yet.
__global__ void __launch_bounds__(256) k_pair_prior_reset(unsigned* prior, uint8_t const* active, int n, int npad) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < npad)
        prior[g] = g < n && active[g] ? kPriorNone : kPriorNever;
}

// ---- mass-sorted layout (SORTED kernels) ----
// Sort key of a gravity factor: the IEEE bit pattern made monotone for the whole real line (negative values reversed).
__global__ void __launch_bounds__(256) k_pair_sort_keys(double4 const* pos, int n, unsigned long long* keys, int* vals) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n)
        return;
    unsigned long long const b = (unsigned long long)__double_as_longlong(pos[g].w);
    keys[g] = (b >> 63) ? ~b : b | 0x8000000000000000ull;
    vals[g] = g;
}
__global__ void __launch_bounds__(256) k_pair_perm_pad(int* perm, int count) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < count)
        perm[g] = -1;
}
// the records of a pass in sorted order; padding (lighter than everything, in front) far away as in the unsorted layout
__global__ void __launch_bounds__(256) k_pair_gather_sorted(double4 const* pos, int const* perm, double4* out, int npad) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npad)
        return;
    int const g = perm[s];
    out[s] = g >= 0 ? pos[g] : make_double4(kPadCoord, kPadCoord, kPadCoord, 0.0);
}
__global__ void __launch_bounds__(256) k_pair_prior_reset_sorted(unsigned* prior, int const* perm, uint8_t const* active, int npad) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < npad)
        prior[s] = perm[s] >= 0 && active[perm[s]] ? kPriorNone : kPriorNever;
}
// ESSA_PAIR_SORTED=0: the guarded filter on the unsorted layout (A/B measurements; what sharded worlds run)
static bool pair_sorted_allowed() { // read per pass (not cached): the tests run both layouts in one process
    char const* e = getenv("ESSA_PAIR_SORTED");
    return e ? atoi(e) != 0 : true;
}

// Can this world go through the pair-shared kernel, and with which block size?  Whole blocks per
// rank, no argmax, slots must fit.  Sharded: rank g owns I blocks [nb g / G, nb (g+1) / G) = its own targets.
// J-side slots of all offsets cost N * (nb / 2) * 24 bytes (6.4 GB at 1 M bodies, 26 GB at 2 M, 103 GB at 4 M).  Above this
// budget a single GPU runs the offsets in batches.  A constant, not a function of the free memory: the batching decides the
// order of the additions, and a world must give the same bits on every run.  (ESSA_PAIR_SLOT_BUDGET_MB: tests.)
static size_t pair_slot_budget() {
    static size_t v = 0;
    if (v == 0) {
        char const* e = getenv("ESSA_PAIR_SLOT_BUDGET_MB");
        v = e && atoll(e) > 0 ? (size_t)atoll(e) << 20 : (size_t)32 << 30;
    }
    return v;
}

static int pair_ri_override() {
    static int v = -1;
    if (v < 0) {
        char const* e = getenv("ESSA_PAIR_RI");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// Picks the block size (2048 or 1024 bodies) and the number of chunks S per I block.  A CTA works through
// whole units -- (offset d) x (512-record part of the J block) -- of equal size, one CTA per SM, so the
// estimate is  waves of working CTAs x units in the longest chunk x time per unit;  the 1024-body kernel
// amortises its shuffles over 4 instead of 8 targets per lane and is ~20 % slower per pair (measured),
// but its units are half as large.  Splitting by parts rather than by whole offsets is what fills the
// machine at mid sizes: 16,384 bodies become 8 blocks x 20 units = 144 working CTAs of one 2048 x 512
// unit each (one wave) instead of 144 CTAs of one 1024 x 1024 unit at the slower rate.
bool paired_plan(essa_ctx const* c, int* nb_out, int* S_out, size_t* slot_doubles, size_t* part_doubles, int* ri_out) {
    int n = c->n;
    if (c->plan.n == n && c->plan.world == c->world && c->plan.rank == c->rank && c->plan.cap == c->cap) { // same world as last time
        if (!c->plan.ok)
            return false;
        *nb_out = c->plan.nb;
        *S_out = c->plan.S;
        *slot_doubles = c->plan.slot_doubles;
        *part_doubles = c->plan.part_doubles;
        if (ri_out)
            *ri_out = c->plan.ri;
        return true;
    }
    c->plan.n = n;
    c->plan.world = c->world;
    c->plan.rank = c->rank;
    c->plan.cap = c->cap;
    c->plan.ok = false;
    long long const slots = c->sm_count; // 1 CTA per SM
    double best = -1;
    int best_ri = 0, best_S = 1, best_nb = 0, best_db = 0, best_batches = 1;
    for (int ri = 8; ri >= 4; ri -= 4) {
        if ((pair_ri_override() == 4 || pair_ri_override() == 8) && ri != pair_ri_override())
            continue;
        int const pb = kPThreads * ri, parts = pb / kPHalf;
        int const nb = (n + pb - 1) / pb;
        if (nb < 2 || (size_t)nb * pb > (size_t)c->cap)
            continue;
        if (c->world > 1 && (n % (pb * c->world)) != 0)
            continue;
        int const own = c->world == 1 ? nb : nb / c->world;
        int const ib0 = (int)((long long)nb * c->rank / c->world);
        int const dmax = nb / 2;
        // offsets per launch: all of them, or batches that keep the slot planes within the budget (single GPU only: a rank
        // of a sharded world holds 1 / world of the slots)
        int batches = 1, db = dmax;
        size_t const slot_bytes = (size_t)own * dmax * 3 * pb * sizeof(double);
        if (c->world == 1 && slot_bytes > pair_slot_budget() && dmax > 1) {
            batches = (int)((slot_bytes + pair_slot_budget() - 1) / pair_slot_budget());
            batches = batches > dmax ? dmax : batches;
            db = (dmax + batches - 1) / batches;
            batches = (dmax + db - 1) / db;
        }
        int const U = ((batches > 1 ? db : dmax) + 1) * parts; // units of (the largest) launch
        double const unit_time = (ri == 8 ? 2.0 : 1.2) * batches;
        int const smax = U < 64 ? U : 64;
        for (int S = 1; S <= smax; S++) {
            // working CTAs and the longest chunk, counting only valid units (for an even block count the
            // offset nb / 2 belongs to the lower half of the I blocks)
            long long working = 0;
            int longest = 0;
            long long const upper = (nb & 1) == 0 ? (ib0 + own > nb / 2 ? ib0 + own - (ib0 > nb / 2 ? ib0 : nb / 2) : 0) : 0;
            for (int half = 0; half < 2; half++) { // I blocks with / without the last offset
                long long const blocks = half ? upper : own - upper;
                if (blocks == 0)
                    continue;
                int const uvalid = half ? dmax * parts : U; // units [uvalid, U) are skipped
                for (int ch = 0; ch < S; ch++) {
                    int lo = (int)((long long)U * ch / S), hi = (int)((long long)U * (ch + 1) / S);
                    hi = hi < uvalid ? hi : uvalid;
                    if (hi > lo) {
                        working += blocks;
                        longest = hi - lo > longest ? hi - lo : longest;
                    }
                }
            }
            long long const waves = (working + slots - 1) / slots;
            double const cost = (double)waves * longest * unit_time;
            if (best < 0 || cost < best) {
                best = cost;
                best_ri = ri;
                best_S = S;
                best_nb = nb;
                best_db = db;
                best_batches = batches;
            }
        }
    }
    if (best < 0)
        return false;
    int const pb = kPThreads * best_ri;
    int const own = c->world == 1 ? best_nb : best_nb / c->world;
    c->plan.ok = true;
    c->plan.nb = *nb_out = best_nb;
    c->plan.S = *S_out = best_S;
    c->plan.db = best_db;
    c->plan.batches = best_batches;
    c->plan.slot_doubles = *slot_doubles = (size_t)own * best_db * 3 * pb;
    // + [3][npad] behind the partials: the sharded fallback's folded slots, or the running sum of a batched pass
    c->plan.part_doubles = *part_doubles = (size_t)best_S * 3 * best_nb * pb + (c->world > 1 || best_batches > 1 ? (size_t)3 * best_nb * pb : 0);
    c->plan.ri = best_ri;
    if (ri_out)
        *ri_out = best_ri;
    return true;
}

// Every body live for good (no creation / deletion instants) and one gravity factor for all: the UNIFORM kernels
// (ESSA_PAIR_UNIFORM=0 keeps the general ones, for A/B measurements).
static bool pair_uniform(essa_ctx const* c) {
    static int allow = -1;
    if (allow < 0) {
        char const* e = getenv("ESSA_PAIR_UNIFORM");
        allow = e ? atoi(e) != 0 : 1;
    }
    return allow && c->gf_uniform && c->events.empty();
}

static void pair_geometry(essa_ctx const* c, int nb, int* ib0, int* ib1) {
    *ib0 = (int)((long long)nb * c->rank / c->world);
    *ib1 = (int)((long long)nb * (c->rank + 1) / c->world);
}

template<int RI, bool ARGMAX, bool UNIFORM, bool SORTED = false>
static cudaError_t launch_pair_kernels(essa_ctx* c, PairArgs const& A, cudaStream_t st, double** acc, unsigned long long** kacc) {
    constexpr int PB = kPThreads * RI;
    constexpr size_t smem = kPSmem + (ARGMAX ? kPSmemKeys : 0);
    // per context: the attribute is per device.  (UNIFORM and SORTED exclude each other: slots 4 and 5 vs 6 and 7)
    bool& attr_set = c->attr_pair[(RI == 8 ? 0 : 1) + (ARGMAX ? 2 : 0) + (UNIFORM || SORTED ? 4 : 0)];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_force_paired<RI, ARGMAX, UNIFORM, SORTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        attr_set = true;
    }
    k_force_paired<RI, ARGMAX, UNIFORM, SORTED><<<(A.ib1 - A.ib0) * A.S, kPThreads, smem, st>>>(A);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    *acc = nullptr;
    *kacc = nullptr;
    if (c->world > 1 && !A.owner_major) {
        *acc = c->pair_part + (size_t)A.S * 3 * A.npad;
        if (ARGMAX)
            *kacc = c->pair_kpart + (size_t)A.S * A.npad;
        k_pair_gather_slots<PB><<<(A.npad + 255) / 256, 256, 0, st>>>(A.nb, A.dmax, A.ib0, A.ib1, A.npad, c->pair_slots, *acc,
            c->pair_kslots, *kacc, A.pos, A.eps2);
        c->launches++;
        e = cudaGetLastError();
    }
    return e;
}


static cudaError_t grow(void** p, size_t* have, size_t want_bytes) {
    if (want_bytes <= *have)
        return cudaSuccess;
    cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, want_bytes);
    if (e == cudaSuccess)
        *have = want_bytes;
    return e;
}

// Buffers of the pair-shared pass for the current plan.  *changed: the slot buffers (what sharded ranks map into each other's
// address space) were (re)allocated.
cudaError_t pair_reserve(essa_ctx* c, bool argmax, bool* changed) {
    int nb, S, ri;
    size_t slot_d, part_d;
    if (changed)
        *changed = false;
    if (!paired_plan(c, &nb, &S, &slot_d, &part_d, &ri))
        return cudaErrorInvalidConfiguration;
    if (slot_d > c->pair_slots_n) {
        cudaFree(c->pair_slots);
        c->pair_slots = nullptr;
        c->pair_slots_n = 0;
        cudaError_t e = cudaMalloc(&c->pair_slots, slot_d * sizeof(double));
        if (e != cudaSuccess)
            return e;
        c->pair_slots_n = slot_d;
        if (changed)
            *changed = true;
    }
    if (part_d > c->pair_part_n) {
        cudaFree(c->pair_part);
        c->pair_part = nullptr;
        c->pair_part_n = 0;
        cudaError_t e = cudaMalloc(&c->pair_part, part_d * sizeof(double));
        if (e != cudaSuccess)
            return e;
        c->pair_part_n = part_d;
    }
    if (argmax) { // candidate keys: one per slot body / per partial body (a third of the acceleration buffers)
        size_t const before = c->pair_kslots_bytes;
        cudaError_t e = grow((void**)&c->pair_kslots, &c->pair_kslots_bytes, slot_d / 3 * sizeof(unsigned long long));
        if (e == cudaSuccess)
            e = grow((void**)&c->pair_kpart, &c->pair_kpart_bytes, (part_d / 3 + 1) * sizeof(unsigned long long));
        if (e != cudaSuccess)
            return e;
        if (changed && c->pair_kslots_bytes != before)
            *changed = true;
    }
    return cudaSuccess;
}

// Would pair_reserve have to (re)allocate a slot buffer?  (Sharded ranks unmap each other's buffers first.)
bool pair_reserve_would_change(essa_ctx const* c, bool argmax) {
    int nb, S, ri;
    size_t slot_d, part_d;
    if (!paired_plan(c, &nb, &S, &slot_d, &part_d, &ri))
        return false;
    return slot_d > c->pair_slots_n || (argmax && slot_d / 3 * sizeof(unsigned long long) > c->pair_kslots_bytes);
}

// Phase 1: the pair kernel over this rank's I blocks; sharded worlds also fold their slots into
// acc[3][npad] (returned through *acc_out) for the caller to reduce-scatter over the ranks.
cudaError_t launch_force_paired_compute(essa_ctx* c, double eps2, cudaStream_t st, double** acc_out, int* npad_out, bool argmax,
    unsigned long long** kacc_out) {
    int nb, S, ri;
    size_t slot_d, part_d;
    if (!paired_plan(c, &nb, &S, &slot_d, &part_d, &ri))
        return cudaErrorInvalidConfiguration;
    {
        cudaError_t e = pair_reserve(c, argmax, nullptr);
        if (e != cudaSuccess)
            return e;
    }
    // mass-sorted layout: tracked passes of an unsharded world (a rank of a sharded world owns a contiguous range of ORIGINAL
    // indices -- essa_upload_shard / essa_download_shard -- which a global sort by mass would scatter over the ranks)
    bool const sorted = argmax && c->world == 1 && c->plan.batches == 1 && pair_sorted_allowed();
    c->pair_sorted_pass = sorted;
    if (argmax) {
        size_t const want = (size_t)nb * kPThreads * ri;
        if (want > c->pair_prior_n) {
            cudaFree(c->pair_prior);
            c->pair_prior = nullptr;
            c->pair_prior_n = 0;
            c->pair_prior_valid = false;
            cudaError_t e = cudaMalloc(&c->pair_prior, want * sizeof(unsigned));
            if (e != cudaSuccess)
                return e;
            c->pair_prior_n = want;
        }
        if (sorted) {
            cudaError_t e = grow((void**)&c->pair_perm, &c->pair_perm_n, want * sizeof(int));
            if (e == cudaSuccess)
                e = grow((void**)&c->pair_pos_s, &c->pair_pos_s_n, want * sizeof(double4));
            if (e != cudaSuccess)
                return e;
        }
        if (!c->pair_prior_valid || c->pair_prior_npad != (int)want) {
            if (sorted) {
                // gravity factors or the active mask may have changed: sort again.  pos[cur].w = active ? gf : 0 is the key
                // (stable radix sort: equal factors stay in index order); the n bodies go behind the npad - n padding slots
                size_t const n = (size_t)c->n, kb = (n * sizeof(unsigned long long) + 255) / 256 * 256, vb = (n * sizeof(int) + 255) / 256 * 256;
                size_t cub_bytes = 0;
                cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long const*)nullptr, (unsigned long long*)nullptr,
                    (int const*)nullptr, (int*)nullptr, (int)n, 0, 64, st);
                if (e == cudaSuccess)
                    e = grow(&c->pair_sort_tmp, &c->pair_sort_tmp_bytes, 2 * kb + vb + cub_bytes);
                if (e != cudaSuccess)
                    return e;
                unsigned char* base = static_cast<unsigned char*>(c->pair_sort_tmp);
                unsigned long long* k_in = reinterpret_cast<unsigned long long*>(base);
                unsigned long long* k_out = reinterpret_cast<unsigned long long*>(base + kb);
                int* v_in = reinterpret_cast<int*>(base + 2 * kb);
                int const front = (int)(want - n);
                k_pair_sort_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->pos[c->cur], (int)n, k_in, v_in);
                if (front > 0)
                    k_pair_perm_pad<<<(unsigned)((front + 255) / 256), 256, 0, st>>>(c->pair_perm, front);
                e = cub::DeviceRadixSort::SortPairs(base + 2 * kb + vb, cub_bytes, k_in, k_out, v_in, c->pair_perm + front, (int)n, 0, 64, st);
                if (e != cudaSuccess)
                    return e;
                k_pair_prior_reset_sorted<<<(unsigned)((want + 255) / 256), 256, 0, st>>>(c->pair_prior, c->pair_perm, c->active, (int)want);
                c->launches += 4;
            } else {
                k_pair_prior_reset<<<(unsigned)((want + 255) / 256), 256, 0, st>>>(c->pair_prior, c->active, c->n, (int)want);
                c->launches++;
            }
            c->pair_prior_valid = true;
            c->pair_prior_npad = (int)want;
        }
        if (sorted) {
            k_pair_gather_sorted<<<(unsigned)((want + 255) / 256), 256, 0, st>>>(c->pos[c->cur], c->pair_perm, c->pair_pos_s, (int)want);
            c->launches++;
        }
    }
    PairArgs A;
    A.prior = c->pair_prior;
    A.pos = sorted ? c->pair_pos_s : c->pos[c->cur];
    A.pos_orig = c->pos[c->cur];
    A.perm = sorted ? c->pair_perm : nullptr;
    A.nb = nb;
    A.dmax = nb / 2;
    A.S = S;
    pair_geometry(c, nb, &A.ib0, &A.ib1);
    A.npad = nb * kPThreads * ri;
    A.part_i = c->pair_part;
    A.slots = c->pair_slots;
    A.part_k = c->pair_kpart;
    A.slots_k = c->pair_kslots;
    A.eps2 = eps2;
    A.n_part_i = c->pair_part_n;
    A.n_slots = c->pair_slots_n;
    A.n_part_k = c->pair_kpart_bytes / sizeof(unsigned long long);
    A.n_slots_k = c->pair_kslots_bytes / sizeof(unsigned long long);
    A.n_world = c->n;
    // J-side slots live with the owner of the J block: own memory on one GPU, peer-mapped memory when sharded (capi.cu set the
    // table up); without peer access the writer-major fallback + NCCL reduce-scatter
    A.owner_major = c->world == 1 || c->pair_peer_ok;
    A.bpr = nb / c->world;
    for (int r = 0; r < kPairMaxPeers; r++) {
        A.slot_peer[r] = c->world == 1 ? c->pair_slots : (r < c->world ? c->pair_slot_peer[r] : nullptr);
        A.slotk_peer[r] = c->world == 1 ? c->pair_kslots : (r < c->world ? c->pair_kslot_peer[r] : nullptr);
    }
    // remote J tiles straight from their owners (the caller has put the barrier in front of this launch): untracked passes only
    // -- the tracked one also stages the owners' filter thresholds, which arrive with their own broadcast
    A.peer_pos = c->world > 1 && c->pair_peer_ok && c->pair_pospeer_ok && c->pair_peer_pos_pass && !argmax ? 1 : 0;
    for (int r = 0; r < kPairMaxPeers; r++)
        A.pos_peer[r] = A.peer_pos && r < c->world ? c->pair_pos_peer[c->cur][r] : nullptr;
    double* acc = nullptr;
    unsigned long long* kacc = nullptr;
    bool const uniform = pair_uniform(c) && !argmax;
    int const batches = c->plan.batches, db = c->plan.db;
    if (batches > 1 && (argmax || c->world > 1))
        return cudaErrorInvalidConfiguration; // candidate keys and peer slots are not batched (the routing never asks for it)
    A.dslots = batches > 1 ? db : A.dmax;
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < batches && e == cudaSuccess; b++) {
        int const last = (b + 1) * db < A.dmax ? (b + 1) * db : A.dmax;
        A.d0 = b == 0 ? 0 : b * db + 1; // the diagonal tile (d = 0, no slot) rides with the first batch
        A.dn = last - A.d0 + 1;
        A.dslot0 = b * db + 1;
        if (sorted)
            e = ri == 8 ? launch_pair_kernels<8, true, false, true>(c, A, st, &acc, &kacc) : launch_pair_kernels<4, true, false, true>(c, A, st, &acc, &kacc);
        else if (ri == 8)
            e = argmax ? launch_pair_kernels<8, true, false>(c, A, st, &acc, &kacc)
                       : (uniform ? launch_pair_kernels<8, false, true>(c, A, st, &acc, &kacc) : launch_pair_kernels<8, false, false>(c, A, st, &acc, &kacc));
        else
            e = argmax ? launch_pair_kernels<4, true, false>(c, A, st, &acc, &kacc)
                       : (uniform ? launch_pair_kernels<4, false, true>(c, A, st, &acc, &kacc) : launch_pair_kernels<4, false, false>(c, A, st, &acc, &kacc));
        if (e == cudaSuccess && batches > 1) {
            double* const jacc = c->pair_part + (size_t)A.S * 3 * A.npad;
            if (ri == 8)
                k_pair_fold<kPThreads * 8><<<(A.npad + 255) / 256, 256, 0, st>>>(nb, A.npad, A.S, A.d0, last, A.dslots, A.dslot0, b == 0, c->pair_part,
                    c->pair_slots, jacc);
            else
                k_pair_fold<kPThreads * 4><<<(A.npad + 255) / 256, 256, 0, st>>>(nb, A.npad, A.S, A.d0, last, A.dslots, A.dslot0, b == 0, c->pair_part,
                    c->pair_slots, jacc);
            c->launches++;
            e = cudaGetLastError();
        }
    }
    if (acc_out)
        *acc_out = acc;
    if (kacc_out)
        *kacc_out = kacc;
    if (npad_out)
        *npad_out = A.npad;
    return e;
}

// Phase 2: ordered sums + kick / drift epilogue for this rank's targets.
cudaError_t launch_force_paired_finish(essa_ctx* c, KickArgs const& k, cudaStream_t st, double const* ajred, bool argmax_noop, bool save_old,
    bool argmax, unsigned long long const* akred, double eps2) {
    int nb, S, ri;
    size_t slot_d, part_d;
    if (!paired_plan(c, &nb, &S, &slot_d, &part_d, &ri))
        return cudaErrorInvalidConfiguration;
    PairFinishArgs F;
    F.w = world_ptrs(c);
    F.k = k;
    F.n = c->n;
    F.i0 = (int)((long long)c->n * c->rank / c->world);
    F.i1 = (int)((long long)c->n * (c->rank + 1) / c->world);
    F.nb = nb;
    F.dmax = nb / 2;
    F.S = S;
    pair_geometry(c, nb, &F.ib0, &F.ib1);
    F.npad = nb * kPThreads * ri;
    F.part_i = c->pair_part;
    F.argmax_noop = argmax_noop ? 1 : 0;
    F.argmax = argmax ? 1 : 0;
    F.part_k = c->pair_kpart;
    F.slots_k = c->pair_kslots;
    F.prior = c->pair_prior;
    F.akred = akred;
    F.world = c->world;
    F.save_old = save_old ? 1 : 0;
    F.eps2 = eps2;
    F.scale = pair_uniform(c) && !argmax ? c->gf_uniform_value : 1.0;
    F.slots = c->pair_slots;
    F.ajred = ajred;
    F.perm = nullptr;
    F.prior_slack = kPriorSlack;
    if (char const* e = getenv("ESSA_PAIR_PRIOR_SLACK_LOG2")) // diagnostics: 18 = a factor 2 in the metric
        F.prior_slack = 1u << atoi(e);
    if (argmax && c->pair_sorted_pass) { // one thread per slot of the sorted layout (padding included: it returns at once)
        F.perm = c->pair_perm;
        F.i0 = 0;
        F.i1 = F.npad;
    }
    if (c->plan.batches > 1) { // the batches left their running sum behind the partials: nothing else to add
        F.ajred = c->pair_part + (size_t)S * 3 * F.npad;
        F.S = 0;
    }
    int cnt = F.i1 - F.i0;
    if (cnt <= 0)
        return cudaSuccess;
    if (ri == 8)
        k_pair_finish<kPThreads * 8><<<(cnt + 255) / 256, 256, 0, st>>>(F);
    else
        k_pair_finish<kPThreads * 4><<<(cnt + 255) / 256, 256, 0, st>>>(F);
    c->launches++;
    return cudaGetLastError();
}

} // namespace essa
