// K6 building blocks -- Object::update (src/Object.cpp:131-153): History::move_forward
// (src/History.cpp:11-31), Object::nonphysical_update / recalculate_trails_with_offset
// (src/Object.cpp:103-124,155-167) and Trail (src/Trail.cpp:49-88,105-110), as device functions
// over the rings that essa_record_configure lays out.  Used by the resident kernels (inline or on
// a recorder warp) and by the per-step record kernel of the tiled path.
//
// Results are bit-identical to the reference's order of operations (IEEE single / double, no
// contraction); the only value that is not bit-reproducible against glibc is atan2, which feeds a
// threshold test and the stored m_last_xy_angle.  What is NOT done the reference's way is the
// amount of work per step -- three things are deferred while a kernel runs, because only their
// last value is observable:
//   * Trail::change_current overwrites the newest vertex on every decimated step; the vertex
//     (two float conversions and a division per component) is materialised only when the next
//     append, an offset change or the end of the kernel needs it;
//   * m_ap / m_pe move on every step of a monotonic stretch of an orbit; the exact distance
//     (IEEE sqrt) and |v| (another one) are formed when the stretch ends.  Decisions are taken on
//     squared distances with a 1e-15 guard band inside which the exact path runs;
//   * the decimation test |atan2(y, x) - last| <= 2 pi / capacity is settled without an arctangent
//     against the unit vector of the last appended direction: "decimate" when the new direction is
//     within (threshold - 1e-9 rad) of it and the two cannot lie on opposite sides of the +-pi cut
//     (|cross| < tan(threshold - 1e-9) dot), "append" when it is more than (threshold + 1e-9 rad)
//     away; m_last_xy_angle itself (an atan2) is formed when the narrow band in between, a cut
//     crossing or the end of the kernel needs its value.
#pragma once

#include "common.cuh"
#include "ctx.hpp"

namespace essa {

// Per-body recording state held in registers while a kernel runs.
struct BodyRec {
    int hist_start, hist_len;
    int append_off, length, cap;
    long long voff;
    double offx, offy, offz;
    double last_angle;
    double thr; // 2 pi / capacity (src/Trail.cpp:58): one IEEE division per kernel, not per step
    double tthr_lo, tthr_hi; // tan(thr -+ 1e-9); -1 / +inf when the shortcuts are not usable
    double dir_c, dir_s;     // a vector along last_angle; both 0 when last_angle is not an atan2 value
    int ang_pend;            // last_angle = atan2(dir_s, dir_c) not formed yet
    double ap, pe, apv, pev;
    // deferred work
    int vpend, ap_pend, pe_pend;
    double vpx, vpy, vpz;              // position of the pending change_current
    double ap_d2, ap_vx, ap_vy, ap_vz; // squared distance / velocity of the pending apoapsis
    double pe_d2, pe_vx, pe_vy, pe_vz;
};

__device__ __forceinline__ void trail_set_angle(BodyRec& b, double ang) {
    b.last_angle = ang;
    b.ang_pend = 0;
    if (fabs(ang) <= 3.14159265358979323846)
        sincos(ang, &b.dir_s, &b.dir_c);
    else
        b.dir_s = b.dir_c = 0.0;
}

// m_last_xy_angle = atan2(y, x), postponed.  The shortcuts compare directions only (both sides of
// |cross| < tan * dot scale with the reference vector's length), so the position itself serves as the
// reference vector.  Overflow / zero give inf / NaN / 0, which fail both shortcuts.
__device__ __forceinline__ void trail_set_direction(BodyRec& b, double px, double py) {
    b.ang_pend = 1;
    b.dir_c = px;
    b.dir_s = py;
}

__device__ __forceinline__ double trail_last_angle(BodyRec& b) {
    if (b.ang_pend) {
        b.last_angle = atan2(b.dir_s, b.dir_c);
        b.ang_pend = 0;
    }
    return b.last_angle;
}

__device__ __forceinline__ void rec_load(RecordPtrs const& R, int k, BodyRec& b) {
    b.hist_start = R.hist_start[k];
    b.hist_len = R.hist_len[k];
    b.append_off = R.append_off[k];
    b.length = R.length[k];
    b.cap = R.cap[k];
    b.thr = __ddiv_rn(2 * 3.14159265358979323846, (double)b.cap);
    b.voff = R.vert_off[k];
    b.offx = R.toff[3 * k];
    b.offy = R.toff[3 * k + 1];
    b.offz = R.toff[3 * k + 2];
    trail_set_angle(b, R.last_angle[k]);
    bool const usable = b.thr > 2e-9 && b.thr < 1.5;
    b.tthr_lo = usable ? tan(b.thr - 1e-9) : -1.0;
    b.tthr_hi = usable ? tan(b.thr + 1e-9) : __longlong_as_double(0x7ff0000000000000LL);
    b.ap = R.ap[k];
    b.pe = R.pe[k];
    b.apv = R.ap_vel[k];
    b.pev = R.pe_vel[k];
    b.vpend = b.ap_pend = b.pe_pend = 0;
    b.vpx = b.vpy = b.vpz = 0;
    b.ap_d2 = b.ap_vx = b.ap_vy = b.ap_vz = 0;
    b.pe_d2 = b.pe_vx = b.pe_vy = b.pe_vz = 0;
}

// pos.cast<float>() / AU  (src/Trail.cpp:53,64): narrow first, divide in double, narrow again.
// The IEEE division is a long software sequence; the quotient is only needed to binary32, so it
// is formed as f * RN(1/AU) (within 3 double ulps of the correctly rounded quotient) and the slow
// exact division runs only when that could change the binary32 rounding, i.e. when the double's 29
// discarded mantissa bits sit within 8 ulps of the rounding boundary.  Bit-identical to the division.
// (kept out of line so that the compiler cannot if-convert it into the common path)
static __device__ __noinline__ double divide_by_au_exact(double f) { return __ddiv_rn(f, ESSA_AU); }

__device__ __forceinline__ float to_vertex1(double p) {
    double f = (double)__double2float_rn(p);
    double qf = __dmul_rn(f, 1.0 / ESSA_AU);
    unsigned lo = (unsigned)__double2loint(qf) & 0x1fffffffu;
    unsigned ex = ((unsigned)__double2hiint(qf) >> 20) & 0x7ffu;
    bool risky = f != 0.0 && ((lo - 0x0ffffff8u) <= 16u || ex < 1023u - 120u || ex > 1023u + 120u);
    if (risky)
        qf = divide_by_au_exact(f);
    return __double2float_rn(qf);
}

__device__ __forceinline__ void vert_write(RecordPtrs const& R, BodyRec const& b, int slot, float x, float y, float z) {
    ESSA_CHECK(slot >= 0 && slot < b.cap && b.cap >= 3);
    float* v = R.verts + (b.voff + slot) * 3;
    v[0] = x;
    v[1] = y;
    v[2] = z;
}

__device__ __forceinline__ double length3(double x, double y, double z) {
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// Trail::change_current (src/Trail.cpp:105-110) for the position remembered by the last decimated push.
__device__ __forceinline__ void trail_flush_vertex(RecordPtrs const& R, BodyRec& b) {
    if (!b.vpend)
        return;
    float fx = to_vertex1(b.vpx), fy = to_vertex1(b.vpy), fz = to_vertex1(b.vpz);
    if (b.append_off == 1)
        vert_write(R, b, b.length - 1, fx, fy, fz);
    vert_write(R, b, b.append_off - 1, fx, fy, fz);
    b.vpend = 0;
}

__device__ __forceinline__ void ap_flush(BodyRec& b) {
    if (b.ap_pend) {
        b.ap = __dsqrt_rn(b.ap_d2);
        b.apv = length3(b.ap_vx, b.ap_vy, b.ap_vz);
        b.ap_pend = 0;
    }
}

__device__ __forceinline__ void pe_flush(BodyRec& b) {
    if (b.pe_pend) {
        b.pe = __dsqrt_rn(b.pe_d2);
        b.pev = length3(b.pe_vx, b.pe_vy, b.pe_vz);
        b.pe_pend = 0;
    }
}

// Settles everything deferred and writes the per-body state back (trail / history part).
__device__ __forceinline__ void rec_store_trail(RecordPtrs const& R, int k, BodyRec& b, bool with_history = true) {
    trail_flush_vertex(R, b);
    if (with_history) {
        R.hist_start[k] = b.hist_start;
        R.hist_len[k] = b.hist_len;
    }
    R.append_off[k] = b.append_off;
    R.length[k] = b.length;
    R.toff[3 * k] = b.offx;
    R.toff[3 * k + 1] = b.offy;
    R.toff[3 * k + 2] = b.offz;
    R.last_angle[k] = trail_last_angle(b);
}

// ... and the ap / pe part.
__device__ __forceinline__ void rec_store_appe(RecordPtrs const& R, int k, BodyRec& b) {
    ap_flush(b);
    pe_flush(b);
    R.ap[k] = b.ap;
    R.pe[k] = b.pe;
    R.ap_vel[k] = b.apv;
    R.pe_vel[k] = b.pev;
}

__device__ __forceinline__ void rec_store(RecordPtrs const& R, int k, BodyRec& b) {
    rec_store_trail(R, k, b);
    rec_store_appe(R, k, b);
}

// History::move_forward on the m_side == 0 branch (the only one World::update reaches).
__device__ __forceinline__ void history_push(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx,
    double vy, double vz) {
    // growing: append behind the newest entry; full: overwrite the oldest and advance the start
    bool const grow = b.hist_len < ESSA_HISTORY_SEGMENTS;
    int slot = grow ? b.hist_start + b.hist_len : b.hist_start;
    slot = slot >= ESSA_HISTORY_SEGMENTS ? slot - ESSA_HISTORY_SEGMENTS : slot;
    int const next = b.hist_start + 1 == ESSA_HISTORY_SEGMENTS ? 0 : b.hist_start + 1;
    b.hist_len = grow ? b.hist_len + 1 : b.hist_len;
    b.hist_start = grow ? b.hist_start : next;
    // 48-byte entries in a 256-byte aligned ring: three 16-byte stores
    ESSA_CHECK(k >= 0 && k < R.tracked && slot >= 0 && slot < ESSA_HISTORY_SEGMENTS && b.hist_len <= ESSA_HISTORY_SEGMENTS);
    double2* e = reinterpret_cast<double2*>(R.hist + ((size_t)k * ESSA_HISTORY_SEGMENTS + slot) * 6);
    e[0] = make_double2(px, py);
    e[1] = make_double2(pz, vx);
    e[2] = make_double2(vy, vz);
}

// Trail::recalculate_with_offset (src/Trail.cpp:82-88)
__device__ __forceinline__ void trail_recalculate(RecordPtrs const& R, BodyRec& b, double ox, double oy, double oz) {
    if (b.offx == ox && b.offy == oy && b.offz == oz)
        return;
    trail_flush_vertex(R, b);
    float ax = to_vertex1(b.offx), ay = to_vertex1(b.offy), az = to_vertex1(b.offz);
    float bx = to_vertex1(ox), by = to_vertex1(oy), bz = to_vertex1(oz);
    ESSA_CHECK(b.length >= 0 && b.length <= b.cap);
    float* v = R.verts + b.voff * 3;
    for (int s = 0; s < b.length; s++) {
        v[3 * s] = __fsub_rn(__fadd_rn(v[3 * s], ax), bx);
        v[3 * s + 1] = __fsub_rn(__fadd_rn(v[3 * s + 1], ay), by);
        v[3 * s + 2] = __fsub_rn(__fadd_rn(v[3 * s + 2], az), bz);
    }
    b.offx = ox;
    b.offy = oy;
    b.offz = oz;
}

// Trail::push_back (src/Trail.cpp:49-75) incl. change_current (:105-110), in two halves so that a warp
// can convert the vertices of an append cooperatively (resident_warp.cu) between them.
//
// First half: everything up to the decision.  Returns true when a vertex has to be appended.
__device__ __forceinline__ bool trail_decide(RecordPtrs const& R, BodyRec& b, double px, double py, double pz) {
    // The common step -- still within the decimation angle of the last appended vertex -- is branch free:
    // change_current(pos) is only remembered (written when something can observe it).
    // same_side: both directions in one open half plane that the +-pi cut does not cross
    bool const same_side = (px > 0.0 && b.dir_c > 0.0) || (py > 0.0 && b.dir_s > 0.0) || (py < 0.0 && b.dir_s < 0.0);
    double const dot = fma(px, b.dir_c, py * b.dir_s);
    double const crs = fabs(fma(px, b.dir_s, -(py * b.dir_c)));
    bool const tested = b.length >= 3 && b.append_off != 0;
    bool const quick = tested && same_side && crs < b.tthr_lo * dot;
    b.vpend = quick ? 1 : b.vpend;
    b.vpx = quick ? px : b.vpx;
    b.vpy = quick ? py : b.vpy;
    b.vpz = quick ? pz : b.vpz;
    if (quick)
        return false;

    if (b.length == 1) {
        vert_write(R, b, b.append_off, to_vertex1(px), to_vertex1(py), to_vertex1(pz));
        b.append_off++;
        b.length++;
    }
    if (b.length >= 3 && b.append_off != 0) {
        // farther than the threshold for sure (a cut crossing only makes |difference| larger)?  else the real thing
        bool const apart = b.tthr_lo > 0.0 && (dot <= 0.0 ? (crs > 0.0 || dot < 0.0) : crs > b.tthr_hi * dot);
        if (!apart && fabs(__dsub_rn(atan2(py, px), trail_last_angle(b))) <= b.thr) {
            b.vpend = 1;
            b.vpx = px;
            b.vpy = py;
            b.vpz = pz;
            return false;
        }
    }
    return true;
}

// Second half: the append.  v[0..2] = the pending change_current vertex (used only if there is one),
// v[3..5] = the new vertex, both already converted with to_vertex1.
__device__ __forceinline__ void trail_append(RecordPtrs const& R, BodyRec& b, float const* v, double px, double py) {
    if (b.vpend) {
        if (b.append_off == 1)
            vert_write(R, b, b.length - 1, v[0], v[1], v[2]);
        vert_write(R, b, b.append_off - 1, v[0], v[1], v[2]);
        b.vpend = 0;
    }
    ESSA_CHECK(b.append_off >= 0 && b.append_off < b.cap && (!b.vpend || b.append_off >= 1));
    vert_write(R, b, b.append_off, v[3], v[4], v[5]);
    b.append_off++;
    if (b.length < b.cap)
        b.length++;
    if (b.append_off == b.cap) {
        vert_write(R, b, 0, v[3], v[4], v[5]);
        b.append_off = 1;
    }
    trail_set_direction(b, px, py);
}

__device__ __forceinline__ void trail_push(RecordPtrs const& R, BodyRec& b, double px, double py, double pz) {
    if (!trail_decide(R, b, px, py, pz))
        return;
    float v[6] = { 0.f, 0.f, 0.f, to_vertex1(px), to_vertex1(py), to_vertex1(pz) };
    if (b.vpend) {
        v[0] = to_vertex1(b.vpx);
        v[1] = to_vertex1(b.vpy);
        v[2] = to_vertex1(b.vpz);
    }
    trail_append(R, b, v, px, py);
}

// m_ap / m_ap_vel (src/Object.cpp:115-118): `if (ap < d) { ap = d; ap_vel = |v| }`, d = RN(sqrt(d2)).
// RN o sqrt is monotone, so d2 <= (pending d2) means "no update"; a d2 more than 1e-15 (relative)
// above the reference square means d is strictly greater; in between, the exact path decides.
__device__ __forceinline__ void apoapsis_step(BodyRec& b, double d2, double vx, double vy, double vz) {
    double const ref2 = b.ap_pend ? b.ap_d2 : __dmul_rn(b.ap, b.ap);
    double const lo = b.ap_pend ? ref2 : ref2 * (1.0 - 1e-15);
    double const hi = ref2 * (1.0 + 1e-15);
    bool const up = d2 > hi;             // strictly farther: remember, settle later
    if (d2 > lo && !up) {                // inside the guard band (rare): the reference's own comparison
        ap_flush(b);
        double d = __dsqrt_rn(d2);
        if (b.ap < d) {
            b.ap = d;
            b.apv = length3(vx, vy, vz);
        }
        return;
    }
    b.ap_pend = up ? 1 : b.ap_pend;
    b.ap_d2 = up ? d2 : b.ap_d2;
    b.ap_vx = up ? vx : b.ap_vx;
    b.ap_vy = up ? vy : b.ap_vy;
    b.ap_vz = up ? vz : b.ap_vz;
}

// m_pe / m_pe_vel (src/Object.cpp:120-123): `if (pe > d) { pe = d; pe_vel = |v| }`
__device__ __forceinline__ void periapsis_step(BodyRec& b, double d2, double vx, double vy, double vz) {
    double const ref2 = b.pe_pend ? b.pe_d2 : __dmul_rn(b.pe, b.pe);
    double const hi = b.pe_pend ? ref2 : ref2 * (1.0 + 1e-15);
    double const lo = ref2 * (1.0 - 1e-15);
    bool const down = d2 < lo;
    if (d2 < hi && !down) {
        pe_flush(b);
        double d = __dsqrt_rn(d2);
        if (b.pe > d) {
            b.pe = d;
            b.pev = length3(vx, vy, vz);
        }
        return;
    }
    b.pe_pend = down ? 1 : b.pe_pend;
    b.pe_d2 = down ? d2 : b.pe_d2;
    b.pe_vx = down ? vx : b.pe_vx;
    b.pe_vy = down ? vy : b.pe_vy;
    b.pe_vz = down ? vz : b.pe_vz;
}

// Object::update(speed > 0) for one non-deleted body, in two independent halves so that they can
// run on different warps.  (mx, my, mz) is the final position of the most attracting object (the
// drift of every body is finished before any update() runs); ignored when most < 0.
//
// Half 1: History::move_forward + recalculate_trails_with_offset / Trail::push_back.
// (returns whether a vertex has to be appended for the trail-relative position (rx, ry, rz))
__device__ __forceinline__ bool record_trail_decide(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx,
    double vy, double vz, double mx, double my, double mz, int most, int old_most, bool offset_trails, bool forward_sim, bool with_history,
    double& rx, double& ry, double& rz) {
    if (with_history && !forward_sim)
        history_push(R, k, b, px, py, pz, vx, vy, vz);
    bool const rel = offset_trails && most >= 0 && !forward_sim;
    double const ox = rel ? mx : 0.0, oy = rel ? my : 0.0, oz = rel ? mz : 0.0;
    bool const follow = rel && most == old_most; // Trail::set_offset: the offset just follows the same body
    if (!follow && (b.offx != ox || b.offy != oy || b.offz != oz))
        trail_recalculate(R, b, ox, oy, oz); // the reference point changed: re-base every vertex
    b.offx = ox;
    b.offy = oy;
    b.offz = oz;
    // pos - most.pos, or pos itself (x - 0.0 == x exactly)
    rx = rel ? __dsub_rn(px, mx) : px;
    ry = rel ? __dsub_rn(py, my) : py;
    rz = rel ? __dsub_rn(pz, mz) : pz;
    return trail_decide(R, b, rx, ry, rz);
}

__device__ __forceinline__ void record_trail(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx, double vy,
    double vz, double mx, double my, double mz, int most, int old_most, bool offset_trails, bool forward_sim, bool with_history = true) {
    double rx, ry, rz;
    if (!record_trail_decide(R, k, b, px, py, pz, vx, vy, vz, mx, my, mz, most, old_most, offset_trails, forward_sim, with_history, rx, ry,
            rz))
        return;
    float v[6] = { 0.f, 0.f, 0.f, to_vertex1(rx), to_vertex1(ry), to_vertex1(rz) };
    if (b.vpend) {
        v[0] = to_vertex1(b.vpx);
        v[1] = to_vertex1(b.vpy);
        v[2] = to_vertex1(b.vpz);
    }
    trail_append(R, b, v, rx, ry);
}

// Half 2: ap / pe against the most attracting object (src/Object.cpp:111-123).
__device__ __forceinline__ void record_appe(BodyRec& b, double px, double py, double pz, double vx, double vy, double vz, double mx,
    double my, double mz, int most) {
    if (most < 0)
        return;
    double dx = __dsub_rn(px, mx), dy = __dsub_rn(py, my), dz = __dsub_rn(pz, mz);
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    apoapsis_step(b, d2, vx, vy, vz);
    periapsis_step(b, d2, vx, vy, vz);
}

template<class PosOf>
__device__ __forceinline__ void record_body(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx, double vy,
    double vz, int most, int old_most, bool offset_trails, bool forward_sim, PosOf pos_of) {
    double mx = 0, my = 0, mz = 0;
    if (most >= 0)
        pos_of(most, mx, my, mz);
    record_trail(R, k, b, px, py, pz, vx, vy, vz, mx, my, mz, most, old_most, offset_trails, forward_sim);
    record_appe(b, px, py, pz, vx, vy, vz, mx, my, mz, most);
}

} // namespace essa
