// K6 building blocks -- Object::update (src/Object.cpp:131-153): History::move_forward
// (src/History.cpp:11-31), Object::nonphysical_update / recalculate_trails_with_offset
// (src/Object.cpp:103-124,155-167) and Trail (src/Trail.cpp:49-88,105-110), as device functions
// over the rings that essa_record_configure lays out.  Used inline by the resident kernel and by
// the per-step record kernel of the tiled path.
//
// All arithmetic here follows the reference's order exactly (IEEE single / double, no
// contraction); the only function that is not bit-reproducible against glibc is atan2, which
// feeds a threshold test and the stored m_last_xy_angle.
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
    double thr; // 2 pi / capacity (src/Trail.cpp:58), one IEEE division per kernel instead of per step
    double ap, pe, apv, pev;
};

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
    b.last_angle = R.last_angle[k];
    b.ap = R.ap[k];
    b.pe = R.pe[k];
    b.apv = R.ap_vel[k];
    b.pev = R.pe_vel[k];
}

__device__ __forceinline__ void rec_store(RecordPtrs const& R, int k, BodyRec const& b) {
    R.hist_start[k] = b.hist_start;
    R.hist_len[k] = b.hist_len;
    R.append_off[k] = b.append_off;
    R.length[k] = b.length;
    R.toff[3 * k] = b.offx;
    R.toff[3 * k + 1] = b.offy;
    R.toff[3 * k + 2] = b.offz;
    R.last_angle[k] = b.last_angle;
    R.ap[k] = b.ap;
    R.pe[k] = b.pe;
    R.ap_vel[k] = b.apv;
    R.pe_vel[k] = b.pev;
}

// pos.cast<float>() / AU  (src/Trail.cpp:53,64): narrow first, divide in double, narrow again.
// The IEEE division is a long software sequence; the quotient is only needed to binary32, so it
// is formed as f * RN(1/AU) (within 3 double ulps of the correctly rounded quotient) and the slow
// exact division runs only when that could change the binary32 rounding, i.e. when the double's 29
// discarded mantissa bits sit within 8 ulps of the rounding boundary.  Bit-identical to the division.
__device__ __forceinline__ float to_vertex1(double p) {
    double f = (double)__double2float_rn(p);
    double qf = __dmul_rn(f, 1.0 / ESSA_AU);
    unsigned lo = (unsigned)__double2loint(qf) & 0x1fffffffu;
    unsigned ex = ((unsigned)__double2hiint(qf) >> 20) & 0x7ffu;
    bool risky = f != 0.0 && ((lo - 0x0ffffff8u) <= 16u || ex < 1023u - 120u || ex > 1023u + 120u);
    if (risky)
        qf = __ddiv_rn(f, ESSA_AU);
    return __double2float_rn(qf);
}

__device__ __forceinline__ void vert_write(RecordPtrs const& R, BodyRec const& b, int slot, float x, float y, float z) {
    float* v = R.verts + (b.voff + slot) * 3;
    v[0] = x;
    v[1] = y;
    v[2] = z;
}

// History::move_forward on the m_side == 0 branch (the only one World::update reaches).
__device__ __forceinline__ void history_push(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx,
    double vy, double vz) {
    int slot;
    if (b.hist_len < ESSA_HISTORY_SEGMENTS) {
        slot = b.hist_start + b.hist_len;
        if (slot >= ESSA_HISTORY_SEGMENTS)
            slot -= ESSA_HISTORY_SEGMENTS;
        b.hist_len++;
    } else {
        slot = b.hist_start;
        b.hist_start = b.hist_start + 1 == ESSA_HISTORY_SEGMENTS ? 0 : b.hist_start + 1;
    }
    // 48-byte entries in a 256-byte aligned ring: three 16-byte stores
    double2* e = reinterpret_cast<double2*>(R.hist + ((size_t)k * ESSA_HISTORY_SEGMENTS + slot) * 6);
    e[0] = make_double2(px, py);
    e[1] = make_double2(pz, vx);
    e[2] = make_double2(vy, vz);
}

// Trail::recalculate_with_offset (src/Trail.cpp:82-88)
__device__ __forceinline__ void trail_recalculate(RecordPtrs const& R, BodyRec& b, double ox, double oy, double oz) {
    if (b.offx == ox && b.offy == oy && b.offz == oz)
        return;
    float ax = to_vertex1(b.offx), ay = to_vertex1(b.offy), az = to_vertex1(b.offz);
    float bx = to_vertex1(ox), by = to_vertex1(oy), bz = to_vertex1(oz);
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

// Trail::push_back (src/Trail.cpp:49-75) incl. change_current (:105-110)
__device__ __forceinline__ void trail_push(RecordPtrs const& R, BodyRec& b, double px, double py, double pz) {
    float fx = to_vertex1(px), fy = to_vertex1(py), fz = to_vertex1(pz);
    if (b.length == 1) {
        vert_write(R, b, b.append_off, fx, fy, fz);
        b.append_off++;
        b.length++;
    }
    double ang = 0;
    bool have_ang = false;
    if (b.length >= 3 && b.append_off != 0) {
        // |atan2(y, x) - last| <= 2 pi / capacity decides between moving the newest vertex and appending.
        // A binary32 estimate (error < 2e-6 rad incl. the narrowing of x, y) settles every case that is
        // not within 8e-6 rad of the threshold; only those, and appends (which must store the angle),
        // pay for the binary64 atan2.
        double const thr = b.thr;
        double est = fabs((double)atan2f((float)py, (float)px) - b.last_angle);
        bool decimate;
        if (est < thr - 8e-6 && fabs(px) < 3e38 && fabs(py) < 3e38)
            decimate = true;
        else {
            ang = atan2(py, px);
            have_ang = true;
            decimate = fabs(__dsub_rn(ang, b.last_angle)) <= thr;
        }
        if (decimate) {
            if (b.append_off == 1)
                vert_write(R, b, b.length - 1, fx, fy, fz);
            vert_write(R, b, b.append_off - 1, fx, fy, fz);
            return;
        }
    }
    vert_write(R, b, b.append_off, fx, fy, fz);
    b.append_off++;
    if (b.length < b.cap)
        b.length++;
    if (b.append_off == b.cap) {
        vert_write(R, b, 0, fx, fy, fz);
        b.append_off = 1;
    }
    b.last_angle = have_ang ? ang : atan2(py, px);
}

__device__ __forceinline__ double length3(double x, double y, double z) {
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// Object::update(speed > 0) for one non-deleted body.  pos_of(j, x, y, z) yields the final
// position of body j (the drift of every body is finished before any update() runs).
template<class PosOf>
__device__ __forceinline__ void record_body(RecordPtrs const& R, int k, BodyRec& b, double px, double py, double pz, double vx, double vy,
    double vz, int most, int old_most, bool offset_trails, bool forward_sim, PosOf pos_of) {
    if (!forward_sim)
        history_push(R, k, b, px, py, pz, vx, vy, vz);
    double mx = 0, my = 0, mz = 0;
    if (most >= 0)
        pos_of(most, mx, my, mz);
    if (offset_trails && most >= 0 && !forward_sim) {
        if (most != old_most)
            trail_recalculate(R, b, mx, my, mz);
        else {
            b.offx = mx;
            b.offy = my;
            b.offz = mz;
        }
        trail_push(R, b, __dsub_rn(px, mx), __dsub_rn(py, my), __dsub_rn(pz, mz));
    } else {
        trail_recalculate(R, b, 0.0, 0.0, 0.0);
        trail_push(R, b, px, py, pz);
    }
    if (most < 0)
        return;
    // ap / pe (src/Object.cpp:111-123).  The exact distance (IEEE sqrt) is needed only when it sets a
    // new extreme; an estimate good to 1e-14 rules that out on most steps.
    double dx = __dsub_rn(px, mx), dy = __dsub_rn(py, my), dz = __dsub_rn(pz, mz);
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double y0 = rsqrt_seed(d2);
    double t = d2 * y0;
    double e = fma(-t, y0, 1.0);
    double est = fma(t * e, fma(0.375, e, 0.5), t); // ~ sqrt(d2)
    if (est * (1.0 + 1e-12) < b.ap && est * (1.0 - 1e-12) > b.pe)
        return;
    double d = __dsqrt_rn(d2);
    if (b.ap < d) {
        b.ap = d;
        b.apv = length3(vx, vy, vz);
    }
    if (b.pe > d) {
        b.pe = d;
        b.pev = length3(vx, vy, vz);
    }
}

} // namespace essa
