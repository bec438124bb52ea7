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
    double ap, pe, apv, pev;
};

__device__ __forceinline__ void rec_load(RecordPtrs const& R, int k, BodyRec& b) {
    b.hist_start = R.hist_start[k];
    b.hist_len = R.hist_len[k];
    b.append_off = R.append_off[k];
    b.length = R.length[k];
    b.cap = R.cap[k];
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
__device__ __forceinline__ float to_vertex1(double p) {
    return __double2float_rn(__ddiv_rn((double)__double2float_rn(p), ESSA_AU));
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
    double* e = R.hist + ((size_t)k * ESSA_HISTORY_SEGMENTS + slot) * 6;
    e[0] = px;
    e[1] = py;
    e[2] = pz;
    e[3] = vx;
    e[4] = vy;
    e[5] = vz;
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
        ang = atan2(py, px);
        have_ang = true;
        if (fabs(__dsub_rn(ang, b.last_angle)) <= __ddiv_rn(2 * 3.14159265358979323846, (double)b.cap)) {
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
    double d = length3(__dsub_rn(px, mx), __dsub_rn(py, my), __dsub_rn(pz, mz));
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
