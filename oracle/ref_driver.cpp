// TEST INFRASTRUCTURE.  C ABI over the REFERENCE's own History and Trail classes (compiled from /root/reference/src through
// oracle/_ref/src, see the Makefile) so that tests can drive them next to the oracle's restatement with the same inputs.
// Access control is lifted for this translation unit only (the reference offers no accessors for the ring contents): the
// headers the two classes include come first, so that the two macros touch nothing but `class History {` / `class Trail {`.
#include <Essa/Engine/3D/Shaders/Basic.hpp>
#include <EssaUtil/Vector.hpp>
#include <list>
#include <optional>
#include <vector>

#include "SimulationView.hpp"

#define private public
#define class struct
#include "History.hpp"
#include "Trail.hpp"
#undef class
#undef private

#include <cstring>

extern "C" {

void* ref_history_new(unsigned segments, double const* first) {
    return new History(segments, { { first[0], first[1], first[2] }, { first[3], first[4], first[5] } });
}
void ref_history_free(void* h) { delete static_cast<History*>(h); }
static int give(std::optional<History::Entry> const& e, double* out) {
    if (!e)
        return 0;
    out[0] = e->pos.x(); out[1] = e->pos.y(); out[2] = e->pos.z();
    out[3] = e->vel.x(); out[4] = e->vel.y(); out[5] = e->vel.z();
    return 1;
}
int ref_history_move_forward(void* h, double const* e, double* out) {
    return give(static_cast<History*>(h)->move_forward({ { e[0], e[1], e[2] }, { e[3], e[4], e[5] } }), out);
}
int ref_history_move_backward(void* h, double const* e, double* out) {
    return give(static_cast<History*>(h)->move_backward({ { e[0], e[1], e[2] }, { e[3], e[4], e[5] } }), out);
}
void ref_history_reset(void* h) { static_cast<History*>(h)->reset(); }
int ref_history_read(void* h, double* out, int max_entries, int* side) {
    auto* H = static_cast<History*>(h);
    int k = 0;
    for (auto const& e : H->m_entry_list) {
        if (k < max_entries) {
            double v[6] = { e.pos.x(), e.pos.y(), e.pos.z(), e.vel.x(), e.vel.y(), e.vel.z() };
            std::memcpy(out + 6 * k, v, sizeof(v));
        }
        k++;
    }
    if (side)
        *side = H->m_side ? 1 : 0;
    return k;
}

void* ref_trail_new(size_t max_trail_size) { return new Trail(max_trail_size, Util::Color {}); }
void ref_trail_free(void* t) { delete static_cast<Trail*>(t); }
void ref_trail_push_back(void* t, double x, double y, double z) { static_cast<Trail*>(t)->push_back({ x, y, z }); }
void ref_trail_change_current(void* t, double x, double y, double z) { static_cast<Trail*>(t)->change_current({ x, y, z }); }
void ref_trail_set_offset(void* t, double x, double y, double z) { static_cast<Trail*>(t)->set_offset({ x, y, z }); }
void ref_trail_recalculate_with_offset(void* t, double x, double y, double z) { static_cast<Trail*>(t)->recalculate_with_offset({ x, y, z }); }
void ref_trail_reset(void* t) { static_cast<Trail*>(t)->reset(); }
// verts: 3 floats per vertex (capacity of them); meta = {capacity, append_offset, length}; real = {offset x y z, last_xy_angle}
void ref_trail_read(void* t, float* verts, int* meta, double* real) {
    auto* T = static_cast<Trail*>(t);
    if (verts)
        for (size_t i = 0; i < T->m_vertexes.size(); i++) {
            auto const& p = T->m_vertexes[i].value<0>();
            verts[3 * i] = p.x(); verts[3 * i + 1] = p.y(); verts[3 * i + 2] = p.z();
        }
    meta[0] = (int)T->m_vertexes.size(); meta[1] = T->m_append_offset; meta[2] = T->m_length;
    real[0] = T->m_offset.x(); real[1] = T->m_offset.y(); real[2] = T->m_offset.z(); real[3] = T->last_xy_angle;
}

} // extern "C"
