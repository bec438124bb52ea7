// include/essa_world.h over the C++ facade: exceptions stop here.
#include <cstring>
#include <string>

#include "../../include/essa_world.h"
#include "facade.hpp"

struct essa_world {
    essa::World world;
    std::string err;
    std::string scratch;
    explicit essa_world(int device)
        : world(device) { }
};

static thread_local std::string g_world_error;

#define GUARD(...)                                                                                       \
    try {                                                                                                \
        __VA_ARGS__                                                                                      \
    } catch (std::exception const& e) {                                                                  \
        w->err = e.what();                                                                               \
        return ESSA_ERR_STATE;                                                                           \
    }

#define WANT(cond, msg)                                                                                  \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            if (w)                                                                                       \
                w->err = msg;                                                                            \
            return ESSA_ERR_ARG;                                                                         \
        }                                                                                                \
    } while (0)

using essa::Object;
using essa::Vec3;

extern "C" {

essa_world* essa_world_create(int device) { return new essa_world(device); }
void essa_world_destroy(essa_world* w) { delete w; }
const char* essa_world_last_error(const essa_world* w) { return w ? w->err.c_str() : g_world_error.c_str(); }

int essa_world_load(essa_world* w, const char* essa_file) {
    WANT(w && essa_file, "null argument");
    GUARD(w->world.reset(std::string(essa_file));)
    return ESSA_OK;
}

int essa_world_load_text(essa_world* w, const char* essa_text) {
    WANT(w && essa_text, "null argument");
    GUARD(w->world.reset(std::nullopt); std::string e = essa::load_essa_text(essa_text, w->world); if (!e.empty()) {
        w->err = "World loading failed: " + e;
        return ESSA_ERR_ARG;
    })
    return ESSA_OK;
}

int essa_world_add_object(essa_world* w, double mass, double radius, const double pos[3], const double vel[3], const uint8_t rgb[3],
    const char* name, unsigned period) {
    WANT(w && pos && vel, "null argument");
    essa::Color c;
    if (rgb)
        c = { rgb[0], rgb[1], rgb[2], 255 };
    GUARD(w->world.add_object(std::make_unique<Object>(mass, radius, Vec3 { pos[0], pos[1], pos[2] }, Vec3 { vel[0], vel[1], vel[2] }, c,
                                  name ? name : "", period));)
    return static_cast<int>(w->world.object_count()) - 1;
}

int essa_world_add_orbiting_ap_pe(essa_world* w, int around, double mass, float radius, float apoapsis, float periapsis, int direction,
    double theta, double alpha, const char* name) {
    WANT(w && around >= 0 && around < (int)w->world.object_count(), "bad `around` index");
    GUARD(w->world.add_object(w->world.object_at(around)->create_object_relative_to_ap_pe(mass, radius, apoapsis, periapsis, direction != 0,
                                  theta, alpha, essa::Color {}, name ? name : "", 0.0));)
    return static_cast<int>(w->world.object_count()) - 1;
}

int essa_world_add_orbiting_maj_ecc(essa_world* w, int around, double mass, float radius, float semi_major, double eccentrity, int direction,
    double theta, double alpha, const char* name) {
    WANT(w && around >= 0 && around < (int)w->world.object_count(), "bad `around` index");
    GUARD(w->world.add_object(w->world.object_at(around)->create_object_relative_to_maj_ecc(mass, radius, semi_major, eccentrity,
                                  direction != 0, theta, alpha, essa::Color {}, name ? name : "", 0.0));)
    return static_cast<int>(w->world.object_count()) - 1;
}

int essa_world_delete_object(essa_world* w, int index) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    GUARD(w->world.delete_object_by_ptr(w->world.object_at(index));)
    return ESSA_OK;
}

int essa_world_mark_deleted(essa_world* w, int index) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    w->world.object_at(index)->delete_object();
    return ESSA_OK;
}

int essa_world_count(const essa_world* w) { return w ? static_cast<int>(w->world.object_count()) : 0; }

int essa_world_find(essa_world* w, const char* name) {
    if (!w || !name)
        return -1;
    Object* o = w->world.get_object_by_name(name);
    return o ? w->world.index_of(o) : -1;
}

int64_t essa_world_date(const essa_world* w) { return w ? w->world.date() : 0; }
int essa_world_set_date(essa_world* w, int64_t date) {
    WANT(w, "null world");
    w->world.set_date(date);
    return ESSA_OK;
}
int essa_world_get_dt(const essa_world* w) { return w ? w->world.simulation_seconds_per_tick() : 0; }
int essa_world_set_dt(essa_world* w, int seconds) {
    WANT(w, "null world");
    w->world.set_simulation_seconds_per_tick(seconds);
    return ESSA_OK;
}
int essa_world_set_flags(essa_world* w, int offset_trails, int exact_order, int two_pass) {
    WANT(w, "null world");
    w->world.offset_trails = offset_trails != 0;
    w->world.exact_order = exact_order != 0;
    w->world.two_pass = two_pass != 0;
    return ESSA_OK;
}

int essa_world_update(essa_world* w, int steps) {
    WANT(w, "null world");
    WANT(steps != 0, "steps must not be 0 (assert in src/World.cpp:87)");
    GUARD(w->world.update(steps);)
    return ESSA_OK;
}

int essa_world_set_forces(essa_world* w) {
    WANT(w, "null world");
    GUARD(w->world.set_forces();)
    return ESSA_OK;
}

double essa_world_last_step_ms(const essa_world* w) { return w ? w->world.last_step_ms() : 0.0; }

int essa_world_debug_stash_last_object(essa_world* w) {
    WANT(w, "null world");
    GUARD(w->world.debug_stash_last_object();)
    return ESSA_OK;
}

int essa_world_object_history_size(const essa_world* w) { return w ? (int)w->world.object_history_size() : 0; }

int essa_world_set_persistent(essa_world* w, int enable, int idle_timeout_us) {
    WANT(w, "null world");
    GUARD(w->world.set_persistent(enable != 0, idle_timeout_us);)
    return ESSA_OK;
}

int essa_world_persist_stats(essa_world* w, int64_t out[6]) {
    WANT(w, "null world");
    WANT(out, "out is null");
    GUARD(w->world.persist_stats(out);)
    return ESSA_OK;
}

int essa_world_get_state(essa_world* w, double* pos, double* vel, double* acc, double* gf, int32_t* most, double* appe, int32_t* active,
    double* maxattr) {
    WANT(w, "null world");
    int k = 0;
    w->world.for_each_object([&](Object& o) {
        Vec3 p = o.pos(), v = o.vel(), a = o.acc();
        if (pos) { pos[3 * k] = p.x; pos[3 * k + 1] = p.y; pos[3 * k + 2] = p.z; }
        if (vel) { vel[3 * k] = v.x; vel[3 * k + 1] = v.y; vel[3 * k + 2] = v.z; }
        if (acc) { acc[3 * k] = a.x; acc[3 * k + 1] = a.y; acc[3 * k + 2] = a.z; }
        if (gf) gf[k] = o.gravity_factor();
        if (most) most[k] = o.most_attracting_object() ? w->world.index_of(o.most_attracting_object()) : -1;
        if (appe) { appe[4 * k] = o.apoapsis(); appe[4 * k + 1] = o.periapsis(); appe[4 * k + 2] = o.apoapsis_velocity(); appe[4 * k + 3] = o.periapsis_velocity(); }
        if (active) active[k] = o.deleted() ? 0 : 1;
        if (maxattr) maxattr[k] = o.max_attraction();
        k++;
    });
    return ESSA_OK;
}

int essa_world_set_object(essa_world* w, int index, const double pos[3], const double vel[3], double gf) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    Object* o = w->world.object_at(index);
    if (pos) o->set_pos({ pos[0], pos[1], pos[2] });
    if (vel) o->set_vel({ vel[0], vel[1], vel[2] });
    if (gf >= 0) o->set_gravity_factor(gf);
    return ESSA_OK;
}

const char* essa_world_object_name(essa_world* w, int index) {
    if (!w || index < 0 || index >= (int)w->world.object_count())
        return "";
    w->scratch = w->world.object_at(index)->name();
    return w->scratch.c_str();
}

int essa_world_set_object_name(essa_world* w, int index, const char* name) {
    WANT(w && name && index >= 0 && index < (int)w->world.object_count(), "bad index");
    w->world.object_at(index)->set_name(name);
    return ESSA_OK;
}

int essa_world_object_props(essa_world* w, int index, double out[7]) {
    WANT(w && out && index >= 0 && index < (int)w->world.object_count(), "bad index");
    Object* o = w->world.object_at(index);
    out[0] = o->radius();
    out[1] = o->orbit_period();
    out[2] = o->eccentrity();
    out[3] = o->trail_capacity();
    out[4] = o->color().r;
    out[5] = o->color().g;
    out[6] = o->color().b;
    return ESSA_OK;
}

int essa_world_set_object_props(essa_world* w, int index, double radius, const uint8_t rgb[3]) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    Object* o = w->world.object_at(index);
    if (radius >= 0) o->set_radius(radius);
    if (rgb) o->set_color({ rgb[0], rgb[1], rgb[2], 255 });
    return ESSA_OK;
}

int essa_world_trail(essa_world* w, int index, float* verts, essa_trail_meta* meta) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    GUARD(essa::Trail const& t = w->world.object_at(index)->trail(); if (meta) {
        meta->capacity = static_cast<int32_t>(t.size());
        meta->append_offset = t.append_offset;
        meta->length = t.length;
        meta->pad = 0;
        meta->offset[0] = t.offset.x;
        meta->offset[1] = t.offset.y;
        meta->offset[2] = t.offset.z;
        meta->last_xy_angle = t.last_xy_angle;
    } if (verts) std::memcpy(verts, t.vertexes.data(), t.vertexes.size() * sizeof(float));)
    return ESSA_OK;
}

int essa_world_history(essa_world* w, int index, double* out) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    int n = 0;
    GUARD(auto const& h = w->world.object_at(index)->history(); n = static_cast<int>(h.size()); if (out) {
        for (int k = 0; k < n; k++) {
            double v[6] = { h[k].pos.x, h[k].pos.y, h[k].pos.z, h[k].vel.x, h[k].vel.y, h[k].vel.z };
            std::memcpy(out + 6 * k, v, sizeof(v));
        }
    })
    return n;
}

int essa_world_reset_history(essa_world* w, int index) {
    WANT(w && index >= 0 && index < (int)w->world.object_count(), "bad index");
    GUARD(w->world.object_at(index)->reset_history();)
    return ESSA_OK;
}

essa_world* essa_world_clone_for_forward_simulation(essa_world* w) {
    if (!w)
        return nullptr;
    essa_world* nw = new essa_world(0);
    try {
        w->world.clone_for_forward_simulation(nw->world);
    } catch (std::exception const& e) {
        w->err = e.what();
        delete nw;
        return nullptr;
    }
    return nw;
}

int essa_world_is_forward_simulated(const essa_world* w) { return w && w->world.is_forward_simulated(); }

int essa_world_closest_approach(essa_world* w, int index, int other, double out[7]) {
    WANT(w && out && index >= 0 && other >= 0 && index < (int)w->world.object_count() && other < (int)w->world.object_count(), "bad index");
    GUARD(auto e = w->world.object_at(index)->closest_approach(*w->world.object_at(other)); if (!e) return 0; out[0] = e->distance;
          out[1] = e->this_position.x; out[2] = e->this_position.y; out[3] = e->this_position.z; out[4] = e->other_object_position.x;
          out[5] = e->other_object_position.y; out[6] = e->other_object_position.z;)
    return 1;
}

essa_ctx* essa_world_context(essa_world* w) {
    if (!w)
        return nullptr;
    try {
        return w->world.context();
    } catch (std::exception const& e) {
        w->err = e.what();
        return nullptr;
    }
}

} // extern "C"
