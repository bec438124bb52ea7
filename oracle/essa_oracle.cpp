// ============================================================================================
// essa_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see essa_oracle.hpp for the rules).
//
//   1. the .essa loader restatement (src/ConfigLoader.cpp),
//   2. a C ABI over oracle::World for ctypes (tests/, bench.py cpu_baseline, smoke()),
//   3. SoA helpers that evaluate the reference's force law on plain arrays for large N
//      (reference order, extended precision, OpenMP-over-rows variant for honest CPU baselines).
// ============================================================================================
#include "essa_oracle.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

namespace {

// Util::TextReader stand-in.
struct Reader {
    std::string const& s;
    size_t p = 0;
    bool eof() const { return p >= s.size(); }
    int peek() const { return eof() ? -1 : static_cast<unsigned char>(s[p]); }
    int consume() { return eof() ? -1 : static_cast<unsigned char>(s[p++]); }
    template<class F> std::string consume_while(F f) {
        size_t b = p;
        while (!eof() && f(static_cast<unsigned char>(s[p])))
            p++;
        return s.substr(b, p - b);
    }
    void skip_ws() { consume_while([](int c) { return isspace(c) != 0; }); }
    std::string where() const {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < p && i < s.size(); i++) {
            if (s[i] == '\n') { line++; col = 1; } else col++;
        }
        return std::to_string(line) + ":" + std::to_string(col);
    }
};

struct Error {
    std::string message;
};

// PropertyMap -- ConfigLoader.cpp:32-108
struct PropertyMap {
    std::map<std::string, std::string> props;
    bool contains(std::string const& k) const { return props.count(k) != 0; }
    std::string get(std::string const& k, std::string const& d = "") const {
        auto it = props.find(k);
        return it == props.end() ? d : it->second;
    }
    double get_double(std::string const& k, double d = 0) const {
        auto it = props.find(k);
        if (it == props.end())
            return d;
        char* end = nullptr;
        double v = std::strtod(it->second.c_str(), &end);
        if (end == it->second.c_str())
            throw Error { "Invalid number" };
        return v;
    }
    // ConfigLoader.cpp:61-78 -- Distance holds a float; unit suffix after '_'.
    float get_distance(std::string const& k) const {
        auto it = props.find(k);
        if (it == props.end())
            return 0.0f;
        auto us = it->second.find('_');
        if (us == std::string::npos)
            return static_cast<float>(get_double(k));
        std::string unit = it->second.substr(us + 1);
        std::string num = it->second.substr(0, us);
        char* end = nullptr;
        float value = std::strtof(num.c_str(), &end);
        if (end == num.c_str())
            throw Error { "Invalid number" };
        if (unit == "m")
            return value;
        if (unit == "km")
            return value * 1000;
        if (unit == "AU")
            return static_cast<float>(value * kAU);
        throw Error { "Invalid unit" };
    }
    int get_int(std::string const& k, int d) const {
        auto it = props.find(k);
        if (it == props.end())
            return d;
        try {
            return std::stoi(it->second);
        } catch (...) {
            throw Error { k + " is an invalid number" };
        }
    }
};

std::pair<std::string, std::string> parse_key_value_pair(Reader& r) { // ConfigLoader.cpp:187-202
    r.skip_ws();
    std::string name = r.consume_while([](int c) { return isalpha(c) || c == '_'; });
    r.skip_ws();
    if (r.consume() != '=')
        throw Error { "Expected '=' at " + r.where() };
    r.skip_ws();
    std::string value = r.consume_while([](int c) { return c != ';' && !isspace(c); });
    return { name, value };
}

PropertyMap read_properties(Reader& r) { // ConfigLoader.cpp:118-132
    PropertyMap pm;
    while (true) {
        auto kv = parse_key_value_pair(r);
        pm.props.insert(kv);
        r.skip_ws();
        int pk = r.peek();
        if (pk < 0)
            throw Error { "Expected ';' after property list at " + r.where() };
        if (!isalpha(pk))
            return pm;
    }
}

double to_rad(double deg, bool as_float) { // Util::Angle::degrees -- representation unverified
    double rad = deg * M_PI / 180.0;
    return as_float ? static_cast<double>(static_cast<float>(rad)) : rad;
}

} // namespace

std::string load_essa(std::string const& text, World& world, bool angle_as_float) {
    Reader r { text };
    try {
        r.skip_ws();
        while (!r.eof()) { // ConfigLoader.cpp:17-30
            // ---- parse_statement, ConfigLoader.cpp:110-185, applied immediately (Config::apply :204-270)
            std::string keyword = r.consume_while([](int c) { return isalpha(c) || c == '_'; });
            if (keyword.empty())
                throw Error { "Expected keyword at " + r.where() };
            r.skip_ws();
            if (keyword == "planet") {
                auto pm = read_properties(r);
                double mass = pm.get_double("mass");
                float radius = pm.get_distance("radius");
                std::string name = pm.get("name");
                double trail_length = pm.get_int("trail_length", 600);
                Vec3 pos { pm.get_double("posx"), pm.get_double("posy"), pm.get_double("posz") };
                Vec3 vel { pm.get_double("velx"), pm.get_double("vely"), pm.get_double("velz") };
                world.add_object(std::make_unique<Object>(mass, radius, pos, vel, name, static_cast<unsigned>(trail_length)));
            } else if (keyword == "orbiting_planet") {
                auto pm = read_properties(r);
                double mass = pm.get_double("mass");
                float radius = pm.get_distance("radius");
                std::string name = pm.get("name");
                double theta = to_rad(pm.get_double("orbit_position"), angle_as_float);
                double alpha = to_rad(pm.get_double("orbit_tilt"), angle_as_float);
                std::string around_name = pm.get("around");
                bool clockwise = !(pm.get("direction") == "left");
                bool ap_pe = pm.contains("apoapsis") && pm.contains("periapsis");
                bool maj_ecc = pm.contains("major_axis") && pm.contains("eccentrity");
                if (!ap_pe && !maj_ecc)
                    throw Error { "orbiting_planet must define apoapsis+periapsis or major_axis+eccentrity" };
                Object* around = world.get_object_by_name(around_name);
                if (!around)
                    throw Error { "Invalid planet for orbiting around: '" + around_name + "'" };
                if (ap_pe) {
                    world.add_object(around->create_object_relative_to_ap_pe(mass, radius, pm.get_distance("apoapsis"),
                        pm.get_distance("periapsis"), clockwise, theta, alpha, name, 0.0));
                } else {
                    world.add_object(around->create_object_relative_to_maj_ecc(mass, radius, pm.get_distance("major_axis"),
                        pm.get_double("eccentrity"), clockwise, theta, alpha, name, 0.0));
                }
            } else if (keyword == "light_source") {
                std::string name = r.consume_while([](int c) { return isalnum(c) != 0; });
                if (!world.get_object_by_name(name))
                    throw Error { "Invalid planet for light source: '" + name + "'" };
            } else {
                throw Error { "Invalid statement keyword: '" + keyword + "'" };
            }
            r.skip_ws();
            if (r.peek() != ';')
                throw Error { "Expected ';' at " + r.where() };
            r.consume();
            r.skip_ws();
        }
    } catch (Error const& e) {
        return e.message;
    }
    return {};
}

} // namespace oracle

// ============================================================================================
// C ABI for ctypes
// ============================================================================================
using oracle::Object;
using oracle::Vec3;
using oracle::World;
using oracle::History;
using oracle::Trail;

extern "C" {

void* ow_create() { return new World(); }
void ow_destroy(void* w) { delete static_cast<World*>(w); }

int ow_load_essa_file(void* w, char const* path, int angle_as_float, char* err, int errlen) {
    std::ifstream f(path);
    if (!f) {
        std::snprintf(err, errlen, "cannot open %s", path);
        return -1;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    std::string e = oracle::load_essa(ss.str(), *static_cast<World*>(w), angle_as_float != 0);
    if (!e.empty()) {
        std::snprintf(err, errlen, "%s", e.c_str());
        return -2;
    }
    return 0;
}

int ow_add_object(void* w, double mass, double radius, double const* pos, double const* vel, char const* name, unsigned period) {
    auto* world = static_cast<World*>(w);
    world->add_object(std::make_unique<Object>(mass, radius, Vec3 { pos[0], pos[1], pos[2] }, Vec3 { vel[0], vel[1], vel[2] }, name, period));
    return static_cast<int>(world->m_object_list.size()) - 1;
}

void ow_set_dt(void* w, int dt) { static_cast<World*>(w)->m_simulation_seconds_per_tick = dt; }
int ow_get_dt(void* w) { return static_cast<World*>(w)->m_simulation_seconds_per_tick; }
void ow_set_flags(void* w, int offset_trails, int legacy_euler) {
    auto* world = static_cast<World*>(w);
    world->offset_trails = offset_trails != 0;
    world->legacy_euler = legacy_euler != 0;
}
void ow_update(void* w, int steps) { static_cast<World*>(w)->update(steps); }
void ow_set_forces(void* w) { static_cast<World*>(w)->set_forces(); }
int ow_count(void* w) { return static_cast<int>(static_cast<World*>(w)->m_object_list.size()); }
long long ow_date(void* w) { return static_cast<World*>(w)->m_date; }
void ow_set_date(void* w, long long d) { static_cast<World*>(w)->m_date = d; }
int ow_is_forward_simulated(void* w) { return static_cast<World*>(w)->m_is_forward_simulated; }

// pos/vel/acc: 3N doubles (xyz interleaved); gf: N; most: N (list index or -1); appe: 4N
// (ap, pe, ap_vel, pe_vel); active: N; maxattr: N.  Any pointer may be null.
void ow_get_state(void* w, double* pos, double* vel, double* acc, double* gf, int* most, double* appe, int* active, double* maxattr) {
    auto* world = static_cast<World*>(w);
    std::map<Object const*, int> index;
    int k = 0;
    for (auto& o : world->m_object_list)
        index[o.get()] = k++;
    k = 0;
    for (auto& o : world->m_object_list) {
        if (pos) { pos[3 * k] = o->m_pos.x; pos[3 * k + 1] = o->m_pos.y; pos[3 * k + 2] = o->m_pos.z; }
        if (vel) { vel[3 * k] = o->m_vel.x; vel[3 * k + 1] = o->m_vel.y; vel[3 * k + 2] = o->m_vel.z; }
        if (acc) { acc[3 * k] = o->m_attraction_factor.x; acc[3 * k + 1] = o->m_attraction_factor.y; acc[3 * k + 2] = o->m_attraction_factor.z; }
        if (gf) gf[k] = o->m_gravity_factor;
        if (most) most[k] = o->m_most_attracting_object ? index.at(o->m_most_attracting_object) : -1;
        if (appe) { appe[4 * k] = o->m_ap; appe[4 * k + 1] = o->m_pe; appe[4 * k + 2] = o->m_ap_vel; appe[4 * k + 3] = o->m_pe_vel; }
        if (active) active[k] = o->deleted() ? 0 : 1;
        if (maxattr) maxattr[k] = o->m_max_attraction;
        k++;
    }
}

void ow_set_object_state(void* w, int idx, double const* pos, double const* vel, double gf) {
    Object* o = static_cast<World*>(w)->object_at(idx);
    if (pos) o->m_pos = { pos[0], pos[1], pos[2] };
    if (vel) o->m_vel = { vel[0], vel[1], vel[2] };
    if (gf >= 0) o->m_gravity_factor = gf;
}
// m_creation_date / m_deletion_date / m_deleted of one object (what add_object / delete_object set).
void ow_set_dates(void* w, int idx, long long creation, long long deletion, int deleted) {
    Object* o = static_cast<World*>(w)->object_at(idx);
    o->m_creation_date = creation;
    o->m_deletion_date = deletion;
    o->m_deleted = deleted != 0;
}
void ow_delete_object(void* w, int idx) { static_cast<World*>(w)->object_at(idx)->delete_object(); }
// World::delete_object_by_ptr (src/World.cpp:208-219)
void ow_remove_object(void* w, int idx) {
    auto* world = static_cast<World*>(w);
    Object* ptr = world->object_at(idx);
    world->m_object_list.remove_if([ptr](std::unique_ptr<Object>& obj) { return obj.get() == ptr; });
    for (auto& o : world->m_object_list)
        if (o->m_most_attracting_object == ptr)
            o->delete_most_attracting_object();
}
void ow_reset_history(void* w, int idx) { static_cast<World*>(w)->object_at(idx)->reset_history(); }
// Seeds the ObjectHistory (src/ObjectHistory.cpp:20-23 on the list's last object): the reference's current revision never does
// this itself (World.cpp:76-81 pushes only when the history already holds an entry), so its date-driven push / pop logic
// (World.cpp:59-84) can only be exercised from outside.
void ow_debug_stash_last(void* w) {
    auto* world = static_cast<World*>(w);
    world->m_object_history.push_to_entry(std::move(world->m_object_list.back()));
    world->m_object_list.pop_back();
}
int ow_object_history_size(void* w) { return (int)static_cast<World*>(w)->m_object_history.size(); }
char const* ow_name(void* w, int idx) { return static_cast<World*>(w)->object_at(idx)->m_name.c_str(); }
double ow_radius(void* w, int idx) { return static_cast<World*>(w)->object_at(idx)->m_radius; }
double ow_orbit_len(void* w, int idx) { return static_cast<World*>(w)->object_at(idx)->m_orbit_len; }
double ow_eccentrity(void* w, int idx) { return static_cast<World*>(w)->object_at(idx)->m_eccentrity; }

int ow_trail_capacity(void* w, int idx) { return static_cast<int>(static_cast<World*>(w)->object_at(idx)->m_trail.vertexes().size()); }
// verts: 3*capacity floats; meta: {append_offset, length}; offset: 3 doubles; last_angle: 1 double
void ow_trail_get(void* w, int idx, float* verts, int* meta, double* offset, double* last_angle) {
    auto const& t = static_cast<World*>(w)->object_at(idx)->m_trail;
    if (verts) {
        size_t k = 0;
        for (auto const& v : t.vertexes()) {
            verts[k++] = v.x;
            verts[k++] = v.y;
            verts[k++] = v.z;
        }
    }
    if (meta) { meta[0] = t.append_offset(); meta[1] = t.length(); }
    if (offset) { offset[0] = t.offset().x; offset[1] = t.offset().y; offset[2] = t.offset().z; }
    if (last_angle) *last_angle = t.last_xy_angle();
}

int ow_history_len(void* w, int idx) { return static_cast<int>(static_cast<World*>(w)->object_at(idx)->m_history.entries().size()); }
// out: 6*len doubles, oldest first: px py pz vx vy vz
void ow_history_get(void* w, int idx, double* out) {
    size_t k = 0;
    for (auto const& e : static_cast<World*>(w)->object_at(idx)->m_history.entries()) {
        out[k++] = e.pos.x; out[k++] = e.pos.y; out[k++] = e.pos.z;
        out[k++] = e.vel.x; out[k++] = e.vel.y; out[k++] = e.vel.z;
    }
}

void* ow_clone_for_forward_simulation(void* w) {
    auto* nw = new World();
    static_cast<World*>(w)->clone_for_forward_simulation(*nw);
    return nw;
}
// out: {distance, this_pos xyz, other_pos xyz}; returns 0 if no entry
int ow_closest_approach(void* w, int idx, int other, double* out) {
    auto* world = static_cast<World*>(w);
    Object* o = world->object_at(idx);
    auto it = o->m_closest_approaches.find(world->object_at(other));
    if (it == o->m_closest_approaches.end())
        return 0;
    auto const& e = it->second;
    out[0] = e.distance;
    out[1] = e.this_position.x; out[2] = e.this_position.y; out[3] = e.this_position.z;
    out[4] = e.other_object_position.x; out[5] = e.other_object_position.y; out[6] = e.other_object_position.z;
    return 1;
}

// Wall-clock of `steps` World::update iterations on this thread (the timed CPU baseline).
double ow_time_update(void* w, int steps) {
    auto t0 = std::chrono::steady_clock::now();
    static_cast<World*>(w)->update(steps);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ============================================================================================
// SoA helpers (large N).  Arithmetic identical to Object::update_forces_against (Object.cpp:64-79).
// ============================================================================================

// World::set_forces order on arrays: pair-shared base, i<j, a_i -= base*gf_j ; a_j += base*gf_i.
void osoa_forces_pairshared(int n, double const* x, double const* y, double const* z, double const* gf, double* ax, double* ay, double* az) {
    for (int i = 0; i < n; i++)
        ax[i] = ay[i] = az[i] = 0;
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            double dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];
            double den = dx * dx + dy * dy + dz * dz;
            den *= std::sqrt(den);
            double bx = dx / den, by = dy / den, bz = dz / den;
            ax[i] -= bx * gf[j]; ay[i] -= by * gf[j]; az[i] -= bz * gf[j];
            ax[j] += bx * gf[i]; ay[j] += by * gf[i]; az[j] += bz * gf[i];
        }
    }
}

// Per-row evaluation in ascending partner order.  Bit-identical to what set_forces leaves in
// row k (partners p<k arrive as `+= base*gf_p` with base from x_p-x_k, partners p>k as
// `-= base*gf_p` with base from x_k-x_p; IEEE negation is exact, so both are the same addend).
static inline void row_force(int n, double const* x, double const* y, double const* z, double const* gf, int k, double* out) {
    double ax = 0, ay = 0, az = 0;
    for (int p = 0; p < k; p++) {
        double dx = x[p] - x[k], dy = y[p] - y[k], dz = z[p] - z[k];
        double den = dx * dx + dy * dy + dz * dz;
        den *= std::sqrt(den);
        ax += (dx / den) * gf[p]; ay += (dy / den) * gf[p]; az += (dz / den) * gf[p];
    }
    for (int p = k + 1; p < n; p++) {
        double dx = x[k] - x[p], dy = y[k] - y[p], dz = z[k] - z[p];
        double den = dx * dx + dy * dy + dz * dz;
        den *= std::sqrt(den);
        ax -= (dx / den) * gf[p]; ay -= (dy / den) * gf[p]; az -= (dz / den) * gf[p];
    }
    out[0] = ax; out[1] = ay; out[2] = az;
}

void osoa_forces_rows(int n, double const* x, double const* y, double const* z, double const* gf, int const* rows, int nrows, double* out) {
    for (int r = 0; r < nrows; r++)
        row_force(n, x, y, z, gf, rows[r], out + 3 * r);
}

// Same, OpenMP over rows: NOT the reference (which is single-threaded); an honest all-cores CPU number.
int osoa_forces_rows_omp(int n, double const* x, double const* y, double const* z, double const* gf, int const* rows, int nrows, double* out) {
    int threads = 1;
#pragma omp parallel
    {
#pragma omp single
        {
#ifdef _OPENMP
            threads = omp_get_num_threads();
#endif
        }
#pragma omp for schedule(dynamic, 1)
        for (int r = 0; r < nrows; r++)
            row_force(n, x, y, z, gf, rows[r], out + 3 * r);
    }
    return threads;
}

// Extended precision (x87 long double, 64-bit mantissa) evaluation of the same rows: the yardstick
// for the self-calibrating tolerance of SURVEY.md section 8c(ii).
void osoa_forces_rows_ld(int n, double const* x, double const* y, double const* z, double const* gf, int const* rows, int nrows, double* out) {
    for (int r = 0; r < nrows; r++) {
        int k = rows[r];
        long double ax = 0, ay = 0, az = 0;
        for (int p = 0; p < n; p++) {
            if (p == k)
                continue;
            long double dx = (long double)x[p] - x[k], dy = (long double)y[p] - y[k], dz = (long double)z[p] - z[k];
            long double r2 = dx * dx + dy * dy + dz * dz;
            long double s = gf[p] / (r2 * sqrtl(r2));
            ax += dx * s; ay += dy * s; az += dz * s;
        }
        out[3 * r] = (double)ax; out[3 * r + 1] = (double)ay; out[3 * r + 2] = (double)az;
    }
}

// Reference-semantics KDK on arrays, all bodies active, no recording: two force passes per step
// (World.cpp:104,119), kick/drift as separate multiply-then-add (World.cpp:112,115,127).
void osoa_step(int n, double* x, double* y, double* z, double* vx, double* vy, double* vz, double const* gf,
    double* ax, double* ay, double* az, int steps, int dt) {
    double step = dt, half = step / 2.0, mul = steps >= 0 ? 1 : -1;
    for (int s = 0; s < std::abs(steps); s++) {
        osoa_forces_pairshared(n, x, y, z, gf, ax, ay, az);
        for (int i = 0; i < n; i++) {
            vx[i] = vx[i] + ax[i] * half * mul; vy[i] = vy[i] + ay[i] * half * mul; vz[i] = vz[i] + az[i] * half * mul;
            x[i] = x[i] + vx[i] * step * mul; y[i] = y[i] + vy[i] * step * mul; z[i] = z[i] + vz[i] * step * mul;
        }
        osoa_forces_pairshared(n, x, y, z, gf, ax, ay, az);
        for (int i = 0; i < n; i++) {
            vx[i] = vx[i] + ax[i] * half * mul; vy[i] = vy[i] + ay[i] * half * mul; vz[i] = vz[i] + az[i] * half * mul;
        }
    }
}

// Kinetic and potential energy (per unit G already folded: gf = G m).  E = sum 1/2 m v^2 - sum_{i<j} G m_i m_j / r
// evaluated in long double.  `gravity` = G so that m = gf / G.
void osoa_energy(int n, double const* x, double const* y, double const* z, double const* vx, double const* vy, double const* vz,
    double const* gf, double gravity, double* kinetic, double* potential, double* momentum) {
    long double ke = 0, pe = 0, px = 0, py = 0, pz = 0;
    for (int i = 0; i < n; i++) {
        long double m = (long double)gf[i] / gravity;
        ke += 0.5L * m * ((long double)vx[i] * vx[i] + (long double)vy[i] * vy[i] + (long double)vz[i] * vz[i]);
        px += m * vx[i]; py += m * vy[i]; pz += m * vz[i];
    }
#pragma omp parallel for reduction(+ : pe) schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        long double row = 0;
        for (int j = i + 1; j < n; j++) {
            long double dx = (long double)x[i] - x[j], dy = (long double)y[i] - y[j], dz = (long double)z[i] - z[j];
            row += (long double)gf[j] / sqrtl(dx * dx + dy * dy + dz * dz);
        }
        pe -= row * ((long double)gf[i] / gravity);
    }
    *kinetic = (double)ke;
    *potential = (double)pe;
    if (momentum) { momentum[0] = (double)px; momentum[1] = (double)py; momentum[2] = (double)pz; }
}

// Timed slice of the reference's pair loop on its own containers (std::list<unique_ptr<Object>>):
// rows [i0, i1) of World::set_forces's outer loop against all later objects.  Returns seconds;
// *pairs = number of update_forces_against calls made.
double ow_time_force_rows(void* w, int i0, int i1, long long* pairs) {
    auto* world = static_cast<World*>(w);
    auto& list = world->m_object_list;
    long long count = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (auto& p : list)
        if (!p->deleted())
            p->clear_forces();
    auto it = list.begin();
    std::advance(it, i0);
    for (int i = i0; i < i1 && it != list.end(); i++, it++) {
        auto it2 = it;
        auto& this_object = **it;
        it2++;
        for (; it2 != list.end(); it2++) {
            if (!this_object.deleted() && !(*it2)->deleted()) {
                this_object.update_forces_against(**it2);
                count++;
            }
        }
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (pairs)
        *pairs = count;
    return dt;
}

// Bulk construction for large synthetic worlds (avoids one ctypes call per body).
void ow_add_bodies(void* w, int n, double const* x, double const* y, double const* z, double const* vx, double const* vy, double const* vz, double const* mass) {
    auto* world = static_cast<World*>(w);
    for (int i = 0; i < n; i++) {
        auto o = std::make_unique<Object>(mass[i], 1.0, Vec3 { x[i], y[i], z[i] }, Vec3 { vx[i], vy[i], vz[i] }, "b" + std::to_string(i), 1);
        o->m_world = world;
        o->m_creation_date = world->m_date;
        world->m_object_list.push_back(std::move(o)); // add_object's remove_if is O(N) per call; same end state
    }
}

// ---- the oracle's own History / Trail classes behind the C ABI of ref_driver.cpp (same signatures, prefix o instead of ref):
// tests drive both with identical inputs and demand identical bits (tests/test_oracle_ref.py) ----
void* o_history_new(unsigned segments, double const* first) {
    return new History(segments, { { first[0], first[1], first[2] }, { first[3], first[4], first[5] } });
}
void o_history_free(void* h) { delete static_cast<History*>(h); }
static int o_give(std::optional<History::Entry> const& e, double* out) {
    if (!e)
        return 0;
    double v[6] = { e->pos.x, e->pos.y, e->pos.z, e->vel.x, e->vel.y, e->vel.z };
    std::memcpy(out, v, sizeof(v));
    return 1;
}
int o_history_move_forward(void* h, double const* e, double* out) {
    return o_give(static_cast<History*>(h)->move_forward({ { e[0], e[1], e[2] }, { e[3], e[4], e[5] } }), out);
}
int o_history_move_backward(void* h, double const* e, double* out) {
    return o_give(static_cast<History*>(h)->move_backward({ { e[0], e[1], e[2] }, { e[3], e[4], e[5] } }), out);
}
void o_history_reset(void* h) { static_cast<History*>(h)->reset(); }
int o_history_read(void* h, double* out, int max_entries, int* side) {
    auto* H = static_cast<History*>(h);
    int k = 0;
    for (auto const& e : H->entries()) {
        if (k < max_entries) {
            double v[6] = { e.pos.x, e.pos.y, e.pos.z, e.vel.x, e.vel.y, e.vel.z };
            std::memcpy(out + 6 * k, v, sizeof(v));
        }
        k++;
    }
    if (side)
        *side = H->side() ? 1 : 0;
    return k;
}
void* o_trail_new(size_t max_trail_size) { return new Trail(max_trail_size); }
void o_trail_free(void* t) { delete static_cast<Trail*>(t); }
void o_trail_push_back(void* t, double x, double y, double z) { static_cast<Trail*>(t)->push_back({ x, y, z }); }
void o_trail_change_current(void* t, double x, double y, double z) { static_cast<Trail*>(t)->change_current({ x, y, z }); }
void o_trail_set_offset(void* t, double x, double y, double z) { static_cast<Trail*>(t)->set_offset({ x, y, z }); }
void o_trail_recalculate_with_offset(void* t, double x, double y, double z) { static_cast<Trail*>(t)->recalculate_with_offset({ x, y, z }); }
void o_trail_reset(void* t) { static_cast<Trail*>(t)->reset(); }
void o_trail_read(void* t, float* verts, int* meta, double* real) {
    auto* T = static_cast<Trail*>(t);
    if (verts)
        for (size_t i = 0; i < T->vertexes().size(); i++) {
            verts[3 * i] = T->vertexes()[i].x;
            verts[3 * i + 1] = T->vertexes()[i].y;
            verts[3 * i + 2] = T->vertexes()[i].z;
        }
    meta[0] = (int)T->vertexes().size();
    meta[1] = T->append_offset();
    meta[2] = T->length();
    real[0] = T->offset().x;
    real[1] = T->offset().y;
    real[2] = T->offset().z;
    real[3] = T->last_xy_angle();
}

} // extern "C"
