// World / Object host facade over the C ABI (see facade.hpp).  Host code only; every numerical
// step of World::update happens in the CUDA kernels behind essa_step().
#include "facade.hpp"

#include <algorithm>
#include <cassert>
#include <cstring>

namespace essa {

// ---------------------------------------------------------------------------------------------
// Trail (host copy).  Only the constructor-time push and the rare host-side edits
// (reset, delete_most_attracting_object) run here; per-step pushes run on the device.
// ---------------------------------------------------------------------------------------------
static float vertex_component(double p) {
    // pos.cast<float>() / AU  (src/Trail.cpp:53,64)
    return static_cast<float>(static_cast<double>(static_cast<float>(p)) / ESSA_AU);
}

Trail::Trail(size_t max_trail_size) {
    if (max_trail_size < 3)
        max_trail_size = 3; // src/Trail.cpp:23-26
    if (max_trail_size > 0xfffff)
        max_trail_size = 0xfffff; // src/Trail.cpp:28-31
    vertexes.assign(max_trail_size * 3, 0.0f);
    length = 1;
}

void Trail::push_back(Vec3 const& pos) {
    auto put = [&](int slot) {
        vertexes[3 * slot] = vertex_component(pos.x);
        vertexes[3 * slot + 1] = vertex_component(pos.y);
        vertexes[3 * slot + 2] = vertex_component(pos.z);
    };
    int const cap = static_cast<int>(size());
    if (length == 1) {
        put(append_offset);
        append_offset++;
        length++;
    }
    if (length >= 3 && append_offset != 0) {
        if (std::fabs(std::atan2(pos.y, pos.x) - last_xy_angle) <= 2 * M_PI / cap) {
            if (append_offset == 1)
                put(length - 1);
            put(append_offset - 1);
            return;
        }
    }
    put(append_offset);
    append_offset++;
    if (length < cap)
        length++;
    if (append_offset == cap) {
        put(0);
        append_offset = 1;
    }
    last_xy_angle = std::atan2(pos.y, pos.x);
}

void Trail::recalculate_with_offset(Vec3 const& o) {
    if (offset == o)
        return;
    float a[3] = { vertex_component(offset.x), vertex_component(offset.y), vertex_component(offset.z) };
    float b[3] = { vertex_component(o.x), vertex_component(o.y), vertex_component(o.z) };
    for (int s = 0; s < length; s++)
        for (int c = 0; c < 3; c++)
            vertexes[3 * s + c] = vertexes[3 * s + c] + a[c] - b[c];
    offset = o;
}

// ---------------------------------------------------------------------------------------------
// Object
// ---------------------------------------------------------------------------------------------
Object::Object(double mass, double radius, Vec3 pos, Vec3 vel, Color color, std::string name, unsigned period)
    : m_trail(std::max(2U, std::max(period * 2, 500U)))
    , m_first_entry { pos, vel }
    , m_pos(pos)
    , m_vel(vel)
    , m_orbit_len(period)
    , m_gravity_factor(mass * ESSA_GRAVITY)
    , m_radius(radius)
    , m_name(std::move(name))
    , m_color(color) {
    m_trail_capacity = static_cast<unsigned>(m_trail.size());
    m_history.push_back(m_first_entry); // History(1000, { pos, vel })
    m_trail.push_back(m_pos);
}

void Object::touch() {
    if (m_world)
        m_world->mark_state_dirty();
}

void Object::set_pos(Vec3 const& pos) {
    m_pos = pos;
    touch();
}

void Object::set_vel(Vec3 const& vel) {
    m_vel = vel;
    touch();
}

void Object::set_gravity_factor(double gf) {
    m_gravity_factor = gf;
    touch();
}

Vec3 Object::attraction(Object const& other) const {
    Vec3 dist = m_pos - other.m_pos;
    double force = other.m_gravity_factor / dist.length_squared();
    return dist.normalized() * force;
}

bool Object::deleted() const { return (m_deleted && m_deletion_date <= m_world->date()) || m_creation_date > m_world->date(); }

void Object::delete_object() {
    m_deletion_date = m_world->date();
    m_deleted = true;
    touch();
}

void Object::delete_most_attracting_object() {
    trail(); // bring the ring to the host, edit, send back on the next sync
    m_most_attracting_object = nullptr;
    m_trail.recalculate_with_offset({});
    m_trail.reset();
    m_rings_on_device = false;
    if (m_world)
        m_world->mark_list_dirty();
}

Object::Info Object::get_info() const {
    Info info { mass(), m_radius, m_vel.length() };
    if (m_most_attracting_object) {
        info.distance_from_most_massive_object = get_distance(m_pos, m_most_attracting_object->m_pos);
        info.apoapsis = m_ap;
        info.apoapsis_velocity = m_ap_vel;
        info.periapsis = m_pe;
        info.periapsis_velocity = m_pe_vel;
        info.orbit_period = m_orbit_len;
        info.orbit_eccentrity = m_eccentrity;
    }
    return info;
}

void Object::reset_history() {
    trail();
    history();
    m_history.assign(1, m_first_entry);
    m_trail.reset();
    m_rings_on_device = false;
    if (m_world)
        m_world->mark_list_dirty();
}

Trail const& Object::trail() {
    if (m_rings_on_device && m_world)
        m_world->pull_rings(*this);
    return m_trail;
}

std::vector<HistoryEntry> const& Object::history() {
    if (m_rings_on_device && m_world)
        m_world->pull_rings(*this);
    return m_history;
}

std::unique_ptr<Object> Object::create_object_relative_to_ap_pe(double mass, float radius, float apoapsis, float periapsis,
    bool direction, double theta, double alpha, Color color, std::string name, double rotation) const {
    double const GM = m_gravity_factor;
    double const a = (apoapsis + periapsis) / 2;
    double const b = std::sqrt(apoapsis * periapsis);
    double const T = 2 * M_PI * std::sqrt((a * a * a) / GM);
    Vec3 pos { std::cos(theta) * std::cos(alpha) * a, std::sin(theta) * std::cos(alpha) * a, std::sin(alpha) * a };
    pos = pos.rotate_z(rotation);
    pos += m_pos;

    auto result = std::make_unique<Object>(mass, radius, pos, Vec3 {}, color, std::move(name), static_cast<unsigned>(T / (3600 * 24)));
    result->m_ap = apoapsis;
    result->m_pe = periapsis;
    result->m_eccentrity = std::sqrt(1 - (b * b) / (a * a));

    double const velocity_constant = (2 * GM) / (apoapsis + periapsis);
    result->m_ap_vel = std::sqrt(2 * GM / apoapsis - velocity_constant);
    result->m_pe_vel = std::sqrt(2 * GM / periapsis - velocity_constant);
    double const velocity = std::sqrt(2 * GM / get_distance(pos, m_pos) - velocity_constant);

    Vec3 vel { std::cos(theta + M_PI / 2) * velocity, std::sin(theta + M_PI / 2) * velocity, 0 };
    if (direction)
        vel = -vel;
    vel += m_vel;
    result->m_vel = vel;
    return result;
}

std::unique_ptr<Object> Object::create_object_relative_to_maj_ecc(double mass, float radius, float semi_major, double ecc,
    bool direction, double theta, double alpha, Color color, std::string name, double rotation) const {
    double const GM = m_gravity_factor;
    double const a = semi_major;
    double const T = 2 * M_PI * std::sqrt((a * a * a) / GM);
    Vec3 pos { std::cos(theta) * std::cos(alpha) * a, std::sin(theta) * std::cos(alpha) * a, std::sin(alpha) * a };
    pos = pos.rotate_z(rotation);
    pos += m_pos;

    auto result = std::make_unique<Object>(mass, radius, pos, Vec3 {}, color, std::move(name), static_cast<unsigned>(T / (3600 * 24)));
    result->m_ap = 0;
    result->m_pe = std::numeric_limits<double>::max();
    result->m_eccentrity = ecc;

    double const velocity_constant = (2 * GM) / (a * 2 - a * ecc);
    result->m_ap_vel = std::sqrt(2 * GM / a - velocity_constant);
    result->m_pe_vel = std::sqrt(2 * GM / (a - a * ecc) - velocity_constant);
    double const velocity = std::sqrt(2 * GM / get_distance(pos, m_pos) - velocity_constant);

    Vec3 vel { std::cos(theta + M_PI / 2) * velocity, std::sin(theta + M_PI / 2) * velocity, 0 };
    if (direction)
        vel = -vel;
    vel += m_vel;
    result->m_vel = vel;
    return result;
}

std::unique_ptr<Object> Object::clone_for_forward_simulation() const {
    Color c { static_cast<uint8_t>(std::min(255, m_color.r + 60)), static_cast<uint8_t>(std::min(255, m_color.g + 60)),
        static_cast<uint8_t>(std::min(255, m_color.b + 60)), m_color.a };
    auto object = std::make_unique<Object>(0, m_radius, m_pos, m_vel, c, m_name, 500);
    object->m_is_forward_simulated = true;
    object->m_gravity_factor = m_gravity_factor;
    return object;
}

std::optional<Object::ClosestApproachEntry> Object::closest_approach(Object const& other) {
    if (!m_world || m_slot < 0 || other.m_slot < 0 || !m_world->m_ctx || !m_world->m_closest_enabled)
        return std::nullopt;
    double out[7];
    int rc = essa_closest_read(m_world->m_ctx, m_slot, other.m_slot, out);
    if (rc <= 0)
        return std::nullopt;
    ClosestApproachEntry e;
    e.distance = out[0];
    e.this_position = { out[1], out[2], out[3] };
    e.other_object_position = { out[4], out[5], out[6] };
    return e;
}

// ---------------------------------------------------------------------------------------------
// ObjectHistory
// ---------------------------------------------------------------------------------------------
void ObjectHistory::clear_history(unsigned long to_index) {
    m_entries.resize(std::min<size_t>(to_index, m_entries.size()));
    m_pos = static_cast<unsigned>(to_index - 1);
}

bool ObjectHistory::set_time(SimulationTime time) {
    m_pos = 0;
    for (auto& e : m_entries) {
        if (e->creation_date() > time)
            return true;
        m_pos++;
    }
    return false;
}

void ObjectHistory::push_to_entry(std::unique_ptr<Object> obj) {
    obj->reset_history();
    m_entries.push_back(std::move(obj));
}

std::unique_ptr<Object> ObjectHistory::pop_from_entries() {
    auto object = std::move(m_entries.back());
    m_entries.pop_back();
    return object;
}

// ---------------------------------------------------------------------------------------------
// World
// ---------------------------------------------------------------------------------------------
World::World(int device)
    : m_device(device)
    , m_start_date(ESSA_START_DATE)
    , m_date(ESSA_START_DATE) { }

World::~World() {
    if (m_ctx)
        essa_ctx_destroy(m_ctx);
}

void World::check(int rc, char const* what) {
    if (rc < 0)
        throw std::runtime_error(std::string(what) + ": " + essa_last_error(m_ctx));
}

essa_ctx* World::context() {
    if (!m_ctx) {
        int rc = essa_ctx_create(m_device, static_cast<int>(std::max<size_t>(m_object_list.size(), 1)), &m_ctx);
        if (rc != ESSA_OK)
            throw std::runtime_error(std::string("essa_ctx_create: ") + essa_last_error(nullptr));
    }
    return m_ctx;
}

double World::last_step_ms() const { return m_ctx ? essa_last_step_ms(m_ctx) : 0.0; }

void World::set_persistent(bool enable, int idle_timeout_us) {
    check(essa_persist_configure(context(), enable ? 1 : 0, idle_timeout_us), "essa_persist_configure");
}

void World::persist_stats(int64_t out[6]) {
    check(essa_persist_stats(context(), out), "essa_persist_stats");
}

void World::set_date(SimulationTime d) {
    m_date = d;
    m_state_dirty = true;
}

void World::add_object(std::unique_ptr<Object> object) {
    object->m_world = this;
    object->m_creation_date = m_date;
    m_object_list.push_back(std::move(object));
    if (m_object_history.set_time(m_date))
        m_object_history.clear_history(m_object_history.get_pos());
    m_object_list.remove_if([&](std::unique_ptr<Object>& obj) { return obj->creation_date() > m_date; });
    m_list_dirty = true;
}

Object* World::get_object_by_name(std::string const& name) {
    for (auto& obj : m_object_list)
        if (obj->name() == name)
            return obj.get();
    return nullptr;
}

bool World::exist_object_with_name(std::string const& name) const {
    for (auto const& obj : m_object_list)
        if (obj->name() == name && !obj->deleted())
            return true;
    return false;
}

bool World::set_light_source(std::string const& name) {
    m_light_source = get_object_by_name(name);
    return m_light_source != nullptr;
}

Object* World::object_at(size_t idx) {
    auto it = m_object_list.begin();
    std::advance(it, idx);
    return it->get();
}

int World::index_of(Object const* o) const {
    int k = 0;
    for (auto const& p : m_object_list) {
        if (p.get() == o)
            return k;
        k++;
    }
    return -1;
}

void World::reset(std::optional<std::string> const& filename) {
    if (on_reset)
        on_reset();
    m_object_list.clear();
    m_date = ESSA_START_DATE;
    m_object_history.clear_history(0);
    m_light_source = nullptr;
    m_list_dirty = true;
    if (filename) {
        std::string err = load_essa_file(*filename, *this);
        if (!err.empty())
            throw std::runtime_error("World loading failed: " + err);
    }
}

void World::reset_all_trails() {
    for_each_object([](Object& o) {
        o.trail();
        o.m_trail.reset();
        o.m_rings_on_device = false;
    });
    m_list_dirty = true;
}

void World::delete_object_by_ptr(Object* ptr) {
    // bring every survivor's rings home first: slots are about to be renumbered
    for (auto& o : m_object_list)
        if (o.get() != ptr)
            pull_rings(*o);
    m_object_list.remove_if([ptr](std::unique_ptr<Object>& obj) { return obj.get() == ptr; });
    for (auto& o : m_object_list)
        if (o->most_attracting_object() == ptr)
            o->delete_most_attracting_object();
    if (m_light_source == ptr)
        m_light_source = nullptr;
    m_list_dirty = true;
}

void World::clone_for_forward_simulation(World& new_world) const {
    // `new_world = World()` in the reference: everything back to defaults
    new_world.m_object_list.clear();
    new_world.m_object_history.clear_history(0);
    new_world.m_date = ESSA_START_DATE;
    new_world.m_start_date = ESSA_START_DATE;
    new_world.m_simulation_seconds_per_tick = 60 * 60 * 12;
    new_world.m_light_source = nullptr;
    new_world.m_is_forward_simulated = true;
    new_world.offset_trails = offset_trails;
    new_world.exact_order = exact_order;
    new_world.m_list_dirty = true;
    for (auto const& object : m_object_list)
        if (!object->deleted())
            new_world.add_object(object->clone_for_forward_simulation());
}

void World::pull_rings(Object& o) {
    if (!o.m_rings_on_device || !m_ctx || o.m_slot < 0)
        return;
    int n = essa_history_read(m_ctx, o.m_slot, nullptr);
    check(n, "essa_history_read");
    std::vector<double> buf(static_cast<size_t>(n) * 6);
    check(essa_history_read(m_ctx, o.m_slot, buf.data()), "essa_history_read");
    o.m_history.resize(n);
    for (int k = 0; k < n; k++)
        o.m_history[k] = { { buf[6 * k], buf[6 * k + 1], buf[6 * k + 2] }, { buf[6 * k + 3], buf[6 * k + 4], buf[6 * k + 5] } };
    essa_trail_meta meta;
    check(essa_trail_read(m_ctx, o.m_slot, o.m_trail.vertexes.data(), &meta), "essa_trail_read");
    o.m_trail.append_offset = meta.append_offset;
    o.m_trail.length = meta.length;
    o.m_trail.offset = { meta.offset[0], meta.offset[1], meta.offset[2] };
    o.m_trail.last_xy_angle = meta.last_xy_angle;
}

essa_step_opts World::step_opts() const {
    essa_step_opts o;
    std::memset(&o, 0, sizeof(o));
    o.dt_seconds = m_simulation_seconds_per_tick;
    o.forward_simulated = m_is_forward_simulated;
    o.offset_trails = offset_trails;
    o.record = 1;
    o.track_most_attracting = 1;
    o.exact_order = exact_order && m_object_list.size() <= ESSA_RESIDENT_MAX;
    o.two_pass = two_pass;
    o.kernel = ESSA_KERNEL_AUTO;
    o.eps2 = 0.0;
    return o;
}

void World::sync_to_device() {
    context();
    size_t const n = m_object_list.size();
    if (m_list_dirty) {
        // Slots are list positions.  Bring the rings of objects that already live on the device
        // home (by their old slot), renumber, upload, lay the rings out again and send them back.
        for (auto& o : m_object_list)
            pull_rings(*o);
        int slot = 0;
        for (auto& o : m_object_list)
            o->m_slot = slot++;
    }
    if (m_list_dirty || m_state_dirty) {
        std::vector<double> f(15 * n);
        std::vector<int64_t> dates(2 * n);
        std::vector<uint8_t> del(n);
        std::vector<int32_t> most(n);
        double *x = f.data(), *y = x + n, *z = y + n, *vx = z + n, *vy = vx + n, *vz = vy + n, *gf = vz + n, *ap = gf + n, *pe = ap + n,
               *apv = pe + n, *pev = apv + n, *ax = pev + n, *ay = ax + n, *az = ay + n, *mx = az + n;
        size_t k = 0;
        for (auto& o : m_object_list) {
            x[k] = o->m_pos.x; y[k] = o->m_pos.y; z[k] = o->m_pos.z;
            vx[k] = o->m_vel.x; vy[k] = o->m_vel.y; vz[k] = o->m_vel.z;
            gf[k] = o->m_gravity_factor;
            ap[k] = o->m_ap; pe[k] = o->m_pe; apv[k] = o->m_ap_vel; pev[k] = o->m_pe_vel;
            ax[k] = o->m_attraction_factor.x; ay[k] = o->m_attraction_factor.y; az[k] = o->m_attraction_factor.z;
            mx[k] = o->m_max_attraction;
            dates[k] = o->m_creation_date;
            dates[n + k] = o->m_deletion_date;
            del[k] = o->m_deleted;
            most[k] = o->m_most_attracting_object ? o->m_most_attracting_object->m_slot : -1;
            k++;
        }
        essa_bodies b;
        std::memset(&b, 0, sizeof(b));
        b.x = x; b.y = y; b.z = z; b.vx = vx; b.vy = vy; b.vz = vz; b.gf = gf;
        b.ap = ap; b.pe = pe; b.ap_vel = apv; b.pe_vel = pev;
        b.ax = ax; b.ay = ay; b.az = az; b.max_attraction = mx; // kept by bodies that deleted() masks
        b.creation_date = dates.data();
        b.deletion_date = dates.data() + n;
        b.deleted = del.data();
        b.most = most.data();
        check(essa_set_date(m_ctx, m_date), "essa_set_date");
        check(essa_upload(m_ctx, static_cast<int>(n), &b), "essa_upload");
    }
    if (m_list_dirty) {
        std::vector<int32_t> caps(n);
        size_t k = 0;
        for (auto& o : m_object_list)
            caps[k++] = static_cast<int32_t>(o->m_trail_capacity);
        check(essa_record_configure(m_ctx, static_cast<int>(n), caps.data(), 0), "essa_record_configure");
        for (auto& o : m_object_list) {
            std::vector<double> h(o->m_history.size() * 6);
            for (size_t e = 0; e < o->m_history.size(); e++) {
                auto const& en = o->m_history[e];
                double v[6] = { en.pos.x, en.pos.y, en.pos.z, en.vel.x, en.vel.y, en.vel.z };
                std::memcpy(&h[6 * e], v, sizeof(v));
            }
            double first[6] = { o->m_first_entry.pos.x, o->m_first_entry.pos.y, o->m_first_entry.pos.z, o->m_first_entry.vel.x,
                o->m_first_entry.vel.y, o->m_first_entry.vel.z };
            check(essa_history_write(m_ctx, o->m_slot, first, 1, 1), "essa_history_write"); // History::m_first_entry
            check(essa_history_write(m_ctx, o->m_slot, h.data(), static_cast<int>(o->m_history.size()), 0), "essa_history_write");
            essa_trail_meta meta;
            meta.capacity = static_cast<int32_t>(o->m_trail.size());
            meta.append_offset = o->m_trail.append_offset;
            meta.length = o->m_trail.length;
            meta.pad = 0;
            meta.offset[0] = o->m_trail.offset.x;
            meta.offset[1] = o->m_trail.offset.y;
            meta.offset[2] = o->m_trail.offset.z;
            meta.last_xy_angle = o->m_trail.last_xy_angle;
            check(essa_trail_write(m_ctx, o->m_slot, o->m_trail.vertexes.data(), &meta), "essa_trail_write");
            o->m_rings_on_device = true;
        }
        if (m_is_forward_simulated) {
            check(essa_closest_enable(m_ctx, 1), "essa_closest_enable");
            m_closest_enabled = true;
        }
    }
    m_list_dirty = false;
    m_state_dirty = false;
}

void World::sync_from_device() {
    size_t const n = m_object_list.size();
    if (n == 0)
        return;
    // scratch kept across calls: this runs once per GUI tick
    std::vector<double>& f = m_scratch_f;
    std::vector<int32_t>& most = m_scratch_most;
    std::vector<Object*>& by_slot = m_scratch_by_slot;
    f.resize(14 * n);
    most.resize(n);
    by_slot.resize(n);
    essa_bodies b;
    std::memset(&b, 0, sizeof(b));
    double* p = f.data();
    b.x = p; b.y = p + n; b.z = p + 2 * n; b.vx = p + 3 * n; b.vy = p + 4 * n; b.vz = p + 5 * n;
    b.ax = p + 6 * n; b.ay = p + 7 * n; b.az = p + 8 * n;
    b.max_attraction = p + 9 * n;
    b.ap = p + 10 * n; b.pe = p + 11 * n; b.ap_vel = p + 12 * n; b.pe_vel = p + 13 * n;
    b.most = most.data();
    check(essa_download(m_ctx, &b), "essa_download");
    size_t k = 0;
    for (auto& o : m_object_list)
        by_slot[k++] = o.get();
    k = 0;
    for (auto& o : m_object_list) {
        o->m_pos = { b.x[k], b.y[k], b.z[k] };
        o->m_vel = { b.vx[k], b.vy[k], b.vz[k] };
        o->m_attraction_factor = { b.ax[k], b.ay[k], b.az[k] };
        o->m_max_attraction = b.max_attraction[k];
        o->m_ap = b.ap[k]; o->m_pe = b.pe[k]; o->m_ap_vel = b.ap_vel[k]; o->m_pe_vel = b.pe_vel[k];
        o->m_most_attracting_object = most[k] >= 0 ? by_slot[most[k]] : nullptr;
        k++;
    }
}

void World::debug_stash_last_object() {
    if (m_object_list.empty())
        return;
    m_object_history.push_to_entry(std::move(m_object_list.back())); // reset_history() brings its rings home first
    m_object_list.pop_back();
    m_list_dirty = true;
}

void World::update_history_and_date(bool reverse) {
    if (m_is_forward_simulated)
        return;
    m_object_history.set_time(m_date);
    if (!reverse) {
        m_date += m_simulation_seconds_per_tick;
        if (m_object_history.size() > 0 && m_object_list.back()->creation_date() <= m_date) {
            m_object_list.push_back(m_object_history.pop_from_entries());
            m_list_dirty = true;
        }
    } else {
        m_date -= m_simulation_seconds_per_tick;
        if (m_object_history.size() > 0 && m_object_list.back()->creation_date() > m_date) {
            m_object_history.push_to_entry(std::move(m_object_list.back()));
            m_object_list.pop_back();
            m_list_dirty = true;
        }
    }
}

void World::update(int steps) {
    assert(steps != 0);
    if (steps == 0)
        return;
    if (m_object_list.empty()) { // nothing to integrate: only the clock moves (src/World.cpp:59-84)
        if (!m_is_forward_simulated)
            m_date += static_cast<SimulationTime>(steps) * m_simulation_seconds_per_tick;
        return;
    }
    if (m_object_history.size() > 0) {
        // Objects may enter / leave the list by date: go one step at a time so that the list the
        // device sees is the list of that step (unreachable in the reference's current revision).
        bool reverse = steps < 0;
        for (int i = 0; i < std::abs(steps); i++) {
            SimulationTime before = m_date;
            update_history_and_date(reverse);
            m_date = before; // essa_step moves the date itself
            sync_to_device();
            essa_step_opts o = step_opts();
            check(essa_step(m_ctx, reverse ? -1 : 1, &o), "essa_step");
            m_date = essa_get_date(m_ctx);
            sync_from_device();
        }
        return;
    }
    sync_to_device();
    essa_step_opts o = step_opts();
    check(essa_step_async(m_ctx, steps, &o), "essa_step"); // the read-back below is the synchronisation point
    m_date = essa_get_date(m_ctx);
    sync_from_device();
}

void World::set_forces() {
    if (m_object_list.empty())
        return;
    sync_to_device();
    essa_step_opts o = step_opts();
    o.record = 0;
    check(essa_set_forces(m_ctx, &o), "essa_set_forces");
    sync_from_device();
}

} // namespace essa
