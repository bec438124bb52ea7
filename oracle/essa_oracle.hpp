// ============================================================================================
// essa_oracle.hpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch CPU restatement of the reference's World::update hot path
// (essa-software/essa, files under /root/reference/src).  It exists so that the CUDA product in
// essa_b200/ can be checked against the reference's algorithm on the same inputs, and so that the
// reference's CPU cost can be timed on the GPU box's host cores (bench.py `cpu_baseline`).
//
//   * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
//     build, link, load or call anything in oracle/.  The product never does.
//   * The reference itself cannot be compiled here: its CMake fetches EssaGUI/EssaUtil from GitHub
//     (CMakeLists.txt:13-19) and needs SDL2/GLEW/fmt; none is present and there is no network.
//     All vector arithmetic therefore lives in an absent third-party dependency
//     (EssaGUI @ 6bc4531055f7732b20748ba53b5bb7c6056e20da).  What is restated here follows the
//     reference's own call sites; the three facts that remain [unverified] are marked below:
//       (U1) vector / scalar is three IEEE divisions (not a reciprocal multiply),
//       (U2) Util::Constants::AU = 149,597,870,700 m,
//       (U3) Point3f / double is evaluated in double and narrowed to float.
//   * PARITY PIN: the force law, G = 6.67408e-11 and gf = m*G are pinned by the reference's own
//     recorded traces tests/accuracy.ods (replayed by tests/test_oracle_golden.py with the
//     `legacy_euler` switch below: 0 mismatching fields of 43,206 + 438).  The current KDK
//     ordering, History, Trail and most-attracting logic are NOT pinned by any reference test
//     ("parity unpinned" for those parts; DESIGN.md says the same).
//
// Compile with:  g++ -std=c++20 -O2 -ffp-contract=off   (no FMA contraction: results must not
// depend on the host ISA; the reference's x86-64 baseline build has no FMA either).
//
// Every function cites the reference file:line it follows.
// ============================================================================================
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace oracle {

// Util::Constants::Gravity -- pinned to 6 s.f. by tests/accuracy.ods.
inline constexpr double kGravity = 6.67408e-11;
// Util::Constants::AU -- (U2) unverified, IAU 2012 value assumed.
inline constexpr double kAU = 149597870700.0;
// Util::SimulationTime::create(1990, 4, 20) as seen in the traces (first row is t0 + dt).
inline constexpr int64_t kStartDate = 640566000;

// --------------------------------------------------------------------------------------------
// Minimal stand-in for Util::DeprecatedVector3d: component-wise binary64, evaluated left to right.
// --------------------------------------------------------------------------------------------
struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3 operator+(Vec3 const& o) const { return { x + o.x, y + o.y, z + o.z }; }
    Vec3 operator-(Vec3 const& o) const { return { x - o.x, y - o.y, z - o.z }; }
    Vec3 operator-() const { return { -x, -y, -z }; }
    Vec3 operator*(double s) const { return { x * s, y * s, z * s }; }
    Vec3 operator/(double s) const { return { x / s, y / s, z / s }; } // (U1)
    Vec3& operator+=(Vec3 const& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3& operator-=(Vec3 const& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    bool operator==(Vec3 const& o) const { return x == o.x && y == o.y && z == o.z; }
    double length_squared() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(length_squared()); }
    Vec3 rotate_z(double a) const {
        double c = std::cos(a), s = std::sin(a);
        return { x * c - y * s, x * s + y * c, z };
    }
};
inline double get_distance(Vec3 const& a, Vec3 const& b) { return (a - b).length(); }

struct Vec3f {
    float x = 0, y = 0, z = 0;
};

// --------------------------------------------------------------------------------------------
// Trail -- src/Trail.hpp:8-44, src/Trail.cpp:19-110 (draw() is OpenGL and out of scope).
// --------------------------------------------------------------------------------------------
class Trail {
public:
    static constexpr size_t kMaxAllocated = 0xfffff; // Trail.cpp:17

    // Trail.cpp:19-34
    explicit Trail(size_t max_trail_size) {
        if (max_trail_size < 3)
            max_trail_size = 3;
        if (max_trail_size > kMaxAllocated)
            max_trail_size = kMaxAllocated;
        m_vertexes.resize(max_trail_size);
        m_length++;
    }

    // Trail.cpp:49-75
    void push_back(Vec3 const& pos) {
        if (m_length == 1) {
            assert(m_append_offset == 1);
            m_vertexes[m_append_offset] = to_vertex(pos);
            m_append_offset++;
            m_length++;
        }
        if (m_length >= 3 && m_append_offset != 0) {
            if (std::fabs(std::atan2(pos.y, pos.x) - m_last_xy_angle) <= 2 * M_PI / m_vertexes.size()) {
                change_current(pos);
                return;
            }
        }
        m_vertexes[m_append_offset] = to_vertex(pos);
        m_append_offset++;
        if (static_cast<size_t>(m_length) < m_vertexes.size())
            m_length++;
        if (static_cast<size_t>(m_append_offset) == m_vertexes.size()) {
            m_vertexes[0] = to_vertex(pos);
            m_append_offset = 1;
        }
        m_last_xy_angle = std::atan2(pos.y, pos.x);
    }

    // Trail.cpp:77-80
    void reset() {
        m_append_offset = 1;
        m_length = 1;
    }

    void set_offset(Vec3 const& o) { m_offset = o; } // Trail.hpp:38

    // Trail.cpp:82-88.  Vertex arithmetic is binary32: v + float(off_old)/AU - float(off_new)/AU.
    void recalculate_with_offset(Vec3 const& offset) {
        if (m_offset == offset)
            return;
        Vec3f a = to_vertex(m_offset), b = to_vertex(offset);
        for (int s = 0; s < m_length; s++) {
            auto& v = m_vertexes[s];
            v.x = v.x + a.x - b.x;
            v.y = v.y + a.y - b.y;
            v.z = v.z + a.z - b.z;
        }
        m_offset = offset;
    }

    // Trail.cpp:105-110
    void change_current(Vec3 const& pos) {
        assert(m_length > 0);
        if (m_append_offset == 1)
            m_vertexes[m_length - 1] = to_vertex(pos);
        m_vertexes[m_append_offset - 1] = to_vertex(pos);
    }

    void set_enable_min_step(bool b) { m_enable_min_step = b; } // set, never read (Trail.hpp:24,40)

    std::vector<Vec3f> const& vertexes() const { return m_vertexes; }
    int append_offset() const { return m_append_offset; }
    int length() const { return m_length; }
    Vec3 offset() const { return m_offset; }
    double last_xy_angle() const { return m_last_xy_angle; }

    // pos.cast<float>() / AU  (Trail.cpp:53,64): cast first, then divide -- (U3) in double, narrowed.
    static Vec3f to_vertex(Vec3 const& p) {
        return { static_cast<float>(static_cast<double>(static_cast<float>(p.x)) / kAU),
            static_cast<float>(static_cast<double>(static_cast<float>(p.y)) / kAU),
            static_cast<float>(static_cast<double>(static_cast<float>(p.z)) / kAU) };
    }

private:
    std::vector<Vec3f> m_vertexes;
    int m_append_offset = 1;
    int m_length = 0;
    Vec3 m_offset;
    bool m_enable_min_step = true;
    double m_last_xy_angle = 0;
};

// --------------------------------------------------------------------------------------------
// History -- src/History.hpp:8-34, src/History.cpp:6-59.  std::list on purpose: the malloc/free
// per body per step is part of the reference's CPU cost that bench.py times.
// --------------------------------------------------------------------------------------------
class History {
public:
    struct Entry {
        Vec3 pos, vel;
    };

    History(unsigned segments, Entry first) // History.cpp:6-9
        : m_segments(segments)
        , m_first_entry(first) {
        m_entry_list.push_back(first);
        m_side = false;
    }

    std::optional<Entry> move_forward(Entry const entry) { // History.cpp:11-31
        if (!m_side) {
            m_entry_list.push_back(entry);
            if (m_entry_list.size() > m_segments)
                m_entry_list.pop_front();
            return {};
        }
        if (!m_entry_list.empty()) {
            auto value = m_entry_list.front();
            m_entry_list.pop_front();
            return value;
        }
        m_side = false;
        m_entry_list.push_back(entry);
        return {};
    }

    std::optional<Entry> move_backward(Entry const entry) { // History.cpp:33-53
        if (m_side) {
            m_entry_list.push_front(entry);
            if (m_entry_list.size() > m_segments)
                m_entry_list.pop_back();
            return {};
        }
        if (!m_entry_list.empty()) {
            auto value = m_entry_list.back();
            m_entry_list.pop_back();
            return value;
        }
        m_side = true;
        m_entry_list.push_front(entry);
        return {};
    }

    void reset() { // History.cpp:55-59
        m_entry_list.clear();
        m_entry_list.push_back(m_first_entry);
        m_side = false;
    }

    std::list<Entry> const& entries() const { return m_entry_list; }
    bool side() const { return m_side; }

private:
    std::list<Entry> m_entry_list;
    bool m_side;
    unsigned const m_segments;
    Entry const m_first_entry;
};

class World;

// --------------------------------------------------------------------------------------------
// Object -- src/Object.hpp:21-191, src/Object.cpp (physics + recording parts only).
// --------------------------------------------------------------------------------------------
class Object {
public:
    // Object.cpp:32-43
    Object(double mass, double radius, Vec3 pos, Vec3 vel, std::string name, unsigned period)
        : m_trail(std::max(2U, std::max(period * 2, 500U)))
        , m_history(1000, { pos, vel })
        , m_pos(pos)
        , m_vel(vel)
        , m_orbit_len(period)
        , m_gravity_factor(mass * kGravity)
        , m_radius(radius)
        , m_name(std::move(name)) {
        m_trail.push_back(m_pos);
    }
    Object(Object const&) = delete;
    Object& operator=(Object const&) = delete;

    bool deleted() const; // Object.cpp:60-62

    // Object.cpp:64-95 -- operation order kept exactly.
    void update_forces_against(Object& o) {
        Vec3 dist = m_pos - o.m_pos;
        double denominator = dist.length_squared();
        denominator *= std::sqrt(denominator);
        Vec3 base = dist / denominator;

        Vec3 this_attraction = base * o.m_gravity_factor;
        m_attraction_factor -= this_attraction;
        Vec3 other_attraction = base * m_gravity_factor;
        o.m_attraction_factor += other_attraction;

        double this_mag = this_attraction.length_squared() / o.m_gravity_factor;
        if (this_mag > m_max_attraction && o.m_gravity_factor > m_gravity_factor) {
            m_max_attraction = this_mag;
            m_most_attracting_object = &o;
        }
        double other_mag = other_attraction.length_squared() / m_gravity_factor;
        if (other_mag > o.m_max_attraction && o.m_gravity_factor < m_gravity_factor) {
            o.m_max_attraction = other_mag;
            o.m_most_attracting_object = this;
        }
    }

    void before_update() { // Object.cpp:97-101
        if (m_is_forward_simulated)
            update_closest_approaches();
        m_old_most_attracting_object = m_most_attracting_object;
    }

    void clear_forces() { // Object.cpp:126-129 -- most_attracting is NOT cleared
        m_attraction_factor = Vec3 {};
        m_max_attraction = 0;
    }

    void update(int speed) { // Object.cpp:131-153
        if (speed > 0) {
            if (!m_is_forward_simulated) {
                auto val = m_history.move_forward({ m_pos, m_vel });
                if (val.has_value()) {
                    m_pos = val->pos;
                    m_vel = val->vel;
                }
            }
        } else if (speed < 0) {
            assert(!m_is_forward_simulated);
            auto val = m_history.move_backward({ m_pos, m_vel });
            if (val.has_value()) {
                m_pos = val->pos;
                m_vel = val->vel;
            }
        }
        nonphysical_update();
    }

    void delete_object(); // Object.cpp:243-246

    void delete_most_attracting_object() { // Object.cpp:169-173
        m_most_attracting_object = nullptr;
        m_trail.recalculate_with_offset({});
        m_trail.reset();
    }

    void reset_history() { // Object.cpp:268-271
        m_history.reset();
        m_trail.reset();
    }

    // Object.cpp:309-344.  `apoapsis`/`periapsis` arrive as Distance::value() (a float).
    std::unique_ptr<Object> create_object_relative_to_ap_pe(double mass, float radius, float apoapsis, float periapsis,
        bool direction, double theta, double alpha, std::string name, double rotation) const {
        double GM = m_gravity_factor;
        double a = (apoapsis + periapsis) / 2; // float + float, as in the reference expression
        double b = std::sqrt(apoapsis * periapsis);
        double T = 2 * M_PI * std::sqrt((a * a * a) / GM);
        Vec3 pos { std::cos(theta) * std::cos(alpha) * a, std::sin(theta) * std::cos(alpha) * a, std::sin(alpha) * a };
        pos = pos.rotate_z(rotation);
        pos += m_pos;

        auto result = std::make_unique<Object>(mass, radius, pos, Vec3 {}, std::move(name), static_cast<unsigned>(T / (3600 * 24)));
        result->m_ap = apoapsis;
        result->m_pe = periapsis;
        result->m_eccentrity = std::sqrt(1 - (b * b) / (a * a));

        double velocity_constant = (2 * GM) / (apoapsis + periapsis);
        result->m_ap_vel = std::sqrt(2 * GM / apoapsis - velocity_constant);
        result->m_pe_vel = std::sqrt(2 * GM / periapsis - velocity_constant);
        double velocity = std::sqrt(2 * GM / get_distance(pos, m_pos) - velocity_constant);

        Vec3 vel { std::cos(theta + M_PI / 2) * velocity, std::sin(theta + M_PI / 2) * velocity, 0 };
        if (direction)
            vel = -vel;
        vel += m_vel;
        result->m_vel = vel;
        return result;
    }

    // Object.cpp:350-383
    std::unique_ptr<Object> create_object_relative_to_maj_ecc(double mass, float radius, float semi_major, double ecc,
        bool direction, double theta, double alpha, std::string name, double rotation) const {
        double GM = m_gravity_factor;
        double a = semi_major;
        double T = 2 * M_PI * std::sqrt((a * a * a) / GM);
        Vec3 pos { std::cos(theta) * std::cos(alpha) * a, std::sin(theta) * std::cos(alpha) * a, std::sin(alpha) * a };
        pos = pos.rotate_z(rotation);
        pos += m_pos;

        auto result = std::make_unique<Object>(mass, radius, pos, Vec3 {}, std::move(name), static_cast<unsigned>(T / (3600 * 24)));
        result->m_ap = 0;
        result->m_pe = std::numeric_limits<double>::max();
        result->m_eccentrity = ecc;

        double velocity_constant = (2 * GM) / (a * 2 - a * ecc);
        result->m_ap_vel = std::sqrt(2 * GM / a - velocity_constant);
        result->m_pe_vel = std::sqrt(2 * GM / (a - a * ecc) - velocity_constant);
        double velocity = std::sqrt(2 * GM / get_distance(pos, m_pos) - velocity_constant);

        Vec3 vel { std::cos(theta + M_PI / 2) * velocity, std::sin(theta + M_PI / 2) * velocity, 0 };
        if (direction)
            vel = -vel;
        vel += m_vel;
        result->m_vel = vel;
        return result;
    }

    // Object.cpp:385-399 (colour handling dropped: drawing only)
    std::unique_ptr<Object> clone_for_forward_simulation() const {
        auto object = std::make_unique<Object>(0, m_radius, m_pos, m_vel, m_name, 500);
        object->m_is_forward_simulated = true;
        object->m_gravity_factor = m_gravity_factor;
        object->m_trail.set_enable_min_step(false);
        return object;
    }

    struct ClosestApproachEntry { // Object.hpp:184-188
        Vec3 this_position;
        Vec3 other_object_position;
        double distance {};
    };

    // --- state (Object.hpp:152-191), public: this is test infrastructure ---
    Trail m_trail;
    History m_history;
    Vec3 m_pos, m_vel, m_attraction_factor;
    bool m_deleted = false;
    bool m_is_forward_simulated = false;
    World* m_world = nullptr;
    double m_ap = 0, m_pe = std::numeric_limits<double>::max();
    double m_ap_vel = 0, m_pe_vel = 0;
    double m_orbit_len = 0, m_eccentrity = 0;
    double m_gravity_factor = 0;
    double m_radius = 0;
    int64_t m_creation_date = 0, m_deletion_date = 0;
    std::string m_name;
    double m_max_attraction = 0;
    Object* m_old_most_attracting_object = nullptr;
    Object* m_most_attracting_object = nullptr;
    std::map<Object*, ClosestApproachEntry> m_closest_approaches;

private:
    void update_closest_approaches(); // Object.cpp:405-417
    void nonphysical_update();        // Object.cpp:103-124
    void recalculate_trails_with_offset() { // Object.cpp:155-167
        if (!m_most_attracting_object || m_is_forward_simulated) {
            m_trail.recalculate_with_offset({});
            m_trail.push_back(m_pos);
            return;
        }
        if (m_most_attracting_object != m_old_most_attracting_object)
            m_trail.recalculate_with_offset(m_most_attracting_object->m_pos);
        else
            m_trail.set_offset(m_most_attracting_object->m_pos);
        m_trail.push_back(m_pos - m_most_attracting_object->m_pos);
    }
};

// --------------------------------------------------------------------------------------------
// ObjectHistory -- src/ObjectHistory.hpp:6-24, src/ObjectHistory.cpp:4-29 (inert in this revision,
// SURVEY.md section 8 a2; restated for fidelity).
// --------------------------------------------------------------------------------------------
class ObjectHistory {
public:
    unsigned get_pos() const { return m_pos; }
    void clear_history(unsigned long to_index) {
        m_entries.resize(std::min<size_t>(to_index, m_entries.size()));
        m_pos = static_cast<unsigned>(to_index - 1);
    }
    bool set_time(int64_t time) {
        m_pos = 0;
        for (auto& e : m_entries) {
            if (e->m_creation_date > time)
                return true;
            m_pos++;
        }
        return false;
    }
    void push_to_entry(std::unique_ptr<Object> obj) {
        obj->reset_history();
        m_entries.push_back(std::move(obj));
    }
    std::unique_ptr<Object> pop_from_entries() {
        auto object = std::move(m_entries.back());
        m_entries.pop_back();
        return object;
    }
    unsigned size() const { return static_cast<unsigned>(m_entries.size()); }

private:
    std::vector<std::unique_ptr<Object>> m_entries;
    unsigned m_pos = 0;
};

// --------------------------------------------------------------------------------------------
// World -- src/World.hpp:19-95, src/World.cpp:22-139,230-238.
// --------------------------------------------------------------------------------------------
class World {
public:
    World()
        : m_date(kStartDate) { m_start_date = m_date; } // World.cpp:22-27

    void add_object(std::unique_ptr<Object> object) { // World.cpp:29-40
        object->m_world = this;
        object->m_creation_date = m_date;
        m_object_list.push_back(std::move(object));
        if (m_object_history.set_time(m_date))
            m_object_history.clear_history(m_object_history.get_pos());
        m_object_list.remove_if([&](std::unique_ptr<Object>& obj) { return obj->m_creation_date > m_date; });
    }

    void set_forces() { // World.cpp:42-57
        for (auto& p : m_object_list) {
            if (!p->deleted())
                p->clear_forces();
        }
        for (auto it = m_object_list.begin(); it != m_object_list.end(); it++) {
            auto it2 = it;
            auto& this_object = **it;
            it2++;
            for (; it2 != m_object_list.end(); it2++) {
                if (!this_object.deleted() && !(*it2)->deleted())
                    this_object.update_forces_against(**it2);
            }
        }
    }

    void update(int steps) { // World.cpp:86-139
        assert(steps != 0);
        bool reverse = steps < 0;
        for (unsigned i = 0; i < static_cast<unsigned>(std::abs(steps)); i++) {
            update_history_and_date(reverse);
            for (auto& obj : m_object_list)
                obj->before_update();

            double step = m_simulation_seconds_per_tick;
            double halfStep = (step / 2.0);
            double mul = (steps >= 0) ? 1 : -1;

            if (legacy_euler) {
                // NOT in the current reference: the kick(dt)->drift(dt) integrator that produced
                // tests/accuracy.ods.  Used only to replay those traces.
                set_forces();
                for (auto& obj : m_object_list) {
                    if (obj->deleted())
                        continue;
                    obj->m_vel = obj->m_vel + obj->m_attraction_factor * step * mul;
                    obj->m_pos = obj->m_pos + obj->m_vel * step * mul;
                }
                continue;
            }

            set_forces();
            for (auto& obj : m_object_list) {
                if (obj->deleted())
                    continue;
                obj->m_vel = obj->m_vel + obj->m_attraction_factor * halfStep * mul;
                obj->m_pos = obj->m_pos + obj->m_vel * step * mul;
            }
            set_forces();
            for (auto& obj : m_object_list) {
                if (obj->deleted())
                    continue;
                obj->m_vel = obj->m_vel + obj->m_attraction_factor * halfStep * mul;
                obj->update(m_simulation_seconds_per_tick);
            }
        }
    }

    Object* get_object_by_name(std::string const& name) { // World.cpp:163-169
        for (auto& obj : m_object_list) {
            if (obj->m_name == name)
                return obj.get();
        }
        return nullptr;
    }

    void clone_for_forward_simulation(World& new_world) const { // World.cpp:230-238
        new_world.m_object_list.clear();
        new_world.m_date = kStartDate;
        new_world.m_start_date = kStartDate;
        new_world.m_is_forward_simulated = true;
        new_world.offset_trails = offset_trails;
        new_world.m_simulation_seconds_per_tick = 60 * 60 * 12; // `new_world = World()` resets it
        for (auto& object : m_object_list) {
            if (!object->deleted())
                new_world.add_object(object->clone_for_forward_simulation());
        }
    }

    Object* object_at(size_t idx) {
        auto it = m_object_list.begin();
        std::advance(it, idx);
        return it->get();
    }
    int index_of(Object const* o) const {
        int k = 0;
        for (auto& p : m_object_list) {
            if (p.get() == o)
                return k;
            k++;
        }
        return -1;
    }

    // --- state (World.hpp:73-80) ---
    int64_t m_start_date;
    int64_t m_date;
    ObjectHistory m_object_history;
    std::list<std::unique_ptr<Object>> m_object_list;
    int m_simulation_seconds_per_tick = 60 * 60 * 12;
    bool m_is_forward_simulated = false;
    // SimulationView::offset_trails() -- a headless world has no view; default true
    // (SimulationView.hpp:182).
    bool offset_trails = true;
    bool legacy_euler = false;

private:
    void update_history_and_date(bool reverse) { // World.cpp:59-84
        if (m_is_forward_simulated)
            return;
        m_object_history.set_time(m_date);
        if (!reverse) {
            m_date += m_simulation_seconds_per_tick;
            if (m_object_history.size() > 0) {
                if (m_object_list.back()->m_creation_date <= m_date)
                    m_object_list.push_back(m_object_history.pop_from_entries());
            }
        } else {
            m_date -= m_simulation_seconds_per_tick;
            if (m_object_history.size() > 0) {
                if (m_object_list.back()->m_creation_date > m_date) {
                    m_object_history.push_to_entry(std::move(m_object_list.back()));
                    m_object_list.pop_back();
                }
            }
        }
    }
};

inline bool Object::deleted() const {
    return (m_deleted && m_deletion_date <= m_world->m_date) || m_creation_date > m_world->m_date;
}

inline void Object::delete_object() {
    m_deletion_date = m_world->m_date;
    m_deleted = true;
}

inline void Object::update_closest_approaches() {
    for (auto& other : m_world->m_object_list) {
        Object& object = *other;
        if (&object == this)
            continue;
        auto& e = m_closest_approaches[&object];
        auto distance = get_distance(object.m_pos, m_pos);
        if (e.distance == 0 || distance < e.distance) {
            e.distance = distance;
            e.this_position = m_pos;
            e.other_object_position = object.m_pos;
        }
    }
}

inline void Object::nonphysical_update() {
    if (m_world->offset_trails)
        recalculate_trails_with_offset();
    else {
        m_trail.recalculate_with_offset({});
        m_trail.push_back(m_pos);
    }
    if (m_most_attracting_object == nullptr)
        return;
    double d = get_distance(m_pos, m_most_attracting_object->m_pos);
    if (m_ap < d) {
        m_ap = d;
        m_ap_vel = m_vel.length();
    }
    if (m_pe > d) {
        m_pe = d;
        m_pe_vel = m_vel.length();
    }
}

// --------------------------------------------------------------------------------------------
// .essa loader -- src/ConfigLoader.cpp:17-30,61-78,110-202,204-270.  Returns "" or an error text.
// `angle_as_float`: whether Util::Angle stores binary32 (unknown; default double).
// --------------------------------------------------------------------------------------------
std::string load_essa(std::string const& text, World& world, bool angle_as_float = false);

} // namespace oracle
