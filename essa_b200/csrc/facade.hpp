// Host-side mirror of the reference's World / Object interface for the World::update path.
//
// Same names, argument meaning and error behaviour as essa-software/essa's src/World.hpp:19-95 and
// src/Object.hpp:21-191 (physics + recording surface; drawing, GUI and the embedded-Python glue
// are out of scope), but World::update / World::set_forces do not run any loop on the CPU: they
// synchronise the object list with an essa_ctx and call the C ABI of include/essa_b200.h.
// There is no CPU fallback: the first update() on a machine without a B200 throws.
//
// State lives in two places:
//   * host mirrors inside Object (pos, vel, acc, gf, dates, most-attracting link, ap/pe) -- what the
//     accessors return; refreshed from the device after every update()/set_forces();
//   * device-resident rings (History, Trail) -- fetched on demand by Object::trail()/history().
// Host-side mutation (set_pos, set_vel, add_object, delete_object ...) marks the world dirty; the
// next update() re-uploads before stepping.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <list>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/essa_b200.h"

namespace essa {

// Stand-in for Util::DeprecatedVector3d (EssaUtil, not vendored by the reference): component-wise
// binary64, evaluated left to right.
struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3 operator+(Vec3 const& o) const { return { x + o.x, y + o.y, z + o.z }; }
    Vec3 operator-(Vec3 const& o) const { return { x - o.x, y - o.y, z - o.z }; }
    Vec3 operator-() const { return { -x, -y, -z }; }
    Vec3 operator*(double s) const { return { x * s, y * s, z * s }; }
    Vec3 operator/(double s) const { return { x / s, y / s, z / s }; }
    Vec3& operator+=(Vec3 const& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3& operator-=(Vec3 const& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    bool operator==(Vec3 const& o) const { return x == o.x && y == o.y && z == o.z; }
    double length_squared() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(length_squared()); }
    Vec3 normalized() const { double l = length(); return { x / l, y / l, z / l }; }
    Vec3 rotate_z(double a) const {
        double c = std::cos(a), s = std::sin(a);
        return { x * c - y * s, x * s + y * c, z };
    }
};
inline double get_distance(Vec3 const& a, Vec3 const& b) { return (a - b).length(); }

struct Color {
    uint8_t r = 255, g = 255, b = 255, a = 255;
};

using SimulationTime = int64_t; // seconds; Util::SimulationClock::time_point in the reference

// Snapshot of Object::m_trail (src/Trail.hpp:8-44) as it stands on the device.
struct Trail {
    std::vector<float> vertexes; // 3 floats per vertex, already divided by AU
    int append_offset = 1;
    int length = 0;
    Vec3 offset;
    double last_xy_angle = 0;
    size_t size() const { return vertexes.size() / 3; }

    explicit Trail(size_t max_trail_size = 3);
    void push_back(Vec3 const& pos);               // src/Trail.cpp:49-75 (host copy: constructor push only)
    void reset() { append_offset = 1; length = 1; } // src/Trail.cpp:77-80
    void recalculate_with_offset(Vec3 const& o);   // src/Trail.cpp:82-88
};

struct HistoryEntry {
    Vec3 pos, vel;
};

class World;

class Object {
public:
    // src/Object.cpp:32-43
    Object(double mass, double radius, Vec3 pos, Vec3 vel, Color color, std::string name, unsigned period);
    Object(Object const&) = delete;
    Object& operator=(Object const&) = delete;

    std::string name() const { return m_name; }
    void set_name(std::string n) { m_name = std::move(n); }

    double mass() const { return m_gravity_factor / ESSA_GRAVITY; }
    double gravity_factor() const { return m_gravity_factor; }
    void set_gravity_factor(double gf);

    Vec3 pos() const { return m_pos; }
    void set_pos(Vec3 const& pos);
    Vec3 vel() const { return m_vel; }
    void set_vel(Vec3 const& vel);
    Vec3 acc() const { return m_attraction_factor; }
    Vec3 attraction(Object const& other) const; // src/Object.cpp:53-58 (kept for Python only)

    Color color() const { return m_color; }
    void set_color(Color c) { m_color = c; }
    double radius() const { return m_radius; }
    void set_radius(double r) { m_radius = r; }

    SimulationTime creation_date() const { return m_creation_date; }
    SimulationTime deletion_date() const { return m_deletion_date; }
    bool deleted() const;  // src/Object.cpp:60-62
    void delete_object();  // src/Object.cpp:243-246

    Object* most_attracting_object() const { return m_most_attracting_object; }
    void delete_most_attracting_object(); // src/Object.cpp:169-173

    struct Info { // src/Object.hpp:102-116
        double mass, radius, absolute_velocity;
        double distance_from_most_massive_object = 0;
        double apoapsis = 0, apoapsis_velocity = 0, periapsis = 0, periapsis_velocity = 0;
        double orbit_period = 0, orbit_eccentrity = 0;
    };
    Info get_info() const; // src/Object.cpp:248-266

    void reset_history(); // src/Object.cpp:268-271

    // Device-resident recording, fetched on access.
    Trail const& trail();
    std::vector<HistoryEntry> const& history();

    // src/Object.cpp:309-344 / :350-383.  Distances arrive as float (Distance::value()), angles in radians.
    std::unique_ptr<Object> create_object_relative_to_ap_pe(double mass, float radius, float apoapsis, float periapsis, bool direction,
        double theta, double alpha, Color color, std::string name, double rotation) const;
    std::unique_ptr<Object> create_object_relative_to_maj_ecc(double mass, float radius, float semi_major, double ecc, bool direction,
        double theta, double alpha, Color color, std::string name, double rotation) const;
    std::unique_ptr<Object> clone_for_forward_simulation() const; // src/Object.cpp:385-399

    struct ClosestApproachEntry { // src/Object.hpp:184-188
        Vec3 this_position, other_object_position;
        double distance {};
    };
    // Forward-simulated worlds only (src/Object.cpp:405-417); nullopt when never evaluated.
    std::optional<ClosestApproachEntry> closest_approach(Object const& other);

    double apoapsis() const { return m_ap; }
    double periapsis() const { return m_pe; }
    double apoapsis_velocity() const { return m_ap_vel; }
    double periapsis_velocity() const { return m_pe_vel; }
    double max_attraction() const { return m_max_attraction; }
    double orbit_period() const { return m_orbit_len; }
    double eccentrity() const { return m_eccentrity; }
    unsigned trail_capacity() const { return m_trail_capacity; }
    bool is_forward_simulated() const { return m_is_forward_simulated; }

private:
    friend class World;
    void touch();

    Trail m_trail;                       // host copy (constructor state until first upload, then a cache)
    std::vector<HistoryEntry> m_history; // host copy, oldest first
    HistoryEntry m_first_entry;          // History::m_first_entry
    bool m_rings_on_device = false;
    int m_slot = -1;

    Vec3 m_pos, m_vel, m_attraction_factor;
    bool m_deleted = false;
    bool m_is_forward_simulated = false;
    World* m_world = nullptr;
    double m_ap = 0, m_pe = std::numeric_limits<double>::max();
    double m_ap_vel = 0, m_pe_vel = 0;
    double m_orbit_len = 0, m_eccentrity = 0;
    double m_gravity_factor = 0;
    double m_radius = 0;
    SimulationTime m_creation_date = 0, m_deletion_date = 0;
    std::string m_name;
    Color m_color;
    unsigned m_trail_capacity = 0;
    double m_max_attraction = 0;
    Object* m_most_attracting_object = nullptr;
};

// src/ObjectHistory.hpp:6-24 -- timeline of not-yet-created objects.  Inert in the reference's
// current revision (never populated); kept on the host for interface fidelity.
class ObjectHistory {
public:
    unsigned get_pos() const { return m_pos; }
    void clear_history(unsigned long to_index);
    bool set_time(SimulationTime time);
    void push_to_entry(std::unique_ptr<Object> obj);
    std::unique_ptr<Object> pop_from_entries();
    unsigned size() const { return static_cast<unsigned>(m_entries.size()); }

private:
    std::vector<std::unique_ptr<Object>> m_entries;
    unsigned m_pos = 0;
};

class World {
public:
    explicit World(int device = 0);
    ~World();
    World(World const&) = delete;
    World& operator=(World const&) = delete;

    void update(int steps);                        // src/World.cpp:86-139
    void add_object(std::unique_ptr<Object>);      // src/World.cpp:29-40
    void reset(std::optional<std::string> const& filename); // src/World.cpp:171-202; throws on loader errors
    Object* get_object_by_name(std::string const& name);
    SimulationTime date() const { return m_date; }
    void set_date(SimulationTime d);

    template<class C>
    void for_each_object(C callback) {
        for (auto& it : m_object_list)
            callback(*it);
    }

    void clone_for_forward_simulation(World& new_world) const; // src/World.cpp:230-238
    int simulation_seconds_per_tick() const { return m_simulation_seconds_per_tick; }
    void set_simulation_seconds_per_tick(int s) { m_simulation_seconds_per_tick = s; }
    void reset_all_trails();
    void delete_object_by_ptr(Object* ptr);        // src/World.cpp:208-219
    std::unique_ptr<Object>& last_object() { return m_object_list.back(); }
    void set_forces();                             // src/World.cpp:42-57
    bool exist_object_with_name(std::string const& name) const;
    bool set_light_source(std::string const& name);
    void set_light_source(Object* obj) { m_light_source = obj; }
    Object* light_source() const { return m_light_source; }
    std::function<void()> on_reset;

    size_t object_count() const { return m_object_list.size(); }
    Object* object_at(size_t idx);
    int index_of(Object const* o) const;
    bool is_forward_simulated() const { return m_is_forward_simulated; }

    // What the reference reads from m_simulation_view->offset_trails() (SimulationView.hpp:182): a
    // headless world has no view, so it is a member here.  Default true as in the GUI.
    bool offset_trails = true;
    // Arithmetic of the resident kernel: true = the reference's operation order, bit for bit.
    bool exact_order = false;
    // Evaluate both set_forces() calls of every step (the reference does; results are identical).
    bool two_pass = false;

    // Seeds the ObjectHistory with the list's last object.  The reference's current revision never does (src/World.cpp:76-81
    // pushes only when the history already holds an entry), so the date-driven push / pop of World.cpp:59-84 can only be
    // exercised from outside: this is the hook the tests use.
    void debug_stash_last_object();
    unsigned object_history_size() const { return m_object_history.size(); }

    essa_ctx* context(); // creates the context on first use; throws std::runtime_error without a GPU
    double last_step_ms() const;
    // essa_persist_configure / essa_persist_stats of the world's context: whether the kernel that served one update()
    // stays resident for the next (the GUI's cadence: one update(iterations) + read-back per tick); default on
    void set_persistent(bool enable, int idle_timeout_us = 0);
    void persist_stats(int64_t out[6]);

private:
    friend class Object;
    void mark_state_dirty() { m_state_dirty = true; }
    void mark_list_dirty() { m_list_dirty = true; }
    void sync_to_device();
    void sync_from_device();
    void pull_rings(Object& o);
    void update_history_and_date(bool reverse); // src/World.cpp:59-84
    essa_step_opts step_opts() const;
    void check(int rc, char const* what);

    int m_device;
    essa_ctx* m_ctx = nullptr;
    std::vector<double> m_scratch_f;
    std::vector<int32_t> m_scratch_most;
    std::vector<Object*> m_scratch_by_slot;
    bool m_state_dirty = true, m_list_dirty = true;
    bool m_closest_enabled = false;

    SimulationTime m_start_date;
    SimulationTime m_date;
    ObjectHistory m_object_history;
    std::list<std::unique_ptr<Object>> m_object_list;
    int m_simulation_seconds_per_tick = 60 * 60 * 12; // src/World.hpp:78
    bool m_is_forward_simulated = false;
    Object* m_light_source = nullptr;
};

// .essa world files (reference src/ConfigLoader.cpp:17-30,61-78,110-270).  Returns "" or an error
// message "<what> at <line>:<column>".
std::string load_essa_text(std::string const& text, World& world);
std::string load_essa_file(std::string const& filename, World& world);

} // namespace essa
