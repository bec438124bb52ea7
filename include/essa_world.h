/* ==========================================================================================
 * essa_world.h -- flat C interface of the World / Object facade (essa_b200/csrc/facade.hpp).
 *
 * The facade mirrors the reference's C++ classes (src/World.hpp:19-95, src/Object.hpp:21-191);
 * this header exposes the same operations to C, ctypes and other FFIs, the way the reference's
 * embedded-Python glue (src/World.cpp:244-285, src/Object.cpp:427-551) exposes them to scripts.
 * Objects are addressed by list index (insertion order, as World::m_object_list).
 * Return codes as in essa_b200.h; essa_world_last_error() gives the message.
 * ========================================================================================== */
#ifndef ESSA_WORLD_H
#define ESSA_WORLD_H

#include <stdint.h>

#include "essa_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct essa_world essa_world;

/* World::World() (src/World.cpp:22-27).  No device is touched until the first update. */
essa_world* essa_world_create(int device);
void essa_world_destroy(essa_world* w);
const char* essa_world_last_error(const essa_world* w);

/* World::reset(filename) (src/World.cpp:171-202) / ConfigLoader (src/ConfigLoader.cpp). */
int essa_world_load(essa_world* w, const char* essa_file);
int essa_world_load_text(essa_world* w, const char* essa_text);
/* World::add_object(make_unique<Object>(mass, radius, pos, vel, color, name, period)); returns the index. */
int essa_world_add_object(essa_world* w, double mass, double radius, const double pos[3], const double vel[3], const uint8_t rgb[3],
                          const char* name, unsigned period);
/* Object::create_object_relative_to_ap_pe / _maj_ecc + add_object (src/Object.cpp:309-383); angles in radians. */
int essa_world_add_orbiting_ap_pe(essa_world* w, int around, double mass, float radius, float apoapsis, float periapsis, int direction,
                                  double theta, double alpha, const char* name);
int essa_world_add_orbiting_maj_ecc(essa_world* w, int around, double mass, float radius, float semi_major, double eccentrity,
                                    int direction, double theta, double alpha, const char* name);
int essa_world_delete_object(essa_world* w, int index);  /* World::delete_object_by_ptr */
int essa_world_mark_deleted(essa_world* w, int index);   /* Object::delete_object */

int essa_world_count(const essa_world* w);
int essa_world_find(essa_world* w, const char* name);    /* World::get_object_by_name; -1 when absent */
int64_t essa_world_date(const essa_world* w);
int essa_world_set_date(essa_world* w, int64_t date);
int essa_world_get_dt(const essa_world* w);              /* simulation_seconds_per_tick */
int essa_world_set_dt(essa_world* w, int seconds);
/* offset_trails (SimulationView::offset_trails()), exact_order, two_pass (see facade.hpp). */
int essa_world_set_flags(essa_world* w, int offset_trails, int exact_order, int two_pass);

/* World::update(steps) / World::set_forces().  steps == 0 is an argument error (assert in the reference). */
int essa_world_update(essa_world* w, int steps);
int essa_world_set_forces(essa_world* w);
double essa_world_last_step_ms(const essa_world* w);
/* ObjectHistory (src/ObjectHistory.cpp, src/World.cpp:59-84): objects leave the list when the date moves below their creation
 * date and come back when it moves forward again -- but only once the history holds an entry, which the reference's current
 * revision never creates.  This seeds it with the list's last object (tests; a future revision's rewind). */
int essa_world_debug_stash_last_object(essa_world* w);
int essa_world_object_history_size(const essa_world* w);
/* essa_persist_configure / essa_persist_stats of the world's context (essa_b200.h): keep the small-world kernel
 * resident between update() calls (default on). */
int essa_world_set_persistent(essa_world* w, int enable, int idle_timeout_us);
int essa_world_persist_stats(essa_world* w, int64_t out[6]);

/* Bulk read of the object list: pos/vel/acc are 3N doubles (xyz interleaved), gf N, most N
 * (index or -1), appe 4N (ap, pe, ap_vel, pe_vel), active N, maxattr N.  Any pointer may be null. */
int essa_world_get_state(essa_world* w, double* pos, double* vel, double* acc, double* gf, int32_t* most, double* appe,
                         int32_t* active, double* maxattr);
/* Object setters (set_pos / set_vel; gf < 0 leaves the gravity factor alone). */
int essa_world_set_object(essa_world* w, int index, const double pos[3], const double vel[3], double gf);
const char* essa_world_object_name(essa_world* w, int index);
int essa_world_set_object_name(essa_world* w, int index, const char* name);
/* out = {radius, orbit_period, eccentrity, trail_capacity, r, g, b} */
int essa_world_object_props(essa_world* w, int index, double out[7]);
int essa_world_set_object_props(essa_world* w, int index, double radius, const uint8_t rgb[3]);

/* Object::trail(): verts = 3 * capacity floats (may be null to query meta only). */
int essa_world_trail(essa_world* w, int index, float* verts, essa_trail_meta* meta);
/* Object history, oldest first, 6 doubles per entry; returns the count (out may be null). */
int essa_world_history(essa_world* w, int index, double* out);
int essa_world_reset_history(essa_world* w, int index);  /* Object::reset_history */

/* World::clone_for_forward_simulation (src/World.cpp:230-238): a new, forward-simulated world. */
essa_world* essa_world_clone_for_forward_simulation(essa_world* w);
int essa_world_is_forward_simulated(const essa_world* w);
/* out = {distance, this xyz, other xyz}; 1 if present, 0 if not (src/Object.cpp:405-417). */
int essa_world_closest_approach(essa_world* w, int index, int other, double out[7]);

/* The device context behind the world (created on demand), for essa_energy() and friends. */
essa_ctx* essa_world_context(essa_world* w);

#ifdef __cplusplus
}
#endif
#endif /* ESSA_WORLD_H */
