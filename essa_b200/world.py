"""World / Object for Python: the PySSA scripting surface (reference docs/PySSA/World.md,
docs/PySSA/Object.md, src/World.cpp:244-285, src/Object.cpp:427-551) over the headless facade.

    world = World.load("worlds/solar.essa")
    world.simulation_seconds_per_tick = 600
    world.update(10)                       # not in PySSA (the GUI drives update there); added for scripts
    earth = world.get_object_by_name("Earth")
    earth.pos, earth.vel, earth.mass

Attribute names and units follow the reference.  One conscious deviation: the reference's Python
`mass` getter returns gf*G and its setter stores value/G (src/Object.cpp:539-549), inverted with
respect to Object::mass() = gf/G (src/Object.hpp:72); here `mass` is kilograms both ways, as the
docs say ("Object mass (in kg)").
"""
import ctypes as C
import weakref

import numpy as np

from . import capi

_P = C.c_void_p
_dp = capi._dp
_u8p = capi._u8p
_i32p = capi._i32p
_fp = capi._fp

WORLD_SYMBOLS = [
    ("essa_world_create", _P, [C.c_int]),
    ("essa_world_destroy", None, [_P]),
    ("essa_world_last_error", C.c_char_p, [_P]),
    ("essa_world_load", C.c_int, [_P, C.c_char_p]),
    ("essa_world_load_text", C.c_int, [_P, C.c_char_p]),
    ("essa_world_add_object", C.c_int, [_P, C.c_double, C.c_double, _dp, _dp, _u8p, C.c_char_p, C.c_uint]),
    ("essa_world_add_orbiting_ap_pe", C.c_int, [_P, C.c_int, C.c_double, C.c_float, C.c_float, C.c_float, C.c_int, C.c_double,
                                                C.c_double, C.c_char_p]),
    ("essa_world_add_orbiting_maj_ecc", C.c_int, [_P, C.c_int, C.c_double, C.c_float, C.c_float, C.c_double, C.c_int, C.c_double,
                                                  C.c_double, C.c_char_p]),
    ("essa_world_delete_object", C.c_int, [_P, C.c_int]),
    ("essa_world_mark_deleted", C.c_int, [_P, C.c_int]),
    ("essa_world_count", C.c_int, [_P]),
    ("essa_world_find", C.c_int, [_P, C.c_char_p]),
    ("essa_world_date", C.c_int64, [_P]),
    ("essa_world_set_date", C.c_int, [_P, C.c_int64]),
    ("essa_world_get_dt", C.c_int, [_P]),
    ("essa_world_set_dt", C.c_int, [_P, C.c_int]),
    ("essa_world_set_flags", C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    ("essa_world_update", C.c_int, [_P, C.c_int]),
    ("essa_world_set_forces", C.c_int, [_P]),
    ("essa_world_last_step_ms", C.c_double, [_P]),
    ("essa_world_debug_stash_last_object", C.c_int, [_P]),
    ("essa_world_object_history_size", C.c_int, [_P]),
    ("essa_world_set_persistent", C.c_int, [_P, C.c_int, C.c_int]),
    ("essa_world_persist_stats", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("essa_world_get_state", C.c_int, [_P, _dp, _dp, _dp, _dp, _i32p, _dp, _i32p, _dp]),
    ("essa_world_set_object", C.c_int, [_P, C.c_int, _dp, _dp, C.c_double]),
    ("essa_world_object_name", C.c_char_p, [_P, C.c_int]),
    ("essa_world_set_object_name", C.c_int, [_P, C.c_int, C.c_char_p]),
    ("essa_world_object_props", C.c_int, [_P, C.c_int, _dp]),
    ("essa_world_set_object_props", C.c_int, [_P, C.c_int, C.c_double, _u8p]),
    ("essa_world_trail", C.c_int, [_P, C.c_int, _fp, C.POINTER(capi.TrailMeta)]),
    ("essa_world_history", C.c_int, [_P, C.c_int, _dp]),
    ("essa_world_reset_history", C.c_int, [_P, C.c_int]),
    ("essa_world_clone_for_forward_simulation", _P, [_P]),
    ("essa_world_is_forward_simulated", C.c_int, [_P]),
    ("essa_world_closest_approach", C.c_int, [_P, C.c_int, C.c_int, _dp]),
    ("essa_world_context", _P, [_P]),
]

_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        for name, res, args in WORLD_SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _bound = True
    return L


class Object:
    """One body.  Either free-standing (created by a script, to be passed to World.add_object) or a
    weak reference to an object of a world; using it after the object was removed raises
    ReferenceError, as in PySSA."""

    def __init__(self, *, mass, radius, pos=(0.0, 0.0, 0.0), vel=(0.0, 0.0, 0.0), color=(255, 255, 255), name="Planet", period=1):
        # keyword-only, as Object.create_for_python (src/Object.cpp:439-470)
        self._pending = dict(mass=float(mass), radius=float(radius), pos=tuple(map(float, pos)), vel=tuple(map(float, vel)),
                             color=tuple(int(c) for c in color), name=str(name), period=int(period))
        self._world = None
        self._uid = None

    @classmethod
    def _wrap(cls, world, uid):
        o = cls.__new__(cls)
        o._pending = None
        o._world = weakref.ref(world)
        o._uid = uid
        return o

    def _loc(self):
        if self._pending is not None:
            raise ReferenceError("object is not in a world yet")
        w = self._world()
        if w is None or self._uid not in w._uids:
            raise ReferenceError("the native object no longer exists")
        return w, w._uids.index(self._uid)

    def _vec(self, which):
        if self._pending is not None:
            return self._pending[which]
        w, i = self._loc()
        return tuple(w.state()[which][i])

    @property
    def pos(self):
        return self._vec("pos")

    @pos.setter
    def pos(self, v):
        if self._pending is not None:
            self._pending["pos"] = tuple(map(float, v))
            return
        w, i = self._loc()
        a = np.ascontiguousarray(v, dtype=np.float64)
        w._ck(w._L.essa_world_set_object(w._h, i, capi._ptr(a, _dp), None, -1.0))

    @property
    def vel(self):
        return self._vec("vel")

    @vel.setter
    def vel(self, v):
        if self._pending is not None:
            self._pending["vel"] = tuple(map(float, v))
            return
        w, i = self._loc()
        a = np.ascontiguousarray(v, dtype=np.float64)
        w._ck(w._L.essa_world_set_object(w._h, i, None, capi._ptr(a, _dp), -1.0))

    @property
    def acc(self):
        return self._vec("acc")

    @property
    def name(self):
        if self._pending is not None:
            return self._pending["name"]
        w, i = self._loc()
        return w._L.essa_world_object_name(w._h, i).decode()

    @name.setter
    def name(self, v):
        if self._pending is not None:
            self._pending["name"] = str(v)
            return
        w, i = self._loc()
        w._ck(w._L.essa_world_set_object_name(w._h, i, str(v).encode()))

    def _props(self):
        w, i = self._loc()
        out = np.zeros(7)
        w._ck(w._L.essa_world_object_props(w._h, i, capi._ptr(out, _dp)))
        return out

    @property
    def radius(self):
        return self._pending["radius"] if self._pending is not None else float(self._props()[0])

    @radius.setter
    def radius(self, v):
        if self._pending is not None:
            self._pending["radius"] = float(v)
            return
        w, i = self._loc()
        w._ck(w._L.essa_world_set_object_props(w._h, i, float(v), None))

    @property
    def color(self):
        if self._pending is not None:
            return self._pending["color"]
        return tuple(int(c) for c in self._props()[4:7])

    @color.setter
    def color(self, v):
        if self._pending is not None:
            self._pending["color"] = tuple(int(c) for c in v)
            return
        w, i = self._loc()
        rgb = np.array([int(c) for c in v], dtype=np.uint8)
        w._ck(w._L.essa_world_set_object_props(w._h, i, -1.0, capi._ptr(rgb, _u8p)))

    @property
    def mass(self):
        if self._pending is not None:
            return self._pending["mass"]
        w, i = self._loc()
        return float(w.state()["gf"][i]) / capi.GRAVITY

    @mass.setter
    def mass(self, v):
        if self._pending is not None:
            self._pending["mass"] = float(v)
            return
        w, i = self._loc()
        w._ck(w._L.essa_world_set_object(w._h, i, None, None, float(v) * capi.GRAVITY))

    def attraction(self, other):
        """Object::attraction (src/Object.cpp:53-58): other.gf / r^2 along (this - other)."""
        d = np.array(self.pos) - np.array(other.pos)
        r2 = float(d @ d)
        gf = other.mass * capi.GRAVITY
        return tuple(d / np.sqrt(r2) * (gf / r2))

    # beyond PySSA: read-only views of what World::update records
    @property
    def most_attracting_object(self):
        w, i = self._loc()
        m = int(w.state()["most"][i])
        return None if m < 0 else Object._wrap(w, w._uids[m])

    @property
    def trail(self):
        w, i = self._loc()
        return w.trail(i)

    @property
    def history(self):
        w, i = self._loc()
        return w.history(i)


class World:
    def __init__(self, device=0, _handle=None):
        self._L = _lib()
        self._h = _P(_handle if _handle is not None else self._L.essa_world_create(device))
        self._uids = []
        self._next_uid = 0
        self._resync_uids()

    def _resync_uids(self):
        n = self._L.essa_world_count(self._h)
        self._uids = list(range(self._next_uid, self._next_uid + n))
        self._next_uid += n

    def close(self):
        if getattr(self, "_h", None):
            self._L.essa_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise capi.EssaError(rc, (self._L.essa_world_last_error(self._h) or b"").decode())
        return rc

    @classmethod
    def load(cls, path, device=0):
        w = cls(device)
        w._ck(w._L.essa_world_load(w._h, str(path).encode()))
        w._resync_uids()
        return w

    @classmethod
    def loads(cls, text, device=0):
        w = cls(device)
        w._ck(w._L.essa_world_load_text(w._h, text.encode()))
        w._resync_uids()
        return w

    # ---- PySSA surface ----
    @property
    def simulation_seconds_per_tick(self):
        return self._L.essa_world_get_dt(self._h)

    @simulation_seconds_per_tick.setter
    def simulation_seconds_per_tick(self, v):
        self._ck(self._L.essa_world_set_dt(self._h, int(v)))

    def add_object(self, obj=None, **kw):
        if obj is None:
            obj = Object(**kw)
        p = obj._pending
        if p is None:
            raise ValueError("object already belongs to a world")
        pos = np.array(p["pos"], dtype=np.float64)
        vel = np.array(p["vel"], dtype=np.float64)
        rgb = np.array(p["color"], dtype=np.uint8)
        before = self._L.essa_world_count(self._h)
        self._ck(self._L.essa_world_add_object(self._h, p["mass"], p["radius"], capi._ptr(pos, _dp), capi._ptr(vel, _dp),
                                               capi._ptr(rgb, _u8p), p["name"].encode(), p["period"]))
        after = self._L.essa_world_count(self._h)
        if after == before + 1:
            uid = self._next_uid
            self._next_uid += 1
            self._uids.append(uid)
        else:  # add_object may also drop objects created in the future (src/World.cpp:37-39)
            self._resync_uids()
            uid = self._uids[-1]
        obj._pending = None
        obj._world = weakref.ref(self)
        obj._uid = uid
        return obj

    def get_object_by_name(self, name):
        i = self._L.essa_world_find(self._h, str(name).encode())
        return None if i < 0 else Object._wrap(self, self._uids[i])

    # ---- stepping (what SimulationView::update does per GUI tick, src/SimulationView.cpp:374-397) ----
    def update(self, steps):
        self._ck(self._L.essa_world_update(self._h, int(steps)))
        if self._L.essa_world_object_history_size(self._h) and self._L.essa_world_count(self._h) != len(self._uids):
            self._resync_uids()  # objects left / rejoined the list with the date (ObjectHistory)

    def set_forces(self):
        self._ck(self._L.essa_world_set_forces(self._h))

    def debug_stash_last_object(self):
        """Seeds the ObjectHistory with the list's last object (see include/essa_world.h)."""
        self._ck(self._L.essa_world_debug_stash_last_object(self._h))
        self._resync_uids()

    @property
    def object_history_size(self):
        return self._L.essa_world_object_history_size(self._h)

    def set_persistent(self, enable=True, idle_timeout_us=0):
        """Keep the small-world kernel resident between update() calls (essa_persist_configure); default on."""
        self._ck(self._L.essa_world_set_persistent(self._h, int(enable), int(idle_timeout_us)))

    def persist_stats(self):
        """{live, ticks served through the mailbox, persistent kernels launched, idle retirements seen}."""
        out = (C.c_int64 * 6)()
        self._ck(self._L.essa_world_persist_stats(self._h, out))
        return dict(zip(("live", "ticks", "launches", "idle_exits", "wait_ns", "host_ns"), (int(v) for v in out)))

    def set_flags(self, offset_trails=True, exact_order=False, two_pass=False):
        self._ck(self._L.essa_world_set_flags(self._h, int(offset_trails), int(exact_order), int(two_pass)))

    @property
    def date(self):
        return self._L.essa_world_date(self._h)

    @date.setter
    def date(self, v):
        self._ck(self._L.essa_world_set_date(self._h, int(v)))

    @property
    def last_step_ms(self):
        return self._L.essa_world_last_step_ms(self._h)

    def __len__(self):
        return self._L.essa_world_count(self._h)

    def objects(self):
        return [Object._wrap(self, u) for u in self._uids]

    def names(self):
        return [self._L.essa_world_object_name(self._h, i).decode() for i in range(len(self))]

    def state(self):
        n = len(self)
        out = dict(pos=np.zeros((n, 3)), vel=np.zeros((n, 3)), acc=np.zeros((n, 3)), gf=np.zeros(n), most=np.zeros(n, dtype=np.int32),
                   appe=np.zeros((n, 4)), active=np.zeros(n, dtype=np.int32), maxattr=np.zeros(n))
        self._ck(self._L.essa_world_get_state(self._h, capi._ptr(out["pos"], _dp), capi._ptr(out["vel"], _dp), capi._ptr(out["acc"], _dp),
                                              capi._ptr(out["gf"], _dp), capi._ptr(out["most"], _i32p), capi._ptr(out["appe"], _dp),
                                              capi._ptr(out["active"], _i32p), capi._ptr(out["maxattr"], _dp)))
        return out

    def props(self, index):
        out = np.zeros(7)
        self._ck(self._L.essa_world_object_props(self._h, index, capi._ptr(out, _dp)))
        return dict(radius=out[0], orbit_period=out[1], eccentrity=out[2], trail_capacity=int(out[3]), color=tuple(int(c) for c in out[4:7]))

    def set_object_state(self, index, pos=None, vel=None, gf=-1.0):
        p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self._ck(self._L.essa_world_set_object(self._h, index, capi._ptr(p, _dp), capi._ptr(v, _dp), float(gf)))

    def delete_object(self, index):
        """World::delete_object_by_ptr: the object leaves the list."""
        self._ck(self._L.essa_world_delete_object(self._h, index))
        del self._uids[index]

    def mark_deleted(self, index):
        """Object::delete_object: stays in the list, masked from the current date on."""
        self._ck(self._L.essa_world_mark_deleted(self._h, index))

    def trail(self, index):
        meta = capi.TrailMeta()
        self._ck(self._L.essa_world_trail(self._h, index, None, C.byref(meta)))
        verts = np.zeros((meta.capacity, 3), dtype=np.float32)
        self._ck(self._L.essa_world_trail(self._h, index, capi._ptr(verts, _fp), C.byref(meta)))
        return dict(capacity=meta.capacity, verts=verts, append_offset=meta.append_offset, length=meta.length,
                    offset=np.array(list(meta.offset)), last_xy_angle=meta.last_xy_angle)

    def history(self, index):
        n = self._ck(self._L.essa_world_history(self._h, index, None))
        out = np.zeros((n, 6))
        self._ck(self._L.essa_world_history(self._h, index, capi._ptr(out, _dp)))
        return out

    def reset_history(self, index):
        self._ck(self._L.essa_world_reset_history(self._h, index))

    def clone_for_forward_simulation(self):
        h = self._L.essa_world_clone_for_forward_simulation(self._h)
        if not h:
            raise capi.EssaError(-4, (self._L.essa_world_last_error(self._h) or b"").decode())
        return World(_handle=h)

    @property
    def is_forward_simulated(self):
        return bool(self._L.essa_world_is_forward_simulated(self._h))

    def closest_approach(self, index, other):
        out = np.zeros(7)
        if self._ck(self._L.essa_world_closest_approach(self._h, index, other, capi._ptr(out, _dp))) == 0:
            return None
        return dict(distance=out[0], this_position=out[1:4].copy(), other_object_position=out[4:7].copy())
