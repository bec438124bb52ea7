"""ctypes binding of the CPU oracle (oracle/libessa_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  Nothing under essa_b200/ imports it.
See oracle/essa_oracle.hpp for what it restates (reference file:line per function).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libessa_oracle.so")

GRAVITY = 6.67408e-11
AU = 149597870700.0
START_DATE = 640566000


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("essa_oracle.cpp", "essa_oracle.hpp")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libessa_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ow_create.restype = C.c_void_p
        L.ow_destroy.argtypes = [C.c_void_p]
        L.ow_load_essa_file.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.ow_add_object.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, _dp, C.c_char_p, C.c_uint]
        L.ow_set_dt.argtypes = [C.c_void_p, C.c_int]
        L.ow_get_dt.argtypes = [C.c_void_p]
        L.ow_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ow_update.argtypes = [C.c_void_p, C.c_int]
        L.ow_set_forces.argtypes = [C.c_void_p]
        L.ow_debug_stash_last.argtypes = [C.c_void_p]
        L.ow_object_history_size.argtypes = [C.c_void_p]
        L.ow_count.argtypes = [C.c_void_p]
        L.ow_date.argtypes = [C.c_void_p]
        L.ow_date.restype = C.c_longlong
        L.ow_set_date.argtypes = [C.c_void_p, C.c_longlong]
        L.ow_is_forward_simulated.argtypes = [C.c_void_p]
        L.ow_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip, _dp, _ip, _dp]
        L.ow_set_object_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double]
        L.ow_delete_object.argtypes = [C.c_void_p, C.c_int]
        L.ow_set_dates.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_int]
        L.ow_remove_object.argtypes = [C.c_void_p, C.c_int]
        L.ow_reset_history.argtypes = [C.c_void_p, C.c_int]
        L.ow_name.argtypes = [C.c_void_p, C.c_int]
        L.ow_name.restype = C.c_char_p
        for f in (L.ow_radius, L.ow_orbit_len, L.ow_eccentrity):
            f.argtypes = [C.c_void_p, C.c_int]
            f.restype = C.c_double
        L.ow_trail_capacity.argtypes = [C.c_void_p, C.c_int]
        L.ow_trail_get.argtypes = [C.c_void_p, C.c_int, _fp, _ip, _dp, _dp]
        L.ow_history_len.argtypes = [C.c_void_p, C.c_int]
        L.ow_history_get.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ow_clone_for_forward_simulation.argtypes = [C.c_void_p]
        L.ow_clone_for_forward_simulation.restype = C.c_void_p
        L.ow_closest_approach.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ow_time_update.argtypes = [C.c_void_p, C.c_int]
        L.ow_time_update.restype = C.c_double
        L.ow_time_force_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.ow_time_force_rows.restype = C.c_double
        L.ow_add_bodies.argtypes = [C.c_void_p, C.c_int] + [_dp] * 7
        L.osoa_forces_pairshared.argtypes = [C.c_int] + [_dp] * 7
        L.osoa_forces_rows.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _ip, C.c_int, _dp]
        L.osoa_forces_rows_omp.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _ip, C.c_int, _dp]
        L.osoa_forces_rows_ld.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _ip, C.c_int, _dp]
        L.osoa_step.argtypes = [C.c_int] + [_dp] * 10 + [C.c_int, C.c_int]
        L.osoa_energy.argtypes = [C.c_int] + [_dp] * 7 + [C.c_double, _dp, _dp, _dp]
        _lib = L
    return _lib


class World:
    """oracle::World (restates src/World.cpp:22-139)."""

    def __init__(self, handle=None):
        self._L = lib()
        self._h = C.c_void_p(handle if handle is not None else self._L.ow_create())

    def __del__(self):
        try:
            if self._h:
                self._L.ow_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def load(cls, path, angle_as_float=False):
        w = cls()
        err = C.create_string_buffer(512)
        rc = w._L.ow_load_essa_file(w._h, os.fsencode(path), int(angle_as_float), err, 512)
        if rc != 0:
            raise ValueError("World loading failed: " + err.value.decode())
        return w

    def add_object(self, mass, radius, pos, vel, name="Planet", period=1):
        p = np.ascontiguousarray(pos, dtype=np.float64)
        v = np.ascontiguousarray(vel, dtype=np.float64)
        return self._L.ow_add_object(self._h, mass, radius, _d(p), _d(v), name.encode(), int(period))

    def add_bodies(self, x, y, z, vx, vy, vz, mass):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, vx, vy, vz, mass)]
        self._L.ow_add_bodies(self._h, len(arrs[0]), *[_d(a) for a in arrs])

    @property
    def dt(self):
        return self._L.ow_get_dt(self._h)

    @dt.setter
    def dt(self, v):
        self._L.ow_set_dt(self._h, int(v))

    def set_flags(self, offset_trails=True, legacy_euler=False):
        self._L.ow_set_flags(self._h, int(offset_trails), int(legacy_euler))

    def update(self, steps):
        assert steps != 0
        self._L.ow_update(self._h, int(steps))

    def set_forces(self):
        self._L.ow_set_forces(self._h)

    def __len__(self):
        return self._L.ow_count(self._h)

    @property
    def date(self):
        return self._L.ow_date(self._h)

    @date.setter
    def date(self, v):
        self._L.ow_set_date(self._h, int(v))

    def state(self):
        n = len(self)
        out = dict(pos=np.zeros((n, 3)), vel=np.zeros((n, 3)), acc=np.zeros((n, 3)), gf=np.zeros(n),
                   most=np.zeros(n, dtype=np.int32), appe=np.zeros((n, 4)), active=np.zeros(n, dtype=np.int32),
                   maxattr=np.zeros(n))
        self._L.ow_get_state(self._h, _d(out["pos"]), _d(out["vel"]), _d(out["acc"]), _d(out["gf"]), _i(out["most"]),
                             _d(out["appe"]), _i(out["active"]), _d(out["maxattr"]))
        return out

    def set_object_state(self, idx, pos=None, vel=None, gf=-1.0):
        p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self._L.ow_set_object_state(self._h, idx, _d(p), _d(v), gf)

    def set_dates(self, idx, creation, deletion=0, deleted=False):
        self._L.ow_set_dates(self._h, idx, int(creation), int(deletion), int(deleted))

    def delete_object(self, idx):
        self._L.ow_delete_object(self._h, idx)

    def remove_object(self, idx):
        """World::delete_object_by_ptr (src/World.cpp:208-219)."""
        self._L.ow_remove_object(self._h, idx)

    def debug_stash_last(self):
        """Moves the list's last object into the ObjectHistory (see ow_debug_stash_last)."""
        self._L.ow_debug_stash_last(self._h)

    @property
    def object_history_size(self):
        return self._L.ow_object_history_size(self._h)

    def reset_history(self, idx):
        self._L.ow_reset_history(self._h, idx)

    @property
    def is_forward_simulated(self):
        return bool(self._L.ow_is_forward_simulated(self._h))

    def name(self, idx):
        return self._L.ow_name(self._h, idx).decode()

    def names(self):
        return [self.name(i) for i in range(len(self))]

    def radius(self, idx):
        return self._L.ow_radius(self._h, idx)

    def orbit_len(self, idx):
        return self._L.ow_orbit_len(self._h, idx)

    def eccentrity(self, idx):
        return self._L.ow_eccentrity(self._h, idx)

    def trail(self, idx):
        cap = self._L.ow_trail_capacity(self._h, idx)
        verts = np.zeros((cap, 3), dtype=np.float32)
        meta = np.zeros(2, dtype=np.int32)
        off = np.zeros(3)
        ang = np.zeros(1)
        self._L.ow_trail_get(self._h, idx, verts.ctypes.data_as(_fp), _i(meta), _d(off), _d(ang))
        return dict(capacity=cap, verts=verts, append_offset=int(meta[0]), length=int(meta[1]), offset=off,
                    last_xy_angle=float(ang[0]))

    def history(self, idx):
        n = self._L.ow_history_len(self._h, idx)
        out = np.zeros((n, 6))
        self._L.ow_history_get(self._h, idx, _d(out))
        return out

    def clone_for_forward_simulation(self):
        return World(self._L.ow_clone_for_forward_simulation(self._h))

    def closest_approach(self, idx, other):
        out = np.zeros(7)
        if not self._L.ow_closest_approach(self._h, idx, other, _d(out)):
            return None
        return dict(distance=out[0], this_position=out[1:4].copy(), other_object_position=out[4:7].copy())

    def time_update(self, steps):
        return self._L.ow_time_update(self._h, int(steps))

    def time_force_rows(self, i0, i1):
        pairs = C.c_longlong(0)
        t = self._L.ow_time_force_rows(self._h, int(i0), int(i1), C.byref(pairs))
        return t, pairs.value


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def forces_pairshared(x, y, z, gf):
    """World::set_forces order on arrays (src/World.cpp:42-57, src/Object.cpp:64-79)."""
    x, y, z, gf = map(_c, (x, y, z, gf))
    n = len(x)
    a = np.zeros((3, n))
    lib().osoa_forces_pairshared(n, _d(x), _d(y), _d(z), _d(gf), _d(a[0]), _d(a[1]), _d(a[2]))
    return a


def forces_rows(x, y, z, gf, rows, mode="ref"):
    """Acceleration of the listed rows; mode 'ref' (reference order, binary64), 'omp', or 'ld'
    (x87 extended precision yardstick)."""
    x, y, z, gf = map(_c, (x, y, z, gf))
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((len(rows), 3))
    f = {"ref": lib().osoa_forces_rows, "omp": lib().osoa_forces_rows_omp, "ld": lib().osoa_forces_rows_ld}[mode]
    r = f(len(x), _d(x), _d(y), _d(z), _d(gf), _i(rows), len(rows), _d(out))
    if mode == "omp":
        return out, r
    return out


def step_soa(x, y, z, vx, vy, vz, gf, steps, dt):
    """Reference-semantics KDK on arrays (two force passes per step); returns new arrays + acc."""
    arrs = [_c(a).copy() for a in (x, y, z, vx, vy, vz)]
    gf = _c(gf)
    n = len(gf)
    acc = np.zeros((3, n))
    lib().osoa_step(n, *[_d(a) for a in arrs], _d(gf), _d(acc[0]), _d(acc[1]), _d(acc[2]), int(steps), int(dt))
    return arrs, acc


def energy(x, y, z, vx, vy, vz, gf):
    arrs = [_c(a) for a in (x, y, z, vx, vy, vz, gf)]
    ke, pe = C.c_double(0), C.c_double(0)
    mom = np.zeros(3)
    lib().osoa_energy(len(arrs[0]), *[_d(a) for a in arrs], GRAVITY, C.byref(ke), C.byref(pe), _d(mom))
    return ke.value, pe.value, mom
