"""TEST INFRASTRUCTURE.  The reference's OWN History and Trail classes (src/History.cpp, src/Trail.cpp compiled where they lie
under /root/reference into oracle/_ref/libessa_ref.so, see oracle/Makefile) next to the oracle's restatement of them, both
behind the same small C ABI, so that a test can feed both the same inputs and demand the same bits.

The rest of the reference cannot be compiled offline (GUI / OpenGL / embedded-Python headers throughout, CMake FetchContent);
for History and Trail the only things substituted are the un-vendored EssaUtil / EssaGUI headers (oracle/ref_stubs)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libessa_ref.so")
REFERENCE = "/root/reference"

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Builds oracle/_ref/libessa_ref.so when the reference tree is present (the authoring container); elsewhere the
    prebuilt library that travelled with the snapshot is used.  Returns the path or None."""
    have_ref = os.path.exists(os.path.join(REFERENCE, "src", "History.cpp"))
    if have_ref and (force or not os.path.exists(REF_LIB)
                     or os.path.getmtime(REF_LIB) < os.path.getmtime(os.path.join(_HERE, "ref_driver.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_ref/libessa_ref.so"], stdout=subprocess.DEVNULL)
    return REF_LIB if os.path.exists(REF_LIB) else None


def _bind(L, prefix):
    f = {}
    sig = {"history_new": (C.c_void_p, [C.c_uint, _dp]), "history_free": (None, [C.c_void_p]),
           "history_move_forward": (C.c_int, [C.c_void_p, _dp, _dp]), "history_move_backward": (C.c_int, [C.c_void_p, _dp, _dp]),
           "history_reset": (None, [C.c_void_p]), "history_read": (C.c_int, [C.c_void_p, _dp, C.c_int, _ip]),
           "trail_new": (C.c_void_p, [C.c_size_t]), "trail_free": (None, [C.c_void_p]),
           "trail_push_back": (None, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
           "trail_change_current": (None, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
           "trail_set_offset": (None, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
           "trail_recalculate_with_offset": (None, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
           "trail_reset": (None, [C.c_void_p]), "trail_read": (None, [C.c_void_p, _fp, _ip, _dp])}
    for name, (res, args) in sig.items():
        fn = getattr(L, prefix + name)
        fn.restype = res
        fn.argtypes = args
        f[name] = fn
    return f


class _Impl:
    def __init__(self, fns):
        self.f = fns


def impls():
    """(reference, oracle) function tables, or None when the reference library is not available."""
    path = build()
    if path is None:
        return None
    import oracle
    return _Impl(_bind(C.CDLL(path), "ref_")), _Impl(_bind(oracle.lib(), "o_"))


class History:
    def __init__(self, impl, segments, first):
        self.f = impl.f
        a = np.ascontiguousarray(first, dtype=np.float64)
        self.h = self.f["history_new"](segments, a.ctypes.data_as(_dp))

    def __del__(self):
        self.f["history_free"](self.h)

    def _move(self, which, e):
        a = np.ascontiguousarray(e, dtype=np.float64)
        out = np.zeros(6)
        got = self.f[which](self.h, a.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        return out if got else None

    def move_forward(self, e):
        return self._move("history_move_forward", e)

    def move_backward(self, e):
        return self._move("history_move_backward", e)

    def reset(self):
        self.f["history_reset"](self.h)

    def read(self):
        side = C.c_int(0)
        n = self.f["history_read"](self.h, None, 0, C.byref(side))
        out = np.zeros((n, 6))
        self.f["history_read"](self.h, out.ctypes.data_as(_dp), n, C.byref(side))
        return out, side.value


class Trail:
    def __init__(self, impl, cap):
        self.f = impl.f
        self.t = self.f["trail_new"](cap)

    def __del__(self):
        self.f["trail_free"](self.t)

    def push_back(self, p):
        self.f["trail_push_back"](self.t, *map(float, p))

    def change_current(self, p):
        self.f["trail_change_current"](self.t, *map(float, p))

    def set_offset(self, p):
        self.f["trail_set_offset"](self.t, *map(float, p))

    def recalculate_with_offset(self, p):
        self.f["trail_recalculate_with_offset"](self.t, *map(float, p))

    def reset(self):
        self.f["trail_reset"](self.t)

    def read(self):
        meta = (C.c_int * 3)()
        real = (C.c_double * 4)()
        self.f["trail_read"](self.t, None, meta, real)
        verts = np.zeros((meta[0], 3), dtype=np.float32)
        self.f["trail_read"](self.t, verts.ctypes.data_as(_fp), meta, real)
        return dict(verts=verts, capacity=meta[0], append_offset=meta[1], length=meta[2], offset=np.array(real[:3]), last_xy_angle=real[3])
