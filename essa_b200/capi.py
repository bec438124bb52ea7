"""ctypes binding of include/essa_b200.h (essa_b200/lib/libessa_b200.so).

Thin by design: every method is one C-ABI call on numpy buffers.  There is no CPU fallback --
loading fails loudly when the library is missing, and Context() fails when there is no B200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESSA_B200_LIB") or os.path.join(_HERE, "lib", "libessa_b200.so")  # (override: kernel experiments)

GRAVITY = 6.67408e-11
AU = 149597870700.0
START_DATE = 640566000
HISTORY_SEGMENTS = 1000
RESIDENT_MAX = 1024
ENSEMBLE_MAX = 320
NCCL_ID_BYTES = 128

KERNEL_AUTO, KERNEL_TILED, KERNEL_RESIDENT, KERNEL_PAIRED, KERNEL_GRID = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class EssaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("essa_b200 error %d: %s" % (code, message))
        self.code = code


class Bodies(C.Structure):
    _fields_ = [(n, _dp) for n in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "gf")] + [
        ("creation_date", _i64p), ("deletion_date", _i64p), ("deleted", _u8p), ("most", _i32p),
        ("max_attraction", _dp), ("ap", _dp), ("pe", _dp), ("ap_vel", _dp), ("pe_vel", _dp)]


class StepOpts(C.Structure):
    _fields_ = [("dt_seconds", C.c_int), ("forward_simulated", C.c_int), ("offset_trails", C.c_int),
                ("record", C.c_int), ("track_most_attracting", C.c_int), ("exact_order", C.c_int),
                ("two_pass", C.c_int), ("kernel", C.c_int), ("eps2", C.c_double)]


class TrailMeta(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("append_offset", C.c_int32), ("length", C.c_int32), ("pad", C.c_int32),
                ("offset", C.c_double * 3), ("last_xy_angle", C.c_double)]


class EnsembleOpts(C.Structure):
    _fields_ = [("copies", C.c_int), ("first_copy", C.c_int64), ("dt_seconds", C.c_int), ("seed", C.c_uint64),
                ("rel_sigma", C.c_double), ("exact_order", C.c_int), ("closest_approaches", C.c_int),
                ("approach_to", C.c_int), ("trail_every", C.c_int), ("trail_capacity", C.c_int)]


class TrailView(C.Structure):
    _fields_ = [("verts_device", C.c_void_p), ("capacity", C.c_int32), ("append_offset", C.c_int32), ("length", C.c_int32),
                ("pad", C.c_int32), ("offset", C.c_double * 3)]


# every symbol include/essa_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("essa_abi_version", C.c_int, []),
    ("essa_ctx_create", C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    ("essa_ctx_destroy", None, [_P]),
    ("essa_last_error", C.c_char_p, [_P]),
    ("essa_launch_count", C.c_int64, [_P]),
    ("essa_pass_count", C.c_int64, [_P]),
    ("essa_interaction_count", C.c_double, [_P]),
    ("essa_last_step_ms", C.c_double, [_P]),
    ("essa_last_force_ms", C.c_double, [_P]),
    ("essa_last_pass_phases", C.c_int, [_P, C.POINTER(C.c_double)]),
    ("essa_last_force_passes", C.c_int, [_P]),
    ("essa_last_step_breakdown", C.c_int, [_P, _dp]),
    ("essa_sync", C.c_int, [_P]),
    ("essa_upload", C.c_int, [_P, C.c_int, C.POINTER(Bodies)]),
    ("essa_download", C.c_int, [_P, C.POINTER(Bodies)]),
    ("essa_upload_shard", C.c_int, [_P, C.POINTER(Bodies)]),
    ("essa_download_shard", C.c_int, [_P, C.POINTER(Bodies)]),
    ("essa_set_date", C.c_int, [_P, C.c_int64]),
    ("essa_get_date", C.c_int64, [_P]),
    ("essa_body_count", C.c_int, [_P]),
    ("essa_step", C.c_int, [_P, C.c_int, C.POINTER(StepOpts)]),
    ("essa_step_async", C.c_int, [_P, C.c_int, C.POINTER(StepOpts)]),
    ("essa_set_forces", C.c_int, [_P, C.POINTER(StepOpts)]),
    ("essa_persist_configure", C.c_int, [_P, C.c_int, C.c_int]),
    ("essa_persist_stats", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("essa_record_configure", C.c_int, [_P, C.c_int, _i32p, C.c_int]),
    ("essa_tracked_count", C.c_int, [_P]),
    ("essa_record_reset", C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    ("essa_history_read", C.c_int, [_P, C.c_int, _dp]),
    ("essa_history_write", C.c_int, [_P, C.c_int, _dp, C.c_int, C.c_int]),
    ("essa_trail_read", C.c_int, [_P, C.c_int, _fp, C.POINTER(TrailMeta)]),
    ("essa_trail_write", C.c_int, [_P, C.c_int, _fp, C.POINTER(TrailMeta)]),
    ("essa_trail_device_view", C.c_int, [_P, C.c_int, C.POINTER(TrailView)]),
    ("essa_snapshot_size", C.c_int64, [_P]),
    ("essa_snapshot_write", C.c_int, [_P, C.c_void_p, C.c_int64]),
    ("essa_snapshot_read", C.c_int, [_P, C.c_void_p, C.c_int64]),
    ("essa_closest_enable", C.c_int, [_P, C.c_int]),
    ("essa_closest_read", C.c_int, [_P, C.c_int, C.c_int, _dp]),
    ("essa_energy", C.c_int, [_P, _dp, _dp, _dp]),
    ("essa_ensemble_init", C.c_int, [_P, C.POINTER(EnsembleOpts)]),
    ("essa_ensemble_step", C.c_int, [_P, C.c_int]),
    ("essa_ensemble_read", C.c_int, [_P, C.c_int, C.c_int] + [_dp] * 7),
    ("essa_ensemble_read_approaches", C.c_int, [_P, C.c_int, C.c_int, _dp, _dp]),
    ("essa_ensemble_read_trail", C.c_int, [_P, C.c_int, C.c_int, _dp, _i32p]),
    ("essa_ensemble_factor", C.c_double, [C.c_uint64, C.c_double, C.c_int64, C.c_int, C.c_int]),
    ("essa_nccl_unique_id", C.c_int, [_u8p]),
    ("essa_shard_attach", C.c_int, [_P, C.c_int, C.c_int, _u8p]),
    ("essa_shard_range", C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("essa_shard_rank", C.c_int, [_P]),
    ("essa_shard_world", C.c_int, [_P]),
    ("essa_measure_fp64_peak", C.c_double, [_P, C.c_double]),
    ("essa_measure_rsqrt_error", C.c_int, [_P, _dp, _dp]),
    ("essa_measure_latency", C.c_int, [_P, _dp]),
    ("essa_flush_l2", C.c_int, [_P]),
]

_lib = None


def lib():
    """Load the shared library (never builds it: run `python -m essa_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m essa_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, ptype):
    return None if a is None else a.ctypes.data_as(ptype)


def ensemble_factor(seed, rel_sigma, copy, body, component):
    return lib().essa_ensemble_factor(seed, rel_sigma, copy, body, component)


def shard_range(n, rank, world):
    a, b = C.c_int(0), C.c_int(0)
    rc = lib().essa_shard_range(n, rank, world, C.byref(a), C.byref(b))
    if rc != 0:
        raise EssaError(rc, "bad shard arguments")
    return a.value, b.value


def nccl_unique_id():
    buf = np.zeros(NCCL_ID_BYTES, dtype=np.uint8)
    rc = lib().essa_nccl_unique_id(_ptr(buf, _u8p))
    if rc != 0:
        raise EssaError(rc, (lib().essa_last_error(None) or b"").decode())
    return buf


class Context:
    """One world resident on one B200 (essa_ctx)."""

    _DOUBLE_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "gf", "max_attraction", "ap", "pe", "ap_vel", "pe_vel")

    def __init__(self, device=0, capacity=0):
        self._L = lib()
        h = _P()
        rc = self._L.essa_ctx_create(device, capacity, C.byref(h))
        if rc != 0:
            raise EssaError(rc, (self._L.essa_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.essa_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise EssaError(rc, (self._L.essa_last_error(self._h) or b"").decode())
        return rc

    # ---- state ----
    def upload(self, x, y, z, vx, vy, vz, gf, creation_date=None, deletion_date=None, deleted=None, most=None,
               ap=None, pe=None, ap_vel=None, pe_vel=None):
        keep = {}
        b = Bodies()
        for name, val in (("x", x), ("y", y), ("z", z), ("vx", vx), ("vy", vy), ("vz", vz), ("gf", gf), ("ap", ap),
                          ("pe", pe), ("ap_vel", ap_vel), ("pe_vel", pe_vel)):
            if val is not None:
                keep[name] = _arr(val, np.float64)
                setattr(b, name, _ptr(keep[name], _dp))
        for name, val, dt, pt in (("creation_date", creation_date, np.int64, _i64p), ("deletion_date", deletion_date, np.int64, _i64p),
                                  ("deleted", deleted, np.uint8, _u8p), ("most", most, np.int32, _i32p)):
            if val is not None:
                keep[name] = _arr(val, dt)
                setattr(b, name, _ptr(keep[name], pt))
        n = len(keep["x"])
        for k, v in keep.items():
            assert len(v) == n, k
        self._ck(self._L.essa_upload(self._h, n, C.byref(b)))

    def download(self, fields=("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "gf", "most", "max_attraction", "ap", "pe",
                               "ap_vel", "pe_vel"), out=None):
        """`out`: a dict of arrays from an earlier call to write into (what a host program with its own SoA buffers does;
        fresh arrays cost a page fault per 4 KiB on first touch)."""
        n = self.n
        out = {} if out is None else out
        b = Bodies()

        def buf(f, dtype):
            a = out.get(f)
            if a is None or a.dtype != dtype or a.shape != (n,) or not a.flags.c_contiguous:
                a = out[f] = np.zeros(n, dtype=dtype)
            return a

        for f in fields:
            if f in self._DOUBLE_FIELDS:
                setattr(b, f, _ptr(buf(f, np.float64), _dp))
            elif f == "most":
                b.most = _ptr(buf(f, np.int32), _i32p)
            elif f in ("creation_date", "deletion_date"):
                setattr(b, f, _ptr(buf(f, np.int64), _i64p))
            elif f == "deleted":
                b.deleted = _ptr(buf(f, np.uint8), _u8p)
            else:
                raise KeyError(f)
        self._ck(self._L.essa_download(self._h, C.byref(b)))
        return out

    def upload_shard(self, x, y, z, vx, vy, vz):
        """Positions / velocities of this rank's own bodies from full-length arrays (only the slice is read)."""
        b = Bodies()
        keep = [_arr(a, np.float64) for a in (x, y, z, vx, vy, vz)]
        for name, a in zip(("x", "y", "z", "vx", "vy", "vz"), keep):
            assert len(a) == self.n, name
            setattr(b, name, _ptr(a, _dp))
        self._ck(self._L.essa_upload_shard(self._h, C.byref(b)))

    def download_shard(self, fields, out):
        """This rank's own bodies into the slice [i0, i1) of the full-length arrays of `out` (created if missing)."""
        n = self.n
        b = Bodies()
        for f in fields:
            if f in self._DOUBLE_FIELDS:
                if f not in out:
                    out[f] = np.zeros(n, dtype=np.float64)
                setattr(b, f, _ptr(out[f], _dp))
            elif f == "most":
                if f not in out:
                    out[f] = np.zeros(n, dtype=np.int32)
                b.most = _ptr(out[f], _i32p)
            else:
                raise KeyError(f)
        self._ck(self._L.essa_download_shard(self._h, C.byref(b)))
        return out

    @property
    def n(self):
        return self._L.essa_body_count(self._h)

    @property
    def date(self):
        return self._L.essa_get_date(self._h)

    @date.setter
    def date(self, v):
        self._ck(self._L.essa_set_date(self._h, int(v)))

    # ---- stepping ----
    @staticmethod
    def opts(dt, forward_simulated=False, offset_trails=True, record=False, track_most_attracting=False, exact_order=False,
             two_pass=False, kernel=KERNEL_AUTO, eps2=0.0):
        return StepOpts(int(dt), int(forward_simulated), int(offset_trails), int(record), int(track_most_attracting),
                        int(exact_order), int(two_pass), int(kernel), float(eps2))

    def step(self, steps, dt=None, opts=None, **kw):
        o = opts if opts is not None else self.opts(dt, **kw)
        self._ck(self._L.essa_step(self._h, int(steps), C.byref(o)))

    def step_async(self, steps, dt=None, opts=None, **kw):
        o = opts if opts is not None else self.opts(dt, **kw)
        self._ck(self._L.essa_step_async(self._h, int(steps), C.byref(o)))

    def set_forces(self, dt=1, opts=None, **kw):
        o = opts if opts is not None else self.opts(dt, **kw)
        self._ck(self._L.essa_set_forces(self._h, C.byref(o)))

    def persist_configure(self, enable=True, idle_timeout_us=0):
        self._ck(self._L.essa_persist_configure(self._h, int(enable), int(idle_timeout_us)))

    def persist_stats(self):
        out = (C.c_int64 * 6)()
        self._ck(self._L.essa_persist_stats(self._h, out))
        return dict(zip(("live", "ticks", "launches", "idle_exits", "wait_ns", "host_ns"), (int(v) for v in out)))

    def sync(self):
        self._ck(self._L.essa_sync(self._h))

    # ---- recording ----
    def record_configure(self, trail_capacity, keep=0):
        caps = _arr(trail_capacity, np.int32)
        self._ck(self._L.essa_record_configure(self._h, len(caps), _ptr(caps, _i32p), int(keep)))

    @property
    def tracked(self):
        return self._L.essa_tracked_count(self._h)

    def record_reset(self, body, history=True, trail=True):
        self._ck(self._L.essa_record_reset(self._h, body, int(history), int(trail)))

    def history(self, body):
        n = self._ck(self._L.essa_history_read(self._h, body, None))
        out = np.zeros((n, 6))
        self._ck(self._L.essa_history_read(self._h, body, _ptr(out, _dp)))
        return out

    def history_write(self, body, entries, set_first=False):
        e = _arr(entries, np.float64).reshape(-1, 6)
        self._ck(self._L.essa_history_write(self._h, body, _ptr(e, _dp), len(e), int(set_first)))

    def trail(self, body):
        meta = TrailMeta()
        self._ck(self._L.essa_trail_read(self._h, body, None, C.byref(meta)))
        verts = np.zeros((meta.capacity, 3), dtype=np.float32)
        self._ck(self._L.essa_trail_read(self._h, body, _ptr(verts, _fp), C.byref(meta)))
        return dict(capacity=meta.capacity, verts=verts, append_offset=meta.append_offset, length=meta.length,
                    offset=np.array(list(meta.offset)), last_xy_angle=meta.last_xy_angle)

    def trail_write(self, body, trail):
        meta = TrailMeta(int(trail["capacity"]), int(trail["append_offset"]), int(trail["length"]), 0,
                         (C.c_double * 3)(*[float(v) for v in trail["offset"]]), float(trail["last_xy_angle"]))
        verts = _arr(trail["verts"], np.float32)
        self._ck(self._L.essa_trail_write(self._h, body, _ptr(verts, _fp), C.byref(meta)))

    def trail_device_view(self, body):
        """Zero-copy view of the Trail ring in device memory (what a renderer maps): dict with the device pointer."""
        v = TrailView()
        self._ck(self._L.essa_trail_device_view(self._h, body, C.byref(v)))
        return dict(verts_device=v.verts_device, capacity=v.capacity, append_offset=v.append_offset, length=v.length,
                    offset=np.array(list(v.offset)))

    def snapshot(self):
        """The whole context as one flat blob (bytes)."""
        n = self._ck(self._L.essa_snapshot_size(self._h))
        buf = np.zeros(n, dtype=np.uint8)
        self._ck(self._L.essa_snapshot_write(self._h, buf.ctypes.data_as(C.c_void_p), n))
        return buf.tobytes()

    def restore(self, blob):
        buf = np.frombuffer(blob, dtype=np.uint8).copy()
        self._ck(self._L.essa_snapshot_read(self._h, buf.ctypes.data_as(C.c_void_p), len(buf)))

    def closest_enable(self, enable=True):
        self._ck(self._L.essa_closest_enable(self._h, int(enable)))

    def closest_approach(self, body, other):
        out = np.zeros(7)
        if self._ck(self._L.essa_closest_read(self._h, body, other, _ptr(out, _dp))) == 0:
            return None
        return dict(distance=out[0], this_position=out[1:4].copy(), other_object_position=out[4:7].copy())

    # ---- diagnostics ----
    def energy(self):
        ke, pe = C.c_double(0), C.c_double(0)
        mom = np.zeros(3)
        self._ck(self._L.essa_energy(self._h, C.byref(ke), C.byref(pe), _ptr(mom, _dp)))
        return ke.value, pe.value, mom

    @property
    def launches(self):
        return self._L.essa_launch_count(self._h)

    @property
    def passes(self):
        return self._L.essa_pass_count(self._h)

    @property
    def interactions(self):
        return self._L.essa_interaction_count(self._h)

    @property
    def last_step_ms(self):
        return self._L.essa_last_step_ms(self._h)

    @property
    def last_force_ms(self):
        return self._L.essa_last_force_ms(self._h)

    @property
    def last_force_passes(self):
        return self._L.essa_last_force_passes(self._h)

    @property
    def last_pass_phases(self):
        """ms of the last pair-shared force pass that was timed by itself: before the kernel, kernel, exchange, finish."""
        out = np.zeros(4)
        self._L.essa_last_pass_phases(self._h, _ptr(out, _dp))
        return dict(zip(("before_kernel_ms", "pair_kernel_ms", "exchange_ms", "finish_ms"), out.tolist()))

    @property
    def last_step_breakdown(self):
        out = np.zeros(4)
        self._L.essa_last_step_breakdown(self._h, _ptr(out, _dp))
        return dict(zip(("step_ms", "force_ms", "pre_ms", "post_ms"), out.tolist()))

    def measure_fp64_peak(self, ms=200.0):
        v = self._L.essa_measure_fp64_peak(self._h, float(ms))
        if v < 0:
            raise EssaError(-2, "fp64 peak kernel failed")
        return v

    def measure_rsqrt_error(self):
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self._L.essa_measure_rsqrt_error(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_latency(self):
        out = np.zeros(8)
        self._ck(self._L.essa_measure_latency(self._h, _ptr(out, _dp)))
        return dict(zip(("dfma", "dadd", "dmul", "rsqrt_seed", "lds", "shfl_f64", "dfma_issue_4chains", "dfma_issue_8chains"), out.tolist()))

    def flush_l2(self):
        self._ck(self._L.essa_flush_l2(self._h))

    # ---- ensemble ----
    def ensemble_init(self, copies, dt, seed=7, rel_sigma=1e-6, first_copy=0, exact_order=False, closest_approaches=False,
                      approach_to=0, trail_every=0, trail_capacity=0):
        o = EnsembleOpts(int(copies), int(first_copy), int(dt), int(seed), float(rel_sigma), int(exact_order),
                         int(closest_approaches), int(approach_to), int(trail_every), int(trail_capacity))
        self._ck(self._L.essa_ensemble_init(self._h, C.byref(o)))
        self._ens = (int(copies), self.n, int(trail_capacity))

    def ensemble_approaches(self, first, count):
        """(this_pos, other_pos): [count][N][3] positions of body k and of the designated body at their closest approach."""
        n = self._ens[1]
        a, b = np.zeros((count, n, 3)), np.zeros((count, n, 3))
        self._ck(self._L.essa_ensemble_read_approaches(self._h, first, count, _ptr(a, _dp), _ptr(b, _dp)))
        return a, b

    def ensemble_trail(self, first, count):
        """(xyz [count][capacity][3], lengths [count]): the designated body's trail per copy."""
        cap = self._ens[2]
        xyz, ln = np.zeros((count, cap, 3)), np.zeros(count, dtype=np.int32)
        self._ck(self._L.essa_ensemble_read_trail(self._h, first, count, _ptr(xyz, _dp), _ptr(ln, _i32p)))
        return xyz, ln

    def ensemble_step(self, steps):
        self._ck(self._L.essa_ensemble_step(self._h, int(steps)))

    def ensemble_read(self, first, count, min_distance=False):
        n = self._ens[1]
        arrs = [np.zeros((count, n)) for _ in range(6)]
        md = np.zeros((count, n)) if min_distance else None
        self._ck(self._L.essa_ensemble_read(self._h, first, count, *[_ptr(a, _dp) for a in arrs], _ptr(md, _dp)))
        out = dict(zip(("x", "y", "z", "vx", "vy", "vz"), arrs))
        if min_distance:
            out["min_distance"] = md
        return out

    # ---- sharding ----
    def shard_attach(self, rank, world, unique_id):
        uid = _arr(unique_id, np.uint8)
        assert len(uid) == NCCL_ID_BYTES
        self._ck(self._L.essa_shard_attach(self._h, rank, world, _ptr(uid, _u8p)))

    @property
    def shard(self):
        return self._L.essa_shard_rank(self._h), self._L.essa_shard_world(self._h)
