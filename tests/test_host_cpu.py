"""CPU-only checks (`-m "not gpu"`): the C-ABI library loads and exports every symbol that
include/*.h declares; the product's .essa loader and orbit builders produce exactly the oracle's
initial conditions; host-side facade logic (object list, dates, dirty tracking, errors); and the
product fails loudly -- no CPU fallback -- when asked to step without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from essa_b200 import capi, synth, world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD_FILES = ["2body.essa", "8body.essa", "light_test.essa", "trail_test.essa", "solar.essa", "jupiter_system.essa"]


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(essa_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["essa_b200.h", "essa_world.h"])
def test_library_exports_every_declared_symbol(header):
    L = ctypes.CDLL(capi.LIB_PATH)
    names = _declared_functions(header)
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    bound = {s[0] for s in capi.SYMBOLS} | {s[0] for s in world.WORLD_SYMBOLS}
    declared = set(_declared_functions("essa_b200.h")) | set(_declared_functions("essa_world.h"))
    assert declared <= bound, sorted(declared - bound)
    assert capi.lib().essa_abi_version() == 2


def test_no_cpu_fallback():
    """In a container without a GPU the product must refuse, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.EssaError) as e:
        capi.Context(0)
    assert e.value.code == -2
    w = world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", "2body.essa"))
    with pytest.raises(capi.EssaError):
        w.update(1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "essa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "essa_oracle" not in text, f


@pytest.mark.parametrize("fname", WORLD_FILES)
def test_loader_matches_oracle_bitwise(worlds_dir, fname):
    """src/ConfigLoader.cpp + src/Object.cpp:309-383: same list, same bits, same constructor-time rings."""
    path = os.path.join(worlds_dir, fname)
    ow = oracle.World.load(path)
    pw = world.World.load(path)
    assert len(pw) == len(ow)
    assert pw.names() == ow.names()
    a, b = pw.state(), ow.state()
    for k in ("pos", "vel", "gf", "appe", "most", "active"):
        assert np.array_equal(a[k], b[k]), k
    for i in range(len(ow)):
        pr = pw.props(i)
        assert pr["radius"] == ow.radius(i)
        assert pr["orbit_period"] == ow.orbit_len(i)
        assert pr["eccentrity"] == ow.eccentrity(i) or (np.isnan(pr["eccentrity"]) and np.isnan(ow.eccentrity(i)))
        ta, tb = pw.trail(i), ow.trail(i)
        assert ta["capacity"] == tb["capacity"] == pr["trail_capacity"]
        assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"]) == (3 % ta["capacity"] or 1, 3)
        assert np.array_equal(ta["verts"], tb["verts"])
        assert ta["last_xy_angle"] == tb["last_xy_angle"]
        assert np.array_equal(pw.history(i), ow.history(i))


def test_loader_sanity_anchors(worlds_dir):
    """gross-error anchors for solar.essa (not golden values)."""
    pw = world.World.load(os.path.join(worlds_dir, "solar.essa"))
    names = pw.names()
    st = pw.state()
    earth, moon, nep = names.index("Earth"), names.index("Moon"), names.index("Neptune")
    assert st["pos"][earth][0] == pytest.approx(-1.505875e11, rel=1e-6)
    assert st["vel"][earth][1] == pytest.approx(-29689.98, rel=1e-6)
    assert st["pos"][moon] == pytest.approx([-1.507061e11, 3.651106e8, 0], rel=1e-6)
    assert st["vel"][nep][:2] == pytest.approx([5298.89, 1196.02], rel=1e-5)
    assert pw.props(nep)["trail_capacity"] == 120442
    assert pw.props(names.index("Sun"))["trail_capacity"] == 1200
    assert pw.date == capi.START_DATE and pw.simulation_seconds_per_tick == 43200


@pytest.mark.parametrize("text,needle", [
    ("planet mass=1 name=A", "Expected ';'"),
    ("moon mass=1;", "Invalid statement keyword"),
    ("planet mass=abc name=A;", "Invalid number"),
    ("planet mass=1 radius=3_parsec name=A;", "Invalid unit"),
    ("orbiting_planet mass=1 name=B around=A apoapsis=1_AU periapsis=1_AU;", "Invalid planet for orbiting around"),
    ("planet mass=1 name=A; orbiting_planet mass=1 name=B around=A;", "apoapsis+periapsis or major_axis+eccentrity"),
    ("planet mass 1;", "Expected '='"),
    ("light_source Nobody;", "Invalid planet for light source"),
])
def test_loader_errors(text, needle):
    with pytest.raises(capi.EssaError) as e:
        world.World.loads(text)
    assert needle in str(e.value)
    ow = oracle.World()
    err = ctypes.create_string_buffer(512)
    tmp = os.path.join(ROOT, "tests", "_tmp_err.essa")
    with open(tmp, "w") as f:
        f.write(text)
    try:
        rc = ow._L.ow_load_essa_file(ow._h, tmp.encode(), 0, err, 512)
    finally:
        os.remove(tmp)
    assert rc != 0 and needle in err.value.decode()


def test_pyssa_surface_on_host():
    w = world.World()
    sun = w.add_object(mass=2e30, radius=7e8, name="Sun", color=(255, 255, 0))
    o = world.Object(mass=6e24, radius=6.4e6, pos=(1.5e11, 0, 0), vel=(0, 29800.0, 0), name="Earth", period=365)
    assert o.pos == (1.5e11, 0.0, 0.0)
    with pytest.raises(ReferenceError):
        o._loc()
    w.add_object(o)
    assert len(w) == 2 and w.names() == ["Sun", "Earth"]
    e = w.get_object_by_name("Earth")
    assert e.pos == (1.5e11, 0.0, 0.0) and e.vel == (0.0, 29800.0, 0.0)
    assert e.mass == pytest.approx(6e24, rel=1e-15) and e.radius == 6.4e6 and e.color == (255, 255, 255)
    assert sun.color == (255, 255, 0)
    e.pos = (1.0e11, 1.0, 2.0)
    e.vel = (1.0, 2.0, 3.0)
    e.mass = 7e24
    e.name = "Terra"
    e.radius = 1.0
    e.color = (1, 2, 3)
    assert w.get_object_by_name("Earth") is None
    t = w.get_object_by_name("Terra")
    assert t.pos == (1.0e11, 1.0, 2.0) and t.vel == (1.0, 2.0, 3.0) and t.radius == 1.0 and t.color == (1, 2, 3)
    assert t.mass == pytest.approx(7e24, rel=1e-15)
    # Object::attraction (src/Object.cpp:53-58): other.gf / r^2 along (this - other)
    a = t.attraction(sun)
    assert a[0] == pytest.approx(2e30 * capi.GRAVITY / 1e22, rel=1e-9) and a[0] > 0
    w.simulation_seconds_per_tick = 600
    assert w.simulation_seconds_per_tick == 600
    w.delete_object(0)
    assert w.names() == ["Terra"]
    with pytest.raises(ReferenceError):
        sun.pos
    assert t.pos == (1.0e11, 1.0, 2.0)
    assert world.World.loads("").names() == []
    with pytest.raises(capi.EssaError):
        w.update(0)


def test_empty_world_update_moves_only_the_clock():
    """World::update on an empty list: update_history_and_date still runs (src/World.cpp:59-84); no device needed."""
    w = world.World()
    w.simulation_seconds_per_tick = 600
    w.update(7)
    assert w.date == capi.START_DATE + 7 * 600 and len(w) == 0
    w.update(-3)
    assert w.date == capi.START_DATE + 4 * 600
    ow = oracle.World()
    ow.dt = 600
    # (the reference dereferences m_object_list.back() on an empty list here; the oracle is not asked to)


def test_host_clone_for_forward_simulation(worlds_dir):
    pw = world.World.load(os.path.join(worlds_dir, "solar.essa"))
    pw.mark_deleted(3)
    fw = pw.clone_for_forward_simulation()
    assert fw.is_forward_simulated and not pw.is_forward_simulated
    assert len(fw) == len(pw) - 1
    assert fw.simulation_seconds_per_tick == 43200
    a, b = fw.state(), pw.state()
    keep = [i for i in range(len(pw)) if i != 3]
    assert np.array_equal(a["pos"], b["pos"][keep]) and np.array_equal(a["gf"], b["gf"][keep])
    assert all(fw.props(i)["trail_capacity"] == 1000 for i in range(len(fw)))


def test_synthetic_worlds_are_deterministic():
    a, b = synth.plummer(4096, seed=3), synth.plummer(4096, seed=3)
    for k in a:
        assert np.array_equal(a[k], b[k])
    g = synth.galaxy_collision(2048)
    m = g["gf"] / capi.GRAVITY
    assert len(g["x"]) == 2048 and abs((m * g["vx"]).sum()) < 1e-6 * (m * np.abs(g["vx"])).sum()
    assert (synth.default_dt(16384), synth.default_dt(262144), synth.default_dt(1 << 20)) == (1000000, 300000, 200000)


def test_ensemble_factor_host_side():
    f = capi.ensemble_factor
    assert f(7, 1e-6, 0, 3, 2) == 1.0
    vals = np.array([f(7, 1e-6, c, b, k) for c in range(1, 200) for b in range(3) for k in range(6)])
    n = (vals - 1.0) / 1e-6
    assert abs(n.mean()) < 0.1 and 0.9 < n.std() < 1.1 and np.abs(n).max() < 6
    assert f(7, 1e-6, 5, 1, 2) == f(7, 1e-6, 5, 1, 2) != f(8, 1e-6, 5, 1, 2)
