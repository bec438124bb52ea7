"""The oracle's History and Trail against the REFERENCE's own src/History.cpp and src/Trail.cpp (oracle/_ref, compiled from the
reference tree where it lies; the un-vendored EssaUtil / EssaGUI headers are replaced by oracle/ref_stubs, which is where the
three arithmetic assumptions U1-U3 of DESIGN.md live).  Same inputs, identical bits: ring contents, append offset / length
bookkeeping, the decimation rule, offset re-basing, wrap-around, reset, and both directions of the History."""
import os

import numpy as np
import pytest

import oracle
from oracle import ref

IMPLS = ref.impls()
pytestmark = pytest.mark.skipif(IMPLS is None, reason="oracle/_ref/libessa_ref.so not built (needs /root/reference)")


def _same_trail(a, b):
    ra, rb = a.read(), b.read()
    assert (ra["capacity"], ra["append_offset"], ra["length"]) == (rb["capacity"], rb["append_offset"], rb["length"])
    assert np.array_equal(ra["verts"].view(np.uint32), rb["verts"].view(np.uint32))
    assert np.array_equal(ra["offset"], rb["offset"]) and ra["last_xy_angle"] == rb["last_xy_angle"]


@pytest.mark.parametrize("cap", [1, 3, 7, 500, 736])
def test_trail_matches_the_reference_on_an_orbit_with_wrap_and_offsets(cap):
    R, O = IMPLS
    a, b = ref.Trail(R, cap), ref.Trail(O, cap)
    _same_trail(a, b)
    rng = np.random.default_rng(cap)
    au = oracle.AU
    # an eccentric, precessing orbit sampled finely enough that most pushes are decimated, for several revolutions (wraps)
    n = 4 * max(cap, 40) * 3
    t = np.linspace(0, 8 * np.pi, n)
    pts = np.stack([au * (1 + 0.3 * np.cos(1.1 * t)) * np.cos(t), au * (1 + 0.3 * np.cos(1.1 * t)) * np.sin(t), 0.05 * au * np.sin(3 * t)], axis=1)
    pts += rng.normal(0, 1e3, pts.shape)
    for i, p in enumerate(pts):
        if i == n // 3:       # the most attracting object changes: every vertex is re-based, then the offset follows
            off = rng.normal(0, 0.2 * au, 3)
            for tr in (a, b):
                tr.recalculate_with_offset(off)
        if i == n // 2:
            off2 = rng.normal(0, 0.1 * au, 3)
            for tr in (a, b):
                tr.set_offset(off2)
        if i == (2 * n) // 3:
            for tr in (a, b):
                tr.reset()
        if i % 97 == 5:
            for tr in (a, b):
                tr.change_current(p * 1.0001)
        a.push_back(p)
        b.push_back(p)
        if i % 61 == 0:
            _same_trail(a, b)
    _same_trail(a, b)


def test_trail_matches_the_reference_on_the_oracles_own_solar_system(worlds_dir):
    """Positions as World::update produces them (relative to the most attracting object, as Object::update feeds the trail)."""
    R, O = IMPLS
    ow = oracle.World.load(os.path.join(worlds_dir, "solar.essa"))
    ow.dt = 43200
    names = ow.names()
    earth, moon = names.index("Earth"), names.index("Moon")
    a, b = ref.Trail(R, 500), ref.Trail(O, 500)
    for _ in range(1500):
        ow.update(1)
        st = ow.state()
        p = st["pos"][moon] - st["pos"][earth]
        a.push_back(p)
        b.push_back(p)
    _same_trail(a, b)
    # ... and the oracle World's own trail of the Moon is that very trail (same class inside the World)
    wt = ow.trail(moon)
    assert wt["length"] <= 500


def test_history_matches_the_reference_in_both_directions():
    R, O = IMPLS
    rng = np.random.default_rng(2)
    first = rng.normal(size=6)
    a, b = ref.History(R, 50, first), ref.History(O, 50, first)

    def same():
        (ea, sa), (eb, sb) = a.read(), b.read()
        assert sa == sb and np.array_equal(ea, eb)

    same()
    for i in range(400):
        e = rng.normal(size=6)
        r = rng.random()
        if r < 0.6:
            x, y = a.move_forward(e), b.move_forward(e)
        elif r < 0.97:
            x, y = a.move_backward(e), b.move_backward(e)
        else:
            a.reset()
            b.reset()
            x = y = None
        assert (x is None) == (y is None) and (x is None or np.array_equal(x, y))
        same()
    # the only branch World::update reaches: forward pushes on a ring of 1000 (src/Object.cpp:34)
    a, b = ref.History(R, 1000, first), ref.History(O, 1000, first)
    for i in range(2300):
        e = rng.normal(size=6)
        assert a.move_forward(e) is None and b.move_forward(e) is None
    same()
    assert len(a.read()[0]) == 1000
