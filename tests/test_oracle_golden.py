"""Pin the oracle against the only golden vectors the reference ships for this path:
tests/accuracy.ods (sheets data-600 and data-60000), extracted by tools/extract_accuracy_ods.py
into tests/golden/accuracy_*.csv.gz.

The traces were produced by an older kick(dt)->drift(dt) build, so they are
replayed with the oracle's `legacy_euler` switch; what they pin is the pair-force law
(src/Object.cpp:64-79), G = 6.67408e-11, gf = m*G and the `date += dt` before integrating cadence.
Every printed field (6 significant figures, C++ ostream default) must match exactly.
"""
import gzip
import os

import pytest

import oracle


def _fmt(v):
    s = "%.6g" % v
    if s == "-0":
        s = "0"
    return s


def _rows(path):
    with gzip.open(path, "rt") as f:
        header = f.readline().strip().split(",")
        assert header == ["time", "name", "px", "py", "pz", "vx", "vy", "vz"]
        for line in f:
            yield line.strip().split(",")


@pytest.mark.parametrize("sheet,dt,steps", [("data-600", 600, 7201), ("data-60000", 60000, 73)])
def test_accuracy_ods_replay(golden_dir, sheet, dt, steps):
    w = oracle.World()
    w.add_object(2e28, 1e5, (-5e9, 0, 0), (0, 5000, 0), "Sun1")
    w.add_object(2e28, 1e5, (5e9, 0, 0), (0, -5000, 0), "Sun2")
    w.dt = dt
    w.set_flags(legacy_euler=True)
    rows = list(_rows(os.path.join(golden_dir, "accuracy_%s.csv.gz" % sheet)))
    assert len(rows) == 2 * steps
    mismatches = 0
    fields = 0
    for s in range(steps):
        w.update(1)
        st = w.state()
        for b in range(2):
            row = rows[2 * s + b]
            assert int(row[0]) == w.date
            assert row[1] == w.name(b)
            got = [_fmt(v) for v in (*st["pos"][b], *st["vel"][b])]
            for g, e in zip(got, row[2:]):
                fields += 1
                if g != e:
                    mismatches += 1
    assert fields == 6 * 2 * steps
    assert mismatches == 0


def test_kdk_does_not_reproduce_the_trace(golden_dir):
    """Documents that the trace pre-dates the KDK integrator (it must NOT match in KDK mode)."""
    w = oracle.World()
    w.add_object(2e28, 1e5, (-5e9, 0, 0), (0, 5000, 0), "Sun1")
    w.add_object(2e28, 1e5, (5e9, 0, 0), (0, -5000, 0), "Sun2")
    w.dt = 60000
    rows = list(_rows(os.path.join(golden_dir, "accuracy_data-60000.csv.gz")))
    bad = 0
    for s in range(73):
        w.update(1)
        st = w.state()
        for b in range(2):
            got = [_fmt(v) for v in (*st["pos"][b], *st["vel"][b])]
            bad += sum(g != e for g, e in zip(got, rows[2 * s + b][2:]))
    assert bad > 100


def test_wrong_gravity_constant_breaks_the_trace(golden_dir):
    """G is pinned: CODATA-2014 6.67408e-11 matches, 6.6743e-11 does not (scale gf accordingly)."""
    w = oracle.World()
    w.add_object(2e28, 1e5, (-5e9, 0, 0), (0, 5000, 0), "Sun1")
    w.add_object(2e28, 1e5, (5e9, 0, 0), (0, -5000, 0), "Sun2")
    for b in range(2):
        w.set_object_state(b, gf=2e28 * 6.6743e-11)
    w.dt = 600
    w.set_flags(legacy_euler=True)
    rows = list(_rows(os.path.join(golden_dir, "accuracy_data-600.csv.gz")))
    bad = 0
    for s in range(600):
        w.update(1)
        st = w.state()
        for b in range(2):
            got = [_fmt(v) for v in (*st["pos"][b], *st["vel"][b])]
            bad += sum(g != e for g, e in zip(got, rows[2 * s + b][2:]))
    assert bad > 100
