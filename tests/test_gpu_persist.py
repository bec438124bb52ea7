"""The persistent resident kernels (essa_persist_configure; resident_quad.cu K3q for up to 32 bodies and resident_cluster.cu K3x
for 33 .. 128, with a mailbox in mapped pinned memory).

What the reference's caller does -- SimulationView::update runs World::update(iterations) once per GUI tick and
reads the objects back (src/SimulationView.cpp:374-397, 10 iterations by default) -- is served by a kernel that
stays on its SMs between calls.  The mode must be invisible apart from its speed, so every test here runs the
same sequence of calls on two worlds, one with the mode on and one with one launch per call, and demands
BIT-IDENTICAL state, History rings and Trail vertices; the one-launch path itself is checked against the oracle in
test_gpu_resident.py / test_gpu_small_worlds.py, and once more here.
"""
import os
import time

import numpy as np
import pytest

import oracle
from essa_b200 import capi, world

pytestmark = pytest.mark.gpu


def _two(worlds_dir, fname="solar.essa", dt=600, idle_us=0):
    path = os.path.join(worlds_dir, fname)
    a, b = world.World.load(path), world.World.load(path)
    for w in (a, b):
        w.simulation_seconds_per_tick = dt
        w.set_flags(offset_trails=True)
    a.set_persistent(True, idle_us)
    b.set_persistent(False)
    return a, b


def _same(a, b, rings=False):
    sa, sb = a.state(), b.state()
    assert a.date == b.date
    for k in ("pos", "vel", "acc", "maxattr", "appe", "most", "active"):
        assert np.array_equal(sa[k], sb[k]), k
    if rings:
        for i in range(len(a)):
            assert np.array_equal(a.history(i), b.history(i)), i
            ta, tb = a.trail(i), b.trail(i)
            assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"]), i
            assert np.array_equal(ta["verts"][:ta["length"]], tb["verts"][:tb["length"]]), i
            assert np.array_equal(ta["offset"], tb["offset"]), i
            assert ta["last_xy_angle"] == tb["last_xy_angle"], i


@pytest.mark.parametrize("fname,dt", [("solar.essa", 600), ("8body.essa", 600), ("2body.essa", 600), ("trail_test.essa", 3600),
                                      ("solar.essa", 43200), ("jupiter_system.essa", 600), ("jupiter_system.essa", 7200)])
def test_gui_cadence_is_bit_identical_to_one_launch_per_call(worlds_dir, fname, dt):
    a, b = _two(worlds_dir, fname, dt)
    for tick in range(120):
        k = 10 if tick % 7 else 1 + tick % 5
        a.update(k)
        b.update(k)
        _same(a, b)
    st = a.persist_stats()
    assert st["launches"] >= 1 and st["ticks"] + st["launches"] == 120, st
    assert b.persist_stats()["launches"] == 0
    _same(a, b, rings=True)       # reading the rings retires the kernel ...
    for _ in range(5):            # ... and the next tick starts another
        a.update(10)
        b.update(10)
    _same(a, b, rings=True)
    a.close()
    b.close()


def test_persistent_ticks_match_the_oracle(worlds_dir):
    path = os.path.join(worlds_dir, "solar.essa")
    ow = oracle.World.load(path)
    pw = world.World.load(path)
    ow.dt = 600
    pw.simulation_seconds_per_tick = 600
    pw.set_flags(offset_trails=True)
    for _ in range(40):
        ow.update(10)
        pw.update(10)
    assert pw.persist_stats()["ticks"] >= 30
    a, b = pw.state(), ow.state()
    assert np.array_equal(a["most"], b["most"])
    rel = np.abs(a["pos"] - b["pos"]).max(axis=1) / np.abs(b["pos"]).max(axis=1)
    assert rel.max() <= 1e-12 * np.sqrt(400)
    for i in range(len(ow)):
        ta, tb = pw.trail(i), ow.trail(i)
        assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"]), i
        assert pw.history(i).shape == ow.history(i).shape
    pw.close()


def test_host_mutation_between_ticks_retires_the_kernel(worlds_dir):
    a, b = _two(worlds_dir)
    for w in (a, b):
        w.update(10)
        w.update(10)
    for w in (a, b):
        earth = w.get_object_by_name("Earth")
        p = np.array(earth.pos)
        earth.pos = tuple(p * (1 + 1e-3))
        earth.vel = tuple(np.array(earth.vel) * 0.999)
    for _ in range(3):
        a.update(10)
        b.update(10)
        _same(a, b)
    for w in (a, b):  # delete an object, then add one: the list changes under a resident kernel
        w.delete_object(w.names().index("Mars"))
        w.update(10)
        w.add_object(world.Object(mass=1e22, radius=1e6, pos=(3e11, 1e10, 0.0), vel=(0.0, 2e4, 0.0), name="Probe"))
        w.update(10)
        w.update(10)
    _same(a, b, rings=True)
    # time runs backwards (another launch configuration), then forwards again
    for k in (-10, -10, 10, 10):
        a.update(k)
        b.update(k)
        _same(a, b)
    a.simulation_seconds_per_tick = b.simulation_seconds_per_tick = 3600
    for _ in range(3):
        a.update(10)
        b.update(10)
    _same(a, b, rings=True)
    assert a.persist_stats()["launches"] >= 5
    a.close()
    b.close()


def test_idle_timeout_and_doorbell_race(worlds_dir):
    """A kernel nobody rings retires by itself; a doorbell that races with the timeout is never lost."""
    a, b = _two(worlds_dir, idle_us=150)
    rng = np.random.default_rng(5)
    for tick in range(300):
        a.update(10)
        b.update(10)
        if tick % 3 == 0:
            # sleep around the timeout so that some doorbells land while the kernel is deciding to leave
            t = 150e-6 * (0.5 + rng.random())
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < t:
                pass
    _same(a, b, rings=True)
    st = a.persist_stats()
    assert st["ticks"] + st["launches"] == 300, st
    assert st["launches"] >= 2, st  # some sleeps were longer than the timeout
    time.sleep(0.01)
    a.update(10)
    b.update(10)
    _same(a, b)
    a.close()
    b.close()


def test_c_abi_async_tick_and_download(worlds_dir):
    """essa_step_async + essa_download through the mailbox; a download that asks for fields the mailbox does not carry
    (gf, dates) retires the kernel and takes the ordinary path; other entry points in between do the same."""
    pw = world.World.load(os.path.join(worlds_dir, "solar.essa"))
    st = pw.state()
    caps = [pw.props(i)["trail_capacity"] for i in range(len(pw))]
    pw.close()
    n = len(caps)
    ctxs = []
    for persist in (True, False):
        c = capi.Context(0)
        c.persist_configure(persist)
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
                 creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
        c.record_configure(caps)
        ctxs.append(c)
    fast = ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "most", "max_attraction", "ap", "pe", "ap_vel", "pe_vel")
    for tick in range(60):
        outs = []
        for c in ctxs:
            c.step_async(10, 600, record=True)
            outs.append(c.download(fast if tick % 10 else fast + ("gf", "creation_date", "deleted")))
        for k in outs[0]:
            assert np.array_equal(outs[0][k], outs[1][k]), (tick, k)
        if tick == 25:
            e = [c.energy() for c in ctxs]   # a kernel launch on the same context: retires the resident one first
            assert e[0][0] == e[1][0] and e[0][1] == e[1][1]
        if tick == 40:
            for c in ctxs:
                c.step(7, 600, record=True)  # synchronous flavour
    s = ctxs[0].persist_stats()
    assert s["ticks"] >= 40 and s["launches"] >= 7, s
    assert ctxs[0].date == ctxs[1].date
    for i in range(n):
        assert np.array_equal(ctxs[0].history(i), ctxs[1].history(i))
        ta, tb = ctxs[0].trail(i), ctxs[1].trail(i)
        assert np.array_equal(ta["verts"], tb["verts"]) and ta["length"] == tb["length"]
    for c in ctxs:
        c.close()


def test_exact_order_and_masked_calls_do_not_use_the_mailbox(worlds_dir):
    path = os.path.join(worlds_dir, "solar.essa")
    pw = world.World.load(path)
    pw.simulation_seconds_per_tick = 600
    pw.set_flags(offset_trails=True, exact_order=True)
    for _ in range(5):
        pw.update(10)
    st = pw.persist_stats()
    assert (st["live"], st["ticks"], st["launches"], st["idle_exits"]) == (0, 0, 0, 0)
    pw.close()
    # an object whose deletion date is crossed during a tick: that tick (and only that one) takes the one-shot kernel
    a, b = _two(worlds_dir)
    for w in (a, b):
        w.update(10)
        w.mark_deleted(w.names().index("Venus"))
    for _ in range(4):
        a.update(10)
        b.update(10)
        _same(a, b)
    _same(a, b, rings=True)
    a.close()
    b.close()
