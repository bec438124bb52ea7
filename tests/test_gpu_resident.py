"""GPU parity of the resident kernel (K3), the recording (K6), the World facade and the ensemble
kernel (K4) against the CPU oracle.

exact_order mode: BIT-IDENTICAL positions, velocities, accelerations, most-attracting links, ap/pe,
History rings and Trail vertices after K steps of the reference's .essa worlds (SURVEY.md 8c i).
The one field that may differ is Trail::m_last_xy_angle (CUDA's atan2 vs glibc's, both < 2 ulp);
it feeds a threshold test, so the vertices themselves must still be equal.

fast mode: relative tolerance 1e-12 * sqrt(K) on positions / velocities (SURVEY.md 8c iii).
"""
import os

import numpy as np
import pytest

import oracle
from essa_b200 import capi, world

pytestmark = pytest.mark.gpu

WORLDS = {"2body.essa": (600, 400), "8body.essa": (600, 300), "trail_test.essa": (3600, 400), "light_test.essa": (600, 50),
          "solar.essa": (600, 300), "jupiter_system.essa": (600, 60)}


def _pair(worlds_dir, fname, dt, exact=True, offset_trails=True):
    path = os.path.join(worlds_dir, fname)
    ow = oracle.World.load(path)
    pw = world.World.load(path)
    ow.dt = dt
    pw.simulation_seconds_per_tick = dt
    ow.set_flags(offset_trails=offset_trails)
    pw.set_flags(offset_trails=offset_trails, exact_order=exact)
    return ow, pw


def _assert_same_state(pw, ow, exact=True, rtol=0.0):
    a, b = pw.state(), ow.state()
    assert pw.date == ow.date
    assert np.array_equal(a["active"], b["active"])
    assert np.array_equal(a["most"], b["most"])
    for k in ("pos", "vel", "acc", "maxattr", "appe"):
        if exact:
            assert np.array_equal(a[k], b[k]), (k, np.abs(a[k] - b[k]).max())
        else:
            num = np.abs(a[k] - b[k]).reshape(len(a[k]), -1).max(axis=1)
            den = np.abs(b[k]).reshape(len(b[k]), -1).max(axis=1) + 1e-300
            tol = rtol * (1e3 if k in ("acc", "maxattr") else 1.0)
            assert (num / den).max() <= tol, (k, (num / den).max())


def _assert_same_rings(pw, ow, exact=True):
    for i in range(len(ow)):
        ha, hb = pw.history(i), ow.history(i)
        assert ha.shape == hb.shape
        ta, tb = pw.trail(i), ow.trail(i)
        assert (ta["capacity"], ta["append_offset"], ta["length"]) == (tb["capacity"], tb["append_offset"], tb["length"]), i
        if exact:
            assert np.array_equal(ha, hb), i
            assert np.array_equal(ta["verts"][:ta["length"]], tb["verts"][:tb["length"]]), i
            assert np.array_equal(ta["offset"], tb["offset"]), i
            assert ta["last_xy_angle"] == pytest.approx(tb["last_xy_angle"], rel=0, abs=4 * np.spacing(abs(tb["last_xy_angle"]) + 1e-300))


@pytest.mark.parametrize("fname", list(WORLDS))
def test_exact_order_is_bit_identical(worlds_dir, fname):
    dt, steps = WORLDS[fname]
    ow, pw = _pair(worlds_dir, fname, dt)
    for chunk in (1, 10, steps - 11):   # GUI-like ticks: several update() calls
        ow.update(chunk)
        pw.update(chunk)
        _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)


def test_exact_default_tick_and_history_wrap(worlds_dir):
    """World's own default tick (43,200 s) for 1,100 steps: the 1000-entry History ring wraps, short
    trails wrap (capacity 500) and Mercury's angular decimation kicks in."""
    ow, pw = _pair(worlds_dir, "solar.essa", 43200)
    ow.update(1100)
    pw.update(1100)
    _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)
    assert pw.history(0).shape == (1000, 6)


def test_exact_without_offset_trails(worlds_dir):
    ow, pw = _pair(worlds_dir, "trail_test.essa", 3600, offset_trails=False)
    ow.update(200)
    pw.update(200)
    _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)


def test_exact_backward_steps(worlds_dir):
    ow, pw = _pair(worlds_dir, "solar.essa", 600)
    for s in (20, -35, 5):
        ow.update(s)
        pw.update(s)
        _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)


def test_exact_host_mutation_between_updates(worlds_dir):
    """GUI / Python edits between ticks: set_pos, set_vel, mark deleted, remove, add, reset_history."""
    ow, pw = _pair(worlds_dir, "solar.essa", 600)
    ow.update(7)
    pw.update(7)
    # move Mars, re-aim Venus
    ow.set_object_state(4, pos=(2.0e11, 1.0e10, 5.0e8))
    pw.set_object_state(4, pos=(2.0e11, 1.0e10, 5.0e8))
    ow.set_object_state(2, vel=(1000.0, -36000.0, 10.0))
    pw.set_object_state(2, vel=(1000.0, -36000.0, 10.0))
    ow.update(5)
    pw.update(5)
    _assert_same_state(pw, ow)
    # Object::delete_object on Jupiter: stays in the list, masked from now on
    ow.delete_object(5)
    pw.mark_deleted(5)
    ow.update(6)
    pw.update(6)
    _assert_same_state(pw, ow)
    # a new object appears (creation date = now)
    ow.add_object(1e26, 1e7, (5e11, 0, 0), (0, 12000.0, 0), "Rogue", 900)
    pw.add_object(mass=1e26, radius=1e7, pos=(5e11, 0, 0), vel=(0, 12000.0, 0), name="Rogue", period=900)
    ow.update(9)
    pw.update(9)
    _assert_same_state(pw, ow)
    # World::delete_object_by_ptr on the Moon's parent: the Moon loses its most-attracting object
    earth = ow.names().index("Earth")
    ow.remove_object(earth)
    pw.delete_object(earth)
    ow.reset_history(1)
    pw.reset_history(1)
    ow.update(11)
    pw.update(11)
    _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)


def test_exact_two_pass_and_set_forces(worlds_dir):
    ow, pw = _pair(worlds_dir, "jupiter_system.essa", 600)
    pw.set_flags(exact_order=True, two_pass=True)
    ow.set_forces()
    pw.set_forces()
    _assert_same_state(pw, ow)
    ow.update(12)
    pw.update(12)
    _assert_same_state(pw, ow)


@pytest.mark.parametrize("fname,steps", [("solar.essa", 2000), ("jupiter_system.essa", 200), ("trail_test.essa", 1000)])
def test_fast_mode_within_tolerance(worlds_dir, fname, steps):
    ow, pw = _pair(worlds_dir, fname, 600, exact=False)
    ow.update(steps)
    pw.update(steps)
    _assert_same_state(pw, ow, exact=False, rtol=1e-12 * np.sqrt(steps))
    _assert_same_rings(pw, ow, exact=False)
    # trail vertices are float32 of the same positions: equal to a few float ulps
    for i in range(len(ow)):
        ta, tb = pw.trail(i), ow.trail(i)
        n = ta["length"]
        assert np.allclose(ta["verts"][:n], tb["verts"][:n], rtol=1e-5, atol=1e-7)


def test_forward_simulation_clone_and_closest_approaches(worlds_dir):
    """World::clone_for_forward_simulation + update(ticks) (EssaCreateObject.cpp:166-189)."""
    ow, pw = _pair(worlds_dir, "solar.essa", 600)
    ow.update(3)
    pw.update(3)
    of, pf = ow.clone_for_forward_simulation(), pw.clone_for_forward_simulation()
    pf.set_flags(exact_order=True)
    assert pf.is_forward_simulated and of.is_forward_simulated
    of.update(250)
    pf.update(250)
    _assert_same_state(pf, of)
    assert pf.date == capi.START_DATE  # the clock of a forward-simulated world does not move
    for i in range(len(of)):
        assert pf.history(i).shape == (1, 6)
        ta, tb = pf.trail(i), of.trail(i)
        assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"])
        assert np.array_equal(ta["verts"][:ta["length"]], tb["verts"][:tb["length"]])
        for j in range(len(of)):
            if i == j:
                continue
            ca, cb = pf.closest_approach(i, j), of.closest_approach(i, j)
            assert ca["distance"] == cb["distance"]
            assert np.array_equal(ca["this_position"], cb["this_position"])
            assert np.array_equal(ca["other_object_position"], cb["other_object_position"])
    # the original world is untouched and still steps
    ow.update(2)
    pw.update(2)
    _assert_same_state(pw, ow)


def test_resident_and_tiled_kernels_agree(worlds_dir):
    """Same world through both kernels of the C ABI, recording on."""
    ow = oracle.World.load(os.path.join(worlds_dir, "jupiter_system.essa"))
    st = ow.state()
    caps = [ow.trail(i)["capacity"] for i in range(len(ow))]
    out = {}
    for kern in (capi.KERNEL_RESIDENT, capi.KERNEL_TILED):
        c = capi.Context(0)
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
                 creation_date=np.full(len(ow), capi.START_DATE, dtype=np.int64))
        c.record_configure(caps)
        c.step(40, 600, kernel=kern, record=True)
        out[kern] = (c.download(), [c.trail(i) for i in range(len(ow))], [c.history(i) for i in range(len(ow))])
        c.close()
    a, b = out[capi.KERNEL_RESIDENT], out[capi.KERNEL_TILED]
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.allclose(a[0][k], b[0][k], rtol=1e-11, atol=0), k
    assert np.array_equal(a[0]["most"], b[0]["most"])
    for i in range(len(ow)):
        assert a[2][i].shape == b[2][i].shape == (41, 6)
        assert (a[1][i]["append_offset"], a[1][i]["length"]) == (b[1][i]["append_offset"], b[1][i]["length"])


def _ensemble_reference(st, copy, steps, dt, seed, sigma):
    """Oracle run of one perturbed copy, initial state rebuilt with the C ABI's host-side factor."""
    n = len(st["gf"])
    w = oracle.World()
    comp = [st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2]]
    pert = [np.array([comp[q][b] * capi.ensemble_factor(seed, sigma, copy, b, q) for b in range(n)]) for q in range(6)]
    w.add_bodies(*pert, st["gf"] / oracle.GRAVITY)
    for b in range(n):
        w.set_object_state(b, gf=st["gf"][b])
    w.dt = dt
    w.update(steps)
    return w.state()


@pytest.mark.parametrize("exact", [True, False])
def test_ensemble_copies_match_oracle(worlds_dir, exact):
    ow = oracle.World.load(os.path.join(worlds_dir, "jupiter_system.essa"))
    st = ow.state()
    copies, steps, dt, seed, sigma = 1000, 25, 600, 7, 1e-6
    c = capi.Context(0)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
    c.ensemble_init(copies, dt, seed=seed, rel_sigma=sigma, exact_order=exact)
    c.ensemble_step(10)
    c.ensemble_step(steps - 10)
    sample = [0, 1, 2, 3, 499, 500, 998, 999] + list(np.random.default_rng(0).choice(copies, 24, replace=False))
    for cp in sample:
        got = c.ensemble_read(int(cp), 1)
        ref = _ensemble_reference(st, int(cp), steps, dt, seed, sigma)
        pos = np.stack([got["x"][0], got["y"][0], got["z"][0]], axis=1)
        vel = np.stack([got["vx"][0], got["vy"][0], got["vz"][0]], axis=1)
        if exact:
            assert np.array_equal(pos, ref["pos"]) and np.array_equal(vel, ref["vel"]), cp
        else:
            assert np.abs(pos - ref["pos"]).max() <= 1e-11 * np.abs(ref["pos"]).max()
            assert np.abs(vel - ref["vel"]).max() <= 1e-10 * np.abs(ref["vel"]).max()
    # copy 0 is the unperturbed template
    got0 = c.ensemble_read(0, 2)
    assert not np.array_equal(got0["x"][0], got0["x"][1])
    c.close()


def _oracle_world_from(comp, gf):
    w = oracle.World()
    w.add_bodies(*comp, gf / oracle.GRAVITY)
    for b in range(len(gf)):
        w.set_object_state(b, gf=gf[b])
    return w


@pytest.mark.parametrize("case", ["2body.essa", "8body.essa", "solar.essa", "synthetic5", "synthetic13", "synthetic16", "synthetic17",
                                  "synthetic20", "synthetic33", "synthetic40", "synthetic64", "jupiter_system.essa", "synthetic100",
                                  "synthetic160", "synthetic161"])
def test_ensemble_pair_shared_kernels(worlds_dir, case):
    """Fast-arithmetic ensembles evaluate each PAIR once (ensemble.cu): templates of up to 16 bodies on one thread per copy
    (k_ensemble_solo), 17 .. 160 bodies on a ring of 2 .. 32 lanes per copy (k_ensemble_ring: every shape of the ring, full and
    padded homes), larger ones on the two-bodies-per-thread kernel.  Perturbed copies against oracle runs of the rebuilt initial
    state (1e-11, the bar of the other fast-arithmetic kernels), several launches, a copy count that is no multiple of the
    CTA, time reversed."""
    if case.endswith(".essa"):
        st = oracle.World.load(os.path.join(worlds_dir, case)).state()
    else:
        n = int(case[len("synthetic"):])
        rng = np.random.default_rng(n)
        ang = rng.uniform(0, 2 * np.pi, n)
        rad = 1e11 * (1 + np.arange(n))
        m = np.concatenate([[2e30], 10 ** rng.uniform(22, 27, n - 1)])
        v = np.sqrt(oracle.GRAVITY * 2e30 / rad)
        pos = np.stack([rad * np.cos(ang), rad * np.sin(ang), 1e9 * rng.standard_normal(n)], axis=1)
        vel = np.stack([-v * np.sin(ang), v * np.cos(ang), 10 * rng.standard_normal(n)], axis=1)
        pos[0] = vel[0] = 0
        st = {"pos": pos, "vel": vel, "gf": m * oracle.GRAVITY}
    copies, dt, seed, sigma = 333, 600, 11, 1e-6
    c = capi.Context(0)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
    c.ensemble_init(copies, dt, seed=seed, rel_sigma=sigma)
    for k in (1, 7, 32):
        c.ensemble_step(k)
    sample = [0, 1, 31, 32, 33, 319, 320, 332]
    refs = {}
    for cp in sample:
        got = c.ensemble_read(cp, 1)
        refs[cp] = ref = _ensemble_reference(st, cp, 40, dt, seed, sigma)
        pos = np.stack([got["x"][0], got["y"][0], got["z"][0]], axis=1)
        vel = np.stack([got["vx"][0], got["vy"][0], got["vz"][0]], axis=1)
        assert np.abs(pos - ref["pos"]).max() <= 1e-11 * np.abs(ref["pos"]).max(), cp
        assert np.abs(vel - ref["vel"]).max() <= 1e-10 * np.abs(ref["vel"]).max(), cp
    # Leapfrog is time reversible: 40 steps back land on the perturbed initial state to rounding
    c.ensemble_step(-40)
    comp = [st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2]]
    for cp in (1, 332):
        got = c.ensemble_read(cp, 1)
        for q, name in enumerate("xyz"):
            want = np.array([comp[q][b] * capi.ensemble_factor(seed, sigma, cp, b, q) for b in range(len(st["gf"]))])
            assert np.abs(got[name][0] - want).max() <= 1e-9 * np.abs(st["pos"]).max(), (cp, name)
    c.close()


@pytest.mark.parametrize("fname", ["8body.essa", "jupiter_system.essa"])
def test_ensemble_skips_a_deleted_body(worlds_dir, fname):
    """A body that Object::deleted() masks (src/Object.hpp deleted()) is neither attracted nor moved and attracts nobody:
    the other bodies move as the oracle world WITHOUT it does, and it stays where it was -- in the pair-shared kernels
    (fast arithmetic: one thread per copy / ring of lanes) and in the exact-order kernel alike."""
    st = oracle.World.load(os.path.join(worlds_dir, fname)).state()
    n, gone = len(st["gf"]), 2
    keep = [b for b in range(n) if b != gone]
    comp = [st["pos"][keep, 0], st["pos"][keep, 1], st["pos"][keep, 2], st["vel"][keep, 0], st["vel"][keep, 1], st["vel"][keep, 2]]
    ow = _oracle_world_from(comp, st["gf"][keep])
    ow.dt = 600
    ow.update(30)
    ref = ow.state()
    deleted = np.zeros(n, dtype=np.uint8)
    deleted[gone] = 1
    for exact in (False, True):
        c = capi.Context(0)
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
                 creation_date=np.full(n, capi.START_DATE, dtype=np.int64), deleted=deleted)
        c.ensemble_init(40, 600, seed=3, rel_sigma=0.0, exact_order=exact)
        c.ensemble_step(10)
        c.ensemble_step(20)
        got = c.ensemble_read(37, 1)
        pos = np.stack([got["x"][0], got["y"][0], got["z"][0]], axis=1)
        vel = np.stack([got["vx"][0], got["vy"][0], got["vz"][0]], axis=1)
        assert np.array_equal(pos[gone], st["pos"][gone]) and np.array_equal(vel[gone], st["vel"][gone])
        if exact:
            assert np.array_equal(pos[keep], ref["pos"]) and np.array_equal(vel[keep], ref["vel"])
        else:
            assert np.abs(pos[keep] - ref["pos"]).max() <= 1e-11 * np.abs(ref["pos"]).max()
            assert np.abs(vel[keep] - ref["vel"]).max() <= 1e-10 * np.abs(ref["vel"]).max()
        c.close()


def test_ensemble_closest_approach_to_body(worlds_dir):
    ow = oracle.World.load(os.path.join(worlds_dir, "solar.essa"))
    st = ow.state()
    c = capi.Context(0)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
    c.ensemble_init(8, 600, seed=1, rel_sigma=0.0, exact_order=True, closest_approaches=True, approach_to=3)
    c.ensemble_step(100)
    got = c.ensemble_read(0, 8, min_distance=True)
    fw = ow.clone_for_forward_simulation()
    fw.dt = 600
    fw.update(100)
    for b in range(len(ow)):
        if b == 3:
            continue
        assert got["min_distance"][5][b] == fw.closest_approach(b, 3)["distance"]
    c.close()


@pytest.mark.parametrize("exact,fname,to", [(True, "solar.essa", 3), (False, "solar.essa", 3), (False, "jupiter_system.essa", 0),
                                            (False, "jupiter_system.essa", 41), (True, "jupiter_system.essa", 76)])
def test_ensemble_candidate_service_approach_table_and_trail(worlds_dir, exact, fname, to):
    """SURVEY 8 f2 -- what src/essagui/EssaCreateObject.cpp:166-189 does for ONE candidate on the GUI thread (clone the world,
    update(ticks), show the candidate's trail and its closest approaches, src/Object.cpp:405-417), for many perturbed copies in
    one launch: per copy, the designated body's closest approach to EVERY other body (distance + both positions) and its trail.
    exact_order: bit-identical to the oracle's forward-simulated clone; fast arithmetic: 1e-11."""
    ow = oracle.World.load(os.path.join(worlds_dir, fname))
    st = ow.state()
    n, steps, every = len(ow), 120, 10
    c = capi.Context(0)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
    c.ensemble_init(70, 600, seed=1, rel_sigma=0.0, exact_order=exact, closest_approaches=True, approach_to=to, trail_every=every,
                    trail_capacity=steps // every + 3)
    c.ensemble_step(50)   # two launches: minima, positions and the trail's sampling carry over
    c.ensemble_step(steps - 50)
    got = c.ensemble_read(0, 70, min_distance=True)
    this_pos, other_pos = c.ensemble_approaches(0, 70)
    trail, tlen = c.ensemble_trail(0, 70)
    fw = ow.clone_for_forward_simulation()
    fw.dt = 600
    want_trail = []
    for s in range(steps):
        fw.update(1)
        if (s + 1) % every == 0:
            want_trail.append(fw.state()["pos"][to].copy())
    want_trail = np.array(want_trail)
    for cp in (0, 33, 69):  # sigma = 0: every copy is the template
        assert tlen[cp] == steps // every
        for b in range(n):
            if b == to:
                continue
            e = fw.closest_approach(b, to)
            if exact:
                assert got["min_distance"][cp][b] == e["distance"]
                assert np.array_equal(this_pos[cp][b], e["this_position"]) and np.array_equal(other_pos[cp][b], e["other_object_position"])
            else:
                assert got["min_distance"][cp][b] == pytest.approx(e["distance"], rel=1e-11)
                assert np.allclose(this_pos[cp][b], e["this_position"], rtol=1e-10, atol=1e-3)
                assert np.allclose(other_pos[cp][b], e["other_object_position"], rtol=1e-10, atol=1e-3)
        if exact:
            assert np.array_equal(trail[cp][:tlen[cp]], want_trail)
        else:
            assert np.allclose(trail[cp][:tlen[cp]], want_trail, rtol=1e-11, atol=1e-3)
    c.close()


def test_ensemble_moves_a_massless_test_particle(worlds_dir):
    """A live body of mass 0 attracts nothing but is attracted and moves (the reference integrates it); only bodies that
    Object::deleted() masks stay put.  The ensemble used to infer the mask from gf == 0."""
    ow = oracle.World.load(os.path.join(worlds_dir, "8body.essa"))
    ow.add_object(0.0, 1.0, (2.5e11, 0.0, 0.0), (0.0, 2.3e4, 0.0), "probe", 100)
    st = ow.state()
    n = len(ow)
    assert st["gf"][n - 1] == 0.0
    for exact in (True, False):
        c = capi.Context(0)
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
                 creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
        c.ensemble_init(4, 600, seed=1, rel_sigma=0.0, exact_order=exact)
        c.ensemble_step(50)
        got = c.ensemble_read(0, 4)
        fw = ow.clone_for_forward_simulation()
        fw.dt = 600
        fw.update(50)
        ref = fw.state()["pos"]
        mine = np.stack([got["x"][2], got["y"][2], got["z"][2]], axis=1)
        assert not np.array_equal(mine[n - 1], st["pos"][n - 1])  # it moved
        if exact:
            assert np.array_equal(mine, ref)
        else:
            assert np.allclose(mine, ref, rtol=1e-11, atol=1e-2)
        c.close()


def test_object_history_moves_objects_out_and_in_with_the_date(worlds_dir):
    """src/World.cpp:59-84 + src/ObjectHistory.cpp:4-29: once the ObjectHistory holds an entry, stepping the date below the
    creation date of the list's last object moves that object into the history (with its History / Trail reset), and stepping
    forward pops entries back, one per step.  The reference's current revision never seeds the history itself, so both
    sides are seeded through their debug hook; after that every step must agree bit for bit (exact_order), object count and
    date included."""
    ow, pw = _pair(worlds_dir, "8body.essa", 600)
    for w in (ow, pw):
        w.update(5)
    ow.add_object(3e24, 1e6, (4e11, 1e10, 0), (0, 17000.0, 0), "Late1", 300)
    pw.add_object(mass=3e24, radius=1e6, pos=(4e11, 1e10, 0), vel=(0, 17000.0, 0), name="Late1", period=300)
    for w in (ow, pw):
        w.update(4)
    ow.add_object(5e23, 1e6, (-3e11, 2e10, 1e9), (0, -19000.0, 0), "Late2", 200)
    pw.add_object(mass=5e23, radius=1e6, pos=(-3e11, 2e10, 1e9), vel=(0, -19000.0, 0), name="Late2", period=200)
    for w in (ow, pw):
        w.update(3)
    _assert_same_state(pw, ow)
    n_full = len(ow)
    ow.debug_stash_last()
    pw.debug_stash_last_object()
    assert len(ow) == len(pw) == n_full - 1 and ow.object_history_size == pw.object_history_size == 1
    # backwards past Late1's creation date: it joins Late2 in the history
    for _ in range(9):
        ow.update(-1)
        pw.update(-1)
        assert len(pw) == len(ow) and pw.object_history_size == ow.object_history_size
        _assert_same_state(pw, ow)
    assert len(ow) == n_full - 2 and ow.object_history_size == 2
    # forwards again: entries come back one per step (masked by Object::deleted() until their creation date), histories reset
    for _ in range(14):
        ow.update(1)
        pw.update(1)
        assert len(pw) == len(ow) and pw.object_history_size == ow.object_history_size
        _assert_same_state(pw, ow)
    assert len(ow) == n_full and ow.object_history_size == 0
    _assert_same_rings(pw, ow)
    # with the history empty again the world is back on the ordinary path (several steps per call)
    ow.update(20)
    pw.update(20)
    _assert_same_state(pw, ow)
    _assert_same_rings(pw, ow)


def test_snapshot_round_trip_and_zero_copy_trail_view(worlds_dir):
    """What comes after the path: the Trail ring as a zero-copy device view for a renderer (src/Trail.cpp:90-103
    draws exactly these vertices in one or two line strips), and a flat snapshot of the whole context -- bodies, date, History
    and Trail rings -- that restores into a fresh context which then continues bit for bit."""
    pw = world.World.load(os.path.join(worlds_dir, "solar.essa"))
    st = pw.state()
    caps = [pw.props(i)["trail_capacity"] for i in range(len(pw))]
    pw.close()
    n = len(caps)
    a = capi.Context(0)
    a.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
             creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
    a.record_configure(caps)
    a.step(1500, 43200, record=True, exact_order=True)  # History rings wrapped (1000), short trails wrapped
    blob = a.snapshot()
    assert blob[:8] == b"ESSASNAP"
    b = capi.Context(0)
    b.restore(blob)
    assert b.date == a.date and b.n == n and b.tracked == n
    for c in (a, b):
        c.step(40, 43200, record=True, exact_order=True)
    fa, fb = a.download(), b.download()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    for i in range(n):
        assert np.array_equal(a.history(i), b.history(i))
        ta, tb = a.trail(i), b.trail(i)
        assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"])
        assert np.array_equal(ta["verts"], tb["verts"]) and np.array_equal(ta["offset"], tb["offset"])
        assert ta["last_xy_angle"] == tb["last_xy_angle"]
    # the device view: same cursor, and the very bytes essa_trail_read copies out
    try:
        from cuda import cudart
    except ImportError:
        cudart = None
    for i in (0, 3, n - 1):
        v = a.trail_device_view(i)
        t = a.trail(i)
        assert v["verts_device"] and (v["capacity"], v["append_offset"], v["length"]) == (t["capacity"], t["append_offset"], t["length"])
        assert np.array_equal(v["offset"], t["offset"])
        if cudart is not None:
            host = np.zeros((v["capacity"], 3), dtype=np.float32)
            err, = cudart.cudaMemcpy(host.ctypes.data, v["verts_device"], host.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            assert int(err) == 0 and np.array_equal(host, t["verts"])
    a.close()
    b.close()
