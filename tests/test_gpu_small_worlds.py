"""Boundary sizes of the small-world kernels: n = 1, 2, 31, 32 (one-warp kernel K3w), 33, 200 (resident
CTA, 255-register variant), 300 (resident CTA, 1024-thread variant), random hierarchical worlds with
recording on, against the oracle: exact_order bit for bit, fast within 1e-12 * sqrt(K)."""
import numpy as np
import pytest

import oracle
from essa_b200 import capi, world

pytestmark = pytest.mark.gpu


def _random_world(n, seed):
    """A heavy primary, a few sub-primaries, light satellites on roughly circular orbits."""
    rng = np.random.default_rng(seed)
    bodies = [(2e30, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 600)]
    for k in range(1, n):
        parent = 0 if k < 6 or rng.random() < 0.5 else int(rng.integers(1, min(k, 6)))
        pm, ppos, pvel, _ = bodies[parent]
        r = (rng.uniform(0.3, 30) * 1.5e11) if parent == 0 else rng.uniform(2e8, 2e9)
        th = rng.uniform(0, 2 * np.pi)
        mass = (10.0 ** rng.uniform(24, 27)) if parent == 0 and k < 6 else 10.0 ** rng.uniform(18, 22)
        v = np.sqrt(oracle.GRAVITY * pm / r)
        pos = (ppos[0] + r * np.cos(th), ppos[1] + r * np.sin(th), ppos[2] + rng.uniform(-1, 1) * 1e-3 * r)
        vel = (pvel[0] - v * np.sin(th), pvel[1] + v * np.cos(th), pvel[2])
        bodies.append((mass, pos, vel, int(rng.integers(1, 900))))
    return bodies


def _pair(n, seed, exact):
    ow, pw = oracle.World(), world.World()
    for k, (m, p, v, period) in enumerate(_random_world(n, seed)):
        ow.add_object(m, 1e6, p, v, "b%d" % k, period)
        pw.add_object(mass=m, radius=1e6, pos=p, vel=v, name="b%d" % k, period=period)
    ow.dt = 600
    pw.simulation_seconds_per_tick = 600
    pw.set_flags(offset_trails=True, exact_order=exact)
    return ow, pw


def _compare(pw, ow, exact, rtol):
    a, b = pw.state(), ow.state()
    assert np.array_equal(a["most"], b["most"])
    for k in ("pos", "vel", "acc", "maxattr", "appe"):
        if exact:
            assert np.array_equal(a[k], b[k]), k
        else:
            num = np.abs(a[k] - b[k]).reshape(len(a[k]), -1).max(axis=1)
            den = np.abs(b[k]).reshape(len(b[k]), -1).max(axis=1) + 1e-300
            assert (num / den).max() <= rtol * (1e3 if k in ("acc", "maxattr") else 1.0), (k, (num / den).max())
    for i in range(len(ow)):
        ta, tb = pw.trail(i), ow.trail(i)
        assert (ta["append_offset"], ta["length"]) == (tb["append_offset"], tb["length"]), i
        assert pw.history(i).shape == ow.history(i).shape
        if exact:
            assert np.array_equal(ta["verts"][:ta["length"]], tb["verts"][:tb["length"]]), i
            assert np.array_equal(pw.history(i), ow.history(i)), i


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 200, 300, 1024])
def test_exact_order_boundary_sizes(n):
    steps = 120 if n <= 33 else (25 if n <= 300 else 4)
    ow, pw = _pair(n, 100 + n, exact=True)
    for chunk in (1, steps - 1):
        ow.update(chunk)
        pw.update(chunk)
    _compare(pw, ow, True, 0.0)


# 1 .. 32: K3q (two-SM cluster); 33 .. 128: K3x (16-SM cluster; 64 | 65 and 96 | 97 switch the partners per lane,
# 33, 49, 113 the bodies per CTA); 129 .. 512: K3xw with 16 lanes per body, 513 .. 1024: 8 lanes per body
@pytest.mark.parametrize("n", [1, 2, 5, 16, 31, 32, 33, 49, 64, 65, 96, 97, 100, 113, 128, 129, 200, 300, 512, 513, 1000, 1024])
def test_fast_boundary_sizes(n):
    steps = 400 if n <= 129 else (60 if n <= 512 else 12)
    ow, pw = _pair(n, 200 + n, exact=False)
    for chunk in (3, steps - 3):
        ow.update(chunk)
        pw.update(chunk)
    _compare(pw, ow, False, 1e-12 * np.sqrt(steps))


@pytest.mark.parametrize("n", [10, 77])
def test_fast_masked_bodies_and_backward_steps(n):
    """Fast-arithmetic cluster kernels (K3q / K3x), recording on: backward steps, then bodies that deleted() masks
    for a whole call -- two created in the future, two marked deleted."""
    ow, pw = _pair(n, 300 + n, exact=False)
    t0 = pw.date
    for w in (ow, pw):
        w.date = t0 + 10 ** 8  # World::add_object stamps the creation date with the world's date
    ow.add_object(1e25, 1e6, (3e11, 1e11, 0.0), (0.0, 2e4, 0.0), "late0", 300)
    pw.add_object(mass=1e25, radius=1e6, pos=(3e11, 1e11, 0.0), vel=(0.0, 2e4, 0.0), name="late0", period=300)
    ow.add_object(1e20, 1e6, (-4e11, 0.0, 1e9), (0.0, -1.8e4, 0.0), "late1", 500)
    pw.add_object(mass=1e20, radius=1e6, pos=(-4e11, 0.0, 1e9), vel=(0.0, -1.8e4, 0.0), name="late1", period=500)
    for w in (ow, pw):
        w.date = t0
    late = [n, n + 1]
    for steps in (100, -60):
        ow.update(steps)
        pw.update(steps)
    ow.delete_object(2)
    pw.mark_deleted(2)
    ow.delete_object(n - 1)
    pw.mark_deleted(n - 1)
    for steps in (150, 30):
        ow.update(steps)
        pw.update(steps)
    a, b = pw.state(), ow.state()
    assert np.array_equal(a["active"], b["active"]) and not a["active"][late + [2, n - 1]].any()
    assert np.array_equal(a["pos"][late], b["pos"][late])  # masked bodies do not move
    _compare(pw, ow, False, 1e-12 * np.sqrt(340) * 4)


def test_warp_kernel_long_run_wraps_rings():
    """K3w over 1,300 steps at the default 12-hour tick: History wraps, 500-vertex trails wrap, the
    snapshot ring between the physics and recorder warps cycles 80 times."""
    ow, pw = _pair(6, 7, exact=False)  # planets only: satellites with day-long periods are chaotic at this tick
    ow.dt = 43200
    pw.simulation_seconds_per_tick = 43200
    ow.update(1300)
    pw.update(1300)
    _compare(pw, ow, False, 1e-12 * np.sqrt(1300) * 10)
    assert pw.history(3).shape == (1000, 6)


def test_ensemble_copies_not_multiple_of_cta():
    st_w = _random_world(9, 3)
    c = capi.Context(0)
    x, y, z = (np.array([b[1][k] for b in st_w]) for k in range(3))
    vx, vy, vz = (np.array([b[2][k] for b in st_w]) for k in range(3))
    gf = np.array([b[0] for b in st_w]) * capi.GRAVITY
    c.upload(x, y, z, vx, vy, vz, gf)
    c.ensemble_init(1001, 600, seed=11, rel_sigma=1e-5, exact_order=True)
    c.ensemble_step(30)
    got = c.ensemble_read(1000, 1)
    ow = oracle.World()
    comp = [x, y, z, vx, vy, vz]
    pert = [np.array([comp[q][b] * capi.ensemble_factor(11, 1e-5, 1000, b, q) for b in range(9)]) for q in range(6)]
    ow.add_bodies(*pert, gf / oracle.GRAVITY)
    for b in range(9):
        ow.set_object_state(b, gf=gf[b])
    ow.dt = 600
    ow.update(30)
    ref = ow.state()
    assert np.array_equal(np.stack([got["x"][0], got["y"][0], got["z"][0]], axis=1), ref["pos"])
    assert np.array_equal(np.stack([got["vx"][0], got["vy"][0], got["vz"][0]], axis=1), ref["vel"])
    c.close()
