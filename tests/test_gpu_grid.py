"""GPU parity of K1g (essa_b200/csrc/force_grid.cu): mid-size worlds stepped by all SMs in one cooperative launch per call,
one grid barrier per step.  Bars as for K1 (tests/test_gpu_tiled.py): trajectories of unsoftened Plummer spheres within 1e-10
of the CPU oracle's World::update (src/World.cpp:86-139), most-attracting links as the oracle's rule gives them, and the
size-independent invariants (split calls == one call bitwise, +K then -K returns, momentum)."""
import numpy as np
import pytest

import oracle
from essa_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _upload(ctx, w, **kw):
    ctx.date = capi.START_DATE
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], **kw)


@pytest.mark.parametrize("n,steps", [(1025, 10), (2048, 6), (2048, -6), (3000, 4), (700, 8)])
def test_grid_kdk_trajectory_matches_oracle(ctx, n, steps):
    w = synth.plummer(n, seed=11)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(steps, dt, kernel=capi.KERNEL_GRID)
    got = ctx.download()
    (ox, oy, oz, ovx, ovy, ovz), oacc = oracle.step_soa(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], steps, dt)
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    vel = np.stack([got["vx"], got["vy"], got["vz"]], axis=1)
    acc = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    assert _rel_rows(pos, np.stack([ox, oy, oz], axis=1)).max() <= 1e-10
    assert _rel_rows(vel, np.stack([ovx, ovy, ovz], axis=1)).max() <= 1e-10
    assert _rel_rows(acc, oacc.T).max() <= 1e-9
    assert ctx.date == capi.START_DATE + steps * dt
    assert ctx.last_force_passes == abs(steps) + 1


@pytest.mark.parametrize("n", [1025, 4097, 5120])
def test_grid_agrees_with_tiled_over_the_size_range(ctx, n):
    """Every CTA / lane split the launcher can pick (T = 7 .. 42 bodies per CTA), ragged last CTA, odd step counts (the
    position buffers end with swapped roles), a following call that reuses the accelerations, unequal masses with links."""
    w = synth.plummer(n, seed=n)
    rng = np.random.default_rng(n)
    gf = w["gf"] * 10.0 ** rng.uniform(-1, 1, n)
    dt = synth.default_dt(n)
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_GRID):
        ctx.date = capi.START_DATE
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, most=np.full(n, -1, dtype=np.int32))
        ctx.step(3, dt, kernel=kern, track_most_attracting=True)
        ctx.step(2, dt, kernel=kern, track_most_attracting=True)
        got[kern] = ctx.download()
    t, g = got[capi.KERNEL_TILED], got[capi.KERNEL_GRID]
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.allclose(g[k], t[k], rtol=1e-11, atol=0.0), k
    acc_t = np.stack([t["ax"], t["ay"], t["az"]], axis=1)
    acc_g = np.stack([g["ax"], g["ay"], g["az"]], axis=1)
    assert _rel_rows(acc_g, acc_t).max() <= 1e-11
    assert np.array_equal(g["most"], t["most"])
    assert np.all(np.abs(g["max_attraction"] - t["max_attraction"]) <= 1e-11 * np.abs(t["max_attraction"]))


def test_grid_most_attracting_links_follow_the_reference_rule(ctx):
    """src/Object.cpp:84-94 on the final positions: the strictly heavier partner with the largest gf / r^4; masked bodies are
    neither candidates nor targets; old_most is the link before the last step."""
    n = 2500
    w = synth.plummer(n, seed=31)
    rng = np.random.default_rng(7)
    gf = w["gf"] * 10.0 ** rng.uniform(-2, 2, n)
    t0 = capi.START_DATE
    cdate = np.full(n, t0, dtype=np.int64)
    cdate[rng.choice(n, 30, replace=False)] = t0 + 10 ** 9
    dt = synth.default_dt(n)
    ctx.date = t0
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, creation_date=cdate, most=np.full(n, -1, dtype=np.int32))
    ctx.step(3, dt, kernel=capi.KERNEL_GRID, track_most_attracting=True)
    p = ctx.download()
    act = cdate <= t0
    assert (p["most"][~act] == -1).all()
    for k in ("x", "y", "z"):
        assert np.array_equal(p[k][~act], w[k][~act])  # masked bodies do not move
    x, y, z = p["x"], p["y"], p["z"]
    for i in rng.choice(np.where(act)[0], 96, replace=False):
        r2 = (x - x[i]) ** 2 + (y - y[i]) ** 2 + (z - z[i]) ** 2
        cand = act & (gf > gf[i])
        cand[i] = False
        if not cand.any():
            assert p["most"][i] == -1
            continue
        r2[i] = np.inf
        metric = np.where(cand, gf / r2 ** 2, 0.0)
        assert p["most"][i] == int(np.argmax(metric)), i
        assert p["max_attraction"][i] == pytest.approx(metric.max(), rel=1e-9)


def test_grid_exact_tie_takes_the_lowest_index(ctx):
    """Two heavier partners at mirrored positions (bit-identical metrics): the lower index wins (strict '>' in ascending
    order, src/Object.cpp:86), whichever lanes of the body's group met them.  Masses so small that nothing moves by half an
    ulp during the one step the kernel needs (it serves essa_step only), so the metrics stay bit-identical."""
    n = 2048
    w = synth.plummer(n, seed=5)
    zero = np.zeros(n)
    for t, a, b in ((100, 1500, 700), (7, 8, 9), (2047, 2046, 3)):
        x, y, z = w["x"] + 1e13, w["y"].copy(), w["z"].copy()
        mass = np.full(n, 1e10)
        x[t], y[t], z[t] = 0.0, 0.0, 0.0
        x[a], y[a], z[a] = 1e9, 0.0, 0.0
        x[b], y[b], z[b] = 0.0, 1e9, 0.0
        mass[a] = mass[b] = 1e12
        ctx.date = capi.START_DATE
        ctx.upload(x, y, z, zero, zero, zero, mass * 6.67408e-11, most=np.full(n, -1, dtype=np.int32))
        ctx.step(1, 1, kernel=capi.KERNEL_GRID, track_most_attracting=True)
        got = ctx.download(("most", "x", "y"))
        assert got["x"][a] == 1e9 and got["y"][b] == 1e9
        assert int(got["most"][t]) == min(a, b)
        assert got["most"][a] == -1 and got["most"][b] == -1


def test_grid_split_calls_equal_one_call_bitwise_and_runs_are_reproducible(ctx):
    n = 1500
    w = synth.plummer(n, seed=4)
    dt = synth.default_dt(n)
    runs = []
    for split in ((7,), (7,), (2, 1, 4)):
        _upload(ctx, w)
        for k in split:
            ctx.step(k, dt, kernel=capi.KERNEL_GRID)
        runs.append(ctx.download())
    for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az"):
        assert np.array_equal(runs[0][k], runs[1][k]), k
        assert np.array_equal(runs[0][k], runs[2][k]), k


def test_grid_reversibility_and_momentum(ctx):
    n = 2048
    w = synth.plummer(n, seed=8)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(9, dt, kernel=capi.KERNEL_GRID)
    mid = ctx.download()
    ctx.step(-9, dt, kernel=capi.KERNEL_GRID)
    back = ctx.download()
    scale = np.abs(w["x"]).max()
    for k in ("x", "y", "z"):
        assert np.abs(mid[k] - w[k]).max() > 1e-6 * scale   # it did move
        assert np.abs(back[k] - w[k]).max() <= 1e-9 * scale
    m = w["gf"]
    for k in ("vx", "vy", "vz"):
        assert abs(np.sum(m * mid[k]) - np.sum(m * w[k])) <= 1e-12 * np.sum(m * np.abs(mid[k]))


def test_auto_takes_the_grid_kernel_and_leaves_what_it_cannot_do_to_the_tiled_one(ctx):
    """AUTO: 1,025 .. 5,120 bodies go through K1g; a call that crosses a creation instant, recording and two_pass stay with K1
    (one launch per step); an explicit request the kernel cannot serve is an error."""
    n = 2048
    w = synth.plummer(n, seed=2)
    dt = synth.default_dt(n)
    t0 = capi.START_DATE
    _upload(ctx, w)
    l0 = ctx.launches
    ctx.step(5, dt)
    assert ctx.launches - l0 == 1  # one cooperative launch for the whole call
    a = ctx.download()
    _upload(ctx, w)
    ctx.step(5, dt, kernel=capi.KERNEL_GRID)
    b = ctx.download()
    for k in ("x", "vx"):
        assert np.array_equal(a[k], b[k])
    cdate = np.full(n, t0, dtype=np.int64)
    cdate[5] = t0 + 2 * dt  # body 5 appears after two steps
    _upload(ctx, w, creation_date=cdate)
    with pytest.raises(capi.EssaError):
        ctx.step(5, dt, kernel=capi.KERNEL_GRID)
    _upload(ctx, w, creation_date=cdate)
    ctx.step(5, dt)  # AUTO: the tiled kernel, with the mask flipping inside the call
    c = ctx.download()
    _upload(ctx, w, creation_date=cdate)
    ctx.step(5, dt, kernel=capi.KERNEL_TILED)
    d = ctx.download()
    for k in ("x", "vx"):
        assert np.array_equal(c[k], d[k])
    assert c["x"][5] != w["x"][5]  # it did join


def test_grid_softening_deleted_bodies_and_forward_simulation_agree_with_tiled(ctx):
    """What else essa_step_opts can ask of the kernel: Plummer softening (eps2, the same for every pair), bodies whose deletion
    date has passed (Object::deleted(), src/Object.cpp:60-62: neither attracted nor attracting nor moved), forward_simulated
    (the date stands still, src/World.cpp:59-61) -- against K1 with the same options, and the date bookkeeping."""
    n = 3000
    w = synth.plummer(n, seed=13)
    rng = np.random.default_rng(14)
    t0 = capi.START_DATE
    dt = synth.default_dt(n)
    ddate = np.full(n, t0 - 1, dtype=np.int64)
    deleted = np.zeros(n, dtype=np.uint8)
    gone = rng.choice(n, 50, replace=False)
    deleted[gone] = 1
    eps2 = float(np.median(np.abs(w["x"])) * 1e-3) ** 2
    for fwd in (False, True):
        got = {}
        for kern in (capi.KERNEL_TILED, capi.KERNEL_GRID):
            ctx.date = t0
            ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], deletion_date=ddate, deleted=deleted)
            ctx.step(4, dt, kernel=kern, eps2=eps2, forward_simulated=fwd)
            got[kern] = ctx.download()
            assert ctx.date == (t0 if fwd else t0 + 4 * dt)
        t, g = got[capi.KERNEL_TILED], got[capi.KERNEL_GRID]
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            assert np.allclose(g[k], t[k], rtol=1e-11, atol=0.0), k
            assert np.array_equal(g[k][gone], w[k][gone])  # deleted bodies stay where they were
        live = deleted == 0
        assert np.abs(g["x"][live] - w["x"][live]).max() > 0


@pytest.mark.parametrize("n", [64, 147, 148, 149, 400])
def test_grid_small_worlds_one_body_per_cta_and_fewer_bodies_than_sms(ctx, n):
    """T = 1 (fewer bodies than SMs: some CTAs own nothing but still take part in every barrier), T = 2 with a ragged last CTA,
    AUTO's lower bound: trajectories against the oracle."""
    w = synth.plummer(n, seed=n)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(7, dt, kernel=capi.KERNEL_GRID)
    got = ctx.download()
    (ox, oy, oz, ovx, ovy, ovz), _ = oracle.step_soa(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], 7, dt)
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    vel = np.stack([got["vx"], got["vy"], got["vz"]], axis=1)
    assert _rel_rows(pos, np.stack([ox, oy, oz], axis=1)).max() <= 1e-10
    assert _rel_rows(vel, np.stack([ovx, ovy, ovz], axis=1)).max() <= 1e-10


def test_grid_refuses_worlds_outside_its_range(ctx):
    for n in (32, 5121):
        w = synth.plummer(n, seed=1)
        _upload(ctx, w)
        with pytest.raises(capi.EssaError):
            ctx.step(1, 1000, kernel=capi.KERNEL_GRID)
