"""GPU parity of K1p, the pair-shared force pass (each unordered pair evaluated once and applied to
both bodies, like the reference's own i<j loop), against the CPU oracle.  Same bars as K1
(tests/test_gpu_tiled.py): single pass within max(1e-13, 4x the oracle's own error vs extended
precision); 10-step trajectories of unsoftened Plummer spheres within 1e-10."""
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


@pytest.fixture(params=["mass_sorted", "unsorted"])
def layout(request, monkeypatch):
    """Tracked passes of an unsharded world run on the mass-sorted layout (filter modes 2 / 3 of k_force_paired);
    ESSA_PAIR_SORTED=0 keeps the guarded filter on the caller's order -- what the ranks of a sharded world run."""
    monkeypatch.setenv("ESSA_PAIR_SORTED", "1" if request.param == "mass_sorted" else "0")
    return request.param


def _upload(ctx, w):
    ctx.date = capi.START_DATE
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])


# blocks of 2048 bodies: nb = 2 (even, d = nb/2 handled by half the ring), 3 (odd), 4, 5, ragged last blocks, many blocks
@pytest.mark.parametrize("n", [2049, 4096, 6000, 8192, 10000, 16384, 20000, 65536])  # 16,384: BASELINE config and AUTO's routing boundary
def test_paired_force_pass_matches_extended_precision(ctx, n):
    w = synth.plummer(n, seed=n)
    _upload(ctx, w)
    ctx.set_forces(kernel=capi.KERNEL_PAIRED)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    rows = np.unique(np.concatenate([np.random.default_rng(1).choice(n, 192, replace=False), [0, 2047, 2048, n - 1]]))
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
    ref = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ref")
    err, err_ref = _rel_rows(a[rows], ld), _rel_rows(ref, ld)
    assert err.max() <= max(1e-13, 4.0 * err_ref.max()), (err.max(), err_ref.max())
    assert np.all(np.isfinite(a))
    # pair-shared evaluation conserves momentum to round-off: sum(m a) ~ 0
    ma = a * w["gf"][:, None]
    assert np.abs(ma.sum(axis=0)).max() <= 1e-12 * np.abs(ma).sum()


def test_paired_agrees_with_tiled(ctx):
    n = 30000
    w = synth.plummer(n, seed=2)
    _upload(ctx, w)
    ctx.set_forces(kernel=capi.KERNEL_TILED)
    t = ctx.download(("ax", "ay", "az"))
    _upload(ctx, w)
    ctx.set_forces(kernel=capi.KERNEL_PAIRED)
    p = ctx.download(("ax", "ay", "az"))
    a = np.stack([t["ax"], t["ay"], t["az"]], axis=1)
    b = np.stack([p["ax"], p["ay"], p["az"]], axis=1)
    assert _rel_rows(b, a).max() <= 1e-12


def test_paired_unequal_masses_and_masked_bodies(ctx):
    """gf varies over 6 decades; some bodies are masked by date (gf_eff = 0 records)."""
    n = 6000
    w = synth.plummer(n, seed=3)
    rng = np.random.default_rng(4)
    w["gf"] = w["gf"] * 10.0 ** rng.uniform(-3, 3, n)
    t0 = capi.START_DATE
    cdate = np.full(n, t0, dtype=np.int64)
    cdate[rng.choice(n, 50, replace=False)] = t0 + 10 ** 9  # not created yet
    ctx.date = t0
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], creation_date=cdate)
    ctx.set_forces(kernel=capi.KERNEL_PAIRED)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    act = cdate <= t0
    gfe = np.where(act, w["gf"], 0.0)
    rows = np.where(act)[0][:200]
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], gfe, rows, mode="ld")
    ref = oracle.forces_rows(w["x"], w["y"], w["z"], gfe, rows, mode="ref")
    assert _rel_rows(a[rows], ld).max() <= max(1e-13, 4.0 * _rel_rows(ref, ld).max())
    assert np.all(a[~act] == 0.0)  # masked bodies keep their (zero) acceleration


@pytest.mark.parametrize("n,steps", [(3000, 10), (3000, -6)])
def test_paired_kdk_trajectory_matches_oracle(ctx, n, steps):
    w = synth.plummer(n, seed=11)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(steps, dt, kernel=capi.KERNEL_PAIRED)
    got = ctx.download()
    (ox, oy, oz, ovx, ovy, ovz), oacc = oracle.step_soa(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], steps, dt)
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    vel = np.stack([got["vx"], got["vy"], got["vz"]], axis=1)
    assert _rel_rows(pos, np.stack([ox, oy, oz], axis=1)).max() <= 1e-10
    assert _rel_rows(vel, np.stack([ovx, ovy, ovz], axis=1)).max() <= 1e-10
    assert ctx.last_force_passes == abs(steps) + 1


def test_paired_is_deterministic_and_reuse_is_exact(ctx):
    n = 9000
    w = synth.plummer(n, seed=5)
    dt = synth.default_dt(n)
    runs = []
    for two_pass in (False, False, True):
        _upload(ctx, w)
        ctx.step(4, dt, kernel=capi.KERNEL_PAIRED, two_pass=two_pass)
        runs.append(ctx.download())
    for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az"):
        assert np.array_equal(runs[0][k], runs[1][k]), k   # no atomics: bitwise reproducible
        assert np.array_equal(runs[0][k], runs[2][k]), k   # reuse == recompute


def test_paired_refuses_what_it_cannot_do(ctx):
    w = synth.plummer(512, seed=1)
    _upload(ctx, w)
    with pytest.raises(capi.EssaError):
        ctx.step(1, 1000, kernel=capi.KERNEL_PAIRED)


@pytest.mark.parametrize("n", [4096, 6000, 16384, 20000])
def test_paired_most_attracting_links_match_tiled_and_oracle(ctx, n, layout):
    """Unequal masses: the pair-shared kernel carries the most-attracting candidates as sortable keys
    (src/Object.cpp:84-94: the heavier partner with the largest gf / r^4, lowest index on ties)."""
    w = synth.plummer(n, seed=31)
    rng = np.random.default_rng(n)
    gf = w["gf"] * 10.0 ** rng.uniform(-2, 2, n)
    t0 = capi.START_DATE
    cdate = np.full(n, t0, dtype=np.int64)
    cdate[rng.choice(n, 40, replace=False)] = t0 + 10 ** 9  # masked bodies are neither candidates nor targets
    dt = synth.default_dt(n)
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
        ctx.date = t0
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, creation_date=cdate,
                   most=np.full(n, -1, dtype=np.int32))
        ctx.step(2, dt, kernel=kern, track_most_attracting=True)
        got[kern] = ctx.download()
    t, p = got[capi.KERNEL_TILED], got[capi.KERNEL_PAIRED]
    assert np.array_equal(t["most"], p["most"])
    act = cdate <= t0
    assert (p["most"][~act] == -1).all() and (p["most"][act] >= 0).sum() >= act.sum() - 1  # all but the heaviest have a link
    m_t, m_p = t["max_attraction"], p["max_attraction"]
    assert np.abs(m_p - m_t).max() <= 1e-12 * np.abs(m_t).max() and np.all(np.abs(m_p[act] - m_t[act]) <= 1e-11 * np.abs(m_t[act]))
    # the links against a direct evaluation of the rule on the final positions
    x, y, z = p["x"], p["y"], p["z"]
    for i in rng.choice(np.where(act)[0], 64, replace=False):
        r2 = (x - x[i]) ** 2 + (y - y[i]) ** 2 + (z - z[i]) ** 2
        cand = act & (gf > gf[i])
        cand[i] = False
        if not cand.any():
            continue
        r2[i] = np.inf
        metric = np.where(cand, gf / r2 ** 2, 0.0)
        assert p["most"][i] == int(np.argmax(metric)), i
        assert m_p[i] == pytest.approx(metric.max(), rel=1e-9)


@pytest.mark.parametrize("track", [False, True])
def test_paired_coincident_bodies_in_different_blocks(ctx, track, layout):
    """The off-diagonal tiles of the pair-shared kernel run without the r = 0 test; bodies that share a position
    across blocks (incl. a body at the origin next to ragged-block padding, and a masked body on top of a live one)
    are repaired in k_pair_finish: the pair contributes nothing, as on every other path (the reference yields NaN)."""
    n = 9000  # blocks of 2048: 4.4 blocks, ragged last block
    w = synth.plummer(n, seed=41)
    rng = np.random.default_rng(42)
    gf = w["gf"] * 10.0 ** rng.uniform(-1, 1, n)
    x, y, z = w["x"].copy(), w["y"].copy(), w["z"].copy()
    pairs = [(5, 5000), (2047, 2048), (100, 8999), (4100, 4200)]  # the last pair shares a block (guarded diagonal tile)
    for a, b in pairs:
        x[b], y[b], z[b] = x[a], y[a], z[a]
    x[7000] = y[7000] = z[7000] = 0.0  # a live body at the origin
    t0 = capi.START_DATE
    cdate = np.full(n, t0, dtype=np.int64)
    cdate[300] = t0 + 10 ** 9  # masked body 300 ...
    x[6500], y[6500], z[6500] = x[300], y[300], z[300]  # ... under live body 6500
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
        ctx.date = t0
        ctx.upload(x, y, z, w["vx"], w["vy"], w["vz"], gf, creation_date=cdate, most=np.full(n, -1, dtype=np.int32))
        ctx.set_forces(kernel=kern, track_most_attracting=track)
        got[kern] = ctx.download()
    t, p = got[capi.KERNEL_TILED], got[capi.KERNEL_PAIRED]
    a = np.stack([t["ax"], t["ay"], t["az"]], axis=1)
    b = np.stack([p["ax"], p["ay"], p["az"]], axis=1)
    assert np.all(np.isfinite(b))
    live = cdate <= t0
    assert _rel_rows(b[live], a[live]).max() <= 1e-12
    assert np.all(b[~live] == 0.0)
    if track:
        assert np.array_equal(t["most"], p["most"])
        m_t, m_p = t["max_attraction"], p["max_attraction"]
        assert np.all(np.abs(m_p[live] - m_t[live]) <= 1e-11 * np.abs(m_t[live]))
        for i, j in pairs:  # a body on top of another one is never its most-attracting object
            assert p["most"][i] != j and p["most"][j] != i


def test_paired_coincident_bodies_uniform_masses(ctx):
    """Same for the UNIFORM instantiation (every body live, one gravity factor: sums scaled once per body)."""
    n = 9000
    w = synth.plummer(n, seed=43)
    x, y, z = w["x"].copy(), w["y"].copy(), w["z"].copy()
    for a, b in [(5, 5000), (2047, 2048), (100, 8999)]:
        x[b], y[b], z[b] = x[a], y[a], z[a]
    x[7000] = y[7000] = z[7000] = 0.0
    assert np.all(w["gf"] == w["gf"][0])
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
        ctx.date = capi.START_DATE
        ctx.upload(x, y, z, w["vx"], w["vy"], w["vz"], w["gf"])
        ctx.set_forces(kernel=kern)
        got[kern] = ctx.download(("ax", "ay", "az"))
    a = np.stack([got[capi.KERNEL_TILED][k] for k in ("ax", "ay", "az")], axis=1)
    b = np.stack([got[capi.KERNEL_PAIRED][k] for k in ("ax", "ay", "az")], axis=1)
    assert np.all(np.isfinite(b))
    assert _rel_rows(b, a).max() <= 1e-12


@pytest.mark.parametrize("dt_scale", [1, 40])
def test_paired_link_thresholds_survive_passes_and_large_moves(ctx, dt_scale, layout):
    """The pair-shared kernel filters candidates against thresholds left by the previous pass; a body whose
    most-attracting partner weakened by more than the slack (here: steps 40x too long, bodies move a lot) is
    re-ranked from scratch in k_pair_finish.  Links and max_attraction must stay those of the tiled kernel."""
    n = 8192
    w = synth.plummer(n, seed=51)
    rng = np.random.default_rng(52)
    gf = w["gf"] * rng.uniform(0.5, 2.0, n)
    dt = synth.default_dt(n) * dt_scale
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
        ctx.date = capi.START_DATE
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, most=np.full(n, -1, dtype=np.int32))
        hist = []
        for _ in range(5):
            ctx.step(1, dt, kernel=kern, track_most_attracting=True)
            d = ctx.download()
            hist.append((d["most"].copy(), d["max_attraction"].copy()))
        got[kern] = hist
    changed = 0
    for (mt, at), (mp, ap_) in zip(got[capi.KERNEL_TILED], got[capi.KERNEL_PAIRED]):
        assert np.array_equal(mt, mp)
        assert np.all(np.abs(ap_ - at) <= 1e-9 * np.abs(at))
    first, last = got[capi.KERNEL_PAIRED][0][0], got[capi.KERNEL_PAIRED][-1][0]
    changed = int((first != last).sum())
    if dt_scale > 1:
        assert changed > 50  # the long steps did reshuffle the links


def test_paired_tracks_most_attracting_when_it_is_a_no_op(ctx):
    """Equal gravity factors: `o.gf > this.gf` (src/Object.cpp:86,91) is false for every pair, so the reference leaves
    most_attracting_object alone and max_attraction at 0.  The pair-shared kernel accepts tracking then and must
    leave exactly what the tiled argmax kernel leaves."""
    n = 5000
    w = synth.plummer(n, seed=21)
    dt = synth.default_dt(n)
    got = {}
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
        ctx.date = capi.START_DATE
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], most=np.arange(n, dtype=np.int32)[::-1].copy())
        ctx.step(3, dt, kernel=kern, track_most_attracting=True)
        got[kern] = ctx.download()
    t, p = got[capi.KERNEL_TILED], got[capi.KERNEL_PAIRED]
    assert np.array_equal(t["most"], p["most"]) and np.array_equal(p["most"], np.arange(n, dtype=np.int32)[::-1])
    assert np.array_equal(t["max_attraction"], p["max_attraction"]) and not p["max_attraction"].any()
    pos_t = np.stack([t["x"], t["y"], t["z"]], axis=1)
    pos_p = np.stack([p["x"], p["y"], p["z"]], axis=1)
    assert _rel_rows(pos_p, pos_t).max() <= 1e-12


def _near_tie_world(n, t, a, b, delta, seed=5):
    """Light bodies of one common mass far away (x + 1e13 m), and near the origin a light target `t` with two heavy
    partners `a`, `b` whose attraction metrics gf / r^4 differ by the relative amount `delta` (a is the stronger one;
    delta == 0: mirrored positions, bit-identical metrics in the reference's arithmetic)."""
    w = synth.plummer(n, seed=seed)
    x, y, z = w["x"] + 1e13, w["y"].copy(), w["z"].copy()
    mass = np.full(n, 1e20)
    d = 1.0e9
    x[t], y[t], z[t] = 0.0, 0.0, 0.0
    x[a], y[a], z[a] = d, 0.0, 0.0
    x[b], y[b], z[b] = 0.0, d * (1.0 + delta / 4.0), 0.0
    mass[a] = mass[b] = 1e26
    return x, y, z, mass


@pytest.mark.parametrize("kernel", [capi.KERNEL_PAIRED, capi.KERNEL_TILED])
@pytest.mark.parametrize("delta", [1e-7, 1e-9, 1e-12, 0.0])
@pytest.mark.parametrize("order", ["stronger_first", "stronger_last"])
def test_most_attracting_near_ties_are_decided_as_the_reference_does(ctx, kernel, delta, order, layout):
    """src/Object.cpp:84-94 decides in full binary64 with a strict '>' in ascending index order.  The pair-shared kernel
    ranks candidates by sortable keys that keep 31 mantissa bits of the metric: two heavier partners within 5e-10 of each
    other (or exactly tied) must be re-decided exactly, wherever in the reduction they meet -- here the three bodies sit
    in three different 2048-body blocks, so the candidates meet in k_pair_finish / in the J-side slot merge."""
    n = 6144
    t, a, b = 100, 2500, 5000
    if order == "stronger_last":
        a, b = b, a  # the stronger partner has the HIGHER index: "lowest index wins" would be the wrong guess
    x, y, z, mass = _near_tie_world(n, t, a, b, delta)
    ow = oracle.World()
    ow.add_bodies(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), mass)
    ow.set_forces()
    ost = ow.state()
    want = int(ost["most"][t])
    assert want == (a if delta > 0 else min(a, b))  # the oracle itself: the stronger one; at the exact tie the lower index
    ctx.date = capi.START_DATE
    ctx.upload(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), ost["gf"], most=np.full(n, -1, dtype=np.int32))
    ctx.set_forces(kernel=kernel, track_most_attracting=True)
    got = ctx.download(("most", "max_attraction"))
    assert int(got["most"][t]) == want
    assert got["most"][a] == -1 and got["most"][b] == -1  # nobody is heavier than the two heavy ones
    # every link, not only the constructed one
    assert np.array_equal(got["most"], ost["most"])
    assert got["max_attraction"][t] == pytest.approx(ost["maxattr"][t], rel=1e-12)
    # ... and again after a second pass (the filter thresholds of the first pass are in force now)
    ctx.set_forces(kernel=kernel, track_most_attracting=True)
    assert np.array_equal(ctx.download(("most",))["most"], ost["most"])


def test_most_attracting_near_tie_inside_one_block_and_one_warp_visit(ctx, layout):
    """The same, with target and both partners inside one 32-body subset of one block (the candidates then meet in a
    lane's own running maximum, pair_exact) and in neighbouring subsets."""
    n = 4096
    for t, a, b in ((5, 9, 17), (5, 17, 9), (40, 70, 1030), (40, 1030, 70)):
        for delta in (1e-9, 1e-12, 0.0):
            x, y, z, mass = _near_tie_world(n, t, a, b, delta)
            ow = oracle.World()
            ow.add_bodies(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), mass)
            ow.set_forces()
            ost = ow.state()
            ctx.date = capi.START_DATE
            ctx.upload(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), ost["gf"], most=np.full(n, -1, dtype=np.int32))
            ctx.set_forces(kernel=capi.KERNEL_PAIRED, track_most_attracting=True)
            got = ctx.download(("most",))
            assert np.array_equal(got["most"], ost["most"]), (t, a, b, delta, got["most"][t], ost["most"][t])


@pytest.mark.parametrize("delta", [1e-7, 1e-9, 1e-12])
@pytest.mark.parametrize("order", ["stronger_first", "stronger_last"])
def test_most_attracting_near_ties_across_blocks_of_the_mass_sorted_layout(ctx, layout, delta, order):
    """Near ties between partners of DIFFERENT masses: in the mass-sorted layout the target, the lighter partner and the
    heavier partner land in different blocks (2,500 distant bodies of an intermediate mass lie between the two partners in
    mass order), so the two candidates meet in the slot / chunk merges of k_pair_finish.  The oracle decides."""
    n = 6144
    rng = np.random.default_rng(77)
    w = synth.plummer(n, seed=9)
    x, y, z = w["x"] + 1e13, w["y"].copy(), w["z"].copy()
    mass = 1e20 * rng.uniform(1.0, 2.0, n)
    far = rng.choice(np.arange(200, n), 2500, replace=False)
    mass[far] = 2e26  # heavier than partner a, lighter than partner b, ten thousand times farther away
    free = np.setdiff1d(np.arange(200, n), far)
    t, a, b = 100, int(free[10]), int(free[-10])
    if order == "stronger_last":
        a, b = b, a
    d = 1.0e9
    x[t], y[t], z[t] = 0.0, 0.0, 0.0
    x[a], y[a], z[a] = d, 0.0, 0.0                                      # mass 1e26 at d: the stronger one by `delta`
    x[b], y[b], z[b] = 0.0, d * np.sqrt(2.0) * (1.0 + delta / 4.0), 0.0  # mass 4e26 at sqrt(2) d (1 + delta / 4)
    mass[a], mass[b] = 1e26, 4e26
    ow = oracle.World()
    ow.add_bodies(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), mass)
    ow.set_forces()
    ost = ow.state()
    assert int(ost["most"][t]) in (a, b)
    ctx.date = capi.START_DATE
    ctx.upload(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), ost["gf"], most=np.full(n, -1, dtype=np.int32))
    for _ in range(2):  # the second pass runs against the thresholds the first one left
        ctx.set_forces(kernel=capi.KERNEL_PAIRED, track_most_attracting=True)
        got = ctx.download(("most", "max_attraction"))
        assert np.array_equal(got["most"], ost["most"])
        assert np.allclose(got["max_attraction"], ost["maxattr"], rtol=1e-11, atol=0.0)


def test_mass_sorted_layout_with_mass_classes_and_massless_bodies(ctx, layout):
    """Worlds of a few mass classes (whole tiles of one class run without a filter), massless test particles and a ragged block: every link and
    every acceleration as the tiled kernel and the oracle's rule give them; the layout follows a change of the masses and of
    the active mask (a body created later joins, another upload with other masses)."""
    n = 7000
    w = synth.plummer(n, seed=21)
    rng = np.random.default_rng(22)
    gf = w["gf"] * rng.choice([1.0, 1.0, 1.0, 3.0, 10.0], n)
    gf[rng.choice(n, 30, replace=False)] = 0.0
    t0 = capi.START_DATE
    dt = synth.default_dt(n)
    cdate = np.full(n, t0, dtype=np.int64)
    late = rng.choice(n, 25, replace=False)
    cdate[late] = t0 + 2 * dt  # these join after two steps: the mask flips, the layout is sorted again
    for round_ in range(2):
        got = {}
        for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED):
            ctx.date = t0
            ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, creation_date=cdate, most=np.full(n, -1, dtype=np.int32))
            ctx.step(4, dt, kernel=kern, track_most_attracting=True)
            got[kern] = ctx.download()
        t, p = got[capi.KERNEL_TILED], got[capi.KERNEL_PAIRED]
        assert np.array_equal(t["most"], p["most"])
        assert (p["most"][late] >= 0).sum() >= 20  # the late bodies take part by now
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            assert np.allclose(p[k], t[k], rtol=1e-10, atol=0.0), k
        m_t, m_p = t["max_attraction"], p["max_attraction"]
        assert np.all(np.abs(m_p - m_t) <= 1e-11 * np.abs(m_t))
        gf = gf * rng.uniform(0.5, 2.0, n)  # second round: every mass different, another order


@pytest.mark.parametrize("budget_mb", ["1", "4", None])
def test_paired_offsets_in_batches_when_the_slots_exceed_the_budget(budget_mb):
    """The J-side slots of all offsets cost N * (nb / 2) * 24 bytes: 103 GB at 4 M bodies.  Above the slot budget (32 GiB) a
    single GPU walks the offsets in batches and folds each batch into a running sum (k_pair_fold).  A small budget pushes
    10,000 .. 65,536-body worlds through 2 .. 16 batches: same bar against the oracle as the unbatched pass, momentum
    conserved, deterministic, trajectories equal to the tiled kernel's.  (Subprocess: the budget is read once per process.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("ESSA_PAIR_SLOT_BUDGET_MB", None)
    if budget_mb:
        env["ESSA_PAIR_SLOT_BUDGET_MB"] = budget_mb
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "batched_check.py")], env=env, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("batched_check")]
    assert p.returncode == 0 and len(lines) == 7 and all(ln.endswith("PASS") for ln in lines), p.stdout + p.stderr
