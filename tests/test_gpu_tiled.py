"""GPU parity of K1 (j-tiled force pass + fused kick/drift) and K5 (energy) against the CPU oracle,
through the C ABI.  Tolerances follow SURVEY.md section 8c:

(ii)  single force pass: ||a_gpu - a_ld|| / ||a_ld|| per body <= max(1e-13, 4 x the oracle's own
      binary64-vs-extended error on the same input)  -- the GPU may differ from the reference about
      as much as the reference differs from exact arithmetic, not more;
(iii) trajectories after K = 10 steps: <= 1e-10 relative on positions and velocities (Plummer
      spheres are unsoftened; close pairs amplify round-off).
"""
import numpy as np
import pytest

import oracle
from essa_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _rel_rows(a, b):
    """Row-wise ||a - b|| / ||b||."""
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def _upload(ctx, w):
    ctx.date = capi.START_DATE
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [1, 2, 3, 255, 257, 1000, 4096, 16384])  # 16,384: BASELINE.json's tiled-kernel config
def test_force_pass_matches_extended_precision(ctx, n):
    w = synth.plummer(n, seed=n)
    _upload(ctx, w)
    ctx.set_forces(kernel=capi.KERNEL_TILED)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    if n == 1:
        assert np.all(a == 0.0)
        return
    rows = np.arange(n) if n <= 1000 else np.random.default_rng(1).choice(n, 256, replace=False)
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
    ref = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ref")
    err_gpu = _rel_rows(a[rows], ld)
    err_ref = _rel_rows(ref, ld)
    tol = max(1e-13, 4.0 * err_ref.max())
    assert err_gpu.max() <= tol, (err_gpu.max(), err_ref.max())


def test_force_pass_large_sampled_rows(ctx):
    """N = 65,536 (several j-chunks, R = 2 targets per thread): 256 sampled rows."""
    n = 65536
    w = synth.plummer(n, seed=5)
    _upload(ctx, w)
    ctx.set_forces(kernel=capi.KERNEL_TILED)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    rows = np.random.default_rng(2).choice(n, 256, replace=False)
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
    ref = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ref")
    assert _rel_rows(a[rows], ld).max() <= max(1e-13, 4.0 * _rel_rows(ref, ld).max())
    # Newton's third law: total momentum change rate sum(m a) ~ 0 relative to sum |m a|
    tot = np.abs((a * w["gf"][:, None]).sum(axis=0)).max()
    assert tot <= 1e-11 * np.abs(a * w["gf"][:, None]).sum()


def test_coincident_and_massless_bodies(ctx):
    """r = 0 pairs contribute nothing (the reference would produce NaN); gf = 0 bodies attract nothing."""
    x = np.array([0.0, 0.0, 1e9, 2e9])
    y = np.zeros(4)
    z = np.zeros(4)
    gf = np.array([1e20, 1e20, 0.0, 1e20])
    ctx.upload(x, y, z, y, y, y, gf)
    ctx.set_forces(kernel=capi.KERNEL_TILED)
    a = ctx.download(("ax", "ay", "az"))
    assert np.all(np.isfinite(a["ax"]))
    assert a["ax"][0] == a["ax"][1] == pytest.approx(1e20 / (2e9) ** 2, rel=1e-14)
    assert a["ax"][2] == pytest.approx(-2e20 / 1e18 + 1e20 / 1e18, rel=1e-14)


@pytest.mark.parametrize("n,steps", [(1024, 10), (1024, -10), (3000, 4)])
def test_kdk_trajectory_matches_oracle(ctx, n, steps):
    w = synth.plummer(n, seed=11)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(steps, dt, kernel=capi.KERNEL_TILED)
    got = ctx.download()
    (ox, oy, oz, ovx, ovy, ovz), oacc = oracle.step_soa(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], steps, dt)
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    vel = np.stack([got["vx"], got["vy"], got["vz"]], axis=1)
    acc = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    assert _rel_rows(pos, np.stack([ox, oy, oz], axis=1)).max() <= 1e-10
    assert _rel_rows(vel, np.stack([ovx, ovy, ovz], axis=1)).max() <= 1e-10
    assert _rel_rows(acc, oacc.T).max() <= 1e-9
    assert ctx.date == capi.START_DATE + steps * dt


def test_two_pass_equals_reuse_bitwise(ctx):
    """Reusing the previous step's second set_forces() is exactly what the reference recomputes."""
    n = 2048
    w = synth.plummer(n, seed=3)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(5, dt, kernel=capi.KERNEL_TILED)
    a = ctx.download()
    p1 = ctx.last_force_passes
    _upload(ctx, w)
    ctx.step(5, dt, kernel=capi.KERNEL_TILED, two_pass=True)
    b = ctx.download()
    assert p1 == 6 and ctx.last_force_passes == 10
    for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az"):
        assert np.array_equal(a[k], b[k]), k


def test_split_calls_equal_one_call_bitwise(ctx):
    n = 1500
    w = synth.plummer(n, seed=4)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(6, dt, kernel=capi.KERNEL_TILED)
    a = ctx.download()
    _upload(ctx, w)
    for _ in range(3):
        ctx.step(2, dt, kernel=capi.KERNEL_TILED)
    b = ctx.download()
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(a[k], b[k]), k


def test_reversibility(ctx):
    """update(+K) then update(-K) returns to the start to round-off (SURVEY.md section 4 iii)."""
    n = 1024
    w = synth.plummer(n, seed=6)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ctx.step(8, dt, kernel=capi.KERNEL_TILED)
    ctx.step(-8, dt, kernel=capi.KERNEL_TILED)
    got = ctx.download(("x", "y", "z"))
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    ref = np.stack([w["x"], w["y"], w["z"]], axis=1)
    assert _rel_rows(pos, ref).max() <= 1e-9
    assert ctx.date == capi.START_DATE


def test_energy_matches_oracle_and_is_conserved(ctx):
    n = 4096
    w = synth.plummer(n, seed=8)
    dt = synth.default_dt(n)
    _upload(ctx, w)
    ke, pe, mom = ctx.energy()
    oke, ope, omom = oracle.energy(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    assert ke == pytest.approx(oke, rel=1e-12)
    assert pe == pytest.approx(ope, rel=1e-12)
    assert 0.4 < -2 * ke / pe < 0.6 or 0.8 < -2 * ke / pe < 1.2  # near virial equilibrium (2K + W ~ 0)
    scale = np.abs(w["vx"]).sum() * synth.BODY_MASS
    assert np.abs(mom).max() <= 1e-9 * scale
    ctx.step(20, dt, kernel=capi.KERNEL_TILED)
    ke2, pe2, mom2 = ctx.energy()
    assert abs((ke2 + pe2) - (ke + pe)) <= 1e-3 * abs(ke + pe)
    assert np.abs(mom2 - mom).max() <= 1e-9 * scale


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 512, 513, 769, 1000, 1025, 1537])
def test_energy_pair_enumeration_edges(ctx, n):
    """K5 meets every unordered pair once (tile T against tiles T + d, d = 0 .. ntiles / 2): one tile, two tiles (the opposite
    tile met from the lower half only), odd and even rings, ragged last tiles -- each against the oracle's long-double sum."""
    w = synth.plummer(n, seed=100 + n)
    _upload(ctx, w)
    ke, pe, mom = ctx.energy()
    oke, ope, omom = oracle.energy(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    assert ke == pytest.approx(oke, rel=1e-12)
    assert pe == pytest.approx(ope, rel=1e-12, abs=0.0)


def test_most_attracting_tiled_matches_oracle(ctx, worlds_dir):
    """Argmax variant of K1 on a world with a mass hierarchy (jupiter_system.essa, 77 bodies)."""
    import os
    ow = oracle.World.load(os.path.join(worlds_dir, "jupiter_system.essa"))
    st = ow.state()
    ctx.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
    ctx.set_forces(kernel=capi.KERNEL_TILED, track_most_attracting=True)
    ow.set_forces()
    ref = ow.state()
    got = ctx.download(("most", "max_attraction", "ax", "ay", "az"))
    assert np.array_equal(got["most"], ref["most"])
    assert np.allclose(got["max_attraction"], ref["maxattr"], rtol=1e-12, atol=0)
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    assert _rel_rows(a[1:], ref["acc"][1:]).max() <= 1e-12


def test_active_mask_by_date(ctx):
    """Object::deleted(): bodies created in the future or deleted in the past neither attract nor move."""
    n = 600
    w = synth.plummer(n, seed=9)
    dt = 1000000
    t0 = capi.START_DATE
    cdate = np.full(n, t0, dtype=np.int64)
    ddate = np.zeros(n, dtype=np.int64)
    deleted = np.zeros(n, dtype=np.uint8)
    cdate[5] = t0 + 3 * dt          # appears at step 3
    deleted[7] = 1
    ddate[7] = t0 + 2 * dt          # disappears at step 2
    ctx.date = t0
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], creation_date=cdate, deletion_date=ddate, deleted=deleted)
    ctx.step(5, dt, kernel=capi.KERNEL_TILED)
    got = ctx.download()
    ow = oracle.World()
    ow.add_bodies(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], np.full(n, synth.BODY_MASS))
    ow.dt = dt
    ow.set_dates(5, cdate[5])
    ow.set_dates(7, cdate[7], ddate[7], True)
    ow.update(5)
    ref = ow.state()
    pos = np.stack([got["x"], got["y"], got["z"]], axis=1)
    vel = np.stack([got["vx"], got["vy"], got["vz"]], axis=1)
    assert _rel_rows(pos, ref["pos"]).max() <= 1e-10
    assert _rel_rows(vel, ref["vel"]).max() <= 1e-10
    # body 7 froze after step 1 (date t0+2dt masks it from step 2 on)
    assert not np.array_equal(pos[7], [w["x"][7], w["y"][7], w["z"][7]])


def test_tracked_prefix_recording_on_a_large_world(ctx):
    """Tiled path with recording: only a tracked prefix of a big world gets History / Trail rings
    (48 kB of History per body would not scale); checked against the oracle, which records every body."""
    n, tracked, steps = 3000, 8, 12
    w = synth.plummer(n, seed=13)
    # make the first body heavy so that the tracked ones have a most-attracting object
    w["gf"] = w["gf"].copy()
    w["gf"][0] *= 500.0
    dt = synth.default_dt(n)
    t0 = capi.START_DATE
    ctx.date = t0
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], creation_date=np.full(n, t0, dtype=np.int64))
    ctx.record_configure([500] * tracked)
    ctx.step(steps, dt, kernel=capi.KERNEL_TILED, record=True)
    got = ctx.download()
    ow = oracle.World()
    ow.add_bodies(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"] / oracle.GRAVITY)
    for b in range(n):
        if b == 0:
            ow.set_object_state(0, gf=w["gf"][0])
    ow.dt = dt
    ow.update(steps)
    ref = ow.state()
    assert np.array_equal(got["most"], ref["most"])
    assert np.count_nonzero(got["most"][:tracked] == 0) >= tracked - 1
    for k in range(tracked):
        h, ho = ctx.history(k), ow.history(k)
        assert h.shape == ho.shape == (steps + 1, 6)
        assert np.allclose(h, ho, rtol=1e-10, atol=0)
        t, to = ctx.trail(k), ow.trail(k)
        assert (t["append_offset"], t["length"]) == (to["append_offset"], to["length"])
        assert np.allclose(t["verts"][:t["length"]], to["verts"][:to["length"]], rtol=1e-5, atol=1e-6)
        assert np.allclose(t["offset"], to["offset"], rtol=1e-10)
    assert np.allclose(got["ap"][1:tracked], ref["appe"][1:tracked, 0], rtol=1e-10)
    assert np.allclose(got["pe"][1:tracked], ref["appe"][1:tracked, 1], rtol=1e-10)
    ctx.record_configure([])
