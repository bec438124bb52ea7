"""BASELINE.json's full sizes (Plummer N = 262,144; galaxy collision N = 1,048,576) through
size-independent properties: sampled target rows against the oracle's extended-precision evaluation
over ALL sources (SURVEY.md 8c iv), Newton's third law (sum m a = 0), +K / -K reversibility, and the
energy diagnostic against the oracle on a sampled sub-problem."""
import numpy as np
import pytest

import oracle
from essa_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.mark.parametrize("name,n,kernel", [("plummer", 16384, capi.KERNEL_AUTO), ("plummer", 16384, capi.KERNEL_TILED),
                                           ("plummer", 16383, capi.KERNEL_AUTO),  # ragged last block of the pair-shared kernel
                                           ("plummer", 12288, capi.KERNEL_AUTO), ("plummer", 12289, capi.KERNEL_AUTO),  # either side of AUTO's switch
                                           ("plummer", 262144, capi.KERNEL_AUTO), ("plummer", 262144, capi.KERNEL_TILED),
                                           ("galaxy", 1048576, capi.KERNEL_AUTO)])
def test_full_size_force_pass_and_reversibility(name, n, kernel):
    w = synth.plummer(n) if name == "plummer" else synth.galaxy_collision(n)
    dt = synth.default_dt(n)
    ctx = capi.Context(0, n)
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    ctx.set_forces(kernel=kernel)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    rows = np.sort(np.random.default_rng(7).choice(n, 48, replace=False))
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
    ref, _ = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="omp")
    err, err_ref = _rel_rows(a[rows], ld), _rel_rows(ref, ld)
    assert err.max() <= max(1e-13, 4.0 * err_ref.max()), (err.max(), err_ref.max())
    ma = a * w["gf"][:, None]
    assert np.abs(ma.sum(axis=0)).max() <= 1e-11 * np.abs(ma).sum()
    # update(+2); update(-2) returns to the start to round-off (close pairs amplify it: 1e-8 bar at this N)
    ctx.step(2, dt, kernel=kernel)
    mid = ctx.download(("x",))
    assert not np.array_equal(mid["x"], w["x"])
    ctx.step(-2, dt, kernel=kernel)
    back = ctx.download(("x", "y", "z"))
    pos = np.stack([back["x"], back["y"], back["z"]], axis=1)
    start = np.stack([w["x"], w["y"], w["z"]], axis=1)
    rel = _rel_rows(pos, start)
    assert np.percentile(rel, 99) <= 1e-12 and rel.max() <= 1e-8, (np.percentile(rel, 99), rel.max())
    assert ctx.date == capi.START_DATE
    ctx.close()


def test_four_million_bodies_take_the_pair_shared_kernel_in_batches():
    """4,194,304 bodies: the J-side slots of all 1,024 offsets would be 103 GB, so the pass runs in four batches of 256 offsets
    against 26 GB of slots (force_paired.cu, k_pair_fold).  Sampled rows against the oracle's extended-precision sums, and
    sum(m a) ~ 0 over all four million bodies -- a missed or doubled batch would show in either."""
    n = 1 << 22
    w = synth.plummer(n)
    ctx = capi.Context(0, n)
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    ctx.set_forces(kernel=capi.KERNEL_PAIRED)
    got = ctx.download(("ax", "ay", "az"))
    a = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
    rows = np.sort(np.random.default_rng(9).choice(n, 24, replace=False))
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
    ref, _ = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="omp")
    err, err_ref = _rel_rows(a[rows], ld), _rel_rows(ref, ld)
    assert err.max() <= max(1e-13, 4.0 * err_ref.max()), (err.max(), err_ref.max())
    assert np.all(np.isfinite(a))
    ma = a * w["gf"][:, None]
    assert np.abs(ma.sum(axis=0)).max() <= 1e-11 * np.abs(ma).sum()
    ctx.close()


def test_energy_diagnostic_on_a_subproblem():
    """K5 against the oracle's long-double energy on the first 20,000 bodies of the 1 M world."""
    w = synth.galaxy_collision(1048576)
    m = 20000
    ctx = capi.Context(0)
    sub = {k: v[:m] for k, v in w.items()}
    ctx.upload(sub["x"], sub["y"], sub["z"], sub["vx"], sub["vy"], sub["vz"], sub["gf"])
    ke, pe, mom = ctx.energy()
    oke, ope, omom = oracle.energy(sub["x"], sub["y"], sub["z"], sub["vx"], sub["vy"], sub["vz"], sub["gf"])
    assert ke == pytest.approx(oke, rel=1e-12) and pe == pytest.approx(ope, rel=1e-12)
    assert np.allclose(mom, omom, rtol=1e-10, atol=1e-12 * np.abs(sub["vx"]).sum() * synth.BODY_MASS)
    ctx.close()


@pytest.mark.parametrize("n", [262144, 1048576])
def test_full_size_most_attracting_links_on_unequal_masses(n, monkeypatch):
    """BASELINE's large worlds with unequal masses, most-attracting links tracked (what every real .essa world asks for):
    the mass-sorted tracked pass against (i) a direct evaluation of src/Object.cpp:84-94 over ALL partners for sampled bodies,
    (ii) the same pass on the caller's order (the two layouts share nothing but the exact tests), link for link over the whole
    world, after two steps (so that the second pass filters against the thresholds the first one left), and (iii) the
    accelerations of sampled rows against the oracle's extended-precision sums."""
    w = synth.plummer(n) if n < (1 << 20) else synth.galaxy_collision(n)
    rng = np.random.default_rng(n)
    gf = w["gf"] * (1.0 + 3.0 * rng.random(n))
    dt = synth.default_dt(n)
    got = {}
    for layout in ("1", "0"):
        monkeypatch.setenv("ESSA_PAIR_SORTED", layout)
        ctx = capi.Context(0, n)
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf, most=np.full(n, -1, dtype=np.int32))
        ctx.step(2, dt, track_most_attracting=True)
        got[layout] = ctx.download(("x", "y", "z", "ax", "ay", "az", "most", "max_attraction"))
        ctx.close()
    s, u = got["1"], got["0"]
    assert np.array_equal(s["most"], u["most"])
    assert (s["most"] >= 0).sum() >= n - 1  # everybody but the heaviest has a link
    assert np.all(np.abs(s["max_attraction"] - u["max_attraction"]) <= 1e-12 * np.abs(u["max_attraction"]))
    for k in ("x", "y", "z"):
        assert np.allclose(s[k], u[k], rtol=1e-13, atol=0.0)
    x, y, z = s["x"], s["y"], s["z"]
    for i in rng.choice(n, 24, replace=False):
        r2 = (x - x[i]) ** 2 + (y - y[i]) ** 2 + (z - z[i]) ** 2
        r2[i] = np.inf
        metric = np.where(gf > gf[i], gf / r2 ** 2, 0.0)
        if metric.max() > 0:
            assert s["most"][i] == int(np.argmax(metric)), i
            assert s["max_attraction"][i] == pytest.approx(metric.max(), rel=1e-9)
    rows = np.sort(rng.choice(n, 16, replace=False))
    ld = oracle.forces_rows(x, y, z, gf, rows, mode="ld")
    a = np.stack([s["ax"], s["ay"], s["az"]], axis=1)
    assert _rel_rows(a[rows], ld).max() <= 1e-12
