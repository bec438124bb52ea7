#!/usr/bin/env python3
"""Pair-shared force pass with its offsets in BATCHES (force_paired.cu: J-side slots over the slot budget).  The default budget
(32 GiB) batches only worlds beyond 2,097,152 bodies; run with ESSA_PAIR_SLOT_BUDGET_MB=<small> to push small worlds through
the same code: every line must say PASS.  tests/test_gpu_paired.py runs this in a subprocess (the budget is read once)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from essa_b200 import capi, synth  # noqa: E402


def rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def acc_of(c):
    g = c.download(("ax", "ay", "az"))
    return np.stack([g["ax"], g["ay"], g["az"]], axis=1)


def main():
    budget = os.environ.get("ESSA_PAIR_SLOT_BUDGET_MB", "(default)")
    c = capi.Context(0)
    ok_all = True
    for n, unequal in ((10000, False), (20000, False), (20000, True), (65536, False), (65536, True), (16384, False)):
        w = synth.plummer(n, seed=n)
        gf = w["gf"].copy()
        if unequal:
            gf *= np.random.default_rng(3).uniform(1.0, 4.0, n)
        c.date = capi.START_DATE
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        c.set_forces(kernel=capi.KERNEL_PAIRED)
        a = acc_of(c)
        rows = np.unique(np.concatenate([np.random.default_rng(1).choice(n, 128, replace=False), [0, 2047, 2048, n - 1]]))
        ld = oracle.forces_rows(w["x"], w["y"], w["z"], gf, rows, mode="ld")
        ref = oracle.forces_rows(w["x"], w["y"], w["z"], gf, rows, mode="ref")
        err, err_ref = rel_rows(a[rows], ld).max(), rel_rows(ref, ld).max()
        ma = a * gf[:, None]
        mom = np.abs(ma.sum(axis=0)).max() / np.abs(ma).sum()
        # determinism: the same pass again, bit for bit
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        c.set_forces(kernel=capi.KERNEL_PAIRED)
        same = np.array_equal(a, acc_of(c))
        # a few KDK steps against the one-sided tiled kernel
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        c.step(3, synth.default_dt(n), kernel=capi.KERNEL_PAIRED)
        p = c.download(("x", "y", "z"))
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        c.step(3, synth.default_dt(n), kernel=capi.KERNEL_TILED)
        t = c.download(("x", "y", "z"))
        traj = max(np.abs(p[k] - t[k]).max() / np.abs(t[k]).max() for k in "xyz")
        good = err <= max(1e-13, 4.0 * err_ref) and np.all(np.isfinite(a)) and mom <= 1e-12 and same and traj <= 1e-12
        ok_all &= bool(good)
        print("batched_check budget_mb=%s n=%d unequal=%d err=%.2e (oracle's own %.2e) momentum=%.1e deterministic=%s traj_vs_tiled=%.1e -> %s"
              % (budget, n, unequal, err, err_ref, mom, same, traj, "PASS" if good else "FAIL"), flush=True)
    # candidate keys are not batched: tracked + unequal masses must be refused by the explicit kernel choice when batching is on,
    # and AUTO must still give the right links (through the tiled kernel)
    n = 65536  # 25 MB of slots (50 MB in 1024-body blocks): batched under any budget this tool is run with
    w = synth.plummer(n, seed=5)
    gf = w["gf"] * np.random.default_rng(4).uniform(1.0, 4.0, n)
    c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
    refused = False
    try:
        c.set_forces(kernel=capi.KERNEL_PAIRED, track_most_attracting=True)
    except capi.EssaError:
        refused = True
    c.set_forces(kernel=capi.KERNEL_AUTO, track_most_attracting=True)
    most_auto = c.download(("most",))["most"].copy()
    c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
    c.set_forces(kernel=capi.KERNEL_TILED, track_most_attracting=True)
    most_tiled = c.download(("most",))["most"]
    batching = budget != "(default)"
    good = np.array_equal(most_auto, most_tiled) and (refused == batching)
    ok_all &= bool(good)
    print("batched_check tracked+unequal refused=%s (batching %s) links_auto==tiled=%s -> %s" % (refused, batching, np.array_equal(most_auto, most_tiled),
                                                                                              "PASS" if good else "FAIL"), flush=True)
    c.close()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
