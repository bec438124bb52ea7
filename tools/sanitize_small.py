#!/usr/bin/env python3
"""A short tour of every kernel family for compute-sanitizer (small sizes, few steps):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth, world  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    return world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", name))


for name, steps in (("solar.essa", 120), ("jupiter_system.essa", 40), ("8body.essa", 30)):
    w = load(name)
    w.simulation_seconds_per_tick = 600
    w.update(steps)          # K3q / K3x with recording
    w.update(-steps // 2)
    w.set_flags(exact_order=True)
    w.update(5)              # K3w exact / K3 exact
    print(name, "ok", w.get_object(0).pos if hasattr(w, "get_object") else "", flush=True)

c = capi.Context(0)
for n in (40, 128, 200):     # K3x PER = 2 / 4, single-CTA K3
    p = synth.plummer(n, seed=n)
    c.date = capi.START_DATE
    c.upload(p["x"], p["y"], p["z"], p["vx"], p["vy"], p["vz"], p["gf"] * np.linspace(1, 3, n))
    c.record_configure([500] * n)
    c.step(25, 1000, record=True)
    print("resident", n, "ok", flush=True)
st = load("solar.essa").state()
c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
c.ensemble_init(100, 600, seed=7, rel_sigma=1e-6, closest_approaches=True, approach_to=0)
c.ensemble_step(20)          # warp-synchronous ensemble
st = load("jupiter_system.essa").state()
c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
c.ensemble_init(10, 600, seed=7, rel_sigma=1e-6)
c.ensemble_step(5)           # CTA ensemble, ring order
print("ensembles ok", flush=True)
for n, kern in ((3000, capi.KERNEL_TILED), (4500, capi.KERNEL_PAIRED)):
    p = synth.plummer(n, seed=n)
    c.date = capi.START_DATE
    c.upload(p["x"], p["y"], p["z"], p["vx"], p["vy"], p["vz"], p["gf"])
    c.step(2, 1000, kernel=kern, track_most_attracting=True)
    print("force pass", n, "ok", flush=True)
print(c.energy())
c.close()
