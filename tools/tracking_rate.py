#!/usr/bin/env python3
"""Force-pass rate at N = 262,144 (or argv[1]) with the most-attracting bookkeeping off / on: equal masses (pair-shared
UNIFORM kernel; the bookkeeping is provably a no-op) vs masses spread over a factor 4 (general kernel; with tracking
the ARGMAX instantiation: candidate keys behind the integer filter)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
only = sys.argv[2] if len(sys.argv) > 2 else None  # e.g. "unequal:1" = unequal masses with tracking only (for ncu)
c = capi.Context(0, n)
peak = c.measure_fp64_peak()
w = synth.plummer(n, seed=5)
dt = synth.default_dt(n)
for label, gf in (("equal masses", w["gf"]), ("unequal masses", w["gf"] * np.random.default_rng(1).uniform(0.5, 2.0, n))):
    for track in (False, True):
        if only and only != "%s:%d" % (label.split()[0], track):
            continue
        c.date = capi.START_DATE
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        c.step(2, dt, track_most_attracting=track)
        c.step(3, dt, track_most_attracting=track)
        ms = c.last_step_ms / 3
        rate = n * (n - 1) / (ms * 1e-3)
        print("%-15s tracking=%d  %.3e interactions/s  %.2f ms/step  %.1f %% of FP64 peak" % (label, track, rate, ms, 100 * rate * 20 / (peak * 1e12)), flush=True)
        c.step(1, dt, track_most_attracting=track)
        print("    phases of one pass (ms): %s" % {k: round(v, 3) for k, v in c.last_pass_phases.items()}, flush=True)
c.close()
