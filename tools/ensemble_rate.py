#!/usr/bin/env python3
"""65,536-copy ensembles of solar.essa and jupiter_system.essa through the C ABI: copy-iterations/s,
interactions/s and the fraction of the FP64 FMA peak measured in the same process (20 flops/interaction)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, world  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
c = capi.Context(0)
peak = c.measure_fp64_peak()
print("fp64 peak measured: %.2f TFLOP/s" % peak)
only = sys.argv[2] if len(sys.argv) > 2 else ""
for fname, steps in (("solar.essa", 2000), ("jupiter_system.essa", 100)):
    if only and not fname.startswith(only):
        continue
    pw = world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", fname))
    st = pw.state()
    n = len(pw)
    for approaches in (False, True):
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
        c.ensemble_init(copies, 600, seed=7, rel_sigma=1e-6, closest_approaches=approaches, approach_to=0)
        c.ensemble_step(2)
        c.ensemble_step(steps)
        ms = c.last_step_ms
        inter = copies * (steps + 1) * n * (n - 1) / (ms * 1e-3)
        print("%-20s approaches=%d  %.3e copy-it/s  %.3e interactions/s  %.1f %% of FP64 peak" % (
            fname, approaches, copies * steps / (ms * 1e-3), inter, 100 * inter * 20 / (peak * 1e12)), flush=True)
c.close()
