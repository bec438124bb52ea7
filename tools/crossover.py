#!/usr/bin/env python3
"""Where the per-step-launch tiled path overtakes the resident kernels: us per step for Plummer spheres of
n bodies, 200 steps per call, fast arithmetic, no recording."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402

c = capi.Context(0)
for n in (40, 64, 96, 128, 160, 192, 256, 384, 512, 513, 768, 1024):
    w = synth.plummer(n, seed=n)
    row = []
    for kern in (capi.KERNEL_RESIDENT, capi.KERNEL_TILED):
        c.date = capi.START_DATE
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        c.step(20, 1000, kernel=kern)
        c.step(200, 1000, kernel=kern)
        row.append(c.last_step_ms * 1e3 / 200)
    print("n=%5d  resident %8.2f us/step   tiled %8.2f us/step" % (n, row[0], row[1]), flush=True)
# mid sizes: one launch per step either way (tiled K1 / pair-shared K1p)
for n in (1025, 2048, 4096, 8192, 12288, 16383, 16384):
    w = synth.plummer(n, seed=n)
    row = []
    for kern in (capi.KERNEL_TILED, capi.KERNEL_PAIRED, capi.KERNEL_AUTO):
        c.date = capi.START_DATE
        c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        try:
            c.step(20, 1000, kernel=kern)
            c.step(200, 1000, kernel=kern)
            row.append(c.last_step_ms * 1e3 / 200)
        except capi.EssaError:
            row.append(float("nan"))
    ideal = n * (n - 1.0) * 9 / 17.1e12 * 1e6  # 9 FP64 instructions per directed interaction at the measured DFMA issue rate
    print("n=%5d  tiled %8.2f us/step   pair-shared %8.2f us/step   auto %8.2f us/step   (FP64-pipe floor %.2f us)" % (n, row[0], row[1], row[2], ideal),
          flush=True)
c.close()
