#!/usr/bin/env python3
"""us per step of a mid-size world through K1 (one launch per step), K1g (one cooperative launch per call, force_grid.cu) and,
up to 1,024 bodies, the resident cluster kernels (AUTO); tracking off / on.   python tools/mid_probe.py 512 1024 2048 4096 6144"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [2048]
c = capi.Context(0)
for n in sizes:
    w = synth.plummer(n, seed=n)
    row = []
    for label, kern in (("tiled", capi.KERNEL_TILED), ("grid", capi.KERNEL_GRID), ("auto", capi.KERNEL_AUTO)):
        for track in (False, True):
            c.date = capi.START_DATE
            c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"] * (1.0 + (track and 1e-3) * (synth.np.arange(n) % 7)))
            c.step(5, 1000, kernel=kern, track_most_attracting=track)
            c.step(100, 1000, kernel=kern, track_most_attracting=track)
            row.append("%s%s %.2f" % (label, "+links" if track else "", c.last_step_ms * 1e3 / 100))
    print("n=%d  us/step:  %s" % (n, "   ".join(row)), flush=True)
c.close()
