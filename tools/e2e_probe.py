#!/usr/bin/env python3
"""Where the end-to-end tick (essa_upload + essa_step(1) + essa_download with host buffers) spends its time."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
w = synth.galaxy_collision(n) if n >= (1 << 20) else synth.plummer(n)
dt = synth.default_dt(n)
c = capi.Context(0, n)
opts = capi.Context.opts(dt, track_most_attracting=True)
host = {k: np.array(w[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "gf")}
for it in range(4):
    t0 = time.perf_counter()
    c.upload(host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"], host["gf"])
    t1 = time.perf_counter()
    c.step(1, opts=opts)
    t2 = time.perf_counter()
    c.download(("x", "y", "z", "vx", "vy", "vz"), out=host)
    t3 = time.perf_counter()
    print("tick %d: upload %.1f ms, step %.1f ms (device %.1f, force %.1f in %d passes), download %.1f ms" % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), c.last_step_ms, c.last_force_ms, c.last_force_passes, 1e3 * (t3 - t2)), flush=True)
for it in range(2):
    t1 = time.perf_counter()
    c.step(1, opts=opts)
    t2 = time.perf_counter()
    print("resident step: %.1f ms (device %.1f, force %.1f in %d passes)" % (1e3 * (t2 - t1), c.last_step_ms, c.last_force_ms, c.last_force_passes), flush=True)
c.close()
