#!/usr/bin/env python3
"""The GUI's cadence (SimulationView::update: `iterations` steps per tick, state read back every tick):
microseconds per World::update(k) call through the facade, solar.essa, recording on."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import world  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fname = sys.argv[1] if len(sys.argv) > 1 else "solar.essa"
pw = world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", fname))
pw.simulation_seconds_per_tick = 600
pw.set_flags(offset_trails=True)
pw.update(1000)
for k in (1, 10, 100, 1000):
    t0 = time.perf_counter()
    for _ in range(200):
        pw.update(k)
    dt = (time.perf_counter() - t0) / 200
    print("%s update(%4d): %7.1f us per call  %9.0f it/s  (kernel %6.1f us)" % (fname, k, dt * 1e6, k / dt, pw.last_step_ms * 1e3), flush=True)
pw.close()
