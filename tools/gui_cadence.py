#!/usr/bin/env python3
"""The GUI's cadence (SimulationView::update: `iterations` steps per tick, state read back every tick,
src/SimulationView.cpp:374-397): microseconds per World::update(k) call through the facade, recording on,
with the persistent resident kernel (mailbox doorbell) and with one launch per call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import world  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fname = sys.argv[1] if len(sys.argv) > 1 else "solar.essa"
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"  # the persistent mode at 1 and 10 iterations per call only
for persist in ((True,) if quick else (True, False)):
    pw = world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", fname))
    pw.simulation_seconds_per_tick = 600
    pw.set_flags(offset_trails=True)
    pw.set_persistent(persist)
    pw.update(1000)
    for k in ((1, 10) if quick else (1, 10, 100, 1000)):
        for _ in range(20):
            pw.update(k)
        s0 = pw.persist_stats()
        t0 = time.perf_counter()
        for _ in range(500):
            pw.update(k)
        dt = (time.perf_counter() - t0) / 500
        s1 = pw.persist_stats()
        print("%s persistent=%d update(%4d): %7.2f us per call  %9.0f it/s  (device %6.2f us; host clock: doorbell->result %5.2f us, "
              "result->next doorbell %5.2f us)  %s"
              % (fname, persist, k, dt * 1e6, k / dt, pw.last_step_ms * 1e3, (s1["wait_ns"] - s0["wait_ns"]) / 500e3,
                 (s1["host_ns"] - s0["host_ns"]) / 500e3, {k2: s1[k2] for k2 in ("ticks", "launches", "idle_exits")}), flush=True)
    pw.close()
