#!/usr/bin/env python3
"""solar.essa through the World facade: `steps` KDK iterations in one World::update call (profiling target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import world  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
exact = len(sys.argv) > 2 and sys.argv[2] == "exact"
fname = sys.argv[3] if len(sys.argv) > 3 else "solar.essa"
pw = world.World.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "worlds", fname))
pw.simulation_seconds_per_tick = 600
pw.set_flags(offset_trails=True, exact_order=exact)
pw.update(1000)
pw.update(steps)
print("%s steps=%d exact=%s it/s=%.0f" % (fname, steps, exact, steps / (pw.last_step_ms * 1e-3)))
