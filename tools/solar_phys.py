#!/usr/bin/env python3
"""solar.essa through the C ABI without recording (physics only): profiling target."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, world  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pw = world.World.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "worlds", "solar.essa"))
st = pw.state()
c = capi.Context(0)
n = len(pw)
c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
         creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
c.step(1000, 600)
c.step(steps, 600)
print("physics only: %.0f it/s" % (steps / (c.last_step_ms * 1e-3)))
