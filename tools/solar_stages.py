#!/usr/bin/env python3
"""solar.essa through the C ABI: physics only / + most-attracting tracking / + recording (stage timing)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, world  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
fname = sys.argv[2] if len(sys.argv) > 2 else "solar.essa"
pw = world.World.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "worlds", fname))
st = pw.state()
caps = [pw.props(i)["trail_capacity"] for i in range(len(pw))]
n = len(pw)
persist = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # 1: the persistent flavour of K3q (same step loop, doorbell between calls)
only = sys.argv[4] if len(sys.argv) > 4 else None  # run one stage only
for label, kw in (("physics", {}), ("physics+attraction", dict(track_most_attracting=True)), ("all", dict(record=True))):
    if only and label != only:
        continue
    c = capi.Context(0)
    c.persist_configure(persist)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
             creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
    c.record_configure(caps)
    c.step(1000, 600, **kw)
    c.step(steps, 600, **kw)
    ms = c.last_step_ms
    print("%-20s %9.0f it/s  %6.0f clk/step" % (label, steps / (ms * 1e-3), ms * 1e-3 * 1.965e9 / steps), flush=True)
    c.close()
