#!/usr/bin/env python3
"""One 65,536-copy ensemble launch of a template world (for ncu): python tools/ensemble_one.py jupiter_system.essa 50"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, world  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fname = sys.argv[1] if len(sys.argv) > 1 else "jupiter_system.essa"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
pw = world.World.load(os.path.join(ROOT, "tests", "golden", "worlds", fname))
st = pw.state()
c = capi.Context(0)
c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
c.ensemble_init(65536, 600, seed=7, rel_sigma=1e-6)
c.ensemble_step(steps)
print("%.3f ms" % c.last_step_ms)
c.close()
