#!/usr/bin/env python3
"""One force pass of an N-body Plummer world (argv[1]) through the C ABI: for `ncu --metrics gpu__time_duration.sum` launch
lists at sizes where a whole bench run under the profiler would take too long."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1])
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = synth.plummer(n, seed=1)
c = capi.Context(0)
c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
for _ in range(passes):
    t0 = time.perf_counter()
    c.set_forces()
    print("n=%d set_forces %.1f ms" % (n, (time.perf_counter() - t0) * 1e3), flush=True)
c.close()
