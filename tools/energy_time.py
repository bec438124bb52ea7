import sys, time
sys.path.insert(0, '/root/repo')
from essa_b200 import capi, synth
import numpy as np
c = capi.Context(0)
for n in (262144, 1048576):
    w = synth.plummer(n, seed=1)
    c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    c.energy()
    t0 = time.perf_counter(); e = c.energy(); t1 = time.perf_counter()
    print("energy n=%d: %.1f ms  ke=%.6e pe=%.6e" % (n, (t1 - t0) * 1e3, e[0], e[1]), flush=True)
