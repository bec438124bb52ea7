import sys
sys.path.insert(0, '/root/repo')
from essa_b200 import capi, synth
n = int(sys.argv[1])
c = capi.Context(0)
w = synth.plummer(n, seed=n)
c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
c.step(5, 1000, kernel=capi.KERNEL_TILED)
c.step(20, 1000, kernel=capi.KERNEL_TILED)
print(c.last_step_ms * 1e3 / 20)
