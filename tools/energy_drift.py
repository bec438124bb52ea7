#!/usr/bin/env python3
"""Energy / momentum drift of the KDK path on BASELINE.json's synthetic worlds:
|dE/E0| as a function of step (no softening, fixed integer dt -- close pairs dominate), and the
GPU-vs-oracle energy difference after 10 steps at N = 16,384 (the parity-relevant quantity)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/energy_drift.jsonl", "a") as f:
            f.write(line + "\n")


def main():
    with_oracle = "--oracle" in sys.argv
    ctx = capi.Context(0)
    for name, n, steps, every in (("plummer", 16384, 200, 20), ("plummer", 262144, 100, 10), ("galaxy", 1048576, 40, 10)):
        w = synth.plummer(n) if name == "plummer" else synth.galaxy_collision(n)
        dt = synth.default_dt(n)
        ctx.date = capi.START_DATE
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        ke0, pe0, p0 = ctx.energy()
        e0 = ke0 + pe0
        pscale = np.abs(w["vx"]).sum() * synth.BODY_MASS
        t0 = time.time()
        for s in range(every, steps + 1, every):
            ctx.step(every, dt)
            ke, pe, p = ctx.energy()
            emit(world=name, n=n, dt=dt, step=s, rel_dE=abs(ke + pe - e0) / abs(e0), virial=-2 * ke / pe,
                 rel_dP=float(np.abs(p - p0).max() / pscale), wall_s=round(time.time() - t0, 2))
    if with_oracle:
        import oracle
        n, steps = 16384, 10
        w = synth.plummer(n)
        dt = synth.default_dt(n)
        ctx.date = capi.START_DATE
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        ctx.step(steps, dt)
        g = ctx.download(("x", "y", "z", "vx", "vy", "vz"))
        t0 = time.time()
        (ox, oy, oz, ovx, ovy, ovz), _ = oracle.step_soa(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"], steps, dt)
        cpu_s = time.time() - t0
        eg = oracle.energy(g["x"], g["y"], g["z"], g["vx"], g["vy"], g["vz"], w["gf"])
        eo = oracle.energy(ox, oy, oz, ovx, ovy, ovz, w["gf"])
        pos_g = np.stack([g["x"], g["y"], g["z"]], axis=1)
        pos_o = np.stack([ox, oy, oz], axis=1)
        rel = np.linalg.norm(pos_g - pos_o, axis=1) / np.linalg.norm(pos_o, axis=1)
        emit(what="gpu_vs_oracle", n=n, steps=steps, dt=dt, rel_dE_gpu_vs_oracle=abs((eg[0] + eg[1]) - (eo[0] + eo[1])) / abs(eo[0] + eo[1]),
             pos_rel_max=float(rel.max()), pos_rel_p99=float(np.percentile(rel, 99)), oracle_cpu_s=round(cpu_s, 1))


if __name__ == "__main__":
    main()
