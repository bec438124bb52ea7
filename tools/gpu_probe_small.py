#!/usr/bin/env python3
"""On-box probe of the small-N paths: dependent-issue latencies, resident kernels on solar.essa and
jupiter_system.essa with recording on/off, fast/exact, and the ensemble kernel."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, world  # noqa: E402

WORLDS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "worlds")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/probe_small.jsonl", "a") as f:
            f.write(line + "\n")


def ctx_from(fname):
    pw = world.World.load(os.path.join(WORLDS, fname))
    st = pw.state()
    caps = [pw.props(i)["trail_capacity"] for i in range(len(pw))]
    c = capi.Context(0)
    n = len(pw)
    c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"],
             creation_date=np.full(n, capi.START_DATE, dtype=np.int64))
    c.record_configure(caps)
    return c, n


def main():
    c0 = capi.Context(0)
    emit(what="latency_cycles", **c0.measure_latency())
    c0.close()
    for fname, steps in (("solar.essa", 200000), ("jupiter_system.essa", 20000)):
        for exact in (False, True):
            for record in (False, True):
                c, n = ctx_from(fname)
                c.step(1000, 600, record=record, exact_order=exact)
                c.step(steps, 600, record=record, exact_order=exact)
                ms = c.last_step_ms
                emit(what="resident", world=fname, n=n, exact=exact, record=record, steps=steps, iters_per_s=steps / (ms * 1e-3),
                     cycles_per_step_at_1965MHz=ms * 1e-3 * 1.965e9 / steps)
                c.close()
    # ensemble: 65,536 copies of jupiter and of solar
    for fname, copies, steps in (("jupiter_system.essa", 65536, 20), ("solar.essa", 65536, 200)):
        c, n = ctx_from(fname)
        c.ensemble_init(copies, 600, seed=7, rel_sigma=1e-6)
        c.ensemble_step(2)
        c.ensemble_step(steps)
        ms = c.last_step_ms
        passes = steps + 1
        emit(what="ensemble", world=fname, n=n, copies=copies, steps=steps, ms=ms, copy_steps_per_s=copies * steps / (ms * 1e-3),
             interactions_per_s=copies * passes * n * (n - 1) / (ms * 1e-3))
        c.close()


if __name__ == "__main__":
    main()
