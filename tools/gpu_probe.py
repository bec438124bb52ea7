#!/usr/bin/env python3
"""Quick on-box probe: FP64 peak, rsqrt seed accuracy, K1 force-pass rate at several N.
Writes one JSON object per line to stdout (and gpurun_out/probe.jsonl when that directory exists)."""
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
        with open("gpurun_out/probe.jsonl", "a") as f:
            f.write(line + "\n")


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 32768, 65536, 262144, 1048576]
    ctx = capi.Context(0)
    peak = ctx.measure_fp64_peak(300.0)
    emit(what="fp64_peak_tflops", value=peak)
    seed_err, ref_err = ctx.measure_rsqrt_error()
    emit(what="rsqrt_error", seed_rel=seed_err, refined_rel=ref_err)
    for n in sizes:
        w = synth.plummer(n, seed=1) if n < (1 << 20) else synth.galaxy_collision(n)
        dt = synth.default_dt(n)
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        steps = 3 if n >= (1 << 20) else (5 if n >= 262144 else 20)
        for kname, kern in (("tiled", capi.KERNEL_TILED), ("paired", capi.KERNEL_PAIRED)):
            if kern == capi.KERNEL_PAIRED and n <= 2048:
                continue
            ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
            ctx.step(1, dt, kernel=kern)  # warm-up (2 passes)
            t0 = time.time()
            ctx.step(steps, dt, kernel=kern)
            wall = time.time() - t0
            fms, passes = ctx.last_force_ms, ctx.last_force_passes
            inter = float(n) * (n - 1) * passes
            rate = inter / (fms * 1e-3)
            emit(what="force_pass", kernel=kname, n=n, steps=steps, passes=passes, force_ms_per_pass=fms / passes,
                 step_ms=ctx.last_step_ms, wall_s=wall, interactions_per_s=rate, frac_of_measured_peak=rate * 20 / (peak * 1e12),
                 frac_of_nominal=rate * 20 / 37.2e12)


if __name__ == "__main__":
    main()
