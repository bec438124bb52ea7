#!/usr/bin/env python3
"""essa_measure_fp64_peak alone (the roofline denominator of bench.py): run it under
    ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,... -k regex:k_dfma_peak
to see whether the measured DFMA rate (~92 % of 148 x 64 x 2 x 1.965 GHz) is the pipe's ceiling or the micro-benchmark's."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi  # noqa: E402

ms = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
c = capi.Context(0)
for _ in range(3):
    print("fp64 peak over %.0f ms: %.2f TFLOP/s" % (ms, c.measure_fp64_peak(ms)), flush=True)
c.close()
