#!/usr/bin/env python3
"""Headline benchmark of the World::update hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): synthetic galaxy collision, N = 1,048,576 bodies, FP64, Leapfrog KDK,
dt = 200,000 s -- BASELINE.json's headline config; it fits one B200 (59 MB of state).  A "step" is
one World::update iteration over the whole world.  With N > 1 GPUs the targets are sharded over the
ranks (strong scaling: the world does not grow) and the drifted positions are all-gathered once per
force pass (NCCL, overlapped with the local-shard tiles).

One JSON line on stdout (rank 0):
  value     directed pairwise interactions / s, whole job, device-timed (CUDA events on the
            launching stream, max over ranks), state resident in HBM, L2 flushed between steps;
  roofline  FP64-pipe roofline of the force kernel: 20 flops per directed interaction
            (BASELINE.json's convention) / average launch duration measured live, against the
            FP64 FMA peak measured in the same run (MEASURED_PEAKS.json has no FP64 figure);
  e2e       the same metric through the public C-ABI call sequence a host program makes per tick
            with HOST buffers: essa_upload + essa_step + essa_download inside the timed region;
  cpu_baseline  the CPU oracle (a restatement of the reference's own loop on the reference's own
            containers) timed on this box's host cores on a bounded slice of the same world;
  extra     Solar System iterations / s (resident kernel through the World facade, History + Trail
            recording on) next to the oracle's, the second half of BASELINE.json's metric.

--impl reference times the reference's CPU implementation (oracle port: the reference itself
cannot be built offline) on the same world; each step is a bounded slice of World::set_forces.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

# NCCL writes its debug output ("NCCL version ..." at NCCL_DEBUG=VERSION and above) to stdout by default;
# stdout carries ONE JSON line, so NCCL's goes to stderr.
# (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, hence WARN.)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "NONE", ""):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_PER_INTERACTION = 20.0
NOMINAL_FP64_TFLOPS = 37.2  # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
METRIC = "pairwise_interactions_per_s"
UNIT = "interactions/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--bodies", dest="n", type=int, default=1 << 20, help="bodies (default: BASELINE.json's 1,048,576)")
    ap.add_argument("--no-extra", action="store_true", help="skip the Solar System / CPU baseline legs")
    ap.add_argument("--kernel", default="auto", choices=["auto", "tiled", "paired"])
    return ap.parse_args()


def workload(n):
    from essa_b200 import synth
    if n >= (1 << 20):
        w = synth.galaxy_collision(n)
        name = "synthetic galaxy collision (2 Plummer spheres) N=%d FP64 Leapfrog KDK" % n
    else:
        w = synth.plummer(n)
        name = "synthetic Plummer sphere N=%d FP64 Leapfrog KDK" % n
    return w, synth.default_dt(n), name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + ",".join(self.FIELDS),
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_mod():
    from essa_b200 import synth
    return synth


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's own loop
# --------------------------------------------------------------------------------------------
def cpu_world(w):
    """oracle::World holding the whole synthetic world in the reference's containers
    (std::list<std::unique_ptr<Object>>, one Trail + History per object)."""
    import oracle
    ow = oracle.World()
    ow.add_bodies(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"] / oracle.GRAVITY)
    return ow


def cpu_slice(ow, n, row0, target_pairs):
    """Time rows [row0, row0 + r) of World::set_forces's outer loop (pair-shared, list order).
    Returns (seconds, directed interactions evaluated, rows)."""
    rows = max(1, int(np.ceil(target_pairs / max(1, n - 1 - row0))))
    rows = min(rows, n - 1 - row0)
    secs, pairs = ow.time_force_rows(row0, row0 + rows)
    return secs, 2.0 * pairs, rows


def cpu_baseline(w, n, seconds=12.0):
    import oracle
    model, cores = host_info()
    t0 = time.time()
    try:
        ow = cpu_world(w)
    except MemoryError:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "could not allocate the oracle world"}
    build_s = time.time() - t0
    # calibrate on a short slice, then size the timed slice for ~`seconds`
    s0, i0, r0 = cpu_slice(ow, n, 0, 2e7)
    rate = i0 / s0
    s1, i1, r1 = cpu_slice(ow, n, r0, rate / 2.0 * seconds)
    # all-cores variant (NOT the reference, which is single-threaded): OpenMP over rows of the same law
    rows = np.arange(0, n, max(1, n // 256), dtype=np.int32)[:256]
    t = time.time()
    _, threads = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="omp")
    omp_s = time.time() - t
    return {"value": i1 / s1, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "rows [%d, %d) of World::set_forces over the same %d-body world: %.3g directed interactions in %.1f s, "
                      "1 thread (the reference is single-threaded), oracle built -O2 -ffp-contract=off" % (r0, r0 + r1, n, i1, s1),
            "host_cpu": model, "host_cores": cores, "oracle_world_build_s": round(build_s, 1),
            "omp_all_cores": {"value": len(rows) * (n - 1) / omp_s, "unit": UNIT, "threads": threads,
                              "note": "OpenMP over rows of the same pair law; not the reference"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    w, dt, name = workload(n)
    ow = cpu_world(w)
    row = 0
    # calibrate (untimed), then size a step so that the whole run stays near two minutes
    s, i, r = cpu_slice(ow, n, row, 1e7)
    row += r
    pairs_per_s = i / 2.0 / s
    per_step = float(min(5e7, max(2e6, 120.0 * pairs_per_s / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        s, i, r = cpu_slice(ow, n, row, per_step)
        row += r
    secs, inter, rows0 = 0.0, 0.0, row
    for _ in range(args.steps):
        s, i, r = cpu_slice(ow, n, row, per_step)
        secs += s
        inter += i
        row += r
    model, cores = host_info()
    value = inter / secs
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "n": n, "dt_seconds": dt,
                       "step": "bounded slice of one World::set_forces pass (rows [%d, %d) of the pair loop)" % (rows0, row)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "%d rows of the pair-shared loop per step, %d-body world, 1 thread (reference is single-threaded); "
                                       "the reference itself cannot be built offline, this is its restatement on its own containers" % ((row - rows0) // max(1, args.steps), n),
                             "host_cpu": model, "host_cores": cores},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# Solar System leg (second half of BASELINE.json's metric)
# --------------------------------------------------------------------------------------------
def solar_leg():
    """Small-world half of BASELINE.json's metric: Solar System iterations/s (single trajectory through
    the World facade, recording on; and the 65,536-copy ensemble aggregate), jupiter_system.essa as a
    resident cluster kernel and as a 65,536-copy ensemble -- each next to the CPU oracle."""
    import oracle
    from essa_b200 import capi, world
    wdir = os.path.join(ROOT, "tests", "golden", "worlds")
    out = {}
    for fname, tag, big, small in (("solar.essa", "solar", 1000000, 200000), ("jupiter_system.essa", "jupiter", 100000, 20000)):
        path = os.path.join(wdir, fname)
        for label, exact, steps in (("fast", False, big), ("exact_order", True, small)):
            pw = world.World.load(path)
            pw.simulation_seconds_per_tick = 600
            pw.set_flags(offset_trails=True, exact_order=exact)
            pw.update(10000 if tag == "solar" else 1000)
            pw.update(steps)
            out["%s_iters_per_s_%s" % (tag, label)] = steps / (pw.last_step_ms * 1e-3)
            # the GUI's cadence (SimulationView::update, src/SimulationView.cpp:374-397): 10 iterations per tick, the objects read
            # back after every tick.  Fast arithmetic: served by the persistent kernel (mailbox doorbell, no launch per tick).
            for _ in range(50):
                pw.update(10)
            s0 = pw.persist_stats()
            t0 = time.perf_counter()
            for _ in range(1000):
                pw.update(10)
            el = time.perf_counter() - t0
            s1 = pw.persist_stats()
            out["%s_iters_per_s_%s_10_per_call" % (tag, label)] = 10000 / el
            out["%s_us_per_update10_%s" % (tag, label)] = el / 1000 * 1e6
            if not exact:
                out["%s_persistent_ticks_of_1000" % tag] = s1["ticks"] - s0["ticks"]
                pw.set_persistent(False)
                pw.update(10)
                t0 = time.perf_counter()
                for _ in range(200):
                    pw.update(10)
                out["%s_us_per_update10_fast_one_launch_per_call" % tag] = (time.perf_counter() - t0) / 200 * 1e6
            st = pw.state()
            pw.close()
        ow = oracle.World.load(path)
        ow.dt = 600
        ow.update(1000)
        nsteps = 300000 if tag == "solar" else 5000
        out["%s_iters_per_s_cpu_oracle" % tag] = nsteps / ow.time_update(nsteps)
        out["%s_speedup_vs_cpu_oracle" % tag] = out["%s_iters_per_s_fast" % tag] / out["%s_iters_per_s_cpu_oracle" % tag]
        out["%s_speedup_vs_cpu_oracle_10_per_call" % tag] = out["%s_iters_per_s_fast_10_per_call" % tag] / out["%s_iters_per_s_cpu_oracle" % tag]
        # 65,536 perturbed copies (forward-simulation semantics: no History)
        c = capi.Context(0)
        n = len(st["gf"])
        c.upload(st["pos"][:, 0], st["pos"][:, 1], st["pos"][:, 2], st["vel"][:, 0], st["vel"][:, 1], st["vel"][:, 2], st["gf"])
        copies, esteps = 65536, (200 if tag == "solar" else 20)
        c.ensemble_init(copies, 600, seed=7, rel_sigma=1e-6)
        c.ensemble_step(2)
        c.ensemble_step(esteps)
        ms = c.last_step_ms
        out["%s_ensemble_copy_iters_per_s" % tag] = copies * esteps / (ms * 1e-3)
        out["%s_ensemble_interactions_per_s" % tag] = copies * (esteps + 1) * n * (n - 1) / (ms * 1e-3)
        out["%s_ensemble_aggregate_vs_cpu_oracle" % tag] = out["%s_ensemble_copy_iters_per_s" % tag] / out["%s_iters_per_s_cpu_oracle" % tag]
        if tag == "solar":
            # 2,048 warps of one-copy-per-thread work leave a sixth of the 592 schedulers a warp short; 1,048,576 copies fill them
            big = 1 << 20
            c.ensemble_init(big, 600, seed=7, rel_sigma=1e-6)
            c.ensemble_step(2)
            c.ensemble_step(esteps)
            out["solar_ensemble_1m_copies_copy_iters_per_s"] = big * esteps / (c.last_step_ms * 1e-3)
        c.close()
    # mid-size worlds (no BASELINE configuration, but the sizes between the resident kernels and the pair-shared one): us per step of
    # a 100-step call, one cooperative launch (K1g) vs one launch per step (K1), unequal masses with most-attracting links
    from essa_b200 import synth
    c = capi.Context(0)
    for m in (1024, 2048, 4096):
        w = synth.plummer(m, seed=m)
        gf = w["gf"] * (1.0 + 1e-3 * (np.arange(m) % 7))
        for label, kern in (("grid", capi.KERNEL_AUTO), ("tiled", capi.KERNEL_TILED)):
            c.date = capi.START_DATE
            c.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
            c.step(5, 1000, kernel=kern, track_most_attracting=True)
            c.step(100, 1000, kernel=kern, track_most_attracting=True)
            out["mid_n%d_us_per_step_%s" % (m, label)] = c.last_step_ms * 1e3 / 100
    c.close()
    out["note"] = ("dt=600 s, History + Trail + ap/pe recording on, offset trails, through the World facade (update(k) + read-back of every "
                   "object per call); *_iters_per_s_<mode> = one call of many steps, *_10_per_call = the GUI's cadence (10 iterations per "
                   "tick), where fast arithmetic is served by the persistent cluster kernel through its mailbox; a single trajectory is "
                   "latency-bound (a step is a dependent chain), so >=100x the CPU is reachable only in aggregate -- the ensemble lines; "
                   "CPU = oracle (reference containers, -O2, 1 thread) on this box; the README's 67,500 it/s (Ryzen 5 5600H, GUI on) is "
                   "published-not-measured")
    return out


def parity_vs_oracle(ctx, w, n, opts, rank, rows=32, gf=None):
    """One force pass on the (possibly sharded) context; rank 0 compares sampled rows with the CPU oracle's extended-precision
    evaluation over ALL sources -- the bar of tests/test_gpu_tiled.py: the GPU may differ from exact arithmetic at most 4x as
    much as the reference's own binary64 loop does (and never more than 1e-13 when that loop is better).  Collective."""
    import oracle
    ctx.set_forces(opts=opts)
    got = ctx.download(("ax", "ay", "az"))
    if rank != 0:
        return None
    g = w["gf"] if gf is None else gf
    idx = np.sort(np.random.default_rng(11).choice(n, rows, replace=False)).astype(np.int32)
    ld = oracle.forces_rows(w["x"], w["y"], w["z"], g, idx, mode="ld")
    ref, _ = oracle.forces_rows(w["x"], w["y"], w["z"], g, idx, mode="omp")
    acc = np.stack([got["ax"], got["ay"], got["az"]], axis=1)[idx]
    err = float((np.linalg.norm(acc - ld, axis=1) / np.linalg.norm(ld, axis=1)).max())
    err_ref = float((np.linalg.norm(ref - ld, axis=1) / np.linalg.norm(ld, axis=1)).max())
    tol = max(1e-13, 4.0 * err_ref)
    return {"parity_rel_err": err, "oracle_own_rel_err": err_ref, "tolerance": tol, "rows": int(rows), "ok": bool(err <= tol),
            "what": "max over sampled target rows of |a_gpu - a_ext| / |a_ext|, a_ext = the oracle's long-double sum over all sources"}


def side_config(capi, synth, local_rank, rank, world_size, uid_fn, barrier, max_over_ranks, n, steps, warmup, peak, unequal=False):
    """Another BASELINE.json configuration on the same ranks: Plummer sphere of n bodies (optionally with masses spread over a
    factor 4 and most-attracting tracking with real candidates), timed like the headline."""
    w = synth.plummer(n) if n < (1 << 20) else synth.galaxy_collision(n)
    gf = w["gf"]
    if unequal:
        gf = gf * (1.0 + 3.0 * np.random.default_rng(3).random(n))
    dt = synth.default_dt(n)
    ctx = capi.Context(local_rank, n)
    if world_size > 1:
        ctx.shard_attach(rank, world_size, uid_fn())
    opts = capi.Context.opts(dt, kernel=capi.KERNEL_AUTO, track_most_attracting=True)
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
    par = parity_vs_oracle(ctx, w, n, opts, rank, rows=16 if n > 300000 else 32, gf=gf)
    for _ in range(warmup):
        ctx.step(1, opts=opts)
    barrier()
    p0 = ctx.passes
    ms = fms = 0.0
    fp = 0
    for _ in range(steps):
        ctx.flush_l2()
        ctx.step(1, opts=opts)
        ms += ctx.last_step_ms
        fms += ctx.last_force_ms
        fp += ctx.last_force_passes
    phases = ctx.last_pass_phases
    barrier()
    ms = max_over_ranks(ms)
    fms = max_over_ranks(fms)
    passes = ctx.passes - p0
    ctx.close()
    if rank != 0:
        return None
    value = float(passes) * n * (n - 1) / (ms * 1e-3)
    paired = n >= 16384 and n % (2048 * world_size) == 0
    uniform = paired and not unequal
    ipi = (9.0 if uniform else 10.0) if paired else 16.0
    tfl = (n // world_size) * float(n - 1) * FLOPS_PER_INTERACTION / (fms / max(1, fp) * 1e-3) / 1e12
    return {"n": n, "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "n_gpus": world_size,
            "masses": ("spread over a factor 4, most-attracting links tracked with candidate keys on the %s" % (
                "mass-sorted layout" if world_size == 1 else "caller's order (sharded ranks own ranges of original indices)")) if unequal else "equal",
            "last_pass_phases_rank0": phases,
            "roofline_frac_20flop": tfl / peak if peak else None, "fp64_instructions_per_interaction": ipi,
            "fp64_pipe_frac": tfl / peak * (2.0 * ipi / FLOPS_PER_INTERACTION) if peak else None, "parity": par}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from essa_b200 import capi

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the essa_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from essa_b200 import dist as edist  # the rendezvous-level plumbing (tests/test_dist_cpu.py exercises it under gloo)

    def max_over_ranks(x):
        return x if world_size == 1 else edist.max_over_ranks(x)

    fresh_uid = edist.broadcast_unique_id

    n = args.n
    w, dt, name = workload(n)
    ctx = capi.Context(local_rank, n)
    if world_size > 1:
        ctx.shard_attach(rank, world_size, fresh_uid())
    kernel = {"auto": capi.KERNEL_AUTO, "tiled": capi.KERNEL_TILED, "paired": capi.KERNEL_PAIRED}[args.kernel]
    paired = args.kernel == "paired" or (args.kernel == "auto" and n >= 16384 and n % (2048 * world_size) == 0)
    # reference semantics: Object::update_forces_against always runs its most-attracting bookkeeping
    opts = capi.Context.opts(dt, kernel=kernel, track_most_attracting=True)

    peak = ctx.measure_fp64_peak(300.0) if rank == 0 else 0.0
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    # parity first: the very configuration that is timed below, on this rank count, against the CPU oracle
    parity = parity_vs_oracle(ctx, w, n, opts, rank, rows=16)
    if rank == 0 and not parity["ok"]:
        raise SystemExit("bench.py: parity failure, %r" % (parity,))
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    e0 = ctx.energy() if n <= (1 << 20) else None

    for _ in range(max(args.warmup, 0)):
        ctx.step(1, opts=opts)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0, passes0 = ctx.launches, ctx.passes
    t_wall0 = time.time()
    dev_ms = force_ms = 0.0
    force_passes = 0
    for _ in range(args.steps):
        ctx.flush_l2()  # outside the event-timed region
        ctx.step(1, opts=opts)
        dev_ms += ctx.last_step_ms
        force_ms += ctx.last_force_ms
        force_passes += ctx.last_force_passes
        breakdown = ctx.last_step_breakdown
        breakdown["pass_phases"] = ctx.last_pass_phases  # where the last force pass spent its time (sharded: the timeline of one rank)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launches - launches0 - args.steps  # the L2 flush is a memset, not a kernel of ours... count kernels only
    launches = ctx.launches - launches0
    passes = ctx.passes - passes0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    force_ms_max = max_over_ranks(force_ms)
    total_interactions = float(passes) * n * (n - 1)  # all ranks together evaluate N(N-1) per pass
    value = total_interactions / (dev_ms * 1e-3)

    # ---- energy drift over everything stepped so far (warm-up + timed) ----
    drift = None
    if e0 is not None:
        e1 = ctx.energy()
        E0, E1 = e0[0] + e0[1], e1[0] + e1[1]
        drift = {"steps": args.warmup + args.steps, "E0_J": E0, "E1_J": E1, "rel": abs(E1 - E0) / abs(E0),
                 "momentum_rel": float(np.abs(e1[2] - e0[2]).max() / (np.abs(w["vx"]).sum() * (w["gf"][0] / capi.GRAVITY)))}

    # ---- end to end through the C ABI with host buffers ----
    e2e_steps = max(1, args.steps)
    host = {k: np.array(w[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "gf")}
    # one untimed tick first: the first download allocates its pinned staging buffer (and, sharded, sets up the gather
    # of velocities / accelerations) -- 70 ms at 1 GPU, 250 ms at 8 that belong to no steady-state tick
    ctx.upload(host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"], host["gf"])
    ctx.step(1, opts=opts)
    ctx.download(("x", "y", "z", "vx", "vy", "vz"))
    barrier()
    p0 = ctx.passes
    t0 = time.perf_counter()
    ticks = []
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        if world_size > 1:  # a sharded host program keeps and moves only the bodies its rank integrates
            ctx.upload_shard(host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"])
        else:
            ctx.upload(host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"], host["gf"])
        tb = time.perf_counter()
        ctx.step(1, opts=opts)
        tc = time.perf_counter()
        if world_size > 1:
            ctx.download_shard(("x", "y", "z", "vx", "vy", "vz"), host)
        else:
            ctx.download(("x", "y", "z", "vx", "vy", "vz"), out=host)  # into the host program's own arrays
        td = time.perf_counter()
        ticks.append({"upload_ms": 1e3 * (tb - ta), "step_ms": 1e3 * (tc - tb), "download_ms": 1e3 * (td - tc),
                      "device_step_ms": ctx.last_step_ms, "force_ms": ctx.last_force_ms, "force_passes": ctx.last_force_passes})
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = float(ctx.passes - p0) * n * (n - 1) / e2e_s
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": int(n * 7 * 8) if world_size == 1 else int(n // world_size * 6 * 8) * world_size,
           "d2h_bytes_per_step": int(n * 6 * 8),
           "steps": e2e_steps, "passes_per_step": (ctx.passes - p0) / e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "ticks_rank0": ticks,
           "note": "after one untimed tick: essa_upload (host SoA -> pinned staging -> HBM) + essa_step(1) + essa_download per step "
                   "(N > 1: essa_upload_shard / essa_download_shard -- every rank moves the bodies it integrates over its own PCIe link, "
                   "the other ranks' records arrive over NVLink; bytes are whole-job totals); an upload invalidates the resident "
                   "acceleration, so each step evaluates both set_forces() passes, as the reference does"}

    ctx.close()
    # ---- the other large-N configurations of BASELINE.json, on the same ranks (collective: every rank takes part) ----
    sides = {}
    if not args.no_extra and n == (1 << 20):
        for tag, sn, sk, unequal in (("n16384", 16384, 20, False), ("n262144", 262144, 3, False),
                                     ("unequal_mass_tracked_n1m", 1 << 20, 2, True)):
            try:
                sides[tag] = side_config(capi, synth_mod(), local_rank, rank, world_size, fresh_uid, barrier, max_over_ranks, sn, sk, 2, peak,
                                         unequal)
            except Exception as e:  # the headline must not die on a side leg
                sides[tag] = {"error": repr(e)}
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    ni = n // world_size
    per_launch_interactions = float(ni) * (n - 1)
    avg_force_ms = force_ms_max / max(1, force_passes)
    achieved_tflops = per_launch_interactions * FLOPS_PER_INTERACTION / (avg_force_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("%s_n%d" % ("paired" if paired else "tiled", n))
        except Exception:
            traffic = None
    uniform = paired and bool(np.all(w["gf"] == w["gf"][0])) and os.environ.get("ESSA_PAIR_UNIFORM", "1") != "0"
    ipi = (9.0 if uniform else 10.0) if paired else 16.0
    # floating-point operations the kernel really executes per directed interaction (an FMA counts 2): per pair 12 DFMA +
    # 3 DADD + 3 DMUL (+ 2 DMUL by the gravity factors unless uniform); one-sided kernel 9 DFMA + 3 DADD + 4 DMUL
    fpi = (15.0 if uniform else 16.0) if paired else 25.0
    roofline = {"bound": "fp64", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tflops / peak,
                "traffic": traffic, "kernel": "k_force_paired (+ k_pair_finish)" if paired else "k_force_tiled",
                "avg_launch_ms": avg_force_ms, "launches_timed": force_passes,
                "algorithmic_flops_per_launch": per_launch_interactions * FLOPS_PER_INTERACTION,
                "peak_source": "measured in this run: register-resident independent DFMA chains on all SMs for 300 ms "
                               "(MEASURED_PEAKS.json carries no FP64 figure)",
                "frac_of_nominal": achieved_tflops / NOMINAL_FP64_TFLOPS, "nominal_peak": NOMINAL_FP64_TFLOPS,
                "fp64_instructions_per_interaction": ipi,
                "fp64_pipe_frac": achieved_tflops / peak * (2.0 * ipi / FLOPS_PER_INTERACTION),
                "executed_flops_per_interaction": fpi,
                "executed_tflops": achieved_tflops * fpi / FLOPS_PER_INTERACTION,
                "note": "the path is FP64-pipe bound (intensity ~0.19*N flop/B), neither HBM nor tensor.  `frac` follows BASELINE.json's "
                        "convention of 20 flops per directed interaction; `fp64_pipe_frac` is the share of the measured DFMA issue rate the "
                        "kernel's FP64 instructions actually occupy (interactions/s x instructions per interaction / peak instruction rate); `frac` can "
                        "exceed 1 because the kernel executes fewer operations (`executed_flops_per_interaction`) than the convention counts.  " + (
                    ("Pair-shared kernel: each unordered pair is evaluated once and applied to both bodies (as the reference's own "
                     "i<j loop does).  %s") % (
                        "Every body of this world is live and has the same gravity factor g, so the kernel accumulates sum |d|^-3 d and "
                        "multiplies by g once per body: 18 FP64 instructions per pair = 9 per directed interaction "
                        "(ESSA_PAIR_UNIFORM=0 runs the general kernel: 20 per pair)" if uniform else
                        "20 FP64 instructions per pair = 10 per directed interaction")
                    if paired else
                    "16 FP64 instructions per directed interaction cap this formulation at 10/16 = 62.5% of the 20-flop convention")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": name, "n": n, "dt_seconds": dt, "sharding": "targets/%d: NCCL gather of drifted positions + reduce-scatter of pair-shared partial accelerations per pass" % world_size
                       if world_size > 1 else "single GPU", "l2": "flushed (256 MiB memset) between timed steps",
                       "passes_per_step": passes / args.steps,
                       "most_attracting": "tracked, as Object::update_forces_against always does (track_most_attracting=1); every body of this workload has the same mass, so the heavier-partner test of src/Object.cpp:86,91 fails for every pair: the pair-shared kernel writes max_attraction = 0 and keeps the links, exactly what the reference leaves (checked against the tiled argmax kernel in tests/test_gpu_paired.py); unequal masses carry sortable candidate keys through the same kernel", "accounting": "value counts only force passes actually evaluated: N(N-1) directed interactions per pass; the reference's "
                                     "first set_forces() of a step recomputes the previous step's second one bit for bit, so the resident "
                                     "acceleration is reused (tests: two_pass == reuse, bitwise) and is NOT counted",
                       "timing": "CUDA events on the launching stream around each essa_step, max over ranks"},
            "roofline": roofline, "e2e": e2e, "parity": parity, "gpu_launches": int(launches), "clocks": clocks, "energy_drift": drift,
            "wall_ms_per_step": 1e3 * (t_wall1 - t_wall0) / args.steps, "last_step_breakdown_rank0": breakdown}
    line["extra"] = dict(sides)
    if world_size == 1 and not args.no_extra:
        line["cpu_baseline"] = cpu_baseline(w, n)
        try:
            line["extra"].update(solar_leg())
        except Exception as e:  # the headline must not die on the secondary leg
            line["extra"]["solar_error"] = repr(e)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
