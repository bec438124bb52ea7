#!/usr/bin/env python3
"""Multi-GPU parity check of the sharded paths (tiled and pair-shared); run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/sharded_check.py

Every rank steps its shard of a ragged world (N not divisible by the rank count); rank 0 also steps
the same world unsharded and compares (different partial-sum grouping -> tolerance, not bits), and checks a
sharded force pass against the CPU oracle's extended-precision rows (oracle.forces_rows) with the bar of
tests/test_gpu_tiled.py.  Then most-attracting links on unequal masses over the WHOLE gathered world (essa_download
gathers links / max_attraction / ap / pe from their owners), and the near-tie worlds of tests/test_gpu_paired.py with
target and partners on different ranks against the oracle's World::set_forces."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (the checker)
from essa_b200 import capi, synth  # noqa: E402
from essa_b200 import dist as edist  # noqa: E402


def fresh_uid(rank):
    return edist.broadcast_unique_id()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.from_numpy(edist.broadcast_unique_id())
    ok = True
    for n, steps, kern in ((4099, 6, capi.KERNEL_TILED), (70001, 3, capi.KERNEL_TILED), (8192 * world, 4, capi.KERNEL_PAIRED),
                           (65536, 3, capi.KERNEL_AUTO)):
        w = synth.plummer(n, seed=21)
        dt = synth.default_dt(n)
        ctx = capi.Context(local)
        ctx.shard_attach(rank, world, uid.cpu().numpy())
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        e0 = ctx.energy()
        # one sharded force pass against the oracle's extended-precision rows (the bar of tests/test_gpu_tiled.py)
        ctx.set_forces(kernel=kern)
        f = ctx.download(("ax", "ay", "az"))
        if rank == 0:
            rows = np.sort(np.random.default_rng(3).choice(n, 96, replace=False))
            ld = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ld")
            refr = oracle.forces_rows(w["x"], w["y"], w["z"], w["gf"], rows, mode="ref")
            acc = np.stack([f["ax"], f["ay"], f["az"]], axis=1)[rows]
            err = (np.linalg.norm(acc - ld, axis=1) / np.linalg.norm(ld, axis=1)).max()
            err_ref = (np.linalg.norm(refr - ld, axis=1) / np.linalg.norm(ld, axis=1)).max()
            good = err <= max(1e-13, 4.0 * err_ref)
            ok = ok and good
            print("sharded_check kernel=%d n=%d ranks=%d force pass vs oracle rows: %.2e (oracle's own %.2e) -> %s"
                  % (kern, n, world, err, err_ref, "PASS" if good else "FAIL"), flush=True)
        ctx.step(steps, dt, kernel=kern)
        ctx.step(-2, dt, kernel=kern)
        got = ctx.download()
        e1 = ctx.energy()
        if rank == 0:
            ref = capi.Context(local)
            ref.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
            r0 = ref.energy()
            ref.step(steps, dt, kernel=capi.KERNEL_TILED)
            ref.step(-2, dt, kernel=capi.KERNEL_TILED)
            exp = ref.download()
            r1 = ref.energy()
            worst = 0.0
            for grp in (("x", "y", "z"), ("vx", "vy", "vz"), ("ax", "ay", "az")):
                a = np.stack([got[k] for k in grp], axis=1)
                b = np.stack([exp[k] for k in grp], axis=1)
                worst = max(worst, (np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max())
            e_ok = abs(e0[1] - r0[1]) <= 1e-12 * abs(r0[1]) and abs(e1[0] - r1[0]) <= 1e-10 * abs(r1[0])
            good = worst <= 1e-10 and e_ok and ctx.date == ref.date
            ok = ok and good
            print("sharded_check kernel=%d n=%d ranks=%d worst_rel=%.3e energy_ok=%s -> %s" % (kern, n, world, worst, e_ok, "PASS" if good else "FAIL"), flush=True)
            ref.close()
        ctx.close()
        # a fresh communicator for the next world
        uid = torch.from_numpy(edist.broadcast_unique_id())
    # pair-shared kernel with most-attracting tracking on unequal masses: candidate keys are max-reduced over the ranks
    n = 8192 * world
    w = synth.plummer(n, seed=23)
    gf = w["gf"] * 10.0 ** np.random.default_rng(7).uniform(-2, 2, n)
    dt = synth.default_dt(n)
    ctx = capi.Context(local)
    ctx.shard_attach(rank, world, uid.cpu().numpy())
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
    ctx.step(3, dt, kernel=capi.KERNEL_PAIRED, track_most_attracting=True)
    got = ctx.download()
    if rank == 0:
        ref = capi.Context(local)
        ref.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], gf)
        ref.step(3, dt, kernel=capi.KERNEL_TILED, track_most_attracting=True)
        exp = ref.download()
        # essa_download gathers links and max_attraction from their owners: the WHOLE world is compared
        same = np.array_equal(got["most"], exp["most"])
        rel = np.abs(got["max_attraction"] - exp["max_attraction"]).max() / np.abs(exp["max_attraction"]).max()
        good = same and rel <= 1e-11 and (got["most"] >= 0).sum() >= n - 1
        ok = ok and good
        print("sharded_check pair-shared tracking n=%d ranks=%d links_equal=%s maxattr_rel=%.2e -> %s" % (n, world, same, rel, "PASS" if good else "FAIL"), flush=True)
        ref.close()
    ctx.close()
    # a sharded host program: every rank uploads / downloads only the bodies it integrates (essa_upload_shard / _download_shard)
    n = 8192 * world
    w = synth.plummer(n, seed=29)
    dt = synth.default_dt(n)
    ctx = capi.Context(local)
    ctx.shard_attach(rank, world, fresh_uid(rank))
    ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
    ctx.step(2, dt)
    host = ctx.download(("x", "y", "z", "vx", "vy", "vz"))
    for k in ("vx", "vy", "vz"):
        host[k] = host[k] * (1.0 + 1e-6)  # the host program changes its bodies between ticks (every rank: the same rule)
    i0, i1 = capi.shard_range(n, rank, world)
    mine = {k: np.where((np.arange(n) >= i0) & (np.arange(n) < i1), host[k], np.nan) for k in host}  # other ranks' bodies: poison
    ctx.upload_shard(mine["x"], mine["y"], mine["z"], mine["vx"], mine["vy"], mine["vz"])
    ctx.step(2, dt)
    out = {k: np.full(n, np.nan) for k in ("x", "y", "z", "vx", "vy", "vz")}
    ctx.download_shard(("x", "y", "z", "vx", "vy", "vz"), out)
    if rank == 0:
        ref = capi.Context(local)
        ref.upload(host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"], w["gf"])
        ref.step(2, dt)
        exp = ref.download(("x", "y", "z", "vx", "vy", "vz"))
        sl = slice(i0, i1)
        worst = max(float(np.abs(out[k][sl] - exp[k][sl]).max() / np.abs(exp[k][sl]).max()) for k in out)
        untouched = all(np.isnan(out[k][i1:]).all() for k in out)
        good = worst <= 1e-12 and untouched
        ok = ok and good
        print("sharded_check shard upload / download n=%d ranks=%d worst_rel=%.2e other_slices_untouched=%s -> %s"
              % (n, world, worst, untouched, "PASS" if good else "FAIL"), flush=True)
        ref.close()
    ctx.close()
    # near ties / exact ties between two heavier partners (src/Object.cpp:84-94 decides in full binary64, strict '>', ascending
    # index): target and partners on different ranks, so the candidates meet in the exchange of per-rank keys
    n = 4096 * world
    blk = n // world
    for delta in (1e-7, 1e-9, 1e-12, 0.0):
        for swap in (False, True):
            t, a, b = 100, (world - 1) * blk + 77, blk // 2 + 1500
            if swap:
                a, b = b, a
            wq = synth.plummer(n, seed=5)
            x, y, z = wq["x"] + 1e13, wq["y"].copy(), wq["z"].copy()
            mass = np.full(n, 1e20)
            x[t], y[t], z[t] = 0.0, 0.0, 0.0
            x[a], y[a], z[a] = 1.0e9, 0.0, 0.0
            x[b], y[b], z[b] = 0.0, 1.0e9 * (1.0 + delta / 4.0), 0.0
            mass[a] = mass[b] = 1e26
            gfq = mass * capi.GRAVITY
            ctx = capi.Context(local)
            ctx.shard_attach(rank, world, fresh_uid(rank))
            zero = np.zeros(n)
            ctx.upload(x, y, z, zero, zero, zero, gfq, most=np.full(n, -1, dtype=np.int32))
            ctx.set_forces(kernel=capi.KERNEL_PAIRED, track_most_attracting=True)
            got = ctx.download(("most",))
            if rank == 0:
                ow = oracle.World()
                ow.add_bodies(x, y, z, zero, zero, zero, mass)
                ow.set_forces()
                ost = ow.state()
                good = bool(np.array_equal(ost["gf"], gfq)) and np.array_equal(got["most"], ost["most"]) \
                    and int(ost["most"][t]) == (a if delta > 0 else min(a, b))
                ok = ok and good
                print("sharded_check near tie delta=%g swap=%d ranks=%d link=%d oracle=%d -> %s"
                      % (delta, swap, world, got["most"][t], ost["most"][t], "PASS" if good else "FAIL"), flush=True)
            ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
