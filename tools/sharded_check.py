#!/usr/bin/env python3
"""Multi-GPU parity check of the sharded tiled path; run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/sharded_check.py

Every rank steps its shard of a ragged world (N not divisible by the rank count); rank 0 also steps
the same world unsharded and compares (different partial-sum grouping -> tolerance, not bits)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from essa_b200 import capi, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.from_numpy(capi.nccl_unique_id()))
    dist.broadcast(uid, 0)
    ok = True
    for n, steps, kern in ((4099, 6, capi.KERNEL_TILED), (70001, 3, capi.KERNEL_TILED), (8192 * world, 4, capi.KERNEL_PAIRED),
                           (65536, 3, capi.KERNEL_AUTO)):
        w = synth.plummer(n, seed=21)
        dt = synth.default_dt(n)
        ctx = capi.Context(local)
        ctx.shard_attach(rank, world, uid.cpu().numpy())
        ctx.upload(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["gf"])
        e0 = ctx.energy()
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
        uid = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(capi.nccl_unique_id()))
        dist.broadcast(uid, 0)
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
        own = slice(0, n // world)  # most / max_attraction are kept for a rank's own targets
        same = np.array_equal(got["most"][own], exp["most"][own])
        rel = np.abs(got["max_attraction"][own] - exp["max_attraction"][own]).max() / np.abs(exp["max_attraction"][own]).max()
        good = same and rel <= 1e-11 and (got["most"][own] >= 0).sum() >= n // world - 1
        ok = ok and good
        print("sharded_check pair-shared tracking n=%d ranks=%d links_equal=%s maxattr_rel=%.2e -> %s" % (n, world, same, rel, "PASS" if good else "FAIL"), flush=True)
        ref.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
