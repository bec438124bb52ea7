"""N > 1 host path on CPU: world_size-2 (and 3) gloo process groups exercise the rendezvous plumbing
the sharded bench uses -- NCCL id broadcast, shard partition, max-over-ranks -- and check that a rank
without a GPU fails loudly instead of computing on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    from essa_b200 import capi
    from essa_b200 import dist as edist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = edist.broadcast_unique_id()
        ranges = edist.gather_ranges(n)
        slowest = edist.max_over_ranks(10.0 + rank)
        refused = False
        if not torch.cuda.is_available():
            try:
                ctx = capi.Context(0)
                edist.attach(ctx)
            except capi.EssaError as e:
                refused = e.code == -2
        q.put((rank, bytes(uid), ranges, slowest, refused))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1048576), (2, 4099), (3, 10)])
def test_gloo_rendezvous_and_partition(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    uids = {g[1] for g in got}
    assert len(uids) == 1 and len(next(iter(uids))) == 128 and any(next(iter(uids)))
    ranges = got[0][2]
    assert all(g[2] == ranges for g in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    assert all(g[3] == 10.0 + world - 1 for g in got)
    import torch
    if not torch.cuda.is_available():
        assert all(g[4] for g in got)


def test_shard_range_is_a_partition():
    from essa_b200 import capi
    for n in (0, 1, 7, 1000, 1048576):
        for world in (1, 2, 3, 8):
            edges = [capi.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
    with pytest.raises(capi.EssaError):
        capi.shard_range(10, 2, 2)
