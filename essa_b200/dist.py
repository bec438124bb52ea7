"""torch.distributed plumbing for the sharded tiled path: one process per GPU.

Only rendezvous-level work happens here (broadcast of the NCCL id that essa_shard_attach needs,
max-over-ranks of a timing); the data path collective -- the per-pass gather of drifted positions --
is issued by the C ABI on its own NCCL communicator."""
import numpy as np
import torch
import torch.distributed as dist

from . import capi


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_unique_id(src=0):
    """Rank `src` asks NCCL for a unique id; every rank returns the same 128 bytes."""
    uid = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        uid.copy_(torch.from_numpy(capi.nccl_unique_id()))
    dist.broadcast(uid, src)
    return uid.cpu().numpy()


def max_over_ranks(x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ranges(n):
    """Every rank's target range, as a list indexed by rank (sanity check of the partition)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.tensor(capi.shard_range(n, rank, world), dtype=torch.int64, device=_device())
    out = [torch.zeros(2, dtype=torch.int64, device=_device()) for _ in range(world)]
    dist.all_gather(out, mine)
    return [tuple(int(v) for v in t.cpu()) for t in out]


def attach(ctx):
    """Join this process's context to the job-wide communicator (no-op for a single process)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    ctx.shard_attach(dist.get_rank(), dist.get_world_size(), broadcast_unique_id())
