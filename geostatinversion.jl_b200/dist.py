"""Multi-GPU bootstrap: one process per GPU (torchrun / Julia `Distributed` workers).
torch.distributed is only the plumbing that carries the 128-byte NCCL id from rank 0 to
the other ranks; the data path (all-gathers of iterates and TSQR R factors, pivot
exchanges) is NCCL called from inside libgsi_b200.so."""
import os

from .core import Context, partition_rows, set_default_context


def broadcast_unique_id(make_id=None):
    """Rank 0 creates the id (Context.unique_id by default), everyone returns it."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = (make_id or Context.unique_id)()
        assert len(raw) == 128
        t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def context_from_torch_distributed(set_default=True):
    """Creates this rank's Context from the torchrun environment."""
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = broadcast_unique_id() if world > 1 else None
    ctx = Context(local_rank, rank, world, uid)
    if set_default:
        set_default_context(ctx)
    return ctx


def my_rows(n, ctx):
    return partition_rows(n, ctx.world, ctx.rank)
