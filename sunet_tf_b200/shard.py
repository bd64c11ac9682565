"""Work partitioning over GPUs: one process per GPU, contiguous shards, NO collective in the forward.

Images of a batch never interact (LayerNorm is per token, attention per window, there is no BatchNorm), and the tiles
of an any-resolution input are independent forwards, so every rank runs the full model on its own shard.  The only
exchange in the whole path is the single sum-reduce of the folded canvases in tiles.denoise_any_resolution.
"""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def batch_range(total, rank, world_size):
    """Contiguous slice [lo, hi) of `total` items for `rank`; sizes differ by at most one (BASELINE config 3: 512/G)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_range(n_tiles, rank, world_size):
    """Row-major tile t goes to rank t * world_size // n_tiles -> contiguous [lo, hi)."""
    lo = -((-rank * n_tiles) // world_size)
    hi = -((-(rank + 1) * n_tiles) // world_size)
    return lo, min(hi, n_tiles)


def init_process_group(backend=None):
    """torch.distributed over 127.0.0.1 (NCCL on GPUs, gloo on CPU); no-op for a single process."""
    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws, local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=ws, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return rank, ws, local


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (timing is always reported as the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
