"""Data-parallel training of the fused VAE: one process per GPU, batch sharded by rank, ONE SUM all-reduce per step
over the flat bucket [network grads | radius grads | ELBO statistics] (NCCL over NVLink / NVSwitch).

The reference has no distributed mode; its ELBO is a SUM over the batch (mt/mvae/stats.py:200-202),
so the cross-GPU reduction is a SUM, not DDP's mean: N ranks with B/N samples each reproduce the single-GPU step at
batch B up to summation order.  Parameters and optimizer state are replicated; every rank applies the same update.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(batch: int, rank: int, world: int) -> tuple:
    """Contiguous batch split: rank r owns rows [r*B/N, (r+1)*B/N) — remainder to the first ranks."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(bucket: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of the data path."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket


def attach(model, group=None) -> None:
    """Make `model.train_step` all-reduce its gradient/statistics bucket before the optimizer step."""
    model._grad_hook = lambda bucket: allreduce_sum_(bucket, group)


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model._flat, src=src, group=group)
        dist.broadcast(model._rflat, src=src, group=group)
        model.mark_parameters_changed()
