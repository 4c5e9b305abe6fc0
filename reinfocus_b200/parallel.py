"""Sharding of vector envs across the GPUs of one box.

Envs are independent, so GPU g simply owns the contiguous env block
[g*n/G, (g+1)*n/G) with its own renderer (its own seed-0 RNG-state cache): a sharded run
equals G independent reference VectorEnvironment(num_envs=n/G) instances. The only
communication is gathering the per-env observations (a few bytes per env) to every rank /
the policy rank, done with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

import os


def shard_bounds(num_envs: int, world_size: int, rank: int) -> tuple[int, int]:
    """[first, last) env indices owned by ``rank``; the first ``num_envs % world_size``
    ranks take one extra env."""

    assert 0 <= rank < world_size and num_envs >= 0
    base, extra = divmod(num_envs, world_size)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None):
    """Initialises torch.distributed from torchrun's environment (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_ADDR, MASTER_PORT). Returns (rank, world_size, local_rank)."""

    import torch
    import torch.distributed as dist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return rank, world_size, local_rank


def gather_observations(local, num_envs: int, group=None):
    """All-gathers per-env rows (local shard ``[n_local, ...]``) into ``[num_envs, ...]`` on
    every rank, in env order. Shards may differ in length by one (see shard_bounds)."""

    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world_size = dist.get_world_size(group)
    sizes = [shard_bounds(num_envs, world_size, r) for r in range(world_size)]
    longest = max(last - first for first, last in sizes)
    padded = local
    if local.shape[0] < longest:
        pad = torch.zeros((longest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world_size * longest,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(last - first == longest for first, last in sizes):
        return out
    pieces = [out[r * longest:r * longest + (last - first)] for r, (first, last) in enumerate(sizes)]
    return torch.cat(pieces, dim=0)
