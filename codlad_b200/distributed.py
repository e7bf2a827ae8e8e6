"""Multi-GPU plumbing of the sampling path: one process per GPU, units (protein frame x ensemble member) sharded across
ranks with NO collective on the data path (nothing in the denoiser, the DDPM update, the VQ-VAE
decode or ic_to_xyz mixes batch rows).  torch.distributed is used for the launch barrier and the max-over-ranks timing only."""
from __future__ import annotations

import os

import torch


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str = "nccl", local_rank: int = 0):
    """Join the job torchrun (or the driver) started: one process per GPU, rendezvous from MASTER_ADDR / MASTER_PORT.  A
    single-process run (WORLD_SIZE unset or 1) does not create a process group."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    """Launch barrier + device drain (both sides of every timed region)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def shard_bounds(n_units: int, rank: int, world: int):
    """Contiguous balanced partition: rank r owns units [lo, hi); sizes differ by at most one."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs, world: int):
    """Longest-processing-time assignment of units with unequal cost (e.g. proteins of different length: cost ~ L*K):
    returns a list of unit-index lists, one per rank.  Deterministic (ties broken by unit index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (loads[q], q))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(v) for v in out]


def max_over_ranks(value: float, device=None) -> float:
    """Device-side timing of a multi-GPU run is the MAX over ranks (the job is done when the slowest rank is)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count: int, device=None):
    """Units processed per rank (whole-job throughput = sum of these / max-over-ranks time)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device if device is not None else "cpu")
    t[dist.get_rank()] = int(count)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]
