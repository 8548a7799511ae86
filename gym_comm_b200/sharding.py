"""Multi-GPU sharding of the env batch: one process per GPU, envs split by index, NO collective
on the step path (envs are independent).  torch.distributed is used only for
rendezvous, barriers and max-over-ranks timing."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global env index range owned by `rank`; sizes differ by <= 1."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seed(seed: int, rank: int) -> int:
    """Per-shard device-RNG seed: shards must not replay each other's placements / actions."""
    return (int(seed) * 0x9E3779B97F4A7C15 + int(rank) * 0xD1B54A32D192ED03 + 1) & 0xFFFFFFFFFFFFFFFF


def max_over_ranks(value: float, device="cpu") -> float:
    """Device/wall time of a region = the slowest rank's."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def make_sharded_env(arglist, total_envs: int, seed: int = 0, **kw):
    """This rank's shard of a `total_envs`-env job on its local GPU."""
    from .vec_env import OvercookedVecEnv
    rank, world, local = rank_world()
    lo, hi = shard_range(total_envs, rank, world)
    return OvercookedVecEnv(arglist, num_envs=hi - lo, device=torch.device("cuda", local),
                            seed=shard_seed(seed, rank), **kw), (lo, hi)


def bind_cpu_to_device(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA
    node / PCIe root).  Matters for the host-buffer path only: page-locked buffers allocated afterwards
    land on that node, so eight ranks copying observations to the host do not all cross the socket
    interconnect.  Returns False (and changes nothing) when NVML is unavailable; `OC_NO_AFFINITY=1`
    disables it."""
    if os.environ.get("OC_NO_AFFINITY") == "1":
        return False
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceSetCpuAffinity(h)
        finally:
            pynvml.nvmlShutdown()
        return True
    except Exception:
        return False
