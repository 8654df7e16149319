"""Multi-GPU plumbing: environments are independent, so a job shards contiguous blocks of envs over
the ranks (one process per GPU) and the ONLY collective is one all-reduce of episode statistics per
evaluation batch — mirroring ``np.mean / np.std(reward_list)`` of the reference's evaluation loops
(MA_inv_management.py:591-595).

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing; the statistics themselves
come from ``imx_return_stats`` on the device.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index: int) -> int:
    """Pins this process to the CPU cores NVML reports as local to ``device_index`` (its NUMA node), so that
    pinned host buffers allocated afterwards (first touch) sit next to the GPU's PCIe root port — what the
    host-buffer path (``imx_step_host``) needs when several ranks share a box.  Returns the number of cores
    bound to, 0 if NVML or the affinity call is unavailable (then nothing changes)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = {64 * w + b for w, bits in enumerate(mask) for b in range(64) if (bits >> b) & 1}
        cores &= set(os.sched_getaffinity(0))
        if not cores:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


def shard_range(num_envs_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of global env indices owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(num_envs_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_config(config: dict, num_envs_total: int, rank: int = None, world_size: int = None) -> dict:
    """Returns a copy of ``config`` for this rank's shard: ``num_envs`` = shard size, ``env_offset`` =
    first global env index (keys the Philox stream, so trajectories do not depend on the shard count)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(num_envs_total, rank, world_size)
    out = dict(config)
    out["num_envs"] = hi - lo
    out["env_offset"] = lo
    return out


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """The single collective: element-wise SUM of [n, Σ, Σ², (Σ_i, Σ_i²)...] over the ranks."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats: torch.Tensor) -> Dict[str, object]:
    """mean / population std (numpy's default ddof=0, as in the reference) from reduced statistics."""
    s = stats.detach().cpu().double().tolist()
    n = s[0]
    mean = s[1] / n
    var = max(s[2] / n - mean * mean, 0.0)
    out = {"n": int(n), "mean": mean, "std": math.sqrt(var)}
    per_agent = []
    for k in range(3, len(s), 2):
        mu = s[k] / n
        per_agent.append({"mean": mu, "std": math.sqrt(max(s[k + 1] / n - mu * mu, 0.0))})
    if per_agent:
        out["per_agent"] = per_agent
    return out
