"""Multi-GPU plumbing: one process per GPU, trajectories sharded by rank, no collective in the
integrator.  torch.distributed is used only to agree on timings and to
gather results after the kernels are done; works with NCCL (GPU tensors) and gloo (CPU tensors,
used by the world-size-2 CPU tests)."""

from __future__ import annotations

import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as closest to the GPU (its NUMA node), so that pinned host buffers
    allocated afterwards (first touch) and the copies that use them stay on the GPU's side of the machine -- what
    ``numactl --cpunodebind`` would do for a one-process-per-GPU job.  Best effort: returns False and changes nothing
    if NVML or the affinity call is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            entry = visible.split(",")[device_index].strip()
            handle = pynvml.nvmlDeviceGetHandleByUUID(entry) if entry.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(entry))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous, near-equal slices: rank r owns [lo, hi); sizes differ by at most one."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(array, rank: int, world: int, axis: int = 0):
    """The rank's slice of a batch-major array (NumPy or torch)."""
    lo, hi = shard_bounds(array.shape[axis], rank, world)
    index = [slice(None)] * array.ndim
    index[axis] = slice(lo, hi)
    return array[tuple(index)]


def reduce_max_sum(max_values, sum_values, device="cpu"):
    """max over ranks of `max_values` (timings) and sum over ranks of `sum_values` (work counts),
    as Python floats on every rank.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist

    mx = torch.tensor(list(max_values), dtype=torch.float64, device=device)
    sm = torch.tensor(list(sum_values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx.tolist(), sm.tolist()


def gather_batch(local, n_total: int):
    """Final host/device gather of per-rank results along the batch axis (every rank gets the
    whole array).  Shards may differ in size by one row; they are padded for the collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
