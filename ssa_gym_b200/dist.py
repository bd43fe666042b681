"""Multi-GPU plumbing: the path shards by independent units, so there is NO data-path collective.

Environments (RL mode) or objects (catalog mode) are split into contiguous blocks, one block per rank / GPU
(one process per GPU, `torchrun`).  The only communication is an optional gather of per-environment rewards and
observations onto one device for a learner, done with a single `all_gather` (NCCL over NVLink on GPUs, gloo in
the CPU tests) on a side stream; NCCL has no native gather, and with equal shards all_gather is the cheapest
correct form.  The reference has no counterpart: its parallelism is one env per Ray worker process
(rl_agents/RLLib_PPO_training.py:17).
"""
import os

import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous block [lo, hi) of `total` units owned by `rank`; the first `total % world` ranks get one more."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(total, world):
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def env_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_info()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return rank, local_rank, world


def gather_rows(local, total_rows, stream=None):
    """All-gather a per-rank block of rows (reward[E_r] or obs[E_r, m*12]) into the full [E, ...] tensor on every
    rank.  Shards may differ by one row; they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total_rows, world)
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    if stream is not None and pad.is_cuda:
        with torch.cuda.stream(stream):
            dist.all_gather(out, pad.contiguous())
    else:
        dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def reduce_catalog_stats(max_dpos, trinary_sum, n_objects, argmax_trace_value, argmax_trace_index):
    """Catalog mode (C4): combine per-shard partial reductions {max dpos, sum of trinary counts, (value, global
    index) of the largest trace} into global values.  First-maximum-wins on ties, like np.argmax."""
    import torch
    import torch.distributed as dist
    vals = torch.tensor([max_dpos, trinary_sum, float(n_objects), argmax_trace_value, float(argmax_trace_index)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            vals = vals.cuda()
        allv = [torch.empty_like(vals) for _ in range(dist.get_world_size())]
        dist.all_gather(allv, vals)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = vals.numpy()[None]
    best = None
    for row in allv:  # rank order == index order, strict '>' keeps the first maximum
        if best is None or row[3] > best[0]:
            best = (row[3], int(row[4]))
    return {"max_delta_pos": float(np.max(allv[:, 0])), "trinary_reward": float(np.sum(allv[:, 1]) / np.sum(allv[:, 2]) / 2),
            "argmax_trace": best[1], "max_trace": float(best[0])}
