"""Multi-GPU plumbing: pairs are independent, so a batch shards by contiguous index
ranges, one process per GPU, with no collective on the data path.
``torch.distributed`` is used only to agree on timing (max over ranks) and,
optionally, to gather the fixed-size per-pair results on rank 0."""

from __future__ import annotations

import numpy as np


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of `n_units` for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_int32(local: np.ndarray, n_total: int, rank: int, world: int, group=None):
    """Gather per-pair int32 results (costs, scores, lengths) of all ranks on every
    rank, in pair order.  Works on any torch.distributed backend (gloo on CPU,
    nccl with device tensors is not needed for a few MB)."""
    import torch
    import torch.distributed as dist

    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = np.zeros(width, dtype=np.int32)
    pad[: len(local)] = local
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty(n_total, dtype=np.int32)
    for (lo, hi), t in zip(sizes, parts):
        out[lo:hi] = t.cpu().numpy()[: hi - lo]
    return out
