"""Multi-GPU plumbing: pairs are independent, so a batch shards by contiguous index
ranges with no collective on the data path.  Two forms:

* one process per GPU (``bench.py`` under torchrun): ``shard_range`` + ``torch.distributed`` only
  to agree on timing (max over ranks) and, optionally, to gather the per-pair results;
* one process, one host thread and one context per GPU (``align_on_devices``): the shards are
  cut so that every GPU gets the same number of DP cells, each thread makes one C-ABI call
  (ctypes releases the GIL), and the host concatenates the results."""

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


def shard_by_cells(off1: np.ndarray, off2: np.ndarray, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous pair ranges [lo, hi) holding (nearly) equal numbers of DP cells."""
    cells = np.diff(off1).astype(np.int64) * np.diff(off2).astype(np.int64)
    cum = np.concatenate(([0], np.cumsum(cells)))
    n = len(cells)
    cuts = [0]
    for k in range(1, n_shards):
        cuts.append(int(np.searchsorted(cum, cum[-1] * k / n_shards, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [(cuts[k], cuts[k + 1]) for k in range(n_shards)]


def align_on_devices(align_shard, s1, off1, s2, off2, devices):
    """Run ``align_shard(device, s1, off1, s2, off2) -> (cost, ops, ops_off, ops_len, status)`` on
    one contiguous shard per device, concurrently (one host thread per device), and merge the
    results in pair order.  Empty shards are skipped."""
    from concurrent.futures import ThreadPoolExecutor

    ranges = shard_by_cells(off1, off2, len(devices))
    jobs = [(dev, lo, hi) for dev, (lo, hi) in zip(devices, ranges) if hi > lo]

    def work(job):
        dev, lo, hi = job
        a0, a1 = int(off1[lo]), int(off1[hi])
        b0, b1 = int(off2[lo]), int(off2[hi])
        return align_shard(dev, s1[a0:a1], off1[lo:hi + 1] - a0, s2[b0:b1], off2[lo:hi + 1] - b0)

    if len(jobs) == 1:
        parts = [work(jobs[0])]
    else:
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            parts = list(pool.map(work, jobs))
    cost = np.concatenate([p[0] for p in parts])
    if parts[0][1] is None:                       # cost-only
        return cost, None, None, None, None
    ops = np.concatenate([p[1] for p in parts])
    ops_len = np.concatenate([p[3] for p in parts])
    status = np.concatenate([p[4] for p in parts])
    ops_off = np.zeros(len(cost) + 1, dtype=np.int64)
    at, k = 0, 0
    for p in parts:
        n = len(p[0])
        ops_off[k:k + n + 1] = p[2] + at
        at += int(p[2][-1])
        k += n
    return cost, ops, ops_off, ops_len, status
