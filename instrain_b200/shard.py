"""Multi-GPU sharding of the hot path: scaffolds -> ranks, and the gather of the final tables.

Every (scaffold, split) is independent in the reference itself (inStrain/profile/profile_utilities.py:115-216), so there
is no data-path collective: each rank (one process per GPU) profiles its own scaffolds and only the final row tables
travel -- the replacement of the reference's Manager.dict / Queue plumbing (profile_controller.py:157-193).
Works on any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
import numpy as np


def lpt_partition(weights, n_ranks):
    """Longest-processing-time bin packing: heaviest scaffold first onto the lightest rank (the reference sorts
    scaffolds by filtered pairs for the same reason, inStrain/profile/fasta.py:103-105).
    Returns a list of index lists, one per rank (indices ascending inside a rank)."""
    order = np.argsort(-np.asarray(weights, dtype=np.float64), kind="stable")
    loads = np.zeros(n_ranks)
    bins = [[] for _ in range(n_ranks)]
    for i in order:
        r = int(np.argmin(loads))
        bins[r].append(int(i))
        loads[r] += weights[i]
    return [sorted(b) for b in bins]


def gather_rows(rows, dst=0, group=None, device=None):
    """Gather variable-length structured row arrays (numpy, any dtype) from all ranks to `dst`.
    One all_gather of the row counts, then one padded gather of the payload bytes.  Returns the concatenated array on
    `dst` (rank order), None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    item = rows.dtype.itemsize
    mx = max(counts) if counts else 0
    payload = torch.zeros(max(mx, 1) * item, dtype=torch.uint8, device=dev)
    if len(rows):
        payload[:len(rows) * item] = torch.from_numpy(np.frombuffer(rows.tobytes(), dtype=np.uint8).copy()).to(dev)
    bufs = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [np.frombuffer(b.cpu().numpy().tobytes()[:c * item], dtype=rows.dtype) for b, c in zip(bufs, counts)]
    return np.concatenate(parts) if parts else rows[:0]


def split_runs(splits, weights, n_ranks):
    """Contiguous runs of the splits of ONE scaffold, one per rank, of about equal weight (a scaffold too large
    for a balanced scaffold-wise partition is sharded by runs of splits; every run is profiled from the reads that overlap
    it, instrain_b200.reads.clip_reads).  splits: [(start, end)] in order; weights: per-split cost (aligned bases, or
    positions).  Returns [(first split, one past the last)] -- at most one run per split."""
    import numpy as np
    n = len(splits)
    k = max(1, min(int(n_ranks), n))
    w = np.asarray(weights, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, k):
        target = cum[-1] * r / k
        i = int(np.searchsorted(cum, target, side="left"))
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        i = max(cuts[-1] + 1, min(i, n - (k - r)))
        cuts.append(i)
    cuts.append(n)
    return [(a, b) for a, b in zip(cuts, cuts[1:])]
