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
