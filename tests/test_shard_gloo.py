"""Multi-process path on CPU: scaffold sharding + gather of the final tables over torch.distributed (gloo, world 2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_balances_and_covers():
    from instrain_b200.shard import lpt_partition
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1000, 57).astype(float)
    for n in (1, 2, 4, 8):
        bins = lpt_partition(w, n)
        assert sorted(i for b in bins for i in b) == list(range(57))
        loads = [w[b].sum() for b in bins]
        assert max(loads) - min(loads) <= w.max()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from instrain_b200 import _cabi
    from instrain_b200.shard import gather_rows, lpt_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = [5, 3, 9, 1, 7, 2]
    mine = lpt_partition(weights, world)[rank]
    snv = np.zeros(sum(weights[i] for i in mine), dtype=_cabi.SNV_DT)          # rows "produced" by this rank
    k = 0
    for i in mine:
        snv["pos"][k:k + weights[i]] = 1000 * i + np.arange(weights[i])
        snv["mm"][k:k + weights[i]] = i
        k += weights[i]
    ld = np.zeros(rank * 3, dtype=_cabi.LD_DT)                               # rank 0 contributes no linkage rows
    ld["r2"] = rank + 0.5
    g_snv, g_ld = gather_rows(snv), gather_rows(ld)
    if rank == 0:
        q.put((g_snv["pos"].tolist(), g_snv["mm"].tolist(), g_ld["r2"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_tables_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    pos, mm, r2 = q.get(timeout=100)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    weights = [5, 3, 9, 1, 7, 2]
    exp = sorted(1000 * i + j for i, w in enumerate(weights) for j in range(w))
    assert sorted(pos) == exp and len(mm) == sum(weights)
    assert r2 == [1.5, 1.5, 1.5]
