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


def _profile_worker(rank, world, port, q, isp):
    """profile_bam_distributed on the bundled BAM subset, world size 2 over gloo; the CUDA engine answered by the oracle
    (test stub: no GPU here -- the kernels have their own parity tests)."""
    import json
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import instrain_b200.profile as P
    from conftest import GOLDEN
    from test_profile_host_cpu import OracleEngine

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    P.Engine = E
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["ISB_NATIVE_STORE"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    out = P.profile_bam_distributed(bam, None, rdic, isp if rank == 0 else None, s2s=seqs, device=0)
    if rank == 0:
        r = out.result
        q.put((sorted(r.scaffold_list), r.failures, len(r.raw_snp_table), len(r.raw_linkage_table), len(r.cumulative_scaffold_table),
               r.raw_snp_table.to_json(), r.raw_linkage_table[["scaffold", "position_A", "position_B", "mm", "countAB", "total"]].to_json(),
               sorted(r.timing)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_profile_bam_distributed_gloo_world2(tmp_path):
    """One BAM, two ranks: the scaffolds are LPT-partitioned, every rank profiles its share, rank 0 ends up with exactly the
    tables (and the SNVprofile directory) of a single-process run."""
    import io
    import json
    import pandas as pd
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import GOLDEN
    from test_profile_host_cpu import OracleEngine
    from instrain_b200.profile import profile_scaffolds
    from instrain_b200.store import SNVprofileStore
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    isp = str(tmp_path / "dist.IS")
    procs = [ctx.Process(target=_profile_worker, args=(r, 2, port, q, isp)) for r in range(2)]
    for p in procs:
        p.start()
    names, failures, n_snv, n_ld, n_sum, snv_json, ld_json, timing = q.get(timeout=280)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    one = profile_scaffolds(os.path.join(GOLDEN, "c1_G1_subset.bam"), rdic, seqs, engine=OracleEngine())
    assert names == sorted(rdic) and failures == [] and timing == ["rank0_profile_scaffolds_s", "rank1_profile_scaffolds_s"]
    assert (n_snv, n_ld, n_sum) == (len(one.raw_snp_table), len(one.raw_linkage_table), len(one.cumulative_scaffold_table))
    key = ["scaffold", "position", "mm"]
    a = pd.read_json(io.StringIO(snv_json)).sort_values(key).reset_index(drop=True)
    b = one.raw_snp_table.sort_values(key).reset_index(drop=True)
    for c in ["scaffold", "position", "mm", "ref_base", "A", "C", "T", "G", "con_base", "var_base", "allele_count", "class", "cryptic"]:
        assert (a[c].values == b[c].values).all(), c
    key = ["scaffold", "position_A", "position_B", "mm"]
    a = pd.read_json(io.StringIO(ld_json)).sort_values(key).reset_index(drop=True)
    b = one.raw_linkage_table.sort_values(key).reset_index(drop=True)
    for c in key + ["countAB", "total"]:
        assert (a[c].values == b[c].values).all(), c
    S = SNVprofileStore(isp)
    assert sorted(S.get("scaffold_list")) == sorted(rdic) and len(S.get("raw_linkage_table")) == n_ld
    covT = S.get("covT")
    for s, sp in one.scaffolds.items():
        assert set(covT[s]) == set(sp.covT)


def _run_worker(rank, world, port, q, bam, seqs_json, isp):
    """profile_bam_distributed on a BAM with ONE scaffold: the scaffold is profiled by runs of splits, one per rank."""
    import json
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import instrain_b200.profile as P
    from instrain_b200.read_filter import filter_reads
    from test_profile_host_cpu import OracleEngine

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    P.Engine = E
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["ISB_NATIVE_STORE"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = json.load(open(seqs_json))
    r2m, _, _ = filter_reads(bam, list(seqs))
    calls = []
    real = P.profile_scaffold_run

    def spy(*a, **k):
        calls.append(a[4])
        return real(*a, **k)

    P.profile_scaffold_run = spy
    out = P.profile_bam_distributed(bam, None, r2m, isp if rank == 0 else None, s2s=seqs, device=0, seed=7)
    assert len(calls) == 1 and len(calls[0]) >= 2                         # every rank got one run of several splits
    if rank == 0:
        r = out.result
        q.put((r.scaffold_list, r.failures, r.raw_snp_table.to_json(), r.raw_linkage_table.to_json(double_precision=15),
               r.cumulative_scaffold_table.to_json(double_precision=15), {int(m): (s.index.tolist(), s.values.tolist()) for m, s in
                                                        r.scaffolds[r.scaffold_list[0]].covT.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_profile_bam_distributed_split_runs_gloo_world2(tmp_path):
    """SURVEY 8(e), the single large scaffold: one scaffold, two ranks -- each rank profiles a contiguous run of its splits
    from the reads that overlap the run; rank 0 ends up with the tables, the per-position series and the summary row of a
    single-process run (engine answered by the oracle: no GPU here)."""
    import io
    import json
    import pandas as pd
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from instrain_b200 import synth_bam
    from instrain_b200.profile import profile_scaffolds
    from instrain_b200.read_filter import filter_reads
    from test_profile_host_cpu import OracleEngine
    bam = str(tmp_path / "one.bam")
    info = synth_bam.write_bam(bam, 52000, 1, 30, 0.02, seed=23)
    seqs_json = str(tmp_path / "seqs.json")
    json.dump(info["seqs"], open(seqs_json, "w"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 11) % 2000
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, q, bam, seqs_json, str(tmp_path / "runs.IS"))) for r in range(2)]
    for p in procs:
        p.start()
    names, failures, snv_json, ld_json, sum_json, cov = q.get(timeout=280)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    r2m, _, _ = filter_reads(bam, info["names"])
    one = profile_scaffolds(bam, r2m, info["seqs"], engine=OracleEngine(), seed=7)
    assert names == info["names"] and failures == []
    a = pd.read_json(io.StringIO(snv_json)).sort_values(["position", "mm"]).reset_index(drop=True)
    b = one.raw_snp_table.sort_values(["position", "mm"]).reset_index(drop=True)
    assert len(a) == len(b) > 300
    for c in ["scaffold", "position", "mm", "ref_base", "A", "C", "T", "G", "con_base", "var_base", "allele_count", "class", "cryptic"]:
        assert (a[c].values == b[c].values).all(), c
    key = ["position_A", "position_B", "mm"]
    a = pd.read_json(io.StringIO(ld_json)).sort_values(key).reset_index(drop=True)
    b = one.raw_linkage_table.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 300
    for c in key + ["countAB", "countAb", "countaB", "countab", "total", "r2", "d_prime", "r2_normalized", "d_prime_normalized"]:
        assert np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=0, atol=1e-12, equal_nan=True), c
    a = pd.read_json(io.StringIO(sum_json)).sort_values("mm").reset_index(drop=True)
    b = one.cumulative_scaffold_table.sort_values("mm").reset_index(drop=True)
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    for c in a.columns:
        if c != "scaffold":
            assert np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=1e-12, atol=0, equal_nan=True), c
    ref_cov = one.scaffolds[info["names"][0]].covT
    assert set(cov) == set(int(m) for m in ref_cov)
    for m, s in ref_cov.items():
        assert cov[int(m)][0] == s.index.tolist() and cov[int(m)][1] == s.values.tolist()
