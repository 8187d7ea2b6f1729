"""The synthetic-BAM writer (instrain_b200/synth_bam.py: bench / test support) against the oracle's independent pure-Python
BAM reader, and the whole host side on the file it writes: C++ read filter = restatement, C++ packer = pileup emulation
record for record, the index leads the threaded packer to the same batches, and profile_bam (engine answered by the
oracle) gives the tables the oracle computes from the emulated events."""
import json
import os

import numpy as np
import pytest

from oracle import bamio, pileup_emul, restate
from test_host_packer import _oracle_filter, compare_bam


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    from instrain_b200 import synth_bam
    path = str(tmp_path_factory.mktemp("synthbam") / "s.bam")
    info = synth_bam.write_bam(path, 24000, 3, 40, 0.02, seed=11)
    return path, info


def test_written_bam_reads_back(bam):
    from instrain_b200 import synth_bam
    path, info = bam
    refs, reads = bamio.read_bam(path)
    assert [n for n, _ in refs] == info["names"] and all(l == 24000 for _, l in refs)
    assert len(reads) == info["n_reads"] == 2 * info["n_pairs"]
    key = [(r.tid, r.pos) for r in reads]
    assert key == sorted(key)                                         # coordinate-sorted
    rng = np.random.Generator(np.random.PCG64(11))
    k = 0
    for tid in range(3):
        ref, rd = synth_bam._scaffold_reads(rng, 24000, 40, 0.02)
        assert "".join("ACTG"[c] for c in ref) == info["seqs"][info["names"][tid]]
        for i in range(0, len(rd["pos"]), 97):
            r = reads[k + i]
            assert (r.tid, r.pos, r.mpos, r.isize, r.nm) == (tid, rd["pos"][i], rd["mpos"][i], rd["isize"][i], rd["nm"][i])
            assert r.flag == (99 if rd["mate"][i] == 0 else 147) and r.cigar == [(0, 150)] and r.mtid == tid
            assert r.seq == "".join("ACTG"[c] for c in rd["base"][i]) and np.array_equal(r.qual, rd["qual"][i])
        k += len(rd["pos"])


def test_host_side_on_the_written_bam(bam):
    from instrain_b200.packer import BamPacker, find_bai, read_bai
    from instrain_b200.profile import iter_batches
    from instrain_b200.read_filter import filter_reads
    path, info = bam
    names, (exp, _, exp_max) = _oracle_filter(path)
    got, _, mx = filter_reads(path, names)
    assert mx == exp_max and got == {s: d for s, d in exp.items() if d} and sum(len(v) for v in got.values()) > 5000
    assert sum(len(v) for v in got.values()) <= info["n_pairs"]
    assert compare_bam(path, got) > 500000                           # C++ packer = emulation, record for record
    first = read_bai(find_bai(path))
    with BamPacker(path) as bp:
        for tid in (2, 0, 1):
            bp.seek(first[tid])
            assert bp.peek_tid() == tid
    one = [b for k, b in iter_batches(path, got, info["seqs"]) if k == "batch"]
    thr = [b for k, b in iter_batches(path, got, info["seqs"], packer_threads=3) if k == "batch"]
    assert len(one) == len(thr) == 1 and one[0]["names"] == thr[0]["names"] == info["names"]
    for a, b in zip(one[0]["parts"], thr[0]["parts"]):
        for key in ("seg_start", "seg_len", "stream", "pair_mm"):
            assert np.array_equal(a[key], b[key]), key


def test_profile_bam_on_the_written_bam(bam, tmp_path, monkeypatch):
    import instrain_b200.profile as P
    from conftest import assert_ld_equal, assert_snv_equal, load_lut
    from test_profile_host_cpu import OracleEngine

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    monkeypatch.setattr(P, "Engine", E)
    monkeypatch.setenv("ISB_NATIVE_STORE", "1")
    path, info = bam
    out = P.profile_bam(path, None, None, str(tmp_path / "s.IS"), s2s=info["seqs"], packer_threads=2)
    res = out.result
    assert not res.failures and res.scaffold_list == info["names"]
    # the same tables from the oracle's own reader + emulation
    refs, reads = bamio.read_bam(path)
    from instrain_b200.read_filter import filter_reads
    r2m, _, _ = filter_reads(path, [n for n, _ in refs])
    lut, dflt = load_lut()
    n_snv = n_ld = 0
    for tid, name in enumerate(info["names"]):
        ev = restate.sort_events(pileup_emul.scaffold_events([r for r in reads if r.tid == tid], r2m[name]))
        exp = restate.profile_events(ev, restate.encode_ref(info["seqs"][name]), lut, dflt, bamio.iterate_splits(24000, 10000))
        t = res.raw_snp_table[res.raw_snp_table["scaffold"] == name]
        assert len(t) == len(exp["snv"]) and sorted(t["position"]) == sorted(exp["snv"]["pos"].tolist())
        l = res.raw_linkage_table[res.raw_linkage_table["scaffold"] == name]
        assert len(l) == len(exp["ld"])
        n_snv += len(t)
        n_ld += len(l)
    assert n_snv > 500 and n_ld > 500
    assert os.path.exists(str(tmp_path / "s.IS" / "output" / "s.IS_SNVs.tsv"))


def test_threaded_read_filter_equals_the_sequential_pass(bam):
    """isb_filter_open_mt (one reader per host thread, scaffolds through the .bai index) gives the same sR2M, tallies, max
    insert and mapping_info report as the sequential pass, on the synthetic file and on the bundled BAM subset."""
    from conftest import GOLDEN
    from instrain_b200.packer import BamPacker
    from instrain_b200.read_filter import filter_reads
    for path in (bam[0], os.path.join(GOLDEN, "c1_G1_subset.bam")):
        with BamPacker(path) as bp:
            names = bp.ref_names
        a = filter_reads(path, names, with_report=True)
        b = filter_reads(path, names, with_report=True, threads=3)
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[3].equals(b[3]) and len(a[0]) > 0
