"""Host packer (instrain_b200/csrc/isb_host.cpp via instrain_b200.packer): BAM -> position-major event columns.
Pure host C++ -- runs without a GPU.  Checked event-for-event against the oracle's htslib-faithful emulation
(oracle/pileup_emul.py), which is itself pinned on the reference's goldens (oracle/validate_against_reference.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import bamio, pileup_emul, restate


def emul_events(reads, r2m):
    ev = restate.sort_events(pileup_emul.scaffold_events(reads, r2m))
    return ev


def compare_bam(bam_path, rdic):
    from instrain_b200 import build
    build.build()
    from instrain_b200.packer import BamPacker
    refs, reads = bamio.read_bam(bam_path)
    by_tid = {}
    for r in reads:
        if r.tid >= 0:
            by_tid.setdefault(r.tid, []).append(r)
    n_ev = 0
    with BamPacker(bam_path) as bp:
        assert bp.ref_names == [n for n, _ in refs] and bp.ref_lens == [l for _, l in refs]
        while True:
            tid = bp.peek_tid()
            if tid < 0:
                break
            name = bp.ref_names[tid]
            r2m = rdic.get(name, {})
            got = bp.pack_scaffold(tid, r2m)
            exp = emul_events(by_tid[tid], r2m)
            for k in ("ref_pos", "base", "qual", "read_id"):
                assert np.array_equal(got[k], exp[k]), (name, k)
            assert np.array_equal(got["pair_mm"], exp["pair_mm"].astype(np.uint8)), name
            assert got["reads_seen"] == len(by_tid[tid])
            n_ev += len(got["ref_pos"])
    return n_ev


def test_packer_matches_emulation_on_fixture_bam():
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    n = compare_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"), rdic)
    assert n > 100000


def test_packer_set_mode_and_offsets():
    """R2M as a set (--skip_mm_profiling: pair_mm == 0) and batch offsets for positions / pair ids."""
    from instrain_b200 import build
    build.build()
    from instrain_b200.packer import BamPacker
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    path = os.path.join(GOLDEN, "c1_G1_subset.bam")
    with BamPacker(path) as bp:
        tid = bp.peek_tid()
        name = bp.ref_names[tid]
        a = bp.pack_scaffold(tid, set(rdic[name].keys()), pos_offset=1000, pair_id_offset=50)
    with BamPacker(path) as bp:
        b = bp.pack_scaffold(bp.peek_tid(), rdic[name])
    assert (a["pair_mm"] == 0).all() and len(a["pair_mm"]) == len(b["pair_mm"])
    assert np.array_equal(a["ref_pos"], b["ref_pos"] + 1000) and np.array_equal(a["read_id"], b["read_id"] + 50)
    assert np.array_equal(a["base"], b["base"]) and np.array_equal(a["qual"], b["qual"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/test_data"), reason="reference tree not present")
def test_packer_matches_emulation_on_reference_bam():
    td = "/root/reference/test/test_data"
    rdic = json.load(open(os.path.join(td, "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.forRC.IS/raw_data/Rdic.json")))
    n = compare_bam(os.path.join(td, "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.sorted.bam"), rdic)
    assert n == 1876216


# ---- read filter: BAM -> sR2M -----------------------------------------------------------------------------
def _oracle_filter(bam_path):
    from oracle import read_filter as orf
    refs, reads = bamio.read_bam(bam_path)
    by = {}
    for r in reads:
        if r.tid >= 0:
            by.setdefault(refs[r.tid][0], []).append(r)
    return [n for n, _ in refs], orf.filter_pairs({s: orf.pair2info(rs) for s, rs in by.items()})


def test_read_filter_matches_restatement_on_fixture_bam():
    from instrain_b200 import build
    build.build()
    from instrain_b200.read_filter import filter_reads
    path = os.path.join(GOLDEN, "c1_G1_subset.bam")
    names, (exp, exp_tal, exp_max) = _oracle_filter(path)
    got, tal, mx = filter_reads(path, names)
    assert mx == exp_max
    assert got == {s: d for s, d in exp.items() if d}
    for s, t in exp_tal.items():
        if t["pass_pairing_filter"]:
            assert tal[s] == t, s
    # non-default thresholds
    from oracle import read_filter as orf
    refs, reads = bamio.read_bam(path)
    by = {}
    for r in reads:
        by.setdefault(refs[r.tid][0], []).append(r)
    exp3, _, _ = orf.filter_pairs({s: orf.pair2info(rs) for s, rs in by.items()}, 0.98, 20, 2, 100)
    got3, _, _ = filter_reads(path, names, 0.98, 20, 2, 100)
    assert got3 == {s: d for s, d in exp3.items() if d}


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/test_data"), reason="reference tree not present")
def test_read_filter_reproduces_reference_rdic():
    """The reference's stored Rdic.json (sR2M) and mapping_info tallies for the bundled BAM, default thresholds."""
    import pandas as pd
    from instrain_b200.packer import BamPacker
    from instrain_b200.read_filter import filter_reads
    td = "/root/reference/test/test_data"
    isd = os.path.join(td, "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.forRC.IS/raw_data")
    bam = os.path.join(td, "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.sorted.bam")
    with BamPacker(bam) as bp:
        names = bp.ref_names
    got, tal, mx = filter_reads(bam, names)
    rdic = json.load(open(os.path.join(isd, "Rdic.json")))
    assert got == rdic and sum(len(v) for v in got.values()) == 7179 and mx == 960.0
    mi = pd.read_csv(os.path.join(isd, "mapping_info.csv.gz"))
    mi = mi[mi.scaffold != "all_scaffolds"].set_index("scaffold")
    for s in mi.index:
        if s in tal:
            assert all(int(mi.loc[s, c]) == tal[s][c] for c in tal[s]), s
