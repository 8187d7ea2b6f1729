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
