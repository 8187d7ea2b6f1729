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
            assert tal[s] == {k: t[k] for k in tal[s]}, s      # the six threshold tallies of the default mode
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


def _canon(ev, keep=None):
    """(position, pair, base) triples of the events, sorted."""
    sel = slice(None) if keep is None else keep
    pos, rid, base = ev["ref_pos"][sel], ev["read_id"][sel], np.minimum(ev["base"][sel], 4)
    o = np.lexsort((base, rid, pos))
    return pos[o], rid[o], base[o]


@pytest.mark.parametrize("bam,r2m_json", [("c1_G1_subset.bam", "c1_G1_subset_r2m.json"), ("small_scaffold.bam", None)])
def test_read_major_packer_matches_event_packer(bam, r2m_json):
    """isb_pack_scaffold_reads (aligned segments, one-hot codes) encodes exactly the events of isb_pack_scaffold that
    pass min_qual -- same positions, pair ids (BAM order of first appearance), bases -- and obeys the layout rules."""
    from instrain_b200 import build, reads
    build.build()
    from instrain_b200.packer import BamPacker
    path = os.path.join(GOLDEN, bam)
    if r2m_json:
        rdic = json.load(open(os.path.join(GOLDEN, r2m_json)))
    else:
        meta = json.load(open(os.path.join(GOLDEN, "small_scaffold.json")))
        rdic = {meta["scaffold"]: meta["r2m"]}
    n_tot = 0
    with BamPacker(path) as pa, BamPacker(path) as pb:
        while True:
            tid = pa.peek_tid()
            if tid < 0:
                break
            name = pa.ref_names[tid]
            r2m = rdic.get(name, {})
            ev = pa.pack_scaffold(tid, r2m, pos_offset=700, pair_id_offset=11)
            pb.peek_tid()
            part = pb.pack_scaffold_reads(tid, r2m, pos_offset=700, pair_id_offset=11, min_qual=30)
            rd = reads.concat_streams([part])
            assert np.array_equal(part["pair_mm"], ev["pair_mm"])
            assert part["reads_seen"] == ev["reads_seen"] and part["reads_packed"] == ev["reads_packed"]
            ok = ev["qual"] >= 30
            assert part["n_events"] == int(ok.sum())
            back = reads.reads_to_events(rd)
            for x, y in zip(_canon(back), _canon(ev, ok)):
                assert np.array_equal(x, y), name
            if rd["n_segs"]:
                nw = ((rd["seg_start"].astype(np.int64) & 7) + rd["seg_len"].astype(np.int64) + 7) // 8   # position-aligned words
                assert rd["seg_word"][0] == 1 and np.all(np.diff(rd["seg_word"]) == nw[:-1] + 1)
                assert np.all(np.diff(rd["seg_start"]) >= 0) and rd["seg_len"].min() >= 1 and rd["max_seg_len"] <= 256
                assert rd["n_words"] % 4 == 0 and rd["seg_word"][-1] + nw[-1] < rd["n_words"]
            n_tot += part["n_events"]
    assert n_tot > 1000


def test_compact_reads_round_trip():
    """reads.compact_reads: decoding the 3-bit units gives back exactly the nibbles of the read-major stream."""
    from oracle import synth
    from instrain_b200 import reads
    batch = synth.make_batch(4000, 40, 0.02, 3, skip_mm=False, n_frac=0.003)
    for kw in (dict(), dict(max_len=37, odd_blocks=True)):
        rd = reads.events_to_reads(batch, **kw)
        c = reads.compact_reads(rd)
        nw = ((rd["seg_start"].astype(np.int64) & 7) + rd["seg_len"].astype(np.int64) + 7) // 8
        assert c["n_units"] == nw.sum() == len(c["base2"]) == len(c["pass"])
        k = np.arange(c["n_units"]) - np.repeat(np.cumsum(nw) - nw, nw)
        w = rd["words"][np.repeat(rd["seg_word"], nw) + k]
        dec = np.zeros(c["n_units"], dtype=np.uint32)
        for t in range(8):
            code = (c["base2"].astype(np.uint32) >> (2 * t)) & 3
            ok = (c["pass"].astype(np.uint32) >> t) & 1
            dec |= (ok << code) << np.uint32(4 * t)
        assert np.array_equal(dec, w)
        # 3 bytes per 8 bases + 10 bytes per segment, against 4 bytes per 8 bases + separators + 18 bytes per segment
        assert 3 * c["n_units"] + 10 * c["n_segs"] < 0.75 * (4 * rd["n_words"] + 18 * rd["n_segs"])


REF_BAM = "/root/reference/test/test_data/N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.sorted.bam"


def _sequential(bam, rdic):
    from instrain_b200.packer import BamPacker
    out = {}
    with BamPacker(bam) as bp:
        names = bp.ref_names
        while True:
            tid = bp.peek_tid()
            if tid < 0:
                break
            out[tid] = bp.pack_scaffold_reads(tid, rdic.get(names[tid], {}))
    return names, out


def _assert_same_pack(a, b):
    for k in ("seg_start", "seg_len", "seg_pair", "seg_word", "stream", "nev_pos", "nev_pair", "pair_mm"):
        assert np.array_equal(a[k], b[k]), k
    assert a["reads_seen"] == b["reads_seen"] and a["n_events"] == b["n_events"] and a["max_seg_len"] == b["max_seg_len"]


def test_parallel_packing_equals_sequential_unindexed():
    """Several packer threads, each seeking to its scaffolds (offsets scanned from the BAM itself: the fixture has no
    .bai), produce exactly the sequential packer's segments, in job order."""
    from instrain_b200 import packer
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    names, seq = _sequential(bam, rdic)
    assert packer.find_bai(bam) is None
    first = packer.scan_scaffold_offsets(bam)
    assert [t for t, f in enumerate(first) if f is not None] == sorted(seq)
    jobs = [(tid, rdic.get(names[tid], {}), 0) for tid in sorted(seq, reverse=True)]      # any order, not only file order
    for threads in (1, 3):
        got = list(packer.pack_scaffolds_parallel(bam, jobs, threads))
        assert len(got) == len(jobs)
        for (tid, _, _), g in zip(jobs, got):
            _assert_same_pack(g, seq[tid])
    shifted = list(packer.pack_scaffolds_parallel(bam, [(jobs[0][0], jobs[0][1], 4096)], 2))[0]   # pos_offset is honoured
    assert np.array_equal(shifted["seg_start"], seq[jobs[0][0]]["seg_start"] + 4096)


@pytest.mark.skipif(not os.path.exists(REF_BAM + ".bai"), reason="reference test data not present")
def test_bai_offsets_and_parallel_packing_on_reference_bam():
    """read_bai (the reference's own samtools index) == offsets scanned from the BAM; 178 scaffolds packed by 4 threads ==
    sequential."""
    from instrain_b200 import packer
    rdic = json.load(open(REF_BAM.replace(".sorted.bam", ".forRC.IS") + "/raw_data/Rdic.json"))
    first = packer.read_bai(REF_BAM + ".bai")
    assert first == packer.scan_scaffold_offsets(REF_BAM)
    names, seq = _sequential(REF_BAM, rdic)
    jobs = [(tid, rdic.get(names[tid], {}), 0) for tid in range(len(names)) if first[tid] is not None]
    got = list(packer.pack_scaffolds_parallel(REF_BAM, jobs, 4))
    for (tid, _, _), g in zip(jobs, got):
        _assert_same_pack(g, seq[tid])


@pytest.mark.parametrize("max_batch_events", [400_000_000, 1])
def test_iter_batches_parallel_equals_sequential(max_batch_events):
    """profile_bam's batch stream: packing the scaffolds on several host threads (index seeks, pair ids shifted afterwards)
    yields the same batches -- segments, word streams, pair ids, reference codes, splits -- as the sequential pass, with
    everything in one batch and with one batch per scaffold."""
    from instrain_b200 import reads as reads_mod
    from instrain_b200.profile import iter_batches
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    a = list(iter_batches(bam, rdic, seqs, max_batch_events=max_batch_events, packer_threads=1))
    b = list(iter_batches(bam, rdic, seqs, max_batch_events=max_batch_events, packer_threads=3))
    assert len(a) == len(b) == (1 if max_batch_events > 1 else len(rdic))
    for (ka, x), (kb, y) in zip(a, b):
        assert ka == kb == "batch" and x["names"] == y["names"] and x["off"] == y["off"] and x["splits"] == y["splits"]
        assert (x["L"], x["n_pairs"], x["n_events"]) == (y["L"], y["n_pairs"], y["n_events"])
        assert np.array_equal(np.concatenate(x["ref"]), np.concatenate(y["ref"]))
        assert np.array_equal(np.concatenate(x["pair_mm"]), np.concatenate(y["pair_mm"]))
        rx, ry = reads_mod.concat_streams(x["parts"]), reads_mod.concat_streams(y["parts"])
        for k in ("seg_start", "seg_len", "seg_pair", "seg_word", "words", "nev_pos", "nev_pair"):
            assert np.array_equal(rx[k], ry[k]), k
        assert rx["n_segs"] == ry["n_segs"] and rx["n_words"] == ry["n_words"] and rx["max_seg_len"] == ry["max_seg_len"]
        assert rx["seg_pair"].max() < x["n_pairs"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/test_data"), reason="reference tree not present")
@pytest.mark.parametrize("bam", ["N5_271_010G1_scaffold_963_Ns.fasta.sorted.bam", "SmallScaffold.fa.sorted.bam",
                                 "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G2.sorted.bam"])
def test_packer_and_filter_on_reference_edge_case_bams(bam):
    """The reference's edge-case inputs (N reference at ~185x with 26 k reads, a tiny scaffold, the second sample): the C++
    read filter gives the restatement's sR2M, and the C++ packer the emulation's events, record for record.  For the Ns
    BAM the packed events are exactly the committed fixture the oracle / CUDA tests profile (tests/golden/c963_Ns.npz)."""
    from instrain_b200.read_filter import filter_reads
    path = os.path.join("/root/reference/test/test_data", bam)
    names, (exp, _, _) = _oracle_filter(path)
    got, _, _ = filter_reads(path, names)
    assert {s: d for s, d in got.items() if d} == {s: d for s, d in exp.items() if d}
    n = compare_bam(path, exp)
    assert n > 1000
    if "963_Ns" in bam:
        z = np.load(os.path.join(GOLDEN, "c963_Ns.npz"))
        from instrain_b200.packer import BamPacker
        with BamPacker(path) as bp:
            ev = bp.pack_scaffold(bp.peek_tid(), exp[names[0]])
        for k in ("ref_pos", "base", "qual", "read_id"):
            assert np.array_equal(ev[k], z[k]), k
        assert np.array_equal(ev["pair_mm"], z["pair_mm"].astype(np.uint8))


def test_profile_scaffolds_host_loop_without_gpu():
    """The host loop of profile_bam (batch stream -> transfer format -> engine call -> failure handling) driven by a stub
    engine that records what it is handed and then fails like a crashed batch: the reference's behaviour is to log the
    scaffolds of the batch as failures and carry on (profile_utilities.py:92-112)."""
    from instrain_b200 import reads as reads_mod
    from instrain_b200.profile import profile_scaffolds
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))

    class StubEngine:
        def __init__(self):
            self.calls = []

        def profile_batch(self, ev, ref_codes, splits, **kw):
            self.calls.append((len(ev["pair_mm"]), len(ref_codes), len(splits), sorted(k for k in ("reads", "cols") if k in kw),
                               {k: v for k, v in kw.items() if k in ("reads", "cols")}))
            raise RuntimeError("no GPU in this test")

    seen = {}
    for transfer, threads in (("segments", 1), ("delta", 1), ("cols", 2)):
        eng = StubEngine()
        res = profile_scaffolds(bam, rdic, seqs, engine=eng, b200_transfer=transfer, packer_threads=threads)
        assert sorted(res.failures) == sorted(rdic) and res.scaffold_list == [] and len(res.raw_snp_table) == 0
        n_pairs, L, n_splits, keys, fmt = eng.calls[0]              # the whole batch; then it is bisected down to
        assert len(eng.calls) == 2 * len(rdic) - 1                    # single scaffolds: 2 n - 1 calls
        assert L == sum(len(seqs[s]) for s in rdic) and n_pairs == sum(len(v) for v in rdic.values())
        seen[transfer] = fmt
    rd = seen["segments"]["reads"]
    assert "mis_word" in seen["delta"]["reads"] and np.array_equal(seen["delta"]["reads"]["seg_start"], rd["seg_start"])
    ref = np.concatenate([__import__("instrain_b200.profile", fromlist=["encode_reference"]).encode_reference(seqs[s])
                          for s in seqs if s in rdic])
    _, _, words = reads_mod.delta_to_words(seen["delta"]["reads"], ref)
    assert int(words.sum()) > 0
    assert seen["cols"]["cols"]["n_chunks"] > 0 and int((seen["cols"]["cols"]["ids"] >= 0).sum()) == int(
        (((rd["seg_start"].astype(np.int64) & 7) + rd["seg_len"].astype(np.int64) + 7) // 8).sum())


def test_failed_batch_is_bisected_to_the_offending_scaffold():
    """One scaffold the engine cannot take (the device's envelope, memory ...) costs that scaffold only: the batch is
    retried in halves and every other scaffold comes out exactly as in an undisturbed run (the reference loses one split to
    an exception, profile_utilities.py:104-111)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_profile_host_cpu import OracleEngine
    from instrain_b200.profile import profile_scaffolds
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    good = profile_scaffolds(bam, rdic, seqs, engine=OracleEngine())
    victim = sorted(rdic, key=lambda s: len(rdic[s]))[len(rdic) // 2]
    L_victim = len(seqs[victim])

    class Picky(OracleEngine):
        def profile_batch(self, ev, ref_codes, splits, **kw):
            # the victim is recognised by its reference sequence sitting somewhere in the batch
            from instrain_b200.profile import encode_reference
            v = encode_reference(seqs[victim])
            ref = np.asarray(ref_codes)
            for off in range(0, len(ref) - L_victim + 1):
                if ref[off] == v[0] and np.array_equal(ref[off:off + L_victim], v):
                    raise RuntimeError("ISB_ERR_UNSUPPORTED (simulated)")
            return super().profile_batch(ev, ref_codes, splits, **kw)

    res = profile_scaffolds(bam, rdic, seqs, engine=Picky())
    assert res.failures == [victim] and sorted(res.scaffold_list) == sorted(s for s in rdic if s != victim)
    key = ["scaffold", "position", "mm"]
    a = res.raw_snp_table.sort_values(key).reset_index(drop=True)
    b = good.raw_snp_table[good.raw_snp_table["scaffold"] != victim].sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 500 and a.equals(b)
    key = ["scaffold", "position_A", "position_B", "mm"]
    a = res.raw_linkage_table.sort_values(key).reset_index(drop=True)
    b = good.raw_linkage_table[good.raw_linkage_table["scaffold"] != victim].sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 500
    for c in key + ["countAB", "countab", "total"]:
        assert (a[c].values == b[c].values).all(), c
    for s in res.scaffold_list:
        for mm, ser in good.scaffolds[s].covT.items():
            assert ser.equals(res.scaffolds[s].covT[mm])


def test_batches_close_on_the_cell_budget():
    """L x M cells bound a batch (dense per-position outputs: 24 B per cell on the device): a small budget splits the
    stream into several batches whose union is the single batch."""
    from instrain_b200.profile import iter_batches
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    one = [b for k, b in iter_batches(bam, rdic, seqs) if k == "batch"]
    assert len(one) == 1
    Lmax = max(len(seqs[s]) for s in rdic)
    for threads in (1, 2):
        many = [b for k, b in iter_batches(bam, rdic, seqs, max_batch_cells=Lmax * 16, packer_threads=threads) if k == "batch"]
        assert len(many) > 1 and [n for b in many for n in b["names"]] == one[0]["names"]
        assert sum(b["n_events"] for b in many) == one[0]["n_events"] and all(b["L"] * b["M"] <= Lmax * 16 or len(b["names"]) == 1 for b in many)


def test_priority_read_files(tmp_path):
    from instrain_b200.read_filter import load_priority_reads
    (tmp_path / "a.fastq").write_text("@r1 extra\nACGT\n+\nIIII\n@r2\nAC\n+\nII\n")
    (tmp_path / "b.txt").write_text("r3\nr4\n")
    import gzip
    with gzip.open(str(tmp_path / "c.txt.gz"), "wt") as f:
        f.write("r5\n")
    assert load_priority_reads(str(tmp_path / "a.fastq")) == {"r1 extra", "r2"}
    assert load_priority_reads(str(tmp_path / "b.txt")) == {"r3", "r4"}
    assert load_priority_reads(str(tmp_path / "c.txt.gz")) == {"r5"}


def test_iter_batches_honours_fdb_splits():
    """The split table the reference hands over (Fdb, inStrain/profile/fasta.py:30-73) is used as given; without it the same
    geometry is derived from window_length."""
    import pandas as pd
    from instrain_b200.profile import iter_batches
    from instrain_b200.synth import iterate_splits
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    rows = []
    for s in rdic:
        L = len(seqs[s])
        cuts = [0, L // 3, 2 * L // 3, L]
        for i in range(3):
            rows.append(dict(scaffold=s, split_number=i, start=cuts[i], end=cuts[i + 1] - 1))
    Fdb = pd.DataFrame(rows).sample(frac=1.0, random_state=0)                     # any row order
    (_, with_fdb), = iter_batches(bam, rdic, seqs, Fdb=Fdb)
    (_, plain), = iter_batches(bam, rdic, seqs, window_length=500)
    exp_fdb, exp_plain = [], []
    for name, off in zip(with_fdb["names"], with_fdb["off"]):
        L = len(seqs[name])
        cuts = [0, L // 3, 2 * L // 3, L]
        exp_fdb += [(cuts[i] + off, cuts[i + 1] - 1 + off) for i in range(3)]
        exp_plain += [(s + off, e + off) for s, e in iterate_splits(L, 500)]
    assert with_fdb["splits"] == exp_fdb and plain["splits"] == exp_plain and len(exp_plain) > len(exp_fdb)


# ---- corrupt / truncated input: a partial pass must raise, never return a partial profile (pysam / htslib raise) ----------

def _damaged_copies(tmp_path):
    src = os.path.join(GOLDEN, "c1_G1_subset.bam")
    raw = open(src, "rb").read()
    out = {}
    cut = tmp_path / "cut_half.bam"
    cut.write_bytes(raw[:len(raw) // 2])                                  # cut inside a BGZF block
    out["cut"] = str(cut)
    # cut exactly at a BGZF block boundary, in the middle of the file: only the missing end-of-file block gives it away
    o, bounds = 0, []
    while o + 18 <= len(raw):
        bsize = (raw[o + 16] | (raw[o + 17] << 8)) + 1
        o += bsize
        bounds.append(o)
    edge = tmp_path / "cut_at_block.bam"
    edge.write_bytes(raw[:bounds[len(bounds) // 2]])
    out["edge"] = str(edge)
    bad = bytearray(raw)
    mid = len(raw) // 2
    rng = np.random.default_rng(1)
    bad[mid:mid + 2000] = rng.integers(0, 256, 2000, dtype=np.uint8).tobytes()
    cor = tmp_path / "corrupt.bam"
    cor.write_bytes(bytes(bad))
    out["corrupt"] = str(cor)
    return out


@pytest.mark.parametrize("kind", ["cut", "edge", "corrupt"])
def test_damaged_bam_raises_everywhere(tmp_path, kind):
    import json
    from instrain_b200.packer import BamPacker
    from instrain_b200.profile import iter_batches
    from instrain_b200.read_filter import filter_reads, mapping_info
    path = _damaged_copies(tmp_path)[kind]
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    with BamPacker(path) as bp:
        names = bp.ref_names                                              # the header is intact in all three
    with pytest.raises(IOError):
        list(iter_batches(path, rdic, seqs))
    with pytest.raises(IOError):
        filter_reads(path, names)
    with pytest.raises(IOError):
        mapping_info(path, names)


def test_intact_bam_still_reads(tmp_path):
    """Control for the test above: the undamaged file gives its batch, and a BAM without the end-of-file block is readable
    on request (ISB_ALLOW_NO_BGZF_EOF), as `samtools` merely warns about it."""
    import json
    from instrain_b200.profile import iter_batches
    src = os.path.join(GOLDEN, "c1_G1_subset.bam")
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    full = [b for k, b in iter_batches(src, rdic, seqs) if k == "batch"]
    assert len(full) == 1 and full[0]["n_events"] > 100000
    raw = open(src, "rb").read()
    assert raw[-28:-26] == b"\x1f\x8b" and raw[-4:] == b"\0\0\0\0"         # the 28-byte empty block
    noeof = tmp_path / "noeof.bam"
    noeof.write_bytes(raw[:-28])
    with pytest.raises(IOError):
        list(iter_batches(str(noeof), rdic, seqs))
    os.environ["ISB_ALLOW_NO_BGZF_EOF"] = "1"
    try:
        again = [b for k, b in iter_batches(str(noeof), rdic, seqs) if k == "batch"]
    finally:
        del os.environ["ISB_ALLOW_NO_BGZF_EOF"]
    assert again[0]["n_events"] == full[0]["n_events"]


def test_mm_levels_beyond_the_device_limit_are_folded():
    from instrain_b200 import _cabi
    from instrain_b200.packer import BamPacker
    mm = BamPacker._mm_levels([0, 3, 63, 64, 300])
    assert mm.dtype == np.uint8 and mm.tolist() == [0, 3, 63, _cabi.ISB_MAX_MM - 1, _cabi.ISB_MAX_MM - 1]
    with pytest.raises(ValueError):
        BamPacker._mm_levels([-1])
