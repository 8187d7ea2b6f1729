"""a single large scaffold is sharded by contiguous runs of splits with a read halo.  instrain_b200.reads.
clip_reads cuts a read-major batch to the positions of a run; instrain_b200.shard.split_runs partitions the splits of one
scaffold into runs of about equal weight.  CPU tests: the clipped batch encodes exactly the events of the run, and the
oracle on the runs, put back together, equals the oracle on the whole scaffold -- SNV rows, linkage rows (none crosses a
split), coverage and clonality at every position."""
import numpy as np

from conftest import assert_ld_equal, assert_snv_equal, load_lut
from instrain_b200 import reads
from instrain_b200.shard import split_runs
from oracle import restate, synth


def _run_batch(batch, rd, lo, hi, run_splits):
    sub, origin = reads.clip_reads(rd, lo, hi)
    ev = reads.reads_to_events(sub)
    ev["pair_mm"] = batch["pair_mm"]
    ref = batch["ref_codes"][origin:hi]
    splits = np.asarray(run_splits, dtype=np.int32)
    return sub, origin, ev, ref, splits


def test_split_runs_cover_and_balance():
    splits = synth.batch_splits(95000, 1, 10000) if hasattr(synth, "batch_splits") else None
    from instrain_b200.synth import iterate_splits
    sp = iterate_splits(95000, 10000)
    w = np.ones(len(sp))
    runs = split_runs(sp, w, 3)
    assert runs[0][0] == 0 and runs[-1][1] == len(sp) and all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
    sizes = [b - a for a, b in runs]
    assert max(sizes) - min(sizes) <= 1 and len(runs) == 3
    assert split_runs(sp[:2], np.ones(2), 8) == [(0, 1), (1, 2)]          # never more runs than splits


def test_clipped_runs_equal_the_whole_scaffold():
    lut, dflt = load_lut()
    batch = synth.make_batch(42000, 60, 0.03, 20260501, skip_mm=False, n_frac=0.002)
    rd = reads.events_to_reads(batch)
    whole = restate.profile_events(batch, batch["ref_codes"], lut, dflt, batch["splits"])
    sp = [tuple(int(v) for v in r) for r in batch["splits"]]
    assert len(sp) >= 4
    runs = split_runs(sp, np.ones(len(sp)), 3)
    snv, ld = [], []
    covT = np.zeros_like(whole["covT"])
    clonT = np.full_like(whole["clonT"], np.nan)
    M = whole["covT"].shape[1]
    for a, b in runs:
        lo, hi = sp[a][0], sp[b - 1][1] + 1
        sub, origin, ev, ref, splits = _run_batch(batch, rd, lo, hi, sp[a:b])
        # the clipped batch holds exactly the passing events of the run
        full_ev = reads.reads_to_events(rd)
        sel = (full_ev["ref_pos"] >= lo) & (full_ev["ref_pos"] < hi)
        for k in ("ref_pos", "base", "read_id"):
            assert np.array_equal(ev[k], full_ev[k][sel]), k
        assert sub["seg_start"].min() >= lo and (sub["seg_start"].astype(np.int64) + sub["seg_len"]).max() <= hi
        got = restate.profile_events(ev, ref, lut, dflt, splits, start=origin, M=M)     # rows come back in whole-batch coordinates
        snv.append(got["snv"]); ld.append(got["ld"])
        covT[lo:hi] = got["covT"][lo - origin:]
        clonT[lo:hi] = got["clonT"][lo - origin:]
        assert not got["covT"][:lo - origin].any()                          # the alignment pad in front of the run: no events
    assert_snv_equal(np.concatenate(snv), whole["snv"])
    assert_ld_equal(np.concatenate(ld), whole["ld"], tol=0.0)
    assert np.array_equal(covT, whole["covT"])
    assert np.array_equal(np.isnan(clonT), np.isnan(whole["clonT"]))
    ok = ~np.isnan(clonT)
    assert np.array_equal(clonT[ok].view(np.uint32), whole["clonT"][ok].view(np.uint32))


def test_clip_reads_property_random_ranges():
    """clip_reads on ragged segments (short pieces, odd block sizes, non-ACGT events) and arbitrary, unaligned ranges --
    empty ones included: the clipped batch obeys the layout rules and encodes exactly the events of the range."""
    rng = np.random.default_rng(20260503)
    batch = synth.make_batch(9000, 25, 0.05, 77, n_scaffolds=2, skip_mm=True, n_frac=0.004)
    rd = reads.events_to_reads(batch, max_len=37, odd_blocks=True)
    full = reads.reads_to_events(rd)
    L = len(batch["ref_codes"])
    ranges = [(0, L), (0, 1), (L - 1, L), (4321, 4322), (17, 17 + 8)] + [tuple(sorted(rng.integers(0, L, 2))) for _ in range(40)]
    for lo, hi in ranges:
        lo, hi = int(lo), int(hi)
        if hi <= lo:
            hi = lo + 1
        sub, origin = reads.clip_reads(rd, lo, hi)
        assert origin == lo & ~7
        ev = reads.reads_to_events(sub)
        sel = (full["ref_pos"] >= lo) & (full["ref_pos"] < hi)
        for k in ("ref_pos", "base", "read_id"):
            assert np.array_equal(ev[k], full[k][sel]), (k, lo, hi)
        s, n, w = sub["seg_start"].astype(np.int64), sub["seg_len"].astype(np.int64), sub["seg_word"]
        assert (np.diff(s) >= 0).all() and (n >= 1).all() and (s >= lo).all() and (s + n <= hi).all()
        nw = ((s & 7) + n + 7) // 8
        assert (w[1:] == w[:-1] + nw[:-1] + 1).all() and (len(w) == 0 or w[0] == 1)
        assert sub["n_words"] % 4 == 0 and (len(w) == 0 or w[-1] + nw[-1] + 1 <= sub["n_words"])
        sep = np.ones(sub["n_words"], dtype=bool)                      # every word outside the segments' data words is zero
        for a, b in zip(w, nw):
            sep[a:a + b] = False
        assert not sub["words"][sep].any()


# ---- on the device --------------------------------------------------------------------------------------------------------
import pytest


@pytest.fixture(scope="module")
def eng(null_lut):
    from instrain_b200.engine import Engine
    e = Engine(0, null_lut[0], null_lut[1])
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_runs_equal_the_whole_scaffold(eng, null_lut):
    """The CUDA path on the clipped runs (start = the run's origin) against the oracle on the whole scaffold."""
    batch = synth.make_batch(52000, 70, 0.03, 20260502, skip_mm=True)
    rd = reads.events_to_reads(batch)
    whole = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"])
    sp = [tuple(int(v) for v in r) for r in batch["splits"]]
    snv, ld = [], []
    covT = np.zeros_like(whole["covT"])
    for a, b in split_runs(sp, np.ones(len(sp)), 4):
        lo, hi = sp[a][0], sp[b - 1][1] + 1
        sub, origin = reads.clip_reads(rd, lo, hi)
        got = eng.profile_batch(batch, batch["ref_codes"][origin:hi], np.asarray(sp[a:b], np.int32), start=origin, M=1, reads=sub,
                                want=("covT", "clonT", "clonTR", "snv", "ld"))
        snv.append(got["snv"]); ld.append(got["ld"])
        covT[lo:hi] = got["covT"][lo - origin:]
    assert_snv_equal(np.concatenate(snv), whole["snv"])
    assert_ld_equal(np.concatenate(ld), whole["ld"], tol=1e-9)
    assert np.array_equal(covT, whole["covT"])


@pytest.mark.gpu
def test_gpu_bench_run_clipping_equals_the_whole_scaffold(eng):
    """bench.py's device-side twin (clip_dataset_to_run) on the generator's data: the runs of one scaffold, profiled one by
    one, give the rows of the whole scaffold."""
    import os
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from instrain_b200 import synth as dsynth
    args = types.SimpleNamespace(L=80000, cov=60, dens=0.03, mm=False, seg_words=None)
    d = bench.build_dataset(dsynth, 0, [0], args)
    npf = lambda t: t.cpu().numpy()
    ev = dict(pair_mm=npf(d["pair_mm"]))
    want = ("covT", "snv", "ld")
    whole = eng.profile_batch(ev, npf(d["ref_codes"]), npf(d["splits"]), M=1, reads=d["reads"], want=want)
    assert len(whole["snv"]) > 500 and len(whole["ld"]) > 500
    sp = [tuple(int(v) for v in r) for r in npf(d["splits"])]
    snv, ld = [], []
    covT = np.zeros_like(whole["covT"])
    for a, b in split_runs(sp, np.ones(len(sp)), 3):
        lo, hi = sp[a][0], sp[b - 1][1] + 1
        dd = bench.clip_dataset_to_run(d, lo, hi)
        assert dd["start"] == lo & ~7 and dd["L_run"] == hi - dd["start"] and len(dd["splits"]) == b - a
        got = eng.profile_batch(ev, npf(dd["ref_codes"]), npf(dd["splits"]), start=dd["start"], M=1, reads=dd["reads"], want=want)
        snv.append(got["snv"]); ld.append(got["ld"])
        covT[lo:hi] = got["covT"][lo - dd["start"]:]
    assert_snv_equal(np.concatenate(snv), whole["snv"])
    assert_ld_equal(np.concatenate(ld), whole["ld"], tol=0.0)
    assert np.array_equal(covT, whole["covT"])


def test_region_packing_equals_clipping_the_whole_scaffold(tmp_path):
    """The host side of a run: BamPacker.pack_scaffold_reads(region=...) after a seek through the .bai LINEAR index reads and
    packs only the reads that overlap the run; cut to the run it encodes the same events as the whole scaffold's pack cut to
    the run (the mate-overlap tweak only needs the reads a region fetch returns: tests/test_polymorpher_region.py)."""
    from instrain_b200 import synth_bam
    from instrain_b200.packer import BamPacker, find_bai, read_bai, read_bai_linear, seek_offset
    from instrain_b200.read_filter import filter_reads
    from instrain_b200.synth import iterate_splits
    bam = str(tmp_path / "r.bam")
    info = synth_bam.write_bam(bam, 70000, 2, 25, 0.02, seed=31)
    r2m, _, _ = filter_reads(bam, info["names"])
    first, linear = read_bai(find_bai(bam)), read_bai_linear(find_bai(bam))
    assert all(len(l) == 5 and l.all() for l in linear)                    # 70 kb = 5 windows of 16 kb, all covered
    for tid, name in enumerate(info["names"]):
        with BamPacker(bam) as bp:
            bp.seek(first[tid])
            whole = reads.concat_streams([bp.pack_scaffold_reads(tid, r2m[name])])
        sp = iterate_splits(70000, 10000)
        seen = 0
        for a, b in split_runs(sp, [e - s + 1 for s, e in sp], 3):
            lo, hi = sp[a][0], sp[b - 1][1] + 1
            with BamPacker(bam) as bp:
                bp.seek(seek_offset(linear[tid], first[tid], lo))
                part = bp.pack_scaffold_reads(tid, r2m[name], region=(lo, hi))
            assert part["reads_seen"] < 0.6 * (info["n_reads"] // 2)        # a third of the scaffold (+ halo), not all of it
            seen += part["reads_seen"]
            got = reads.reads_to_events(reads.clip_reads(reads.concat_streams([part]), lo, hi)[0])
            exp = reads.reads_to_events(reads.clip_reads(whole, lo, hi)[0])
            assert len(exp["ref_pos"]) > 100000
            # pair ids are numbered per pack (order inside a position differs): compare (position, base, mm of the pair) sets
            wmm = whole_pair_mm(bam, tid, r2m[name])
            ga = np.stack([got["ref_pos"], got["base"], part["pair_mm"][got["read_id"]]], axis=1)
            ea = np.stack([exp["ref_pos"], exp["base"], wmm[exp["read_id"]]], axis=1)
            ga, ea = ga[np.lexsort(ga.T[::-1])], ea[np.lexsort(ea.T[::-1])]
            assert np.array_equal(ga, ea), (name, lo)
            # and the pairing itself: events of one pair in the region pack are events of one pair in the whole pack
            gk = np.unique(np.stack([got["read_id"], got["ref_pos"]], axis=1), axis=0)
            ek = np.unique(np.stack([exp["read_id"], exp["ref_pos"]], axis=1), axis=0)
            assert len(gk) == len(ek) and len(np.unique(gk[:, 0])) == len(np.unique(ek[:, 0]))
        assert seen < 1.5 * (info["n_reads"] // 2)                         # the seek lands on a 16 kb window start: some reads read twice


def whole_pair_mm(bam, tid, r2m):
    from instrain_b200.packer import BamPacker, find_bai, read_bai
    with BamPacker(bam) as bp:
        bp.seek(read_bai(find_bai(bam))[tid])
        return bp.pack_scaffold_reads(tid, r2m)["pair_mm"]
