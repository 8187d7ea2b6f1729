"""CPU mirrors of reference tests that have an INDEPENDENT known answer for the hot path's outputs (they need the
reference's test data, so they run in the build container only).

test_profile_3 (test/tests/test_profile.py:347-378): `inStrain profile ... -l 0.98`, then
  * _internal_verify_Sdb     (test_utils.py:265-298): per scaffold, breadth_minCov / coverage / coverage_median never
                             decrease with mm; conANI_reference != 0 wherever there are consensus-divergent sites;
  * coverage and breadth at the highest mm against the output of an independent tool (calculate_breadth, stored as
    `...bam.CB`), one-sided within 0.1 / 0.01 (the profile only sees the reads that pass the filter);
  * _internal_verify_OdbSdb  (test_utils.py:300-317): divergent_site_count == number of SNV-table positions, at the
                             lowest and at the highest mm.
Here the path is: C++ read filter (min_read_ani 0.98) -> C++ packer -> oracle (C restatement) -> summary restatement;
the CUDA path equals the oracle on the same events (tests/test_gpu_*.py)."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import load_lut
from oracle import bamio, ref_harness, restate, summary

TD = "/root/reference/test/test_data/"
BAM = TD + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.sorted.bam"
pytestmark = pytest.mark.skipif(not os.path.exists(BAM), reason="reference test data not present")


def test_profile_3_mirror():
    from instrain_b200.packer import BamPacker
    from instrain_b200.read_filter import filter_reads
    seqs = bamio.read_fasta(TD + "N5_271_010G1_scaffold_min1000.fa")
    with BamPacker(BAM) as bp:
        names = bp.ref_names
    r2m, _, _ = filter_reads(BAM, names, min_read_ani=0.98)
    r2m_095, _, _ = filter_reads(BAM, names)
    assert 0 < sum(len(v) for v in r2m.values()) < sum(len(v) for v in r2m_095.values())     # -l 0.98 drops pairs
    lut, dflt = load_lut()
    odb, sdb = [], []
    with BamPacker(BAM) as bp:
        while True:
            tid = bp.peek_tid()
            if tid < 0:
                break
            name = bp.ref_names[tid]
            ev = bp.pack_scaffold(tid, r2m.get(name, {}))
            if not r2m.get(name):
                continue
            L = len(seqs[name])
            out = restate.profile_events(ev, restate.encode_ref(seqs[name]), lut, dflt, np.array([[0, L - 1]], np.int32),
                                         do_linkage=False)
            for r in summary.scaffold_summary(out["covT"], out["clonT"], out["nmask"], out["snv"], 0):
                odb.append(dict(r, scaffold=name))
            for row in out["snv"]:
                sdb.append(dict(scaffold=name, position=int(row["pos"]), mm=int(row["mm"])))
    Odb, Sdb = pd.DataFrame(odb), pd.DataFrame(sdb)
    assert Odb["scaffold"].nunique() > 150 and len(Sdb) > 300
    # _internal_verify_Sdb
    for scaff, d in Odb.groupby("scaffold"):
        d = d.sort_values("mm")
        for thing in ("breadth_minCov", "coverage", "coverage_median"):
            assert d[thing].tolist() == sorted(d[thing].tolist()), (scaff, thing)
    assert (Odb["conANI_reference"][Odb["consensus_divergent_sites"] > 0] != 0).all()
    # against calculate_breadth
    Cdb = pd.read_csv(TD + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.bam.CB")
    s2c, s2b = Cdb.set_index("scaffold")["coverage"].to_dict(), Cdb.set_index("scaffold")["breadth"].to_dict()
    for scaff, db in Odb.groupby("scaffold"):
        top = db.sort_values("mm", ascending=False).iloc[0]
        assert top["coverage"] - s2c[scaff] < 0.1, (scaff, top["coverage"], s2c[scaff])
        assert top["breadth"] - s2b[scaff] < 0.01, (scaff, top["breadth"], s2b[scaff])
    # _internal_verify_OdbSdb
    low_mm = Sdb["mm"].min()
    for scaff, db in Sdb[Sdb["mm"] == low_mm].groupby("scaffold"):
        snps = Odb["divergent_site_count"][(Odb["scaffold"] == scaff) & (Odb["mm"] == low_mm)].fillna(0).tolist()[0]
        assert snps == len(db), (scaff, snps, len(db))
    top = Odb.sort_values("mm").drop_duplicates(subset="scaffold", keep="last")
    for scaff, db in Sdb.sort_values("mm").drop_duplicates(subset=["scaffold", "position"], keep="last").groupby("scaffold"):
        assert top["divergent_site_count"][top["scaffold"] == scaff].fillna(0).tolist()[0] == len(db), scaff


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_mapping_info_reproduces_reference_table(which):
    """instrain_b200.read_filter.mapping_info (C++ filter) against the reference's stored mapping_info of both samples:
    all 19 columns of all 179 rows (per-scaffold tallies, means, median insert, and the weighted all_scaffolds row)."""
    from instrain_b200.packer import BamPacker
    from instrain_b200.read_filter import MAPPING_INFO_COLUMNS, mapping_info
    bam = TD + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010%s.sorted.bam" % which
    ref = pd.read_csv(TD + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010%s.forRC.IS/raw_data/mapping_info.csv.gz" % which,
                      index_col=0)
    with BamPacker(bam) as bp:
        names = bp.ref_names
    got = mapping_info(bam, names)
    assert list(got.columns) == list(ref.columns) == MAPPING_INFO_COLUMNS and len(got) == len(ref) == 179
    assert got["scaffold"][0] == ref["scaffold"][0] == "all_scaffolds" and set(got["scaffold"]) == set(ref["scaffold"])
    a, b = got.set_index("scaffold").loc[ref["scaffold"]], ref.set_index("scaffold")
    for c in MAPPING_INFO_COLUMNS[1:]:
        assert np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=0, atol=1e-9, equal_nan=True), c


def _ref_filter(s2p2i, pairing_filter, priority, **thr):
    """The reference's OWN paired_read_filter + filter_scaff2pair2info (imported from /root/reference) on pair infos."""
    from oracle import ref_harness
    ref_harness.load_reference()
    import inStrain.filter_reads as fr
    as_ref = {s: {p: np.array(list(i) + [0, 0], dtype="int64") for p, i in d.items()} for s, d in s2p2i.items()}
    tallys = {}
    filt = fr.paired_read_filter(as_ref, priority_reads_set=set(priority), tallys=tallys, pairing_filter=pairing_filter)
    r2m, Rdb = fr.filter_scaff2pair2info(filt, tallys, priority_reads_set=set(priority), pairing_filter=pairing_filter, **thr)
    return {s: {p: int(v) for p, v in d.items()} for s, d in r2m.items()}, Rdb


@pytest.mark.parametrize("pairing_filter,use_priority", [("paired_only", True), ("non_discordant", False),
                                                         ("non_discordant", True), ("all_reads", False)])
def test_read_filter_modes_against_reference_functions(pairing_filter, use_priority):
    """The non-default pairing filters and priority reads: restatement (oracle/read_filter.py) and the C++ filter against
    the reference's own paired_read_filter / filter_scaff2pair2info run on the same pair infos -- sR2M and every tally."""
    from oracle import read_filter as orf
    from instrain_b200.packer import BamPacker
    from instrain_b200.read_filter import filter_reads
    refs, reads = bamio.read_bam(BAM)
    by = {}
    for r in reads:
        if r.tid >= 0:
            by.setdefault(refs[r.tid][0], []).append(r)
    s2i = {s: orf.pair2info(rs) for s, rs in by.items()}
    singles = [p for d in s2i.values() for p, i in d.items() if i[4] == 1]
    priority = singles[::7] if use_priority else []
    thr = dict(min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50)
    exp, Rdb = _ref_filter(s2i, pairing_filter, priority, **thr)
    got, tal, _ = orf.filter_pairs(s2i, pairing_filter=pairing_filter, priority_reads=priority, **thr)
    assert got == exp and sum(len(v) for v in exp.values()) > 7000
    R = Rdb[Rdb["scaffold"] != "all_scaffolds"].set_index("scaffold")
    for s, t in tal.items():
        for c, v in t.items():
            assert int(R.loc[s, c]) == int(v), (s, c)
    with BamPacker(BAM) as bp:
        names = bp.ref_names
    cpp, cpp_tal, _ = filter_reads(BAM, names, pairing_filter=pairing_filter, priority_reads=priority, **thr)
    assert cpp == {s: d for s, d in exp.items() if d}
    for s, t in cpp_tal.items():
        for c, v in t.items():
            assert int(R.loc[s, c]) == int(v), (s, c)
    # the whole read report (mapping_info) of this mode: every column of every row, the weighted all_scaffolds row included
    from instrain_b200.read_filter import MAPPING_INFO_COLUMNS, mapping_info
    mi = mapping_info(BAM, names, pairing_filter=pairing_filter, priority_reads=priority, **thr).set_index("scaffold")
    Rall = Rdb.set_index("scaffold")
    assert sorted(mi.index) == sorted(Rall.index) and "all_scaffolds" in mi.index
    for c in MAPPING_INFO_COLUMNS[1:]:
        a, b = mi.loc[Rall.index, c].values.astype(float), Rall[c].values.astype(float)
        assert np.allclose(a, b, rtol=0, atol=1e-9, equal_nan=True), c


def test_hot_path_with_singletons_in_r2m_against_reference_functions():
    """pairing_filter='all_reads' puts singletons and pairs split over two scaffolds into R2M.  The hot path on such an
    R2M: the oracle (and with it the CUDA path) against the reference's own process_bam_sites / calculate_ld fed with the
    same emulated columns, on the 8 best-covered scaffolds -- SNV rows, linkage rows, covT."""
    from conftest import assert_ld_equal, assert_snv_equal
    from oracle import pileup_emul, read_filter as orf, ref_harness
    from test_oracle_golden import expected_rows
    refs, reads = bamio.read_bam(BAM)
    seqs = bamio.read_fasta(TD + "N5_271_010G1_scaffold_min1000.fa")
    by = {}
    for r in reads:
        if r.tid >= 0:
            by.setdefault(refs[r.tid][0], []).append(r)
    r2m_all, _, _ = orf.filter_pairs({s: orf.pair2info(rs) for s, rs in by.items()}, pairing_filter="all_reads")
    r2m_po, _, _ = orf.filter_pairs({s: orf.pair2info(rs) for s, rs in by.items()})
    lut, dflt = load_lut()
    model = ref_harness.null_model(1e-6)
    from instrain_b200.packer import BamPacker, read_bai
    first = read_bai(BAM + ".bai")
    B = {b: i for i, b in enumerate("ACTG")}
    CLS = {n: i for i, n in enumerate(restate.CLASS_NAMES)}
    n_single = 0
    for name in sorted(r2m_all, key=lambda s: -len(r2m_all[s]))[:8]:
        r2m = r2m_all[name]
        n_single += len(set(r2m) - set(r2m_po.get(name, {})))
        ev = pileup_emul.scaffold_events(by[name], r2m)
        seq = seqs[name]
        out = ref_harness.run_split(ev, seq, 0, len(seq) - 1, r2m, model, scaffold=name)
        z = dict(
            snv_pos=np.array([r["position"] for r in out["snp"]], np.int32), snv_mm=np.array([r["mm"] for r in out["snp"]], np.int32),
            snv_cnt=np.array([[r["A"], r["C"], r["T"], r["G"]] for r in out["snp"]], np.int32).reshape(-1, 4),
            snv_con=np.array([B[r["con_base"]] for r in out["snp"]], np.uint8), snv_var=np.array([B[r["var_base"]] for r in out["snp"]], np.uint8),
            snv_allele_count=np.array([r["allele_count"] for r in out["snp"]], np.uint8),
            snv_cls=np.array([CLS[r["class"]] for r in out["snp"]], np.uint8), snv_cryptic=np.array([r["cryptic"] for r in out["snp"]], np.uint8),
            ld_pos_a=np.array([r["position_A"] for r in out["ld"]], np.int32), ld_pos_b=np.array([r["position_B"] for r in out["ld"]], np.int32),
            ld_mm=np.array([r["mm"] for r in out["ld"]], np.int32),
            ld_counts=np.array([[r["countAB"], r["countAb"], r["countaB"], r["countab"]] for r in out["ld"]], np.int32).reshape(-1, 4),
            ld_alleles=np.array([[B[r[c]] for c in ("allele_A", "allele_a", "allele_B", "allele_b")] for r in out["ld"]], np.uint8).reshape(-1, 4),
            ld_r2=np.array([r["r2"] for r in out["ld"]], np.float64), ld_d_prime=np.array([r["d_prime"] for r in out["ld"]], np.float64))
        sev = restate.sort_events(ev)
        ref_codes = restate.encode_ref(seq)
        got = restate.profile_events(sev, ref_codes, lut, dflt, np.array([[0, len(seq) - 1]], np.int32))
        snv, ld = expected_rows(z, ref_codes)
        assert_snv_equal(got["snv"], snv)
        assert_ld_equal(got["ld"], ld, tol=1e-9)
        for mm, arr in out["covT"].items():
            assert np.array_equal(got["covT"][:, mm], arr), (name, mm)
        with BamPacker(BAM) as bp:                               # the C++ packer on the same R2M: the emulation's events
            tid = bp.ref_names.index(name)
            bp.seek(first[tid])
            pk = bp.pack_scaffold(tid, r2m)
        for k in ("ref_pos", "base", "qual", "read_id"):
            assert np.array_equal(pk[k], sev[k]), (name, k)
        assert np.array_equal(pk["pair_mm"], sev["pair_mm"].astype(np.uint8))
    assert n_single > 100                                        # the case is exercised: names beyond the paired_only set


@pytest.mark.parametrize("seed,cov,dens,n_frac,skip_mm", [(1, 25, 0.03, 0.0, False), (2, 60, 0.05, 0.004, False),
                                                           (3, 12, 0.02, 0.0, True), (4, 90, 0.08, 0.002, False),
                                                           (5, 40, 0.04, 0.01, True)])
def test_oracle_equals_reference_functions_on_random_inputs(seed, cov, dens, n_frac, skip_mm):
    """Beyond the bundled BAMs: random synthetic read sets (haplotype mixtures, sequencing errors, low-quality bases,
    non-ACGT read bases, several mm levels or set mode) through the reference's own process_bam_sites / calculate_ld and
    through the oracle -- SNV rows, linkage rows, covT, clonT."""
    from conftest import assert_ld_equal, assert_snv_equal
    from oracle import ref_harness, synth
    from test_oracle_golden import expected_rows
    L = 900
    batch = synth.make_batch(L, cov, dens, 4000 + seed, skip_mm=skip_mm, n_frac=n_frac)
    n_pairs = len(batch["pair_mm"])
    names = ["r%d" % i for i in range(n_pairs)]
    ev = dict(batch, names=names)
    seq = "".join("ACTGN"[c] for c in batch["ref_codes"])
    r2m = set(names) if skip_mm else {n: int(m) for n, m in zip(names, batch["pair_mm"])}
    model = ref_harness.null_model(1e-6)
    lut, dflt = load_lut()
    (s0, e0), = [tuple(x) for x in batch["splits"]] if len(batch["splits"]) == 1 else [(0, L - 1)]
    out = ref_harness.run_split(ev, seq, int(s0), int(e0), r2m, model, min_cov=5, min_freq=0.05, min_snp=10)
    got = restate.profile_events(batch, batch["ref_codes"], lut, dflt, batch["splits"], min_snp=10)
    B = {b: i for i, b in enumerate("ACTG")}
    CLS = {n: i for i, n in enumerate(restate.CLASS_NAMES)}
    z = dict(
        snv_pos=np.array([r["position"] for r in out["snp"]], np.int32), snv_mm=np.array([r["mm"] for r in out["snp"]], np.int32),
        snv_cnt=np.array([[r["A"], r["C"], r["T"], r["G"]] for r in out["snp"]], np.int32).reshape(-1, 4),
        snv_con=np.array([B[r["con_base"]] for r in out["snp"]], np.uint8), snv_var=np.array([B[r["var_base"]] for r in out["snp"]], np.uint8),
        snv_allele_count=np.array([r["allele_count"] for r in out["snp"]], np.uint8),
        snv_cls=np.array([CLS[r["class"]] for r in out["snp"]], np.uint8), snv_cryptic=np.array([r["cryptic"] for r in out["snp"]], np.uint8),
        ld_pos_a=np.array([r["position_A"] for r in out["ld"]], np.int32), ld_pos_b=np.array([r["position_B"] for r in out["ld"]], np.int32),
        ld_mm=np.array([r["mm"] for r in out["ld"]], np.int32),
        ld_counts=np.array([[r["countAB"], r["countAb"], r["countaB"], r["countab"]] for r in out["ld"]], np.int32).reshape(-1, 4),
        ld_alleles=np.array([[B[r[c]] for c in ("allele_A", "allele_a", "allele_B", "allele_b")] for r in out["ld"]], np.uint8).reshape(-1, 4),
        ld_r2=np.array([r["r2"] for r in out["ld"]], np.float64), ld_d_prime=np.array([r["d_prime"] for r in out["ld"]], np.float64))
    snv, ld = expected_rows(z, batch["ref_codes"])
    assert len(snv) > 5
    assert_snv_equal(got["snv"], snv)
    assert_ld_equal(got["ld"], ld, tol=1e-9)
    for mm, arr in out["covT"].items():
        assert np.array_equal(got["covT"][:, mm], arr), mm
    for mm, arr in out["clonT"].items():
        mine = got["clonT"][:, mm]
        assert np.array_equal(np.isnan(mine), np.isnan(arr)) and np.array_equal(mine[~np.isnan(arr)], arr[~np.isnan(arr)].astype(np.float32)), mm


# ---- the re-drawn outputs: same DISTRIBUTION as the reference's own (unseeded) functions ----------------------------------

def _mean_ci_overlap(a, b, z=5.0):
    """|mean(a) - mean(b)| within z standard errors of the difference."""
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    return abs(a.mean() - b.mean()) <= z * se + 1e-12


@pytest.mark.parametrize("counts,n", [((96, 3, 1, 0), 50), ((40, 35, 20, 5), 50), ((0, 7, 0, 3), 10), ((500, 0, 12, 0), 50),
                                      ((25, 25, 25, 25), 30)])
def test_rarefied_clonality_has_the_reference_distribution(counts, n):
    """calculate_rarefied_clonality (snv_utilities.py:233-247) draws with an unseeded np.random.choice; the restatement
    of the CUDA path's counter-based draws (oracle/rarefied.py, bit-exact with the kernels: tests/test_gpu_*.py) must have
    the same distribution: mean and variance over 20000 sites of the same counts, and the same support."""
    from oracle import rarefied
    _, su, _ = ref_harness.load_reference()
    N = 20000
    rng_state = np.random.get_state()
    np.random.seed(12345)
    try:
        ref = np.array([su.calculate_rarefied_clonality(list(counts), rarefied_coverage=n) for _ in range(N)])
    finally:
        np.random.set_state(rng_state)
    c = np.zeros((N, 1, 4), np.int32)
    c[:, 0, :] = counts
    own = rarefied.clonTR(c, np.zeros(N, np.uint64), rarefied_coverage=n, seed=99)[:, 0].astype(np.float64)
    assert not np.isnan(own).any()
    assert _mean_ci_overlap(own, ref), (own.mean(), ref.mean())
    assert _mean_ci_overlap((own - own.mean()) ** 2, (ref - ref.mean()) ** 2), (own.var(), ref.var())
    support = lambda x: set(np.round(x * n * n).astype(np.int64))          # clonality * n^2 = sum of squared counts: integers
    assert support(own) <= support(ref) | support(own) and len(support(own) ^ support(ref)) <= 0.2 * len(support(ref)) + 2
    below = rarefied.clonTR(c, np.zeros(N, np.uint64), rarefied_coverage=sum(counts) + 1, seed=99)
    assert np.isnan(below).all()                                           # set only where the coverage reaches it


@pytest.mark.parametrize("AB,Ab,aB,ab", [(30, 5, 4, 25), (50, 0, 0, 14), (10, 10, 10, 10), (200, 3, 2, 1)])
def test_normalized_linkage_has_the_reference_distribution(AB, Ab, aB, ab):
    """r2_normalized / d_prime_normalized of _calc_ld_single (linkage.py:200-228: min_snp haplotypes re-drawn from the four
    frequencies, unseeded) against the restatement of the CUDA path's draws: NaN rate, mean and variance over 20000 rows."""
    from oracle import rarefied, restate
    _, _, lk = ref_harness.load_reference()
    N, min_snp = 20000, 20
    total = AB + Ab + aB + ab
    rng_state = np.random.get_state()
    np.random.seed(54321)
    try:
        ref = [lk._calc_ld_single(min_snp, A="A", a="C", B="G", b="T", AB=AB, Ab=Ab, aB=aB, ab=ab, total=total) for _ in range(N)]
    finally:
        np.random.set_state(rng_state)
    rows = np.zeros(N, dtype=restate.LD_DT)
    rows["pos_a"] = np.arange(N)
    rows["pos_b"] = np.arange(N) + 7
    rows["c_AB"], rows["c_Ab"], rows["c_aB"], rows["c_ab"] = AB, Ab, aB, ab
    r2n, dpn = rarefied.normalized_ld(rows, min_snp, seed=5)
    for name, own in (("r2_normalized", r2n), ("d_prime_normalized", dpn)):
        rf = np.array([r[name] for r in ref], dtype=np.float64)
        nan_o, nan_r = np.isnan(own).mean(), np.isnan(rf).mean()
        assert abs(nan_o - nan_r) <= 5 * np.sqrt(max(nan_r * (1 - nan_r), 1e-4) * 2 / N) + 1e-3, (name, nan_o, nan_r)
        o, r = own[~np.isnan(own)], rf[~np.isnan(rf)]
        if len(o) > 100 and len(r) > 100:
            assert _mean_ci_overlap(o, r), (name, o.mean(), r.mean())
            assert _mean_ci_overlap((o - o.mean()) ** 2, (r - r.mean()) ** 2), (name, o.var(), r.var())
