"""profile_bam shim end to end on a real BAM: host packer -> K1/K2/K3 -> reference-shaped tables, against the reference's
stored raw_snp_table / raw_linkage_table rows of the same scaffolds (golden fixtures)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, load_batch

pytestmark = pytest.mark.gpu


def golden_tables(names):
    from instrain_b200._cabi import CLASS_NAMES
    batch, exp = load_batch("G1")
    sn = list(batch["scaffold_names"])
    off = batch["scaffold_off"].astype(np.int64)
    sidx = np.searchsorted(off, exp["snv_pos"], side="right") - 1
    snv = pd.DataFrame({"scaffold": np.array(sn, dtype=object)[sidx], "position": exp["snv_pos"] - off[sidx],
                        "mm": exp["snv_mm"], "A": exp["snv_cnt"][:, 0], "C": exp["snv_cnt"][:, 1],
                        "T": exp["snv_cnt"][:, 2], "G": exp["snv_cnt"][:, 3],
                        "con_base": np.array(list("ACTG"))[exp["snv_con"]], "var_base": np.array(list("ACTG"))[exp["snv_var"]],
                        "allele_count": exp["snv_allele_count"], "class": np.array(CLASS_NAMES, dtype=object)[exp["snv_cls"]],
                        "cryptic": exp["snv_cryptic"].astype(bool)})
    lidx = np.searchsorted(off, exp["ld_pos_a"], side="right") - 1
    ld = pd.DataFrame({"scaffold": np.array(sn, dtype=object)[lidx], "position_A": exp["ld_pos_a"] - off[lidx],
                       "position_B": exp["ld_pos_b"] - off[lidx], "mm": exp["ld_mm"], "countAB": exp["ld_counts"][:, 0],
                       "countAb": exp["ld_counts"][:, 1], "countaB": exp["ld_counts"][:, 2], "countab": exp["ld_counts"][:, 3],
                       "r2": exp["ld_r2"], "d_prime": exp["ld_d_prime"]})
    return snv[snv["scaffold"].isin(names)], ld[ld["scaffold"].isin(names)]


def test_profile_bam_matches_reference_goldens(tmp_path):
    from instrain_b200.profile import profile_bam
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    isp = str(tmp_path / "subset.IS")
    res = profile_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"), None, rdic, isp, s2s=seqs,
                      min_cov=5, min_freq=0.05, min_snp=20, window_length=10000)
    res = getattr(res, "result", res)          # the on-disk object carries the in-memory tables as .result
    assert sorted(res.scaffold_list) == sorted(rdic)
    g_snv, g_ld = golden_tables(set(rdic))
    key = ["scaffold", "position", "mm"]
    a = res.raw_snp_table.sort_values(key).reset_index(drop=True)
    b = g_snv.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) and len(a) > 1000
    for c in ["scaffold", "position", "mm", "A", "C", "T", "G", "con_base", "var_base", "allele_count", "class", "cryptic"]:
        assert (a[c].values == b[c].values).all(), c
    assert (a["position_coverage"] == a[["A", "C", "T", "G"]].sum(axis=1)).all()
    for s in seqs:                                             # ref_base column comes from the FASTA
        m = a["scaffold"] == s
        assert all(seqs[s][p] == r for p, r in zip(a["position"][m], a["ref_base"][m]))
    key = ["scaffold", "position_A", "position_B", "mm"]
    a = res.raw_linkage_table.sort_values(key).reset_index(drop=True)
    b = g_ld.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) and len(a) > 1000
    for c in key + ["countAB", "countAb", "countaB", "countab"]:
        assert (a[c].values == b[c].values).all(), c
    assert (a["distance"] == a["position_B"] - a["position_A"]).all()
    for c in ("r2", "d_prime"):
        assert np.allclose(a[c].values, b[c].values, rtol=0, atol=1e-6, equal_nan=True), c
    # per-scaffold basewise tables: test_profile_13 -- position_coverage == sum of covT[mm' <= mm][pos]
    for s, sp in res.scaffolds.items():
        t = sp.raw_snp_table
        for pos, mm, cov in zip(t["position"], t["mm"], t["position_coverage"]):
            tot = sum(int(sp.covT[m].get(pos, 0)) for m in sp.covT if m <= mm)
            assert tot == cov
        break
    # the SNVprofile directory profile_bam left at ISP_loc: tables and covT / clonT read back equal the in-memory result,
    # and the covT / clonT datasets equal the reference's stored ones for these scaffolds (digests of its .hd5 files)
    from conftest import basewise_digest
    from instrain_b200.store import SNVprofileStore
    S = SNVprofileStore(isp)
    assert S.get("object_type") == "profile" and S.get("scaffold_list") == res.scaffold_list
    back = S.get("raw_snp_table")
    assert len(back) == len(res.raw_snp_table) and list(back.columns[-3:]) == ["var_freq", "con_freq", "ref_freq"]
    assert len(S.get("raw_linkage_table")) == len(res.raw_linkage_table)
    assert len(S.get("cumulative_scaffold_table")) == len(res.cumulative_scaffold_table)
    z = np.load(os.path.join(GOLDEN, "c1_G1_hd5_digest.npz"))
    gold = {str(k): (bytes(a), bytes(b)) for k, a, b in zip(z["names"], z["cov_sha"], z["clon_sha"])}
    covT, clonT = S.get("covT"), S.get("clonT")
    n = 0
    for s in rdic:
        want = {int(k.rsplit("::", 1)[1]) for k in gold if k.rsplit("::", 1)[0] == s}
        assert set(covT[s]) == set(clonT[s]) == want, s
        for mm in want:
            assert bytes(basewise_digest(covT[s][mm].values, covT[s][mm].index.values)) == gold["%s::%d" % (s, mm)][0]
            assert bytes(basewise_digest(clonT[s][mm].values, clonT[s][mm].index.values)) == gold["%s::%d" % (s, mm)][1]
            n += 1
    assert n > 50


def test_profile_bam_tiny_scaffold_vs_reference_functions(tmp_path):
    """The reference's test_profile_18 input (one 126 bp scaffold).  Expected tables were produced by the reference's OWN
    process_bam_sites / calculate_ld (oracle/ref_harness.py) on the emulated pileup -- tests/golden/make_golden.py."""
    from instrain_b200.profile import profile_bam
    fx = json.load(open(os.path.join(GOLDEN, "small_scaffold.json")))
    name = fx["scaffold"]
    res = profile_bam(os.path.join(GOLDEN, "small_scaffold.bam"), None, {name: fx["r2m"]}, str(tmp_path / "tiny.IS"),
                      s2s={name: fx["seq"]}, min_cov=5, min_freq=0.05, min_snp=20)
    res = getattr(res, "result", res)          # the on-disk object carries the in-memory tables as .result
    assert res.scaffold_list == [name] and not res.failures
    exp = pd.DataFrame(fx["snp"]).sort_values(["position", "mm"]).reset_index(drop=True)
    got = res.raw_snp_table.sort_values(["position", "mm"]).reset_index(drop=True)
    assert len(got) == len(exp) == 48
    for c in ["position", "mm", "ref_base", "A", "C", "T", "G", "con_base", "var_base", "allele_count", "class", "cryptic"]:
        assert (got[c].values == exp[c].values).all(), c
    key = ["position_A", "position_B", "mm"]
    exp = pd.DataFrame(fx["ld"]).sort_values(key).reset_index(drop=True)
    got = res.raw_linkage_table.sort_values(key).reset_index(drop=True)
    assert len(got) == len(exp) == 39
    for c in key + ["countAB", "countAb", "countaB", "countab", "allele_A", "allele_a", "allele_B", "allele_b"]:
        assert (got[c].values == exp[c].values).all(), c
    for c in ("r2", "d_prime"):
        assert np.allclose(got[c].values.astype(float), exp[c].values.astype(float), rtol=0, atol=1e-6, equal_nan=True), c
    sp = res.scaffolds[name]
    for mm, dense in fx["covT"].items():                       # exact-mm coverage arrays of the reference (pre-shrink)
        dense = np.asarray(dense)
        s = sp.covT.get(int(mm))
        mine = np.zeros(len(dense), dtype=np.int64)
        if s is not None:
            mine[s.index.values] = s.values
        assert np.array_equal(mine, dense), mm
    assert len(res.cumulative_scaffold_table) == len([m for m in fx["covT"] if np.asarray(fx["covT"][m]).sum() > 0])


def test_polymorpher_extract_snvs_from_bam(null_lut):
    """extract_SNVS_from_bam (second caller of the pileup primitive) against the oracle's counts summed over
    mm, and against the reference's stored SNV rows where the row's mm is the scaffold's top level."""
    from instrain_b200.polymorpher import extract_SNVS_from_bam
    from oracle import bamio, pileup_emul, restate
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    refs, reads = bamio.read_bam(bam)
    tid = 2
    name = refs[tid][0]
    ev = restate.sort_events(pileup_emul.scaffold_events([r for r in reads if r.tid == tid], rdic[name]))
    L = refs[tid][1]
    counts, _ = restate.pileup_counts(ev, 0, L, int(ev["pair_mm"].max()) + 1)
    total = counts.sum(1)
    g_snv, _ = golden_tables({name})
    positions = sorted(set(int(p) for p in g_snv["position"]))[:200] + [0, L - 1]
    got = extract_SNVS_from_bam(bam, rdic[name], positions, name)
    assert set(got) == set(positions)
    for p in positions:
        assert np.array_equal(got[p], total[p]), p
    top = g_snv[g_snv["mm"] == g_snv.groupby("position")["mm"].transform("max")]
    n_chk = 0
    for _, r in top.iterrows():
        p = int(r["position"])
        if p in got and counts[p, int(r["mm"]) + 1:].sum() == 0:
            assert list(got[p]) == [r["A"], r["C"], r["T"], r["G"]]
            n_chk += 1
    assert n_chk > 20
    assert extract_SNVS_from_bam(bam, rdic[name], [], name) == {}


def test_profile_scaffold_run_equals_whole_scaffold(null_lut):
    """SURVEY 8(e) on the device: the largest scaffold of the bundled BAM profiled by two contiguous runs of its splits
    (profile_scaffold_run: packer -> clip_reads -> CUDA path with start = the run's origin) gives the rows and the
    per-position coverage of profiling it whole."""
    import json
    from instrain_b200.engine import Engine
    from instrain_b200.profile import _SplitTable, profile_scaffold_run, profile_scaffolds
    from instrain_b200.shard import split_runs
    name = "N5_271_010G1_scaffold_963"
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    eng = Engine(0, null_lut[0], null_lut[1])
    try:
        whole = profile_scaffolds(bam, {name: rdic[name]}, seqs, engine=eng, window_length=300)
        sp = _SplitTable(None, 300)(name, len(seqs[name]))
        assert len(sp) >= 2
        snv_pos, ld_keys = [], []
        cov = {}
        for a, b in split_runs(sp, [e - s + 1 for s, e in sp], 2):
            rr = profile_scaffold_run(bam, name, rdic[name], seqs[name], sp[a:b], eng)
            snv_pos += [(int(p), int(m)) for p, m in zip(rr["snv"]["pos"], rr["snv"]["mm"])]
            ld_keys += [(int(x), int(y), int(m), int(c)) for x, y, m, c in zip(rr["ld"]["pos_a"], rr["ld"]["pos_b"], rr["ld"]["mm"], rr["ld"]["c_AB"])]
            for m in range(rr["M"]):
                nz = np.nonzero(rr["covT"][:, m])[0]
                cov.setdefault(m, []).extend(zip((nz + rr["lo"]).tolist(), rr["covT"][nz, m].tolist()))
        t = whole.raw_snp_table
        assert sorted(snv_pos) == sorted(zip(t["position"].astype(int), t["mm"].astype(int))) and len(t) > 100
        l = whole.raw_linkage_table
        assert sorted(ld_keys) == sorted(zip(l["position_A"].astype(int), l["position_B"].astype(int), l["mm"].astype(int),
                                            l["countAB"].astype(int))) and len(l) > 100
        for m, s in whole.scaffolds[name].covT.items():
            assert sorted(cov.get(int(m), [])) == list(zip(s.index.tolist(), s.values.tolist())), m
    finally:
        eng.close()
