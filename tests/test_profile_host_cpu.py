"""profile_bam's HOST side end to end without a GPU: BAM -> C++ packer -> batch stream -> transfer format -> [engine] ->
reference-shaped tables -> SNVprofile directory, with the two engine calls answered by the oracle (a test stub: the CUDA
engine gives the same answers, tests/test_gpu_*.py).  Everything around the kernels -- batching, offsets, formats,
tables, the merge-stage summary table, the on-disk store -- is checked against the reference's stored goldens."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, load_batch, load_lut
from instrain_b200 import _cabi, cols as cols_mod, reads as reads_mod
from oracle import restate
from oracle import summary as osum


class OracleEngine:
    """Same two methods as instrain_b200.engine.Engine, computed by the oracle (test infrastructure)."""

    def __init__(self):
        self.lut, self.dflt = load_lut()
        self.formats = []

    def profile_batch(self, ev, ref_codes, splits, min_cov=5, min_freq=0.05, min_snp=20, want=(), reads=None, cols=None, **kw):
        L = len(ref_codes)
        if cols is not None:
            self.formats.append("cols")
            events = cols_mod.cols_to_events(cols, L)
        elif "mis_word" in reads:
            self.formats.append("delta")
            seg_word, n_words, words = reads_mod.delta_to_words(reads, ref_codes)
            events = reads_mod.reads_to_events(dict(reads, seg_word=seg_word, n_words=n_words, words=words))
        else:
            self.formats.append("segments")
            events = reads_mod.reads_to_events(reads)
        events["pair_mm"] = np.asarray(ev["pair_mm"])
        out = restate.profile_events(events, ref_codes, self.lut, self.dflt, splits, start=int(kw.get("start", 0)), min_cov=min_cov,
                                     min_freq=min_freq, min_snp=min_snp, rarefied_coverage=int(kw.get("rarefied_coverage", 50)),
                                     seed=int(kw.get("seed", 0)))
        out["M"] = out["counts"].shape[1]
        for k in ("n_snv", "n_ld"):
            out[k] = len(out[k[2:]])
        return out

    def scaffold_summary(self, covT, clonT, nmask, bounds):
        """What K4 returns (include/instrain_b200.h, isb_summary_row), with numpy."""
        L, M = covT.shape
        n_sc = len(bounds) - 1
        rows = np.zeros(n_sc * M, dtype=_cabi.SUMMARY_DT)
        for s in range(n_sc):
            sl = slice(int(bounds[s]), int(bounds[s + 1]))
            cum = np.zeros(sl.stop - sl.start, dtype=np.int64)
            last = np.full(sl.stop - sl.start, np.nan, dtype=np.float32)
            for m in range(M):
                cum = cum + covT[sl, m]
                c = clonT[sl, m]
                last = np.where(np.isnan(c), last, c)
                r = rows[s * M + m]
                n = len(cum)
                vals = last[~np.isnan(last)]
                r["length"], r["nonzero"], r["sum_cov"] = n, np.count_nonzero(cum), cum.sum()
                r["sum_cov2"] = int((cum.astype(object) ** 2).sum())
                r["counted"], r["sum_clon"] = len(vals), float(vals.astype(np.float64).sum())
                sc = np.sort(cum)
                r["cov_med_lo"], r["cov_med_hi"] = sc[(n - 1) // 2], sc[n // 2]
                if len(vals):
                    sv = np.sort(vals)
                    r["clon_med_lo"], r["clon_med_hi"] = sv[(len(sv) - 1) // 2], sv[len(sv) // 2]
                else:
                    r["clon_med_lo"] = r["clon_med_hi"] = np.nan
                r["present"] = int((covT[sl, m] > 0).any() or (((nmask[sl] >> np.uint64(m)) & np.uint64(1)) != 0).any())
        return rows


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
    summ = pd.DataFrame(exp["sum_values"], columns=list(exp["sum_columns"]))
    summ.insert(0, "scaffold", np.array(sn, dtype=object)[exp["sum_scaffold"]])
    return snv[snv["scaffold"].isin(names)], ld[ld["scaffold"].isin(names)], summ[summ["scaffold"].isin(names)]


@pytest.mark.parametrize("transfer,threads", [("segments", 1), ("delta", 2), ("cols", 3)])
def test_profile_bam_host_side_against_goldens(tmp_path, transfer, threads):
    from instrain_b200.profile import profile_scaffolds
    from instrain_b200.store import SNVprofileStore, store_profile
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    eng = OracleEngine()
    res = profile_scaffolds(bam, rdic, seqs, engine=eng, b200_transfer=transfer, packer_threads=threads)
    assert eng.formats == [transfer] and not res.failures and sorted(res.scaffold_list) == sorted(rdic)
    g_snv, g_ld, g_sum = golden_tables(set(rdic))
    key = ["scaffold", "position", "mm"]
    a, b = res.raw_snp_table.sort_values(key).reset_index(drop=True), g_snv.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 1000
    for c in ["scaffold", "position", "mm", "A", "C", "T", "G", "con_base", "var_base", "allele_count", "class", "cryptic"]:
        assert (a[c].values == b[c].values).all(), c
    key = ["scaffold", "position_A", "position_B", "mm"]
    a, b = res.raw_linkage_table.sort_values(key).reset_index(drop=True), g_ld.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 1000
    for c in key + ["countAB", "countAb", "countaB", "countab"]:
        assert (a[c].values == b[c].values).all(), c
    for c in ("r2", "d_prime"):
        assert np.allclose(a[c].values, b[c].values, rtol=0, atol=1e-9, equal_nan=True), c
    key = ["scaffold", "mm"]
    a = res.cumulative_scaffold_table.sort_values(key).reset_index(drop=True)
    b = g_sum.sort_values(key).reset_index(drop=True)
    assert len(a) == len(b) > 30 and (a["scaffold"].values == b["scaffold"].values).all()
    for c in osum.COLUMNS:
        assert np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=0, atol=1e-9, equal_nan=True), c
    # the SNVprofile directory
    S = store_profile(str(tmp_path / "p.IS"), bam, res)
    S2 = SNVprofileStore(str(tmp_path / "p.IS"))
    assert S2.get("scaffold_list") == res.scaffold_list and len(S2.get("raw_snp_table")) == len(res.raw_snp_table)
    covT = S2.get("covT")
    for s, sp in res.scaffolds.items():
        assert set(covT[s]) == set(sp.covT)
        for mm in sp.covT:
            assert np.array_equal(covT[s][mm].values, sp.covT[mm].values) and np.array_equal(covT[s][mm].index.values, sp.covT[mm].index.values)
    assert S is not None


def test_profile_bam_runs_its_own_read_filter_and_reports_it(tmp_path, monkeypatch):
    """profile_bam(sR2M=None): the C++ read filter supplies sR2M, its report becomes the stored `mapping_info` attribute and
    output/*_mapping_info.tsv (settings as a `#` header line, the reference's leading columns), next to the SNVs /
    scaffold_info / linkage tables."""
    import instrain_b200.profile as P
    from instrain_b200.store import SNVprofileStore

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    monkeypatch.setattr(P, "Engine", E)
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    isp = str(tmp_path / "own_filter.IS")
    res = P.profile_bam(bam, None, None, isp, s2s=seqs, min_read_ani=0.95)
    res = getattr(res, "result", res)          # the on-disk object carries the in-memory tables as .result
    assert not res.failures and len(res.raw_snp_table) > 1000
    S = SNVprofileStore(isp)
    mi = S.get("mapping_info")
    assert mi["scaffold"].tolist()[0] == "all_scaffolds" and set(mi["scaffold"][1:]) >= set(res.scaffold_list)
    assert int(mi["filtered_pairs"][0]) == int(mi["filtered_pairs"][1:].sum()) > 1000
    out = os.path.join(isp, "output")
    base = "own_filter.IS_"
    assert {base + n + ".tsv" for n in ("SNVs", "scaffold_info", "linkage", "mapping_info")} <= set(os.listdir(out))
    lines = open(os.path.join(out, base + "mapping_info.tsv")).read().splitlines()
    assert lines[0] == "# min_read_ani:0.95 max_insert_relative:3 min_insert:50 min_mapq:-1 pairing_filter:paired_only"
    assert lines[1].split("\t")[:3] == ["scaffold", "pass_pairing_filter", "filtered_pairs"]
    snvs = pd.read_csv(os.path.join(out, base + "SNVs.tsv"), sep="\t")
    assert not snvs.duplicated(["scaffold", "position"]).any() and list(snvs.columns[:4]) == ["scaffold", "position", "position_coverage", "allele_count"]


def test_run_log_lines_follow_the_reference_format(tmp_path, monkeypatch, caplog):
    """profile_bam emits the reference's run-log lines at DEBUG level: "WorkerLog SplitProfile <scaffold>.<split> start|end
    RAM time PID" for every split, "WorkerLog MergeProfile <scaffold> ..." for every scaffold (inStrain/logUtils.py:940-975;
    parsed at :167 by splitting on whitespace) and "Checkpoint Profile <task> start|end RAM" (:903-937)."""
    import logging
    import instrain_b200.profile as P
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    monkeypatch.setattr(P, "Engine", E)
    monkeypatch.setenv("ISB_NATIVE_STORE", "1")
    name = "N5_271_010G1_scaffold_963"
    with caplog.at_level(logging.DEBUG):
        P.profile_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"), None, {name: rdic[name]}, str(tmp_path / "l.IS"), s2s=seqs,
                      window_length=300)
    lines = [l for r in caplog.records for l in r.getMessage().split("\n") if l.startswith(("WorkerLog", "Checkpoint"))]
    wl = [l.split() for l in lines if l.startswith("WorkerLog")]
    assert all(len(w) == 7 and w[3] in ("start", "end") and float(w[5]) > 0 and int(w[6]) == os.getpid() for w in wl)
    splits = sorted({w[2] for w in wl if w[1] == "SplitProfile"})
    assert splits == ["%s.%d" % (name, k) for k in range(len(splits))] and len(splits) >= 3
    assert [w[2] for w in wl if w[1] == "MergeProfile"] == [name, name]
    cp = [l.split() for l in lines if l.startswith("Checkpoint")]
    assert [(c[1], c[2], c[3]) for c in cp] == [("Profile", "B200_profile_scaffolds", "start"), ("Profile", "B200_profile_scaffolds", "end"),
                                                ("Profile", "B200_store", "start"), ("Profile", "B200_store", "end")]


def test_store_everything_keeps_the_pileup_counts(tmp_path, monkeypatch):
    """--store_everything: the SNVprofile gets `counts_table`, one [length, 4] array of A,C,T,G counts over all mm levels per
    scaffold (profile_utilities.py:167-168, 258-259, 709-715)."""
    import instrain_b200.profile as P
    from instrain_b200.store import SNVprofileStore
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))

    class E(OracleEngine):
        def __init__(self, *a, **k):
            super().__init__()

        def close(self):
            pass

    monkeypatch.setattr(P, "Engine", E)
    monkeypatch.setenv("ISB_NATIVE_STORE", "1")
    names = ["N5_271_010G1_scaffold_963", "N5_271_010G1_scaffold_62"]
    out = P.profile_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"), None, {n: rdic[n] for n in names}, str(tmp_path / "e.IS"), s2s=seqs,
                        store_everything=True)
    tab = SNVprofileStore(str(tmp_path / "e.IS")).get("counts_table")
    res = out.result
    assert len(tab) == 2 and [t.shape for t in tab] == [(len(seqs[n]), 4) for n in res.scaffold_list]
    for n, t in zip(res.scaffold_list, tab):
        cov = np.zeros(len(seqs[n]), dtype=np.int64)
        for mm, ser in res.scaffolds[n].covT.items():
            cov[ser.index.values] += ser.values
        assert np.array_equal(t.sum(axis=1), cov) and t.sum() > 1000      # total counts = coverage summed over the levels
    plain = P.profile_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"), None, {n: rdic[n] for n in names}, str(tmp_path / "p.IS"), s2s=seqs)
    assert SNVprofileStore(str(tmp_path / "p.IS")).get("counts_table") is None and plain.result.scaffolds[names[0]].pileup_counts is None
