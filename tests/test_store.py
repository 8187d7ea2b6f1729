"""the SNVprofile directory written natively (instrain_b200/store.py) -- CPU test.

The tables come from the ORACLE here (no GPU in this tier of tests): oracle rows -> instrain_b200.tables -> store_profile
-> read back through SNVprofileStore.get, and -- where the reference's test data is on this machine -- compared with the
reference's own stored raw_data/*.csv.gz (inStrain v1.7.0), column for column."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import assert_basewise_matches_digest, load_batch, load_lut
from instrain_b200 import tables
from instrain_b200.profile import ProfileResult, ScaffoldProfile
from instrain_b200.store import SNVprofileStore, store_profile
from oracle import restate

REF_RAW = "/root/reference/test/test_data/N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.forRC.IS/raw_data"


@pytest.fixture(scope="module")
def g1_result():
    lut, dflt = load_lut()
    b, _ = load_batch("G1")
    exp = restate.profile_events(b, b["ref_codes"], lut, dflt, b["splits"])
    names, offs, lens = list(b["scaffold_names"]), b["scaffold_off"].astype(np.int64), b["scaffold_len"]
    bases = np.array(list("ACTGN"))
    seqs = {n: "".join(bases[b["ref_codes"][o:o + l]]) for n, o, l in zip(names, offs, lens)}
    res = ProfileResult()
    res.raw_snp_table = tables.snv_table(exp["snv"], names, offs, seqs)
    res.raw_linkage_table = tables.linkage_table(exp["ld"], names, offs)
    res.cumulative_snv_table = tables.cumulative_snv_table(res.raw_snp_table)
    res.cumulative_scaffold_table = pd.DataFrame({"scaffold": names, "length": lens})
    for n, o, l in zip(names, offs, lens):
        sp = ScaffoldProfile(n, int(l))
        sl = slice(int(o), int(o + l))
        lv = tables.present_levels(exp["covT"][sl], exp["nmask"][sl])
        sp.covT = tables.basewise(exp["covT"][sl], "coverage", lv)
        sp.clonT = tables.basewise(exp["clonT"][sl], "clonality", lv)
        res.scaffolds[n] = sp
        res.scaffold_list.append(n)
    return res, b


def test_store_profile_layout_and_round_trip(tmp_path, g1_result):
    res, b = g1_result
    isp = str(tmp_path / "out.IS")
    S = store_profile(isp, "/some/where.bam", res)
    for lvl in ("output", "raw_data", "log", "figures"):
        assert os.path.isdir(os.path.join(isp, lvl))
    adb = pd.read_csv(os.path.join(isp, "raw_data", "attributes.tsv"), sep="\t", index_col="name")
    assert list(adb.columns) == ["value", "type", "description"]
    for name, typ in [("location", "value"), ("version", "value"), ("object_type", "value"), ("bam_loc", "value"),
                      ("scaffold_list", "list"), ("raw_linkage_table", "pandas"), ("raw_snp_table", "pandas"),
                      ("cumulative_scaffold_table", "pandas"), ("cumulative_snv_table", "pandas"),
                      ("scaffold_2_mm_2_read_2_snvs", "pickle"), ("covT", "special"), ("clonT", "special")]:
        assert adb.loc[name, "type"] == typ, name
        if typ != "value":
            assert os.path.exists(os.path.join(isp, "raw_data", os.path.basename(adb.loc[name, "value"]))), name
    S2 = SNVprofileStore(isp)                                   # re-open an existing directory
    assert S2.get("object_type") == "profile" and S2.get("bam_loc") == "/some/where.bam"
    assert S2.get("scaffold_list") == res.scaffold_list
    assert S2.get("scaffold_2_mm_2_read_2_snvs") == {} and S2.get("nope") is None
    snp = S2.get("raw_snp_table")
    assert list(snp.columns) == list(res.cumulative_snv_table.columns)
    for c in ("scaffold", "position", "mm", "A", "C", "T", "G", "con_base", "var_base", "class", "cryptic"):
        assert (snp[c].values == res.cumulative_snv_table[c].values).all(), c
    assert np.allclose(snp["var_freq"].values, res.cumulative_snv_table["var_freq"].values, rtol=0, atol=1e-15)
    ld = S2.get("raw_linkage_table")
    assert len(ld) == len(res.raw_linkage_table) == 14136
    # covT / clonT through the .hd5 files == the reference's stored ones (digests), every level and position
    covT, clonT = S2.get("covT"), S2.get("clonT")
    L, M = len(b["ref_codes"]), 15
    cov = np.zeros((L, M), np.int32)
    clon = np.full((L, M), np.nan, np.float32)
    nmask = np.zeros(L, np.uint64)
    for n, o in zip(b["scaffold_names"], b["scaffold_off"]):
        for mm, s in covT[str(n)].items():
            cov[o + s.index.values, mm] = s.values
            if len(s) == 0:
                nmask[o] |= np.uint64(1) << np.uint64(mm)     # level present without coverage: keep it a key
        for mm, s in clonT[str(n)].items():
            clon[o + s.index.values, mm] = s.values.astype(np.float32)
    assert assert_basewise_matches_digest("G1", b["scaffold_names"], b["scaffold_off"], b["scaffold_len"], cov, clon, nmask) == 1267
    only = S2.get("covT", scaffolds=[res.scaffold_list[3]])
    assert list(only) == [res.scaffold_list[3]]
    # overwrite with a different type is refused, same type is replaced (SNVprofile.store, SNVprofile.py:95-113)
    S2.store("bam_loc", "/other.bam", "value", "Location of .bam file")
    assert SNVprofileStore(isp).get("bam_loc") == "/other.bam"
    S2.store("bam_loc", ["x"], "list", "Location of .bam file")
    assert SNVprofileStore(isp).get("bam_loc") == "/other.bam"


@pytest.mark.skipif(not os.path.exists(REF_RAW), reason="reference test data not present on this machine")
def test_tables_equal_reference_stored_csv(g1_result):
    res, _ = g1_result
    key = ["scaffold", "position", "mm"]
    gold = pd.read_csv(os.path.join(REF_RAW, "cumulative_snv_table.csv.gz"), index_col=0).sort_values(key).reset_index(drop=True)
    mine = res.cumulative_snv_table.sort_values(key).reset_index(drop=True)
    assert list(mine.columns) == list(gold.columns)
    for c in gold.columns:
        if c.endswith("_freq"):
            assert np.allclose(mine[c].values.astype(float), gold[c].values.astype(float), rtol=0, atol=1e-12, equal_nan=True), c
        else:
            assert (mine[c].values == gold[c].values).all(), c
    raw = pd.read_csv(os.path.join(REF_RAW, "raw_snp_table.csv.gz"), index_col=0)
    assert list(raw.columns) == list(mine.columns)              # the reference's raw table carries the freq columns too
    key = ["scaffold", "position_A", "position_B", "mm"]
    gold = pd.read_csv(os.path.join(REF_RAW, "raw_linkage_table.csv.gz"), index_col=0).sort_values(key).reset_index(drop=True)
    mine = res.raw_linkage_table.sort_values(key).reset_index(drop=True)
    assert sorted(mine.columns) == sorted(gold.columns)
    for c in gold.columns:
        if c in ("r2_normalized", "d_prime_normalized"):        # unseeded random in the reference
            continue
        if c in ("r2", "d_prime"):
            assert np.allclose(mine[c].values, gold[c].values, rtol=0, atol=1e-9, equal_nan=True), c
        else:
            assert (mine[c].values == gold[c].values).all(), c


REF_IS = "/root/reference/test/test_data/N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.forRC.IS"


@pytest.mark.skipif(not os.path.isdir(REF_IS), reason="reference test data not present")
@pytest.mark.parametrize("name,key", [("SNVs", ["scaffold", "position"]), ("scaffold_info", ["scaffold"]),
                                      ("linkage", ["scaffold", "position_A", "position_B"])])
def test_generate_reproduces_reference_output_tables(tmp_path, name, key):
    """SNVprofileStore.generate (the user-facing output/*.tsv of the hot path's tables; SNVprofile.generate,
    SNVprofile.py:192-442) applied to the reference's OWN stored raw tables reproduces the reference's stored output file:
    same rows, same values, the reference's column order (columns the profile hot path does not own -- gene annotations
    -- excepted)."""
    import shutil
    from instrain_b200.store import SNVprofileStore
    loc = tmp_path / os.path.basename(REF_IS)
    os.makedirs(loc / "raw_data")
    for f in ("attributes.tsv", "cumulative_snv_table.csv.gz", "cumulative_scaffold_table.csv.gz", "raw_linkage_table.csv.gz"):
        shutil.copy(os.path.join(REF_IS, "raw_data", f), loc / "raw_data" / f)
    S = SNVprofileStore(str(loc))
    got = S.generate(name, return_table=True)
    out = str(loc / "output" / (os.path.basename(REF_IS) + "_" + name + ".tsv"))
    assert os.path.exists(out)
    ref = pd.read_csv(os.path.join(REF_IS, "output", os.path.basename(REF_IS) + "_" + name + ".tsv"), sep="\t")
    back = pd.read_csv(out, sep="\t")
    assert len(back) == len(got) == len(ref) > 100
    gene_cols = {"gene", "mutation", "mutation_type"}                  # merged in from SNP_mutation_types (gene module)
    common = [c for c in ref.columns if c in back.columns and c not in gene_cols]
    assert set(ref.columns) - set(back.columns) <= gene_cols
    assert [c for c in back.columns if c in common][:len(key)] == key  # the reference's leading column order
    a = back.sort_values(key).reset_index(drop=True)
    b = ref.sort_values(key).reset_index(drop=True)
    for c in common:
        if a[c].dtype.kind == "f" or b[c].dtype.kind == "f":
            assert np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=0, atol=1e-12, equal_nan=True), c
        else:
            assert (a[c].values == b[c].values).all(), c
