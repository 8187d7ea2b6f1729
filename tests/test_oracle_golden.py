"""Pin the oracle (oracle/oracle.c via oracle/restate.py) against the reference's golden vectors.

The expected arrays are the reference's OWN stored outputs (`raw_snp_table`, `raw_linkage_table` of the two
`forRC.IS` profile directories, inStrain v1.7.0) -- see tests/golden/make_golden.py.  This mirrors the reference's
exact-match regression test (test/tests/test_profile.py:831-1206) for the hot path's two tables.
"""
import os

import numpy as np
import pytest

from conftest import assert_ld_equal, assert_snv_equal, load_batch
from oracle import restate


def expected_rows(exp, ref_codes):
    snv = np.zeros(len(exp["snv_pos"]), dtype=restate.SNV_DT)
    snv["pos"], snv["mm"], snv["cnt"] = exp["snv_pos"], exp["snv_mm"], exp["snv_cnt"]
    snv["ref"] = ref_codes[exp["snv_pos"]]
    snv["con"], snv["var"] = exp["snv_con"], exp["snv_var"]
    snv["allele_count"], snv["cls"], snv["cryptic"] = exp["snv_allele_count"], exp["snv_cls"], exp["snv_cryptic"]
    ld = np.zeros(len(exp["ld_pos_a"]), dtype=restate.LD_DT)
    ld["pos_a"], ld["pos_b"], ld["mm"] = exp["ld_pos_a"], exp["ld_pos_b"], exp["ld_mm"]
    for i, f in enumerate(("c_AB", "c_Ab", "c_aB", "c_ab")):
        ld[f] = exp["ld_counts"][:, i]
    for i, f in enumerate(("allele_A", "allele_a", "allele_B", "allele_b")):
        ld[f] = exp["ld_alleles"][:, i]
    ld["r2"], ld["d_prime"] = exp["ld_r2"], exp["ld_d_prime"]
    return snv, ld


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_oracle_reproduces_reference_goldens(which, null_lut):
    batch, exp = load_batch(which)
    lut, dflt = null_lut
    out = restate.profile_events(batch, batch["ref_codes"], lut, dflt, batch["splits"])
    snv, ld = expected_rows(exp, batch["ref_codes"])
    assert_snv_equal(out["snv"], snv)
    assert_ld_equal(out["ld"], ld, tol=1e-9)
    # test_profile_13 (test/tests/test_profile.py:726-750): position_coverage == sum covT[mm' <= mm][pos]
    cov_cum = np.cumsum(out["covT"], axis=1)
    assert np.array_equal(cov_cum[out["snv"]["pos"], out["snv"]["mm"]], out["snv"]["cnt"].sum(1))


def test_null_lut_shape(null_lut):
    lut, dflt = null_lut
    assert dflt == 17 and lut[0] == -1 and lut[5] == 2 and len(lut) == 10000


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_summary_restatement_reproduces_cumulative_scaffold_table(which, null_lut):
    """oracle/summary.py (make_coverage_table restated) vs the reference's stored cumulative_scaffold_table
    (non-random columns), cf. the reference's test_profile_3 / test_profile_16."""
    from oracle import summary
    batch, exp = load_batch(which)
    lut, dflt = null_lut
    out = restate.profile_events(batch, batch["ref_codes"], lut, dflt, batch["splits"], do_linkage=False)
    rows, sidx = [], []
    for i in range(len(batch["scaffold_names"])):
        lo = int(batch["scaffold_off"][i])
        hi = lo + int(batch["scaffold_len"][i])
        sn = out["snv"][(out["snv"]["pos"] >= lo) & (out["snv"]["pos"] < hi)]
        for r in summary.scaffold_summary(out["covT"][lo:hi], out["clonT"][lo:hi], out["nmask"][lo:hi], sn, lo):
            rows.append([float(r[c]) for c in summary.COLUMNS])
            sidx.append(i)
    got = np.array(rows)
    assert list(exp["sum_columns"]) == summary.COLUMNS
    assert np.array_equal(np.array(sidx), exp["sum_scaffold"])
    assert np.allclose(got, exp["sum_values"], rtol=0, atol=1e-9, equal_nan=True)


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/test_data"), reason="reference tree not present")
@pytest.mark.parametrize("which,n_pairs", [("G1", 7179), ("G2", 9435)])
def test_read_filter_restatement_reproduces_rdic(which, n_pairs):
    """oracle/read_filter.py vs the reference's stored Rdic.json (sR2M) and mapping_info tallies."""
    import json
    import pandas as pd
    from oracle import bamio, read_filter
    td = "/root/reference/test/test_data/"
    bam = td + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010%s.sorted.bam" % which
    isd = td + "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010%s.forRC.IS/raw_data/" % which
    refs, reads = bamio.read_bam(bam)
    by = {}
    for r in reads:
        if r.tid >= 0:
            by.setdefault(refs[r.tid][0], []).append(r)
    out, tal, _ = read_filter.filter_pairs({s: read_filter.pair2info(rs) for s, rs in by.items()})
    rdic = json.load(open(isd + "Rdic.json"))
    assert {s: d for s, d in out.items() if d} == rdic and sum(len(v) for v in rdic.values()) == n_pairs
    mi = pd.read_csv(isd + "mapping_info.csv.gz")
    mi = mi[mi.scaffold != "all_scaffolds"].set_index("scaffold")
    assert all(int(mi.loc[s, c]) == tal[s][c] for s in mi.index if s in tal for c in tal[s])


def _param_case(name):
    """The reference's own outputs for non-default settings (tests/golden/make_param_goldens.py) + the oracle's inputs."""
    import hashlib
    z = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_G1_%s.npz" % name)))
    batch, _ = load_batch("G1")
    if bool(z["set_mode"]):
        batch["pair_mm"] = np.zeros_like(batch["pair_mm"])
    off = dict(zip(batch["scaffold_names"], batch["scaffold_off"].astype(np.int64)))
    ln = dict(zip(batch["scaffold_names"], batch["scaffold_len"].astype(np.int64)))
    keep = np.zeros(len(batch["ref_codes"]), dtype=bool)
    for s in z["scaffolds"]:
        keep[off[s]:off[s] + ln[s]] = True
    sha = lambda a: np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    return z, batch, off, ln, keep, sha


def check_against_param_golden(out, z, batch, off, ln, keep, sha):
    """`out` = profile of the WHOLE G1 batch (oracle or CUDA path) with the fixture's settings: its rows inside the chosen
    scaffolds and their dense covT / clonT must equal what the reference's own functions returned."""
    snv, ld = expected_rows(z, batch["ref_codes"])
    assert_snv_equal(out["snv"][keep[out["snv"]["pos"]]], snv)
    assert_ld_equal(out["ld"][keep[out["ld"]["pos_a"]]], ld, tol=1e-9)
    for s, M, c_sha, l_sha in zip(z["scaffolds"], z["M"], z["cov_sha"], z["clon_sha"]):
        sl = slice(int(off[s]), int(off[s] + ln[s]))
        cov, clon = out["covT"][sl, :M], np.array(out["clonT"][sl, :M], dtype=np.float32)
        clon[np.isnan(clon)] = np.float32(np.nan)          # one NaN bit pattern (the CUDA path's differs from numpy's)
        assert not out["covT"][sl, M:].any() and np.isnan(out["clonT"][sl, M:]).all(), s
        assert np.array_equal(sha(cov.astype(np.int32)), c_sha), "covT " + s
        assert np.array_equal(sha(clon.astype(np.float32)), l_sha), "clonT " + s


@pytest.mark.parametrize("name", ["params_P1", "params_P2", "params_P3"])
def test_oracle_reproduces_reference_with_other_settings(name):
    """Pins the oracle's PARAMETER handling on the reference: min_cov / min_freq / min_snp / fdr away from the defaults
    and set-mode R2M (--skip_mm_profiling), against outputs of the reference's own functions run on the same reads."""
    z, batch, off, ln, keep, sha = _param_case(name)
    out = restate.profile_events(batch, batch["ref_codes"], z["lut"], int(z["lut_default"]), batch["splits"],
                                 min_cov=int(z["min_cov"]), min_freq=float(z["min_freq"]), min_snp=int(z["min_snp"]))
    assert len(z["snv_pos"]) > 1000 and len(z["ld_pos_a"]) > 1000
    check_against_param_golden(out, z, batch, off, ln, keep, sha)


def load_ns_case():
    z = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c963_Ns.npz")))
    batch = {k: z[k] for k in ("ref_codes", "ref_pos", "base", "qual", "read_id", "pair_mm", "splits")}
    return z, batch


def check_ns_case(out, z, batch):
    snv, ld = expected_rows(z, batch["ref_codes"])
    assert_snv_equal(out["snv"], snv)
    assert_ld_equal(out["ld"], ld, tol=1e-9)
    assert np.array_equal(out["covT"], z["covT"])
    ok = ~np.isnan(z["clonT"])
    assert np.array_equal(np.isnan(out["clonT"]), ~ok)
    assert np.array_equal(out["clonT"][ok].view(np.uint32), z["clonT"][ok].view(np.uint32))


def test_oracle_reproduces_reference_on_n_reference_case():
    """The reference's edge-case BAM with an N in the reference sequence (scaffold_963_Ns, ~185x): rows (incl. the
    AmbiguousReference class at the N), covT and clonT at every position, against the reference's own functions
    (tests/golden/make_param_goldens.py ns)."""
    from conftest import load_lut
    z, batch = load_ns_case()
    lut, dflt = load_lut()
    out = restate.profile_events(batch, batch["ref_codes"], lut, dflt, batch["splits"])
    assert len(out["snv"]) == 686 and (out["snv"]["cls"] == 0).sum() >= 1          # class 0 = AmbiguousReference
    check_ns_case(out, z, batch)
