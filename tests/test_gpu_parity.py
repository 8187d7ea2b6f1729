"""GPU parity tests proper: the CUDA path (through the C-ABI, instrain_b200.engine.Engine) against the oracle
(oracle/restate.py -> oracle/oracle.c) on identical inputs, and against the reference's golden tables.

Bar (BASELINE.json north_star): per-position counts and SNV calls bit-exact; r2 / d_prime within 1e-6
(the comparisons below use 1e-9; the integer fields of linkage rows are compared exactly).
"""
import numpy as np
import pytest

from conftest import assert_basewise_matches_digest, assert_clontr_equal, assert_ld_equal, assert_snv_equal, load_batch
from oracle import restate, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(null_lut):
    from instrain_b200.engine import Engine
    e = Engine(0, null_lut[0], null_lut[1])
    yield e
    e.close()


def oracle_all(batch, null_lut, **kw):
    return restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], **kw)


def check_batch(eng, batch, null_lut, tol=1e-9, **kw):
    exp = oracle_all(batch, null_lut, **kw)
    M = exp["counts"].shape[1]
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M,
                            want=("counts", "nmask", "covT", "clonT", "clonTR", "site_flags", "snv", "ld"), **kw)
    assert np.array_equal(got["counts"], exp["counts"])
    assert np.array_equal(got["nmask"], exp["nmask"])
    assert np.array_equal(got["covT"], exp["covT"])
    assert_clontr_equal(got["clonTR"], exp["clonTR"])
    # clonality: float32 bit patterns identical (NaN = unset)
    assert np.array_equal(got["clonT"].view(np.uint32) == 0x7FC00000, np.isnan(exp["clonT"])) or \
        np.array_equal(np.isnan(got["clonT"]), np.isnan(exp["clonT"]))
    ok = ~np.isnan(exp["clonT"])
    assert np.array_equal(got["clonT"][ok].view(np.uint32), exp["clonT"][ok].view(np.uint32))
    assert np.array_equal(got["site_flags"], exp["site_flags"])
    assert_snv_equal(got["snv"], exp["snv"])
    assert_ld_equal(got["ld"], exp["ld"], tol=tol)
    return got, exp


# ---- golden vectors of the reference -------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["G1", "G2"])
def test_golden_tables_through_cabi(eng, which, null_lut):
    """The CUDA path reproduces the reference's stored raw_snp_table / raw_linkage_table (inStrain v1.7.0)."""
    from test_oracle_golden import expected_rows
    batch, exp = load_batch(which)
    got, _ = check_batch(eng, batch, null_lut)
    snv, ld = expected_rows(exp, batch["ref_codes"])
    assert_snv_equal(got["snv"], snv)
    assert_ld_equal(got["ld"], ld, tol=1e-6)
    # test_profile_13 (reference test/tests/test_profile.py:726-750)
    cov_cum = np.cumsum(got["covT"], axis=1)
    assert np.array_equal(cov_cum[got["snv"]["pos"], got["snv"]["mm"]], got["snv"]["cnt"].sum(1))
    # the reference's stored covT.hd5 / clonT.hd5 (digests): every position, every mm level, empty levels included
    n = assert_basewise_matches_digest(which, batch["scaffold_names"], batch["scaffold_off"], batch["scaffold_len"],
                                       got["covT"], got["clonT"], got["nmask"])
    assert n == {"G1": 1267, "G2": 1429}[which]


# ---- stage entry points ----------------------------------------------------------------------------------------------
def test_stage_entry_points(eng, null_lut):
    batch, _ = load_batch("G1")
    exp = oracle_all(batch, null_lut)
    L, M = exp["counts"].shape[:2]
    counts, nmask = eng.pileup_counts(batch, 0, L, M)
    assert np.array_equal(counts, exp["counts"]) and np.array_equal(nmask, exp["nmask"])
    covT, clonT, flags, snv = eng.call_snvs(counts, nmask, batch["ref_codes"], cap=16)   # forces the capacity retry
    assert np.array_equal(covT, exp["covT"]) and np.array_equal(flags, exp["site_flags"])
    assert_snv_equal(snv, exp["snv"])
    ld = eng.linkage(batch, counts, nmask, flags, batch["splits"], cap=16)
    assert_ld_equal(ld, exp["ld"], tol=1e-9)


def test_any_order_counts(eng, null_lut):
    """ISB_K1_ANY_ORDER: BAM-order (unsorted) events give the same counts through the atomic path."""
    batch, _ = load_batch("G1")
    exp = oracle_all(batch, null_lut, do_linkage=False)
    L, M = exp["counts"].shape[:2]
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(batch["ref_pos"]))
    ev = {k: np.ascontiguousarray(batch[k][perm]) for k in ("ref_pos", "base", "qual", "read_id")}
    ev["pair_mm"] = batch["pair_mm"]
    counts, nmask = eng.pileup_counts(ev, 0, L, M, any_order=True)
    assert np.array_equal(counts, exp["counts"]) and np.array_equal(nmask, exp["nmask"])


def test_unsorted_events_rejected(eng):
    from instrain_b200 import _cabi
    batch, _ = load_batch("G1")
    ev = {k: batch[k].copy() for k in ("ref_pos", "base", "qual", "read_id", "pair_mm")}
    ev["ref_pos"][1000:200000] = ev["ref_pos"][1000:200000][::-1].copy()
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_counts(ev, 0, len(batch["ref_codes"]), int(batch["pair_mm"].max()) + 1)
    assert ei.value.code == _cabi.ISB_ERR_ORDER


def test_argument_errors(eng):
    from instrain_b200 import _cabi
    ev = dict(ref_pos=np.zeros(4, np.int32), base=np.zeros(4, np.uint8), qual=np.full(4, 40, np.uint8),
              read_id=np.zeros(4, np.int32), pair_mm=np.array([70], np.uint8))
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_counts(ev, 0, 8, 65)
    assert ei.value.code == _cabi.ISB_ERR_ARG
    with pytest.raises(_cabi.IsbError) as ei:            # mm value >= M
        eng.pileup_counts(ev, 0, 8, 3)
    assert ei.value.code == _cabi.ISB_ERR_ARG


# ---- synthetic configs -------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,cov,dens,nsc,skip_mm,n_frac,seed", [
    (30000, 50, 0.01, 2, False, 0.0, 20260102),     # config-2 shaped, M ~ 12
    (30000, 50, 0.01, 1, True, 0.0, 20260102),      # --skip_mm_profiling, M = 1
    (12000, 100, 0.05, 1, False, 0.002, 20260105),  # LD-stress shaped + non-ACGT read bases (nmask)
    (700, 30, 0.02, 3, False, 0.0, 7),              # tiny scaffolds (test_profile_18), one split each
    (25000, 8, 0.01, 1, False, 0.0, 11),            # low coverage around min_cov
    (12000, 500, 0.05, 1, True, 0.0, 20260105),     # BASELINE configs[4] shaped (LD stress): 500x, 5 % SNVs, wide bit rows
])
def test_synthetic_parity(eng, null_lut, L, cov, dens, nsc, skip_mm, n_frac, seed):
    batch = synth.make_batch(L, cov, dens, seed, n_scaffolds=nsc, skip_mm=skip_mm, n_frac=n_frac)
    got, exp = check_batch(eng, batch, null_lut)
    assert len(exp["snv"]) > 0


def test_thresholds_are_parameters(eng, null_lut):
    batch = synth.make_batch(20000, 40, 0.02, 3)
    check_batch(eng, batch, null_lut, min_cov=10, min_freq=0.1, min_snp=5)
    check_batch(eng, batch, null_lut, min_cov=1, min_freq=0.01, min_snp=50)


def test_ambiguous_reference_and_empty_regions(eng, null_lut):
    """N in the reference (test_special_2), zero-coverage stretches, events outside [start, start+L)."""
    batch = synth.make_batch(20000, 40, 0.02, 5)
    batch["ref_codes"] = batch["ref_codes"].copy()
    batch["ref_codes"][::37] = 4
    keep = (batch["ref_pos"] < 5000) | (batch["ref_pos"] > 9000)          # a coverage hole
    for k in ("ref_pos", "base", "qual", "read_id"):
        batch[k] = np.ascontiguousarray(batch[k][keep])
    got, exp = check_batch(eng, batch, null_lut)
    assert (exp["snv"]["cls"] == 0).any()


def test_empty_batch(eng, null_lut):
    L = 1000
    ev = dict(ref_pos=np.zeros(0, np.int32), base=np.zeros(0, np.uint8), qual=np.zeros(0, np.uint8),
              read_id=np.zeros(0, np.int32), pair_mm=np.zeros(0, np.uint8))
    got = eng.profile_batch(ev, np.zeros(L, np.uint8), [(0, L - 1)], M=1, want=("counts", "covT", "clonT", "snv", "ld"))
    assert got["counts"].sum() == 0 and got["covT"].sum() == 0 and np.isnan(got["clonT"]).all()
    assert len(got["snv"]) == 0 and len(got["ld"]) == 0


def test_double_entries_and_self_edges(eng, null_lut):
    """Both mates of a pair counted on one site (htslib's overlap quirk produces this): multiplicity-2 combos and
    self edges (p, p) must match the reference's itertools.combinations semantics."""
    batch = synth.make_batch(6000, 120, 0.03, 9, skip_mm=False)
    exp0 = oracle_all(batch, null_lut)
    sites = np.nonzero(exp0["site_flags"] & 0x10)[0]
    rng = np.random.default_rng(3)
    pos = batch["ref_pos"]
    dup_idx = []
    for p in sites[::3]:
        lo, hi = np.searchsorted(pos, p), np.searchsorted(pos, p + 1)
        cand = np.arange(lo, hi)[batch["qual"][lo:hi] >= 30]
        dup_idx.extend(rng.choice(cand, size=min(len(cand), 40), replace=False).tolist())
    dup_idx = np.array(sorted(dup_idx))
    new = {k: np.concatenate([batch[k], batch[k][dup_idx]]) for k in ("ref_pos", "base", "qual", "read_id")}
    # half of the duplicated entries show a different base than the original mate
    flip = rng.random(len(dup_idx)) < 0.5
    nb = new["base"][len(pos):]
    nb[flip] = (nb[flip] + 1) % 4
    order = np.argsort(new["ref_pos"], kind="stable")
    for k in new:
        batch[k] = np.ascontiguousarray(new[k][order])
    got, exp = check_batch(eng, batch, null_lut, min_snp=10)
    assert (exp["ld"]["pos_a"] == exp["ld"]["pos_b"]).any(), "test did not produce a self edge"


def test_triple_entry_is_reported(eng, null_lut):
    from instrain_b200 import _cabi
    batch = synth.make_batch(3000, 60, 0.03, 4, skip_mm=True)
    exp0 = oracle_all(batch, null_lut)
    p = int(np.nonzero(exp0["site_flags"] & 0x10)[0][0])
    pos = batch["ref_pos"]
    lo, hi = np.searchsorted(pos, p), np.searchsorted(pos, p + 1)
    e = lo + int(np.nonzero((batch["qual"][lo:hi] >= 30) & ((exp0["site_flags"][p] >> batch["base"][lo:hi]) & 1 > 0))[0][0])
    for k in ("ref_pos", "base", "qual", "read_id"):
        batch[k] = np.ascontiguousarray(np.insert(batch[k], [e, e], [batch[k][e], batch[k][e]]))
    with pytest.raises(_cabi.IsbError) as ei:
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1)
    assert ei.value.code == _cabi.ISB_ERR_UNSUPPORTED


# ---- size-independent properties at larger size (no oracle) -----------------------------------------------------------
def test_properties_large(eng, null_lut):
    batch = synth.make_batch(400000, 60, 0.01, 20260103, skip_mm=True)
    L = len(batch["ref_codes"])
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, want=("counts", "covT", "snv", "ld"))
    qual_ok = (batch["qual"] >= 30) & (batch["base"] < 4)
    assert got["counts"].sum() == qual_ok.sum()                                    # checksum of checksums
    assert np.array_equal(got["counts"].sum((1, 2)), np.bincount(batch["ref_pos"][qual_ok], minlength=L))
    assert np.array_equal(got["covT"][:, 0], got["counts"].sum((1, 2)))
    # linearity: counts(evA + evB) == counts(evA) + counts(evB)
    half = (batch["read_id"] % 2) == 0
    sub = lambda m: {**{k: np.ascontiguousarray(batch[k][m]) for k in ("ref_pos", "base", "qual", "read_id")},
                     "pair_mm": batch["pair_mm"]}
    ca, _ = eng.pileup_counts(sub(half), 0, L, 1)
    cb, _ = eng.pileup_counts(sub(~half), 0, L, 1)
    assert np.array_equal(ca + cb, got["counts"])
    # idempotence / determinism of the whole path
    again = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, want=("snv", "ld"))
    assert_snv_equal(again["snv"], got["snv"])
    assert_ld_equal(again["ld"], got["ld"], tol=0)
    # every linkage row: pos_a <= pos_b, same split, total > min_snp
    ld = got["ld"]
    assert (ld["pos_a"] <= ld["pos_b"]).all()
    sp = np.searchsorted(batch["splits"][:, 0], ld["pos_a"], side="right")
    assert np.array_equal(sp, np.searchsorted(batch["splits"][:, 0], ld["pos_b"], side="right"))
    assert ((ld["c_AB"] + ld["c_Ab"] + ld["c_aB"] + ld["c_ab"]) > 20).all()


# ---- device-generated data (the bench's generator) ---------------------------------------------------------------------
@pytest.mark.parametrize("skip_mm", [True, False])
def test_device_generated_batch_parity(eng, null_lut, skip_mm):
    """instrain_b200.synth (bench data generator): columns are position-major, and sampled scaffolds of it give the
    same answer on the CUDA path and the oracle ('sampled splits vs oracle')."""
    import torch
    from instrain_b200 import synth as dsynth
    d = dsynth.generate(0, 60000, 3, 60, 0.01, 20260103, skip_mm=skip_mm)
    pos = d["ref_pos"].cpu().numpy()
    assert (np.diff(pos) >= 0).all() and pos.min() >= 0 and pos.max() < 180000
    cov = len(pos) / 180000.0
    assert 45 < cov < 66, cov
    hb = dsynth.to_host_batch(d, 1, 2)                      # scaffolds 1..2 as a self-contained host batch
    assert hb["ref_pos"].min() >= 0 and hb["read_id"].min() == 0
    got, exp = check_batch(eng, hb, null_lut)
    assert len(exp["snv"]) > 100 and len(exp["ld"]) > 10
    # whole device-resident data set through device pointers == per-slice host results
    M = int(d["pair_mm"].max().item()) + 1 if d["pair_mm"].numel() else 1
    full = eng.profile_batch(dict(ref_pos=d["ref_pos"], base=d["base"], qual=d["qual"], read_id=d["read_id"],
                                  pair_mm=d["pair_mm"].cpu().numpy()), d["ref_codes"].cpu().numpy(),
                             d["splits"].cpu().numpy(), M=M, want=("snv", "ld"))
    sel = full["snv"][full["snv"]["pos"] >= 60000].copy()
    sel["pos"] -= 60000
    if M == got["M"]:
        assert_snv_equal(sel, got["snv"])


# ---- packed transfer format (K0 expansion) -------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["G1", "synthetic_mm", "escapes_and_N"])
def test_packed_transfer_format_parity(eng, null_lut, case):
    """isb_profile_batch_packed (compact host->device format, expanded by K0) gives the same tables as the columnar
    entry point and the oracle -- including pair-id jumps > 14 (escape path) and non-ACGT bases."""
    from instrain_b200.packed import encode_packed
    if case == "G1":
        batch, _ = load_batch("G1")
    elif case == "synthetic_mm":
        batch = synth.make_batch(30000, 60, 0.02, 77)
    else:
        batch = synth.make_batch(20000, 40, 0.02, 78, n_frac=0.003)
        keep = (batch["read_id"] % 23) < 2                     # thin the pairs: large id deltas inside positions
        for k in ("ref_pos", "base", "qual", "read_id"):
            batch[k] = np.ascontiguousarray(batch[k][keep])
    L = len(batch["ref_codes"])
    pk = encode_packed(batch, 0, L, 30)
    if case == "escapes_and_N":
        assert len(pk["esc_evt"]) > 100
    exp = oracle_all(batch, null_lut)
    M = exp["counts"].shape[1]
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, packed=pk,
                            want=("counts", "nmask", "covT", "clonT", "site_flags", "snv", "ld"))
    assert np.array_equal(got["counts"], exp["counts"]) and np.array_equal(got["nmask"], exp["nmask"])
    assert np.array_equal(got["covT"], exp["covT"]) and np.array_equal(got["site_flags"], exp["site_flags"])
    assert_snv_equal(got["snv"], exp["snv"])
    assert_ld_equal(got["ld"], exp["ld"], tol=1e-9)


# ---- K4: merge-stage summary -------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["G1", "G2"])
def test_scaffold_summary_matches_golden_cumulative_scaffold_table(eng, null_lut, which):
    """isb_scaffold_summary + instrain_b200.summary vs the reference's stored cumulative_scaffold_table (non-random
    columns; 1e-9) and vs the oracle's numpy restatement."""
    from instrain_b200 import summary as psum
    from oracle import summary as osum
    batch, exp = load_batch(which)
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], want=("covT", "clonT", "nmask", "snv"),
                            skip_linkage=True)
    bounds = np.append(batch["scaffold_off"], len(batch["ref_codes"])).astype(np.int32)
    rows = eng.scaffold_summary(got["covT"], got["clonT"], got["nmask"], bounds)
    names = list(batch["scaffold_names"])
    tab = psum.summary_table(rows, got["snv"], names, batch["scaffold_off"], got["M"])
    assert np.array_equal(tab["scaffold"].map({n: i for i, n in enumerate(names)}).values, exp["sum_scaffold"])
    cols = list(exp["sum_columns"])
    assert np.allclose(tab[cols].values.astype(float), exp["sum_values"], rtol=0, atol=1e-9, equal_nan=True)
    # raw reductions are exact integers: compare with numpy on the same arrays
    M = got["M"]
    r = rows.reshape(len(names), M)
    s = 7
    lo, hi = int(bounds[s]), int(bounds[s + 1])
    cum = np.cumsum(got["covT"][lo:hi].astype(np.int64), axis=1)
    assert np.array_equal(r[s]["sum_cov"], cum.sum(0)) and np.array_equal(r[s]["sum_cov2"].astype(np.int64), (cum * cum).sum(0))
    assert np.array_equal(r[s]["nonzero"], (cum > 0).sum(0))


def test_scaffold_summary_synthetic(eng, null_lut):
    from instrain_b200 import summary as psum
    from oracle import summary as osum
    batch = synth.make_batch(7001, 45, 0.02, 5, n_scaffolds=3)          # odd length: single middle element
    exp = oracle_all(batch, null_lut, do_linkage=False)
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], want=("covT", "clonT", "nmask", "snv"), skip_linkage=True)
    bounds = np.array([0, 7001, 14002, 21003], np.int32)
    rows = eng.scaffold_summary(got["covT"], got["clonT"], got["nmask"], bounds)
    tab = psum.summary_table(rows, got["snv"], ["a", "b", "c"], bounds[:-1], got["M"])
    ref = []
    for i in range(3):
        lo, hi = int(bounds[i]), int(bounds[i + 1])
        sn = exp["snv"][(exp["snv"]["pos"] >= lo) & (exp["snv"]["pos"] < hi)]
        for r in osum.scaffold_summary(exp["covT"][lo:hi], exp["clonT"][lo:hi], exp["nmask"][lo:hi], sn, lo):
            ref.append([float(r[c]) for c in osum.COLUMNS])
    assert np.allclose(tab[osum.COLUMNS].values.astype(float), np.array(ref), rtol=0, atol=1e-9, equal_nan=True)


# ---- chunk pipeline (K1 of chunk c+1 overlapping K2/K3 of chunk c) -----------------------------------------------------
@pytest.mark.parametrize("skip_mm", [True, False])
def test_chunk_pipeline_equals_single_pass(eng, skip_mm):
    """With ISB_PIPELINE, batches >= 2^22 positions are cut at split boundaries and pipelined over two streams; the tables
    must equal the single-pass ones exactly."""
    from instrain_b200 import synth as dsynth
    d = dsynth.generate(0, 300000, 15, 12, 0.01, 99, skip_mm=skip_mm)            # 4.5e6 positions, ~5e7 events
    ev = dict(ref_pos=d["ref_pos"], base=d["base"], qual=d["qual"], read_id=d["read_id"], pair_mm=d["pair_mm"].cpu().numpy())
    ref, spl = d["ref_codes"].cpu().numpy(), d["splits"].cpu().numpy()
    M = int(ev["pair_mm"].max()) + 1 if len(ev["pair_mm"]) else 1
    want = ("counts", "covT", "clonT", "site_flags", "snv", "ld")
    a = eng.profile_batch(ev, ref, spl, M=M, want=want, min_cov=3, min_snp=3)
    b = eng.profile_batch(ev, ref, spl, M=M, want=want, min_cov=3, min_snp=3, pipeline=True)
    assert len(a["snv"]) > 1000 and len(a["ld"]) > 100
    for k in ("counts", "covT", "site_flags"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
    assert_snv_equal(a["snv"], b["snv"])
    assert_ld_equal(a["ld"], b["ld"], tol=0)
    assert (a["n_sites"], a["n_site_pairs"]) == (b["n_sites"], b["n_site_pairs"])


@pytest.mark.parametrize("min_qual", [0, 1, 41, 129, 200])
def test_min_base_quality_is_a_parameter(eng, null_lut, min_qual):
    """K1's byte-SIMD quality test covers 1..128; 0 and > 128 take the scalar compare -- all must equal the oracle."""
    for skip_mm in (True, False):
        batch = synth.make_batch(9000, 70, 0.02, 31, skip_mm=skip_mm)
        check_batch(eng, batch, null_lut, min_qual=min_qual)


def test_fast_quotient_is_exactly_ieee_division(eng):
    """K2 computes the four base frequencies of a site from one reciprocal (Markstein correction).  Exhaustive check of
    every 0 <= c <= s for all s in the range K2 uses it for (s <= 65536; 2.1e9 pairs), bit for bit against __ddiv_rn."""
    assert eng.lib.isb_selftest_division(eng.ctx, 1, 65536) == 0
