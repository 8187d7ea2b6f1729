"""GPU parity of the COLUMN-WORD path (isb_pileup_cols / isb_profile_cols: K1c -> K2 -> K3, and the fused K1c + SNV call
at M = 1) against the oracle on the event columns the words encode, against the reference's golden tables, and against
the read-major CUDA path; plus the device-side layout conversion against the host packer's."""
import numpy as np
import pytest

from conftest import assert_clontr_equal, assert_ld_equal, assert_snv_equal, load_batch
from oracle import restate, synth
from instrain_b200 import cols, reads

pytestmark = pytest.mark.gpu

FULL = ("counts", "nmask", "covT", "clonT", "clonTR", "site_flags", "snv", "ld")
LEAN = ("covT", "clonT", "clonTR", "site_flags", "snv", "ld")  # no counts / nmask: the fused kernel at M = 1


@pytest.fixture(scope="module")
def eng(null_lut):
    from instrain_b200.engine import Engine
    e = Engine(0, null_lut[0], null_lut[1])
    yield e
    e.close()


def check_cols(eng, batch, null_lut, tol=1e-9, rd=None, cd=None, wants=(FULL, LEAN), **kw):
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], **kw)
    L, M = exp["counts"].shape[:2]
    if cd is None:
        if rd is None:
            rd = reads.events_to_reads(batch, kw.get("min_qual", 30))
        cd = cols.reads_to_cols(rd, L)
    got = None
    for want in wants:
        got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, cols=cd, want=want, **kw)
        if "counts" in want:
            assert np.array_equal(got["counts"], exp["counts"])
            assert np.array_equal(got["nmask"], exp["nmask"])
        assert np.array_equal(got["covT"], exp["covT"])
        ok = ~np.isnan(exp["clonT"])
        assert np.array_equal(np.isnan(got["clonT"]), ~ok)
        assert np.array_equal(got["clonT"][ok].view(np.uint32), exp["clonT"][ok].view(np.uint32))
        assert np.array_equal(got["site_flags"], exp["site_flags"])
        assert_clontr_equal(got["clonTR"], exp["clonTR"])
        assert_snv_equal(got["snv"], exp["snv"])
        assert_ld_equal(got["ld"], exp["ld"], tol=tol)
    return got, exp


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_golden_tables_cols(eng, which, null_lut):
    """Column words rebuilt from the golden event columns reproduce the reference's raw_snp_table / raw_linkage_table."""
    from test_oracle_golden import expected_rows
    batch, exp = load_batch(which)
    got, _ = check_cols(eng, batch, null_lut, wants=(FULL,))
    snv, ld = expected_rows(exp, batch["ref_codes"])
    assert_snv_equal(got["snv"], snv)
    assert_ld_equal(got["ld"], ld, tol=1e-6)


def test_golden_collapsed_fused(eng, null_lut):
    """The golden batch with every pair at mm = 0 (what --skip_mm_profiling feeds): the fused M = 1 kernel on real reads
    (indels, clipped ends, double entries -> self edges)."""
    batch, _ = load_batch("G1")
    batch = dict(batch)
    batch["pair_mm"] = np.zeros_like(batch["pair_mm"])
    check_cols(eng, batch, null_lut)


@pytest.mark.parametrize("L,cov,dens,nsc,skip_mm,n_frac,seed", [
    (30000, 50, 0.01, 2, False, 0.0, 20260102),
    (30000, 50, 0.01, 1, True, 0.0, 20260102),
    (12000, 100, 0.05, 1, False, 0.002, 20260105),   # non-ACGT read bases -> nmask
    (12000, 100, 0.05, 1, True, 0.002, 20260105),    # same at M = 1 (fused kernel reads nmask at empty positions)
    (700, 30, 0.02, 3, False, 0.0, 7),
    (1001, 30, 0.02, 3, True, 0.0, 7),               # L not a multiple of 8 / 64 / 256: ragged last column, group and warp
    (25000, 8, 0.01, 1, False, 0.0, 11),
    (25000, 3, 0.01, 1, True, 0.0, 12),              # coverage below min_cov almost everywhere, empty columns
    (9000, 400, 0.01, 1, True, 0.0, 3),              # > 255 words per column: 8 -> 32 bit widening
    (9000, 400, 0.01, 1, False, 0.0, 3),             # same through the shared-memory counters (mid-run flush)
    (12000, 500, 0.05, 1, True, 0.0, 20260105),      # BASELINE configs[4] shaped (LD stress): 500x, 5 % SNVs
])
def test_synthetic_parity_cols(eng, null_lut, L, cov, dens, nsc, skip_mm, n_frac, seed):
    batch = synth.make_batch(L, cov, dens, seed, n_scaffolds=nsc, skip_mm=skip_mm, n_frac=n_frac)
    check_cols(eng, batch, null_lut)


def test_pileup_cols_stage(eng, null_lut):
    batch, _ = load_batch("G1")
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], do_linkage=False)
    L, M = exp["counts"].shape[:2]
    cd = cols.reads_to_cols(reads.events_to_reads(batch), L)
    counts, nmask = eng.pileup_cols(cd, batch["pair_mm"], 0, L, M)
    assert np.array_equal(counts, exp["counts"]) and np.array_equal(nmask, exp["nmask"])


def test_many_mm_levels_cols(eng, null_lut):
    """M > 32: two level groups of the shared-memory counters."""
    batch = synth.make_batch(6000, 60, 0.02, 5, skip_mm=False)
    rng = np.random.default_rng(0)
    batch["pair_mm"] = rng.integers(0, 50, len(batch["pair_mm"])).astype(batch["pair_mm"].dtype)
    check_cols(eng, batch, null_lut)


def test_device_conversion_equals_host(eng):
    """isb_cols_from_reads (device) and isb_cols_from_reads_host (C++ packer) build the same arrays, also from segments
    of every length, clustered starts and blocks split at 256."""
    rng = np.random.default_rng(4)
    L = 5000
    starts, lens, pairs, codes = [], [], [], []
    for i in range(3000):
        n = int(rng.integers(1, 41)) if i % 7 else int(rng.integers(257, 700))
        s = int(rng.integers(0, L - n))
        starts.append(s); lens.append(n); pairs.append(i // 2)
        c = rng.integers(0, 5, n).astype(np.uint8)
        c[rng.random(n) < 0.2] = reads.NO_EVENT
        codes.append(c)
    rd1 = reads.build_reads(starts, lens, pairs, np.concatenate(codes))
    batch, _ = load_batch("G2")
    rd2 = reads.events_to_reads(batch)
    for rd, Lx in ((rd1, L), (rd2, len(batch["ref_codes"]))):
        h = cols.reads_to_cols(rd, Lx)
        d = eng.cols_from_reads(rd, Lx)
        assert h["n_chunks"] == d["n_chunks"] and np.array_equal(h["grp_off"], d["grp_off"])
        assert np.array_equal(h["words"], d["words"]) and np.array_equal(h["ids"], d["ids"])


def test_short_segments_cols(eng, null_lut):
    rng = np.random.default_rng(5)
    L = 5000
    starts, lens, pairs, codes = [], [], [], []
    for i in range(3000):
        n = int(rng.integers(1, 41)) if i % 7 else int(rng.integers(257, 700))
        s = int(rng.integers(0, L - n))
        starts.append(s); lens.append(n); pairs.append(i // 2)
        c = rng.integers(0, 5, n).astype(np.uint8)
        c[rng.random(n) < 0.2] = reads.NO_EVENT
        codes.append(c)
    rd = reads.build_reads(starts, lens, pairs, np.concatenate(codes))
    ev = reads.reads_to_events(rd)
    ev["pair_mm"] = rng.integers(0, 6, 1500).astype(np.uint8)
    ev["ref_codes"] = rng.integers(0, 4, L).astype(np.uint8)
    ev["splits"] = np.array([[0, L - 1]], dtype=np.int32)
    check_cols(eng, ev, null_lut, rd=rd)
    ev["pair_mm"][:] = 0
    check_cols(eng, ev, null_lut, rd=rd)


def test_empty_and_sparse_cols(eng, null_lut):
    L = 3000
    ref = np.zeros(L, np.uint8)
    cd = cols.reads_to_cols(reads.build_reads([], [], [], []), L)
    assert cd["n_chunks"] == 0
    for want in (("counts", "covT", "snv", "ld"), ("covT", "snv", "ld")):
        got = eng.profile_batch(dict(pair_mm=np.zeros(0, np.uint8)), ref, np.array([[0, L - 1]], np.int32), M=1, cols=cd,
                                want=want)
        assert got["covT"].sum() == 0 and got["n_snv"] == 0 and got["n_ld"] == 0
    rd = reads.build_reads([2990], [10], [0], np.full(10, 2, np.uint8))     # one segment touching the last position
    cd = cols.reads_to_cols(rd, L)
    counts, nmask = eng.pileup_cols(cd, np.zeros(1, np.uint8), 0, L, 1)
    assert counts[2990:, 0, 2].tolist() == [1] * 10 and counts.sum() == 10 and nmask.sum() == 0


def test_cols_layout_violations_rejected(eng):
    from instrain_b200 import _cabi
    batch, _ = load_batch("G1")
    L, M = len(batch["ref_codes"]), int(batch["pair_mm"].max()) + 1
    cd = cols.reads_to_cols(reads.events_to_reads(batch), L)
    bad = dict(cd); bad["grp_off"] = cd["grp_off"].copy(); bad["grp_off"][100:200] = bad["grp_off"][100:200][::-1]
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_cols(bad, batch["pair_mm"], 0, L, M)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    bad = dict(cd); bad["grp_off"] = cd["grp_off"].copy(); bad["grp_off"][-1] += 7       # beyond n_chunks
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_cols(bad, batch["pair_mm"], 0, L, M)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    with pytest.raises(_cabi.IsbError) as ei:                                            # n_groups does not match L
        eng.pileup_cols(cd, batch["pair_mm"], 0, L - 5000, M)
    assert ei.value.code == _cabi.ISB_ERR_ARG
    with pytest.raises(_cabi.IsbError) as ei:                                            # mm value >= M
        eng.pileup_cols(cd, batch["pair_mm"], 0, L, M - 1)
    assert ei.value.code == _cabi.ISB_ERR_ARG
    bad = dict(cd); bad["ids"] = cd["ids"].copy(); bad["ids"][bad["ids"] >= 0] += len(batch["pair_mm"])   # ids out of range
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_cols(bad, batch["pair_mm"], 0, L, M)
    assert ei.value.code == _cabi.ISB_ERR_ORDER


def test_cols_equal_read_major_larger(eng, null_lut):
    """Column-word and read-major CUDA paths on a batch the oracle would take minutes for: identical tables, fused and
    unfused."""
    for skip_mm in (True, False):
        batch = synth.make_batch(400000, 60, 0.01, 99, n_scaffolds=2, skip_mm=skip_mm)
        rd = reads.events_to_reads(batch, max_len=150 if skip_mm else 256)
        L = len(batch["ref_codes"])
        cd = cols.reads_to_cols(rd, L)
        M = int(batch["pair_mm"].max()) + 1
        a = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=FULL, reads=rd)
        for want in (FULL, LEAN):
            b = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want, cols=cd)
            for k in ("counts", "nmask", "covT", "site_flags"):
                if k in want:
                    assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
            assert_snv_equal(a["snv"], b["snv"])
            assert_ld_equal(a["ld"], b["ld"], tol=0.0)


def test_device_generator_cols(eng, null_lut):
    """bench.py's resident data set: the generator's read-major batch converted on the device (isb_cols_from_reads with
    device pointers) gives the same tables as the read-major path on the same data."""
    from instrain_b200 import synth as dsynth
    for skip_mm in (True, False):
        d = dsynth.generate(0, 50000, 3, 60, 0.01, 77, skip_mm=skip_mm, events=False, reads=True)
        cd = dsynth.reads_to_cols_device(eng, d)
        M = int(d["pair_mm"].max().item()) + 1
        ev = dict(pair_mm=d["pair_mm"].cpu().numpy())
        ref, splits = d["ref_codes"].cpu().numpy(), d["splits"].cpu().numpy()
        a = eng.profile_batch(ev, ref, splits, M=M, want=FULL, reads=d["reads"])
        for want in (FULL, LEAN):
            b = eng.profile_batch(ev, ref, splits, M=M, want=want, cols=cd)
            for k in ("counts", "nmask", "covT", "site_flags"):
                if k in want:
                    assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
            assert_snv_equal(a["snv"], b["snv"])
            assert_ld_equal(a["ld"], b["ld"], tol=0.0)
        assert a["n_snv"] > 100 and a["n_ld"] > 10
