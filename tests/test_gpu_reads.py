"""GPU parity of the READ-MAJOR path (isb_pileup_reads / isb_profile_reads: K1r -> K2 -> K3 with site events gathered
from the aligned segments) against the oracle on the event columns the segments encode, against the reference's golden
tables, and against the position-major CUDA path."""
import numpy as np
import pytest

from conftest import assert_clontr_equal, assert_ld_equal, assert_snv_equal, load_batch
from oracle import restate, synth
from instrain_b200 import reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(null_lut):
    from instrain_b200.engine import Engine
    e = Engine(0, null_lut[0], null_lut[1])
    yield e
    e.close()


def check_reads(eng, batch, null_lut, tol=1e-9, rd=None, **kw):
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], **kw)
    M = exp["counts"].shape[1]
    if rd is None:
        rd = reads.events_to_reads(batch, kw.get("min_qual", 30))
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, reads=rd,
                            want=("counts", "nmask", "covT", "clonT", "clonTR", "site_flags", "snv", "ld"), **kw)
    assert_clontr_equal(got["clonTR"], exp["clonTR"])
    assert np.array_equal(got["counts"], exp["counts"])
    assert np.array_equal(got["nmask"], exp["nmask"])
    assert np.array_equal(got["covT"], exp["covT"])
    ok = ~np.isnan(exp["clonT"])
    assert np.array_equal(np.isnan(got["clonT"]), ~ok)
    assert np.array_equal(got["clonT"][ok].view(np.uint32), exp["clonT"][ok].view(np.uint32))
    assert np.array_equal(got["site_flags"], exp["site_flags"])
    assert_snv_equal(got["snv"], exp["snv"])
    assert_ld_equal(got["ld"], exp["ld"], tol=tol)
    return got, exp


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_golden_tables_read_major(eng, which, null_lut):
    """Segments rebuilt from the golden event columns reproduce the reference's raw_snp_table / raw_linkage_table."""
    from test_oracle_golden import expected_rows
    batch, exp = load_batch(which)
    got, _ = check_reads(eng, batch, null_lut)
    snv, ld = expected_rows(exp, batch["ref_codes"])
    assert_snv_equal(got["snv"], snv)
    assert_ld_equal(got["ld"], ld, tol=1e-6)


@pytest.mark.parametrize("L,cov,dens,nsc,skip_mm,n_frac,seed", [
    (30000, 50, 0.01, 2, False, 0.0, 20260102),
    (30000, 50, 0.01, 1, True, 0.0, 20260102),
    (12000, 100, 0.05, 1, False, 0.002, 20260105),   # non-ACGT read bases -> nmask
    (12000, 100, 0.05, 1, True, 0.002, 20260105),    # same at M = 1
    (700, 30, 0.02, 3, False, 0.0, 7),
    (25000, 8, 0.01, 1, False, 0.0, 11),
    (9000, 400, 0.01, 1, True, 0.0, 3),              # > 255 candidates per position: 8 -> 32 bit widening, several chunks
    (9000, 400, 0.01, 1, False, 0.0, 3),             # same through the shared-memory counters (mid-run flush)
    (12000, 500, 0.05, 1, True, 0.0, 20260105),      # BASELINE configs[4] shaped (LD stress): 500x, 5 % SNVs, wide bit rows
])
def test_synthetic_parity_read_major(eng, null_lut, L, cov, dens, nsc, skip_mm, n_frac, seed):
    batch = synth.make_batch(L, cov, dens, seed, n_scaffolds=nsc, skip_mm=skip_mm, n_frac=n_frac)
    check_reads(eng, batch, null_lut)


def test_pileup_reads_stage(eng, null_lut):
    batch, _ = load_batch("G1")
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], do_linkage=False)
    L, M = exp["counts"].shape[:2]
    rd = reads.events_to_reads(batch)
    counts, nmask = eng.pileup_reads(rd, batch["pair_mm"], 0, L, M)
    assert np.array_equal(counts, exp["counts"]) and np.array_equal(nmask, exp["nmask"])


def test_many_mm_levels(eng, null_lut):
    """M > 32: the shared-memory level groups of K1r (two passes over the candidates)."""
    batch = synth.make_batch(6000, 60, 0.02, 5, skip_mm=False)
    rng = np.random.default_rng(0)
    batch["pair_mm"] = rng.integers(0, 50, len(batch["pair_mm"])).astype(batch["pair_mm"].dtype)
    check_reads(eng, batch, null_lut)


def test_short_and_split_segments(eng, null_lut):
    """Segments of every length 1..40 plus blocks longer than 256 (split by the encoder), starts clustered."""
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
    rd = reads.build_reads(starts, lens, pairs, np.concatenate(codes))
    assert rd["max_seg_len"] == 256
    ev = reads.reads_to_events(rd)
    n_pairs = 1500
    ev["pair_mm"] = rng.integers(0, 6, n_pairs).astype(np.uint8)
    ev["ref_codes"] = rng.integers(0, 4, L).astype(np.uint8)
    ev["splits"] = np.array([[0, L - 1]], dtype=np.int32)
    check_reads(eng, ev, null_lut, rd=rd)
    ev["pair_mm"][:] = 0
    check_reads(eng, ev, null_lut, rd=rd)


def test_empty_and_sparse(eng, null_lut):
    L = 3000
    ref = np.zeros(L, np.uint8)
    rd = reads.build_reads([], [], [], [])
    got = eng.profile_batch(dict(pair_mm=np.zeros(0, np.uint8)), ref, np.array([[0, L - 1]], np.int32), M=1, reads=rd,
                            want=("counts", "covT", "snv", "ld"))
    assert got["counts"].sum() == 0 and got["n_snv"] == 0 and got["n_ld"] == 0
    rd = reads.build_reads([2990], [10], [0], np.full(10, 2, np.uint8))     # one segment touching the last position
    counts, nmask = eng.pileup_reads(rd, np.zeros(1, np.uint8), 0, L, 1)
    assert counts[2990:, 0, 2].tolist() == [1] * 10 and counts.sum() == 10 and nmask.sum() == 0


def test_layout_violations_rejected(eng):
    from instrain_b200 import _cabi
    batch, _ = load_batch("G1")
    L, M = len(batch["ref_codes"]), int(batch["pair_mm"].max()) + 1
    rd = reads.events_to_reads(batch)
    bad = dict(rd); bad["seg_start"] = rd["seg_start"].copy(); bad["seg_start"][100:5000] = bad["seg_start"][100:5000][::-1]
    with pytest.raises(_cabi.IsbError) as ei:
        eng.pileup_reads(bad, batch["pair_mm"], 0, L, M)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    bad = dict(rd); bad["seg_word"] = rd["seg_word"].copy(); bad["seg_word"][50:] += 40000
    with pytest.raises(_cabi.IsbError):
        eng.pileup_reads(bad, batch["pair_mm"], 0, L, M)
    with pytest.raises(_cabi.IsbError) as ei:                                      # a segment beyond start + L
        eng.pileup_reads(rd, batch["pair_mm"], 0, L - 50000, M)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    with pytest.raises(_cabi.IsbError) as ei:                                      # mm value >= M
        eng.pileup_reads(rd, batch["pair_mm"], 0, L, M - 1)
    assert ei.value.code == _cabi.ISB_ERR_ARG


def test_read_major_equals_position_major_larger(eng, null_lut):
    """Both CUDA paths on a batch the oracle would take minutes for: identical tables."""
    for skip_mm in (True, False):
        batch = synth.make_batch(400000, 60, 0.01, 99, n_scaffolds=2, skip_mm=skip_mm)
        rd = reads.events_to_reads(batch, max_len=150 if skip_mm else 256, odd_blocks=skip_mm)
        M = int(batch["pair_mm"].max()) + 1
        want = ("counts", "nmask", "covT", "clonT", "site_flags", "snv", "ld")
        a = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want)
        b = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want, reads=rd)
        for k in ("counts", "nmask", "covT", "site_flags"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
        assert_snv_equal(a["snv"], b["snv"])
        assert_ld_equal(a["ld"], b["ld"], tol=0.0)


def test_device_generator_reads_equal_events(eng, null_lut):
    """The bench's device generator emits the SAME fragments as event columns and as aligned segments: both CUDA paths
    give identical tables on the device-resident data, and a host slice of the segments matches the oracle."""
    from instrain_b200 import synth as dsynth
    for skip_mm in (True, False):
        d = dsynth.generate(0, 50000, 3, 60, 0.01, 77, skip_mm=skip_mm, events=True, reads=True,
                            seg_words=22 if skip_mm else None)
        rd = d["reads"]
        assert rd["n_segs"] == 2 * d["pair_mm"].numel() and int(rd["seg_len"].min()) == 150
        M = int(d["pair_mm"].max().item()) + 1
        ev = dict(ref_pos=d["ref_pos"], base=d["base"], qual=d["qual"], read_id=d["read_id"], pair_mm=d["pair_mm"].cpu().numpy())
        ref, splits = d["ref_codes"].cpu().numpy(), d["splits"].cpu().numpy()
        want = ("counts", "nmask", "covT", "clonT", "site_flags", "snv", "ld")
        a = eng.profile_batch(ev, ref, splits, M=M, want=want)
        b = eng.profile_batch(ev, ref, splits, M=M, want=want, reads=rd)
        for k in ("counts", "nmask", "covT", "site_flags"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
        assert_snv_equal(a["snv"], b["snv"])
        assert_ld_equal(a["ld"], b["ld"], tol=0.0)
        assert a["n_snv"] > 100 and a["n_ld"] > 10
        hb = dsynth.to_host_batch(d, 1, 2)
        hr = dsynth.reads_to_host(d, 1, 2)
        assert np.array_equal(hb["pair_mm"], hr["pair_mm"])
        check_reads(eng, hb, null_lut, rd=hr["reads"])


@pytest.mark.parametrize("which", ["G1", "synth_mm", "synth_m1", "synth_n"])
def test_compact_transfer_format(eng, which, null_lut):
    """isb_profile_reads_compact (3 bits per aligned base, no word offsets; K0r rebuilds the stream on the device) gives
    the oracle's tables, from segments with odd block sizes, short pieces and non-ACGT bases alike."""
    if which == "G1":
        batch, _ = load_batch("G1")
        rd = reads.events_to_reads(batch)
    elif which == "synth_mm":
        batch = synth.make_batch(30000, 50, 0.01, 20260102, n_scaffolds=2, skip_mm=False)
        rd = reads.events_to_reads(batch, max_len=37, odd_blocks=True)          # many short pieces, two separators
    elif which == "synth_m1":
        batch = synth.make_batch(20000, 300, 0.02, 5, skip_mm=True)
        rd = reads.events_to_reads(batch, max_len=150)
    else:
        batch = synth.make_batch(12000, 100, 0.05, 20260105, skip_mm=False, n_frac=0.002)
        rd = reads.events_to_reads(batch)
    check_reads(eng, batch, null_lut, rd=reads.compact_reads(rd))


@pytest.mark.parametrize("which", ["G1", "synth_mm", "synth_m1", "synth_n"])
def test_delta_transfer_format(eng, which, null_lut):
    """isb_profile_reads_delta (event bits + one entry per base that differs from the reference; K0d rebuilds the stream
    on the device from the reference) gives the oracle's tables: real reads with indels, odd block sizes and short
    pieces, deep coverage, non-ACGT read bases and an N in the reference."""
    if which == "G1":
        batch, _ = load_batch("G1")
        rd = reads.events_to_reads(batch)
    elif which == "synth_mm":
        batch = synth.make_batch(30000, 50, 0.01, 20260102, n_scaffolds=2, skip_mm=False)
        rd = reads.events_to_reads(batch, max_len=37, odd_blocks=True)
    elif which == "synth_m1":
        batch = synth.make_batch(20000, 300, 0.02, 5, skip_mm=True)
        rd = reads.events_to_reads(batch, max_len=150)
    else:
        batch = synth.make_batch(12001, 100, 0.05, 20260105, skip_mm=False, n_frac=0.002)   # L not a multiple of 8
        rd = reads.events_to_reads(batch)
        batch["ref_codes"] = batch["ref_codes"].copy()
        batch["ref_codes"][100:140] = 4                                          # reference N run: every base is a mismatch
    check_reads(eng, batch, null_lut, rd=reads.delta_reads(rd, batch["ref_codes"]))


def test_delta_format_rejects_bad_entries(eng, null_lut):
    from instrain_b200 import _cabi
    batch = synth.make_batch(5000, 30, 0.01, 1, skip_mm=True)
    dl = reads.delta_reads(reads.events_to_reads(batch), batch["ref_codes"])
    bad = dict(dl); bad["mis_word"] = dl["mis_word"].copy(); bad["mis_word"][0] = 0xfffffff0     # beyond the stream
    with pytest.raises(_cabi.IsbError) as ei:
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=bad)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    bad = dict(dl); bad["n_units"] = dl["n_units"] - 3
    with pytest.raises(_cabi.IsbError) as ei:
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=bad)
    assert ei.value.code == _cabi.ISB_ERR_ORDER


def test_compact_format_rejects_inconsistent_units(eng, null_lut):
    from instrain_b200 import _cabi
    batch = synth.make_batch(5000, 30, 0.01, 1, skip_mm=True)
    rd = reads.compact_reads(reads.events_to_reads(batch))
    rd["n_units"] -= 3                                                            # table needs more units than given
    with pytest.raises(_cabi.IsbError) as ei:
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=rd)
    assert ei.value.code == _cabi.ISB_ERR_ORDER


def test_row_storage_regrow(null_lut):
    """The fused K3 front end sizes its bit-row storage by a guess and regrows it from the counted need: force the
    smallest guess (separate process: the guess is read once per process)."""
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from conftest import load_batch, load_lut, assert_ld_equal\n"
            "from oracle import restate\n"
            "from instrain_b200 import reads\n"
            "from instrain_b200.engine import Engine\n"
            "lut = load_lut(); b, _ = load_batch('G1')\n"
            "exp = restate.profile_events(b, b['ref_codes'], lut[0], lut[1], b['splits'])\n"
            "e = Engine(0, lut[0], lut[1])\n"
            "for _ in range(2):\n"
            "    got = e.profile_batch(b, b['ref_codes'], b['splits'], M=exp['counts'].shape[1], reads=reads.events_to_reads(b), want=('ld',))\n"
            "    assert_ld_equal(got['ld'], exp['ld'], tol=1e-9)\n"
            "print('regrow ok', len(got['ld']))\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ISB_K3_ROWS_INIT="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "regrow ok" in r.stdout, r.stdout + r.stderr


# ---- the fused read-major kernel (K1f: pileup + SNV call + linkage site rows in one launch; M = 1, no raw counts asked) ----

def check_fused(eng, batch, null_lut, tol=1e-9, rd=None, nmask=False, **kw):
    okw = {k: v for k, v in kw.items() if k != "skip_linkage"}
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"],
                                 do_linkage=not kw.get("skip_linkage", False), **okw)
    assert exp["counts"].shape[1] == 1
    if rd is None:
        rd = reads.events_to_reads(batch, kw.get("min_qual", 30))
    want = ("covT", "clonT", "clonTR", "site_flags", "snv", "ld") + (("nmask",) if nmask else ())
    n0 = eng.launch_count
    got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=rd, want=want, **kw)
    launches = eng.launch_count - n0
    assert_clontr_equal(got["clonTR"], exp["clonTR"])
    assert np.array_equal(got["covT"], exp["covT"])
    ok = ~np.isnan(exp["clonT"])
    assert np.array_equal(np.isnan(got["clonT"]), ~ok)
    assert np.array_equal(got["clonT"][ok].view(np.uint32), exp["clonT"][ok].view(np.uint32))
    assert np.array_equal(got["site_flags"], exp["site_flags"])
    if nmask:
        assert np.array_equal(got["nmask"], exp["nmask"])
    assert_snv_equal(got["snv"], exp["snv"])
    if not kw.get("skip_linkage"):
        assert_ld_equal(got["ld"], exp["ld"], tol=tol)
        assert got["n_sites"] == int((exp["site_flags"] & 0x10).astype(bool).sum())
    return got, exp, launches


@pytest.mark.parametrize("L,cov,dens,nsc,n_frac,seed,max_len", [
    (30000, 50, 0.01, 1, 0.0, 20260102, 256),
    (12000, 100, 0.05, 1, 0.002, 20260105, 256),     # non-ACGT read bases
    (700, 30, 0.02, 3, 0.0, 7, 256),                 # tiny scaffolds: tiles across scaffold / split boundaries
    (25000, 8, 0.01, 1, 0.0, 11, 256),               # coverage around min_cov
    (9000, 400, 0.01, 1, 0.0, 3, 150),               # > 255 candidates per position: plane flushes into the tile
    (12000, 500, 0.05, 1, 0.0, 20260105, 256),       # BASELINE configs[4] shaped: wide bit rows (overflow region)
    (20001, 60, 0.02, 2, 0.0, 5, 37),                # L not a multiple of 8, many short pieces
    (3000, 2500, 0.02, 1, 0.0, 9, 256),              # several staging chunks per tile: global candidate search for the sites
])
def test_fused_parity(eng, null_lut, L, cov, dens, nsc, n_frac, seed, max_len):
    batch = synth.make_batch(L, cov, dens, seed, n_scaffolds=nsc, skip_mm=True, n_frac=n_frac)
    rd = reads.events_to_reads(batch, max_len=max_len, odd_blocks=max_len != 256)
    _, _, launches = check_fused(eng, batch, null_lut, rd=rd, nmask=n_frac > 0)
    assert launches <= 8                            # tile bounds, K1f, enumeration, statistics, self edges (+ N events, threshold table)


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_fused_golden_reads_set_mode(eng, which, null_lut):
    """The bundled BAM's reads (indels, overlap-quirk double entries -> self edges) with every pair at level 0."""
    batch, _ = load_batch(which)
    batch = dict(batch)
    batch["pair_mm"] = np.zeros_like(batch["pair_mm"])
    check_fused(eng, batch, null_lut)


def test_fused_other_settings(eng, null_lut):
    batch = synth.make_batch(20000, 40, 0.03, 21, skip_mm=True)
    for kw in (dict(min_cov=1, min_freq=0.10, min_snp=5), dict(min_cov=10, min_freq=0.02, min_snp=40),
               dict(min_cov=0), dict(skip_linkage=True)):
        check_fused(eng, batch, null_lut, **kw)


def test_fused_short_and_split_segments(eng, null_lut):
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
    rd = reads.build_reads(starts, lens, pairs, np.concatenate(codes))
    ev = reads.reads_to_events(rd)
    ev["pair_mm"] = np.zeros(1500, np.uint8)
    ev["ref_codes"] = rng.integers(0, 4, L).astype(np.uint8)
    ev["splits"] = np.array([[0, 1999], [2000, L - 1]], dtype=np.int32)
    check_fused(eng, ev, null_lut, rd=rd, nmask=True)


def test_fused_empty_and_errors(eng, null_lut):
    from instrain_b200 import _cabi
    L = 3000
    ref = np.zeros(L, np.uint8)
    rd = reads.build_reads([], [], [], [])
    got = eng.profile_batch(dict(pair_mm=np.zeros(0, np.uint8)), ref, np.array([[0, L - 1]], np.int32), M=1, reads=rd)
    assert got["covT"].sum() == 0 and got["n_snv"] == 0 and got["n_ld"] == 0 and got["n_sites"] == 0
    batch = synth.make_batch(20000, 30, 0.01, 1, skip_mm=True)
    rd = reads.events_to_reads(batch)
    bad = dict(rd); bad["seg_start"] = rd["seg_start"].copy(); bad["seg_start"][100:900] = bad["seg_start"][100:900][::-1]
    with pytest.raises(_cabi.IsbError) as ei:
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=bad)
    assert ei.value.code == _cabi.ISB_ERR_ORDER
    bad = dict(rd); bad["seg_word"] = rd["seg_word"].copy(); bad["seg_word"][50:] += 1 << 30
    with pytest.raises(_cabi.IsbError):
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=bad)
    bad = dict(rd); bad["seg_pair"] = rd["seg_pair"].copy(); bad["seg_pair"][::3] = 1 << 29      # pair ids beyond n_pairs
    with pytest.raises(_cabi.IsbError):
        eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, reads=bad)
    check_fused(eng, batch, null_lut, rd=rd)                                   # the context is usable afterwards


def test_fused_equals_unfused_larger(eng, null_lut):
    """Fused kernel against K1 -> K2 -> K3 on event columns at a size the oracle would take minutes for, through the
    device generator's segments (uniform 150-base reads) and through the transfer formats."""
    batch = synth.make_batch(600000, 80, 0.01, 99, n_scaffolds=2, skip_mm=True)
    rd = reads.events_to_reads(batch, max_len=150, odd_blocks=True)
    want = ("covT", "clonT", "site_flags", "snv", "ld")
    a = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, want=want)
    for fmt in (rd, reads.delta_reads(rd, batch["ref_codes"]), reads.compact_reads(rd)):
        b = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=1, want=want, reads=fmt)
        for k in ("covT", "site_flags"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
        assert_snv_equal(a["snv"], b["snv"])
        assert_ld_equal(a["ld"], b["ld"], tol=0.0)
        assert a["n_sites"] == b["n_sites"] and a["n_site_pairs"] == b["n_site_pairs"]


def test_fused_scratch_regrow(null_lut):
    """Site slots, bit-row storage and the pair list start from a forced tiny guess and are regrown from the counted need
    (separate process: the guess is read once per process)."""
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from conftest import load_lut, assert_ld_equal, assert_snv_equal\n"
            "from oracle import restate, synth\n"
            "from instrain_b200 import reads\n"
            "from instrain_b200.engine import Engine\n"
            "lut = load_lut(); b = synth.make_batch(12000, 500, 0.05, 20260105, skip_mm=True)\n"
            "exp = restate.profile_events(b, b['ref_codes'], lut[0], lut[1], b['splits'])\n"
            "e = Engine(0, lut[0], lut[1])\n"
            "for _ in range(2):\n"
            "    got = e.profile_batch(b, b['ref_codes'], b['splits'], M=1, reads=reads.events_to_reads(b), want=('snv', 'ld'))\n"
            "    assert_snv_equal(got['snv'], exp['snv']); assert_ld_equal(got['ld'], exp['ld'], tol=1e-9)\n"
            "print('regrow ok', len(got['ld']))\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ISB_K1F_SITES_INIT="16", ISB_K1F_QUEUE_INIT="64")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "regrow ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("skip_mm", [True, False])
def test_redrawn_outputs_all_layouts(eng, null_lut, skip_mm):
    """clonTR and the normalized linkage columns (re-drawn quantities: unseeded in the reference, counter-based here) with a
    non-default seed and rarefied coverage: bit-exact against the oracle's restatement of the same draws in every input
    layout, reproducible, and different under another seed."""
    from instrain_b200 import cols
    batch = synth.make_batch(30000, 70, 0.03, 31, skip_mm=skip_mm)
    rd = reads.events_to_reads(batch)
    cd = cols.reads_to_cols(rd, len(batch["ref_codes"]))
    M = int(batch["pair_mm"].max()) + 1
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"], rarefied_coverage=30,
                                 seed=20260103, min_snp=12)
    assert (~np.isnan(exp["clonTR"])).sum() > 20000 and ((exp["clonTR"] < 1) & ~np.isnan(exp["clonTR"])).sum() > 500
    assert (~np.isnan(exp["ld"]["r2_normalized"])).sum() > 100
    want = ("covT", "clonT", "clonTR", "snv", "ld")
    outs = []
    for layout in (dict(reads=rd), dict(cols=cd), {}, dict(reads=reads.delta_reads(rd, batch["ref_codes"]))):
        got = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want, rarefied_coverage=30, seed=20260103,
                                min_snp=12, **layout)
        assert_clontr_equal(got["clonTR"], exp["clonTR"])
        assert_ld_equal(got["ld"], exp["ld"], tol=1e-12)
        outs.append(got)
    other = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want, rarefied_coverage=30, seed=1, min_snp=12, reads=rd)
    drawn = ~np.isnan(exp["clonTR"]) & (exp["clonTR"] < 1)
    assert (other["clonTR"][drawn] != exp["clonTR"][drawn]).mean() > 0.3
    assert np.array_equal(np.isnan(other["clonTR"]), np.isnan(exp["clonTR"]))
    none = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=want, rarefied_coverage=0, reads=rd)
    assert np.isnan(none["clonTR"]).all()


def test_async_steps_on_device_buffers(eng, null_lut):
    """The streaming mode bench.py's timed loops use: inputs and outputs resident on the device, three steps enqueued with
    ISB_NO_SYNC (no host round trip per call), the row counts fetched through isb_row_counts_async, one isb_synchronize at the
    end -- tables identical to the synchronous call and to the oracle."""
    import ctypes as C
    import torch
    from instrain_b200 import _cabi
    batch = synth.make_batch(40000, 60, 0.02, 4242, skip_mm=True)
    exp = restate.profile_events(batch, batch["ref_codes"], null_lut[0], null_lut[1], batch["splits"])
    rd = reads.events_to_reads(batch)
    L = len(batch["ref_codes"])
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: t(rd[k]) for k in ("seg_start", "seg_len", "seg_pair", "seg_word", "words")}
    d["seg_len"] = t(np.asarray(rd["seg_len"]).view(np.int16))
    pair_mm, ref, splits = t(np.zeros(len(batch["pair_mm"]), np.uint8)), t(batch["ref_codes"]), t(batch["splits"].astype(np.int32))
    p = _cabi.ptr
    bt = _cabi.IsbReadsBatch(int(rd["n_segs"]), p(d["seg_start"]), p(d["seg_len"]), p(d["seg_pair"]), p(d["seg_word"]),
                             int(rd["n_words"]), p(d["words"]), int(rd["max_seg_len"]), 0, 0, None, None, pair_mm.numel(), p(pair_mm),
                             0, L, p(ref), len(batch["splits"]), p(splits), 1, 0)
    snv_cap, ld_cap = len(exp["snv"]) + 64, len(exp["ld"]) + 64
    out = dict(covT=torch.empty((L, 1), dtype=torch.int32, device=dev), clonT=torch.empty((L, 1), dtype=torch.float32, device=dev),
               flags=torch.empty(L, dtype=torch.uint8, device=dev), snv=torch.zeros(snv_cap * 32, dtype=torch.uint8, device=dev),
               ld=torch.zeros(ld_cap * 64, dtype=torch.uint8, device=dev))
    res = _cabi.IsbResult(None, None, p(out["covT"]), p(out["clonT"]), p(out["flags"]), p(out["snv"]), snv_cap, p(out["ld"]), ld_cap,
                          0, 0, 0, 0, None)
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    lib, ctx = eng.lib, eng.ctx
    sync_prm = _cabi.IsbParams(5, 20, 30, 0, 0.05, 0, 0, 0)
    assert lib.isb_profile_reads(ctx, C.byref(bt), C.byref(sync_prm), C.byref(res)) == 0      # warm-up: scratch sized, synchronous
    n_sync = (int(res.n_snv), int(res.n_ld), int(res.n_sites), int(res.n_site_pairs))
    assert n_sync[:2] == (len(exp["snv"]), len(exp["ld"]))
    out["snv"].zero_(); out["ld"].zero_(); out["covT"].zero_()
    torch.cuda.synchronize()
    prm = _cabi.IsbParams(5, 20, 30, _cabi.ISB_NO_SYNC, 0.05, 0, 0, 0)
    for _ in range(3):
        assert lib.isb_profile_reads(ctx, C.byref(bt), C.byref(prm), C.byref(res)) == 0
        assert int(res.n_snv) == -1                                                            # not known without a round trip
    assert lib.isb_row_counts_async(ctx, p(counts)) == 0
    assert lib.isb_synchronize(ctx) == 0
    assert tuple(counts.tolist()) == n_sync
    snv = out["snv"].cpu().numpy().view(_cabi.SNV_DT)[:n_sync[0]]
    ld = out["ld"].cpu().numpy().view(_cabi.LD_DT)[:n_sync[1]]
    assert_snv_equal(snv.copy(), exp["snv"])
    assert_ld_equal(ld.copy(), exp["ld"], tol=1e-9)
    assert np.array_equal(out["covT"].cpu().numpy(), exp["covT"]) and np.array_equal(out["flags"].cpu().numpy(), exp["site_flags"])
