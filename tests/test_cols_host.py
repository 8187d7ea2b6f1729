"""Column-word layout on the CPU: the C++ host conversion (isb_cols_from_reads_host) against an independent numpy
restatement, and the round trip column words -> events == read-major segments -> events (no GPU needed)."""
import numpy as np
import pytest

from conftest import load_batch
from oracle import synth
from instrain_b200 import cols, reads


def _check(rd, L):
    a = cols.reads_to_cols(rd, L)
    n = cols.reads_to_cols_numpy(rd, L)
    assert a["n_chunks"] == n["n_chunks"] and np.array_equal(a["grp_off"], n["grp_off"])
    assert np.array_equal(a["words"], n["words"]) and np.array_equal(a["ids"], n["ids"])
    assert a["grp_off"][0] == 0 and np.all(np.diff(a["grp_off"]) >= 0) and a["grp_off"][-1] == a["n_chunks"]
    ev, ev2 = cols.cols_to_events(a, L), reads.reads_to_events(rd)
    for k in ("ref_pos", "base", "read_id"):
        assert np.array_equal(ev[k], ev2[k]), k
    return a


@pytest.mark.parametrize("L,cov,dens,nsc,skip_mm,n_frac,seed", [
    (30000, 50, 0.01, 2, False, 0.0, 1),
    (700, 30, 0.02, 3, False, 0.002, 7),
    (1001, 5, 0.01, 1, True, 0.0, 4),
    (9000, 300, 0.01, 1, True, 0.0, 3),
])
def test_host_conversion_synthetic(L, cov, dens, nsc, skip_mm, n_frac, seed):
    b = synth.make_batch(L, cov, dens, seed, n_scaffolds=nsc, skip_mm=skip_mm, n_frac=n_frac)
    _check(reads.events_to_reads(b), len(b["ref_codes"]))


def test_host_conversion_golden_and_column_order():
    """Real reads (indels, clipped ends); inside a column the words keep the table order of their segments."""
    b, _ = load_batch("G1")
    L = len(b["ref_codes"])
    rd = reads.events_to_reads(b)
    a = _check(rd, L)
    # column order: ids of a column, read in slot order, are the pair ids of its covering segments in table order
    s = rd["seg_start"].astype(np.int64)
    e = s + rd["seg_len"].astype(np.int64) - 1
    w = a["ids"].reshape(-1, cols.LANES, cols.UNIT)
    rng = np.random.default_rng(0)
    for c in rng.integers(0, (L + 7) // 8, 200):
        g, lane = c // cols.LANES, c % cols.LANES
        col_ids = w[a["grp_off"][g]:a["grp_off"][g + 1], lane, :].reshape(-1)
        col_ids = col_ids[col_ids >= 0]
        cover = np.nonzero(((s >> 3) <= c) & ((e >> 3) >= c))[0]
        assert np.array_equal(col_ids, rd["seg_pair"][cover])


def test_host_conversion_edge_cases():
    L = 3000
    a = cols.reads_to_cols(reads.build_reads([], [], [], []), L)
    assert a["n_chunks"] == 0 and len(a["grp_off"]) == (L + cols.GROUP - 1) // cols.GROUP + 1
    rd = reads.build_reads([2990], [10], [0], np.full(10, 2, np.uint8))
    a = _check(rd, L)
    assert a["n_chunks"] == 1                                   # two column words of one group, one slot each
    bad = dict(rd); bad["seg_start"] = np.array([2995], np.int32)            # runs past L
    with pytest.raises(ValueError):
        cols.reads_to_cols(bad, L)


def _random_reads(rng, L, n_seg, max_len):
    starts, lens, pairs, codes = [], [], [], []
    for i in range(n_seg):
        n = int(rng.integers(1, max_len + 1))
        if n > L:
            n = L
        s = int(rng.integers(0, L - n + 1))
        starts.append(s); lens.append(n); pairs.append(int(rng.integers(0, max(1, n_seg // 2))))
        c = rng.integers(0, 5, n).astype(np.uint8)
        c[rng.random(n) < 0.25] = reads.NO_EVENT
        codes.append(c)
    return reads.build_reads(starts, lens, pairs, np.concatenate(codes) if codes else [], odd_blocks=bool(rng.integers(0, 2)))


@pytest.mark.parametrize("seed", range(12))
def test_layout_round_trips_on_random_segments(seed):
    """Property check over random segment sets (every length 1..700 incl. blocks split at 256, clustered and sparse starts,
    L of any residue mod 8 / 64, duplicate pairs, non-ACGT bases): column words and the reference-delta format both decode
    to exactly the events of the read-major batch, the C++ host routines equal the numpy restatements."""
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.integers(9, 3000))
    rd = _random_reads(rng, L, int(rng.integers(1, 400)), int(rng.choice([3, 40, 150, 256, 700])))
    _check(rd, L)
    ref = rng.integers(0, 5, L).astype(np.uint8)
    a, h = reads.delta_reads(rd, ref), reads.delta_reads_host(rd, ref)
    assert np.array_equal(a["pass"], h["pass"])
    key = lambda d: np.sort(d["mis_word"].astype(np.int64) * 256 + d["mis_code"])
    assert np.array_equal(key(a), key(h))
    seg_word, n_words, words = reads.delta_to_words(h, ref)
    ev = reads.reads_to_events(dict(rd, seg_word=seg_word, n_words=n_words, words=words))
    ev0 = reads.reads_to_events(rd)
    for k in ("ref_pos", "base", "read_id"):
        assert np.array_equal(ev[k], ev0[k]), k
