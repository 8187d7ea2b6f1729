"""Host-side encoders of the transfer formats on the CPU: the reference-delta format (instrain_b200.reads.delta_reads)
decoded by a numpy restatement of K0d gives back the canonical nibble stream of the same segments (no GPU needed)."""
import numpy as np
import pytest

from conftest import load_batch
from oracle import synth
from instrain_b200 import reads


def canonical_stream(rd):
    """[leading zero word][data words + one zero word per segment] -- what K0r / K0d rebuild on the device."""
    s = rd["seg_start"].astype(np.int64)
    nw = ((s & 7) + rd["seg_len"].astype(np.int64) + 7) // 8
    n = len(s)
    seg_word = 1 + (np.cumsum(nw) - nw) + np.arange(n)
    n_words = (1 + int(nw.sum()) + n + 3) // 4 * 4
    words = np.zeros(n_words, np.uint32)
    k = np.arange(int(nw.sum())) - np.repeat(np.cumsum(nw) - nw, nw)
    words[np.repeat(seg_word, nw) + k] = rd["words"][np.repeat(np.asarray(rd["seg_word"], np.int64), nw) + k]
    return seg_word, n_words, words


@pytest.mark.parametrize("which", ["G1", "short_odd", "n_bases", "ragged"])
def test_delta_round_trip(which):
    if which == "G1":
        b, _ = load_batch("G1")
        rd = reads.events_to_reads(b)
    elif which == "short_odd":
        b = synth.make_batch(30000, 50, 0.01, 1, n_scaffolds=2, skip_mm=False)
        rd = reads.events_to_reads(b, max_len=37, odd_blocks=True)
    elif which == "n_bases":
        b = synth.make_batch(12000, 100, 0.05, 5, skip_mm=True, n_frac=0.002)
        rd = reads.events_to_reads(b)
        b["ref_codes"] = b["ref_codes"].copy()
        b["ref_codes"][500:520] = 4
    else:
        b = synth.make_batch(1001, 20, 0.02, 3, skip_mm=True)
        rd = reads.events_to_reads(b)
    dl = reads.delta_reads(rd, b["ref_codes"])
    assert dl["n_units"] == reads.compact_reads(rd)["n_units"] and np.array_equal(dl["pass"], reads.compact_reads(rd)["pass"])
    sw, n_words, words = reads.delta_to_words(dl, b["ref_codes"])
    sw2, n_words2, words2 = canonical_stream(rd)
    assert n_words == n_words2 and np.array_equal(sw, sw2) and np.array_equal(words, words2)
    assert len(dl["mis_word"]) < 0.2 * int(rd["seg_len"].sum())       # a delta: far fewer entries than bases


@pytest.mark.parametrize("which", ["G2", "short_odd", "ragged"])
def test_delta_host_encoder_equals_numpy(which):
    """isb_reads_delta_host (C++, what the host pipeline runs) and the numpy encoder agree: same event bits, same set of
    mismatch entries; its output decodes to the canonical stream."""
    if which == "G2":
        b, _ = load_batch("G2")
        rd = reads.events_to_reads(b)
    elif which == "short_odd":
        b = synth.make_batch(20000, 40, 0.02, 9, n_scaffolds=2, skip_mm=False, n_frac=0.001)
        rd = reads.events_to_reads(b, max_len=37, odd_blocks=True)
        b["ref_codes"] = b["ref_codes"].copy()
        b["ref_codes"][1000:1040] = 4
    else:
        b = synth.make_batch(1001, 20, 0.02, 3, skip_mm=True)
        rd = reads.events_to_reads(b)
    a = reads.delta_reads(rd, b["ref_codes"])
    h = reads.delta_reads_host(rd, b["ref_codes"])
    assert h["n_units"] == a["n_units"] and np.array_equal(h["pass"], a["pass"])
    key = lambda d: np.sort(d["mis_word"].astype(np.int64) * 256 + d["mis_code"])
    assert np.array_equal(key(h), key(a))
    sw, n_words, words = reads.delta_to_words(h, b["ref_codes"])
    sw2, n_words2, words2 = canonical_stream(rd)
    assert n_words == n_words2 and np.array_equal(words, words2)
    bad = dict(rd)
    bad["seg_start"] = rd["seg_start"][::-1].copy()
    with pytest.raises(ValueError):
        reads.delta_reads_host(bad, b["ref_codes"])
