"""inStrain.polymorpher.extract_SNVS_from_bam (polymorpher.py:275-316) piles up only the region
[min(positions) - 1, max(positions) + 1) -- pysam fetches just the reads that overlap it, so htslib's mate-overlap
handling (overlap_push / tweak_overlap_quality) sees fewer reads than in a whole-scaffold pass.  instrain_b200.polymorpher
packs the whole scaffold and gathers the positions.  This test pins that the two are the same thing under the repo's
restatement of htslib's pileup (oracle/pileup_emul.py, itself pinned on the reference's stored tables): the emulator
restricted to the reads a region fetch returns gives, at every asked position, exactly the counts of the whole-scaffold
emulation -- for multi-position regions and for the tightest possible ones (a single position).  (A mate can only change
a base quality inside the overlap of the two reads, and a position of that overlap inside the region makes BOTH mates
overlap the region.)"""
import json
import os

import numpy as np

from conftest import GOLDEN
from oracle import bamio, pileup_emul as pe


def _counts(ev, L):
    ok = (ev["qual"] >= 30) & (ev["base"] < 4)
    out = np.zeros((L, 4), dtype=np.int64)
    np.add.at(out, (ev["ref_pos"][ok], ev["base"][ok]), 1)
    return out


def test_region_limited_pileup_equals_whole_scaffold_pileup():
    refs, reads = bamio.read_bam(os.path.join(GOLDEN, "c1_G1_subset.bam"))
    r2m = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    by = {}
    for r in reads:
        if r.tid >= 0:
            by.setdefault(r.tid, []).append(r)
    rng = np.random.default_rng(7)
    n_checked = n_nonzero = 0
    for tid, rs in sorted(by.items(), key=lambda kv: -len(kv[1]))[:3]:
        name, L = refs[tid][0], refs[tid][1]
        full = _counts(pe.scaffold_events(rs, r2m[name]), L)
        starts = np.array([r.pos for r in rs])
        ends = np.array([r.pos + pe.ref_len(r.cigar) for r in rs])
        regions = [sorted(int(x) for x in rng.integers(0, L, int(rng.integers(2, 6)))) for _ in range(10)]
        regions += [[int(p)] for p in rng.integers(0, L, 120)]
        for positions in regions:
            lo, hi = max(min(positions) - 1, 0), max(positions) + 1            # start / stop of the reference's pileup call
            sub = [r for r, s, e in zip(rs, starts, ends) if s < hi and e > lo]  # what the index fetch hands the pileup
            part = _counts(pe.scaffold_events(sub, r2m[name]), L)
            for p in positions:
                assert np.array_equal(part[p], full[p]), (name, p, positions)
                n_checked += 1
                n_nonzero += int(full[p].sum() > 0)
    assert n_checked > 400 and n_nonzero > 300
