"""Second caller of the pileup primitive: base counts at given positions of one scaffold.

Drop-in for inStrain.polymorpher.extract_SNVS_from_bam (inStrain/polymorpher.py:275-316, used by `compare --bams` to pool
SNVs): the reference re-runs the identical pysam pileup over [min(positions)-1, max(positions)] and sums
get_base_counts_mm over all mm levels (get_pooling_counts, :312-316).  Here: host packer -> K1 with the mm dimension
collapsed (M = 1) -> gather.  The reference piles up only the reads its index fetch returns for that region; under the
pinned restatement of htslib's pileup the mate-overlap tweak gives the same counts either way (a mate can only change
qualities inside the overlap of the two reads, and such a position inside the region makes both mates overlap it):
tests/test_polymorpher_region.py holds the region-limited emulation against the whole-scaffold one at > 400 positions,
single-position regions included, so packing the whole scaffold is the reference's semantics.
"""
import numpy as np

from .engine import Engine
from .packer import BamPacker


def extract_SNVS_from_bam(bam_loc, R2M, positions, scaffold, engine=None, device=0, **kwargs):
    """Returns {position: np.array([A, C, T, G])} like the reference (zeros where nothing is counted)."""
    positions = [int(p) for p in positions]
    if len(positions) == 0:
        return {}
    own = engine is None
    if own:
        engine = Engine(device)
    try:
        with BamPacker(bam_loc) as bp:
            if scaffold not in bp.ref_names:
                raise ValueError("scaffold %s is not in the .bam file %s" % (scaffold, bam_loc))
            want = bp.ref_names.index(scaffold)
            ev = None
            from .packer import find_bai, read_bai
            bai = find_bai(bam_loc)
            if bai is not None:                                  # indexed BAM: go straight to the scaffold's first record
                first = read_bai(bai)[want]
                if first is None:
                    return {p: np.zeros(4, dtype=int) for p in set(positions)}
                bp.seek(first)
            while True:
                tid = bp.peek_tid()
                if tid < 0 or tid > want:
                    break
                if tid == want:
                    ev = bp.pack_scaffold(tid, set(R2M.keys()) if isinstance(R2M, dict) else R2M)
                    break
                bp.pack_scaffold(tid, {})                        # no index: the sequential reader skips earlier scaffolds
            L = bp.ref_lens[want]
        if ev is None or len(ev["ref_pos"]) == 0:
            return {p: np.zeros(4, dtype=int) for p in set(positions)}
        lo, hi = max(min(positions) - 1, 0), min(max(positions), L - 1)
        sel = slice(int(np.searchsorted(ev["ref_pos"], lo)), int(np.searchsorted(ev["ref_pos"], hi + 1)))
        sub = dict(ref_pos=np.ascontiguousarray(ev["ref_pos"][sel]), base=np.ascontiguousarray(ev["base"][sel]),
                   qual=np.ascontiguousarray(ev["qual"][sel]), read_id=np.ascontiguousarray(ev["read_id"][sel]), pair_mm=None)
        counts, _ = engine.pileup_counts(sub, lo, hi - lo + 1, 1)
        out = {}
        for p in set(positions):
            out[p] = counts[p - lo, 0].astype(int) if lo <= p <= hi else np.zeros(4, dtype=int)
        return out
    finally:
        if own:
            engine.close()
