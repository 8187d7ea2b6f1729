"""htslib-faithful pileup emulation (oracle / test infrastructure only).

The reference obtains its pileup from a third-party dependency that is NOT in /root/reference:
pysam (setup.py:24, `pysam>=0.15`; the stored goldens were produced with pysam 0.16.0.1 which
bundles htslib 1.10.2).  Call site:

    samfile.pileup(scaffold, truncate=True, max_depth=100000, stepper='nofilter', compute_baq=True,
                   ignore_orphans=True, ignore_overlaps=True, min_base_quality=30,
                   start=start, stop=end+1)                 inStrain/profile/profile_utilities.py:150-153

This module restates the published htslib algorithm for exactly that configuration
(bam_plp `overlap_push` + `tweak_overlap_quality` + `cigar_iref2iseq_set/next`, and pysam's
per-column base-quality filter), see SURVEY.md section 8c / Appendix A.  It is anchored on the reference's
golden tables: feeding these columns to the reference's own `process_bam_sites`/`calculate_ld`
reproduces raw_snp_table / raw_linkage_table of the `forRC.IS` fixtures row for row
(oracle/ref_harness.py, tests/test_oracle_golden.py).
"""
import numpy as np

# CIGAR op codes (BAM): M I D N S H P = X
_M, _I, _D, _N, _S, _H, _P, _EQ, _X = range(9)
_MATCH = (_M, _EQ, _X)

# inStrain base order is A, C, T, G (inStrain/profile/profile_utilities.py:34-35)
BASE_CODE = {"A": 0, "C": 1, "T": 2, "G": 3}
BASE_OTHER = 4          # in-alignment base that is not A/C/G/T (e.g. N): KeyError path at :278-285
CODE_BASE = "ACTG"

MIN_BASE_QUAL = 30      # min_base_quality=30 at profile_utilities.py:152

FLAG_PROPER = 0x2
FLAG_UNMAP = 0x4
FLAG_MUNMAP = 0x8


def ref_len(cigar):
    return sum(n for op, n in cigar if op in (_M, _D, _N, _EQ, _X))


class _Walker:
    """State of htslib's cigar_iref2iseq_{set,next} for one read."""
    __slots__ = ("cigar", "ci", "icig", "iseq", "iref")

    def __init__(self, cigar):
        self.cigar = cigar
        self.ci = 0
        self.icig = 0
        self.iseq = 0
        self.iref = 0

    def set(self, pos):
        """Position on the first M/=/X base at reference offset >= pos. Returns 0 or -1."""
        if pos < 0:
            return -1
        self.icig = self.iseq = self.iref = 0
        cig = self.cigar
        while self.ci < len(cig):
            op, n = cig[self.ci]
            if op == _S or op == _I:
                self.ci += 1
                self.iseq += n
                self.icig = 0
            elif op == _H or op == _P:
                self.ci += 1
                self.icig = 0
            elif op in _MATCH:
                pos -= n
                if pos < 0:
                    self.icig = n + pos
                    self.iseq += self.icig
                    self.iref += self.icig
                    return 0
                self.ci += 1
                self.iseq += n
                self.icig = 0
                self.iref += n
            elif op == _D or op == _N:
                pos -= n
                if pos < 0:
                    pos = 0
                self.ci += 1
                self.iref += n
                self.icig = 0
            else:
                return -2
        self.iseq = -1
        return -1

    def next(self):
        """Advance to the next M/=/X base (with htslib 1.10's in-block counter behaviour)."""
        cig = self.cigar
        while self.ci < len(cig):
            op, n = cig[self.ci]
            if op in _MATCH:
                if self.icig >= n - 1:
                    self.icig = 0
                    self.ci += 1
                    continue
                self.iseq += 1
                self.icig += 1
                self.iref += 1
                return 0
            if op == _D or op == _N:
                self.ci += 1
                self.iref += n
                self.icig = 0
            elif op == _I or op == _S:
                self.ci += 1
                self.iseq += n
                self.icig = 0
            elif op == _H or op == _P:
                self.ci += 1
                self.icig = 0
            else:
                return -2
        self.iseq = -1
        self.iref = -1
        return -1


def tweak_overlap_quality(a_pos, a_cigar, a_seq, a_qual, b_pos, b_cigar, b_seq, b_qual):
    """htslib tweak_overlap_quality(a, b); a arrived first. Modifies a_qual / b_qual in place."""
    wa, wb = _Walker(a_cigar), _Walker(b_cigar)
    iref = b_pos
    a_ret = wa.set(iref - a_pos)
    if a_ret < 0:
        return
    b_ret = wb.set(iref - b_pos)
    if b_ret < 0:
        return
    while True:
        while a_ret >= 0 and wa.iref >= 0 and wa.iref < iref - a_pos:
            a_ret = wa.next()
        if a_ret < 0:
            return
        if iref < wa.iref + a_pos:
            iref = wa.iref + a_pos
        while b_ret >= 0 and wb.iref >= 0 and wb.iref < iref - b_pos:
            b_ret = wb.next()
        if b_ret < 0:
            return
        if iref < wb.iref + b_pos:
            iref = wb.iref + b_pos
        iref += 1
        if wa.iref + a_pos != wb.iref + b_pos:
            continue
        ia, ib = wa.iseq, wb.iseq
        if a_seq[ia] == b_seq[ib]:
            q = int(a_qual[ia]) + int(b_qual[ib])
            a_qual[ia] = 200 if q > 200 else q
            b_qual[ib] = 0
        elif a_qual[ia] >= b_qual[ib]:
            a_qual[ia] = int(0.8 * int(a_qual[ia]))
            b_qual[ib] = 0
        else:
            b_qual[ib] = int(0.8 * int(b_qual[ib]))
            a_qual[ia] = 0


def tweak_scaffold(reads):
    """Apply bam_plp overlap handling (ignore_overlaps=True) to the mapped reads of ONE scaffold,
    given in file order.  Returns a list of post-tweak quality arrays (copies), one per read."""
    quals = [r.qual.copy() for r in reads]
    pending = {}
    for i, r in enumerate(reads):
        if r.flag & FLAG_MUNMAP or not (r.flag & FLAG_PROPER):
            continue
        end = r.pos + ref_len(r.cigar)
        if (r.mtid >= 0 and r.tid != r.mtid) or (abs(r.isize) >= 2 * len(r.seq) and r.mpos >= end):
            continue
        j = pending.get(r.name)
        if j is None:
            if r.mpos >= r.pos:
                pending[r.name] = i
        else:
            a = reads[j]
            tweak_overlap_quality(a.pos, a.cigar, a.seq, quals[j], r.pos, r.cigar, r.seq, quals[i])
            del pending[r.name]
    return quals


def read_events(read, qual):
    """(ref_pos[], qpos[]) of the M/=/X bases of one read (true alignment, no quirk)."""
    pos, q = read.pos, 0
    rp, qp = [], []
    for op, n in read.cigar:
        if op in _MATCH:
            rp.append(np.arange(pos, pos + n, dtype=np.int64))
            qp.append(np.arange(q, q + n, dtype=np.int64))
            pos += n
            q += n
        elif op == _I or op == _S:
            q += n
        elif op == _D or op == _N:
            pos += n
    if not rp:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(rp), np.concatenate(qp)


_LUT = np.full(256, BASE_OTHER, dtype=np.uint8)
for _b, _c in BASE_CODE.items():
    _LUT[ord(_b)] = _c


def scaffold_events(reads, r2m, tweak=True):
    """Columnar events of one scaffold, BAM (file) order.

    reads : mapped reads of the scaffold in file order
    r2m   : dict name -> mm  (the hot path's sR2M[scaffold]; profile_utilities.py:268-286), or a set
    Returns dict with
      ref_pos int32[n], base uint8[n] (0..3 = A,C,T,G ; 4 = other), qual uint8[n] (post-tweak),
      read_id int32[n] (pair index into pair_mm / names), pair_mm int32[n_pairs], names list[str]
    Only reads whose name is in r2m are emitted (others can never be counted: :277-283).
    Events below the base-quality threshold are kept -- filtering is the kernel's job.
    """
    reads = [r for r in reads if not (r.flag & FLAG_UNMAP) and r.tid >= 0]
    quals = tweak_scaffold(reads) if tweak else [r.qual for r in reads]
    is_dict = isinstance(r2m, dict)
    names, name2id, mm = [], {}, []
    e_pos, e_base, e_qual, e_rid = [], [], [], []
    for r, q in zip(reads, quals):
        if r.name not in r2m:
            continue
        pid = name2id.get(r.name)
        if pid is None:
            pid = name2id[r.name] = len(names)
            names.append(r.name)
            mm.append(int(r2m[r.name]) if is_dict else 0)
        rp, qp = read_events(r, q)
        if len(rp) == 0:
            continue
        sb = np.frombuffer(r.seq.encode(), dtype=np.uint8)
        e_pos.append(rp.astype(np.int32))
        e_base.append(_LUT[sb[qp]])
        e_qual.append(q[qp])
        e_rid.append(np.full(len(rp), pid, dtype=np.int32))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return dict(ref_pos=cat(e_pos, np.int32), base=cat(e_base, np.uint8), qual=cat(e_qual, np.uint8),
                read_id=cat(e_rid, np.int32), pair_mm=np.asarray(mm, dtype=np.int32), names=names)
