"""Seeded synthetic metagenome -> position-major event batch (test infrastructure; numpy, CPU).

Follows SURVEY.md section 8d: iid uniform reference, K=4 haplotypes with abundances (0.4,0.3,0.2,0.1), Bernoulli SNV
sites whose alt base is carried by a random non-empty proper subset of haplotypes, 2x150 read pairs with fragment
length N(350,30) clipped to [200,500], base qualities from the bundled BAM's empirical bins, substitution errors
with p = 10^(-q/10), mm = mismatches of the pair vs the reference, pairs with 1 - mm/300 <= 0.95 dropped
(evaluate_pair, inStrain/filter_reads.py:406-408).  Mates that overlap go through htslib's overlap-quality tweak
(the all-M special case of oracle/pileup_emul.tweak_overlap_quality: no indels, so no walker quirk).
"""
import numpy as np

QUAL_BINS = np.array([8, 12, 22, 27, 32, 37, 41], dtype=np.uint8)
QUAL_P = np.array([0.002, 0.045, 0.034, 0.047, 0.085, 0.164, 0.623])
HAP_ABUND = np.array([0.4, 0.3, 0.2, 0.1])
READLEN = 150


def make_batch(L, coverage, snv_density, seed, n_scaffolds=1, window=10000, skip_mm=False, n_frac=0.0):
    """Return a batch dict (same keys as tests/golden/*_batch.npz) for n_scaffolds scaffolds of length L each.

    skip_mm: R2M is a set (--skip_mm_profiling): pair_mm == 0 for every kept pair (M = 1).
    n_frac:  fraction of read bases replaced by a non-ACGT base (code 4) keeping their quality.
    """
    from .bamio import iterate_splits
    rng = np.random.Generator(np.random.PCG64(seed))
    out = dict(ref_codes=[], ref_pos=[], base=[], qual=[], read_id=[], pair_mm=[], splits=[])
    pair_off = 0
    for s in range(n_scaffolds):
        off = s * L
        ref = rng.integers(0, 4, L, dtype=np.uint8)
        is_snv = rng.random(L) < snv_density
        alt = ((ref + 1 + rng.integers(0, 3, L)) % 4).astype(np.uint8)
        carriers = rng.integers(1, 15, L)                       # non-empty proper subset of 4 haplotypes
        hap = np.empty((4, L), dtype=np.uint8)
        for h in range(4):
            hap[h] = np.where(is_snv & ((carriers >> h) & 1).astype(bool), alt, ref)
        n_pairs = int(coverage * L / (2 * READLEN))
        frag = np.clip(np.rint(rng.normal(350, 30, n_pairs)), 200, 500).astype(np.int64)
        frag = np.minimum(frag, L)
        start = (rng.random(n_pairs) * (L - frag + 1)).astype(np.int64)
        order = np.argsort(start, kind="stable")                # pair id = BAM order of the first mate
        frag, start = frag[order], start[order]
        hp = rng.choice(4, n_pairs, p=HAP_ABUND)
        offs = np.arange(READLEN, dtype=np.int64)
        pos1 = start[:, None] + offs[None, :]
        pos2 = (start + frag - READLEN)[:, None] + offs[None, :]
        pos = np.stack([pos1, pos2], 1)                          # [pair, mate, offset]
        true = hap[hp[:, None, None], pos]
        q = rng.choice(QUAL_BINS, size=pos.shape, p=QUAL_P / QUAL_P.sum())
        err = rng.random(pos.shape) < 10.0 ** (-q.astype(np.float64) / 10.0)
        sub = ((true + 1 + rng.integers(0, 3, pos.shape)) % 4).astype(np.uint8)
        base = np.where(err, sub, true).astype(np.uint8)
        mm = (base != ref[pos]).sum((1, 2))
        keep = (1.0 - mm / float(2 * READLEN)) > 0.95
        # mate-overlap tweak (a = mate 1 arrived first): overlap offsets of mate 1 are [frag-150, 150)
        ov = np.maximum(0, 2 * READLEN - frag)                   # overlap length
        i1 = offs[None, :] + (frag - READLEN)[:, None]           # mate-1 offset matching mate-2 offset j
        valid = offs[None, :] < ov[:, None]
        rows = np.nonzero(valid)
        pi, j2 = rows
        j1 = i1[pi, j2]
        qa = q[pi, 0, j1].astype(np.int64)
        qb = q[pi, 1, j2].astype(np.int64)
        same = base[pi, 0, j1] == base[pi, 1, j2]
        a_wins = qa >= qb
        new_a = np.where(same, np.minimum(200, qa + qb), np.where(a_wins, np.floor(0.8 * qa), 0))
        new_b = np.where(same, 0, np.where(a_wins, 0, np.floor(0.8 * qb)))
        q = q.copy()
        q[pi, 0, j1] = new_a.astype(np.uint8)
        q[pi, 1, j2] = new_b.astype(np.uint8)
        if n_frac > 0:
            base = np.where(rng.random(pos.shape) < n_frac, 4, base).astype(np.uint8)
        # keep filtered pairs, re-number ids in BAM order, emit reads in file order (sorted by read start)
        kp = np.nonzero(keep)[0]
        new_id = np.full(n_pairs, -1, dtype=np.int64)
        new_id[kp] = np.arange(len(kp))
        r_start = np.concatenate([start[kp], start[kp] + frag[kp] - READLEN])
        r_pair = np.concatenate([kp, kp])
        r_mate = np.concatenate([np.zeros(len(kp), np.int64), np.ones(len(kp), np.int64)])
        fo = np.argsort(r_start, kind="stable")
        r_pair, r_mate = r_pair[fo], r_mate[fo]
        e_pos = pos[r_pair, r_mate].reshape(-1)
        e_base = base[r_pair, r_mate].reshape(-1)
        e_qual = q[r_pair, r_mate].reshape(-1)
        e_rid = np.repeat(new_id[r_pair], READLEN)
        po = np.argsort(e_pos, kind="stable")                    # position-major, column order = file order
        out["ref_codes"].append(ref)
        out["ref_pos"].append((e_pos[po] + off).astype(np.int32))
        out["base"].append(e_base[po])
        out["qual"].append(e_qual[po].astype(np.uint8))
        out["read_id"].append((e_rid[po] + pair_off).astype(np.int32))
        out["pair_mm"].append(np.zeros(len(kp), np.int32) if skip_mm else mm[kp].astype(np.int32))
        out["splits"].extend([(a + off, b + off) for a, b in iterate_splits(L, window)])
        pair_off += len(kp)
    cat = np.concatenate
    return dict(ref_codes=cat(out["ref_codes"]), ref_pos=cat(out["ref_pos"]), base=cat(out["base"]),
                qual=cat(out["qual"]), read_id=cat(out["read_id"]), pair_mm=cat(out["pair_mm"]),
                splits=np.array(out["splits"], np.int32))
