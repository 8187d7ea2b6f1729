"""Read-major batch ("aligned segments") -- host-side encoders for isb_profile_reads / isb_pileup_reads.

Layout and rules: include/instrain_b200.h (isb_reads_batch).  One 4-bit ONE-HOT code per aligned base, stored once per
read: A=1, C=2, T=4, G=8 for a base that is an event (quality >= min_qual after htslib's mate-overlap tweak), 0 otherwise;
passing non-ACGT bases go to the separate (nev_pos, nev_pair) list.  The stream is POSITION-ALIGNED: the words of a
segment cover the 8-position columns of the batch coordinate system (word 0 = coordinates [start & ~7, (start & ~7) + 8)),
so the nibble of coordinate p is nibble (p - start) + (start & 7) of the segment's words.

`events_to_reads` rebuilds segments from position-major event columns (what the test fixtures and the oracle use):
the events of a pair form runs of consecutive positions; a pair that enters the same column twice (both mates, htslib's
overlap quirk) gets its second entries in a second layer of runs.  `build_reads` assembles the arrays from segments.
"""
import numpy as np

MAX_SEG_LEN = 256


NO_EVENT = 255


def build_reads(seg_start, seg_len, seg_pair, codes, code_off=None, max_len=MAX_SEG_LEN, odd_blocks=False):
    """seg_*: per-segment arrays in ANY order; codes: one uint8 per aligned base of all segments -- 0..3 = A,C,T,G event,
    4 = passing non-ACGT base, NO_EVENT (255) = not an event -- segment i occupying codes[code_off[i] : code_off[i] +
    seg_len[i]] (default: back to back).  Splits segments longer than MAX_SEG_LEN, sorts by start (stable), lays out the
    word stream and the N-event list."""
    seg_start = np.asarray(seg_start, dtype=np.int64)
    seg_len = np.asarray(seg_len, dtype=np.int64)
    seg_pair = np.asarray(seg_pair, dtype=np.int64)
    codes = np.asarray(codes, dtype=np.uint8)
    if code_off is None:
        code_off = np.concatenate([[0], np.cumsum(seg_len)[:-1]]) if len(seg_len) else np.zeros(0, np.int64)
    code_off = np.asarray(code_off, dtype=np.int64)
    if not 1 <= max_len <= MAX_SEG_LEN:
        raise ValueError("max_len must be in [1, %d]" % MAX_SEG_LEN)
    if len(seg_len) and seg_len.max() > max_len:                           # split long blocks
        MAX_SEG_LEN_ = max_len
        n_parts = (seg_len + MAX_SEG_LEN_ - 1) // MAX_SEG_LEN_
        rep = np.repeat(np.arange(len(seg_len)), n_parts)
        part = np.arange(len(rep)) - np.repeat(np.cumsum(n_parts) - n_parts, n_parts)
        seg_start, seg_pair = seg_start[rep] + part * MAX_SEG_LEN_, seg_pair[rep]
        code_off = code_off[rep] + part * MAX_SEG_LEN_
        seg_len = np.minimum(seg_len[rep] - part * MAX_SEG_LEN_, MAX_SEG_LEN_)
    order = np.argsort(seg_start, kind="stable")
    seg_start, seg_len, seg_pair, code_off = seg_start[order], seg_len[order], seg_pair[order], code_off[order]
    n = len(seg_start)
    sh = seg_start & 7                                                     # position-aligned words: leading zero nibbles
    nw = (sh + seg_len + 7) // 8
    blk = nw + 1                                                           # data words + separator(s)
    if odd_blocks:
        blk = blk + (1 - blk % 2)                                          # odd block sizes spread K1r's shared-memory banks
    seg_word = np.ones(n, dtype=np.int64)
    if n:
        seg_word[1:] = 1 + np.cumsum(blk[:-1])
    n_words = int(seg_word[-1] + blk[-1]) if n else 1
    n_words = (n_words + 3) // 4 * 4
    # base j of segment i -> nibble j + sh[i]: word seg_word[i] + (j + sh) // 8, bits 4 * ((j + sh) % 8)
    tot = int(seg_len.sum())
    seg_of = np.repeat(np.arange(n), seg_len)
    j = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(seg_len) - seg_len, seg_len)
    c = codes[np.repeat(code_off, seg_len) + j]
    onehot = np.where(c < 4, np.uint8(1) << np.minimum(c, 3), 0).astype(np.uint8)
    is_n = c == 4
    nev_pos = (np.repeat(seg_start, seg_len) + j)[is_n]
    nev_pair = np.repeat(seg_pair, seg_len)[is_n]
    nib = np.zeros(n_words * 8, dtype=np.uint8)
    nib[(np.repeat(seg_word * 8 + sh, seg_len) + j)] = onehot
    nib = nib.reshape(-1, 8).astype(np.uint32)
    words = np.zeros(n_words, dtype=np.uint32)
    for k in range(8):
        words |= nib[:, k] << np.uint32(4 * k)
    return dict(n_segs=n, seg_start=seg_start.astype(np.int32), seg_len=seg_len.astype(np.uint16),
                seg_pair=seg_pair.astype(np.int32), seg_word=seg_word, n_words=n_words, words=words,
                max_seg_len=int(seg_len.max()) if n else 1, nev_pos=nev_pos.astype(np.int32),
                nev_pair=nev_pair.astype(np.int32))


def events_to_reads(ev, min_qual=30, max_len=MAX_SEG_LEN, odd_blocks=False):
    """Position-major (or any-order) event columns -> read-major batch arrays.  Events below min_qual keep their place in
    their run as NO_EVENT."""
    pos = np.asarray(ev["ref_pos"], dtype=np.int64)
    rid = np.asarray(ev["read_id"], dtype=np.int64)
    base = np.minimum(np.asarray(ev["base"]), 4).astype(np.uint8)
    ok = np.asarray(ev["qual"]) >= min_qual
    n = len(pos)
    if n == 0:
        return build_reads([], [], [], [])
    order = np.lexsort((pos, rid))                                        # by pair, then position (stable)
    pos, rid, base, ok = pos[order], rid[order], base[order], ok[order]
    same = np.zeros(n, dtype=bool)
    same[1:] = (rid[1:] == rid[:-1]) & (pos[1:] == pos[:-1])
    # layer = index of the event among the events of its (pair, position)
    grp_start = np.maximum.accumulate(np.where(~same, np.arange(n), 0))
    layer = np.arange(n) - grp_start
    order2 = np.lexsort((pos, layer, rid))
    pos, rid, base, ok, layer = pos[order2], rid[order2], base[order2], ok[order2], layer[order2]
    brk = np.ones(n, dtype=bool)
    brk[1:] = (rid[1:] != rid[:-1]) | (layer[1:] != layer[:-1]) | (pos[1:] != pos[:-1] + 1)
    starts = np.nonzero(brk)[0]
    seg_len = np.diff(np.append(starts, n))
    codes = np.where(ok, base, NO_EVENT).astype(np.uint8)
    return build_reads(pos[starts], seg_len, rid[starts], codes, starts, max_len, odd_blocks)


def reads_to_events(rd, min_qual=30):
    """Inverse view for tests: the passing events (position-major) a read-major batch encodes."""
    seg_len = rd["seg_len"].astype(np.int64)
    tot = int(seg_len.sum())
    j = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(seg_len) - seg_len, seg_len)
    jn = j + np.repeat(rd["seg_start"].astype(np.int64) & 7, seg_len)      # nibble index inside the segment's words
    w = rd["words"][np.repeat(rd["seg_word"], seg_len) + jn // 8]
    code = (w >> (4 * (jn % 8)).astype(np.uint32)) & 15
    keep = code != 0
    pos = np.concatenate([(np.repeat(rd["seg_start"].astype(np.int64), seg_len) + j)[keep], rd["nev_pos"].astype(np.int64)])
    rid = np.concatenate([np.repeat(rd["seg_pair"], seg_len)[keep], rd["nev_pair"]])
    base = np.concatenate([np.log2(code[keep]).astype(np.uint8), np.full(len(rd["nev_pos"]), 4, np.uint8)])
    order = np.lexsort((rid, pos))
    return dict(ref_pos=pos[order].astype(np.int32), base=base[order], qual=np.full(len(order), 255, np.uint8),
                read_id=rid[order].astype(np.int32))


def compact_reads(rd):
    """Read-major batch -> compact TRANSFER format (include/instrain_b200.h, isb_reads_compact): per unit of 8 bases one
    uint16 of 2-bit base codes and one uint8 of event bits; no word offsets (the device rebuilds the canonical stream)."""
    seg_len = rd["seg_len"].astype(np.int64)
    nw = ((rd["seg_start"].astype(np.int64) & 7) + seg_len + 7) // 8       # units = the segment's (position-aligned) words
    n_units = int(nw.sum())
    k = np.arange(n_units, dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
    w = rd["words"][np.repeat(np.asarray(rd["seg_word"], dtype=np.int64), nw) + k].astype(np.uint32)
    base2 = np.zeros(n_units, dtype=np.uint16)
    ps = np.zeros(n_units, dtype=np.uint8)
    for t in range(8):
        nib = (w >> np.uint32(4 * t)) & np.uint32(15)
        code = np.select([nib == 2, nib == 4, nib == 8], [1, 2, 3], 0).astype(np.uint16)
        base2 |= code << np.uint16(2 * t)
        ps |= (nib != 0).astype(np.uint8) << np.uint8(t)
    out = {k_: rd[k_] for k_ in ("n_segs", "seg_start", "seg_len", "seg_pair", "max_seg_len", "nev_pos", "nev_pair")}
    out.update(n_units=n_units, base2=base2, **{"pass": ps})
    return out


def _units(rd):
    """(unit count per segment, flat unit -> segment index, flat unit -> index inside its segment, source word index)."""
    seg_len = rd["seg_len"].astype(np.int64)
    nw = ((rd["seg_start"].astype(np.int64) & 7) + seg_len + 7) // 8
    n_units = int(nw.sum())
    seg = np.repeat(np.arange(len(nw)), nw)
    k = np.arange(n_units, dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
    src = np.repeat(np.asarray(rd["seg_word"], dtype=np.int64), nw) + k
    return nw, seg, k, src


def delta_reads(rd, ref_codes, start=0):
    """Read-major batch -> reference-delta TRANSFER format (include/instrain_b200.h, isb_reads_delta): per unit of 8 bases
    the event bits only; per passing base that differs from the reference one (mis_word, mis_code) entry.  ref_codes[i] is
    the reference code (0..3 = A,C,T,G, 4 = other) of batch coordinate start + i."""
    ref_codes = np.asarray(ref_codes, dtype=np.uint8)
    nw, seg, k, src = _units(rd)
    n_units = len(seg)
    w = rd["words"][src].astype(np.uint32)
    ps = np.zeros(n_units, dtype=np.uint8)
    pos0 = (np.repeat(rd["seg_start"].astype(np.int64) & ~np.int64(7), nw) + 8 * k) - start    # ref index of nibble 0
    canon_word = 1 + np.arange(n_units, dtype=np.int64) + seg                                  # canonical stream index
    mis_word, mis_code = [], []
    L = len(ref_codes)
    for t in range(8):
        nib = (w >> np.uint32(4 * t)) & np.uint32(15)
        has = nib != 0
        ps |= has.astype(np.uint8) << np.uint8(t)
        r = np.full(n_units, 4, dtype=np.uint8)
        inside = (pos0 + t >= 0) & (pos0 + t < L)
        r[inside] = ref_codes[(pos0 + t)[inside]]
        ref_hot = np.where(r < 4, np.uint32(1) << np.minimum(r, 3).astype(np.uint32), 0).astype(np.uint32)
        diff = has & (nib != ref_hot)
        mis_word.append(canon_word[diff])
        mis_code.append(((nib[diff] ^ ref_hot[diff]) | np.uint32(t << 4)).astype(np.uint8))
    out = {k_: rd[k_] for k_ in ("n_segs", "seg_start", "seg_len", "seg_pair", "max_seg_len", "nev_pos", "nev_pair")}
    mw = np.concatenate(mis_word) if mis_word else np.zeros(0, np.int64)
    if len(mw) and mw.max() > 0xffffffff:
        raise ValueError("batch too large for 32-bit word indices")
    out.update(n_units=n_units, mis_word=mw.astype(np.uint32), mis_code=np.concatenate(mis_code) if mis_code else np.zeros(0, np.uint8),
               **{"pass": ps})
    return out


def delta_reads_host(rd, ref_codes, start=0):
    """delta_reads through the C++ host routine (isb_reads_delta_host): what the host pipeline uses.  Same result as
    delta_reads up to the order of the mismatch entries (stream order here)."""
    from . import _cabi
    lib = _cabi.load()
    p = _cabi.ptr
    ref_codes = np.ascontiguousarray(ref_codes, dtype=np.uint8)
    seg_start = np.ascontiguousarray(rd["seg_start"], dtype=np.int32)
    seg_len = np.ascontiguousarray(rd["seg_len"], dtype=np.uint16)
    seg_word = np.ascontiguousarray(rd["seg_word"], dtype=np.int64)
    win = np.ascontiguousarray(rd["words"], dtype=np.uint32)
    n_units = int((((seg_start.astype(np.int64) & 7) + seg_len.astype(np.int64) + 7) // 8).sum())
    ps = np.zeros(n_units, dtype=np.uint8)
    cap = max(1024, n_units // 8)
    while True:
        mw, mc = np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.uint8)
        n = lib.isb_reads_delta_host(int(rd["n_segs"]), p(seg_start), p(seg_len), p(seg_word), p(win), len(win), start,
                                     len(ref_codes), p(ref_codes), p(ps), n_units, p(mw), p(mc), cap)
        if n < 0:
            raise ValueError("read-major batch violates its layout rules")
        if n <= cap:
            break
        cap = int(n)
    out = {k_: rd[k_] for k_ in ("n_segs", "seg_start", "seg_len", "seg_pair", "max_seg_len", "nev_pos", "nev_pair")}
    out.update(n_units=n_units, mis_word=mw[:n].copy(), mis_code=mc[:n].copy(), **{"pass": ps})
    return out


def delta_to_words(dl, ref_codes, start=0):
    """What K0d computes (numpy restatement, for the CPU round-trip test): the canonical nibble stream of a delta batch
    -> (seg_word, n_words, words)."""
    ref_codes = np.asarray(ref_codes, dtype=np.uint8)
    s = dl["seg_start"].astype(np.int64)
    nw = ((s & 7) + dl["seg_len"].astype(np.int64) + 7) // 8
    n = len(s)
    seg_word = 1 + (np.cumsum(nw) - nw) + np.arange(n)
    n_words = (1 + int(nw.sum()) + n + 3) // 4 * 4
    words = np.zeros(n_words, dtype=np.uint32)
    seg = np.repeat(np.arange(n), nw)
    k = np.arange(int(nw.sum()), dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
    pos0 = (np.repeat(s & ~np.int64(7), nw) + 8 * k) - start
    L = len(ref_codes)
    w = np.zeros(len(seg), dtype=np.uint32)
    for t in range(8):
        r = np.full(len(seg), 4, dtype=np.uint8)
        inside = (pos0 + t >= 0) & (pos0 + t < L)
        r[inside] = ref_codes[(pos0 + t)[inside]]
        hot = np.where(r < 4, np.uint32(1) << np.minimum(r, 3).astype(np.uint32), 0).astype(np.uint32)
        w |= np.where((dl["pass"] >> np.uint8(t)) & 1, hot, 0).astype(np.uint32) << np.uint32(4 * t)
    words[np.repeat(seg_word, nw) + k] = w
    np.bitwise_xor.at(words, dl["mis_word"].astype(np.int64),
                      (dl["mis_code"].astype(np.uint32) & 15) << (4 * (dl["mis_code"].astype(np.uint32) >> 4)))
    return seg_word, n_words, words


def concat_streams(parts):
    """Scaffold-wise packer outputs (BamPacker.pack_scaffold_reads; coordinates / pair ids already offset) -> one batch:
    one leading zero word + the scaffolds' streams + zero padding to a multiple of 4 words."""
    n_words = 1 + sum(len(p["stream"]) for p in parts)
    n_words = (n_words + 3) // 4 * 4
    words = np.zeros(n_words, dtype=np.uint32)
    seg_word, base = [], 1
    for p in parts:
        words[base:base + len(p["stream"])] = p["stream"]
        seg_word.append(p["seg_word"] + base)
        base += len(p["stream"])
    cat = lambda k, dt: (np.concatenate([p[k] for p in parts]) if parts else np.zeros(0, dt)).astype(dt, copy=False)
    return dict(n_segs=sum(len(p["seg_start"]) for p in parts), seg_start=cat("seg_start", np.int32),
                seg_len=cat("seg_len", np.uint16), seg_pair=cat("seg_pair", np.int32),
                seg_word=np.concatenate(seg_word) if seg_word else np.zeros(0, np.int64), n_words=n_words, words=words,
                max_seg_len=max([p["max_seg_len"] for p in parts] + [1]), nev_pos=cat("nev_pos", np.int32),
                nev_pair=cat("nev_pair", np.int32))


def clip_reads(rd, lo, hi):
    """The part of a read-major batch that lies in the coordinates [lo, hi): every segment that overlaps the range, cut to
    it (first / last word masked), re-laid-out as a batch of its own.  Coordinates and pair ids are KEPT; the batch is to be
    profiled with `start = origin = lo & ~7` and `L = hi - origin` (a batch must start on an 8-position column; the up to 7
    positions in front of `lo` carry no events), so rows, and the keys of the re-drawn outputs, come out in the
    coordinates of the whole.  Returns (batch, origin).

    This is the read halo of SURVEY 8(e): a contiguous run of splits of ONE large scaffold becomes a batch of its own --
    the reads that reach into the run from outside are cut at its borders, exactly what a position of the run sees of them
    -- so the runs of a scaffold can be profiled on different GPUs (linkage never crosses a split,
    inStrain/profile/profile_utilities.py:165,184-185; the mate-overlap tweak is already in the codes)."""
    lo, hi = int(lo), int(hi)
    origin = lo & ~7
    s = rd["seg_start"].astype(np.int64)
    n = rd["seg_len"].astype(np.int64)
    keep = (s < hi) & (s + n > lo)
    s, n = s[keep], n[keep]
    sw = np.asarray(rd["seg_word"], dtype=np.int64)[keep]
    s2, e2 = np.maximum(s, lo), np.minimum(s + n, hi)
    c_first, c_last = (s2 >> 3) - (s >> 3), ((e2 - 1) >> 3) - (s >> 3)       # word range inside the original segment
    nw = c_last - c_first + 1
    m = len(s)
    seg_word = np.ones(m, dtype=np.int64)
    if m:
        seg_word[1:] = 1 + np.cumsum(nw[:-1] + 1)
    n_words = int(seg_word[-1] + nw[-1] + 1) if m else 1
    n_words = (n_words + 3) // 4 * 4
    words = np.zeros(n_words, dtype=np.uint32)
    k = np.arange(int(nw.sum()), dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
    w = rd["words"][np.repeat(sw + c_first, nw) + k].astype(np.uint64)
    first = k == 0
    last = k == np.repeat(nw, nw) - 1
    lo_nib = np.repeat(s2 & 7, nw)                                            # nibbles below s2 in the first word: cut
    hi_nib = np.repeat(((e2 - 1) & 7) + 1, nw)                                # nibbles kept in the last word
    w = np.where(first, w & ~((np.uint64(1) << (4 * lo_nib).astype(np.uint64)) - np.uint64(1)), w)
    w = np.where(last, w & ((np.uint64(1) << (4 * hi_nib).astype(np.uint64)) - np.uint64(1)), w)
    words[np.repeat(seg_word, nw) + k] = (w & np.uint64(0xffffffff)).astype(np.uint32)
    nk = (rd["nev_pos"] >= lo) & (rd["nev_pos"] < hi)
    out = dict(n_segs=m, seg_start=s2.astype(np.int32), seg_len=(e2 - s2).astype(np.uint16),
               seg_pair=np.asarray(rd["seg_pair"])[keep].astype(np.int32), seg_word=seg_word, n_words=n_words, words=words,
               max_seg_len=int(rd["max_seg_len"]), nev_pos=rd["nev_pos"][nk].astype(np.int32),
               nev_pair=rd["nev_pair"][nk].astype(np.int32))
    return out, origin
