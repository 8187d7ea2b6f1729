"""Column-word batch ("pileup-major" form of the aligned segments) -- host-side builders for isb_profile_cols /
isb_pileup_cols.

Layout and rules: include/instrain_b200.h (isb_cols_batch).  The one-hot nibble words of a read-major batch
(instrain_b200/reads.py), regrouped per COLUMN WORD (8 positions): group g = 64 positions = 8 column words, chunk =
8 lanes x 8 words; slot i of column (g, lane) is word ((grp_off[g] + i // 8) * 8 + lane) * 8 + i % 8, its read-pair
id the same element of `ids` (-1 = padding).  Words of a column keep the table order of their segments.

`reads_to_cols` calls the C++ host packer routine (isb_cols_from_reads_host, no GPU); `reads_to_cols_numpy` is an
independent vectorised restatement used by the tests to pin it.
"""
import numpy as np

LANES = 8               # column words per group (ISB_COLS_LANES)
UNIT = 8                # consecutive slots of one column per chunk (ISB_COLS_UNIT)
GROUP = 8 * LANES       # positions per group
CHUNK = UNIT * LANES    # words per chunk


def _seg_columns(rd, start):
    """(segment index, column word index, source word index) of every data word, in table order."""
    s = rd["seg_start"].astype(np.int64)
    n = rd["seg_len"].astype(np.int64)
    nw = ((s & 7) + n + 7) // 8
    seg = np.repeat(np.arange(len(s)), nw)
    k = np.arange(int(nw.sum()), dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
    col = np.repeat((s - start) >> 3, nw) + k
    src = np.repeat(np.asarray(rd["seg_word"], dtype=np.int64), nw) + k
    return seg, col, src


def reads_to_cols_numpy(rd, L, start=0):
    if start & 7:
        raise ValueError("start must be a multiple of 8")
    n_groups = (L + GROUP - 1) // GROUP
    seg, col, src = _seg_columns(rd, start)
    cnt = np.bincount(col, minlength=n_groups * LANES).astype(np.int64) if len(col) else np.zeros(n_groups * LANES, np.int64)
    depth_chunks = (cnt.reshape(n_groups, LANES).max(axis=1) + UNIT - 1) // UNIT if n_groups else np.zeros(0, np.int64)
    grp_off = np.zeros(n_groups + 1, dtype=np.int64)
    grp_off[1:] = np.cumsum(depth_chunks)
    n_chunks = int(grp_off[-1])
    words = np.zeros(n_chunks * CHUNK, dtype=np.uint32)
    ids = np.full(n_chunks * CHUNK, -1, dtype=np.int32)
    if len(col):
        order = np.argsort(col, kind="stable")                 # table order inside a column
        col_s, seg_s, src_s = col[order], seg[order], src[order]
        first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        slot = np.arange(len(col_s), dtype=np.int64) - first[col_s]
        idx = ((grp_off[col_s // LANES] + slot // UNIT) * LANES + (col_s % LANES)) * UNIT + slot % UNIT
        words[idx] = rd["words"][src_s]
        ids[idx] = rd["seg_pair"][seg_s]
    return dict(n_groups=n_groups, grp_off=grp_off, n_chunks=n_chunks, words=words, ids=ids,
                nev_pos=rd["nev_pos"], nev_pair=rd["nev_pair"])


def reads_to_cols(rd, L, start=0):
    """Read-major batch (host numpy arrays) -> column-word batch through the C++ host routine."""
    from . import _cabi
    lib = _cabi.load()
    p = _cabi.ptr
    n_groups = (L + GROUP - 1) // GROUP
    seg_start = np.ascontiguousarray(rd["seg_start"], dtype=np.int32)
    seg_len = np.ascontiguousarray(rd["seg_len"], dtype=np.uint16)
    seg_pair = np.ascontiguousarray(rd["seg_pair"], dtype=np.int32)
    seg_word = np.ascontiguousarray(rd["seg_word"], dtype=np.int64)
    win = np.ascontiguousarray(rd["words"], dtype=np.uint32)
    grp_off = np.zeros(n_groups + 1, dtype=np.int64)
    args = (int(rd["n_segs"]), p(seg_start), p(seg_len), p(seg_pair), p(seg_word), p(win), len(win), start, L, p(grp_off))
    n_chunks = lib.isb_cols_from_reads_host(*args, None, None, 0)
    if n_chunks < 0:
        raise ValueError("read-major batch violates its layout rules")
    words = np.empty(n_chunks * CHUNK, dtype=np.uint32)
    ids = np.empty(n_chunks * CHUNK, dtype=np.int32)
    if lib.isb_cols_from_reads_host(*args, p(words), p(ids), n_chunks) != n_chunks:
        raise RuntimeError("isb_cols_from_reads_host: inconsistent sizing")
    return dict(n_groups=n_groups, grp_off=grp_off, n_chunks=int(n_chunks), words=words, ids=ids,
                nev_pos=rd["nev_pos"], nev_pair=rd["nev_pair"])


def cols_to_events(cd, L, start=0):
    """Inverse view for tests: the passing events (position-major, column order) a column-word batch encodes."""
    pos, rid, base = [], [], []
    w = cd["words"].reshape(-1, LANES, UNIT)
    i = cd["ids"].reshape(-1, LANES, UNIT)
    for g in range(cd["n_groups"]):
        c0, c1 = int(cd["grp_off"][g]), int(cd["grp_off"][g + 1])
        if c1 == c0:
            continue
        ww = w[c0:c1].transpose(1, 0, 2).reshape(LANES, -1)       # [lane][slot]
        ii = i[c0:c1].transpose(1, 0, 2).reshape(LANES, -1)
        for k in range(8):
            code = (ww >> np.uint32(4 * k)) & np.uint32(15)
            lane, slot = np.nonzero((code != 0) & (ii >= 0))
            pos.append(start + g * GROUP + lane * 8 + k)
            rid.append(ii[lane, slot])
            base.append(np.log2(code[lane, slot]).astype(np.uint8))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    pos, rid, base = cat(pos, np.int64), cat(rid, np.int64), cat(base, np.uint8)
    pos = np.concatenate([pos, cd["nev_pos"].astype(np.int64)])
    rid = np.concatenate([rid, cd["nev_pair"].astype(np.int64)])
    base = np.concatenate([base, np.full(len(cd["nev_pos"]), 4, np.uint8)])
    order = np.lexsort((rid, pos))
    return dict(ref_pos=pos[order].astype(np.int32), base=base[order], qual=np.full(len(order), 255, np.uint8),
                read_id=rid[order].astype(np.int32))
