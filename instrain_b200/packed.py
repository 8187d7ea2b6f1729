"""Packed host->device transfer format of the event columns (host-side encoder; decoded on the GPU by kernel K0).

~1 byte per event + 12 bytes per position instead of 10 bytes per event: PCIe, not the kernels, bounds the end-to-end
rate.  Layout: include/instrain_b200.h (isb_packed_batch) and instrain_b200/csrc/isb_k0_expand.cu.
Information kept: position (CSR offsets), base code, "quality >= min_qual" (all K1/K3 ever ask of the quality), pair id
(delta-coded inside a position, where events are sorted by pair id -- a stable sort of the BAM-order column, so the
results are identical).
"""
import numpy as np


def encode_packed(ev, start, L, min_qual=30):
    """ev: position-major event columns (dict of numpy arrays: ref_pos, base, qual, read_id).  Returns the dict of arrays
    isb_profile_batch_packed takes (pos_off, id_base, bqd, esc_evt, esc_id) + n_events."""
    pos = np.asarray(ev["ref_pos"], dtype=np.int64) - start
    n = len(pos)
    if n and (pos.min() < 0 or pos.max() >= L):
        raise ValueError("events outside [start, start+L)")
    rid = np.asarray(ev["read_id"], dtype=np.int64)
    base, qual = np.asarray(ev["base"]), np.asarray(ev["qual"])
    if n > 1:
        d_pos, d_rid = np.diff(pos), np.diff(rid)
        in_order = bool(np.all((d_pos > 0) | ((d_pos == 0) & (d_rid >= 0))))
        del d_pos, d_rid
    else:
        in_order = True
    if not in_order:                                             # by position, then pair id; stable
        order = np.lexsort((rid, pos))
        pos, rid, base, qual = pos[order], rid[order], base[order], qual[order]
    base = np.minimum(base, 4).astype(np.uint8)
    ok = qual >= min_qual
    pos_off = np.searchsorted(pos, np.arange(L + 1, dtype=np.int64)).astype(np.int64)
    first = np.zeros(n, dtype=bool)
    nonempty = pos_off[:-1] < pos_off[1:]
    first[pos_off[:-1][nonempty]] = True
    delta = np.diff(rid, prepend=rid[:1] if n else np.zeros(0, np.int64))
    delta[first] = 0
    esc = delta > 14
    id_base = np.zeros(L, dtype=np.int32)
    id_base[nonempty] = rid[pos_off[:-1][nonempty]]
    bqd = (np.where(esc, 15, delta).astype(np.uint8) | (base << 4) | (ok.astype(np.uint8) << 7)).astype(np.uint8)
    return dict(n_events=n, pos_off=pos_off, id_base=id_base, bqd=np.ascontiguousarray(bqd),
                esc_evt=np.nonzero(esc)[0].astype(np.int64), esc_id=rid[esc].astype(np.int32), min_qual=min_qual)
