"""BAM -> sR2M with the reference's default read filter (ctypes face of the C++ filter in csrc/isb_host.cpp).

Mirrors inStrain.filter_reads.load_paired_reads (inStrain/filter_reads.py:157-199) for pairing_filter='paired_only'
without priority reads: returns (scaffold -> {read-pair name -> summed NM}, per-scaffold tallies, max_insert).
"""
import ctypes as C

import numpy as np

from . import _cabi

TALLY_COLUMNS = ["pass_pairing_filter", "pass_min_read_ani", "pass_max_insert", "pass_min_insert", "pass_min_mapq",
                 "filtered_pairs"]


def _lib():
    L = _cabi.load()
    if not getattr(L, "_filter_ready", False):
        vp, i64 = C.c_void_p, C.c_int64
        L.isb_filter_open.restype = vp
        L.isb_filter_open.argtypes = [C.c_char_p]
        L.isb_filter_apply.restype = i64
        L.isb_filter_apply.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_int]
        L.isb_filter_n_refs.restype = C.c_int
        L.isb_filter_n_refs.argtypes = [vp]
        L.isb_filter_max_insert.restype = C.c_double
        L.isb_filter_max_insert.argtypes = [vp]
        L.isb_filter_tally.argtypes = [vp, C.c_int, vp]
        L.isb_filter_stats.argtypes = [vp, C.c_int, vp]
        L.isb_filter_n_pairs.restype = i64
        L.isb_filter_n_pairs.argtypes = [vp, C.c_int]
        L.isb_filter_names_bytes.restype = i64
        L.isb_filter_names_bytes.argtypes = [vp, C.c_int]
        L.isb_filter_copy.argtypes = [vp, C.c_int, vp, vp, vp]
        L.isb_filter_free.argtypes = [vp]
        L._filter_ready = True
    return L


def filter_reads(bam, ref_names, min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50, **_):
    """ref_names: the BAM's reference names in header order (BamPacker(bam).ref_names).  Returns (sR2M, tallies, max_insert);
    scaffolds without kept pairs are absent from sR2M (as parse_filter_reads drops them, controller.py:260-322)."""
    lib = _lib()
    h = lib.isb_filter_open(bam.encode())
    if not h:
        raise IOError("cannot read BAM %s" % bam)
    try:
        lib.isb_filter_apply(h, float(min_read_ani), int(min_mapq), float(max_insert_relative), int(min_insert))
        sr2m, tallies = {}, {}
        for tid in range(lib.isb_filter_n_refs(h)):
            t = np.zeros(6, dtype=np.int64)
            lib.isb_filter_tally(h, tid, t.ctypes.data)
            if t[0]:
                tallies[ref_names[tid]] = dict(zip(TALLY_COLUMNS, (int(x) for x in t)))
            n = int(lib.isb_filter_n_pairs(h, tid))
            if n == 0:
                continue
            blob = C.create_string_buffer(int(lib.isb_filter_names_bytes(h, tid)) + 1)
            off = np.zeros(n + 1, dtype=np.int64)
            mm = np.zeros(n, dtype=np.int32)
            lib.isb_filter_copy(h, tid, blob, off.ctypes.data, mm.ctypes.data)
            raw = blob.raw
            sr2m[ref_names[tid]] = {raw[off[i]:off[i + 1]].decode(): int(mm[i]) for i in range(n)}
        return sr2m, tallies, float(lib.isb_filter_max_insert(h))
    finally:
        lib.isb_filter_free(h)


MAPPING_INFO_COLUMNS = ["scaffold", "unfiltered_reads", "unfiltered_pairs", "unfiltered_singletons", "unfiltered_priority_reads",
                        "pass_pairing_filter", "pass_min_read_ani", "pass_max_insert", "pass_min_insert", "pass_min_mapq",
                        "filtered_pairs", "filtered_singletons", "filtered_priority_reads", "mean_mistmaches",
                        "mean_insert_distance", "mean_mapq_score", "mean_pair_length", "mean_PID", "median_insert"]


def mapping_info(bam, ref_names, min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50, **_):
    """The reference's `mapping_info` table (the read report of filter_scaff2pair2info, filter_reads.py:230-298, with the
    pairing tallies of paired_read_filter, :484-502) for the default pairing filter: one row per scaffold with reads,
    preceded by the `all_scaffolds` row (sums; means weighted by pass_pairing_filter)."""
    import pandas as pd
    lib = _lib()
    h = lib.isb_filter_open(bam.encode())
    if not h:
        raise IOError("cannot read BAM %s" % bam)
    try:
        lib.isb_filter_apply(h, float(min_read_ani), int(min_mapq), float(max_insert_relative), int(min_insert))
        rows = []
        for tid in range(lib.isb_filter_n_refs(h)):
            t = np.zeros(6, dtype=np.int64)
            s = np.zeros(10, dtype=np.float64)
            lib.isb_filter_tally(h, tid, t.ctypes.data)
            lib.isb_filter_stats(h, tid, s.ctypes.data)
            if s[0] == 0:
                continue                                           # no read of this scaffold in the BAM
            rows.append([ref_names[tid], int(s[0]), int(s[1]), int(s[2]), 0, int(t[0]), int(t[1]), int(t[2]), int(t[3]),
                         int(t[4]), int(t[5]), 0, 0, s[3], s[4], s[5], s[6], s[7], s[8]])
    finally:
        lib.isb_filter_free(h)
    Adb = pd.DataFrame(rows, columns=MAPPING_INFO_COLUMNS)
    C_ = Adb[Adb["pass_pairing_filter"] > 0]
    total = C_["pass_pairing_filter"].sum()
    top = {"scaffold": "all_scaffolds"}
    for c in MAPPING_INFO_COLUMNS[1:]:
        if c.startswith("mean_") or c.startswith("median_"):
            top[c] = float((C_[c] * C_["pass_pairing_filter"]).sum() / total) if total else float("nan")
        else:
            top[c] = int(C_[c].sum())
    return pd.concat([pd.DataFrame([top], columns=MAPPING_INFO_COLUMNS), Adb]).reset_index(drop=True)
