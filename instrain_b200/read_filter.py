"""BAM -> sR2M with the reference's read filter (ctypes face of the C++ filter in csrc/isb_host.cpp).

Mirrors inStrain.filter_reads.load_paired_reads (inStrain/filter_reads.py:157-199): every pairing filter ('paired_only',
'non_discordant', 'all_reads') and priority reads; returns (scaffold -> {read-pair name -> summed NM}, per-scaffold
tallies, max_insert).
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
        L.isb_filter_open_mt.restype = vp
        L.isb_filter_open_mt.argtypes = [C.c_char_p, C.c_int, C.c_int, vp]
        L.isb_host_last_error.restype = C.c_char_p
        L.isb_host_last_error.argtypes = []
        L.isb_filter_apply.restype = i64
        L.isb_filter_apply.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_int]
        L.isb_filter_apply2.restype = i64
        L.isb_filter_apply2.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, i64, C.c_char_p, vp]
        L.isb_filter_tally2.argtypes = [vp, C.c_int, vp]
        L.isb_filter_n_refs.restype = C.c_int
        L.isb_filter_n_refs.argtypes = [vp]
        L.isb_filter_max_insert.restype = C.c_double
        L.isb_filter_max_insert.argtypes = [vp]
        L.isb_filter_tally.argtypes = [vp, C.c_int, vp]
        L.isb_filter_stats.argtypes = [vp, C.c_int, vp]
        L.isb_filter_stats2.argtypes = [vp, C.c_int, vp]
        L.isb_filter_n_pairs.restype = i64
        L.isb_filter_n_pairs.argtypes = [vp, C.c_int]
        L.isb_filter_names_bytes.restype = i64
        L.isb_filter_names_bytes.argtypes = [vp, C.c_int]
        L.isb_filter_copy.argtypes = [vp, C.c_int, vp, vp, vp]
        L.isb_filter_free.argtypes = [vp]
        L._filter_ready = True
    return L


PAIRING_MODES = {"paired_only": 0, "non_discordant": 1, "all_reads": 2}


def _apply(lib, h, min_read_ani, min_mapq, max_insert_relative, min_insert, pairing_filter, priority_reads):
    if pairing_filter not in PAIRING_MODES:
        raise ValueError("Do not know paired read filter %r" % (pairing_filter,))
    if pairing_filter == "paired_only" and not priority_reads:
        lib.isb_filter_apply(h, float(min_read_ani), int(min_mapq), float(max_insert_relative), int(min_insert))
        return
    enc = [s.encode() for s in priority_reads]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        off[1:] = np.cumsum([len(b) for b in enc])
    if lib.isb_filter_apply2(h, float(min_read_ani), int(min_mapq), float(max_insert_relative), int(min_insert),
                             PAIRING_MODES[pairing_filter], len(enc), b"".join(enc), off.ctypes.data) < 0:
        raise ValueError("isb_filter_apply2 failed")


def _open(lib, bam, threads=1):
    """One pass over the BAM -> filter handle; with threads > 1 and a .bai index the scaffolds are read concurrently."""
    h = None
    if threads and threads > 1:
        from .packer import find_bai, read_bai
        bai = find_bai(bam)
        if bai is not None:
            first = np.array([v or 0 for v in read_bai(bai)], dtype=np.uint64)
            h = lib.isb_filter_open_mt(bam.encode(), int(threads), len(first), first.ctypes.data)
    if h is None:
        h = lib.isb_filter_open(bam.encode())
    if not h:
        raise IOError("error reading BAM: " + lib.isb_host_last_error().decode())
    return h


def _collect_r2m(lib, h, ref_names, general):
    sr2m, tallies = {}, {}
    for tid in range(lib.isb_filter_n_refs(h)):
        t = np.zeros(6, dtype=np.int64)
        lib.isb_filter_tally(h, tid, t.ctypes.data)
        if t[0]:
            tallies[ref_names[tid]] = dict(zip(TALLY_COLUMNS, (int(x) for x in t)))
            if general:
                t2 = np.zeros(3, dtype=np.int64)
                lib.isb_filter_tally2(h, tid, t2.ctypes.data)
                tallies[ref_names[tid]].update(unfiltered_priority_reads=int(t2[0]), filtered_singletons=int(t2[1]),
                                               filtered_priority_reads=int(t2[2]))
        n = int(lib.isb_filter_n_pairs(h, tid))
        if n == 0:
            continue
        blob = C.create_string_buffer(int(lib.isb_filter_names_bytes(h, tid)) + 1)
        off = np.zeros(n + 1, dtype=np.int64)
        mm = np.zeros(n, dtype=np.int32)
        lib.isb_filter_copy(h, tid, blob, off.ctypes.data, mm.ctypes.data)
        raw = blob.raw
        sr2m[ref_names[tid]] = {raw[off[i]:off[i + 1]].decode(): int(mm[i]) for i in range(n)}
    return sr2m, tallies


def filter_reads(bam, ref_names, min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50,
                 pairing_filter="paired_only", priority_reads=(), with_report=False, threads=1, **_):
    """ref_names: the BAM's reference names in header order (BamPacker(bam).ref_names).  Returns (sR2M, tallies, max_insert);
    scaffolds without kept pairs are absent from sR2M (as parse_filter_reads drops them, controller.py:260-322).
    pairing_filter: 'paired_only' (default), 'non_discordant' or 'all_reads'; priority_reads: names that pass the pairing
    filter regardless (filter_reads.py:471-532).  with_report=True appends the `mapping_info` table (see mapping_info) made
    from the SAME pass over the BAM."""
    lib = _lib()
    h = _open(lib, bam, threads)
    try:
        priority_reads = list(priority_reads)
        general = pairing_filter != "paired_only" or bool(priority_reads)
        _apply(lib, h, min_read_ani, min_mapq, max_insert_relative, min_insert, pairing_filter, priority_reads)
        sr2m, tallies = _collect_r2m(lib, h, ref_names, general)
        out = (sr2m, tallies, float(lib.isb_filter_max_insert(h)))
        if with_report:
            out += (_report(lib, h, ref_names, general),)
        return out
    finally:
        lib.isb_filter_free(h)


def load_priority_reads(file_loc):
    """Names of the reads that pass the pairing filter regardless (--priority_reads; load_priority_reads,
    filter_reads.py:428-469): a FASTQ (names from the @ lines) or a plain list of names, optionally gzipped."""
    import gzip
    opener = gzip.open if str(file_loc).endswith(".gz") else open
    with opener(file_loc, "rt") as f:
        lines = f.readlines()
    if lines and lines[0].startswith("@"):
        return {ln[1:].strip() for ln in lines if ln.startswith("@")}
    return {ln.strip() for ln in lines}


MAPPING_INFO_COLUMNS = ["scaffold", "unfiltered_reads", "unfiltered_pairs", "unfiltered_singletons", "unfiltered_priority_reads",
                        "pass_pairing_filter", "pass_min_read_ani", "pass_max_insert", "pass_min_insert", "pass_min_mapq",
                        "filtered_pairs", "filtered_singletons", "filtered_priority_reads", "mean_mistmaches",
                        "mean_insert_distance", "mean_mapq_score", "mean_pair_length", "mean_PID", "median_insert"]


def _report(lib, h, ref_names, general):
    """mapping_info rows of a filter handle the thresholds have been applied to."""
    import pandas as pd
    rows = []
    for tid in range(lib.isb_filter_n_refs(h)):
        t = np.zeros(6, dtype=np.int64)
        t2 = np.zeros(3, dtype=np.int64)
        s = np.zeros(10, dtype=np.float64)
        lib.isb_filter_tally(h, tid, t.ctypes.data)
        if general:                                            # means over what the pairing filter selected
            lib.isb_filter_stats2(h, tid, s.ctypes.data)
            lib.isb_filter_tally2(h, tid, t2.ctypes.data)
        else:
            lib.isb_filter_stats(h, tid, s.ctypes.data)
        if s[0] == 0:
            continue                                           # no read of this scaffold in the BAM
        rows.append([ref_names[tid], int(s[0]), int(s[1]), int(s[2]), int(t2[0]), int(t[0]), int(t[1]), int(t[2]), int(t[3]),
                     int(t[4]), int(t[5]), int(t2[1]), int(t2[2]), s[3], s[4], s[5], s[6], s[7], s[8]])
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


def mapping_info(bam, ref_names, min_read_ani=0.95, min_mapq=-1, max_insert_relative=3, min_insert=50,
                 pairing_filter="paired_only", priority_reads=(), **_):
    """The reference's `mapping_info` table (the read report of filter_scaff2pair2info, filter_reads.py:230-298, with the
    pairing tallies of paired_read_filter, :484-502) for every pairing filter and with priority reads: one row per
    scaffold with reads, preceded by the `all_scaffolds` row (sums; means weighted by pass_pairing_filter)."""
    lib = _lib()
    h = lib.isb_filter_open(bam.encode())
    if not h:
        raise IOError("error reading BAM: " + lib.isb_host_last_error().decode())
    try:
        priority_reads = list(priority_reads)
        general = pairing_filter != "paired_only" or bool(priority_reads)
        _apply(lib, h, min_read_ani, min_mapq, max_insert_relative, min_insert, pairing_filter, priority_reads)
        return _report(lib, h, ref_names, general)
    finally:
        lib.isb_filter_free(h)
