"""ctypes binding of libinstrain_b200.so (include/instrain_b200.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libinstrain_b200.so")

ISB_OK = 0
ISB_ERR_CUDA, ISB_ERR_ARG, ISB_ERR_CAPACITY, ISB_ERR_ORDER, ISB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
ISB_K1_ANY_ORDER = 0x1
ISB_SKIP_LINKAGE = 0x2
ISB_NO_SYNC = 0x4
ISB_PIPELINE = 0x8
ISB_SITE_ANYSNP = 0x10
ISB_MAX_MM = 64

SNV_DT = np.dtype([("pos", "<i4"), ("cnt", "<i4", (4,)), ("mm", "<i4"), ("ref", "u1"), ("con", "u1"),
                   ("var", "u1"), ("allele_count", "u1"), ("cls", "u1"), ("cryptic", "u1"), ("pad", "u1", (2,))])
LD_DT = np.dtype([("pos_a", "<i4"), ("pos_b", "<i4"), ("mm", "<i4"), ("c_AB", "<i4"), ("c_Ab", "<i4"),
                  ("c_aB", "<i4"), ("c_ab", "<i4"), ("allele_A", "u1"), ("allele_a", "u1"), ("allele_B", "u1"),
                  ("allele_b", "u1"), ("r2", "<f8"), ("d_prime", "<f8"), ("r2_normalized", "<f8"),
                  ("d_prime_normalized", "<f8")])
assert SNV_DT.itemsize == 32 and LD_DT.itemsize == 64

SUMMARY_DT = np.dtype([("length", "<i8"), ("nonzero", "<i8"), ("sum_cov", "<i8"), ("sum_cov2", "<u8"), ("counted", "<i8"),
                       ("sum_clon", "<f8"), ("cov_med_lo", "<i4"), ("cov_med_hi", "<i4"), ("clon_med_lo", "<f4"),
                       ("clon_med_hi", "<f4"), ("present", "<i4"), ("pad", "<i4")])
assert SUMMARY_DT.itemsize == 72

CLASS_NAMES = ["AmbiguousReference", "DivergentSite", "SNS", "SNV", "con_SNV", "pop_SNV"]

# every symbol include/instrain_b200.h declares (tests/test_cabi_symbols.py checks the list against the header)
EXPORTS = ["isb_create", "isb_destroy", "isb_last_error", "isb_abi_version", "isb_set_stream", "isb_synchronize", "isb_row_counts_async",
           "isb_pileup_counts", "isb_call_snvs", "isb_linkage", "isb_profile_batch", "isb_profile_batch_packed",
           "isb_pileup_reads", "isb_profile_reads", "isb_profile_reads_compact", "isb_profile_reads_delta", "isb_reads_delta_host",
           "isb_cols_from_reads", "isb_cols_from_reads_host", "isb_pileup_cols", "isb_profile_cols",
           "isb_scaffold_summary", "isb_launch_count",
           "isb_enable_timing", "isb_stage_times", "isb_selftest_division",
           "isb_bam_open", "isb_bam_close", "isb_bam_n_refs", "isb_bam_ref_name", "isb_bam_ref_len", "isb_bam_error",
           "isb_bam_peek_tid", "isb_host_last_error", "isb_bam_seek", "isb_pack_scaffold", "isb_events_count", "isb_events_pairs", "isb_events_reads_seen",
           "isb_events_reads_packed", "isb_events_copy", "isb_events_free",
           "isb_pack_scaffold_reads", "isb_pack_scaffold_reads_region", "isb_reads_segs", "isb_reads_stream_words", "isb_reads_pairs", "isb_reads_n_events",
           "isb_reads_nev", "isb_reads_max_len", "isb_reads_reads_seen", "isb_reads_reads_packed", "isb_reads_copy",
           "isb_reads_free",
           "isb_filter_open", "isb_filter_open_mt", "isb_filter_apply", "isb_filter_apply2", "isb_filter_tally2", "isb_filter_n_refs", "isb_filter_max_insert", "isb_filter_tally", "isb_filter_stats", "isb_filter_stats2",
           "isb_filter_n_pairs", "isb_filter_names_bytes", "isb_filter_copy", "isb_filter_free"]


class IsbBatch(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("ref_pos", C.c_void_p), ("base", C.c_void_p), ("qual", C.c_void_p),
                ("read_id", C.c_void_p), ("n_pairs", C.c_int64), ("pair_mm", C.c_void_p), ("start", C.c_int32),
                ("L", C.c_int32), ("ref", C.c_void_p), ("n_splits", C.c_int32), ("splits", C.c_void_p),
                ("M", C.c_int32)]


class IsbPackedBatch(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("pos_off", C.c_void_p), ("id_base", C.c_void_p), ("bqd", C.c_void_p),
                ("n_esc", C.c_int64), ("esc_evt", C.c_void_p), ("esc_id", C.c_void_p), ("n_pairs", C.c_int64),
                ("pair_mm", C.c_void_p), ("start", C.c_int32), ("L", C.c_int32), ("ref", C.c_void_p),
                ("n_splits", C.c_int32), ("splits", C.c_void_p), ("M", C.c_int32), ("min_qual", C.c_int32)]


class IsbReadsBatch(C.Structure):
    _fields_ = [("n_segs", C.c_int64), ("seg_start", C.c_void_p), ("seg_len", C.c_void_p), ("seg_pair", C.c_void_p),
                ("seg_word", C.c_void_p), ("n_words", C.c_int64), ("words", C.c_void_p), ("max_seg_len", C.c_int32),
                ("pad", C.c_int32), ("n_nev", C.c_int64), ("nev_pos", C.c_void_p), ("nev_pair", C.c_void_p),
                ("n_pairs", C.c_int64), ("pair_mm", C.c_void_p), ("start", C.c_int32),
                ("L", C.c_int32), ("ref", C.c_void_p), ("n_splits", C.c_int32), ("splits", C.c_void_p),
                ("M", C.c_int32), ("pad2", C.c_int32)]


class IsbReadsCompact(C.Structure):
    _fields_ = [("n_segs", C.c_int64), ("seg_start", C.c_void_p), ("seg_len", C.c_void_p), ("seg_pair", C.c_void_p),
                ("n_units", C.c_int64), ("base2", C.c_void_p), ("pass_", C.c_void_p), ("max_seg_len", C.c_int32),
                ("pad", C.c_int32), ("n_nev", C.c_int64), ("nev_pos", C.c_void_p), ("nev_pair", C.c_void_p),
                ("n_pairs", C.c_int64), ("pair_mm", C.c_void_p), ("start", C.c_int32),
                ("L", C.c_int32), ("ref", C.c_void_p), ("n_splits", C.c_int32), ("splits", C.c_void_p),
                ("M", C.c_int32), ("pad2", C.c_int32)]


class IsbReadsDelta(C.Structure):
    _fields_ = [("n_segs", C.c_int64), ("seg_start", C.c_void_p), ("seg_len", C.c_void_p), ("seg_pair", C.c_void_p),
                ("n_units", C.c_int64), ("pass_", C.c_void_p), ("n_mis", C.c_int64), ("mis_word", C.c_void_p),
                ("mis_code", C.c_void_p), ("max_seg_len", C.c_int32), ("pad", C.c_int32), ("n_nev", C.c_int64),
                ("nev_pos", C.c_void_p), ("nev_pair", C.c_void_p), ("n_pairs", C.c_int64), ("pair_mm", C.c_void_p),
                ("start", C.c_int32), ("L", C.c_int32), ("ref", C.c_void_p), ("n_splits", C.c_int32),
                ("splits", C.c_void_p), ("M", C.c_int32), ("pad2", C.c_int32)]


class IsbColsBatch(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("grp_off", C.c_void_p), ("n_chunks", C.c_int64), ("words", C.c_void_p),
                ("ids", C.c_void_p), ("n_nev", C.c_int64), ("nev_pos", C.c_void_p), ("nev_pair", C.c_void_p),
                ("n_pairs", C.c_int64), ("pair_mm", C.c_void_p), ("start", C.c_int32), ("L", C.c_int32),
                ("ref", C.c_void_p), ("n_splits", C.c_int32), ("splits", C.c_void_p), ("M", C.c_int32),
                ("pad", C.c_int32)]


class IsbParams(C.Structure):
    _fields_ = [("min_cov", C.c_int32), ("min_snp", C.c_int32), ("min_qual", C.c_int32), ("flags", C.c_uint32),
                ("min_freq", C.c_double), ("rarefied_cov", C.c_int32), ("pad", C.c_int32), ("seed", C.c_uint64)]


class IsbResult(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("nmask", C.c_void_p), ("covT", C.c_void_p), ("clonT", C.c_void_p),
                ("site_flags", C.c_void_p), ("snv", C.c_void_p), ("snv_cap", C.c_int64), ("ld", C.c_void_p),
                ("ld_cap", C.c_int64), ("n_snv", C.c_int64), ("n_ld", C.c_int64), ("n_sites", C.c_int64),
                ("n_site_pairs", C.c_int64), ("clonTR", C.c_void_p)]


class IsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libinstrain_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen the in-tree library and declare prototypes.  Raises if it was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ISB_LIB_PATH", LIB_PATH)     # A/B builds of the same sources (tools/gpu_ab.sh)
    if not os.path.exists(path):
        raise ImportError("libinstrain_b200.so is not built (%s). Run `python -m instrain_b200.build` (needs nvcc); "
                          "instrain_b200 has no CPU fallback." % path)
    L = C.CDLL(path)
    vp, i32, i64, u32, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double
    L.isb_create.restype = vp
    L.isb_create.argtypes = [C.c_int, vp, C.c_int, C.c_int]
    L.isb_destroy.restype = None
    L.isb_destroy.argtypes = [vp]
    L.isb_last_error.restype = C.c_char_p
    L.isb_last_error.argtypes = [vp]
    L.isb_abi_version.restype = C.c_int
    L.isb_set_stream.restype = C.c_int
    L.isb_set_stream.argtypes = [vp, vp]
    L.isb_synchronize.restype = C.c_int
    L.isb_synchronize.argtypes = [vp]
    L.isb_row_counts_async.restype = C.c_int
    L.isb_row_counts_async.argtypes = [vp, vp]
    L.isb_launch_count.restype = i64
    L.isb_launch_count.argtypes = [vp]
    L.isb_pileup_counts.restype = C.c_int
    L.isb_pileup_counts.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, i32, i32, C.c_int, C.c_int, u32, vp, vp]
    L.isb_call_snvs.restype = C.c_int
    L.isb_call_snvs.argtypes = [vp, i32, C.c_int, vp, vp, vp, i32, C.c_int, dbl, vp, vp, vp, vp, i64, C.POINTER(i64)]
    L.isb_linkage.restype = C.c_int
    L.isb_linkage.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, i32, i32, C.c_int, C.c_int, vp, vp, vp, i32, vp,
                              C.c_int, vp, i64, C.POINTER(i64)]
    L.isb_selftest_division.restype = i64
    L.isb_selftest_division.argtypes = [vp, C.c_int, C.c_int]
    L.isb_enable_timing.restype = C.c_int
    L.isb_enable_timing.argtypes = [vp, C.c_int]
    L.isb_stage_times.restype = C.c_int
    L.isb_stage_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.isb_scaffold_summary.restype = C.c_int
    L.isb_scaffold_summary.argtypes = [vp, i32, C.c_int, vp, vp, vp, i32, vp, vp]
    L.isb_profile_batch.restype = C.c_int
    L.isb_profile_batch.argtypes = [vp, C.POINTER(IsbBatch), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    L.isb_profile_batch_packed.restype = C.c_int
    L.isb_profile_batch_packed.argtypes = [vp, C.POINTER(IsbPackedBatch), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    L.isb_pileup_reads.restype = C.c_int
    L.isb_pileup_reads.argtypes = [vp, C.POINTER(IsbReadsBatch), vp, vp]
    L.isb_profile_reads.restype = C.c_int
    L.isb_profile_reads.argtypes = [vp, C.POINTER(IsbReadsBatch), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    L.isb_profile_reads_compact.restype = C.c_int
    L.isb_profile_reads_compact.argtypes = [vp, C.POINTER(IsbReadsCompact), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    L.isb_profile_reads_delta.restype = C.c_int
    L.isb_profile_reads_delta.argtypes = [vp, C.POINTER(IsbReadsDelta), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    L.isb_reads_delta_host.restype = i64
    L.isb_reads_delta_host.argtypes = [i64, vp, vp, vp, vp, i64, i32, i32, vp, vp, i64, vp, vp, i64]
    L.isb_cols_from_reads.restype = C.c_int
    L.isb_cols_from_reads.argtypes = [vp, C.POINTER(IsbReadsBatch), vp, C.POINTER(i64), vp, vp, i64]
    L.isb_cols_from_reads_host.restype = i64
    L.isb_cols_from_reads_host.argtypes = [i64, vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, i64]
    L.isb_pileup_cols.restype = C.c_int
    L.isb_pileup_cols.argtypes = [vp, C.POINTER(IsbColsBatch), vp, vp]
    L.isb_profile_cols.restype = C.c_int
    L.isb_profile_cols.argtypes = [vp, C.POINTER(IsbColsBatch), C.POINTER(IsbParams), C.POINTER(IsbResult)]
    _lib = L
    return L


def ptr(a):
    """Host numpy array / torch tensor (host or CUDA) / int / None -> c_void_p value."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError("unsupported buffer type %r" % type(a))
