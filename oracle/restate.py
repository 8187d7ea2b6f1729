"""numpy/ctypes front-end of the C restatement in oracle/oracle.c (test infrastructure only).

`profile_events` is the oracle for the whole path on one coordinate space ("batch"):
events (position-major) + reference codes + null-model LUT + split table  ->
counts[L,M,4], nmask[L], covT[L,M], clonT[L,M], site_flags[L], SNV rows, LD rows.
Each stage follows the reference functions cited in oracle/oracle.c's header.
"""
import ctypes as C

import numpy as np

from . import build as _build

SNV_DT = np.dtype([("pos", "<i4"), ("cnt", "<i4", (4,)), ("mm", "<i4"), ("ref", "u1"), ("con", "u1"),
                   ("var", "u1"), ("allele_count", "u1"), ("cls", "u1"), ("cryptic", "u1"), ("pad", "u1", (2,))])
LD_DT = np.dtype([("pos_a", "<i4"), ("pos_b", "<i4"), ("mm", "<i4"), ("c_AB", "<i4"), ("c_Ab", "<i4"),
                  ("c_aB", "<i4"), ("c_ab", "<i4"), ("allele_A", "u1"), ("allele_a", "u1"), ("allele_B", "u1"),
                  ("allele_b", "u1"), ("r2", "<f8"), ("d_prime", "<f8")])
assert SNV_DT.itemsize == 32 and LD_DT.itemsize == 48
# rows as the CUDA library returns them: + the two re-drawn columns (oracle/rarefied.py)
LD_DT_FULL = np.dtype(LD_DT.descr + [("r2_normalized", "<f8"), ("d_prime_normalized", "<f8")])
assert LD_DT_FULL.itemsize == 64

CLASS_NAMES = ["AmbiguousReference", "DivergentSite", "SNS", "SNV", "con_SNV", "pop_SNV"]
BASES = "ACTG"
SITE_ANYSNP = 0x10

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        L.orc_pileup_counts.restype = C.c_int
        L.orc_call_snvs.restype = C.c_int64
        L.orc_linkage.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sort_events(ev):
    """Stable position-major ordering of BAM-order events (column order = file order)."""
    order = np.argsort(ev["ref_pos"], kind="stable")
    out = dict(ev)
    for k in ("ref_pos", "base", "qual", "read_id"):
        out[k] = np.ascontiguousarray(ev[k][order])
    return out


def encode_ref(seq):
    """Reference string -> codes 0..3 (A,C,T,G), 4 otherwise."""
    lut = np.full(256, 4, dtype=np.uint8)
    for i, b in enumerate(BASES):
        lut[ord(b)] = i
    return lut[np.frombuffer(seq.encode(), dtype=np.uint8)]


def lut_from_model(model, n_lut=10000):
    """Null-model dict (snv_utilities.py:14-38) -> (int32 lut[n_lut] with -1 = key absent, default)."""
    lut = np.full(n_lut, -1, dtype=np.int32)
    for k, v in model.items():
        if 0 <= k < n_lut:
            lut[k] = v
        elif k >= n_lut:
            raise ValueError("null model key beyond n_lut")
    return lut, int(model[-1])


def pileup_counts(ev, start, L, M, min_qual=30):
    n = len(ev["ref_pos"])
    counts = np.zeros((L, M, 4), dtype=np.int32)
    nmask = np.zeros(L, dtype=np.uint64)
    rc = lib().orc_pileup_counts(C.c_int64(n), _p(ev["ref_pos"]), _p(ev["base"]), _p(ev["qual"]), _p(ev["read_id"]),
                                 _p(np.ascontiguousarray(ev["pair_mm"], dtype=np.int32)), C.c_int32(start),
                                 C.c_int32(L), C.c_int(M), C.c_int(min_qual), _p(counts), _p(nmask))
    if rc != 0:
        raise ValueError("pair_mm outside [0, M)")
    return counts, nmask


def call_snvs(counts, nmask, ref_codes, lut, lut_default, start=0, min_cov=5, min_freq=0.05):
    L, M, _ = counts.shape
    covT = np.zeros((L, M), dtype=np.int32)
    clonT = np.zeros((L, M), dtype=np.float32)
    flags = np.zeros(L, dtype=np.uint8)
    cap = max(1024, L * 2)
    while True:
        rows = np.zeros(cap, dtype=SNV_DT)
        n = lib().orc_call_snvs(C.c_int32(L), C.c_int(M), _p(counts), _p(nmask), _p(ref_codes), _p(lut),
                                C.c_int(len(lut)), C.c_int(lut_default), C.c_int(min_cov), C.c_double(min_freq),
                                C.c_int32(start), _p(covT), _p(clonT), _p(flags), _p(rows), C.c_int64(cap))
        if n >= 0:
            return covT, clonT, flags, rows[:n].copy()
        cap = -n


def linkage(ev, counts, nmask, flags, splits, start=0, min_snp=20, min_qual=30):
    L, M, _ = counts.shape
    splits = np.ascontiguousarray(splits, dtype=np.int32).reshape(-1, 2)
    pair_mm = np.ascontiguousarray(ev["pair_mm"], dtype=np.int32)
    cap = 1 << 16
    while True:
        rows = np.zeros(cap, dtype=LD_DT)
        n = lib().orc_linkage(C.c_int64(len(ev["ref_pos"])), _p(ev["ref_pos"]), _p(ev["base"]), _p(ev["qual"]),
                              _p(ev["read_id"]), _p(pair_mm), C.c_int64(len(pair_mm)), C.c_int32(start),
                              C.c_int32(L), C.c_int(M), C.c_int(min_qual), _p(counts), _p(nmask), _p(flags),
                              C.c_int(len(splits)), _p(splits), C.c_int(min_snp), _p(rows), C.c_int64(cap))
        if n >= 0:
            return rows[:n].copy()
        cap = -n


def profile_mt(ev, ref_codes, lut, lut_default, splits, M=None, ref_start=0, min_cov=5, min_freq=0.05, min_snp=20,
               min_qual=30, chunk_splits=4, n_threads=0):
    """All three stages over split chunks on `n_threads` OpenMP threads (CPU-baseline driver): returns (#snv, #ld)."""
    pair_mm = np.ascontiguousarray(ev["pair_mm"], dtype=np.int32)
    if M is None:
        M = int(pair_mm.max()) + 1 if len(pair_mm) else 1
    splits = np.ascontiguousarray(splits, dtype=np.int32).reshape(-1, 2)
    L = lib()
    L.orc_profile_mt.restype = C.c_int
    n_snv, n_ld = C.c_int64(0), C.c_int64(0)
    rc = L.orc_profile_mt(C.c_int64(len(ev["ref_pos"])), _p(ev["ref_pos"]), _p(ev["base"]), _p(ev["qual"]),
                          _p(ev["read_id"]), _p(pair_mm), C.c_int64(len(pair_mm)), C.c_int(M), C.c_int(min_qual),
                          _p(ref_codes), C.c_int32(ref_start), _p(lut), C.c_int(len(lut)), C.c_int(lut_default),
                          C.c_int(min_cov), C.c_double(min_freq), C.c_int(len(splits)), _p(splits), C.c_int(min_snp),
                          C.c_int(chunk_splits), C.c_int(n_threads), C.byref(n_snv), C.byref(n_ld))
    if rc:
        raise RuntimeError("orc_profile_mt failed (%d)" % rc)
    return n_snv.value, n_ld.value


def profile_events(ev, ref_codes, lut, lut_default, splits, start=0, M=None, min_cov=5, min_freq=0.05,
                   min_snp=20, min_qual=30, do_linkage=True, rarefied_coverage=50, seed=0):
    """Whole hot path on one coordinate space. `ev` must be position-major (see sort_events).  The re-drawn outputs
    (clonTR, r2_normalized / d_prime_normalized of the LD rows) come from oracle/rarefied.py with the given seed."""
    from . import rarefied
    L = len(ref_codes)
    if M is None:
        M = int(ev["pair_mm"].max()) + 1 if len(ev["pair_mm"]) else 1
    counts, nmask = pileup_counts(ev, start, L, M, min_qual)
    covT, clonT, flags, snv = call_snvs(counts, nmask, ref_codes, lut, lut_default, start, min_cov, min_freq)
    ld = linkage(ev, counts, nmask, flags, splits, start, min_snp, min_qual) if do_linkage else np.zeros(0, LD_DT)
    full = np.zeros(len(ld), dtype=LD_DT_FULL)
    for k in LD_DT.names:
        full[k] = ld[k]
    full["r2_normalized"], full["d_prime_normalized"] = rarefied.normalized_ld(ld, min_snp, seed)
    return dict(counts=counts, nmask=nmask, covT=covT, clonT=clonT, site_flags=flags, snv=snv, ld=full,
                clonTR=rarefied.clonTR(counts, nmask, rarefied_coverage, seed, start))
