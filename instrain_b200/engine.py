"""Thin Python face of the C-ABI: one `Engine` per GPU.

Everything here goes through libinstrain_b200.so (instrain_b200/_cabi.py); there is no other compute path.
Buffers may be numpy arrays (host; staged by the library) or CUDA torch tensors (used in place).
"""
import ctypes as C

import numpy as np

from . import _cabi
from .null_model import load_lut


def _is_cuda(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _pair_mm_u8(pair_mm):
    """The ABI takes uint8 mm values (ISB_MAX_MM = 64); host arrays of another integer dtype are narrowed."""
    if pair_mm is None or not isinstance(pair_mm, np.ndarray) or pair_mm.dtype == np.uint8:
        return pair_mm
    if len(pair_mm) and (pair_mm.min() < 0 or pair_mm.max() > 255):
        raise ValueError("pair_mm outside [0, 255]")
    return np.ascontiguousarray(pair_mm, dtype=np.uint8)


class Engine:
    """Owns an isb_ctx (device context + scratch).  Mirrors the per-worker state of the reference's
    split_profile_worker (inStrain/profile/profile_utilities.py:37-90): null model + open handle."""

    def __init__(self, device=0, null_lut=None, lut_default=None, model_file=None, fdr=1e-6):
        self.lib = _cabi.load()
        if null_lut is None:
            null_lut, lut_default = load_lut(model_file, fdr)
        null_lut = np.ascontiguousarray(null_lut, dtype=np.int32)
        self.device = device
        self.ctx = self.lib.isb_create(device, null_lut.ctypes.data, len(null_lut), int(lut_default))
        if not self.ctx:
            raise _cabi.IsbError(_cabi.ISB_ERR_CUDA, self.lib.isb_last_error(None).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.isb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise _cabi.IsbError(rc, self.lib.isb_last_error(self.ctx).decode())
        return rc

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.isb_set_stream(self.ctx, cuda_stream_ptr))

    def synchronize(self):
        self._check(self.lib.isb_synchronize(self.ctx))

    @property
    def launch_count(self):
        return int(self.lib.isb_launch_count(self.ctx))

    def enable_timing(self, on=True):
        self._check(self.lib.isb_enable_timing(self.ctx, 1 if on else 0))

    def stage_times(self):
        """(ms[3], calls[3]) of K1 / K2 / K3 device time since the last call (CUDA events on the context stream)."""
        ms = (C.c_double * 3)()
        calls = (C.c_int64 * 3)()
        self._check(self.lib.isb_stage_times(self.ctx, ms, calls))
        return list(ms), list(calls)

    # ---- stage K1 ----------------------------------------------------------------------------------------------
    def pileup_counts(self, ev, start, L, M, min_qual=30, any_order=False, counts=None, nmask=None):
        """ev: dict(ref_pos, base, qual, read_id, pair_mm). Returns (counts[L,M,4] int32, nmask[L] uint64)."""
        n = len(ev["ref_pos"])
        pair_mm = _pair_mm_u8(ev.get("pair_mm"))
        if counts is None:
            counts = np.empty((L, M, 4), dtype=np.int32)
        if nmask is None:
            nmask = np.empty(L, dtype=np.uint64)
        p = _cabi.ptr
        self._check(self.lib.isb_pileup_counts(
            self.ctx, n, p(ev["ref_pos"]), p(ev["base"]), p(ev["qual"]), p(ev.get("read_id")),
            0 if pair_mm is None else len(pair_mm), p(pair_mm), start, L, M, min_qual,
            _cabi.ISB_K1_ANY_ORDER if any_order else 0, p(counts), p(nmask)))
        return counts, nmask

    # ---- stage K2 ----------------------------------------------------------------------------------------------
    def _reads_batch(self, rd, pair_mm, start, L, ref_codes, splits, M):
        p = _cabi.ptr
        n_pairs = len(pair_mm) if pair_mm is not None else 0
        return _cabi.IsbReadsBatch(int(rd["n_segs"]), p(rd["seg_start"]), p(rd["seg_len"]), p(rd["seg_pair"]),
                                   p(rd["seg_word"]), int(rd["n_words"]), p(rd["words"]), int(rd["max_seg_len"]), 0,
                                   len(rd["nev_pos"]), p(rd["nev_pos"]), p(rd["nev_pair"]), n_pairs, p(pair_mm), start, L, p(ref_codes), 0 if splits is None else len(splits),
                                   p(splits), M, 0)

    def pileup_reads(self, rd, pair_mm, start, L, M, counts=None, nmask=None):
        """K1r alone: counts[L, M, 4] (+ nmask[L]) from a read-major batch (instrain_b200.reads)."""
        pair_mm = _pair_mm_u8(pair_mm)
        if counts is None:
            counts = np.empty((L, M, 4), dtype=np.int32)
        if nmask is None:
            nmask = np.empty(L, dtype=np.uint64)
        batch = self._reads_batch(rd, pair_mm, start, L, None, None, M)
        self._check(self.lib.isb_pileup_reads(self.ctx, C.byref(batch), _cabi.ptr(counts), _cabi.ptr(nmask)))
        return counts, nmask

    # ---- column words (instrain_b200.cols) ---------------------------------------------------------------------------
    def _cols_batch(self, cd, pair_mm, start, L, ref_codes, splits, M):
        p = _cabi.ptr
        n_pairs = len(pair_mm) if pair_mm is not None else 0
        return _cabi.IsbColsBatch(int(cd["n_groups"]), p(cd["grp_off"]), int(cd["n_chunks"]), p(cd["words"]), p(cd["ids"]),
                                  len(cd["nev_pos"]), p(cd["nev_pos"]), p(cd["nev_pair"]), n_pairs, p(pair_mm), start, L,
                                  p(ref_codes), 0 if splits is None else len(splits), p(splits), M, 0)

    def pileup_cols(self, cd, pair_mm, start, L, M, counts=None, nmask=None):
        """K1c alone: counts[L, M, 4] (+ nmask[L]) from a column-word batch (instrain_b200.cols)."""
        pair_mm = _pair_mm_u8(pair_mm)
        if counts is None:
            counts = np.empty((L, M, 4), dtype=np.int32)
        if nmask is None:
            nmask = np.empty(L, dtype=np.uint64)
        batch = self._cols_batch(cd, pair_mm, start, L, None, None, M)
        self._check(self.lib.isb_pileup_cols(self.ctx, C.byref(batch), _cabi.ptr(counts), _cabi.ptr(nmask)))
        return counts, nmask

    def cols_from_reads(self, rd, L, start=0):
        """Device-side layout conversion (isb_cols_from_reads): read-major batch -> column-word batch (numpy arrays)."""
        batch = self._reads_batch(rd, None, start, L, None, None, 1)
        from .cols import GROUP, CHUNK
        n_groups = (L + GROUP - 1) // GROUP
        grp_off = np.zeros(n_groups + 1, dtype=np.int64)
        n = C.c_int64(0)
        p = _cabi.ptr
        self._check(self.lib.isb_cols_from_reads(self.ctx, C.byref(batch), p(grp_off), C.byref(n), None, None, 0))
        words = np.empty(n.value * CHUNK, dtype=np.uint32)
        ids = np.empty(n.value * CHUNK, dtype=np.int32)
        self._check(self.lib.isb_cols_from_reads(self.ctx, C.byref(batch), p(grp_off), C.byref(n), p(words), p(ids), n.value))
        return dict(n_groups=n_groups, grp_off=grp_off, n_chunks=int(n.value), words=words, ids=ids,
                    nev_pos=rd["nev_pos"], nev_pair=rd["nev_pair"])

    def call_snvs(self, counts, nmask, ref_codes, start=0, min_cov=5, min_freq=0.05, cap=None):
        L, M = counts.shape[0], counts.shape[1]
        covT = np.empty((L, M), dtype=np.int32)
        clonT = np.empty((L, M), dtype=np.float32)
        flags = np.empty(L, dtype=np.uint8)
        cap = max(1024, L // 8) if cap is None else cap
        p = _cabi.ptr
        while True:
            rows = np.empty(cap, dtype=_cabi.SNV_DT)
            n = C.c_int64(0)
            rc = self._check(self.lib.isb_call_snvs(self.ctx, L, M, p(counts), p(nmask), p(ref_codes), start, min_cov,
                                                    float(min_freq), p(covT), p(clonT), p(flags), p(rows), cap,
                                                    C.byref(n)), allow=(_cabi.ISB_ERR_CAPACITY,))
            if rc == 0:
                return covT, clonT, flags, rows[:n.value].copy()
            cap = int(n.value)

    # ---- stage K3 ----------------------------------------------------------------------------------------------
    def linkage(self, ev, counts, nmask, flags, splits, start=0, min_snp=20, min_qual=30, cap=None):
        L, M = counts.shape[0], counts.shape[1]
        splits = np.ascontiguousarray(splits, dtype=np.int32).reshape(-1, 2)
        pair_mm = _pair_mm_u8(ev.get("pair_mm"))
        cap = 1 << 16 if cap is None else cap
        p = _cabi.ptr
        while True:
            rows = np.empty(cap, dtype=_cabi.LD_DT)
            n = C.c_int64(0)
            rc = self._check(self.lib.isb_linkage(
                self.ctx, len(ev["ref_pos"]), p(ev["ref_pos"]), p(ev["base"]), p(ev["qual"]), p(ev["read_id"]),
                0 if pair_mm is None else len(pair_mm), p(pair_mm), start, L, M, min_qual, p(counts), p(nmask),
                p(flags), len(splits), p(splits), min_snp, p(rows), cap, C.byref(n)),
                allow=(_cabi.ISB_ERR_CAPACITY,))
            if rc == 0:
                return rows[:n.value].copy()
            cap = int(n.value)

    # ---- stage K4 (merge-stage summary reductions) -----------------------------------------------------------------
    def scaffold_summary(self, covT, clonT, nmask, scaffold_off):
        """covT int32[L,M], clonT float32[L,M], nmask uint64[L] or None, scaffold_off int32[n+1] -> SUMMARY_DT[n*M]."""
        L, M = covT.shape
        off = np.ascontiguousarray(scaffold_off, dtype=np.int32)
        n_sc = len(off) - 1
        if off[0] != 0 or off[-1] != L:
            raise ValueError("scaffold_off must start at 0 and end at L")
        rows = np.zeros(n_sc * M, dtype=_cabi.SUMMARY_DT)
        p = _cabi.ptr
        self._check(self.lib.isb_scaffold_summary(self.ctx, L, M, p(covT), p(clonT), p(nmask), n_sc, p(off), p(rows)))
        return rows

    # ---- whole path --------------------------------------------------------------------------------------------
    def profile_batch(self, ev, ref_codes, splits, start=0, M=None, min_cov=5, min_freq=0.05, min_snp=20,
                      min_qual=30, skip_linkage=False, want=("covT", "clonT", "site_flags", "snv", "ld"),
                      snv_cap=None, ld_cap=None, packed=None, pipeline=False, reads=None, cols=None, rarefied_coverage=50,
                      seed=0):
        """Run K1 -> K2 -> K3 on one batch with HOST (numpy) or CUDA-tensor inputs; numpy outputs.

        `want` selects which outputs are copied back ("counts", "nmask", "covT", "clonT", "clonTR", "site_flags", "snv",
        "ld").  "clonTR" = rarefied clonality at `rarefied_coverage` (--rarefied_coverage, argumentParser.py:168); it and
        the normalized linkage columns are drawn from a counter-based generator keyed by `seed` (reproducible).
        """
        L = len(ref_codes)
        pair_mm = _pair_mm_u8(ev["pair_mm"])
        if M is None:
            M = int(pair_mm.max()) + 1 if len(pair_mm) else 1
        if M > _cabi.ISB_MAX_MM:
            raise ValueError("mm levels M=%d exceeds ISB_MAX_MM=%d" % (M, _cabi.ISB_MAX_MM))
        splits = np.ascontiguousarray(splits, dtype=np.int32).reshape(-1, 2)
        p = _cabi.ptr
        if cols is not None:                        # column words (instrain_b200.cols)
            batch = self._cols_batch(cols, pair_mm, start, L, ref_codes, splits, M)
            entry = self.lib.isb_profile_cols
        elif reads is not None and "mis_word" in reads:   # reference-delta transfer format (instrain_b200.reads.delta_reads)
            batch = _cabi.IsbReadsDelta(int(reads["n_segs"]), p(reads["seg_start"]), p(reads["seg_len"]), p(reads["seg_pair"]),
                                        int(reads["n_units"]), p(reads["pass"]), len(reads["mis_word"]), p(reads["mis_word"]),
                                        p(reads["mis_code"]), int(reads["max_seg_len"]), 0, len(reads["nev_pos"]),
                                        p(reads["nev_pos"]), p(reads["nev_pair"]), len(pair_mm), p(pair_mm), start, L,
                                        p(ref_codes), len(splits), p(splits), M, 0)
            entry = self.lib.isb_profile_reads_delta
        elif reads is not None and "base2" in reads:  # compact transfer format (instrain_b200.reads.compact_reads)
            batch = _cabi.IsbReadsCompact(int(reads["n_segs"]), p(reads["seg_start"]), p(reads["seg_len"]),
                                          p(reads["seg_pair"]), int(reads["n_units"]), p(reads["base2"]), p(reads["pass"]),
                                          int(reads["max_seg_len"]), 0, len(reads["nev_pos"]), p(reads["nev_pos"]),
                                          p(reads["nev_pair"]), len(pair_mm), p(pair_mm), start, L, p(ref_codes),
                                          len(splits), p(splits), M, 0)
            entry = self.lib.isb_profile_reads_compact
        elif reads is not None:                    # read-major aligned segments (instrain_b200.reads)
            batch = self._reads_batch(reads, pair_mm, start, L, ref_codes, splits, M)
            entry = self.lib.isb_profile_reads
        elif packed is not None:                   # packed transfer format (instrain_b200.packed.encode_packed)
            if packed["min_qual"] != min_qual:
                raise ValueError("packed batch was encoded with min_qual=%d" % packed["min_qual"])
            batch = _cabi.IsbPackedBatch(packed["n_events"], p(packed["pos_off"]), p(packed["id_base"]), p(packed["bqd"]),
                                         len(packed["esc_evt"]), p(packed["esc_evt"]), p(packed["esc_id"]), len(pair_mm),
                                         p(pair_mm), start, L, p(ref_codes), len(splits), p(splits), M, min_qual)
            entry = self.lib.isb_profile_batch_packed
        else:
            batch = _cabi.IsbBatch(len(ev["ref_pos"]), p(ev["ref_pos"]), p(ev["base"]), p(ev["qual"]), p(ev["read_id"]),
                                   len(pair_mm), p(pair_mm), start, L, p(ref_codes), len(splits), p(splits), M)
            entry = self.lib.isb_profile_batch
        prm = _cabi.IsbParams(min_cov, min_snp, min_qual, (_cabi.ISB_SKIP_LINKAGE if skip_linkage else 0) |
                              (_cabi.ISB_PIPELINE if pipeline else 0), float(min_freq),
                              int(rarefied_coverage) if "clonTR" in want else 0, 0, int(seed) & 0xFFFFFFFFFFFFFFFF)
        out = {}
        if "counts" in want:
            out["counts"] = np.empty((L, M, 4), dtype=np.int32)
        if "nmask" in want:
            out["nmask"] = np.empty(L, dtype=np.uint64)
        if "covT" in want:
            out["covT"] = np.empty((L, M), dtype=np.int32)
        if "clonT" in want:
            out["clonT"] = np.empty((L, M), dtype=np.float32)
        if "site_flags" in want:
            out["site_flags"] = np.empty(L, dtype=np.uint8)
        if "clonTR" in want:
            out["clonTR"] = np.empty((L, M), dtype=np.float32)
        snv_cap = max(1024, L // 8) if snv_cap is None else snv_cap
        ld_cap = (1 << 16) if ld_cap is None else ld_cap
        while True:
            snv = np.empty(snv_cap, dtype=_cabi.SNV_DT) if "snv" in want else None
            ld = np.empty(ld_cap, dtype=_cabi.LD_DT) if ("ld" in want and not skip_linkage) else None
            res = _cabi.IsbResult(p(out.get("counts")), p(out.get("nmask")), p(out.get("covT")), p(out.get("clonT")),
                                  p(out.get("site_flags")), p(snv), snv_cap if snv is not None else 0, p(ld),
                                  ld_cap if ld is not None else 0, 0, 0, 0, 0, p(out.get("clonTR")))
            rc = self._check(entry(self.ctx, C.byref(batch), C.byref(prm), C.byref(res)),
                             allow=(_cabi.ISB_ERR_CAPACITY,))
            if rc == 0:
                break
            snv_cap = max(snv_cap, int(res.n_snv))
            ld_cap = max(ld_cap, int(res.n_ld))
        if snv is not None:
            out["snv"] = snv[:res.n_snv].copy()
        if ld is not None:
            out["ld"] = ld[:res.n_ld].copy()
        elif "ld" in want:
            out["ld"] = np.zeros(0, dtype=_cabi.LD_DT)
        out["n_snv"], out["n_ld"], out["n_sites"], out["n_site_pairs"] = (int(res.n_snv), int(res.n_ld),
                                                                          int(res.n_sites), int(res.n_site_pairs))
        out["M"] = M
        return out
