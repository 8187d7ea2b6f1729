"""Device-side synthetic metagenome (bench / test support; ctypes face of lib/libisb_synth.so).

Generates BASELINE.json's synthetic configurations directly in HBM as position-major event columns
(instrain_b200/csrc_synth/isb_synth.cu).  Not part of the hot path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libisb_synth.so")


class _Params(C.Structure):
    _fields_ = [("L", C.c_int64), ("n_scaffolds", C.c_int32), ("coverage", C.c_int32), ("snv_density", C.c_double),
                ("seed", C.c_uint64), ("skip_mm", C.c_int32), ("pad", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libisb_synth.so is not built; run `python -m instrain_b200.build`")
        _lib = C.CDLL(LIB_PATH)
        _lib.isbs_last_error.restype = C.c_char_p
        _lib.isbs_plan.restype = C.c_int
        _lib.isbs_plan.argtypes = [C.c_int, C.POINTER(_Params), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.isbs_fill.restype = C.c_int
        _lib.isbs_fill.argtypes = [C.c_void_p] * 6
        _lib.isbs_fill_reads.restype = C.c_int
        _lib.isbs_fill_reads.argtypes = [C.c_void_p] * 7 + [C.c_int]
        _lib.isbs_reads_words.restype = C.c_int64
        _lib.isbs_reads_words.argtypes = [C.c_int64]
        _lib.isbs_free.restype = C.c_int
        _lib.isbs_set_seg_words.restype = C.c_int
        _lib.isbs_set_seg_words.argtypes = [C.c_int]
    return _lib


def iterate_splits(s_len, window_len=10000):
    """Split geometry of the reference (inStrain/profile/fasta.py:56-73): 0-based, double-inclusive."""
    n_chunks = s_len // window_len + 1
    chunk_len = int(s_len / n_chunks)
    out, start, end = [], 0, 0
    for i in range(n_chunks):
        if i + 1 == n_chunks:
            out.append((start, s_len - 1))
        else:
            end += chunk_len
            out.append((start, end - 1))
            start += chunk_len
    return out


def batch_splits(L, n_scaffolds, window_len=10000):
    one = np.array(iterate_splits(L, window_len), dtype=np.int64)
    return np.concatenate([one + s * L for s in range(n_scaffolds)]).astype(np.int32)


READLEN = 150


def generate(device, L, n_scaffolds, coverage, snv_density, seed, skip_mm=True, window_len=10000, events=True, reads=False,
             min_qual=30, seg_words=None):
    """Returns a dict of CUDA torch tensors + splits (CUDA int32).  events=True: position-major event columns (ref_pos,
    base, qual, read_id); reads=True: the SAME fragments as a read-major batch under key "reads" (seg_start, seg_len,
    seg_pair, seg_word, words, ...; instrain_b200/reads.py layout).  Always: pair_mm, ref_codes."""
    import torch
    lib = _load()
    spw = (READLEN + 14) // 8 + 1 if seg_words is None else int(seg_words)      # position-aligned data words + separator
    if lib.isbs_set_seg_words(spw) != 0:
        raise ValueError("seg_words must be %d or %d" % ((READLEN + 14) // 8 + 1, (READLEN + 14) // 8 + 2))
    if reads and L % 8:
        raise ValueError("scaffold length must be a multiple of 8 (the read-major stream is position-aligned)")
    prm = _Params(L, n_scaffolds, coverage, float(snv_density), seed, 1 if skip_mm else 0, 0)
    n_ev, n_pairs = C.c_int64(0), C.c_int64(0)
    if lib.isbs_plan(device, C.byref(prm), C.byref(n_ev), C.byref(n_pairs)) != 0:
        raise RuntimeError("isbs_plan: " + lib.isbs_last_error().decode())
    dev = torch.device("cuda", device)
    n, npairs, Ltot = n_ev.value, n_pairs.value, L * n_scaffolds
    pad = 16                                    # keep every column readable in whole 16-byte granules
    rd = None
    if reads:
        n_segs = 2 * npairs
        n_words = int(lib.isbs_reads_words(n_segs))
        rd = dict(n_segs=n_segs, n_words=n_words, max_seg_len=READLEN, seg_words=spw,
                  seg_start=torch.empty(max(n_segs, 1), dtype=torch.int32, device=dev)[:n_segs],
                  seg_len=torch.empty(max(n_segs, 1), dtype=torch.int16, device=dev)[:n_segs],
                  seg_pair=torch.empty(max(n_segs, 1), dtype=torch.int32, device=dev)[:n_segs],
                  seg_word=torch.empty(max(n_segs, 1), dtype=torch.int64, device=dev)[:n_segs],
                  words=torch.empty(n_words, dtype=torch.int32, device=dev),
                  nev_pos=torch.empty(0, dtype=torch.int32, device=dev), nev_pair=torch.empty(0, dtype=torch.int32, device=dev))
        pair_mm = torch.empty(max(npairs, 1), dtype=torch.uint8, device=dev)[:npairs]
        ref_codes = torch.empty(Ltot, dtype=torch.uint8, device=dev)
        if lib.isbs_fill_reads(rd["seg_start"].data_ptr(), rd["seg_len"].data_ptr(), rd["seg_pair"].data_ptr(),
                               rd["seg_word"].data_ptr(), rd["words"].data_ptr(), pair_mm.data_ptr(),
                               ref_codes.data_ptr(), min_qual) != 0:
            raise RuntimeError("isbs_fill_reads: " + lib.isbs_last_error().decode())
        if not events:
            lib.isbs_free()
            return dict(reads=rd, pair_mm=pair_mm, ref_codes=ref_codes, L=L, n_scaffolds=n_scaffolds, n_events=n,
                        splits=torch.from_numpy(batch_splits(L, n_scaffolds, window_len)).to(dev))
    out = dict(
        ref_pos=torch.empty(n + pad, dtype=torch.int32, device=dev)[:n],
        base=torch.empty(n + pad, dtype=torch.uint8, device=dev)[:n],
        qual=torch.empty(n + pad, dtype=torch.uint8, device=dev)[:n],
        read_id=torch.empty(n + pad, dtype=torch.int32, device=dev)[:n],
        pair_mm=torch.empty(max(npairs, 1), dtype=torch.uint8, device=dev)[:npairs],
        ref_codes=torch.empty(Ltot, dtype=torch.uint8, device=dev),
    )
    if lib.isbs_fill(out["ref_pos"].data_ptr(), out["base"].data_ptr(), out["qual"].data_ptr(),
                     out["read_id"].data_ptr(), out["pair_mm"].data_ptr(), out["ref_codes"].data_ptr()) != 0:
        raise RuntimeError("isbs_fill: " + lib.isbs_last_error().decode())
    out["splits"] = torch.from_numpy(batch_splits(L, n_scaffolds, window_len)).to(dev)
    out["L"], out["n_scaffolds"], out["n_events"] = L, n_scaffolds, n
    if rd is not None:
        out["reads"] = rd
    return out


def reads_to_host(d, lo_scaffold=0, n_scaffolds=1):
    """Cut scaffolds [lo, lo+n) out of a generated read-major data set as a self-contained host (numpy) batch, coordinates,
    pair ids and word offsets re-based to 0 (uniform READLEN segments: the word stream of a range is contiguous)."""
    import torch
    L, rd = d["L"], d["reads"]
    p_lo, p_hi = lo_scaffold * L, (lo_scaffold + n_scaffolds) * L
    bounds = torch.searchsorted(rd["seg_start"], torch.tensor([p_lo, p_hi], dtype=torch.int32, device=rd["seg_start"].device))
    s_lo, s_hi = int(bounds[0]), int(bounds[1])
    n = s_hi - s_lo
    pair = rd["seg_pair"][s_lo:s_hi]
    id_lo, id_hi = (int(pair.min()), int(pair.max()) + 1) if n else (0, 0)
    w_lo = int(rd["seg_word"][s_lo]) - 1 if n else 0
    spw = rd["seg_words"]
    n_words = (1 + n * spw + 3) // 4 * 4
    words = np.zeros(n_words, dtype=np.uint32)
    words[:1 + n * spw] = rd["words"][w_lo:w_lo + 1 + n * spw].cpu().numpy().view(np.uint32)
    spl = d["splits"]
    sel = (spl[:, 0] >= p_lo) & (spl[:, 0] < p_hi)
    return dict(reads=dict(n_segs=n, n_words=n_words, max_seg_len=READLEN,
                           seg_start=(rd["seg_start"][s_lo:s_hi] - p_lo).cpu().numpy(),
                           seg_len=rd["seg_len"][s_lo:s_hi].cpu().numpy().view(np.uint16),
                           seg_pair=(pair - id_lo).cpu().numpy(),
                           seg_word=(rd["seg_word"][s_lo:s_hi] - w_lo).cpu().numpy(), words=words,
                           nev_pos=np.zeros(0, np.int32), nev_pair=np.zeros(0, np.int32)),
                pair_mm=d["pair_mm"][id_lo:id_hi].cpu().numpy(), ref_codes=d["ref_codes"][p_lo:p_hi].cpu().numpy(),
                splits=(spl[sel] - p_lo).cpu().numpy())


def to_host_batch(d, lo_scaffold=0, n_scaffolds=1):
    """Cut scaffolds [lo, lo+n) out of a generated data set as a self-contained host (numpy) batch with coordinates and
    pair ids re-based to 0 -- the form the oracle and the host-buffer (e2e) path take."""
    import torch
    L = d["L"]
    p_lo, p_hi = lo_scaffold * L, (lo_scaffold + n_scaffolds) * L
    bounds = torch.searchsorted(d["ref_pos"], torch.tensor([p_lo, p_hi], dtype=torch.int32, device=d["ref_pos"].device))
    e_lo, e_hi = int(bounds[0]), int(bounds[1])
    rid = d["read_id"][e_lo:e_hi]
    id_lo, id_hi = (int(rid.min()), int(rid.max()) + 1) if e_hi > e_lo else (0, 0)
    spl = d["splits"]
    sel = (spl[:, 0] >= p_lo) & (spl[:, 0] < p_hi)
    return dict(
        ref_pos=(d["ref_pos"][e_lo:e_hi] - p_lo).cpu().numpy(), base=d["base"][e_lo:e_hi].cpu().numpy(),
        qual=d["qual"][e_lo:e_hi].cpu().numpy(), read_id=(rid - id_lo).cpu().numpy(),
        pair_mm=d["pair_mm"][id_lo:id_hi].cpu().numpy(), ref_codes=d["ref_codes"][p_lo:p_hi].cpu().numpy(),
        splits=(spl[sel] - p_lo).cpu().numpy())


def reads_to_cols_device(eng, d):
    """Lay the generated read-major data set out as COLUMN WORDS in HBM (include/instrain_b200.h, isb_cols_batch) with
    the library's device-side conversion (isb_cols_from_reads, device pointers in and out; two calls: sizing, fill)."""
    import torch
    from . import _cabi
    from .cols import CHUNK, GROUP
    rd = d["reads"]
    Ltot = d["L"] * d["n_scaffolds"]
    dev = rd["words"].device
    p = _cabi.ptr
    batch = _cabi.IsbReadsBatch(int(rd["n_segs"]), p(rd["seg_start"]), p(rd["seg_len"]), p(rd["seg_pair"]), p(rd["seg_word"]),
                                int(rd["n_words"]), p(rd["words"]), int(rd["max_seg_len"]), 0, 0, None, None, 0, None, 0,
                                Ltot, None, 0, None, 1, 0)
    n_groups = (Ltot + GROUP - 1) // GROUP
    grp_off = torch.empty(n_groups + 1, dtype=torch.int64, device=dev)
    n = C.c_int64(0)
    rc = eng.lib.isb_cols_from_reads(eng.ctx, C.byref(batch), p(grp_off), C.byref(n), None, None, 0)
    if rc != 0:
        raise RuntimeError("isb_cols_from_reads: " + eng.lib.isb_last_error(eng.ctx).decode())
    words = torch.empty(max(n.value, 1) * CHUNK, dtype=torch.int32, device=dev)[:n.value * CHUNK]
    ids = torch.empty(max(n.value, 1) * CHUNK, dtype=torch.int32, device=dev)[:n.value * CHUNK]
    rc = eng.lib.isb_cols_from_reads(eng.ctx, C.byref(batch), p(grp_off), C.byref(n), p(words), p(ids), n.value)
    if rc != 0:
        raise RuntimeError("isb_cols_from_reads: " + eng.lib.isb_last_error(eng.ctx).decode())
    torch.cuda.synchronize()
    return dict(n_groups=n_groups, grp_off=grp_off, n_chunks=int(n.value), words=words, ids=ids,
                nev_pos=rd["nev_pos"], nev_pair=rd["nev_pair"])
