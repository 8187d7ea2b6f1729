"""Host packer: BAM -> read-major aligned segments or position-major event columns (ctypes face of csrc/isb_host.cpp).

Replaces the reference's pysam.AlignmentFile + samfile.pileup(...) (inStrain/profile/profile_utilities.py:56,150-153)
for the hot path: decodes the BAM once, applies htslib's mate-overlap quality tweak, expands CIGARs and emits the
columnar (ref_pos, base, qual, read_id) arrays + pair_mm that the kernels consume.  Host-only C++; no GPU needed.
"""
import ctypes as C

import numpy as np

from . import _cabi


def _lib():
    L = _cabi.load()
    if not getattr(L, "_packer_ready", False):
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.isb_bam_open.restype = vp
        L.isb_bam_open.argtypes = [C.c_char_p]
        L.isb_bam_close.argtypes = [vp]
        L.isb_bam_n_refs.restype = C.c_int
        L.isb_bam_n_refs.argtypes = [vp]
        L.isb_bam_ref_name.restype = C.c_char_p
        L.isb_bam_ref_name.argtypes = [vp, C.c_int]
        L.isb_bam_ref_len.restype = i64
        L.isb_bam_ref_len.argtypes = [vp, C.c_int]
        L.isb_bam_error.restype = C.c_char_p
        L.isb_bam_error.argtypes = [vp]
        L.isb_bam_peek_tid.restype = C.c_int
        L.isb_bam_peek_tid.argtypes = [vp]
        L.isb_bam_seek.restype = C.c_int
        L.isb_bam_seek.argtypes = [vp, C.c_uint64]
        L.isb_pack_scaffold.restype = vp
        L.isb_pack_scaffold.argtypes = [vp, C.c_int, i64, C.c_char_p, vp, vp, i32, i32]
        for f in ("isb_events_count", "isb_events_pairs", "isb_events_reads_seen", "isb_events_reads_packed"):
            getattr(L, f).restype = i64
            getattr(L, f).argtypes = [vp]
        L.isb_events_copy.argtypes = [vp] * 6
        L.isb_events_free.argtypes = [vp]
        L.isb_pack_scaffold_reads.restype = vp
        L.isb_pack_scaffold_reads.argtypes = [vp, C.c_int, i64, C.c_char_p, vp, vp, i32, i32, C.c_int]
        L.isb_pack_scaffold_reads_region.restype = vp
        L.isb_pack_scaffold_reads_region.argtypes = [vp, C.c_int, i64, C.c_char_p, vp, vp, i32, i32, C.c_int, i64, i64]
        for f in ("isb_reads_segs", "isb_reads_stream_words", "isb_reads_pairs", "isb_reads_n_events", "isb_reads_nev",
                  "isb_reads_reads_seen", "isb_reads_reads_packed"):
            getattr(L, f).restype = i64
            getattr(L, f).argtypes = [vp]
        L.isb_reads_max_len.restype = C.c_int
        L.isb_reads_max_len.argtypes = [vp]
        L.isb_reads_copy.argtypes = [vp] * 9 + [i64]
        L.isb_reads_free.argtypes = [vp]
        L._packer_ready = True
    return L


class BamPacker:
    """Streams a coordinate-sorted BAM scaffold by scaffold."""

    def __init__(self, path):
        self.lib = _lib()
        self.h = self.lib.isb_bam_open(path.encode())
        if not self.h:
            raise IOError("cannot open BAM %s" % path)
        n = self.lib.isb_bam_n_refs(self.h)
        self.ref_names = [self.lib.isb_bam_ref_name(self.h, i).decode() for i in range(n)]
        self.ref_lens = [int(self.lib.isb_bam_ref_len(self.h, i)) for i in range(n)]

    def close(self):
        if self.h:
            self.lib.isb_bam_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def seek(self, voffset):
        """Reposition at a BGZF virtual offset (read_bai): the next record read is the one stored there."""
        if self.lib.isb_bam_seek(self.h, int(voffset)) != 0:
            raise IOError("isb_bam_seek failed")

    def peek_tid(self):
        """tid of the next record: >= 0; -1 unmapped tail; -2 end of file.  A corrupt or truncated BAM (the C side's -3)
        raises IOError, as pysam / htslib do: a partial pass must never look like a complete one."""
        t = int(self.lib.isb_bam_peek_tid(self.h))
        if t == -3:
            raise IOError("error reading BAM: " + self.lib.isb_bam_error(self.h).decode())
        return t

    @staticmethod
    def _mm_levels(values):
        """R2M values -> uint8 levels.  The reference keys its tables by any mm (dicts); the device keeps M <= ISB_MAX_MM
        dense levels, so mismatch counts beyond the last level are folded into it (with a warning): such pairs -- summed
        NM >= 64 on a read pair -- keep being counted instead of aborting the run."""
        mm = np.asarray(values, dtype=np.int64)
        if len(mm) and mm.min() < 0:
            raise ValueError("R2M mismatch count < 0")
        top = _cabi.ISB_MAX_MM - 1
        n_over = int((mm > top).sum()) if len(mm) else 0
        if n_over:
            import logging
            logging.warning("instrain_b200: %d read pairs with more than %d mismatches are profiled at mm level %d",
                            n_over, top, top)
            mm = np.minimum(mm, top)
        return mm.astype(np.uint8)

    @staticmethod
    def _names(r2m):
        names = list(r2m.keys()) if isinstance(r2m, dict) else list(r2m)
        enc = [s.encode() for s in names]
        off = np.zeros(len(enc) + 1, dtype=np.int64)
        if enc:
            off[1:] = np.cumsum([len(b) for b in enc])
        if isinstance(r2m, dict):
            mm = BamPacker._mm_levels(np.fromiter((r2m[k] for k in names), dtype=np.int64, count=len(names)))
        else:
            mm = np.zeros(len(names), dtype=np.uint8)
        return len(names), b"".join(enc), off, mm

    def pack_scaffold_reads(self, tid, r2m, pos_offset=0, pair_id_offset=0, min_qual=30, region=None):
        """Consume the records of scaffold `tid` as READ-MAJOR aligned segments (instrain_b200/reads.py layout, the
        scaffold's own word stream: `stream` = [data words + one zero word] per segment, seg_word relative to it).
        region = (lo, hi): only the reads that overlap the scaffold positions [lo, hi) (an index fetch of that region);
        reading stops behind it, the reader is left inside the scaffold."""
        n_names, blob, off, mm = self._names(r2m)
        lo, hi = (int(region[0]), int(region[1])) if region is not None else (0, -1)
        r = self.lib.isb_pack_scaffold_reads_region(self.h, tid, n_names, blob, off.ctypes.data, mm.ctypes.data, pos_offset,
                                                    pair_id_offset, min_qual, lo, hi)
        if not r:
            raise IOError("isb_pack_scaffold_reads failed: " + self.lib.isb_bam_error(self.h).decode())
        try:
            n, nw = int(self.lib.isb_reads_segs(r)), int(self.lib.isb_reads_stream_words(r))
            npairs, nev = int(self.lib.isb_reads_pairs(r)), int(self.lib.isb_reads_nev(r))
            out = dict(seg_start=np.empty(n, np.int32), seg_len=np.empty(n, np.uint16), seg_pair=np.empty(n, np.int32),
                       seg_word=np.empty(n, np.int64), stream=np.empty(nw, np.uint32), nev_pos=np.empty(nev, np.int32),
                       nev_pair=np.empty(nev, np.int32), pair_mm=np.empty(npairs, np.uint8))
            self.lib.isb_reads_copy(r, *[out[k].ctypes.data for k in ("seg_start", "seg_len", "seg_pair", "seg_word", "stream",
                                                                      "nev_pos", "nev_pair", "pair_mm")], 0)
            out["max_seg_len"] = int(self.lib.isb_reads_max_len(r))
            out["n_events"] = int(self.lib.isb_reads_n_events(r))
            out["reads_seen"] = int(self.lib.isb_reads_reads_seen(r))
            out["reads_packed"] = int(self.lib.isb_reads_reads_packed(r))
        finally:
            self.lib.isb_reads_free(r)
        return out

    def pack_scaffold(self, tid, r2m, pos_offset=0, pair_id_offset=0):
        """Consume the records of scaffold `tid`; r2m is the reference's sR2M[scaffold]: dict name -> mm, or a set."""
        names = list(r2m.keys()) if isinstance(r2m, dict) else list(r2m)
        enc = [s.encode() for s in names]
        off = np.zeros(len(enc) + 1, dtype=np.int64)
        if enc:
            off[1:] = np.cumsum([len(b) for b in enc])
        blob = b"".join(enc)
        if isinstance(r2m, dict):
            mm = BamPacker._mm_levels(np.fromiter((r2m[k] for k in names), dtype=np.int64, count=len(names)))
        else:
            mm = np.zeros(len(names), dtype=np.uint8)
        e = self.lib.isb_pack_scaffold(self.h, tid, len(names), blob, off.ctypes.data, mm.ctypes.data, pos_offset,
                                       pair_id_offset)
        if not e:
            raise IOError("isb_pack_scaffold failed: " + self.lib.isb_bam_error(self.h).decode())
        try:
            n, npairs = int(self.lib.isb_events_count(e)), int(self.lib.isb_events_pairs(e))
            out = dict(ref_pos=np.empty(n, np.int32), base=np.empty(n, np.uint8), qual=np.empty(n, np.uint8),
                       read_id=np.empty(n, np.int32), pair_mm=np.empty(npairs, np.uint8))
            self.lib.isb_events_copy(e, out["ref_pos"].ctypes.data, out["base"].ctypes.data, out["qual"].ctypes.data,
                                     out["read_id"].ctypes.data, out["pair_mm"].ctypes.data)
            out["reads_seen"] = int(self.lib.isb_events_reads_seen(e))
            out["reads_packed"] = int(self.lib.isb_events_reads_packed(e))
        finally:
            self.lib.isb_events_free(e)
        return out


def read_bai(path):
    """First-alignment virtual offset of every reference of a .bai index (SAM spec 5.2): the smallest chunk begin over the
    reference's bins (the metadata pseudo-bin 37450 excluded); None for references without alignments."""
    import struct
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"BAI\1":
        raise IOError("not a BAI index: %s" % path)
    n_ref, = struct.unpack_from("<i", data, 4)
    o, out = 8, []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", data, o)
        o += 4
        first = None
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", data, o)
            o += 8
            if b != 37450 and n_chunk > 0:
                begs = np.frombuffer(data, dtype="<u8", count=2 * n_chunk, offset=o)[0::2]
                m = int(begs.min())
                first = m if first is None else min(first, m)
            o += 16 * n_chunk
        n_intv, = struct.unpack_from("<i", data, o)
        o += 4 + 8 * n_intv
        out.append(first)
    return out


def read_bai_linear(path):
    """The linear index of a .bai (SAM spec 5.2): per reference the virtual offset of the first alignment that overlaps each
    16 kb window (0 = no entry).  seek_offset(ioffset, first, lo) below picks where a region fetch starts reading."""
    import struct
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"BAI\1":
        raise IOError("not a BAI index: %s" % path)
    n_ref, = struct.unpack_from("<i", data, 4)
    o, out = 8, []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", data, o)
        o += 4
        for _ in range(n_bin):
            _, n_chunk = struct.unpack_from("<Ii", data, o)
            o += 8 + 16 * n_chunk
        n_intv, = struct.unpack_from("<i", data, o)
        out.append(np.frombuffer(data, dtype="<u8", count=n_intv, offset=o + 4).copy())
        o += 4 + 8 * n_intv
    return out


def seek_offset(ioffset, first, lo):
    """Virtual offset a fetch of positions >= lo starts at: the linear-index entry of lo's 16 kb window (every read that
    overlaps lo overlaps that window), falling back to earlier windows and to the reference's first alignment."""
    k = min(int(lo) >> 14, len(ioffset) - 1)
    while k >= 0:
        if ioffset[k]:
            return int(ioffset[k])
        k -= 1
    return first


def scan_scaffold_offsets(bam):
    """The same list as read_bai, computed by walking the BAM itself (pure Python: every BGZF block is inflated once) --
    for BAMs that come without an index.  inStrain requires indexed BAMs, so this is a fallback for small inputs."""
    import struct
    import zlib
    with open(bam, "rb") as f:
        raw = f.read()
    blocks, o = [], 0                                      # (compressed offset, inflated bytes)
    while o + 18 <= len(raw):
        xlen = raw[o + 10] | (raw[o + 11] << 8)
        extra, bsize, i = raw[o + 12:o + 12 + xlen], None, 0
        while i + 4 <= xlen:
            slen = extra[i + 2] | (extra[i + 3] << 8)
            if extra[i:i + 2] == b"BC" and slen == 2:
                bsize = extra[i + 4] | (extra[i + 5] << 8)
            i += 4 + slen
        if bsize is None:
            raise IOError("BGZF block without BC field")
        blocks.append((o, zlib.decompress(raw[o + 12 + xlen:o + bsize + 1 - 8], -15)))
        o += bsize + 1
    data = b"".join(b for _, b in blocks)
    starts = np.cumsum([0] + [len(b) for _, b in blocks])  # inflated offset of every block

    def voffset(u):
        k = int(np.searchsorted(starts, u, side="right")) - 1
        while k + 1 < len(blocks) and len(blocks[k][1]) == 0:
            k += 1
        return (blocks[k][0] << 16) | (u - int(starts[k]))

    l_text, = struct.unpack_from("<i", data, 4)
    u = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, u)
    u += 4
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, u)
        u += 8 + l_name
    first = [None] * n_ref
    while u + 4 <= len(data):
        block_size, = struct.unpack_from("<i", data, u)
        tid, = struct.unpack_from("<i", data, u + 4)
        if 0 <= tid < n_ref and first[tid] is None:
            first[tid] = voffset(u)
        u += 4 + block_size
    return first


def find_bai(bam):
    for cand in (bam + ".bai", bam[:-4] + ".bai" if bam.endswith(".bam") else None):
        if cand and __import__("os").path.exists(cand):
            return cand
    return None


def pack_scaffolds_parallel(bam, jobs, threads, min_qual=30, window=None):
    """Pack several scaffolds of one indexed BAM concurrently: `jobs` = [(tid, r2m, pos_offset), ...]; yields the
    pack_scaffold_reads results in job order.  Every host thread owns a BamPacker and seeks to its scaffold through the
    .bai index; the C++ packer releases the GIL (ctypes), so plain threads scale.  Pair ids of every result start at 0
    (pair_id_offset is applied by the consumer: it depends on the scaffolds before it)."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    bai = find_bai(bam)
    first = read_bai(bai) if bai is not None else scan_scaffold_offsets(bam)
    tls = threading.local()
    opened, lock = [], threading.Lock()
    jobs = list(jobs)
    # runs of consecutive jobs: one seek per run, then the reader simply continues into the next scaffold when it is the
    # one wanted (small scaffolds share BGZF blocks: seeking to each would inflate every block several times)
    run_len = max(1, len(jobs) // (threads * 8))
    runs = [jobs[i:i + run_len] for i in range(0, len(jobs), run_len)]

    def work(run):
        bp = getattr(tls, "bp", None)
        if bp is None:
            bp = tls.bp = BamPacker(bam)
            with lock:
                opened.append(bp)
        out, at = [], None                                   # at = tid the reader is positioned on, when known
        for tid, r2m, pos_offset in run:
            if first[tid] is None:
                out.append(None)
                continue
            if at != tid:
                bp.seek(first[tid])
                if bp.peek_tid() != tid:
                    raise IOError("index does not lead to scaffold %d" % tid)
            out.append(bp.pack_scaffold_reads(tid, r2m, pos_offset=pos_offset, pair_id_offset=0, min_qual=min_qual))
            nxt = bp.peek_tid()
            at = nxt if nxt >= 0 else None
        return out

    window = window or 2 * threads
    try:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            pending = []
            for run in runs:
                pending.append(ex.submit(work, run))
                if len(pending) >= window:
                    yield from pending.pop(0).result()
            for fut in pending:
                yield from fut.result()
    finally:
        for bp in opened:
            bp.close()
