"""Synthetic metagenome written as a real, coordinate-sorted, indexed BAM (bench / test support; numpy + zlib, no GPU).

The from-BAM leg of bench.py and the end-to-end tests need an input FILE of BASELINE.json's synthetic shape: pysam /
samtools are not available, so this module writes BGZF / BAM / BAI itself (SAM/BAM specification sections 4.1, 4.2, 5.2).
Not part of the hot path; the product reads BAMs (instrain_b200/csrc/isb_host.cpp), it never writes them.

Model = SURVEY.md section 8(d) (the one of the device generator and of the parity tests' numpy generator): iid uniform
reference, 4 haplotypes with abundances (0.4, 0.3, 0.2, 0.1), Bernoulli SNV sites whose alternative base is carried by a
random non-empty proper subset of haplotypes, 2 x 150 bp pairs (CIGAR 150M, flags 99 / 147), fragment length N(350, 30)
clipped to [200, 500], base qualities from the bundled BAM's empirical bins, substitution errors with p = 10^(-q/10), NM
tag = mismatches of the read against the reference.  EVERY pair is written (the read filter decides which count).
"""
import os
import struct
import zlib

import numpy as np

QUAL_BINS = np.array([8, 12, 22, 27, 32, 37, 41], dtype=np.uint8)
QUAL_P = np.array([0.002, 0.045, 0.034, 0.047, 0.085, 0.164, 0.623])
HAP_ABUND = np.array([0.4, 0.3, 0.2, 0.1])
READLEN = 150
NAME_LEN = 12                                  # "p" + 10 digits + NUL
REC = 4 + 32 + NAME_LEN + 4 + (READLEN + 1) // 2 + READLEN + 4     # bytes per alignment record (block_size field included)
_NT16 = np.array([1, 2, 8, 4], dtype=np.uint8)                      # A, C, T, G (inStrain's base order) -> BAM 4-bit codes
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class _BgzfWriter:
    """BGZF: gzip members of <= 64 KiB of payload each, 'BC' extra field with the member size; virtual offsets for the index."""

    def __init__(self, path, level=1):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.level = level
        self.block_off = 0                     # file offset of the member the buffer will become

    def tell(self):
        return (self.block_off << 16) | len(self.buf)

    def _flush(self, n):
        data = bytes(self.buf[:n])
        del self.buf[:n]
        co = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        bsize = len(comp) + 25
        self.f.write(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, bsize))
        self.f.write(comp)
        self.f.write(struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))
        self.block_off += bsize + 1

    def write(self, data, boundaries=None):
        """Append bytes; returns the virtual offset of every position listed in `boundaries` (offsets into `data`)."""
        out = []
        pos, nb = 0, 0
        boundaries = [] if boundaries is None else list(boundaries)
        mv = memoryview(data)
        while pos < len(data):
            room = 0xff00 - len(self.buf)
            take = min(room, len(data) - pos)
            while nb < len(boundaries) and boundaries[nb] < pos + take:
                out.append((self.block_off << 16) | (len(self.buf) + boundaries[nb] - pos))
                nb += 1
            self.buf += mv[pos:pos + take]
            pos += take
            if len(self.buf) >= 0xff00:
                self._flush(len(self.buf))
        while nb < len(boundaries):            # a boundary at the very end of the data
            out.append(self.tell())
            nb += 1
        return out

    def close(self):
        if self.buf:
            self._flush(len(self.buf))
        self.f.write(_BGZF_EOF)
        self.f.close()


def _scaffold_reads(rng, L, coverage, snv_density):
    """One scaffold: reference codes + per-read arrays in file order (sorted by position, stable)."""
    ref = rng.integers(0, 4, L, dtype=np.uint8)
    is_snv = rng.random(L) < snv_density
    alt = ((ref + 1 + rng.integers(0, 3, L)) % 4).astype(np.uint8)
    carriers = rng.integers(1, 15, L)
    hap = np.empty((4, L), dtype=np.uint8)
    for h in range(4):
        hap[h] = np.where(is_snv & ((carriers >> h) & 1).astype(bool), alt, ref)
    n_pairs = int(coverage * L / (2 * READLEN))
    frag = np.minimum(np.clip(np.rint(rng.normal(350, 30, n_pairs)), 200, 500).astype(np.int64), L)
    start = (rng.random(n_pairs) * (L - frag + 1)).astype(np.int64)
    hp = rng.choice(4, n_pairs, p=HAP_ABUND)
    offs = np.arange(READLEN, dtype=np.int64)
    pos = np.stack([start[:, None] + offs[None, :], (start + frag - READLEN)[:, None] + offs[None, :]], 1)
    true = hap[hp[:, None, None], pos]
    q = rng.choice(QUAL_BINS, size=pos.shape, p=QUAL_P / QUAL_P.sum())
    err = rng.random(pos.shape) < 10.0 ** (-q.astype(np.float64) / 10.0)
    sub = ((true + 1 + rng.integers(0, 3, pos.shape)) % 4).astype(np.uint8)
    base = np.where(err, sub, true).astype(np.uint8)
    nm = (base != ref[pos]).sum(2)                                  # [pair, mate]
    r_pos = np.concatenate([start, start + frag - READLEN])
    r_mate = np.concatenate([np.zeros(n_pairs, np.int64), np.ones(n_pairs, np.int64)])
    r_pair = np.concatenate([np.arange(n_pairs), np.arange(n_pairs)])
    order = np.argsort(r_pos, kind="stable")
    r_pos, r_mate, r_pair = r_pos[order], r_mate[order], r_pair[order]
    mpos = np.where(r_mate == 0, start[r_pair] + frag[r_pair] - READLEN, start[r_pair])
    isize = np.where(r_mate == 0, frag[r_pair], -frag[r_pair])
    return ref, dict(pos=r_pos, mate=r_mate, pair=r_pair, mpos=mpos, isize=isize, base=base[r_pair, r_mate],
                     qual=q[r_pair, r_mate], nm=nm[r_pair, r_mate], n_pairs=n_pairs)


def _records(tid, rd, name_off):
    """The scaffold's alignment records as one uint8 array [n, REC] (every record has the same size)."""
    n = len(rd["pos"])
    out = np.zeros((n, REC), dtype=np.uint8)
    core = np.zeros((n, 9), dtype="<i4")
    core[:, 0] = REC - 4                                            # block_size
    core[:, 1] = tid
    core[:, 2] = rd["pos"]
    core[:, 3] = NAME_LEN | (42 << 8) | (4681 << 16)                # l_read_name, mapq, bin (not used by this package's reader)
    flag = np.where(rd["mate"] == 0, 99, 147)
    core[:, 4] = 1 | (flag << 16)                                   # n_cigar_op, flag
    core[:, 5] = READLEN
    core[:, 6] = tid
    core[:, 7] = rd["mpos"]
    core[:, 8] = rd["isize"]
    out[:, :36] = core.view(np.uint8).reshape(n, 36)
    ids = (rd["pair"] + name_off).astype(np.int64)
    out[:, 36] = ord("p")
    for d in range(10):
        out[:, 37 + d] = 48 + (ids // 10 ** (9 - d)) % 10
    o = 36 + NAME_LEN
    out[:, o:o + 4] = np.frombuffer(struct.pack("<I", READLEN << 4), dtype=np.uint8)   # 150M
    o += 4
    codes = _NT16[rd["base"]]
    out[:, o:o + READLEN // 2] = (codes[:, 0::2] << 4) | codes[:, 1::2]
    o += (READLEN + 1) // 2
    out[:, o:o + READLEN] = rd["qual"]
    o += READLEN
    out[:, o], out[:, o + 1], out[:, o + 2] = ord("N"), ord("M"), ord("C")
    out[:, o + 3] = np.minimum(rd["nm"], 255)
    return out


def write_bam(path, L, n_scaffolds, coverage, snv_density, seed, level=1, name_prefix="synth_scaffold_"):
    """Write <path> (+ <path>.bai).  Returns dict(seqs={name: sequence}, n_reads, n_pairs, aligned_bases, bytes)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    names = ["%s%d" % (name_prefix, i) for i in range(n_scaffolds)]
    w = _BgzfWriter(path, level)
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (nm, L) for nm in names)
    hdr = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", n_scaffolds))
    for nm in names:
        hdr += struct.pack("<i", len(nm) + 1) + nm.encode() + b"\0" + struct.pack("<i", L)
    w.write(bytes(hdr))
    seqs, spans, linear, n_reads, name_off = {}, [], [], 0, 0
    letters = np.frombuffer(b"ACTG", dtype=np.uint8)
    for tid, nm in enumerate(names):
        ref, rd = _scaffold_reads(rng, L, coverage, snv_density)
        seqs[nm] = letters[ref].tobytes().decode()
        rec = _records(tid, rd, name_off)
        beg = w.tell()
        # linear index: for every 16 kb window the first record that overlaps it (records are sorted, all READLEN long)
        n_win = (L + 16383) >> 14
        first_rec = np.searchsorted(rd["pos"] + READLEN, np.arange(n_win) * 16384, side="right")
        has = first_rec < len(rec)
        voffs = w.write(rec.reshape(-1), boundaries=(first_rec[has] * REC).tolist())
        lin = np.zeros(n_win, dtype="<u8")
        lin[has] = voffs
        linear.append(lin)
        spans.append((beg, w.tell()))
        n_reads += len(rec)
        name_off += rd["n_pairs"]
    w.close()
    with open(path + ".bai", "wb") as f:                            # one bin with one chunk per reference
        f.write(b"BAI\1" + struct.pack("<i", n_scaffolds))
        for (beg, end), lin in zip(spans, linear):                   # + the linear index (16 kb windows) for region fetches
            f.write(struct.pack("<i", 1) + struct.pack("<Ii", 0, 1) + struct.pack("<QQ", beg, end) + struct.pack("<i", len(lin)) + lin.tobytes())
    return dict(seqs=seqs, names=names, n_reads=n_reads, n_pairs=name_off, aligned_bases=n_reads * READLEN,
                bytes=os.path.getsize(path))
