"""Pure-Python BAM reader (oracle / test infrastructure only).

pysam/htslib are not available in the build container, so the oracle decodes BAM itself:
BGZF is a series of gzip members (stdlib `gzip` reads concatenated members), the record
layout follows the SAM/BAM specification section 4.2.

Replaces, for the oracle only, the reference's `pysam.AlignmentFile(bam)` at
inStrain/profile/profile_utilities.py:56 and inStrain/filter_reads.py:137.
"""
import gzip
import struct
from collections import namedtuple

import numpy as np

SEQ_NT16 = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"

Read = namedtuple(
    "Read",
    "name flag tid pos mapq cigar seq qual mtid mpos isize nm",
)
# cigar: list of (op:int, length:int) with op index into CIGAR_OPS
# seq:   python str (upper-case nt16 letters);  qual: np.uint8 array (phred)


def _parse_tags_nm(buf, off, end):
    """Return the NM tag value (int) or None."""
    while off < end:
        tag = buf[off:off + 2]
        typ = chr(buf[off + 2])
        off += 3
        if typ == "A":
            size, val = 1, None
        elif typ in "cC":
            size = 1
            val = struct.unpack_from("<b" if typ == "c" else "<B", buf, off)[0]
        elif typ in "sS":
            size = 2
            val = struct.unpack_from("<h" if typ == "s" else "<H", buf, off)[0]
        elif typ in "iI":
            size = 4
            val = struct.unpack_from("<i" if typ == "i" else "<I", buf, off)[0]
        elif typ == "f":
            size, val = 4, None
        elif typ in "ZH":
            e = buf.index(b"\0", off)
            size, val = e - off + 1, None
        elif typ == "B":
            sub = chr(buf[off])
            n = struct.unpack_from("<i", buf, off + 1)[0]
            size = 5 + n * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            val = None
        else:
            raise ValueError("bad BAM tag type %r" % typ)
        if tag == b"NM":
            return val
        off += size
    return None


def read_bam(path):
    """Return (refs, reads): refs = [(name, length)], reads = list[Read] in file order."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file: %s" % path)
    off = 4
    (l_text,) = struct.unpack_from("<i", data, off)
    off += 4 + l_text
    (n_ref,) = struct.unpack_from("<i", data, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, off)
        off += 4
        name = data[off:off + l_name - 1].decode()
        off += l_name
        (l_ref,) = struct.unpack_from("<i", data, off)
        off += 4
        refs.append((name, l_ref))
    reads = []
    n = len(data)
    while off < n:
        (block_size,) = struct.unpack_from("<i", data, off)
        off += 4
        end = off + block_size
        (tid, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, mtid, mpos, isize) = struct.unpack_from(
            "<iiBBHHHiiii", data, off)
        p = off + 32
        name = data[p:p + l_read_name - 1].decode()
        p += l_read_name
        cig = struct.unpack_from("<%dI" % n_cigar, data, p)
        cigar = [(c & 0xF, c >> 4) for c in cig]
        p += 4 * n_cigar
        nb = (l_seq + 1) // 2
        packed = np.frombuffer(data, dtype=np.uint8, count=nb, offset=p)
        codes = np.empty(nb * 2, dtype=np.uint8)
        codes[0::2] = packed >> 4
        codes[1::2] = packed & 0xF
        seq = "".join(SEQ_NT16[c] for c in codes[:l_seq])
        p += nb
        qual = np.frombuffer(data, dtype=np.uint8, count=l_seq, offset=p).copy()
        p += l_seq
        nm = _parse_tags_nm(data, p, end)
        reads.append(Read(name, flag, tid, pos, mapq, cigar, seq, qual, mtid, mpos, isize, nm))
        off = end
    return refs, reads


def read_fasta(path):
    """scaffold -> upper-cased sequence (mirrors inStrain/profile/fasta.py:25-27)."""
    seqs, name, chunks = {}, None, []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks).upper()
                name, chunks = line[1:].split()[0], []
            elif line:
                chunks.append(line)
    if name is not None:
        seqs[name] = "".join(chunks).upper()
    return seqs


def iterate_splits(s_len, window_len):
    """Split geometry, 0-based double-inclusive (restates inStrain/profile/fasta.py:56-73)."""
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


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path, refs, reads):
    """Minimal BGZF/BAM writer for test fixtures: refs = [(name, length)], reads = iterable of Read (file order kept)."""
    import zlib
    out = bytearray()
    out += b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", len(refs))
    for name, length in refs:
        nb = name.encode() + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", length)
    code = {c: i for i, c in enumerate(SEQ_NT16)}
    for r in reads:
        nb = r.name.encode() + b"\0"
        rl = sum(n for op, n in r.cigar if op in (0, 2, 3, 7, 8)) or 1
        cig = b"".join(struct.pack("<I", (n << 4) | op) for op, n in r.cigar)
        l_seq = len(r.seq)
        nib = [code[c] for c in r.seq] + [0]
        seq = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, l_seq, 2))
        tags = b"" if r.nm is None else b"NMC" + struct.pack("<B", min(int(r.nm), 255))
        body = struct.pack("<iiBBHHHiiii", r.tid, r.pos, len(nb), r.mapq, _reg2bin(r.pos, r.pos + rl), len(r.cigar),
                           r.flag, l_seq, r.mtid, r.mpos, r.isize) + nb + cig + seq + bytes(r.qual) + tags
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        data = bytes(out)
        for i in range(0, max(len(data), 1), 0xff00):
            chunk = data[i:i + 0xff00]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp
                     + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
