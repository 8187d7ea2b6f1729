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
