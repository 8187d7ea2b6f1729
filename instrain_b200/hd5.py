"""Native writer / reader for the `special` HDF5 stores of an SNVprofile (covT.hd5, clonT.hd5) -- SURVEY.md 8(f4).

The reference writes these with h5py (inStrain/SNVprofile.py:717-733): one dataset per (scaffold, mm) named
"<scaffold>::<mm>", holding `np.array([series.values, series.index])` (a 2 x N array: int64 for coverage, float64 for
clonality) with gzip compression; and reads them back dataset by dataset (SNVprofile.py:750-786).  h5py / libhdf5 are
not part of this image, so the container format is produced here directly, in the same on-disk dialect libhdf5 emits
for h5py's default `libver='earliest'` -- the one the reference's own stored files use (checked structure by structure
against test/test_data/*.forRC.IS/raw_data/covT.hd5):

  superblock v0 (8-byte offsets/lengths, group leaf K = 4, internal K = 16)
  root group: v1 object header with a symbol-table message -> v1 B-tree (node type 0) -> SNOD leaves + local heap
  dataset:    v1 object header {dataspace v1 (rank 2, max dims), datatype v1 (int64 LE / IEEE float64 LE),
              fill value v2, filter pipeline v1 (deflate, level 4), layout v3 chunked -> v1 B-tree (node type 1)}
  chunks:     h5py's guess_chunk shapes (R x C), zlib streams, edge chunks padded with the fill value (0)

`read_hd5` parses exactly that dialect (it reads the reference's stored files, which is how tests pin the covT / clonT
arrays of the hot path at EVERY position, not only at SNV sites) and `write_hd5` emits it.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K, CHUNK_K = 4, 16, 32
_DT_INT64 = bytes([0x10, 0x08, 0x00, 0x00]) + struct.pack("<IHH", 8, 0, 64)
_DT_FLOAT64 = bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<IHHBBBBI", 8, 0, 64, 52, 11, 0, 52, 1023)


# ----------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
            raise ValueError("unsupported HDF5 dialect (superblock v%d, offsets %d, lengths %d)" % (buf[8], buf[13], buf[14]))
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, 16)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", buf, 24)
        # root symbol table entry
        _, self.root_header, cache_type = struct.unpack_from("<QQI", buf, 56)
        self.root_btree = self.root_heap = None
        if cache_type == 1:
            self.root_btree, self.root_heap = struct.unpack_from("<QQ", buf, 56 + 24)

    def messages(self, addr):
        """(type, flags, payload) of every message of a v1 object header, following continuation blocks."""
        b = self.b
        version, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise ValueError("object header v%d not supported" % version)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                payload = b[p + 8:p + 8 + msize]
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", payload, 0))
                out.append((mtype, flags, payload))
                p += 8 + msize
        return out

    def heap_name(self, heap_addr, off):
        b = self.b
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        e = b.index(b"\0", data + off)
        return bytes(b[data + off:e]).decode()

    def group_entries(self, btree, heap):
        """name -> object header address, walking the group's v1 B-tree."""
        b = self.b
        out = {}
        stack = [btree]
        while stack:
            a = stack.pop()
            if b[a:a + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
                if ntype != 0:
                    raise ValueError("group B-tree expected")
                p = a + 24
                for i in range(used):
                    stack.append(struct.unpack_from("<Q", b, p + 8 + i * 16)[0])
            elif b[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, a + 6)[0]
                for i in range(n):
                    noff, haddr = struct.unpack_from("<QQ", b, a + 8 + i * 40)
                    out[self.heap_name(heap, noff)] = haddr
            else:
                raise ValueError("bad group node at %d" % a)
        return out

    def chunks(self, btree, rank1):
        """[(offsets, nbytes, filter_mask, address)] of a chunked dataset's v1 B-tree (node type 1)."""
        b = self.b
        out = []
        if btree == UNDEF:
            return out
        stack = [btree]
        ksize = 8 + 8 * rank1
        while stack:
            a = stack.pop()
            if b[a:a + 4] != b"TREE":
                raise ValueError("bad chunk B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
            p = a + 24
            for i in range(used):
                q = p + i * (ksize + 8)
                nbytes, mask = struct.unpack_from("<II", b, q)
                offs = struct.unpack_from("<%dQ" % rank1, b, q + 8)
                child = struct.unpack_from("<Q", b, q + ksize)[0]
                if level > 0:
                    stack.append(child)
                else:
                    out.append((offs, nbytes, mask, child))
        return out

    def dataset(self, addr):
        dims = dtype = None
        layout = None
        deflate = False
        for mtype, _, m in self.messages(addr):
            if mtype == 1:
                if m[0] != 1:
                    raise ValueError("dataspace v%d not supported" % m[0])
                rank = m[1]
                dims = struct.unpack_from("<%dQ" % rank, m, 8)
            elif mtype == 3:
                cls, size = m[0] & 0x0F, struct.unpack_from("<I", m, 4)[0]
                if m[1] & 1:
                    raise ValueError("big-endian data not supported")
                if cls == 0:
                    dtype = np.dtype("<%s%d" % ("i" if m[1] & 8 else "u", size))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % size)
                else:
                    raise ValueError("datatype class %d not supported" % cls)
            elif mtype == 0x0B:
                nf = m[1]
                p = 8
                for _ in range(nf):
                    fid, nlen, _, ncv = struct.unpack_from("<HHHH", m, p)
                    if fid != 1:
                        raise ValueError("filter %d not supported" % fid)
                    deflate = True
                    p += 8 + ((nlen + 7) & ~7) + 4 * (ncv + (ncv & 1))
            elif mtype == 8:
                if m[0] != 3:
                    raise ValueError("layout v%d not supported" % m[0])
                if m[1] == 2:
                    r1 = m[2]
                    layout = ("chunked", struct.unpack_from("<Q", m, 3)[0], struct.unpack_from("<%dI" % r1, m, 11))
                elif m[1] == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", m, 2)
                else:
                    raise ValueError("compact layout not supported")
        if dims is None or dtype is None or layout is None:
            raise ValueError("incomplete dataset header")
        out = np.zeros(dims, dtype=dtype)
        if out.size == 0:
            return out
        if layout[0] == "contiguous":
            a, n = layout[1], layout[2]
            if a != UNDEF:
                out[...] = np.frombuffer(self.b, dtype=dtype, count=out.size, offset=a).reshape(dims)
            return out
        _, btree, cdims = layout
        cshape = cdims[:-1]
        for offs, nbytes, mask, a in self.chunks(btree, len(cdims)):
            raw = bytes(self.b[a:a + nbytes])
            if deflate and not (mask & 1):
                raw = zlib.decompress(raw)
            blk = np.frombuffer(raw, dtype=dtype).reshape(cshape)
            sl_out = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cshape, dims))
            sl_blk = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = blk[sl_blk]
        return out


def read_hd5(path, names=None):
    """{dataset name -> ndarray} of every dataset in the root group (or only `names`)."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    if r.root_btree is None:
        for mtype, _, m in r.messages(r.root_header):
            if mtype == 0x11:
                r.root_btree, r.root_heap = struct.unpack_from("<QQ", m, 0)
    entries = r.group_entries(r.root_btree, r.root_heap)
    return {n: r.dataset(a) for n, a in entries.items() if names is None or n in names}


def describe_hd5(path, max_datasets=None):
    """Address-free structural description of a file in this dialect: everything libhdf5 looks at before it touches data
    -- superblock fields, the root group's header messages / B-tree / symbol nodes / local heap, and per dataset the
    header messages with their version bytes and decoded fields (dataspace, datatype, fill value, filter pipeline, layout,
    chunk B-tree node type / depth / key layout).  Two files that agree here are read by the same libhdf5 code paths:
    tests/test_hd5.py compares a file written by write_hd5 with the reference's own stored covT.hd5 / clonT.hd5."""
    with open(path, "rb") as f:
        b = f.read()
    r = _Reader(b)
    sb = dict(signature=bytes(b[:8]), superblock_version=b[8], free_space_version=b[9], root_group_version=b[10],
              shared_header_version=b[12], size_of_offsets=b[13], size_of_lengths=b[14], leaf_k=r.leaf_k, internal_k=r.internal_k,
              file_consistency_flags=struct.unpack_from("<I", b, 20)[0], base_address=r.base,
              free_space_address_undefined=struct.unpack_from("<Q", b, 32)[0] == UNDEF,
              driver_info_address_undefined=struct.unpack_from("<Q", b, 48)[0] == UNDEF,
              eof_is_file_size=r.eof == len(b), root_cache_type=struct.unpack_from("<I", b, 72)[0])
    root_msgs = [(t, fl, len(m)) for t, fl, m in r.messages(r.root_header)]
    if r.root_btree is None:
        for mtype, _, m in r.messages(r.root_header):
            if mtype == 0x11:
                r.root_btree, r.root_heap = struct.unpack_from("<QQ", m, 0)
    hv, = struct.unpack_from("<B", b, r.root_heap + 4)
    heap = dict(signature=bytes(b[r.root_heap:r.root_heap + 4]), version=hv)
    # group B-tree: depth, node signatures / types, entries per symbol node
    depth, a, snod_versions, snod_fill = 0, r.root_btree, set(), []
    stack = [(r.root_btree, 0)]
    while stack:
        a, d = stack.pop()
        if b[a:a + 4] == b"TREE":
            ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
            assert ntype == 0
            depth = max(depth, level + 1)
            for i in range(used):
                stack.append((struct.unpack_from("<Q", b, a + 24 + 8 + i * 16)[0], d + 1))
        else:
            assert b[a:a + 4] == b"SNOD"
            snod_versions.add(b[a + 4])
            snod_fill.append(struct.unpack_from("<H", b, a + 6)[0])
    group = dict(btree_depth=depth, snod_versions=sorted(snod_versions), max_entries_per_snod=max(snod_fill) if snod_fill else 0,
                 snod_entries_bound=2 * r.leaf_k)
    entries = r.group_entries(r.root_btree, r.root_heap)
    dsets = {}
    for name in sorted(entries)[:max_datasets]:
        addr = entries[name]
        version, _, nmsg, _, _ = struct.unpack_from("<BBHII", b, addr)
        msgs, chunk_tree = [], None
        for mtype, flags, m in r.messages(addr):
            if mtype == 0:                                                   # NIL padding: size is layout noise
                continue
            d = dict(type=mtype, flags=flags)
            if mtype == 1:
                rank = m[1]
                d.update(version=m[0], rank=rank, dim_flags=m[2], dims=struct.unpack_from("<%dQ" % rank, m, 8),
                         max_dims=struct.unpack_from("<%dQ" % rank, m, 8 + 8 * rank) if m[2] & 1 else None)
            elif mtype == 3:
                size = struct.unpack_from("<I", m, 4)[0]
                d.update(class_and_version=m[0], bits=bytes(m[1:4]), size=size, properties=bytes(m[8:8 + (12 if (m[0] & 15) == 1 else 4)]))
            elif mtype == 5:
                d.update(version=m[0], alloc_time=m[1], write_time=m[2], defined=m[3],
                         size=struct.unpack_from("<I", m, 4)[0] if m[0] >= 2 and m[3] else None)
            elif mtype == 0x0B:
                nf = m[1]
                fl, q = [], 8
                for _ in range(nf):
                    fid, nlen, fflags, ncv = struct.unpack_from("<HHHH", m, q)
                    nm = bytes(m[q + 8:q + 8 + nlen]).rstrip(b"\0")
                    q += 8 + ((nlen + 7) & ~7)
                    fl.append(dict(id=fid, name=nm, flags=fflags, client_data=struct.unpack_from("<%dI" % ncv, m, q)))
                    q += 4 * (ncv + (ncv & 1))
                d.update(version=m[0], filters=fl)
            elif mtype == 8:
                d.update(version=m[0], layout_class=m[1])
                if m[1] == 2:
                    r1 = m[2]
                    bt = struct.unpack_from("<Q", m, 3)[0]
                    d.update(rank_plus_1=r1, chunk_dims=struct.unpack_from("<%dI" % r1, m, 11), allocated=bt != UNDEF)
                    if bt != UNDEF:
                        ntype, level, used = struct.unpack_from("<BBH", b, bt + 4)
                        ch = r.chunks(bt, r1)
                        chunk_tree = dict(signature=bytes(b[bt:bt + 4]), node_type=ntype, root_level=level, n_chunks=len(ch),
                                          filter_masks=sorted({c[2] for c in ch}),
                                          chunk_offsets=sorted(c[0] for c in ch))
            else:
                d.update(size=len(m))
            msgs.append(d)
        dsets[name] = dict(header_version=version, messages=msgs, chunk_tree=chunk_tree)
    return dict(superblock=sb, root_messages=root_msgs, heap=heap, group=group, n_datasets=len(entries), datasets=dsets)


def load_special(path, scaffolds=()):
    """Mirror of SNVprofile._load_special for covT / clonT (SNVprofile.py:750-786): scaffold -> mm -> pandas Series
    (values indexed by position; the reference rebuilds `pd.Series(data=arr[0], index=np.array(arr[1].astype('int')))`)."""
    import pandas as pd
    out = {}
    want = set(scaffolds)
    for key, arr in read_hd5(path).items():
        scaff, mm = key.rsplit("::", 1)
        if want and scaff not in want:
            continue
        out.setdefault(scaff, {})[int(mm)] = pd.Series(data=arr[0], index=np.array(arr[1].astype("int")))
    return out


# ----------------------------------------------------------------------------------------------------------- writer
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, payload, flags=0):
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), flags) + payload


class _FileBuf:
    """The part of a bytearray the writer uses (len, +=, patching the first bytes), backed by a file: the covT / clonT
    stores of a 100 Mb profile are gigabytes of deflated chunks, which need not sit in memory before they are written."""

    def __init__(self, fh, n_zero):
        self.fh, self.n = fh, 0
        self += b"\0" * n_zero

    def __len__(self):
        return self.n

    def __iadd__(self, b):
        self.fh.write(b)
        self.n += len(b)
        return self

    def patch_head(self, b):
        self.fh.seek(0)
        self.fh.write(b)
        self.fh.seek(0, 2)


def _guess_chunk(rows, n, typesize):
    """h5py's chunk-shape heuristic (h5py/_hl/filters.py: guess_chunk) for a 2-D dataset without maxshape: halve the
    dimensions in turn until a chunk is within 50 % of a target size between 8 KiB and 1 MiB that grows with the
    dataset.  The reference stores covT / clonT through h5py's create_dataset(compression='gzip'): same chunk shapes."""
    import math
    base, cmin, cmax = 16 * 1024, 8 * 1024, 1024 * 1024
    chunks = [float(rows), float(n)]
    target = base * (2 ** math.log10(rows * n * typesize / (1024.0 * 1024.0)))
    target = cmax if target > cmax else cmin if target < cmin else target
    idx = 0
    while True:
        nbytes = chunks[0] * chunks[1] * typesize
        if (nbytes < target or abs(nbytes - target) / target < 0.5) and nbytes < cmax:
            break
        if chunks[0] * chunks[1] == 1:
            break
        chunks[idx % 2] = math.ceil(chunks[idx % 2] / 2.0)
        idx += 1
    return int(chunks[0]), int(chunks[1])


class _Writer:
    def __init__(self, fh=None):
        # superblock + root object header come first and are patched at the end
        self.buf = _FileBuf(fh, 96 + 40) if fh is not None else bytearray(96 + 40)
        self.entries = []                       # (name bytes, object header address)

    def _append(self, b):
        a = len(self.buf)
        self.buf += b
        return a

    @staticmethod
    def prepare(arr, level=4):
        """The CPU-heavy, state-free half of add_dataset: dtype widening, chunking and deflate of one dataset (zlib
        releases the GIL: write_hd5 runs this on a thread pool).  Returns (rows, n, C, datatype message, chunks)."""
        arr = np.ascontiguousarray(arr)
        if arr.ndim != 2:
            raise ValueError("2-D arrays only")
        if arr.dtype.kind in "iu" or arr.dtype.kind == "b":
            arr, dt = arr.astype("<i8", copy=False), _DT_INT64
        elif arr.dtype.kind == "f":
            arr, dt = arr.astype("<f8", copy=False), _DT_FLOAT64
        else:
            raise ValueError("dtype %s not supported" % arr.dtype)
        rows, n = arr.shape
        # Chunk shape: h5py's guess_chunk (what the reference's files were written with) as long as the chunk index stays
        # a single B-tree node (<= 2 * CHUNK_K chunks); larger datasets: 1 x C chunks, 2 * CHUNK_K of them.
        R, C = _guess_chunk(rows, n, 8) if n else (1, 1024)          # empty dataset: no chunk index (libhdf5 does the same)
        if n and (-(-rows // R)) * (-(-n // C)) > 2 * CHUNK_K:
            R = 1
            per_row = max(1, (2 * CHUNK_K) // max(rows, 1))
            C = max(1, -(-n // per_row))
            if rows * (-(-n // C)) > 2 * CHUNK_K:
                raise ValueError("too many rows for a single-node chunk index")
        chunks = []
        for r0 in range(0, rows if n else 0, R):
            for c0 in range(0, n, C):
                blk = np.zeros((R, C), dtype=arr.dtype)
                seg = arr[r0:r0 + R, c0:c0 + C]
                blk[:seg.shape[0], :seg.shape[1]] = seg
                chunks.append(((r0, c0, 0), zlib.compress(blk.tobytes(), level)))
        return rows, n, (R, C), dt, chunks

    def add_dataset(self, name, arr, level=4, prepared=None):
        rows, n, (R, C), dt, chunks = prepared if prepared is not None else self.prepare(arr, level)
        keys = [(offs, len(z), self._append(z)) for offs, z in chunks]
        btree = UNDEF
        if keys:
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF))
            # the rightmost key as libhdf5's insertion history leaves it: a chunk at or beyond the current right key moves
            # the key to (that chunk's offset + the chunk dimensions); a chunk below it splits a child and leaves it alone
            right = None
            for offs, nbytes, a in keys:
                node += struct.pack("<II3QQ", nbytes, 0, *offs, a)
                if right is None or offs >= right:
                    right = (offs[0] + R, offs[1] + C, 8)
            node += struct.pack("<II3Q", 0, 0, *right)
            node += b"\0" * (24 + (2 * CHUNK_K + 1) * 32 + 2 * CHUNK_K * 8 - len(node))
            self.buf += b"\0" * (-len(self.buf) % 8)
            btree = self._append(bytes(node))
        msgs = [
            _msg(1, struct.pack("<BBB5x", 1, 2, 1) + struct.pack("<4Q", rows, n, rows, n)),
            _msg(3, dt, flags=1),
            _msg(5, bytes([2, 3, 0, 1]) + struct.pack("<I", 0), flags=1),
            _msg(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", level, 0), flags=1),
            _msg(8, struct.pack("<BBBQ3I", 3, 2, 3, btree, R, C, 8)),
        ]
        body = b"".join(msgs)
        self.buf += b"\0" * (-len(self.buf) % 8)
        haddr = self._append(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)
        self.entries.append((name.encode(), haddr))

    def finish(self):
        ents = sorted(self.entries)              # bytewise order == strcmp order, what the group B-tree is keyed on
        for i in range(1, len(ents)):
            if ents[i][0] == ents[i - 1][0]:
                raise ValueError("duplicate dataset name %r" % ents[i][0])
        # local heap data segment: "" at offset 0, then the names, NUL-terminated and 8-byte aligned
        heap = bytearray(8)
        offs = []
        for name, _ in ents:
            offs.append(len(heap))
            heap += _pad8(name + b"\0")
        self.buf += b"\0" * (-len(self.buf) % 8)
        # symbol-table leaves (<= 2 * LEAF_K entries each)
        level = []                               # (address, heap offset of the largest name below)
        for i in range(0, len(ents), 2 * LEAF_K):
            grp = range(i, min(i + 2 * LEAF_K, len(ents)))
            node = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(grp)))
            for j in grp:
                node += struct.pack("<QQII16x", offs[j], ents[j][1], 0, 0)
            node += b"\0" * (8 + 2 * LEAF_K * 40 - len(node))
            level.append((self._append(bytes(node)), offs[grp[-1]]))
        # B-tree levels above the leaves (<= 2 * INTERNAL_K children per node)
        lvl = 0
        node_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        while True:
            nxt = []
            groups = [level[i:i + 2 * INTERNAL_K] for i in range(0, len(level), 2 * INTERNAL_K)] or [[]]
            addrs = [len(self.buf) + k * node_size for k in range(len(groups))]
            left_key = 0
            for k, grp in enumerate(groups):
                node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, lvl, len(grp), addrs[k - 1] if k else UNDEF,
                                                       addrs[k + 1] if k + 1 < len(groups) else UNDEF))
                node += struct.pack("<Q", left_key)
                for a, key in grp:
                    node += struct.pack("<QQ", a, key)
                    left_key = key
                node += b"\0" * (node_size - len(node))
                nxt.append((self._append(bytes(node)), left_key))
            level = nxt
            lvl += 1
            if len(level) == 1:
                break
        btree = level[0][0]
        heap_addr = self._append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, len(self.buf) + 32))
        self._append(bytes(heap))
        eof = len(self.buf)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, 96, 1, 0) + struct.pack("<QQ", btree, heap_addr)
        assert len(sb) == 96
        root = struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + _msg(0x11, struct.pack("<QQ", btree, heap_addr))
        assert len(root) == 40
        if isinstance(self.buf, _FileBuf):
            self.buf.patch_head(sb + root)
            return len(self.buf)
        self.buf[0:96] = sb
        self.buf[96:136] = root
        return bytes(self.buf)


def write_hd5(path, datasets, level=4, threads=None):
    """datasets: {name -> 2-D array}; integer arrays are stored as int64, float arrays as float64 (what
    `np.array([values, index])` gives the reference).  The deflate of the datasets runs on `threads` host threads
    (default: the cores of the process, at most 16); the file is byte-identical for every thread count."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    if threads is None:
        try:
            threads = min(16, len(os.sched_getaffinity(0)))
        except AttributeError:
            threads = min(16, os.cpu_count() or 1)
    items = list(datasets.items())
    with open(path, "wb") as f:                                   # chunks go to the file as they are produced
        w = _Writer(f)
        if threads <= 1 or len(items) < 2:
            for name, arr in items:
                w.add_dataset(name, arr, level)
        else:
            with ThreadPoolExecutor(max_workers=threads) as ex:
                window, pending = 4 * threads, []                  # bounded look-ahead: compressed chunks wait in memory
                for name, arr in items:
                    pending.append((name, ex.submit(_Writer.prepare, arr, level)))
                    if len(pending) >= window:
                        nm, fut = pending.pop(0)
                        w.add_dataset(nm, None, level, prepared=fut.result())
                for nm, fut in pending:
                    w.add_dataset(nm, None, level, prepared=fut.result())
        return w.finish()


def store_special(path, obj, threads=None):
    """Mirror of SNVprofile._store_special for covT / clonT (SNVprofile.py:717-733): obj = scaffold -> mm -> Series."""
    ds = {}
    for scaff, clon in obj.items():
        for mm, arr in clon.items():
            ds["{0}::{1}".format(scaff, mm)] = np.array([np.asarray(arr.values), np.asarray(arr.index)])
    return write_hd5(path, ds, threads=threads)
