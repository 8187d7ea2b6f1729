// isb_host.cpp -- host side of the drop-in boundary: BAM -> position-major event columns (C++; no GPU involved).
//
// Replaces what the reference delegates to pysam/htslib for this path:
//   pysam.AlignmentFile(bam)                              inStrain/profile/profile_utilities.py:56
//   samfile.pileup(..., stepper='nofilter', ignore_overlaps=True, min_base_quality=30, ...)
//                                                         inStrain/profile/profile_utilities.py:150-153
// i.e. BGZF inflate + BAM record decode (SAM/BAM spec 4.2), htslib's mate-overlap quality tweak
// (bam_plp overlap_push / tweak_overlap_quality / cigar_iref2iseq_set,next of htslib 1.10, including the in-block
// counter behaviour the reference's goldens pin -- SURVEY.md Appendix A), CIGAR expansion of M/=/X bases and a stable
// counting sort into position-major order (= pileup column order).  Only reads whose name is in R2M are packed
// (others can never be counted: profile_utilities.py:277-283); the quality threshold itself is applied by the K1 kernel.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/instrain_b200.h"

namespace {

struct BamRec {
    int32_t tid, pos, mtid, mpos, isize, l_seq;
    uint16_t flag, n_cigar;
    uint32_t cig_off;     // into cigar pool
    uint64_t seq_off;     // into seq/qual pools (one byte per base)
    int32_t name_idx;     // index into the caller's name list
};

struct Bam {
    FILE *fp = nullptr;
    std::vector<uint8_t> buf;     // decompressed bytes not yet consumed
    size_t off = 0;
    bool eof = false;
    bool saw_eof_marker = false;  // the last block read was BGZF's empty end-of-file block
    bool bad = false;             // a block or record could not be read completely: `err` says why
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    char err[256] = "";
    // look-ahead record (raw bytes) of the next scaffold
    std::vector<uint8_t> pending;
    bool has_pending = false;
};

struct Events {
    std::vector<int32_t> ref_pos, read_id;
    std::vector<uint8_t> base, qual, pair_mm;
    int64_t n_reads_seen = 0, n_reads_packed = 0;
};

static bool bam_fail(Bam *b, const char *msg)
{
    if (!b->err[0]) snprintf(b->err, sizeof(b->err), "%s", msg);
    b->bad = true;
    b->eof = true;
    return false;
}

bool read_block(Bam *b)
{
    uint8_t h[18];
    size_t got = fread(h, 1, 18, b->fp);
    if (got == 0) {                                // end of the file: htslib expects the empty BGZF block right before it
        b->eof = true;
        if (!b->saw_eof_marker && !getenv("ISB_ALLOW_NO_BGZF_EOF"))
            return bam_fail(b, "no BGZF end-of-file block: the BAM is truncated (set ISB_ALLOW_NO_BGZF_EOF=1 to read it anyway)");
        return false;
    }
    if (got != 18) return bam_fail(b, "truncated BGZF block header");
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return bam_fail(b, "not a BGZF block (corrupt file?)");
    const int xlen = h[10] | (h[11] << 8);
    if (xlen < 6) return bam_fail(b, "BGZF block without BC field");
    std::vector<uint8_t> extra(xlen);
    memcpy(extra.data(), h + 12, 6);
    if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, b->fp) != (size_t)(xlen - 6)) return bam_fail(b, "truncated BGZF block header");
    int bsize = -1;
    for (int i = 0; i + 4 <= xlen;) {
        const int slen = extra[i + 2] | (extra[i + 3] << 8);
        if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = extra[i + 4] | (extra[i + 5] << 8);
        i += 4 + slen;
    }
    if (bsize < 0) return bam_fail(b, "BGZF block without BC field");
    const int clen = bsize - xlen - 19;           // compressed payload
    if (clen < 0) return bam_fail(b, "corrupt BGZF block size");
    std::vector<uint8_t> comp(clen + 8);
    if (fread(comp.data(), 1, clen + 8, b->fp) != (size_t)(clen + 8)) return bam_fail(b, "truncated BGZF block");
    const uint32_t isize = comp[clen + 4] | (comp[clen + 5] << 8) | (comp[clen + 6] << 16) | ((uint32_t)comp[clen + 7] << 24);
    b->saw_eof_marker = isize == 0;
    if (isize == 0) return true;                  // empty block (EOF marker)
    if (isize > (1u << 16)) return bam_fail(b, "corrupt BGZF block (uncompressed size > 64 KiB)");
    // compact the buffer before appending
    if (b->off > (1u << 20)) { b->buf.erase(b->buf.begin(), b->buf.begin() + b->off); b->off = 0; }
    const size_t old = b->buf.size();
    b->buf.resize(old + isize);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return bam_fail(b, "inflateInit2 failed");
    zs.next_in = comp.data(); zs.avail_in = clen;
    zs.next_out = b->buf.data() + old; zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool full = zs.avail_out == 0;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || !full) {
        b->buf.resize(old);
        char msg[96];
        snprintf(msg, sizeof(msg), "inflate failed (%d): corrupt BGZF block", rc);
        return bam_fail(b, msg);
    }
    const uint32_t crc_stored = comp[clen] | (comp[clen + 1] << 8) | (comp[clen + 2] << 16) | ((uint32_t)comp[clen + 3] << 24);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), b->buf.data() + old, isize) != crc_stored) {
        b->buf.resize(old);
        return bam_fail(b, "BGZF block CRC mismatch: corrupt file");
    }
    return true;
}

// make at least `need` unread bytes available; false at EOF
bool ensure(Bam *b, size_t need)
{
    while (b->buf.size() - b->off < need) {
        if (b->eof) return false;
        if (!read_block(b) && b->eof) return b->buf.size() - b->off >= need;
    }
    return true;
}

bool read_exact(Bam *b, void *dst, size_t n)
{
    if (!ensure(b, n)) return false;
    memcpy(dst, b->buf.data() + b->off, n);
    b->off += n;
    return true;
}

// ---- htslib overlap tweak (restated from htslib 1.10 sam.c; the parity tests hold a Python twin) ---------------
enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
inline bool is_match(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

struct Walker {
    const uint32_t *cig;
    int n, ci = 0;
    int64_t icig = 0, iseq = 0, iref = 0;
    int set(int64_t pos)
    {
        if (pos < 0) return -1;
        icig = iseq = iref = 0;
        while (ci < n) {
            const int op = cig[ci] & 0xf;
            const int64_t len = cig[ci] >> 4;
            if (op == OP_S || op == OP_I) { ++ci; iseq += len; icig = 0; }
            else if (op == OP_H || op == OP_P) { ++ci; icig = 0; }
            else if (is_match(op)) {
                pos -= len;
                if (pos < 0) { icig = len + pos; iseq += icig; iref += icig; return 0; }
                ++ci; iseq += len; icig = 0; iref += len;
            } else if (op == OP_D || op == OP_N) {
                pos -= len;
                if (pos < 0) pos = 0;
                ++ci; iref += len; icig = 0;
            } else return -2;
        }
        iseq = -1;
        return -1;
    }
    int next()
    {
        while (ci < n) {
            const int op = cig[ci] & 0xf;
            const int64_t len = cig[ci] >> 4;
            if (is_match(op)) {
                if (icig >= len - 1) { icig = 0; ++ci; continue; }
                ++iseq; ++icig; ++iref;
                return 0;
            }
            if (op == OP_D || op == OP_N) { ++ci; iref += len; icig = 0; }
            else if (op == OP_I || op == OP_S) { ++ci; iseq += len; icig = 0; }
            else if (op == OP_H || op == OP_P) { ++ci; icig = 0; }
            else return -2;
        }
        iseq = -1; iref = -1;
        return -1;
    }
};

void tweak_overlap(const BamRec &a, const BamRec &b, const uint32_t *cigs, const uint8_t *seq, uint8_t *qual)
{
    Walker wa{cigs + a.cig_off, a.n_cigar}, wb{cigs + b.cig_off, b.n_cigar};
    int64_t iref = b.pos;
    int ar = wa.set(iref - a.pos);
    if (ar < 0) return;
    int br = wb.set(iref - b.pos);
    if (br < 0) return;
    const uint8_t *sa = seq + a.seq_off, *sb = seq + b.seq_off;
    uint8_t *qa = qual + a.seq_off, *qb = qual + b.seq_off;
    for (;;) {
        while (ar >= 0 && wa.iref >= 0 && wa.iref < iref - a.pos) ar = wa.next();
        if (ar < 0) return;
        if (iref < wa.iref + a.pos) iref = wa.iref + a.pos;
        while (br >= 0 && wb.iref >= 0 && wb.iref < iref - b.pos) br = wb.next();
        if (br < 0) return;
        if (iref < wb.iref + b.pos) iref = wb.iref + b.pos;
        ++iref;
        if (wa.iref + a.pos != wb.iref + b.pos) continue;
        const int64_t ia = wa.iseq, ib = wb.iseq;
        if (sa[ia] == sb[ib]) {
            const int q = (int)qa[ia] + (int)qb[ib];
            qa[ia] = (uint8_t)(q > 200 ? 200 : q);
            qb[ib] = 0;
        } else if (qa[ia] >= qb[ib]) {
            qa[ia] = (uint8_t)(0.8 * qa[ia]);
            qb[ib] = 0;
        } else {
            qb[ib] = (uint8_t)(0.8 * qb[ib]);
            qa[ia] = 0;
        }
    }
}

int64_t ref_len_of(const uint32_t *cig, int n)
{
    int64_t l = 0;
    for (int i = 0; i < n; ++i) {
        const int op = cig[i] & 0xf;
        if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) l += cig[i] >> 4;
    }
    return l;
}

// nt16 code (=ACMGRSVTWYHKDBN) -> inStrain base code (A,C,T,G = 0..3; 4 otherwise)
const uint8_t NT16_TO_CODE[16] = {4, 0, 1, 4, 3, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4};

}  // namespace

extern "C" {

void *isb_bam_open(const char *path)
{
    Bam *b = new Bam();
    b->fp = fopen(path, "rb");
    if (!b->fp) { delete b; return nullptr; }
    char magic[4];
    int32_t l_text = 0, n_ref = 0;
    if (!read_exact(b, magic, 4) || memcmp(magic, "BAM\1", 4) != 0 || !read_exact(b, &l_text, 4)) goto fail;
    if (!ensure(b, (size_t)l_text)) goto fail;
    b->off += l_text;
    if (!read_exact(b, &n_ref, 4)) goto fail;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name = 0, l_ref = 0;
        if (!read_exact(b, &l_name, 4) || !ensure(b, (size_t)l_name)) goto fail;
        b->ref_names.emplace_back((const char *)b->buf.data() + b->off, l_name > 0 ? l_name - 1 : 0);
        b->off += l_name;
        if (!read_exact(b, &l_ref, 4)) goto fail;
        b->ref_lens.push_back(l_ref);
    }
    return b;
fail:
    fclose(b->fp);
    delete b;
    return nullptr;
}

void isb_bam_close(void *h)
{
    Bam *b = (Bam *)h;
    if (!b) return;
    if (b->fp) fclose(b->fp);
    delete b;
}

// Reposition the reader at a BGZF virtual offset (compressed block offset << 16 | offset inside the inflated block), as
// stored in a .bai index: lets several readers (one per host thread) each stream their own scaffolds of one BAM.
int isb_bam_seek(void *h, uint64_t voffset)
{
    Bam *b = (Bam *)h;
    if (!b || fseeko(b->fp, (off_t)(voffset >> 16), SEEK_SET) != 0) return -1;
    b->buf.clear();
    b->off = 0;
    b->eof = false;
    b->bad = false;
    b->err[0] = 0;
    b->saw_eof_marker = false;
    b->has_pending = false;
    b->pending.clear();
    const size_t u = (size_t)(voffset & 0xffff);
    if (u) {
        if (!ensure(b, u)) return -1;
        b->off = u;
    }
    return 0;
}

int isb_bam_n_refs(void *h) { return (int)((Bam *)h)->ref_names.size(); }
const char *isb_bam_ref_name(void *h, int tid) { return ((Bam *)h)->ref_names[tid].c_str(); }
int64_t isb_bam_ref_len(void *h, int tid) { return ((Bam *)h)->ref_lens[tid]; }
const char *isb_bam_error(void *h) { return ((Bam *)h)->err; }

// tid of the next alignment record in the (coordinate-sorted) file: >= 0, -1 = unmapped tail, -2 = clean end of file,
// -3 = read error (corrupt or truncated file; isb_bam_error has the reason -- pysam / htslib raise in this case)
int isb_bam_peek_tid(void *h)
{
    Bam *b = (Bam *)h;
    if (b->bad) return -3;
    if (!b->has_pending) {
        int32_t block_size = 0;
        if (!read_exact(b, &block_size, 4)) {
            if (b->bad) return -3;
            if (b->buf.size() != b->off) { bam_fail(b, "truncated alignment record"); return -3; }
            return -2;
        }
        if (block_size < 32 || block_size > (1 << 28)) { bam_fail(b, "corrupt alignment record (block_size)"); return -3; }
        b->pending.resize(block_size);
        if (!read_exact(b, b->pending.data(), block_size)) { bam_fail(b, "truncated alignment record"); return -3; }
        int32_t core[8];
        memcpy(core, b->pending.data(), 32);
        const int l_read_name = (uint32_t)core[2] & 0xff;
        const int64_t n_cig = (uint32_t)core[3] & 0xffff, l_seq = core[4];
        if (core[0] >= (int32_t)b->ref_names.size() || l_seq < 0 ||
            32 + l_read_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq > (int64_t)block_size) {
            bam_fail(b, "corrupt alignment record (field sizes)");
            return -3;
        }
        b->has_pending = true;
    }
    int32_t tid;
    memcpy(&tid, b->pending.data(), 4);
    return tid < 0 ? -1 : tid;
}

struct Loaded {
    std::vector<BamRec> recs;
    std::vector<uint32_t> cigs;
    std::vector<uint8_t> seq, qual;
    int64_t n_reads_seen = 0;
};

// Consume every record of scaffold `tid` (the file is coordinate-sorted, so they are contiguous), keep the reads whose
// name is in the list and apply the mate-overlap quality tweak in file order (bam_plp overlap_push, ignore_overlaps=True).
// Region form (lo < hi): only the reads that overlap the scaffold positions [lo, hi) are kept -- what an index fetch of that
// region hands htslib's pileup (inStrain/polymorpher.py:287-293; the split-run sharding of one large scaffold) -- and
// reading stops at the first record that starts at or behind hi (the reader is then left INSIDE the scaffold).
static void load_scaffold(Bam *b, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off, Loaded &ld,
                          int64_t lo = 0, int64_t hi = -1)
{
    std::unordered_map<std::string, int32_t> name2idx;
    name2idx.reserve((size_t)n_names * 2 + 16);
    for (int64_t i = 0; i < n_names; ++i)
        name2idx.emplace(std::string(names_blob + name_off[i], (size_t)(name_off[i + 1] - name_off[i])), (int32_t)i);
    std::vector<BamRec> &recs = ld.recs;
    std::vector<uint32_t> &cigs = ld.cigs;
    std::vector<uint8_t> &seq = ld.seq, &qual = ld.qual;
    for (;;) {
        const int t = isb_bam_peek_tid(b);
        if (t != tid) break;
        const uint8_t *p = b->pending.data();
        int32_t core[8];
        memcpy(core, p, 32);
        if (hi > lo && (int64_t)core[1] >= hi) break;                // sorted by position: nothing further overlaps the region
        b->has_pending = false;
        ld.n_reads_seen++;
        BamRec r;
        r.tid = core[0]; r.pos = core[1];
        const uint32_t bmq = (uint32_t)core[2];
        const int l_read_name = bmq & 0xff;
        const uint32_t fnc = (uint32_t)core[3];
        r.n_cigar = fnc & 0xffff; r.flag = fnc >> 16;
        r.l_seq = core[4]; r.mtid = core[5]; r.mpos = core[6]; r.isize = core[7];
        if (r.flag & 0x4) continue;                                  // unmapped
        const char *name = (const char *)p + 32;
        auto it = name2idx.find(std::string(name, l_read_name > 0 ? l_read_name - 1 : 0));
        if (it == name2idx.end()) continue;                          // not in R2M: can never be counted
        r.name_idx = it->second;
        const uint8_t *q = p + 32 + l_read_name;
        if (hi > lo) {                                               // ends in front of the region: not fetched
            uint32_t cg[64];
            int64_t rl = 0;
            for (uint32_t c0 = 0; c0 < r.n_cigar; c0 += 64) {
                const int nc = (int)std::min<uint32_t>(64, r.n_cigar - c0);
                memcpy(cg, q + 4u * c0, 4u * nc);
                rl += ref_len_of(cg, nc);
            }
            if ((int64_t)r.pos + rl <= lo) continue;
        }
        r.cig_off = (uint32_t)cigs.size();
        cigs.resize(cigs.size() + r.n_cigar);
        memcpy(cigs.data() + r.cig_off, q, 4u * r.n_cigar);
        q += 4u * r.n_cigar;
        r.seq_off = seq.size();
        seq.resize(seq.size() + r.l_seq);
        qual.resize(qual.size() + r.l_seq);
        for (int i = 0; i < r.l_seq; ++i) seq[r.seq_off + i] = (q[i >> 1] >> ((~i & 1) << 2)) & 0xf;
        q += (r.l_seq + 1) >> 1;
        memcpy(qual.data() + r.seq_off, q, r.l_seq);
        recs.push_back(r);
    }
    std::unordered_map<int32_t, int32_t> pending;                    // name idx -> first mate record
    for (size_t i = 0; i < recs.size(); ++i) {
        const BamRec &r = recs[i];
        if ((r.flag & 0x8) || !(r.flag & 0x2)) continue;
        const int64_t end = r.pos + ref_len_of(cigs.data() + r.cig_off, r.n_cigar);
        const int64_t aisize = r.isize < 0 ? -(int64_t)r.isize : r.isize;
        if ((r.mtid >= 0 && r.tid != r.mtid) || (aisize >= 2 * (int64_t)r.l_seq && r.mpos >= end)) continue;
        auto it = pending.find(r.name_idx);
        if (it == pending.end()) {
            if (r.mpos >= r.pos) pending.emplace(r.name_idx, (int32_t)i);
        } else {
            tweak_overlap(recs[it->second], r, cigs.data(), seq.data(), qual.data());
            pending.erase(it);
        }
    }
}

// Pack the reads of scaffold `tid` whose name is in the given list as position-major event columns.
// names: n_names NUL-free strings concatenated in names_blob, name_off[n_names+1];
// name_mm[i] = R2M value (0 in set mode).  Coordinates are shifted by pos_offset, pair ids start at pair_id_offset
// and follow BAM order of first appearance.  Returns an opaque result (isb_events_*), or NULL on error.
void *isb_pack_scaffold(void *h, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                        const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset)
{
    Bam *b = (Bam *)h;
    Loaded ld;
    load_scaffold(b, tid, n_names, names_blob, name_off, ld);
    if (b->bad) return nullptr;                                   // corrupt / truncated BAM: isb_bam_error() has the reason
    std::vector<BamRec> &recs = ld.recs;
    std::vector<uint32_t> &cigs = ld.cigs;
    std::vector<uint8_t> &seq = ld.seq, &qual = ld.qual;
    Events *ev = new Events();
    ev->n_reads_seen = ld.n_reads_seen;
    // expand M/=/X bases in file order, then stable counting sort by position
    const int64_t L = tid >= 0 && tid < (int)b->ref_lens.size() ? b->ref_lens[tid] : 0;
    std::vector<int32_t> pair_of_name((size_t)n_names, -1);
    std::vector<int64_t> cnt((size_t)L + 1, 0);
    int32_t next_pair = 0;
    for (const BamRec &r : recs) {                                   // pass 1: coverage histogram + pair ids
        if (pair_of_name[r.name_idx] < 0) {
            pair_of_name[r.name_idx] = next_pair++;
            ev->pair_mm.push_back(name_mm ? name_mm[r.name_idx] : 0);
        }
        int64_t pos = r.pos;
        for (int c = 0; c < r.n_cigar; ++c) {
            const int op = cigs[r.cig_off + c] & 0xf;
            const int64_t len = cigs[r.cig_off + c] >> 4;
            if (is_match(op)) {
                for (int64_t k = 0; k < len; ++k)
                    if (pos + k >= 0 && pos + k < L) cnt[pos + k]++;
                pos += len;
            } else if (op == OP_D || op == OP_N) pos += len;
        }
        ev->n_reads_packed++;
    }
    int64_t total = 0;
    for (int64_t i = 0; i <= L; ++i) { const int64_t c = cnt[i]; cnt[i] = total; total += c; }
    ev->ref_pos.resize(total); ev->read_id.resize(total); ev->base.resize(total); ev->qual.resize(total);
    for (const BamRec &r : recs) {                                   // pass 2: scatter (file order => stable)
        int64_t pos = r.pos, qp = 0;
        const int32_t pid = pair_of_name[r.name_idx] + pair_id_offset;
        for (int c = 0; c < r.n_cigar; ++c) {
            const int op = cigs[r.cig_off + c] & 0xf;
            const int64_t len = cigs[r.cig_off + c] >> 4;
            if (is_match(op)) {
                for (int64_t k = 0; k < len; ++k) {
                    if (pos + k < 0 || pos + k >= L) continue;
                    const int64_t w = cnt[pos + k]++;
                    ev->ref_pos[w] = (int32_t)(pos + k) + pos_offset;
                    ev->read_id[w] = pid;
                    ev->base[w] = NT16_TO_CODE[seq[r.seq_off + qp + k]];
                    ev->qual[w] = qual[r.seq_off + qp + k];
                }
                pos += len; qp += len;
            } else if (op == OP_I || op == OP_S) qp += len;
            else if (op == OP_D || op == OP_N) pos += len;
        }
    }
    return ev;
}

int64_t isb_events_count(void *e) { return (int64_t)((Events *)e)->ref_pos.size(); }
int64_t isb_events_pairs(void *e) { return (int64_t)((Events *)e)->pair_mm.size(); }
int64_t isb_events_reads_seen(void *e) { return ((Events *)e)->n_reads_seen; }
int64_t isb_events_reads_packed(void *e) { return ((Events *)e)->n_reads_packed; }
// copy the columns out (any pointer may be NULL)
void isb_events_copy(void *e, int32_t *ref_pos, uint8_t *base, uint8_t *qual, int32_t *read_id, uint8_t *pair_mm)
{
    Events *ev = (Events *)e;
    const size_t n = ev->ref_pos.size();
    if (ref_pos) memcpy(ref_pos, ev->ref_pos.data(), n * 4);
    if (base) memcpy(base, ev->base.data(), n);
    if (qual) memcpy(qual, ev->qual.data(), n);
    if (read_id) memcpy(read_id, ev->read_id.data(), n * 4);
    if (pair_mm) memcpy(pair_mm, ev->pair_mm.data(), ev->pair_mm.size());
}
void isb_events_free(void *e) { delete (Events *)e; }

// ---- read-major packer: the same reads as aligned segments (include/instrain_b200.h, isb_reads_batch) -------------------
// One segment per CIGAR M/=/X block (clipped to the scaffold, split at 256 bases), sorted by start; one-hot 4-bit codes
// (A=1,C=2,T=4,G=8) for the bases whose quality after the overlap tweak is >= min_qual, 0 otherwise; passing non-ACGT
// bases go to the N-event list.  The word stream of the scaffold is [data words + one zero word] per segment, the data
// words position-aligned in batch coordinates (word 0 of a segment covers [start & ~7, (start & ~7) + 8)); the caller
// concatenates scaffolds behind one leading zero word (isb_reads_copy's word_base).
struct ReadsOut {
    std::vector<int32_t> seg_start, seg_pair, nev_pos, nev_pair;
    std::vector<uint16_t> seg_len;
    std::vector<int64_t> seg_word;                                   // relative to the scaffold's stream
    std::vector<uint32_t> words;
    std::vector<uint8_t> pair_mm;
    int64_t n_reads_seen = 0, n_reads_packed = 0, n_events = 0;
    int max_len = 1;
};

void *isb_pack_scaffold_reads_region(void *h, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                                     const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset, int min_qual,
                                     int64_t lo, int64_t hi);

void *isb_pack_scaffold_reads(void *h, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                              const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset, int min_qual)
{
    return isb_pack_scaffold_reads_region(h, tid, n_names, names_blob, name_off, name_mm, pos_offset, pair_id_offset, min_qual, 0, -1);
}

void *isb_pack_scaffold_reads_region(void *h, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                                     const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset, int min_qual,
                                     int64_t lo, int64_t hi)
{
    Bam *b = (Bam *)h;
    Loaded ld;
    load_scaffold(b, tid, n_names, names_blob, name_off, ld, lo, hi);
    if (b->bad) return nullptr;                                   // corrupt / truncated BAM: isb_bam_error() has the reason
    ReadsOut *out = new ReadsOut();
    out->n_reads_seen = ld.n_reads_seen;
    const int64_t L = tid >= 0 && tid < (int)b->ref_lens.size() ? b->ref_lens[tid] : 0;
    std::vector<int32_t> pair_of_name((size_t)n_names, -1);
    int32_t next_pair = 0;
    struct Seg { int32_t start; uint16_t len; int32_t pair; int64_t src; };   // src: offset into the record's seq/qual
    std::vector<Seg> segs;
    for (const BamRec &r : ld.recs) {
        if (pair_of_name[r.name_idx] < 0) {
            pair_of_name[r.name_idx] = next_pair++;
            out->pair_mm.push_back(name_mm ? name_mm[r.name_idx] : 0);
        }
        const int32_t pid = pair_of_name[r.name_idx] + pair_id_offset;
        int64_t pos = r.pos, qp = 0;
        for (int c = 0; c < r.n_cigar; ++c) {
            const int op = ld.cigs[r.cig_off + c] & 0xf;
            const int64_t len = ld.cigs[r.cig_off + c] >> 4;
            if (is_match(op)) {
                int64_t lo = pos < 0 ? 0 : pos, hi = pos + len > L ? L : pos + len;     // clip to the scaffold
                for (int64_t s0 = lo; s0 < hi; s0 += 256) {
                    const int64_t n = hi - s0 < 256 ? hi - s0 : 256;
                    segs.push_back(Seg{(int32_t)s0, (uint16_t)n, pid, (int64_t)r.seq_off + qp + (s0 - pos)});
                }
                pos += len; qp += len;
            } else if (op == OP_I || op == OP_S) qp += len;
            else if (op == OP_D || op == OP_N) pos += len;
        }
        out->n_reads_packed++;
    }
    std::vector<int32_t> order(segs.size());
    for (size_t i = 0; i < segs.size(); ++i) order[i] = (int32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return segs[x].start < segs[y].start; });
    const size_t n = segs.size();
    out->seg_start.resize(n); out->seg_len.resize(n); out->seg_pair.resize(n); out->seg_word.resize(n);
    int64_t w = 0;
    for (size_t i = 0; i < n; ++i) {
        const Seg &sg = segs[order[i]];
        out->seg_start[i] = sg.start + pos_offset;
        out->seg_len[i] = sg.len;
        out->seg_pair[i] = sg.pair;
        out->seg_word[i] = w;
        w += ((int64_t)((sg.start + pos_offset) & 7) + sg.len + 7) / 8 + 1;      // position-aligned words + one zero word
        if (sg.len > out->max_len) out->max_len = sg.len;
    }
    out->words.assign((size_t)w, 0u);
    for (size_t i = 0; i < n; ++i) {
        const Seg &sg = segs[order[i]];
        uint32_t *dst = out->words.data() + out->seg_word[i];
        const int sh = (sg.start + pos_offset) & 7;                     // nibble of the first base in word 0
        for (int j = 0; j < sg.len; ++j) {
            if ((int)ld.qual[sg.src + j] < min_qual) continue;
            const int code = NT16_TO_CODE[ld.seq[sg.src + j]];
            out->n_events++;
            if (code < 4) dst[(j + sh) >> 3] |= (1u << code) << (((j + sh) & 7) << 2);
            else { out->nev_pos.push_back(sg.start + j + pos_offset); out->nev_pair.push_back(sg.pair); }
        }
    }
    return out;
}

int64_t isb_reads_segs(void *r) { return (int64_t)((ReadsOut *)r)->seg_start.size(); }
int64_t isb_reads_stream_words(void *r) { return (int64_t)((ReadsOut *)r)->words.size(); }
int64_t isb_reads_pairs(void *r) { return (int64_t)((ReadsOut *)r)->pair_mm.size(); }
int64_t isb_reads_n_events(void *r) { return ((ReadsOut *)r)->n_events; }
int64_t isb_reads_nev(void *r) { return (int64_t)((ReadsOut *)r)->nev_pos.size(); }
int isb_reads_max_len(void *r) { return ((ReadsOut *)r)->max_len; }
int64_t isb_reads_reads_seen(void *r) { return ((ReadsOut *)r)->n_reads_seen; }
int64_t isb_reads_reads_packed(void *r) { return ((ReadsOut *)r)->n_reads_packed; }
// copy out; seg_word values are rebased to word_base (the index in the batch stream where this scaffold's stream starts)
void isb_reads_copy(void *r, int32_t *seg_start, uint16_t *seg_len, int32_t *seg_pair, int64_t *seg_word, uint32_t *words,
                    int32_t *nev_pos, int32_t *nev_pair, uint8_t *pair_mm, int64_t word_base)
{
    ReadsOut *o = (ReadsOut *)r;
    const size_t n = o->seg_start.size();
    if (seg_start) memcpy(seg_start, o->seg_start.data(), n * 4);
    if (seg_len) memcpy(seg_len, o->seg_len.data(), n * 2);
    if (seg_pair) memcpy(seg_pair, o->seg_pair.data(), n * 4);
    if (seg_word) for (size_t i = 0; i < n; ++i) seg_word[i] = o->seg_word[i] + word_base;
    if (words) memcpy(words, o->words.data(), o->words.size() * 4);
    if (nev_pos) memcpy(nev_pos, o->nev_pos.data(), o->nev_pos.size() * 4);
    if (nev_pair) memcpy(nev_pair, o->nev_pair.data(), o->nev_pair.size() * 4);
    if (pair_mm) memcpy(pair_mm, o->pair_mm.data(), o->pair_mm.size());
}
void isb_reads_free(void *r) { delete (ReadsOut *)r; }

}  // extern "C"

// ---- read filter: BAM -> sR2M (the hot path's input) ---------------------------------------------------------------
// C++ restatement of the reference's default read filter:
//   get_paired_reads        inStrain/filter_reads.py:885-956   per scaffold, name -> (NM sum, insert, max mapq, length, #reads)
//   paired_read_filter      inStrain/filter_reads.py:471-532   pairing_filter = 'paired_only' (exactly two reads of the name on
//                                                              the scaffold), no priority reads
//   filter_scaff2pair2info  inStrain/filter_reads.py:201-300   max_insert = max_insert_relative * median insert over ALL pairs
//   evaluate_pair           inStrain/filter_reads.py:387-426   1 - nm/length > min_read_ani, mapq > min_mapq,
//                                                              min_insert < insert < max_insert
// One sequential pass over the BAM (it shares the BGZF reader with the packer, so a second pysam pass disappears).
namespace {

struct PairInfo {
    int64_t nm = 0, insert = -1, mapq = 0, length = 0;
    int32_t reads = 0;
    int32_t first = 0, last = 0;       // aligned span of the first read
    uint8_t keep = 0;
    // what the pairing filter hands to the thresholds (isb_filter_apply2): selected, and the possibly merged info
    uint8_t sel = 0;
    int64_t e_nm = 0, e_insert = -1, e_mapq = 0, e_length = 0;
    int32_t e_reads = 0;
};

struct ScaffoldPairs {
    std::vector<std::string> names;                  // insertion (file) order
    std::vector<PairInfo> info;
    std::unordered_map<std::string, int32_t> index;
    int64_t tally[6] = {0, 0, 0, 0, 0, 0};           // pass_pairing_filter, pass_min_read_ani, pass_max_insert, pass_min_insert, pass_min_mapq, filtered_pairs
    int64_t tally2[3] = {0, 0, 0};                   // unfiltered_priority_reads, filtered_singletons, filtered_priority_reads
};

struct Filter {
    std::vector<std::string> ref_names;
    std::vector<ScaffoldPairs> sc;
    double max_insert = 0.0;
    char err[256] = "";
};

int nm_tag(const uint8_t *p, const uint8_t *end)
{
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        int64_t val = 0;
        size_t size = 0;
        switch (ty) {
            case 'A': size = 1; break;
            case 'c': size = 1; val = (int8_t)p[0]; break;
            case 'C': size = 1; val = p[0]; break;
            case 's': size = 2; { int16_t v; memcpy(&v, p, 2); val = v; } break;
            case 'S': size = 2; { uint16_t v; memcpy(&v, p, 2); val = v; } break;
            case 'i': size = 4; { int32_t v; memcpy(&v, p, 4); val = v; } break;
            case 'I': size = 4; { uint32_t v; memcpy(&v, p, 4); val = v; } break;
            case 'f': size = 4; break;
            case 'Z': case 'H': { const uint8_t *q = p; while (q < end && *q) ++q; size = (size_t)(q - p) + 1; } break;
            case 'B': {
                const char sub = (char)p[0];
                int32_t n; memcpy(&n, p + 1, 4);
                const int w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                size = 5 + (size_t)n * w;
            } break;
            default: return 0;
        }
        if (t0 == 'N' && t1 == 'M') return (int)val;
        p += size;
    }
    return 0;
}

}  // namespace

extern "C" {

static thread_local char g_host_err[320] = "";
const char *isb_host_last_error(void) { return g_host_err; }

// one alignment record (the reader's pending record) into its scaffold's pair table: get_paired_reads (filter_reads.py:885-956)
static void filter_consume(Bam *b, ScaffoldPairs *spp)
{
    const uint8_t *p = b->pending.data();
    const uint8_t *end = p + b->pending.size();
    b->has_pending = false;
    int32_t core[8];
    memcpy(core, p, 32);
    const int l_read_name = (uint32_t)core[2] & 0xff;
    const int mapq = ((uint32_t)core[2] >> 8) & 0xff;
    const int n_cigar = (uint32_t)core[3] & 0xffff, flag = (uint32_t)core[3] >> 16;
    const int l_seq = core[4];
    if (flag & 0x4 || n_cigar == 0) return;
    const uint8_t *q = p + 32 + l_read_name;
    int64_t pos = core[1], first = -1, last = -1, qlen = 0;
    for (int c = 0; c < n_cigar; ++c) {
        uint32_t cg; memcpy(&cg, q + 4 * c, 4);
        const int op = cg & 0xf;
        const int64_t len = cg >> 4;
        if (is_match(op)) { if (first < 0) first = pos; last = pos + len - 1; pos += len; qlen += len; }
        else if (op == OP_D || op == OP_N) pos += len;
        else if (op == OP_I || op == OP_S) qlen += len;
    }
    if (first < 0) return;                                      // get_reference_positions() == []
    const uint8_t *tags = q + 4 * n_cigar + ((l_seq + 1) >> 1) + l_seq;
    const int nm = nm_tag(tags, end);
    ScaffoldPairs &sp = *spp;
    std::string name((const char *)p + 32, l_read_name > 0 ? l_read_name - 1 : 0);
    auto it = sp.index.find(name);
    if (it == sp.index.end()) {
        PairInfo pi;
        pi.nm = nm; pi.insert = -1; pi.mapq = mapq; pi.length = qlen; pi.reads = 1;
        pi.first = (int32_t)first; pi.last = (int32_t)last;
        sp.index.emplace(name, (int32_t)sp.info.size());
        sp.names.push_back(std::move(name));
        sp.info.push_back(pi);
    } else {
        PairInfo &pi = sp.info[it->second];
        pi.nm += nm;
        pi.reads += 1;
        pi.length += qlen;
        if (mapq > pi.mapq) pi.mapq = mapq;
        if (pi.reads == 2) pi.insert = last > pi.first ? last - pi.first : (int64_t)pi.last - first;
        else pi.insert = -1;
        pi.first = pi.last = 0;
    }
}

// Pass over the whole BAM; returns a filter handle (NULL on failure: isb_host_last_error() says why).
void *isb_filter_open(const char *bam_path)
{
    g_host_err[0] = 0;
    Bam *b = (Bam *)isb_bam_open(bam_path);
    if (!b) { snprintf(g_host_err, sizeof(g_host_err), "cannot open BAM %s", bam_path); return nullptr; }
    Filter *f = new Filter();
    f->ref_names = b->ref_names;
    f->sc.resize(b->ref_names.size());
    for (;;) {
        const int t = isb_bam_peek_tid(b);
        if (t < 0) break;
        filter_consume(b, &f->sc[t]);
    }
    if (b->bad) {                                                 // a partial pass must not look like a complete one
        snprintf(g_host_err, sizeof(g_host_err), "%s: %s", bam_path, b->err);
        isb_bam_close(b);
        delete f;
        return nullptr;
    }
    isb_bam_close(b);
    return f;
}

// The same pass on n_threads host threads: scaffolds are independent until the thresholds are applied, so every thread
// owns a reader, takes the next scaffold with alignments (first_voffset[tid] = BGZF virtual offset of its first record from
// the .bai index, 0 = none) and fills that scaffold's pair table.  Identical tables to isb_filter_open.
void *isb_filter_open_mt(const char *bam_path, int n_threads, int n_refs, const uint64_t *first_voffset)
{
    g_host_err[0] = 0;
    if (n_threads <= 1 || !first_voffset) return isb_filter_open(bam_path);
    Bam *b0 = (Bam *)isb_bam_open(bam_path);
    if (!b0) { snprintf(g_host_err, sizeof(g_host_err), "cannot open BAM %s", bam_path); return nullptr; }
    if ((int)b0->ref_names.size() != n_refs) {
        snprintf(g_host_err, sizeof(g_host_err), "%s: index lists %d references, the header %d", bam_path, n_refs, (int)b0->ref_names.size());
        isb_bam_close(b0);
        return nullptr;
    }
    Filter *f = new Filter();
    f->ref_names = b0->ref_names;
    f->sc.resize(b0->ref_names.size());
    isb_bam_close(b0);
    std::atomic<int> next(0);
    std::mutex err_mu;
    std::string err;
    auto work = [&]() {
        Bam *b = (Bam *)isb_bam_open(bam_path);
        if (!b) { std::lock_guard<std::mutex> g(err_mu); err = "cannot open BAM"; return; }
        for (;;) {
            const int tid = next.fetch_add(1);
            if (tid >= n_refs) break;
            if (!first_voffset[tid]) continue;
            if (isb_bam_seek(b, first_voffset[tid]) != 0) { std::lock_guard<std::mutex> g(err_mu); err = "seek failed"; break; }
            for (;;) {
                const int t = isb_bam_peek_tid(b);
                if (t != tid) break;                              // next scaffold, unmapped tail, end of file or an error
                filter_consume(b, &f->sc[tid]);
            }
            if (b->bad) { std::lock_guard<std::mutex> g(err_mu); err = b->err; break; }
        }
        isb_bam_close(b);
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
    for (auto &t : th) t.join();
    if (!err.empty()) {
        snprintf(g_host_err, sizeof(g_host_err), "%s: %s", bam_path, err.c_str());
        delete f;
        return nullptr;
    }
    return f;
}

// Pairing filter + thresholds, all modes of the reference (paired_read_filter, filter_reads.py:471-532, then
// filter_scaff2pair2info, :201-300).  pairing_mode: 0 = paired_only, 1 = non_discordant, 2 = all_reads; priority reads
// (n_priority names in names_blob / name_off) pass the pairing filter regardless.  Scaffolds in header order, names in
// file order -- the reference's dict orders.  Returns the number of kept names, -1 for an unknown mode.
int64_t isb_filter_apply2(void *h, double min_read_ani, int min_mapq, double max_insert_relative, int min_insert,
                          int pairing_mode, int64_t n_priority, const char *names_blob, const int64_t *name_off)
{
    Filter *f = (Filter *)h;
    if (pairing_mode < 0 || pairing_mode > 2) return -1;
    std::unordered_map<std::string, int> priority;
    for (int64_t i = 0; i < n_priority; ++i)
        priority.emplace(std::string(names_blob + name_off[i], (size_t)(name_off[i + 1] - name_off[i])), 1);
    std::unordered_map<std::string, std::pair<int32_t, int32_t>> where;          // name -> (scaffold, index) of its first copy
    for (size_t s = 0; s < f->sc.size(); ++s) {
        ScaffoldPairs &sp = f->sc[s];
        sp.tally2[0] = sp.tally2[1] = sp.tally2[2] = 0;
        for (size_t k = 0; k < sp.info.size(); ++k) {
            PairInfo &pi = sp.info[k];
            pi.sel = 0; pi.keep = 0;
            pi.e_nm = pi.nm; pi.e_insert = pi.insert; pi.e_mapq = pi.mapq; pi.e_length = pi.length; pi.e_reads = pi.reads;
            const bool prio = !priority.empty() && priority.count(sp.names[k]);
            sp.tally2[0] += prio;
            if (pairing_mode == 0) {
                pi.sel = pi.reads == 2 || prio;
            } else if (pairing_mode == 1) {
                auto it = where.find(sp.names[k]);
                if (it == where.end() || prio) {
                    pi.sel = 1;
                    where[sp.names[k]] = {(int32_t)s, (int32_t)k};
                } else {                                                         // discordant: drop the earlier copy too
                    f->sc[it->second.first].info[it->second.second].sel = 0;
                }
            } else {
                auto it = where.find(sp.names[k]);
                if (it != where.end()) {                                         // _merge_info with the first copy
                    PairInfo &o = f->sc[it->second.first].info[it->second.second];
                    const int64_t nm = pi.e_nm + o.e_nm, mq = pi.e_mapq + o.e_mapq, ln = pi.e_length + o.e_length;
                    const int32_t rd = pi.e_reads + o.e_reads;
                    pi.e_nm = o.e_nm = nm; pi.e_insert = o.e_insert = -2; pi.e_mapq = o.e_mapq = mq;
                    pi.e_length = o.e_length = ln; pi.e_reads = o.e_reads = rd;
                    pi.sel = 1;
                } else {
                    where[sp.names[k]] = {(int32_t)s, (int32_t)k};
                    pi.sel = 1;
                }
            }
        }
    }
    std::vector<int64_t> ins;
    for (auto &sp : f->sc)
        for (auto &pi : sp.info)
            if (pi.sel && pi.e_reads == 2) ins.push_back(pi.e_insert);
    double median = 0.0;
    if (!ins.empty()) {                                               // np.median
        const size_t n = ins.size(), k = n / 2;
        std::nth_element(ins.begin(), ins.begin() + k, ins.end());
        median = (double)ins[k];
        if (n % 2 == 0) median = ((double)*std::max_element(ins.begin(), ins.begin() + k) + (double)ins[k]) / 2.0;
    }
    f->max_insert = median * max_insert_relative;
    int64_t kept = 0;
    for (auto &sp : f->sc) {
        for (int i = 0; i < 6; ++i) sp.tally[i] = 0;
        for (size_t k = 0; k < sp.info.size(); ++k) {
            PairInfo &pi = sp.info[k];
            if (!pi.sel) continue;
            sp.tally[0]++;
            const bool f_ani = (1.0 - (double)pi.e_nm / (double)pi.e_length) > min_read_ani;
            const bool f_mapq = pi.e_mapq > min_mapq;
            bool f_min = true, f_max = true;
            if (pi.e_reads == 2 && pi.e_insert != -1) { f_min = pi.e_insert > min_insert; f_max = (double)pi.e_insert < f->max_insert; }
            sp.tally[1] += f_ani; sp.tally[2] += f_max; sp.tally[3] += f_min; sp.tally[4] += f_mapq;
            if (f_ani && f_mapq && f_min && f_max) {
                pi.keep = 1;
                sp.tally[5]++;
                ++kept;
                sp.tally2[1] += pi.e_reads == 1;
                sp.tally2[2] += !priority.empty() && priority.count(sp.names[k]);
            }
        }
    }
    return kept;
}

void isb_filter_tally2(void *h, int tid, int64_t out[3]) { memcpy(out, ((Filter *)h)->sc[tid].tally2, sizeof(int64_t) * 3); }

// Apply the thresholds (reference defaults: 0.95, -1, 3, 50).  Returns the number of kept pairs.
int64_t isb_filter_apply(void *h, double min_read_ani, int min_mapq, double max_insert_relative, int min_insert)
{
    Filter *f = (Filter *)h;
    std::vector<int64_t> ins;
    for (auto &sp : f->sc)
        for (auto &pi : sp.info)
            if (pi.reads == 2) ins.push_back(pi.insert);
    double median = 0.0;
    if (!ins.empty()) {                                               // np.median
        std::vector<int64_t> v = ins;
        const size_t n = v.size(), k = n / 2;
        std::nth_element(v.begin(), v.begin() + k, v.end());
        median = (double)v[k];
        if (n % 2 == 0) {
            const int64_t lo = *std::max_element(v.begin(), v.begin() + k);
            median = ((double)lo + (double)v[k]) / 2.0;
        }
    }
    f->max_insert = median * max_insert_relative;
    int64_t kept = 0;
    for (auto &sp : f->sc) {
        for (int i = 0; i < 6; ++i) sp.tally[i] = 0;
        for (auto &pi : sp.info) {
            pi.keep = 0;
            if (pi.reads != 2) continue;                              // paired_only
            sp.tally[0]++;
            const bool f_ani = (1.0 - (double)pi.nm / (double)pi.length) > min_read_ani;
            const bool f_mapq = pi.mapq > min_mapq;
            bool f_min = true, f_max = true;
            if (pi.insert != -1) { f_min = pi.insert > min_insert; f_max = (double)pi.insert < f->max_insert; }
            sp.tally[1] += f_ani; sp.tally[2] += f_max; sp.tally[3] += f_min; sp.tally[4] += f_mapq;
            if (f_ani && f_mapq && f_min && f_max) { pi.keep = 1; pi.e_nm = pi.nm; sp.tally[5]++; ++kept; }
        }
    }
    return kept;
}

// Per-scaffold columns of the reference's mapping_info table that are not threshold tallies (paired_read_filter tallies,
// filter_reads.py:484-502, and the means / median of filter_scaff2pair2info, :253-266, taken over the pairs that pass the
// pairing filter): out[10] = unfiltered_reads, unfiltered_pairs, unfiltered_singletons, mean_mistmaches,
// mean_insert_distance, mean_mapq_score, mean_pair_length, mean_PID, median_insert, number of pairs averaged over.
void isb_filter_stats(void *h, int tid, double out[10])
{
    const ScaffoldPairs &sp = ((Filter *)h)->sc[tid];
    double reads = 0, pairs = 0, singles = 0, s_nm = 0, s_ins = 0, s_mapq = 0, s_len = 0, s_pid = 0;
    std::vector<int64_t> ins;
    for (const PairInfo &pi : sp.info) {
        reads += pi.reads;
        if (pi.reads == 1) singles += 1;
        if (pi.reads != 2) continue;
        pairs += 1;
        s_nm += (double)pi.nm; s_ins += (double)pi.insert; s_mapq += (double)pi.mapq; s_len += (double)pi.length;
        s_pid += 1.0 - (double)pi.nm / (double)pi.length;
        ins.push_back(pi.insert);
    }
    double median = 0.0 / 0.0;
    if (!ins.empty()) {
        const size_t n = ins.size(), k = n / 2;
        std::nth_element(ins.begin(), ins.begin() + k, ins.end());
        median = (double)ins[k];
        if (n % 2 == 0) median = ((double)*std::max_element(ins.begin(), ins.begin() + k) + (double)ins[k]) / 2.0;
    }
    const double nan = 0.0 / 0.0;
    out[0] = reads; out[1] = pairs; out[2] = singles;
    out[3] = pairs ? s_nm / pairs : nan; out[4] = pairs ? s_ins / pairs : nan; out[5] = pairs ? s_mapq / pairs : nan;
    out[6] = pairs ? s_len / pairs : nan; out[7] = pairs ? s_pid / pairs : nan; out[8] = median; out[9] = pairs;
}

// The same columns for ANY pairing filter: means / median over the entries the last isb_filter_apply2 selected (the
// scaffold's pair2info after paired_read_filter: singletons and priority reads included when the mode keeps them, values
// merged over scaffolds in all_reads mode), as filter_scaff2pair2info takes them (filter_reads.py:262-276).
void isb_filter_stats2(void *h, int tid, double out[10])
{
    const ScaffoldPairs &sp = ((Filter *)h)->sc[tid];
    double reads = 0, pairs = 0, singles = 0, n = 0, s_nm = 0, s_ins = 0, s_mapq = 0, s_len = 0, s_pid = 0;
    std::vector<int64_t> ins;
    for (const PairInfo &pi : sp.info) {
        reads += pi.reads;
        if (pi.reads == 1) singles += 1;
        if (pi.reads == 2) pairs += 1;
        if (!pi.sel) continue;
        n += 1;
        s_nm += (double)pi.e_nm; s_ins += (double)pi.e_insert; s_mapq += (double)pi.e_mapq; s_len += (double)pi.e_length;
        s_pid += 1.0 - (double)pi.e_nm / (double)pi.e_length;
        ins.push_back(pi.e_insert);
    }
    double median = 0.0 / 0.0;
    if (!ins.empty()) {
        const size_t m = ins.size(), k = m / 2;
        std::nth_element(ins.begin(), ins.begin() + k, ins.end());
        median = (double)ins[k];
        if (m % 2 == 0) median = ((double)*std::max_element(ins.begin(), ins.begin() + k) + (double)ins[k]) / 2.0;
    }
    const double nan = 0.0 / 0.0;
    out[0] = reads; out[1] = pairs; out[2] = singles;
    out[3] = n ? s_nm / n : nan; out[4] = n ? s_ins / n : nan; out[5] = n ? s_mapq / n : nan;
    out[6] = n ? s_len / n : nan; out[7] = n ? s_pid / n : nan; out[8] = median; out[9] = n;
}

int isb_filter_n_refs(void *h) { return (int)((Filter *)h)->sc.size(); }
double isb_filter_max_insert(void *h) { return ((Filter *)h)->max_insert; }
void isb_filter_tally(void *h, int tid, int64_t out[6]) { memcpy(out, ((Filter *)h)->sc[tid].tally, sizeof(int64_t) * 6); }
int64_t isb_filter_n_pairs(void *h, int tid) { return ((Filter *)h)->sc[tid].tally[5]; }
int64_t isb_filter_names_bytes(void *h, int tid)
{
    const ScaffoldPairs &sp = ((Filter *)h)->sc[tid];
    int64_t n = 0;
    for (size_t i = 0; i < sp.info.size(); ++i) if (sp.info[i].keep) n += (int64_t)sp.names[i].size();
    return n;
}
// kept pairs of scaffold tid in file order: concatenated names, offsets [n+1], summed NM
void isb_filter_copy(void *h, int tid, char *names_blob, int64_t *name_off, int32_t *mm)
{
    const ScaffoldPairs &sp = ((Filter *)h)->sc[tid];
    int64_t off = 0, k = 0;
    for (size_t i = 0; i < sp.info.size(); ++i) {
        if (!sp.info[i].keep) continue;
        name_off[k] = off;
        memcpy(names_blob + off, sp.names[i].data(), sp.names[i].size());
        off += (int64_t)sp.names[i].size();
        mm[k] = (int32_t)sp.info[i].e_nm;                         // summed NM (of both scaffolds' reads under all_reads)
        ++k;
    }
    name_off[k] = off;
}
void isb_filter_free(void *h) { delete (Filter *)h; }

}  // extern "C"

// ---- read-major segments -> column words (host side of isb_cols_from_reads; include/instrain_b200.h, isb_cols_batch) ----
// The packer's transposition at word granularity: every data word of a segment goes to the list of the column word it
// covers, lists keep table (= BAM) order, the 8 lists of a group are padded to the group's depth and interleaved in
// 32-byte units.  Two passes over the segment table (count, fill).
extern "C" int64_t isb_cols_from_reads_host(int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len,
                                            const int32_t *seg_pair, const int64_t *seg_word, const uint32_t *words_in,
                                            int64_t n_words_in, int32_t start, int32_t L, int64_t *grp_off, uint32_t *words,
                                            int32_t *ids, int64_t cap_chunks)
{
    if ((start & 7) || L < 0 || !grp_off || n_segs < 0) return -1;
    const int LN = ISB_COLS_LANES;
    const int64_t n_groups = ((int64_t)L + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP;
    std::vector<int32_t> cnt((size_t)n_groups * LN, 0);
    for (int64_t i = 0; i < n_segs; ++i) {
        const int64_t s = seg_start[i], n = seg_len[i];
        const int64_t nw = ((s & 7) + n + 7) >> 3;
        if (n < 1 || n > 256 || s < start || s + n > (int64_t)start + L || seg_word[i] < 0 || seg_word[i] + nw > n_words_in ||
            (i > 0 && seg_start[i - 1] > s))
            return -1;
        const int64_t c_lo = (s - start) >> 3;
        for (int64_t k = 0; k < nw; ++k) cnt[(size_t)(c_lo + k)]++;
    }
    grp_off[0] = 0;
    for (int64_t g = 0; g < n_groups; ++g) {
        int mx = 0;
        for (int l = 0; l < LN; ++l) mx = std::max(mx, (int)cnt[(size_t)g * LN + l]);
        grp_off[g + 1] = grp_off[g] + (mx + ISB_COLS_UNIT - 1) / ISB_COLS_UNIT;
    }
    const int64_t n_chunks = grp_off[n_groups];
    if (!words || !ids || n_chunks > cap_chunks) return n_chunks;
    std::fill(words, words + n_chunks * ISB_COLS_CHUNK, 0u);
    std::fill(ids, ids + n_chunks * ISB_COLS_CHUNK, -1);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int64_t i = 0; i < n_segs; ++i) {
        const int64_t s = seg_start[i], n = seg_len[i];
        const int64_t nw = ((s & 7) + n + 7) >> 3;
        const int64_t c_lo = (s - start) >> 3;
        for (int64_t k = 0; k < nw; ++k) {
            const int64_t c = c_lo + k;
            const int slot = cnt[(size_t)c]++;
            const int64_t idx = ((grp_off[c / LN] + slot / ISB_COLS_UNIT) * LN + (c % LN)) * ISB_COLS_UNIT + slot % ISB_COLS_UNIT;
            words[idx] = words_in[seg_word[i] + k];
            ids[idx] = seg_pair[i];
        }
    }
    return n_chunks;
}

// ---- read-major segments -> reference-delta transfer format (host side of isb_profile_reads_delta) -------------------
// Per 8-base unit the event bits; per passing base whose one-hot code differs from the reference's (reference not
// A/C/T/G: code 0) one (word index in the canonical stream, nibble position | XOR of the codes) entry.  Entries are
// produced in stream order.  Returns the number of entries (only the first cap_mis are stored: call again with larger
// buffers when it exceeds cap_mis), or -1 when the segments violate the layout rules / n_units does not match.
extern "C" int64_t isb_reads_delta_host(int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len,
                                        const int64_t *seg_word, const uint32_t *words_in, int64_t n_words_in, int32_t start,
                                        int32_t L, const uint8_t *ref, uint8_t *pass, int64_t n_units, uint32_t *mis_word,
                                        uint8_t *mis_code, int64_t cap_mis)
{
    if ((start & 7) || L < 0 || n_segs < 0 || !ref || (n_units > 0 && !pass)) return -1;
    int64_t u = 0, n_mis = 0;
    for (int64_t i = 0; i < n_segs; ++i) {
        const int64_t s = seg_start[i], n = seg_len[i];
        const int64_t nw = ((s & 7) + n + 7) >> 3;
        if (n < 1 || n > 256 || s < start || s + n > (int64_t)start + L || seg_word[i] < 0 || seg_word[i] + nw > n_words_in ||
            (i > 0 && seg_start[i - 1] > s) || u + nw > n_units)
            return -1;
        const int64_t r0 = (s & ~(int64_t)7) - start;                  // reference index of nibble 0 of the first unit
        for (int64_t k = 0; k < nw; ++k, ++u) {
            const uint32_t w = words_in[seg_word[i] + k];
            uint32_t ps = 0;
            for (int t = 0; t < 8; ++t) {
                const uint32_t nib = (w >> (4 * t)) & 15u;
                if (!nib) continue;
                ps |= 1u << t;
                const int64_t r = r0 + 8 * k + t;
                const uint32_t rc = (r >= 0 && r < L) ? ref[r] : 4u;
                const uint32_t hot = rc < 4u ? (1u << rc) : 0u;
                if (nib != hot) {
                    const int64_t cw = 1 + u + i;                      // canonical stream: leading zero word + one separator per segment
                    if (cw > 0xffffffffll) return -1;
                    if (n_mis < cap_mis && mis_word && mis_code) {
                        mis_word[n_mis] = (uint32_t)cw;
                        mis_code[n_mis] = (uint8_t)((nib ^ hot) | ((uint32_t)t << 4));
                    }
                    ++n_mis;
                }
            }
            pass[u] = (uint8_t)ps;
        }
    }
    return u == n_units ? n_mis : -1;
}
