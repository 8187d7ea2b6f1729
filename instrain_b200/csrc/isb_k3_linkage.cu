// isb_k3_linkage.cu -- K3: pairwise SNV linkage (r^2, D') from allele co-occurrence on shared read pairs, sm_100a.
//
// Replaces update_linked_reads (inStrain/profile/linkage.py:254-283), calc_mm_SNV_linkage_network (:14-44),
// calculate_ld / _iterator_ld_sites (:46-131), major_minor_allele (:133-136) and the deterministic part of
// _calc_ld_single (:138-198).  Integer popcount work, no tensor cores.
//
// Formulation.  For a linkage-eligible site s (anySNP, allele set `bases`) and allele b, let n_s,b(r) in {0,1,2} be
// the number of qualifying events of read pair r showing b at s.  The reference's combo count of an edge is
//      K[b1,b2](s,t; mm<=m) = sum_{r: mm(r)<=m} n_s,b1(r) * n_t,b2(r)              (s < t, same split)
// because read_to_snvs[mm][name] lists every (site, base) entry of the pair and itertools.combinations pairs them.
// Pair ids are assigned in BAM order, so the pairs covering a site fall in a narrow id window; per site we keep
// bit rows over that window only:   any | ge1[b] for b in bases | ge2[b] for b in bases   (n>=1, n>=2 planes).
// Then K = popc(ge1 & ge1 & mask_m) (+ the ge2 cross terms when a site has a double entry), where mask_m marks the
// pairs with mm <= m.  A row is emitted for level m only if some pair with mm == m links the two sites
// (sorted(mm2combo2counts.items()), linkage.py:93).  Self edges (a pair entered twice on ONE site, which htslib's
// overlap quirk can produce) are handled by a slow exact path.
#include "isb_common.cuh"
#include "isb_scan.cuh"
#include "isb_k3_dev.cuh"
#include "isb_k2_site.cuh"
#include <math_constants.h>
#include <cstdlib>

#define K3_THREADS 256
#define K3_ROW_SCRATCH ISB_K3_ROW_SLOT // words of per-warp shared scratch for assembling a site's bit rows (= the fixed row slot)

struct k3_args {
    // events
    int64_t n;
    const int32_t *ref_pos;
    const uint8_t *base;
    const uint8_t *qual;
    const int32_t *read_id;
    int64_t n_pairs;
    const uint8_t *pair_mm;
    int32_t start, L;
    int32_t col_shift;             // column words: position of `start` relative to the batch the column lists index (chunked calls)
    int M, min_qual, min_snp;
    const int32_t *counts;
    const unsigned long long *nmask;
    const uint8_t *site_flags;
    int32_t n_splits;
    const int32_t *splits;
    // sites
    int64_t S;
    const int32_t *site_pos;       // relative position index
    int64_t *site_ev;              // [2S] event range of the site
    isb_site_meta *meta;
    int32_t *site_words;
    const int64_t *row_off;
    uint32_t *rows;
    uint8_t *has2;
    const uint32_t *mle;           // [M][nwp] pairs with mm <= m (M > 1 only)
    int64_t nwp;
    const int64_t *tile_off;       // event offsets of position tiles of tile_tp positions (bounds the per-site search)
    int tile_tp;
    int32_t *pair_i, *pair_j;      // linked site pairs (site indices), filled by k3_enum_pairs
    int64_t pair_cap;
    // per-tile site slots (fused read-major path, isb_k1f_fused.cu): sites are not globally ordered, a tile's are
    int n_tiles;
    const int32_t *tile_first, *tile_cnt;
    int64_t sites_cap;
    const int4 *site_counts;       // counts per site slot (M = 1) instead of counts[p]
    const unsigned long long *n_sites_dev;   // device-side site / listed-pair counts (no host round trip)
    uint64_t seed;                 // key of the re-drawn ("normalized") linkage columns
    // output
    isb_ld_row *out;
    int64_t cap;
    unsigned long long *n_ld;
    unsigned long long *n_site_pairs;
    unsigned int *d_err;
};

struct FlagFn {
    const uint8_t *flags;
    __device__ int operator()(int64_t i) const { return (flags[i] & ISB_SITE_ANYSNP) ? 1 : 0; }
};
struct SitePosSink {
    int32_t *site_pos;
    __device__ void operator()(int64_t i, int64_t prefix, int v) const { if (v) site_pos[prefix] = (int32_t)i; }
};
struct WordsFn {
    const int32_t *w;
    __device__ int operator()(int64_t i) const { return w[i]; }
};
struct RowOffSink {
    int64_t *row_off;
    __device__ void operator()(int64_t i, int64_t prefix, int) const { row_off[i] = prefix; }
};

__device__ __forceinline__ bool k3_qualifies(const k3_args &a, int64_t e, unsigned bases)
{
    const int b = a.base[e];
    return a.qual[e] >= a.min_qual && b < 4 && ((bases >> b) & 1u);
}

__device__ __forceinline__ int k3_site_split(const k3_args &a, int64_t abs_pos)
{
    return isb_site_split(a.splits, a.n_splits, abs_pos);
}

// ---- per-site event range, split --------------------------------------------------------------------------------
// One THREAD per site: the two lower bounds are scalar binary searches bounded by the site's position tile
// (tile_off from a k1_tile_offsets launch: ~1e5 events => 17 steps).  The searches are pure latency (dependent
// scattered loads), so what matters is how many are in flight: the first version ran them one site per WARP over the
// whole 1e10-event column (34 steps, 1 site in flight per warp) and took 180 us for 1e5 sites.
__global__ void __launch_bounds__(256) k3_site_ranges(k3_args a)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.S) return;
    const int32_t p = a.site_pos[k];
    const int64_t abs_pos = (int64_t)p + a.start;
    const int t = p / a.tile_tp;
    const int64_t t_lo = a.tile_off[t], t_hi = a.tile_off[t + 1];
    const int64_t lo = isb_lower_bound(a.ref_pos, t_lo, t_hi, abs_pos);
    const int64_t hi = isb_lower_bound(a.ref_pos, lo, t_hi, abs_pos + 1);
    a.site_ev[2 * k] = lo;
    a.site_ev[2 * k + 1] = hi;
    a.meta[k].split = k3_site_split(a, abs_pos);
}

// ---- read-major front end: the events of the eligible sites, materialised from the aligned segments ---------------
// K3 only ever looks at the events of linkage-eligible sites (~1 % of the positions), so with a read-major batch the
// (base, pair id) lists of exactly those sites are gathered from the segments that cover them into compact scratch
// columns, and the rest of K3 runs unchanged on those.  One thread per site finds its candidate segments
// (seg_start in (p - max_seg_len, p]) inside the site's K1r tile range; a scan sizes the scratch; one warp per site
// then tests the candidates and writes the passing ones.
__global__ void __launch_bounds__(256) k3r_site_cand(k3_args a, isb_reads_dev rd, int64_t *__restrict__ cand_lo,
                                                     int32_t *__restrict__ n_cand)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.S) return;
    const int32_t p = a.site_pos[k];
    const int64_t abs_pos = (int64_t)p + a.start;
    const int t = p / K1R_TILE;
    const int64_t t_lo = rd.tile_lo[t], t_hi = rd.tile_hi[t];
    const int64_t lo = isb_lower_bound(rd.seg_start, t_lo, t_hi, abs_pos - rd.max_seg_len + 1);
    const int64_t hi = isb_lower_bound(rd.seg_start, lo, t_hi, abs_pos + 1);
    cand_lo[k] = lo;
    n_cand[k] = (int32_t)(hi - lo);
    a.meta[k].split = k3_site_split(a, abs_pos);
}

// ---- column-word front end (isb_cols_batch): the entries of a site are ONE nibble of every word of its column list ----
// Site p lives in column word c = p / 8, lane c % 8 of group c / 8; slot i of its list is word
// ((grp_off[g] + i / 8) * 8 + lane) * 8 + i % 8 of `words` / `ids` (table order of the segments = column order).  No
// candidate search and no dependent address chain: every load address follows from p alone.
struct k3c_column {
    int64_t base;   // index of slot 0
    int depth;      // slots (incl. padding)
    int sh;         // bit offset of the site's nibble
};

__device__ __forceinline__ k3c_column k3c_site_column(const isb_cols_dev &cd, int32_t p)
{
    k3c_column col;
    const int64_t c = p >> 3, g = c / ISB_COLS_LANES;
    const int64_t c0 = __ldg(cd.grp_off + g), c1 = __ldg(cd.grp_off + g + 1);
    const bool ok = c0 >= 0 && c1 >= c0 && c1 <= cd.n_chunks && c1 - c0 <= (1 << 24);   // K1c has flagged a violation
    col.base = (c0 * ISB_COLS_LANES + (c % ISB_COLS_LANES)) * ISB_COLS_UNIT;
    col.depth = ok ? (int)(c1 - c0) * ISB_COLS_UNIT : 0;
    col.sh = (p & 7) << 2;
    return col;
}

__device__ __forceinline__ int64_t k3c_slot_index(const k3c_column &col, int i)
{
    return col.base + (int64_t)(i / ISB_COLS_UNIT) * ISB_COLS_CHUNK + (i % ISB_COLS_UNIT);
}

// entry of slot i from an already loaded (word, id) pair
__device__ __forceinline__ bool k3c_decode(const k3c_column &col, uint32_t w, int id_in, int64_t n_pairs, int &b, int &id)
{
    const uint32_t code = (w >> col.sh) & 15u;
    if (!code || id_in < 0 || (int64_t)id_in >= n_pairs) return false;   // no event / padding / invalid id
    id = id_in;
    b = __ffs((int)code) - 1;                                         // one-hot A,C,T,G
    return true;
}

__device__ __forceinline__ bool k3c_candidate(const isb_cols_dev &cd, const k3c_column &col, int i, int64_t n_pairs, int &b,
                                              int &id)
{
    const int64_t idx = k3c_slot_index(col, i);
    const uint32_t w = __ldg(cd.words + idx);
    if (!((w >> col.sh) & 15u)) return false;
    return k3c_decode(col, w, __ldg(cd.ids + idx), n_pairs, b, id);
}

// Everything the warp-per-site kernel needs to know about a site before it can request the site's words, computed by
// one THREAD per site (site_pos -> flags / group offsets / split are dependent loads: here every site's chain is in
// flight at once, the warp kernel then starts with ONE 32-byte load).
struct __align__(16) k3c_site_rec {
    int64_t base;        // index of slot 0 of the site's column list
    int32_t depth;       // slots (incl. padding)
    int32_t p;           // relative position
    int32_t split;       // split index, -1 if the position is in no split
    uint32_t bases_sh;   // bits 0-3: the site's allele set, bits 8-15: bit offset of its nibble
    int32_t pad[2];
};

__global__ void __launch_bounds__(256) k3c_site_prep(k3_args a, isb_cols_dev cd, k3c_site_rec *__restrict__ recs)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.S) return;
    const int32_t p = a.site_pos[k];
    const k3c_column col = k3c_site_column(cd, p + a.col_shift);
    k3c_site_rec r;
    r.base = col.base;
    r.depth = col.depth;
    r.p = p;
    r.split = k3_site_split(a, (int64_t)p + a.start);
    r.bases_sh = (uint32_t)(a.site_flags[p] & 0xF) | ((uint32_t)col.sh << 8);
    r.pad[0] = r.pad[1] = 0;
    recs[k] = r;
    a.meta[k].split = r.split;
}

// Fused read-major front end (warp per site): gather the site's qualifying (pair id, base) entries from its candidate
// segments (or, column words: from its column list) into a per-warp shared-memory list, derive the pair-id window,
// assemble the bit rows (shared-memory scratch for the common small windows) and store them in the site's fixed row
// slot.  The events never touch global memory; sites with more qualifying entries than the list holds gather a second
// time.
#define K3R_EV_CAP 320
template <bool kCols>
__global__ void __launch_bounds__(K3_THREADS) k3r_site_rows(k3_args a, isb_reads_dev rd, isb_cols_dev cd,
                                                            const k3c_site_rec *__restrict__ recs,
                                                            const int64_t *__restrict__ cand_lo,
                                                            const int32_t *__restrict__ n_cand, int64_t *__restrict__ row_off,
                                                            unsigned long long *__restrict__ row_words_total, int64_t row_cap)
{
    __shared__ uint32_t s_rows[K3_THREADS / 32][K3_ROW_SCRATCH];
    __shared__ int32_t s_id[K3_THREADS / 32][K3R_EV_CAP];
    __shared__ uint8_t s_b[K3_THREADS / 32][K3R_EV_CAP];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    for (int64_t k = warp0; k < a.S; k += n_warps) {
        k3c_column col = {0, 0, 0};
        int32_t p, split;
        unsigned bases;
        if (kCols) {                                                   // one 32-byte record (k3c_site_prep)
            const int4 r0 = __ldg(reinterpret_cast<const int4 *>(recs + k));
            const int4 r1 = __ldg(reinterpret_cast<const int4 *>(recs + k) + 1);
            col.base = ((int64_t)(uint32_t)r0.x) | ((int64_t)r0.y << 32);
            col.depth = r0.z;
            p = r0.w;
            split = r1.x;
            bases = (unsigned)r1.y & 0xFu;
            col.sh = ((unsigned)r1.y >> 8) & 0xFFu;
        } else {
            p = a.site_pos[k];
            split = a.meta[k].split;                                   // set by k3r_site_cand
            bases = a.site_flags[p] & 0xF;
        }
        const int64_t abs_pos = (int64_t)p + a.start;
        const int na = __popc(bases);
        const int64_t clo = kCols ? 0 : cand_lo[k];
        const int nc = kCols ? col.depth : n_cand[k];
        auto candidate = [&](int i, int &b, int &id) -> bool {
            return kCols ? k3c_candidate(cd, col, i, a.n_pairs, b, id) : k3r_candidate(rd, clo + i, abs_pos, b, id);
        };
        int cnt = 0, idmin = INT_MAX, idmax = -1;
        auto append = [&](bool ok, int b, int id) {
            const unsigned mask = __ballot_sync(ISB_FULL, ok);
            if (ok) {
                const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                if (slot < K3R_EV_CAP) { s_id[wib][slot] = id; s_b[wib][slot] = (uint8_t)b; }
                idmin = min(idmin, id);
                idmax = max(idmax, id);
            }
            cnt += __popc(mask);
        };
        if (kCols) {
            // every address follows from the position: the words and ids of 4 x 32 slots are requested before the first is
            // used (one memory latency per 128 slots instead of two dependent ones per 32)
            for (int i0 = 0; i0 < nc; i0 += 128) {
                uint32_t w[4];
                int idv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    w[u] = 0u;
                    idv[u] = -1;
                    if (i < nc) {
                        const int64_t idx = k3c_slot_index(col, i);
                        w[u] = __ldg(cd.words + idx);
                        idv[u] = __ldg(cd.ids + idx);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (i0 + u * 32 >= nc) break;                      // warp-uniform
                    int b = 0, id = 0;
                    const bool ok = k3c_decode(col, w[u], idv[u], a.n_pairs, b, id) && ((bases >> b) & 1u);
                    append(ok, b, id);
                }
            }
        } else {
            for (int i0 = 0; i0 < nc; i0 += 32) {
                int b = 0, id = 0;
                bool ok = (i0 + lane < nc) && candidate(i0 + lane, b, id);
                ok = ok && ((bases >> b) & 1u);
                append(ok, b, id);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            idmin = min(idmin, __shfl_xor_sync(ISB_FULL, idmin, d));
            idmax = max(idmax, __shfl_xor_sync(ISB_FULL, idmax, d));
        }
        isb_site_meta m;
        m.split = split;
        m.ev_lo_rel = 0;
        m.wlo = idmax >= 0 ? (idmin >> 5) : 0;
        m.nw = idmax >= 0 ? (idmax >> 5) - (idmin >> 5) + 1 : 0;
        const int n_words = (1 + 2 * na) * m.nw;
        // Row storage: site k owns the fixed slot [k * K3_ROW_SCRATCH, + K3_ROW_SCRATCH) -- no allocation, nothing to
        // wait for; only rows wider than a slot (deep coverage, many alleles) are allocated with an atomic behind the
        // S fixed slots.
        unsigned long long off = (unsigned long long)k * K3_ROW_SCRATCH;
        if (n_words > K3_ROW_SCRATCH) {
            if (lane == 0) off = (unsigned long long)a.S * K3_ROW_SCRATCH + atomicAdd(row_words_total, (unsigned long long)n_words);
            off = __shfl_sync(ISB_FULL, off, 0);
        }
        const bool fits = (int64_t)(off + n_words) <= row_cap;
        if (!fits) {                                                   // host grows the row storage and re-runs
            if (lane == 0) atomicOr(a.d_err, ISB_DEV_ERR_ROWBUF);
            m.nw = 0;
        }
        if (lane == 0) {
            a.meta[k] = m;
            row_off[k] = (int64_t)off;
            a.has2[k] = 0;
        }
        if (m.nw == 0) continue;
        const bool in_smem = n_words <= K3_ROW_SCRATCH;
        uint32_t *g_any = a.rows + off;
        uint32_t *any = in_smem ? s_rows[wib] : g_any;
        for (int i = lane; i < n_words; i += 32) any[i] = 0u;
        __syncwarp();
        bool dbl = false;
        if (cnt <= K3R_EV_CAP) {
            for (int e = lane; e < cnt; e += 32) dbl |= k3_row_set(any, m, na, bases, s_b[wib][e], s_id[wib][e], a.d_err);
        } else {
            for (int i0 = 0; i0 < nc; i0 += 32) {
                int b = 0, id = 0;
                if (i0 + lane < nc && candidate(i0 + lane, b, id) && ((bases >> b) & 1u))
                    dbl |= k3_row_set(any, m, na, bases, b, id, a.d_err);
            }
        }
        if (__any_sync(ISB_FULL, dbl) && lane == 0) a.has2[k] = 1;
        __syncwarp();
        if (in_smem) {
            for (int i = lane; i < n_words; i += 32) g_any[i] = any[i];
            __syncwarp();
        }
    }
}

// ---- per-site pair-id window (warp per site, coalesced) -----------------------------------------------------------
__global__ void __launch_bounds__(K3_THREADS) k3_site_windows(k3_args a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    for (int64_t k = warp0; k < a.S; k += n_warps) {
        const int32_t p = a.site_pos[k];
        const int64_t lo = a.site_ev[2 * k], hi = a.site_ev[2 * k + 1];
        const unsigned bases = a.site_flags[p] & 0xF;
        int idmin = INT_MAX, idmax = -1;
        for (int64_t e = lo + lane; e < hi; e += 32)
            if (k3_qualifies(a, e, bases)) {
                const int id = a.read_id[e];
                idmin = min(idmin, id);
                idmax = max(idmax, id);
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            idmin = min(idmin, __shfl_xor_sync(ISB_FULL, idmin, d));
            idmax = max(idmax, __shfl_xor_sync(ISB_FULL, idmax, d));
        }
        if (lane == 0) {
            isb_site_meta m = a.meta[k];
            m.ev_lo_rel = 0;
            m.wlo = idmax >= 0 ? (idmin >> 5) : 0;
            m.nw = idmax >= 0 ? (idmax >> 5) - (idmin >> 5) + 1 : 0;
            a.meta[k] = m;
            a.site_words[k] = (1 + 2 * __popc(bases)) * m.nw;
            a.has2[k] = 0;
        }
    }
}

// ---- bit rows ---------------------------------------------------------------------------------------------------
// Warp per site.  Rows of a site that fit in the warp's 64-word shared-memory scratch (the common case: 1 + 2*|bases|
// rows x <= 12 words) are assembled there with shared-memory atomics and stored once, coalesced; larger windows fall
// back to global atomics on the (pre-zeroed) row storage.
__global__ void __launch_bounds__(K3_THREADS) k3_build_rows(k3_args a)
{
    __shared__ uint32_t s_rows[K3_THREADS / 32][K3_ROW_SCRATCH];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    for (int64_t k = warp0; k < a.S; k += n_warps) {
        const isb_site_meta m = a.meta[k];
        if (m.nw == 0) continue;
        const int32_t p = a.site_pos[k];
        const unsigned bases = a.site_flags[p] & 0xF;
        const int na = __popc(bases);
        const int n_words = (1 + 2 * na) * m.nw;
        const bool in_smem = n_words <= K3_ROW_SCRATCH;
        uint32_t *g_any = a.rows + a.row_off[k];
        uint32_t *any = in_smem ? s_rows[wib] : g_any;
        if (in_smem) {
            for (int i = lane; i < n_words; i += 32) any[i] = 0u;
            __syncwarp();
        }
        const int64_t lo = a.site_ev[2 * k], hi = a.site_ev[2 * k + 1];
        bool dbl = false;
        for (int64_t e = lo + lane; e < hi; e += 32) {
            if (!k3_qualifies(a, e, bases)) continue;
            const int b = a.base[e];
            const int id = a.read_id[e];
            const int w = (id >> 5) - m.wlo;
            const uint32_t bit = 1u << (id & 31);
            const int r = __popc(bases & ((1u << b) - 1u));
            if (atomicOr(any + w, bit) & bit) dbl = true;             // second entry of this pair on this site
            uint32_t *ge1 = any + (size_t)(1 + r) * m.nw;
            if (atomicOr(ge1 + w, bit) & bit) {
                uint32_t *ge2 = any + (size_t)(1 + na + r) * m.nw;
                if (atomicOr(ge2 + w, bit) & bit) atomicOr(a.d_err, ISB_DEV_ERR_MULT);
            }
        }
        if (__any_sync(ISB_FULL, dbl) && lane == 0) a.has2[k] = 1;
        if (in_smem) {
            __syncwarp();
            for (int i = lane; i < n_words; i += 32) g_any[i] = any[i];
            __syncwarp();
        }
    }
}

// mle[m][w] = bit i set iff pair (32 w + i) has mm <= m
__global__ void __launch_bounds__(256) k3_mm_masks(const uint8_t *__restrict__ pair_mm, int64_t n_pairs, int M,
                                                   int64_t nwp, uint32_t *__restrict__ mle)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwp) return;
    for (int m = 0; m < M; ++m) {
        uint32_t word = 0;
        for (int i = 0; i < 32; ++i) {
            const int64_t id = (w << 5) + i;
            if (id < n_pairs && pair_mm[id] <= m) word |= 1u << i;
        }
        mle[(size_t)m * nwp + w] = word;
    }
}

// ---- LD statistics ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void k3_major_minor(const int *c, int &maj, int &mnr)
{
    // sorted(d, key=d.get, reverse=True)[:2]: stable, ties keep A,C,T,G order (linkage.py:133-136)
    int idx[4] = {0, 1, 2, 3};
#pragma unroll
    for (int x = 1; x < 4; ++x) {
        const int v = idx[x];
        int k = x - 1;
        while (k >= 0 && c[idx[k]] < c[v]) { idx[k + 1] = idx[k]; --k; }
        idx[k + 1] = v;
    }
    maj = idx[0];
    mnr = idx[1];
}

__device__ __forceinline__ void k3_emit(const k3_args &a, int32_t p1, int32_t p2, int m, int A, int al, int B, int bl,
                                        int cAB, int cAb, int caB, int cab)
{
    const int total = cAB + cAb + caB + cab;
    if (!(total > a.min_snp)) return;                                  // strict (linkage.py:165)
    const double tot = (double)total;
    const double fAB = __ddiv_rn((double)cAB, tot), fAb = __ddiv_rn((double)cAb, tot);
    const double faB = __ddiv_rn((double)caB, tot), fab = __ddiv_rn((double)cab, tot);
    const double fA = __dadd_rn(fAB, fAb), fa = __dadd_rn(fab, faB);
    const double fB = __dadd_rn(fAB, faB), fb = __dadd_rn(fab, fAb);
    const double linkD = __dsub_rn(fAB, __dmul_rn(fA, fB));
    double r2 = CUDART_NAN, dp = CUDART_NAN;
    if (!(fa == 0.0 || fA == 0.0 || fB == 0.0 || fb == 0.0))
        r2 = __ddiv_rn(__dmul_rn(linkD, linkD), __dmul_rn(__dmul_rn(__dmul_rn(fA, fa), fB), fb));
    const double linkd = __dsub_rn(fab, __dmul_rn(fa, fb));
    if (linkd < 0.0) {
        const double d1 = __dmul_rn(-fA, fB), d2 = __dmul_rn(-fa, fb);
        dp = __ddiv_rn(linkd, d1 > d2 ? d1 : d2);
    } else if (linkD > 0.0) {
        const double d1 = __dmul_rn(fA, fb), d2 = __dmul_rn(fa, fB);
        dp = __ddiv_rn(linkd, d1 < d2 ? d1 : d2);
    }
    // r2_normalized / d_prime_normalized (linkage.py:200-228): the same statistics on min_snp haplotypes re-drawn from
    // the four observed frequencies (counter-based draws keyed by the row, isb_k2_site.cuh)
    double r2n = CUDART_NAN, dpn = CUDART_NAN;
    if (a.min_snp >= 1) {
        const uint64_t key = ((uint64_t)(uint32_t)(p1 + a.start) << 32) | (uint64_t)(uint32_t)(p2 + a.start);
        const int hc[4] = {cAB, cAb, caB, cab};                       // np.random.choice(['AB','Ab','aB','ab'], p=..., size=min_snp)
        int hn[4];
        isb_redraw4(hc, total, a.min_snp, isb_rng_base(a.seed, ISB_RNG_TAG_LD, key, (uint64_t)m), hn);
        const int n1 = hn[0], n2 = hn[0] + hn[1], n3 = hn[0] + hn[1] + hn[2];
        const double ns = (double)a.min_snp;
        const double gAB = __ddiv_rn((double)n1, ns), gAb = __ddiv_rn((double)(n2 - n1), ns);
        const double gaB = __ddiv_rn((double)(n3 - n2), ns), gab = __ddiv_rn((double)(a.min_snp - n3), ns);
        const double gA = __dadd_rn(gAB, gAb), ga = __dadd_rn(gab, gaB), gB = __dadd_rn(gAB, gaB), gb = __dadd_rn(gab, gAb);
        const double ldn = __dsub_rn(gab, __dmul_rn(ga, gb));
        if (!(ga == 0.0 || gA == 0.0 || gB == 0.0 || gb == 0.0))
            r2n = __ddiv_rn(__dmul_rn(ldn, ldn), __dmul_rn(__dmul_rn(__dmul_rn(gA, ga), gB), gb));
        if (ldn < 0.0) {
            const double d1 = __dmul_rn(-gA, gB), d2 = __dmul_rn(-ga, gb);
            dpn = __ddiv_rn(ldn, d1 > d2 ? d1 : d2);
        } else if (ldn > 0.0) {
            const double d1 = __dmul_rn(gA, gb), d2 = __dmul_rn(ga, gB);
            dpn = __ddiv_rn(ldn, d1 < d2 ? d1 : d2);
        }
    }
    const unsigned long long slot = atomicAdd(a.n_ld, 1ull);
    if ((int64_t)slot < a.cap) {
        isb_ld_row r;
        r.pos_a = p1 + a.start; r.pos_b = p2 + a.start; r.mm = m;
        r.c_AB = cAB; r.c_Ab = cAb; r.c_aB = caB; r.c_ab = cab;
        r.allele_A = (uint8_t)A; r.allele_a = (uint8_t)al; r.allele_B = (uint8_t)B; r.allele_b = (uint8_t)bl;
        r.r2 = r2; r.d_prime = dp;
        r.r2_normalized = r2n; r.d_prime_normalized = dpn;
        a.out[slot] = r;
    }
}

struct k3_site {
    int32_t p;
    int wlo, nw, na, has2;
    unsigned bases;
    const uint32_t *rows;   // `any` row; ge1[r] at rows + (1+r)*nw; ge2[r] at rows + (1+na+r)*nw
};

__device__ __forceinline__ k3_site k3_load_site(const k3_args &a, int64_t k)
{
    k3_site s;
    const isb_site_meta m = a.meta[k];
    s.p = a.site_pos[k];
    s.wlo = m.wlo; s.nw = m.nw;
    s.bases = a.site_flags[s.p] & 0xF;
    s.na = __popc(s.bases);
    s.has2 = a.has2[k];
    s.rows = a.rows + a.row_off[k];
    return s;
}

__device__ __forceinline__ int k3_rank(unsigned bases, int b)
{
    return ((bases >> b) & 1u) ? __popc(bases & ((1u << b) - 1u)) : -1;
}

// K[b1,b2] over absolute pair-id words [lo, hi), optionally masked (mask indexed by absolute word)
__device__ __forceinline__ int k3_pair_count(const k3_site &si, int ra, const k3_site &sj, int rb, int lo, int hi,
                                             const uint32_t *__restrict__ mask)
{
    if (ra < 0 || rb < 0) return 0;
    const uint32_t *a1 = si.rows + (size_t)(1 + ra) * si.nw - si.wlo;
    const uint32_t *b1 = sj.rows + (size_t)(1 + rb) * sj.nw - sj.wlo;
    int c = 0;
    if (!(si.has2 | sj.has2)) {
        for (int w = lo; w < hi; ++w) {
            uint32_t x = a1[w] & b1[w];
            if (mask) x &= __ldg(mask + w);
            c += __popc(x);
        }
    } else {
        const uint32_t *a2 = si.rows + (size_t)(1 + si.na + ra) * si.nw - si.wlo;
        const uint32_t *b2 = sj.rows + (size_t)(1 + sj.na + rb) * sj.nw - sj.wlo;
        for (int w = lo; w < hi; ++w) {
            const uint32_t mk = mask ? __ldg(mask + w) : 0xffffffffu;
            const uint32_t x1 = a1[w] & mk, x2 = a2[w] & mk, y1 = b1[w], y2 = b2[w];
            c += __popc(x1 & y1) + __popc(x2 & y1) + __popc(x1 & y2) + __popc(x2 & y2);   // (g1+g2)*(g1+g2)
        }
    }
    return c;
}

// exact-mm counts of a site at level m: per site slot (fused read-major path, M = 1) or from the dense counts array
__device__ __forceinline__ int4 k3_site_counts(const k3_args &a, int64_t k, int32_t p, int m)
{
    if (a.site_counts) return __ldg(a.site_counts + k);
    return __ldg(reinterpret_cast<const int4 *>(a.counts) + (size_t)p * a.M + m);
}

__device__ __forceinline__ bool k3_level_present(const k3_args &a, int32_t p, int m, const int4 &E)
{
    if (E.x + E.y + E.z + E.w > 0) return true;
    return a.nmask && ((a.nmask[p] >> m) & 1ull);
}

// Linked site pairs are first ENUMERATED (warp per site, lanes over the partner sites of the same split: window overlap
// + one AND over the `any` rows), then evaluated one THREAD per pair.  The evaluation is a chain of dependent global
// loads (rows, counts, masks); with one pair per lane of a site-warp only ~3.7 linked pairs per site kept the machine
// busy (300 us for 3.7e5 pairs), one thread per pair puts every pair in flight at once.
__global__ void __launch_bounds__(K3_THREADS) k3_enum_pairs(k3_args a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * K3_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * K3_THREADS) >> 5;
    for (int64_t k = warp0; k < a.S; k += n_warps) {
        const isb_site_meta mi = a.meta[k];
        if (mi.split < 0 || mi.nw == 0) continue;
        const uint32_t *any_i = a.rows + a.row_off[k] - mi.wlo;
        for (int64_t jb = k + 1;; jb += 32) {
            const int64_t j = jb + lane;
            isb_site_meta mj;
            mj.ev_lo_rel = 0; mj.nw = 0; mj.wlo = 0; mj.split = -2;
            if (j < a.S) mj = a.meta[j];
            const bool cand = j < a.S && mj.split == mi.split;
            if (!__any_sync(ISB_FULL, cand)) break;
            bool linked = false;
            if (cand && mj.nw > 0) {
                const int lo = max(mi.wlo, mj.wlo), hi = min(mi.wlo + mi.nw, mj.wlo + mj.nw);
                if (lo < hi) {
                    const uint32_t *any_j = a.rows + a.row_off[j] - mj.wlo;
                    for (int w = lo; w < hi; ++w)
                        if (any_i[w] & any_j[w]) { linked = true; break; }
                }
            }
            const unsigned m = __ballot_sync(ISB_FULL, linked);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.n_site_pairs, (unsigned long long)__popc(m));
                base = __shfl_sync(ISB_FULL, base, 0);
                if (linked) {
                    const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
                    if ((int64_t)slot < a.pair_cap) { a.pair_i[slot] = (int32_t)k; a.pair_j[slot] = (int32_t)j; }
                }
            }
        }
    }
}

// Few-threads-per-site enumeration (the default; ISB_K3_ENUM=0 selects the warp-per-site kernel above): every site's
// partner scan is in flight at once, split over K3_ENUM_LANES threads.  The warp-per-site kernel keeps ONE site per warp in flight and is bound by the latency of
// its dependent loads (meta -> row offset -> `any` words): 71 us per 1e5 sites against ~45 us here (B200, K3 stage
// 0.62 -> 0.57 ms per 2e7 positions).  Slots are allocated with one atomic per group of lanes that found a link in the
// same trip.
#ifndef K3_ENUM_LANES
#define K3_ENUM_LANES 4              // threads per site: lane s of a site takes its partners k + 1 + s, k + 1 + s + LANES, ...
#endif
__global__ void __launch_bounds__(256) k3_enum_pairs_t(k3_args a)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = t / K3_ENUM_LANES;
    if (k >= a.S) return;
    const isb_site_meta mi = a.meta[k];
    if (mi.split < 0 || mi.nw == 0) return;
    const uint32_t *any_i = a.rows + a.row_off[k] - mi.wlo;
    const int i_hi = mi.wlo + mi.nw;
    for (int64_t j = k + 1 + (t % K3_ENUM_LANES); j < a.S; j += K3_ENUM_LANES) {
        const isb_site_meta mj = a.meta[j];
        if (mj.split != mi.split) break;
        const int lo = max(mi.wlo, mj.wlo), hi = min(i_hi, mj.wlo + mj.nw);
        bool linked = false;
        if (lo < hi) {
            const uint32_t *any_j = a.rows + a.row_off[j] - mj.wlo;
            for (int w = lo; w < hi; ++w)
                if (any_i[w] & any_j[w]) { linked = true; break; }
        }
        if (linked) {
            const unsigned act = __activemask();
            const int leader = __ffs(act) - 1, lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(a.n_site_pairs, (unsigned long long)__popc(act));
            base = __shfl_sync(act, base, leader);
            const unsigned long long slot = base + __popc(act & ((1u << lane) - 1u));
            if ((int64_t)slot < a.pair_cap) { a.pair_i[slot] = (int32_t)k; a.pair_j[slot] = (int32_t)j; }
        }
    }
}

#ifndef K3_STATS_MINB
#define K3_STATS_MINB 4              // 64 registers: 4 blocks of 256 threads per SM (measured: K3 stage 0.394 -> 0.367 ms per 2e7 positions)
#endif
__device__ __forceinline__ void k3_pair_stats_one(const k3_args &a, int64_t t);

__global__ void __launch_bounds__(256, K3_STATS_MINB) k3_pair_stats(k3_args a, int64_t n_pairs_listed)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs_listed) return;
    k3_pair_stats_one(a, t);
}

// the same over a device-side pair count (grid-stride): no host round trip between enumeration and statistics
__global__ void __launch_bounds__(256, K3_STATS_MINB) k3_pair_stats_dev(k3_args a)
{
    const unsigned long long listed = *a.n_site_pairs;
    const int64_t n = (int64_t)(listed < (unsigned long long)a.pair_cap ? listed : (unsigned long long)a.pair_cap);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) k3_pair_stats_one(a, t);
}

__device__ __forceinline__ void k3_pair_stats_one(const k3_args &a, int64_t t)
{
    const int64_t ki = a.pair_i[t], kj = a.pair_j[t];
    const k3_site si = k3_load_site(a, ki), sj = k3_load_site(a, kj);
    const int lo = max(si.wlo, sj.wlo), hi = min(si.wlo + si.nw, sj.wlo + sj.nw);
    unsigned long long pm = 1ull;                                      // mm levels having >= 1 linking pair
    if (a.M > 1) {
        const uint32_t *any_i = si.rows - si.wlo, *any_j = sj.rows - sj.wlo;
        pm = 0;
        for (int w = lo; w < hi; ++w) {
            uint32_t x = any_i[w] & any_j[w];
            while (x) {
                const int b = __ffs(x) - 1;
                x &= x - 1;
                pm |= 1ull << __ldg(a.pair_mm + (((int64_t)w << 5) + b));
            }
        }
        if (!pm) return;
    }
    int C1[4] = {0, 0, 0, 0}, C2[4] = {0, 0, 0, 0};
    const int m_last = 63 - __clzll(pm);
    for (int m = 0; m <= m_last; ++m) {
        const int4 E1 = k3_site_counts(a, ki, si.p, m);
        const int4 E2 = k3_site_counts(a, kj, sj.p, m);
        C1[0] += E1.x; C1[1] += E1.y; C1[2] += E1.z; C1[3] += E1.w;
        C2[0] += E2.x; C2[1] += E2.y; C2[2] += E2.z; C2[3] += E2.w;
        if (!((pm >> m) & 1ull)) continue;
        if (!(k3_level_present(a, si.p, m, E1) && k3_level_present(a, sj.p, m, E2))) continue;   // updateMMs
        const int s1 = C1[0] + C1[1] + C1[2] + C1[3], s2 = C2[0] + C2[1] + C2[2] + C2[3];
        if (s1 + s2 < a.min_snp) continue;
        int A, al, B, bl;
        k3_major_minor(C1, A, al);
        k3_major_minor(C2, B, bl);
        if (C1[A] == 0 || C1[al] == 0 || C2[B] == 0 || C2[bl] == 0) continue;
        const uint32_t *mask = a.M > 1 ? a.mle + (size_t)m * a.nwp : nullptr;
        const int rA = k3_rank(si.bases, A), ra = k3_rank(si.bases, al);
        const int rB = k3_rank(sj.bases, B), rb = k3_rank(sj.bases, bl);
        const int cAB = k3_pair_count(si, rA, sj, rB, lo, hi, mask);
        const int cAb = k3_pair_count(si, rA, sj, rb, lo, hi, mask);
        const int caB = k3_pair_count(si, ra, sj, rB, lo, hi, mask);
        const int cab = k3_pair_count(si, ra, sj, rb, lo, hi, mask);
        k3_emit(a, si.p, sj.p, m, A, al, B, bl, cAB, cAb, caB, cab);
    }
}

// Self edges: a pair with two entries (first, second in column order) on ONE site gives the combo
// "b_first:b_second" on the edge (p, p) (itertools.combinations over the pair's entry list).  Rare: one thread per
// flagged site, exact and slow.
template <int kMode>   // 0 = position-major events, 1 = read-major segments, 2 = column words, 3 = read-major, per-tile site slots
__global__ void __launch_bounds__(128) k3_self_edges(k3_args a, isb_reads_dev rd, isb_cols_dev cd,
                                                     const int64_t *__restrict__ cand_lo, const int32_t *__restrict__ n_cand)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t S = a.S;
    if (kMode == 3) {
        const unsigned long long ns = *a.n_sites_dev;
        S = (int64_t)(ns < (unsigned long long)a.sites_cap ? ns : (unsigned long long)a.sites_cap);
    }
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < S; k += stride) {
        if (!a.has2[k] || a.meta[k].split < 0) continue;
        const int32_t p = a.site_pos[k];
        const int64_t abs_pos = (int64_t)p + a.start;
        const unsigned bases = a.site_flags[p] & 0xF;
        // entries of the site in column order: event range (position-major), candidate segments (read-major) or the
        // slots of the site's column list (column words)
        k3c_column col = {0, 0, 0};
        if (kMode == 2) col = k3c_site_column(cd, p + a.col_shift);
        int64_t lo = kMode == 2 ? 0 : (kMode == 1 ? cand_lo[k] : (kMode == 0 ? a.site_ev[2 * k] : 0));
        int64_t hi = kMode == 2 ? col.depth : (kMode == 1 ? lo + n_cand[k] : (kMode == 0 ? a.site_ev[2 * k + 1] : 0));
        if (kMode == 3) {                                          // candidate segments inside the site's tile range
            const int tl = p / K1R_TILE;
            lo = isb_lower_bound(rd.seg_start, rd.tile_lo[tl], rd.tile_hi[tl], abs_pos - rd.max_seg_len + 1);
            hi = isb_lower_bound(rd.seg_start, lo, rd.tile_hi[tl], abs_pos + 1);
        }
        auto entry = [&](int64_t e, int &b, int &rid) -> bool {
            if (kMode == 2) return k3c_candidate(cd, col, (int)e, a.n_pairs, b, rid) && ((bases >> b) & 1u);
            if (kMode == 1 || kMode == 3) return k3r_candidate(rd, e, abs_pos, b, rid) && ((bases >> b) & 1u);
            if (!k3_qualifies(a, e, bases)) return false;
            b = a.base[e];
            rid = a.read_id[e];
            return true;
        };
        int K[16];
        for (int i = 0; i < 16; ++i) K[i] = 0;
        int C[4] = {0, 0, 0, 0};
        for (int m = 0; m < a.M; ++m) {
            const int4 E = k3_site_counts(a, k, p, m);
            C[0] += E.x; C[1] += E.y; C[2] += E.z; C[3] += E.w;
            bool added = false;
            for (int64_t e2 = lo; e2 < hi; ++e2) {
                int b2, rid;
                if (!entry(e2, b2, rid)) continue;
                if ((a.M > 1 ? a.pair_mm[rid] : 0) != m) continue;
                for (int64_t e1 = lo; e1 < e2; ++e1) {
                    int b1, rid1;
                    if (entry(e1, b1, rid1) && rid1 == rid) {
                        K[b1 * 4 + b2] += 1;
                        added = true;
                    }
                }
            }
            if (!added) continue;
            if (!k3_level_present(a, p, m, E)) continue;
            const int s1 = C[0] + C[1] + C[2] + C[3];
            if (s1 + s1 < a.min_snp) continue;
            int A, al;
            k3_major_minor(C, A, al);
            if (C[A] == 0 || C[al] == 0) continue;
            k3_emit(a, p, p, m, A, al, A, al, K[A * 4 + A], K[A * 4 + al], K[al * 4 + A], K[al * 4 + al]);
        }
    }
}


// Enumeration over PER-TILE site slots (fused read-major path): the sites of a tile of K1R_TILE positions occupy the
// contiguous, position-ordered slots [tile_first, tile_first + tile_cnt); tiles got their slots in completion order, so
// the partners of a site are the later slots of its own tile, then the slots of the following tiles up to the end of
// the site's split.  Grid-stride over the device-side site count.
__global__ void __launch_bounds__(256) k3_enum_pairs_tiles(k3_args a)
{
    const unsigned long long ns = *a.n_sites_dev;
    const int64_t S = (int64_t)(ns < (unsigned long long)a.sites_cap ? ns : (unsigned long long)a.sites_cap);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < S * K3_ENUM_LANES; t += stride) {
        const int64_t k = t / K3_ENUM_LANES;
        const int sub = (int)(t % K3_ENUM_LANES);
        const isb_site_meta mi = a.meta[k];
        if (mi.split < 0 || mi.nw == 0) continue;
        const uint32_t *any_i = a.rows + a.row_off[k] - mi.wlo;
        const int i_hi = mi.wlo + mi.nw;
        const int32_t p_i = a.site_pos[k];
        int tl = p_i / K1R_TILE;
        const int64_t split_end = (int64_t)__ldg(a.splits + 2 * mi.split + 1) - a.start;   // relative, inclusive
        int last_tile = (int)(split_end / K1R_TILE);
        if (last_tile > a.n_tiles - 1) last_tile = a.n_tiles - 1;
        int64_t j_lo = k + 1, j_hi = (int64_t)a.tile_first[tl] + a.tile_cnt[tl];
        for (;;) {
            if (j_hi > S) j_hi = S;
            // partners in batches of 4 per thread: their records are requested together (the scan is a chain of dependent
            // loads; one record per trip left the kernel at 24 % issue-active)
            for (int64_t j0 = j_lo + sub; j0 < j_hi; j0 += 4 * K3_ENUM_LANES) {
                int4 rec[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t j = j0 + (int64_t)u * K3_ENUM_LANES;
                    rec[u] = make_int4(0, 0, 0, -2);
                    if (j < j_hi) rec[u] = __ldg(reinterpret_cast<const int4 *>(a.meta + j));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t j = j0 + (int64_t)u * K3_ENUM_LANES;
                    const int wlo_j = rec[u].y, nw_j = rec[u].z;                 // isb_site_meta: ev_lo_rel, wlo, nw, split
                    if (j >= j_hi || rec[u].w != mi.split) continue;
                    const int lo = max(mi.wlo, wlo_j), hi = min(i_hi, wlo_j + nw_j);
                    bool linked = false;
                    if (lo < hi) {
                        const uint32_t *any_j = a.rows + a.row_off[j] - wlo_j;
                        for (int w = lo; w < hi; ++w)
                            if (any_i[w] & any_j[w]) { linked = true; break; }
                    }
                    if (linked) {
                        const unsigned act = __activemask();
                        const int leader = __ffs(act) - 1, lane = threadIdx.x & 31;
                        unsigned long long base = 0;
                        if (lane == leader) base = atomicAdd(a.n_site_pairs, (unsigned long long)__popc(act));
                        base = __shfl_sync(act, base, leader);
                        const unsigned long long slot = base + __popc(act & ((1u << lane) - 1u));
                        if ((int64_t)slot < a.pair_cap) { a.pair_i[slot] = (int32_t)k; a.pair_j[slot] = (int32_t)j; }
                    }
                }
            }
            if (++tl > last_tile) break;
            j_lo = a.tile_first[tl];
            j_hi = j_lo + a.tile_cnt[tl];
        }
    }
}

// Linkage back end of the fused read-major path (M = 1): enumeration, statistics and self edges on the site slots K1f
// filled.  Site and pair counts stay on the device; the caller checks the capacities afterwards (isb_k1f_grow).
int isb_k3_backend_tiles(isb_ctx *ctx, const isb_reads_dev *rd, const isb_k3_tiles *ts, int64_t n_pairs, int32_t start, int32_t L,
                         const unsigned long long *nmask, const uint8_t *site_flags, int32_t n_splits, const int32_t *splits,
                         int min_snp, isb_ld_row *rows, int64_t cap)
{
    cudaStream_t st = ctx->stream;
    int rc;
    int64_t cap_pairs = (int64_t)(ctx->buf[SL_PAIRS].cap / (2 * sizeof(int32_t)));
    if (cap_pairs < 6 * ts->sites_cap) {
        if ((rc = isb_ensure(ctx, SL_PAIRS, 2 * sizeof(int32_t) * (size_t)(6 * ts->sites_cap)))) return rc;
        cap_pairs = (int64_t)(ctx->buf[SL_PAIRS].cap / (2 * sizeof(int32_t)));
    }
    k3_args a;
    memset(&a, 0, sizeof(a));
    a.n_pairs = n_pairs; a.start = start; a.L = L; a.M = 1; a.min_snp = min_snp;
    a.nmask = nmask; a.site_flags = site_flags; a.n_splits = n_splits; a.splits = splits;
    a.site_pos = ts->site_pos; a.meta = ts->meta; a.row_off = ts->row_off; a.has2 = ts->has2; a.rows = ts->rows;
    a.site_counts = ts->site_counts; a.n_tiles = ts->n_tiles; a.tile_first = ts->tile_first; a.tile_cnt = ts->tile_cnt;
    a.sites_cap = ts->sites_cap; a.n_sites_dev = ctx->d_counters + 2;
    a.pair_cap = cap_pairs;
    a.pair_i = (int32_t *)ctx->buf[SL_PAIRS].p;
    a.pair_j = a.pair_i + cap_pairs;
    a.out = rows; a.cap = cap;
    a.n_ld = ctx->d_counters + 1; a.n_site_pairs = ctx->d_counters + 3; a.d_err = ctx->d_err;
    a.seed = ctx->seed;
    const int grid = ctx->sm_count * 8;
    // enumeration: one thread per (site, lane) over the site SLOTS (the site count stays on the device; unused slots exit at
    // once) instead of a fixed grid-stride grid: with 1 - 2 trips per thread the last trip left the SMs half empty
    const int64_t enum_blocks = (ts->sites_cap * K3_ENUM_LANES + 255) / 256;
    k3_enum_pairs_tiles<<<(unsigned)(enum_blocks > grid ? (enum_blocks < (1 << 30) ? enum_blocks : (1 << 30)) : grid), 256, 0, st>>>(a);
    ISB_LAUNCH_CHECK();
    k3_pair_stats_dev<<<grid, 256, 0, st>>>(a);
    ISB_LAUNCH_CHECK();
    const isb_cols_dev cd_none = {};
    k3_self_edges<3><<<ctx->sm_count * 2, 128, 0, st>>>(a, *rd, cd_none, nullptr, nullptr);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

static int k3_run(isb_ctx *ctx, const isb_reads_dev *rd, const isb_cols_dev *cd, int64_t n, const int32_t *ref_pos, const uint8_t *base,
                  const uint8_t *qual, const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start,
                  int32_t L, int M, int min_qual, const int32_t *counts, const unsigned long long *nmask,
                  const uint8_t *site_flags, int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows,
                  int64_t cap, int32_t col_shift = 0)
{
    cudaStream_t st = ctx->stream;
    int rc;
    // [1] = linkage rows (appended across the chunks of a pipelined batch), [2] sites, [3] linked pairs, [4] row words
    if (ctx->keep_counters) ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 2, 0, 3 * sizeof(unsigned long long), st));
    else ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 1, 0, 4 * sizeof(unsigned long long), st));
    ctx->h_counters[2] = ctx->h_counters[3] = 0;
    if (L <= 0 || (!rd && !cd && n <= 0) || (rd && rd->n_segs <= 0) || (cd && cd->n_chunks <= 0)) return ISB_OK;

    // 1. ordered list of linkage-eligible sites
    const int nb = (int)(((int64_t)L + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if ((rc = isb_ensure(ctx, SL_SCAN_TMP, sizeof(int64_t) * (size_t)nb))) return rc;
    int64_t *block_sums = (int64_t *)ctx->buf[SL_SCAN_TMP].p;
    FlagFn ff{site_flags};
    scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(ff, (int64_t)L, block_sums);
    ISB_LAUNCH_CHECK();
    scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nb, ctx->d_counters + 2);
    ISB_LAUNCH_CHECK();
    ISB_CUDA(cudaMemcpyAsync(ctx->h_counters + 2, ctx->d_counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    const int64_t S = (int64_t)ctx->h_counters[2];
    if (S == 0) return ISB_OK;
    if ((rc = isb_ensure(ctx, SL_SITE_POS, sizeof(int32_t) * (size_t)S))) return rc;
    if ((rc = isb_ensure(ctx, SL_SITE_META, (sizeof(isb_site_meta) + 2 * sizeof(int64_t)) * (size_t)S))) return rc;
    if ((rc = isb_ensure(ctx, SL_SITE_WORDS, sizeof(int32_t) * (size_t)S))) return rc;
    if ((rc = isb_ensure(ctx, SL_ROW_OFF, sizeof(int64_t) * (size_t)S))) return rc;
    if ((rc = isb_ensure(ctx, SL_HAS2, (size_t)S))) return rc;
    int32_t *site_pos = (int32_t *)ctx->buf[SL_SITE_POS].p;
    SitePosSink sps{site_pos};
    scan_scatter<<<nb, SCAN_THREADS, 0, st>>>(ff, (int64_t)L, block_sums, sps);
    ISB_LAUNCH_CHECK();

    k3_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.ref_pos = ref_pos; a.base = base; a.qual = qual; a.read_id = read_id;
    a.n_pairs = n_pairs; a.pair_mm = pair_mm; a.start = start; a.L = L; a.M = M; a.min_qual = min_qual;
    a.min_snp = min_snp; a.counts = counts; a.nmask = nmask; a.site_flags = site_flags;
    a.n_splits = n_splits; a.splits = splits; a.col_shift = col_shift;
    a.S = S; a.site_pos = site_pos;
    a.site_ev = (int64_t *)ctx->buf[SL_SITE_META].p;
    a.meta = (isb_site_meta *)((int64_t *)ctx->buf[SL_SITE_META].p + 2 * S);
    a.site_words = (int32_t *)ctx->buf[SL_SITE_WORDS].p;
    a.row_off = (int64_t *)ctx->buf[SL_ROW_OFF].p;
    a.has2 = (uint8_t *)ctx->buf[SL_HAS2].p;
    a.out = rows; a.cap = cap;
    a.n_ld = ctx->d_counters + 1; a.n_site_pairs = ctx->d_counters + 3; a.d_err = ctx->d_err;
    a.seed = ctx->seed;

    // 2.-5. bit rows of the sites, then linked pairs
    const int64_t site_warps_blocks = (S * 32 + K3_THREADS - 1) / K3_THREADS;
    const int grid_sites = (int)(site_warps_blocks < (int64_t)ctx->sm_count * 32 ? site_warps_blocks : (int64_t)ctx->sm_count * 32);
    const int nbs = (int)((S + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if ((rc = isb_ensure(ctx, SL_SCAN_TMP, sizeof(int64_t) * (size_t)(nb > nbs ? nb : nbs)))) return rc;
    block_sums = (int64_t *)ctx->buf[SL_SCAN_TMP].p;
    if (M > 1) {                                                // mm <= m masks
        a.nwp = (n_pairs + 31) / 32;
        if ((rc = isb_ensure(ctx, SL_MM_MASK, sizeof(uint32_t) * (size_t)a.nwp * M))) return rc;
        k3_mm_masks<<<(int)((a.nwp + 255) / 256), 256, 0, st>>>(pair_mm, n_pairs, M, a.nwp, (uint32_t *)ctx->buf[SL_MM_MASK].p);
        ISB_LAUNCH_CHECK();
        a.mle = (const uint32_t *)ctx->buf[SL_MM_MASK].p;
    }
    int64_t *cand_lo = nullptr;
    int32_t *n_cand = nullptr;
    k3c_site_rec *recs = nullptr;
    const isb_reads_dev rd_none = {};
    const isb_cols_dev cd_none = {};
    if (!rd && !cd) {                                           // position-major columns
        const int tp = 1024;                                    // searches bounded by position-tile event offsets
        const int n_tiles = (L + tp - 1) / tp;
        if ((rc = isb_tile_offsets(ctx, ref_pos, n, start, L, tp, n_tiles))) return rc;
        a.tile_off = (const int64_t *)ctx->buf[SL_K3_TILE_OFF].p;
        a.tile_tp = tp;
        k3_site_ranges<<<(int)((S + 255) / 256), 256, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
        k3_site_windows<<<grid_sites, K3_THREADS, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
        WordsFn wf{a.site_words};                               // bit-row storage offsets
        scan_reduce<<<nbs, SCAN_THREADS, 0, st>>>(wf, S, block_sums);
        ISB_LAUNCH_CHECK();
        scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nbs, ctx->d_counters + 4);
        ISB_LAUNCH_CHECK();
        RowOffSink ros{(int64_t *)ctx->buf[SL_ROW_OFF].p};
        scan_scatter<<<nbs, SCAN_THREADS, 0, st>>>(wf, S, block_sums, ros);
        ISB_LAUNCH_CHECK();
        ISB_CUDA(cudaMemcpyAsync(ctx->h_counters + 4, ctx->d_counters + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        ISB_CUDA(cudaStreamSynchronize(st));
        const size_t total_words = (size_t)ctx->h_counters[4];
        if ((rc = isb_ensure(ctx, SL_ROWS, sizeof(uint32_t) * (total_words + 4)))) return rc;
        a.rows = (uint32_t *)ctx->buf[SL_ROWS].p;
        ISB_CUDA(cudaMemsetAsync(a.rows, 0, sizeof(uint32_t) * (total_words + 4), st));
        k3_build_rows<<<grid_sites, K3_THREADS, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
    } else {
        if (rd) {                                               // read-major segments: candidate ranges of the sites
            if ((rc = isb_ensure(ctx, SL_RD_CAND, (sizeof(int64_t) + sizeof(int32_t)) * (size_t)S))) return rc;
            cand_lo = (int64_t *)ctx->buf[SL_RD_CAND].p;
            n_cand = (int32_t *)(cand_lo + S);
            k3r_site_cand<<<(int)((S + 255) / 256), 256, 0, st>>>(a, *rd, cand_lo, n_cand);
        } else {                                                // column words: one record per site
            if ((rc = isb_ensure(ctx, SL_RD_CAND, sizeof(k3c_site_rec) * (size_t)S))) return rc;
            recs = (k3c_site_rec *)ctx->buf[SL_RD_CAND].p;
            k3c_site_prep<<<(int)((S + 255) / 256), 256, 0, st>>>(a, *cd, recs);
        }
        ISB_LAUNCH_CHECK();
        // initial guess of the row storage (words per site); the fused kernel reports the exact need if it is too small
        // row storage: S fixed slots + an initial guess of the overflow region (words per site; the fused kernel reports
        // the exact need if it is too small)
        static const int rows_init = getenv("ISB_K3_ROWS_INIT") ? atoi(getenv("ISB_K3_ROWS_INIT")) : 8;
        if ((rc = isb_ensure(ctx, SL_ROWS, sizeof(uint32_t) * ((size_t)S * (size_t)(K3_ROW_SCRATCH + (rows_init > 0 ? rows_init : 1)) + 64)))) return rc;
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (rd || cd) {                                         // fused gather + window + rows (no host sync needed)
            const int64_t row_cap = (int64_t)(ctx->buf[SL_ROWS].cap / sizeof(uint32_t));
            a.rows = (uint32_t *)ctx->buf[SL_ROWS].p;
            ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 4, 0, sizeof(unsigned long long), st));
            if (rd) k3r_site_rows<false><<<grid_sites, K3_THREADS, 0, st>>>(a, *rd, cd_none, nullptr, cand_lo, n_cand, (int64_t *)ctx->buf[SL_ROW_OFF].p,
                                                                          ctx->d_counters + 4, row_cap);
            else k3r_site_rows<true><<<grid_sites, K3_THREADS, 0, st>>>(a, rd_none, *cd, recs, nullptr, nullptr, (int64_t *)ctx->buf[SL_ROW_OFF].p,
                                                                      ctx->d_counters + 4, row_cap);
            ISB_LAUNCH_CHECK();
        }
        int64_t cap_pairs = (int64_t)(ctx->buf[SL_PAIRS].cap / (2 * sizeof(int32_t)));
        if (cap_pairs < 8 * S) {
            if ((rc = isb_ensure(ctx, SL_PAIRS, 2 * sizeof(int32_t) * (size_t)(8 * S)))) return rc;
            cap_pairs = (int64_t)(ctx->buf[SL_PAIRS].cap / (2 * sizeof(int32_t)));
        }
        a.pair_cap = cap_pairs;
        a.pair_i = (int32_t *)ctx->buf[SL_PAIRS].p;
        a.pair_j = a.pair_i + cap_pairs;
        ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 3, 0, sizeof(unsigned long long), st));
        static const int enum_variant = getenv("ISB_K3_ENUM") ? atoi(getenv("ISB_K3_ENUM")) : 1;
        if (enum_variant == 1) k3_enum_pairs_t<<<(int)((S * K3_ENUM_LANES + 255) / 256), 256, 0, st>>>(a);
        else k3_enum_pairs<<<grid_sites, K3_THREADS, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
        ISB_CUDA(cudaMemcpyAsync(ctx->h_counters + 3, ctx->d_counters + 3, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        ISB_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        ISB_CUDA(cudaStreamSynchronize(st));
        if ((rd || cd) && (*ctx->h_err & ISB_DEV_ERR_ROWBUF)) { // row storage too small: grow to the counted size, redo
            if (*ctx->h_err & ~ISB_DEV_ERR_ROWBUF) return ISB_OK;   // another error is pending: let the caller report it
            ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), st));
            const size_t need = (size_t)S * K3_ROW_SCRATCH + (size_t)ctx->h_counters[4];   // fixed slots + counted overflow
            if ((rc = isb_ensure(ctx, SL_ROWS, sizeof(uint32_t) * (need + need / 8 + 1024)))) return rc;
            continue;
        }
        const int64_t n_listed = (int64_t)ctx->h_counters[3];
        if (n_listed <= cap_pairs) {
            if (n_listed > 0) {
                k3_pair_stats<<<(int)((n_listed + 255) / 256), 256, 0, st>>>(a, n_listed);
                ISB_LAUNCH_CHECK();
            }
            break;
        }
        if ((rc = isb_ensure(ctx, SL_PAIRS, 2 * sizeof(int32_t) * (size_t)(n_listed + n_listed / 8)))) return rc;
        if (attempt == 3) return isb_fail(ctx, ISB_ERR_CUDA, "linkage: scratch sizing did not converge");
    }
    const int grid_self = (int)((S + 127) / 128 < (int64_t)ctx->sm_count * 8 ? (S + 127) / 128 : (int64_t)ctx->sm_count * 8);
    if (rd) k3_self_edges<1><<<grid_self, 128, 0, st>>>(a, *rd, cd_none, cand_lo, n_cand);
    else if (cd) k3_self_edges<2><<<grid_self, 128, 0, st>>>(a, rd_none, *cd, nullptr, nullptr);
    else k3_self_edges<0><<<grid_self, 128, 0, st>>>(a, rd_none, cd_none, nullptr, nullptr);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

int isb_k3_launch(isb_ctx *ctx, int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                  const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start, int32_t L, int M,
                  int min_qual, const int32_t *counts, const unsigned long long *nmask, const uint8_t *site_flags,
                  int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows, int64_t cap)
{
    return k3_run(ctx, nullptr, nullptr, n, ref_pos, base, qual, read_id, n_pairs, pair_mm, start, L, M, min_qual, counts, nmask,
                  site_flags, n_splits, splits, min_snp, rows, cap);
}

int isb_k3_launch_reads(isb_ctx *ctx, const isb_reads_dev *rd, int64_t n_pairs, const uint8_t *pair_mm, int32_t start,
                        int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                        const uint8_t *site_flags, int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows,
                        int64_t cap)
{
    return k3_run(ctx, rd, nullptr, 0, nullptr, nullptr, nullptr, nullptr, n_pairs, pair_mm, start, L, M, 0, counts, nmask,
                  site_flags, n_splits, splits, min_snp, rows, cap);
}

int isb_k3_launch_cols(isb_ctx *ctx, const isb_cols_dev *cd, int64_t n_pairs, const uint8_t *pair_mm, int32_t start,
                       int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                       const uint8_t *site_flags, int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows,
                       int64_t cap, int32_t col_shift)
{
    if (cd->n_chunks > 0 && !cd->ids) return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: ids is required for linkage");
    return k3_run(ctx, nullptr, cd, 0, nullptr, nullptr, nullptr, nullptr, n_pairs, pair_mm, start, L, M, 0, counts, nmask,
                  site_flags, n_splits, splits, min_snp, rows, cap, col_shift);
}
