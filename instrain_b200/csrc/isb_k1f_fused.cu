// isb_k1f_fused.cu -- K1f: the per-position part of the path straight from BAM-ORDER aligned segments, sm_100a.
//
//   k1f_pileup<M1, fused>   pileup counts (+ at M = 1, fused: the per-site SNV call and the bit rows of the linkage sites)
//
// Replaces, for a whole batch of scaffolds at once: samfile.pileup(...) column iteration
// (inStrain/profile/profile_utilities.py:150-153) + get_base_counts_mm (:268-286); fused: update_covT (:288-295),
// update_snp_table / call_snv_site / calc_snp_class / calculate_clonality (inStrain/profile/snv_utilities.py:40-231) and
// update_linked_reads (inStrain/profile/linkage.py:254-283).
//
// Input = isb_reads_batch (include/instrain_b200.h): one 4-bit one-hot code per aligned base, stored once per READ, in
// BAM order.  The pileup is a transposition (reads x offsets -> positions); it is done on the fly, in registers:
//
//   * a block owns a tile of 1024 positions, a thread one column word (8 positions); the segment TABLE of the tile
//     (seg_start in (tile_first - max_seg_len, tile_end): a contiguous range of the start-sorted table) is staged in
//     shared memory as one packed word per segment; the nibble WORDS are never staged: word k of a segment is needed by
//     exactly one thread of the grid, so it goes from HBM to that thread's register with one coalesced load;
//   * CIRCULAR SCHEDULE.  Thread t needs the segments [cl_t, ch_t) (~coverage many).  A warp steps through them
//     together: with Pm = max_t (ch_t - cl_t) rounded up to 8 and base = min_t cl_t, lane t visits at step s the one
//     index j in [cl_t, cl_t + Pm) with j = base + s (mod Pm) -- a rotation of its own range.  Lanes whose ranges
//     overlap therefore visit the SAME segment at the same step: the shared-memory read of its table word is a
//     broadcast and their word loads are consecutive addresses (the lanes of a warp fall into <= 3 such bands).  The
//     free-running per-thread loops of the first K1r put every lane on its own segment at every step (3.1x bank
//     conflicts, 32 sectors per load).  A lock-step loop over the union of the ranges would be conflict-free too, but
//     runs 2.6x as many steps;
//   * counting is bit-sliced (isb_bitslice.cuh): 24 logic ops per 8 words;
//   * fused epilogue (M = 1): the warp's 256 count quads are transposed through shared memory to "lane = position";
//     sites with one base only (or below min_cov) are finished with integer compares and coalesced stores, the ~12 %
//     that need the double-precision clonality / the allele tests are compacted into dense batches of 32 first;
//   * linkage front end: the tile's anySNP sites get contiguous site slots (one atomic per tile); one warp per site
//     gathers the site's (pair id, base) entries from the staged segment table + the words (L1 / L2 hits: the block
//     has just streamed them) and assembles the bit rows  any | ge1[b] | ge2[b]  with ballots and REDUX.OR (no atomics).
//
// HBM traffic, fused: 0.5 B per aligned base + 14 B per segment + 1 B per position in, 9 B per position + 32 B per SNV
// row + ~300 B per linkage site out; the counts of a position never leave the SM.
#include "isb_common.cuh"
#include "isb_bitslice.cuh"
#include "isb_k2_site.cuh"
#include "isb_k3_dev.cuh"
#include <climits>
#include <cstdlib>

#define K1F_TILE K1R_TILE
#define K1F_THREADS (K1F_TILE / 8)     // one thread per column word
#define K1F_WARPS (K1F_THREADS / 32)
#define K1F_MAXLEN 256                 // hard cap of max_seg_len
#ifndef K1F_LEVELS
#define K1F_LEVELS 32                  // mm levels per pass of the M > 1 kernel (shared-memory accumulators)
#endif
#ifndef K1F_STAGE_IT
#define K1F_STAGE_IT 2                 // segment-table elements per thread and staging pass (measured 1 / 2 / 4 / 8: K1f 0.432 / 0.411 / 0.427 / 0.478 ms per 2e7 positions)
#endif
#define K1F_SEG_CAP_MAX 6144           // segments staged per chunk at most (8 B each: one chunk per tile up to ~700x coverage)
#define K1F_TILE4 (256 + 32)           // count quads of a warp's 256 positions + one pad quad per 8 positions
#define K1F_CODE_IDS_MIN 512           // pair-id window (ids) of a site the ballot row builder handles at least; sized per batch
#define K1F_CODE_IDS_MAX 4096          // from the pair density (k1f_args.code_ids); wider windows: atomics
#ifndef K1F_MINB
#define K1F_MINB 8                     // __launch_bounds__ min blocks per SM of the M = 1 kernels (64 registers; ~28 KB shared)
#endif
#ifndef K1F_MM_MINB
#define K1F_MM_MINB 1                  // the same for the M > 1 kernel (shared-memory accumulators bound its occupancy, not registers)
#endif
static_assert(K1F_WARPS == 4, "the site bookkeeping of the fused epilogue assumes 4 warps per tile");

struct k1f_args {
    isb_reads_dev rd;
    const uint8_t *pair_mm;
    int64_t n_pairs;
    int32_t start, L;
    int M;
    int seg_cap;                       // segments staged per chunk
    int pf_ahead;                      // L2 prefetch distance in tiles (0: own tile only)
    int32_t *counts;
    unsigned int *d_err;
    // fused epilogue (M = 1)
    isb_k2_fuse k2;
    const int32_t *thr2;
    int n_lut;
    // site queue: the sites the epilogue does not finish itself (k2q_sites does)
    int32_t *q_first, *q_cnt;          // [n_tiles] queue slots of a tile (contiguous, position-ordered)
    int32_t *q_pos;                    // [q_cap] relative position
    int4 *q_E;                         // [q_cap] A,C,T,G counts
    uint32_t *q_cand;                  // [q_cap] candidate segments of the site's column: first (relative to tile_lo) | count << 16
    int64_t q_cap;
    unsigned long long *n_queue;
};
#define K1F_NO_CAND 0xffffffffu        // q_cand: not known (tile staged in several chunks): k3f_site_rows searches

// First index in [lo, hi) with a[idx * stride] >= key, searched by a whole WARP: 32 probes per round (a 32-ary search: 5
// dependent memory round trips over 1e7 entries instead of 24).  All lanes return the result.
__device__ __forceinline__ int64_t k1f_warp_lower_bound(const int32_t *__restrict__ a, int stride, int64_t lo, int64_t hi, int64_t key,
                                                        int lane)
{
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 31) >> 5;
        int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
        if (idx > hi - 1) idx = hi - 1;
        const bool below = (int64_t)__ldg(a + idx * stride) < key;
        const int c = __popc(__ballot_sync(ISB_FULL, below));     // probes are ascending: the first c are below the key
        int64_t i_c = lo + (int64_t)(c + 1) * step - 1;           // first probe at or above the key (if c < 32)
        if (i_c > hi - 1) i_c = hi - 1;
        int64_t i_p = lo + (int64_t)c * step - 1;                 // last probe below it (if c > 0)
        if (i_p > hi - 1) i_p = hi - 1;
        if (c < 32) hi = i_c;
        if (c > 0) lo = i_p + 1;
    }
    const bool below = lo + lane < hi && (int64_t)__ldg(a + (lo + lane) * stride) < key;
    return lo + __popc(__ballot_sync(ISB_FULL, below));
}

// tile t covers relative positions [t * TILE, (t + 1) * TILE): its candidate segment range, the piece of the word stream
// those segments occupy (16-byte aligned: what K1f prefetches into L2), and (linkage) the index of the last split that
// starts at or before the tile's first position (-1: none) -- the per-site split lookup walks forward from there instead
// of searching the whole table.  One warp per tile.
__global__ void __launch_bounds__(256)
k1f_tile_bounds(const int32_t *__restrict__ seg_start, const int64_t *__restrict__ seg_word, int64_t n_segs, int64_t n_words,
                int32_t start, int n_tiles, int max_seg_len, int64_t *__restrict__ tile_lo, int64_t *__restrict__ tile_hi,
                int64_t *__restrict__ tile_wlo, int64_t *__restrict__ tile_whi, const int32_t *__restrict__ splits, int n_splits,
                int32_t *__restrict__ tile_split)
{
    const int lane = threadIdx.x & 31;
    const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= n_tiles) return;                                     // warp-uniform
    const int64_t first = (int64_t)start + (int64_t)t * K1F_TILE;
    const int64_t lo = k1f_warp_lower_bound(seg_start, 1, 0, n_segs, first - max_seg_len + 1, lane);
    const int64_t hi = k1f_warp_lower_bound(seg_start, 1, lo, n_segs, first + K1F_TILE, lane);
    int sp = -1;
    if (tile_split) sp = (int)k1f_warp_lower_bound(splits, 2, 0, n_splits, first + 1, lane) - 1;   // last split with start <= first
    if (lane == 0) {
        tile_lo[t] = lo;
        tile_hi[t] = hi;
        int64_t b0 = 0, b1 = 0;
        if (hi > lo) {
            const int64_t w_first = __ldg(seg_word + lo), w_last = __ldg(seg_word + hi - 1) + (K1F_MAXLEN / 8 + 2);
            b0 = (w_first - 1 > 0 ? w_first - 1 : 0) & ~(int64_t)3;
            b1 = (w_last < n_words ? w_last : n_words) & ~(int64_t)3;
            if (b1 < b0 || b0 > n_words) b0 = b1 = 0;             // a table that breaks the layout rules: K1f reports it
        }
        tile_wlo[t] = b0;
        tile_whi[t] = b1;
        if (tile_split) tile_split[t] = sp;
    }
}

// warp reductions of 32-bit integers: one REDUX instruction each (sm_80+), not a 5-step shuffle ladder
__device__ __forceinline__ int k1f_warp_max(int v) { return __reduce_max_sync(ISB_FULL, v); }
__device__ __forceinline__ int k1f_warp_sum(int v) { return __reduce_add_sync(ISB_FULL, v); }
__device__ __forceinline__ int k1f_warp_min(int v) { return __reduce_min_sync(ISB_FULL, v); }

// M > 1: write (or add, once counts hold a partial sum) the thread's shared 8-bit counters to its cells of `counts`
__device__ __forceinline__ void k1f_flush_levels(const k1f_args &a, uint32_t *s_acc, int t, int Mg, int m_base, int32_t P,
                                                 bool add, bool clear)
{
    int4 *c4 = reinterpret_cast<int4 *>(a.counts);
    for (int m = 0; m < Mg; ++m) {
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            w8[j] = s_acc[(size_t)(m * 8 + j) * K1F_THREADS + t];
            if (clear) s_acc[(size_t)(m * 8 + j) * K1F_THREADS + t] = 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (P + k >= a.L) break;
            const int sh = (k >> 1) * 8, h = k & 1;
            int4 val;
            val.x = (w8[0 + h] >> sh) & 0xff; val.y = (w8[2 + h] >> sh) & 0xff;
            val.z = (w8[4 + h] >> sh) & 0xff; val.w = (w8[6 + h] >> sh) & 0xff;
            int4 *dst = c4 + ((size_t)(P + k) * a.M + m_base + m);
            if (add) { const int4 o = *dst; val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w; }
            *dst = val;
        }
    }
}

// M = 1: add the thread's vertical counter planes into its 8 count quads of the warp's shared-memory tile and clear them.
// The 32-bit counts live in shared memory, not in registers: they are touched once per <= 248 steps, and 32 registers
// less per thread is one more resident block per SM.
__device__ __forceinline__ void k1f_flush_planes(int4 *tile_lane, uint32_t (&pl)[8])
{
    int c[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) c[k][b] = 0;
    k1r_planes_to_counts(c, pl);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int4 v = tile_lane[k];
        v.x += c[k][0]; v.y += c[k][1]; v.z += c[k][2]; v.w += c[k][3];
        tile_lane[k] = v;
    }
}

template <bool kM1, bool kFuse>
__global__ void __launch_bounds__(K1F_THREADS, kM1 ? K1F_MINB : K1F_MM_MINB) k1f_pileup(k1f_args a)
{
    static_assert(kM1 || !kFuse, "the fused epilogue is the M = 1 SNV call");
    extern __shared__ __align__(16) unsigned char k1f_smem[];
    // one packed word per staged segment: bits 11.. = index (relative to the chunk's word base wb) of the word that covers
    // tile column 0, + 160; bits 0..10 = tile-relative end + 256
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(k1f_smem);
    const uint32_t s_meta_sa = isb_smem_u32(s_meta);                      // its shared-space byte address
    int32_t *s_start = reinterpret_cast<int32_t *>(s_meta + a.seg_cap);   // start relative to a.start (sorted): the search key
    uint8_t *s_mm = reinterpret_cast<uint8_t *>(s_start + a.seg_cap);     // M > 1 only
    uint32_t *s_zero = reinterpret_cast<uint32_t *>(k1f_smem + ((((size_t)a.seg_cap * (kM1 ? 8 : 9)) + 15) & ~(size_t)15));   // one zero table word
    unsigned char *s_x = reinterpret_cast<unsigned char *>(s_zero) + 16;
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(s_x);                  // M > 1: [Mg * 8][K1F_THREADS]
    // fused epilogue scratch
    int4 *s_tile = reinterpret_cast<int4 *>(s_x);                         // [K1F_WARPS][K1F_TILE4]
    int32_t *s_cl = reinterpret_cast<int32_t *>(s_tile + K1F_WARPS * K1F_TILE4);       // [K1F_THREADS] candidate range per column
    int32_t *s_ch = s_cl + K1F_THREADS;
    int32_t *s_misc = s_ch + K1F_THREADS;                                 // [16]: queued sites per warp, offsets, slot base
    uint8_t *s_q = reinterpret_cast<uint8_t *>(s_misc + 16);              // [K1F_WARPS][256]: positions for the site queue

    const int t = threadIdx.x;
    const int lane = t & 31, wib = t >> 5;
    const int tile = blockIdx.x;
    const int32_t T0 = tile * K1F_TILE;
    const int32_t P = T0 + t * 8;                               // first of the thread's 8 positions (relative)
    const bool active = P < a.L;
    const int maxlen = a.rd.max_seg_len;
    const int64_t lo = a.rd.tile_lo[tile], hi = a.rd.tile_hi[tile];
    const int m_base = kM1 ? 0 : (int)blockIdx.y * K1F_LEVELS;
    const int Mg = kM1 ? 1 : min(K1F_LEVELS, a.M - m_base);
    const bool single = hi - lo <= (int64_t)a.seg_cap;            // the whole tile in one chunk (the usual case)

    uint32_t pl[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};            // M = 1: vertical counter planes (weights 1 .. 128)
    int n8 = 0;                                                   // steps since the last flush (counters hold <= 255)
    bool spilled = false;                                         // M > 1: counts already hold a partial sum
    int4 *tile4 = s_tile + wib * K1F_TILE4;                       // M = 1: the warp's count quads (pad quad per 8 positions)
    int4 *tile_lane = tile4 + lane * 9;                           //        this thread's 8
    if (kM1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tile_lane[k] = make_int4(0, 0, 0, 0);
        if (kFuse) { s_cl[t] = 0; s_ch[t] = 0; }
    } else {
        for (int w = 0; w < Mg * 8; ++w) s_acc[w * K1F_THREADS + t] = 0u;
    }
    if (t == 0) *s_zero = 0u;                                     // visible after the first barrier of the staging loop
    const uint32_t zero_sa = isb_smem_u32(s_zero);
    unsigned err = 0;
    int64_t wb = 0;                                               // word base of the (last) chunk
    const int P_end = t * 8 + 256;

    // The words of a tile are one contiguous piece of the stream: the TMA engine pulls it into L2 ahead of the loads.  Each
    // block asks for the tile that starts `pf_ahead` tiles later (blocks start in index order: ~1/8 of a block's lifetime
    // ahead = several microseconds, many DRAM latencies; ~12 % more data resident in L2) together with its piece of the
    // segment table, and for its own words in case nobody did (the first tiles of the grid).
    if (t == 0) {
        auto pf = [](const void *ptr, int64_t bytes) {
            const uintptr_t a0 = (uintptr_t)ptr & ~(uintptr_t)15, a1 = ((uintptr_t)ptr + (uintptr_t)bytes + 15) & ~(uintptr_t)15;
            if (bytes > 0 && a1 - a0 < (1u << 24))
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
        };
        const int t2 = tile + a.pf_ahead;
        int64_t w0 = __ldg(a.rd.tile_wlo + tile), w1 = __ldg(a.rd.tile_whi + tile);
        int64_t lo2 = 0, hi2 = 0, v0 = 0, v1 = 0;
        if (a.pf_ahead > 0 && t2 < a.rd.n_tiles) {
            lo2 = __ldg(a.rd.tile_lo + t2); hi2 = __ldg(a.rd.tile_hi + t2);
            v0 = __ldg(a.rd.tile_wlo + t2); v1 = __ldg(a.rd.tile_whi + t2);
        }
        if (tile < a.pf_ahead || a.pf_ahead <= 0) pf(a.rd.words + w0, (w1 - w0) * 4);
        if (hi2 > lo2) {
            pf(a.rd.words + v0, (v1 - v0) * 4);
            pf(a.rd.seg_start + lo2, (hi2 - lo2) * 4);
            pf(a.rd.seg_len + lo2, (hi2 - lo2) * 2);
            pf(a.rd.seg_word + lo2, (hi2 - lo2) * 8);
            if (!kM1) pf(a.rd.seg_pair + lo2, (hi2 - lo2) * 4);
        }
    }
    for (int64_t c0 = lo; c0 < hi; c0 += a.seg_cap) {
        const int nc = (int)min((int64_t)a.seg_cap, hi - c0);
        __syncthreads();                                          // previous chunk consumed
        wb = __ldg(a.rd.seg_word + c0) - 1;                       // the separator in front of the chunk's first segment
        if (wb < 0 || wb >= a.rd.n_words) {                       // layout rules violated (block-uniform)
            err |= ISB_DEV_ERR_SEG;
            continue;
        }
        // Segment table of the chunk -> shared memory, K1F_STAGE_IT * 128 segments per pass.  All global loads of the
        // elements a thread handles in a pass are issued before the first is used: one memory latency per pass (the other
        // resident blocks cover it); fewer elements in flight = fewer registers = one more resident block.
        for (int pb = 0; pb < nc; pb += K1F_STAGE_IT * K1F_THREADS) {
            int32_t r_s[K1F_STAGE_IT], r_prev[K1F_STAGE_IT], r_pid[K1F_STAGE_IT];
            int r_n[K1F_STAGE_IT];
            int64_t r_w[K1F_STAGE_IT];
#pragma unroll
            for (int k = 0; k < K1F_STAGE_IT; ++k) {
                const int i = pb + t + k * K1F_THREADS;
                const int64_t g = c0 + i;
                r_s[k] = 0; r_prev[k] = INT_MIN; r_n[k] = 1; r_w[k] = wb + 1; r_pid[k] = 0;
                if (pb + k * K1F_THREADS >= nc) break;             // block-uniform
                if (i < nc) {
                    r_s[k] = __ldg(a.rd.seg_start + g);
                    r_n[k] = __ldg(a.rd.seg_len + g);
                    r_w[k] = __ldg(a.rd.seg_word + g);
                    if (g > 0) r_prev[k] = __ldg(a.rd.seg_start + g - 1);
                    if (!kM1) r_pid[k] = __ldg(a.rd.seg_pair + g);
                }
            }
            int r_mm[K1F_STAGE_IT];
            if (!kM1) {
#pragma unroll
                for (int k = 0; k < K1F_STAGE_IT; ++k) {
                    r_mm[k] = 255;
                    if (pb + t + k * K1F_THREADS < nc && r_pid[k] >= 0 && (int64_t)r_pid[k] < a.n_pairs)
                        r_mm[k] = __ldg(a.pair_mm + r_pid[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K1F_STAGE_IT; ++k) {
                const int i = pb + t + k * K1F_THREADS;
                if (i >= nc) break;
                const int64_t s64 = (int64_t)r_s[k] - a.start;
                const int n = r_n[k];
                const int64_t wl = r_w[k] - wb;
                const int nw = (int)(((s64 & 7) + n + 7) >> 3);
                bool bad = n < 1 || n > maxlen || s64 < 0 || s64 + n > (int64_t)a.L || wl < 1 || wl > (1 << 20) ||
                           r_w[k] + nw + 1 > a.rd.n_words || r_prev[k] > r_s[k];
                if (bad) err |= ISB_DEV_ERR_SEG;
                const int32_t s = bad ? T0 : (int32_t)s64;
                const int n_c = bad ? 0 : n;                      // a segment that breaks the rules covers nothing
                const int wl_c = bad ? 1 : (int)wl;
                const int s_rel = min(max(s - T0, -255), K1F_TILE - 1);   // candidates start in (T0 - 256, T0 + 1024)
                s_meta[i] = ((uint32_t)(wl_c - (s_rel >> 3) + 160) << 11) | (uint32_t)(s_rel + n_c + 256);
                s_start[i] = s;
                if (!kM1) {
                    if (r_mm[k] >= a.M) { err |= ISB_DEV_ERR_MM; r_mm[k] = 255; }
                    s_mm[i] = (uint8_t)r_mm[k];
                }
            }
        }
        __syncthreads();

        // candidates of this thread inside the chunk: seg_start in (P - maxlen, P + 8)
        int cl = 0, ch = 0;
        if (active) {
            int l = 0, h = nc;
            const int key = P - maxlen + 1;
            while (l < h) { const int mid = (l + h) >> 1; if (s_start[mid] < key) l = mid + 1; else h = mid; }
            cl = l;
            h = nc;
            const int key2 = P + 8;
            while (l < h) { const int mid = (l + h) >> 1; if (s_start[mid] < key2) l = mid + 1; else h = mid; }
            ch = l;
        }
        if (kFuse) { s_cl[t] = cl; s_ch[t] = ch; }
        // circular schedule of the warp (see the header)
        const int len = ch - cl;
        const int Pm = (k1f_warp_max(len) + 7) & ~7;
        if (Pm == 0) continue;                                    // warp-uniform; no block barrier inside the loop body below
        const int base = k1f_warp_min(len > 0 ? cl : INT_MAX);
        int j = cl;
        if (len > 0) {
            const int r = (cl - base) % Pm;
            j = cl + (r ? Pm - r : 0);
        }
        const int wrap = cl + Pm;
        // the 8 one-hot nibbles of segment i at the thread's positions (0 where the segment does not reach).  The stream
        // is position-aligned: the thread's column is ONE word of the segment, no shift.
        const uint32_t *wsrc = a.rd.words + wb + (t - 160);
        asm volatile("" : "+l"(wsrc));                            // keep the pointer in registers (ptxas re-derived it per load: 5 instructions)
        // the schedule runs on shared-memory BYTE addresses of the table words (explicit ld.shared: through a generic
        // pointer every predicated load re-derived the shared window base, 4 extra instructions each)
        const uint32_t ja_cl = s_meta_sa + 4u * (uint32_t)cl, ja_ch = s_meta_sa + 4u * (uint32_t)ch, ja_wrap = s_meta_sa + 4u * (uint32_t)wrap;
        uint32_t ja = s_meta_sa + 4u * (uint32_t)j;
        // table word of the lane's current segment; a step beyond the lane's range (jj may run past ch, into the start
        // column) reads the zero word instead: end field 0 = "covers nothing"
        auto meta_at = [&](uint32_t &jj) -> uint32_t {
            uint32_t md;
            const uint32_t src = jj < ja_ch ? jj : zero_sa;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(md) : "r"(src));
            jj = (jj + 4u == ja_wrap) ? ja_cl : jj + 4u;
            return md;
        };
        auto word_of = [&](uint32_t md) -> uint32_t {
            const bool ok = P_end < (int)(md & 0x7ffu);
            return ok ? __ldg(wsrc + (md >> 11)) : 0u;
        };
        if (kM1) {
            // Software pipeline: the 8 loads of block b + 1 are issued before the carry-save adds of block b, so a warp
            // always has 8 .. 16 word loads in flight (the words come from HBM: ~1 us; with the loads of one block at a
            // time the kernel ran at a third of its issue rate).
            auto load8 = [&](uint32_t (&x)[8]) {
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = meta_at(ja);
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = word_of(x[u]);
            };
            auto flush_if = [&](int added) {
                n8 += added;
                if (n8 > 239) {                                    // the next 16 could overflow 255
                    k1f_flush_planes(tile_lane, pl);
                    n8 = 0;
                }
            };
            const int nb = Pm >> 3;                                // Pm is a multiple of 8
            uint32_t x[8], y[8];
            load8(x);                                              // block 0
            int b = 1;
            for (; b + 1 < nb; b += 2) {                           // invariant: x holds block b - 1, not yet added
                load8(y);
                k1r_add8(pl, x);
                load8(x);
                k1r_add8(pl, y);
                flush_if(16);
            }
            if (b < nb) {
                load8(y);
                k1r_add8(pl, x);
                k1r_add8(pl, y);
                flush_if(16);
            } else {
                k1r_add8(pl, x);
                flush_if(8);
            }
        } else {
            for (int s = 0; s < Pm; s += 8) {
                uint32_t x[8];
                int lv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int jcur = (int)((ja - s_meta_sa) >> 2);
                    lv[u] = jcur < ch ? (int)s_mm[jcur] - m_base : -1;
                    x[u] = meta_at(ja);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = word_of(x[u]);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if ((unsigned)lv[u] < (unsigned)Mg) {          // 8-bit counters per (level, base, even/odd position)
                        uint32_t *acc = s_acc + (size_t)(lv[u] * 8) * K1F_THREADS + t;
                        acc[0 * K1F_THREADS] += x[u] & 0x01010101u;
                        acc[1 * K1F_THREADS] += (x[u] >> 4) & 0x01010101u;
                        acc[2 * K1F_THREADS] += (x[u] >> 1) & 0x01010101u;
                        acc[3 * K1F_THREADS] += (x[u] >> 5) & 0x01010101u;
                        acc[4 * K1F_THREADS] += (x[u] >> 2) & 0x01010101u;
                        acc[5 * K1F_THREADS] += (x[u] >> 6) & 0x01010101u;
                        acc[6 * K1F_THREADS] += (x[u] >> 3) & 0x01010101u;
                        acc[7 * K1F_THREADS] += (x[u] >> 7) & 0x01010101u;
                    }
                }
                n8 += 8;
                if (n8 > 240) {                                    // flush before a byte can overflow (warp-uniform)
                    if (active) k1f_flush_levels(a, s_acc, t, Mg, m_base, P, spilled, true);
                    spilled = true;
                    n8 = 0;
                }
            }
        }
    }
    if (err) atomicOr(a.d_err, err);

    if (!kFuse) {
        if (!active) return;
        if (kM1) {
            k1f_flush_planes(tile_lane, pl);
            int4 *c4 = reinterpret_cast<int4 *>(a.counts) + P;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (P + k >= a.L) break;
                c4[k] = tile_lane[k];
            }
        } else {
            k1f_flush_levels(a, s_acc, t, Mg, m_base, P, spilled, false);
        }
        return;
    }

    // ---- fused epilogue (M = 1) ---------------------------------------------------------------------------------------
    // Coverage of every position; sites with one base only (or below min_cov) are finished here with integer compares and
    // coalesced stores.  The others (~12 % at 100x: a second base, usually a sequencing error) go to the tile's slots of
    // the SITE QUEUE (position, the four counts, candidate segment range of the column) and are finished by k2q_sites.
    // Keeping the double-precision site arithmetic, the re-drawn clonality and the linkage row builder out of this kernel
    // keeps its code inside the instruction caches: the one-kernel version spent a third of its issue slots waiting for
    // instruction fetches (ncu: stall_no_inst 5.2 per issue, 138 KB of SASS).
    const int32_t W0 = T0 + wib * 256;
    unsigned long long ref8 = 0ull;                               // reference bases of the lane's 8 epilogue positions: requested
#pragma unroll                                                    // now, used after the flush (the loads were the longest stall)
    for (int rd = 0; rd < 8; ++rd) {
        const int32_t p = W0 + rd * 32 + lane;
        if (p < a.L) ref8 |= (unsigned long long)__ldg(a.k2.ref + p) << (8 * rd);
    }
    k1f_flush_planes(tile_lane, pl);
    __syncwarp();
    int4 *counts4 = reinterpret_cast<int4 *>(a.counts);
    const bool full_counts = a.counts != nullptr && a.k2.full_counts;
    uint8_t *q_list = s_q + wib * 256;
    int nq = 0;                                                   // positions of the warp for the queue (warp-uniform)
    const bool all_general = a.k2.min_cov < 1;                    // then "below min_cov" is not a simple case
    const int cov_r = a.k2.clonTR ? a.k2.cov_r : 0;               // rarefied clonality wanted where the coverage reaches it
#pragma unroll 2
    for (int rd = 0; rd < 8; ++rd) {
        const int q = rd * 32 + lane;
        const int32_t p = W0 + q;
        const bool in = p < a.L;
        const int4 E = tile4[q + (q >> 3)];
        const int T = E.x + E.y + E.z + E.w;
        const int mx = max(max(E.x, E.y), max(E.z, E.w));
        bool simple = false;
        float clon = CUDART_NAN_F, clonr = CUDART_NAN_F;
        if (in && !all_general) {
            if (T < a.k2.min_cov) {
                simple = true;                                    // call_snv_site -> (None, 0): coverage only
            } else if (mx == T && T < a.n_lut) {                  // one base only: clonality exactly 1; a row unless it is the
                const int con = E.x == T ? 0 : (E.y == T ? 1 : (E.z == T ? 2 : 3));   // reference base and passes the threshold
                if (T >= __ldg(a.thr2 + T) && con == (int)((ref8 >> (8 * rd)) & 0xffull)) { simple = true; clon = 1.0f; }
            }
            if (simple && cov_r > 0 && T >= cov_r) {              // rarefied clonality: 1 with one base only, else drawn (queue)
                if (mx == T) clonr = 1.0f; else simple = false;
            }
        }
        if (in) {
            a.k2.covT[p] = T;
            if (simple) {
                a.k2.clonT[p] = clon;
                a.k2.site_flags[p] = 0;
                if (a.k2.clonTR) a.k2.clonTR[p] = clonr;
                if (full_counts) counts4[p] = E;
            }
        }
        const bool need = in && !simple;
        const unsigned nm = __ballot_sync(ISB_FULL, need);
        if (need) q_list[nq + __popc(nm & ((1u << lane) - 1u))] = (uint8_t)q;
        nq += __popc(nm);
    }
    // queue slots of the tile: contiguous and position-ordered (warp after warp), one atomic per tile
    if (lane == 0) s_misc[wib] = nq;
    __syncthreads();
    if (t == 0) {
        int tot = 0;
        for (int w = 0; w < K1F_WARPS; ++w) { s_misc[4 + w] = tot; tot += s_misc[w]; }
        unsigned long long b0 = 0;
        if (tot) b0 = atomicAdd(a.n_queue, (unsigned long long)tot);
        s_misc[9] = (int32_t)(b0 < (unsigned long long)INT_MAX ? b0 : (unsigned long long)INT_MAX);
        a.q_first[tile] = s_misc[9];
        a.q_cnt[tile] = tot;
    }
    __syncthreads();
    const int64_t qb = (int64_t)s_misc[9] + s_misc[4 + wib];
    for (int i = lane; i < nq; i += 32) {
        const int64_t slot = qb + i;
        if (slot >= a.q_cap) {                                    // host grows the queue and re-runs
            atomicOr(a.d_err, ISB_DEV_ERR_QCAP);
            break;
        }
        const int q = q_list[i];
        const int tc = wib * 32 + (q >> 3);
        a.q_pos[slot] = W0 + q;
        a.q_E[slot] = tile4[q + (q >> 3)];
        // candidate segments of the site's column word, relative to the tile's first candidate (single-chunk tiles)
        a.q_cand[slot] = single ? ((uint32_t)s_cl[tc] | ((uint32_t)(s_ch[tc] - s_cl[tc]) << 16)) : K1F_NO_CAND;
    }
}

// ---- k2q_sites: the general SNV call on the queued sites (K2 at M = 1) + the linkage site slots -----------------------------
// One warp per tile, lane = one queued site, batches of 32: the reference's call_snv_site / update_snp_table / clonality in
// double precision (isb_k2_site.cuh), the re-drawn clonality, the raw_snp_table rows (one row allocation per batch).  The
// tile's anySNP sites get contiguous, position-ordered site slots (one atomic per tile) with their counts and candidate
// range: the input of k3f_site_rows and of the linkage back end.
struct k2q_args {
    int n_tiles;
    const int32_t *q_first, *q_cnt;
    const int32_t *q_pos;
    const int4 *q_E;
    const uint32_t *q_cand;
    int64_t q_cap;
    int32_t start, L;
    const unsigned long long *nmask;
    isb_k2_fuse k2;
    const int32_t *thr2;
    int n_lut, lut_default;
    unsigned long long *n_rows;
    int32_t *counts;
    unsigned int *d_err;
    int do_ld;
    int32_t *tile_first, *tile_cnt;
    int64_t sites_cap;
    int32_t *site_pos;
    int4 *site_counts;
    int4 *site_rec;                    // position, candidate range, allele set, split: all k3f_site_rows needs, one load
    int32_t n_splits;
    const int32_t *splits;
    const int32_t *tile_split;
    const int64_t *tile_lo;
    unsigned long long *n_sites;
};

#define K2Q_WARPS 4
__global__ void __launch_bounds__(K2Q_WARPS * 32) k2q_sites(k2q_args a)
{
    __shared__ uint16_t s_sites[K2Q_WARPS][K1F_TILE];             // index in the tile's queue range | allele set << 12
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int4 *counts4 = reinterpret_cast<int4 *>(a.counts);
    const bool full_counts = a.counts != nullptr && a.k2.full_counts;
    const int cov_r = a.k2.clonTR ? a.k2.cov_r : 0;
    for (int tile = blockIdx.x * K2Q_WARPS + wib; tile < a.n_tiles; tile += gridDim.x * K2Q_WARPS) {
        const int64_t first = a.q_first[tile];
        int cnt = a.q_cnt[tile];
        if (first + cnt > a.q_cap) cnt = first < a.q_cap ? (int)(a.q_cap - first) : 0;   // overflow: flagged by K1f, re-run follows
        int ns = 0;                                               // anySNP sites of the tile (warp-uniform)
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const bool on = b0 + lane < cnt;
            int4 E = make_int4(0, 0, 0, 0);
            int32_t p = 0;
            if (on) {
                p = __ldg(a.q_pos + first + b0 + lane);
                E = __ldg(a.q_E + first + b0 + lane);
            }
            int C[4] = {E.x, E.y, E.z, E.w};
            int r = 0;
            k2_m1_site s;
            s.T = 0; s.clon = CUDART_NAN_F; s.flags = 0u; s.is_row = false; s.i = 0; s.con = 0; s.thr = 0;
            if (on) {
                r = a.k2.ref[p];
                const int T = E.x + E.y + E.z + E.w;
                const int thr_T = (T >= a.k2.min_cov && T < a.n_lut) ? __ldg(a.thr2 + T) : a.lut_default;
                const bool nm0 = T == 0 && a.nmask && (a.nmask[p] & 1ull);
                s = k2_site_m1(C, r, nm0, thr_T, a.n_lut, a.lut_default, a.k2.min_cov, a.k2.min_freq);
                a.k2.clonT[p] = s.clon;
                a.k2.site_flags[p] = (uint8_t)s.flags;
                if (a.k2.clonTR)
                    a.k2.clonTR[p] = (cov_r > 0 && T >= cov_r) ? k2_rarefied_clon(C, T, cov_r, a.k2.seed, (int64_t)p + a.start, 0) : CUDART_NAN_F;
                if (full_counts) counts4[p] = E;
            }
            const bool row = on && s.is_row;
            const unsigned rm = __ballot_sync(ISB_FULL, row);
            if (rm) {                                             // one row allocation per batch of 32
                unsigned long long base_slot = 0;
                if (lane == 0) base_slot = atomicAdd(a.n_rows, (unsigned long long)__popc(rm));
                base_slot = __shfl_sync(ISB_FULL, base_slot, 0);
                const int64_t slot = (int64_t)base_slot + __popc(rm & ((1u << lane) - 1u));
                if (row && slot < a.k2.cap) k2_write_row_m1(a.k2.rows + slot, p + a.start, C, r, s, a.n_lut, a.k2.min_freq);
            }
            const bool st = on && (s.flags & ISB_SITE_ANYSNP);
            const unsigned sm = __ballot_sync(ISB_FULL, st);
            if (st) s_sites[wib][ns + __popc(sm & ((1u << lane) - 1u))] = (uint16_t)((b0 + lane) | ((s.flags & 0xFu) << 12));
            ns += __popc(sm);
        }
        if (!a.do_ld) continue;
        unsigned long long sb = 0;
        if (lane == 0) {
            if (ns) sb = atomicAdd(a.n_sites, (unsigned long long)ns);
            a.tile_first[tile] = (int32_t)(sb < (unsigned long long)INT_MAX ? sb : (unsigned long long)INT_MAX);
            a.tile_cnt[tile] = ns;
        }
        sb = __shfl_sync(ISB_FULL, sb, 0);
        __syncwarp();
        for (int i = lane; i < ns; i += 32) {
            const int64_t slot = (int64_t)sb + i;
            if (slot >= a.sites_cap) {                            // host grows the site slots and re-runs
                atomicOr(a.d_err, ISB_DEV_ERR_SITECAP);
                break;
            }
            const unsigned ent = s_sites[wib][i];
            const int64_t idx = first + (ent & 0x3ffu);
            const int32_t p = __ldg(a.q_pos + idx);
            const int64_t abs_pos = (int64_t)p + a.start;
            int sp = __ldg(a.tile_split + tile);                  // split of the site: walk forward from the tile's
            while (sp + 1 < a.n_splits && (int64_t)__ldg(a.splits + 2 * (sp + 1)) <= abs_pos) ++sp;
            if (!(sp >= 0 && abs_pos <= (int64_t)__ldg(a.splits + 2 * sp + 1))) sp = -1;
            a.site_pos[slot] = p;
            a.site_counts[slot] = __ldg(a.q_E + idx);
            // record: position | first candidate segment (low 32 bits) | its bits 32..39, allele set << 8, candidates << 12
            // (bit 31: range not known, search) | split
            const uint32_t cand = __ldg(a.q_cand + idx);
            int4 rec = make_int4(p, 0, (int)(0x80000000u | ((ent >> 12) << 8)), sp);
            if (cand != K1F_NO_CAND) {
                const int64_t glo = __ldg(a.tile_lo + tile) + (int64_t)(cand & 0xffffu);
                rec.y = (int)(uint32_t)(glo & 0xffffffffll);
                rec.z = (int)((uint32_t)((glo >> 32) & 0xff) | ((ent >> 12) << 8) | ((cand >> 16) << 12));
            }
            a.site_rec[slot] = rec;
        }
        __syncwarp();
    }
}

// ---- k3f_site_rows: linkage front end, one warp per site slot ---------------------------------------------------------------
// update_linked_reads (inStrain/profile/linkage.py:254-283): the (pair id, base) entries of a site, gathered from its
// candidate segments (table entries are contiguous: coalesced; one 32-byte sector of the stream per covering segment) and
// turned into the bit rows  any | ge1[b] | ge2[b]  over the site's pair-id window.
struct k3f_args {
    isb_reads_dev rd;
    int64_t n_pairs;
    int32_t start, L;
    int64_t sites_cap;
    const int4 *site_rec;              // position, candidate range (K1F_NO_CAND: search), allele set, split
    isb_site_meta *meta;
    int64_t *row_off;
    uint8_t *has2;
    uint32_t *rows;
    int64_t row_cap;
    const unsigned long long *n_sites;
    unsigned long long *row_words_total;
    int code_ids;                      // ids of the per-warp "pair id -> allele" byte map
    unsigned int *d_err;
};

// The exact row builder with the multiplicity planes (atomics on the row words): windows wider than the byte map, and
// sites where a pair has two entries (htslib's overlap quirk).  Rare: kept out of line.
__device__ __noinline__ void k3f_slow_rows(const int32_t *__restrict__ seg_start, const uint16_t *__restrict__ seg_len,
                                           const int64_t *__restrict__ seg_word, const int32_t *__restrict__ seg_pair,
                                           const uint32_t *__restrict__ words, int64_t n_words_stream, int64_t n_pairs,
                                           unsigned int *d_err, uint32_t *g_any, int wlo, int nw, int na, unsigned bases,
                                           int64_t glo, int nc, int64_t abs_pos, int lane)
{
    isb_site_meta m;
    m.ev_lo_rel = 0; m.wlo = wlo; m.nw = nw; m.split = 0;
    const int n_words = (1 + 2 * na) * nw;
    for (int i = lane; i < n_words; i += 32) g_any[i] = 0u;
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {
        const int64_t g = glo + i;
        const int32_t s = __ldg(seg_start + g);
        const int64_t j = abs_pos - (int64_t)s;
        if (j < 0 || j >= (int64_t)__ldg(seg_len + g)) continue;
        const int jn = (int)j + (s & 7);
        const int64_t wi = __ldg(seg_word + g) + (jn >> 3);
        if (wi < 0 || wi >= n_words_stream) continue;
        const uint32_t c = (__ldg(words + wi) >> ((jn & 7) << 2)) & 15u;
        if (!c) continue;
        const int b = __ffs((int)c) - 1, id = __ldg(seg_pair + g);
        if (id >= 0 && (int64_t)id < n_pairs && ((bases >> b) & 1u)) k3_row_set(g_any, m, na, bases, b, id, d_err);
    }
    __syncwarp();
}

#define K3F_WARPS 8
#ifndef K3F_MINB
#define K3F_MINB 4                    // resident blocks of 256 threads per SM the register budget is set for
#endif
__global__ void __launch_bounds__(K3F_WARPS * 32, K3F_MINB) k3f_site_rows(k3f_args a)
{
    extern __shared__ __align__(16) unsigned char k3f_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t *code = k3f_smem + (size_t)wib * a.code_ids;          // pair id -> allele code of the warp's site
    uint32_t *code4 = reinterpret_cast<uint32_t *>(code);
    const unsigned long long ns_dev = *a.n_sites;
    const int64_t S = (int64_t)(ns_dev < (unsigned long long)a.sites_cap ? ns_dev : (unsigned long long)a.sites_cap);
    const int maxlen = a.rd.max_seg_len;
    const int64_t k0 = (int64_t)blockIdx.x * K3F_WARPS + wib, k_step = (int64_t)gridDim.x * K3F_WARPS;
    int4 rec_next = k0 < S ? __ldg(a.site_rec + k0) : make_int4(0, 0, 0, 0);
    for (int64_t k = k0; k < S; k += k_step) {
        const int4 rec = rec_next;
        if (k + k_step < S) rec_next = __ldg(a.site_rec + k + k_step);   // the next site's record travels while this one is built
        const int32_t p = rec.x;
        const uint32_t rz = (uint32_t)rec.z;
        const int64_t abs_pos = (int64_t)p + a.start;
        const unsigned bases = (rz >> 8) & 0xFu;
        const int na = __popc(bases);
        int64_t glo = (int64_t)(uint32_t)rec.y | ((int64_t)(rz & 0xffu) << 32);
        int nc = (int)((rz >> 12) & 0x1fffu);
        if (rz & 0x80000000u) {                                    // tile staged in several chunks (deep coverage): search
            const int tile = p / K1F_TILE;
            const int64_t lo = __ldg(a.rd.tile_lo + tile), hi = __ldg(a.rd.tile_hi + tile);
            glo = isb_lower_bound(a.rd.seg_start, lo, hi, abs_pos - maxlen + 1);
            nc = (int)(isb_lower_bound(a.rd.seg_start, glo, hi, abs_pos + 1) - glo);
        }
        if (glo < 0 || glo + nc > a.rd.n_segs) nc = 0;
        // Candidates in groups of 4 x 32: the table loads of a whole group are issued together, then its word loads (two
        // memory latencies per group instead of three per 32 candidates); group 0 -- all of them up to ~120x coverage --
        // stays decoded in registers for the second pass.
        auto group = [&](int g0, int (&b4)[4], int (&id4)[4]) {
            int32_t s4[4], n4[4], pid[4];
            int64_t w4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = g0 + u * 32 + lane;
                s4[u] = 0; n4[u] = 0; pid[u] = -1; w4[u] = 0;
                if (i < nc) {
                    const int64_t g = glo + i;
                    s4[u] = __ldg(a.rd.seg_start + g);
                    n4[u] = (int)__ldg(a.rd.seg_len + g);
                    w4[u] = __ldg(a.rd.seg_word + g);
                    pid[u] = __ldg(a.rd.seg_pair + g);
                }
            }
            uint32_t wd[4];
            int sh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t j = abs_pos - (int64_t)s4[u];
                wd[u] = 0u; sh[u] = 0;
                if (j >= 0 && j < (int64_t)n4[u]) {
                    const int jn = (int)j + (s4[u] & 7);              // position-aligned stream: nibble index in the segment's words
                    const int64_t wi = w4[u] + (jn >> 3);
                    if (wi >= 0 && wi < a.rd.n_words) wd[u] = __ldg(a.rd.words + wi);
                    sh[u] = (jn & 7) << 2;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t c = (wd[u] >> sh[u]) & 15u;
                b4[u] = -1; id4[u] = 0;
                if (c) {
                    const int b = __ffs((int)c) - 1;                   // one-hot A,C,T,G
                    if (pid[u] < 0 || (int64_t)pid[u] >= a.n_pairs) atomicOr(a.d_err, ISB_DEV_ERR_SEG);
                    else if ((bases >> b) & 1u) { b4[u] = b; id4[u] = pid[u]; }
                }
            }
        };
        int gb[4], gid[4];                                         // group 0
        group(0, gb, gid);
        int idmin = INT_MAX, idmax = -1, n_ent = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (gb[u] >= 0) { idmin = min(idmin, gid[u]); idmax = max(idmax, gid[u]); ++n_ent; }
        for (int g0 = 128; g0 < nc; g0 += 128) {                   // deep coverage: further groups
            int b4[4], id4[4];
            group(g0, b4, id4);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (b4[u] >= 0) { idmin = min(idmin, id4[u]); idmax = max(idmax, id4[u]); ++n_ent; }
        }
        idmin = k1f_warp_min(idmin);
        idmax = k1f_warp_max(idmax);
        isb_site_meta m;
        m.split = rec.w;
        m.ev_lo_rel = 0;
        m.wlo = idmax >= 0 ? (idmin >> 5) : 0;
        m.nw = idmax >= 0 ? (idmax >> 5) - (idmin >> 5) + 1 : 0;
        const int n_words = (1 + 2 * na) * m.nw;
        // Row storage: slot k owns the fixed region [k * ROW_SLOT, + ROW_SLOT); only rows wider than that (deep coverage)
        // are allocated with an atomic behind the sites_cap fixed regions.
        unsigned long long off = (unsigned long long)k * ISB_K3_ROW_SLOT;
        if (n_words > ISB_K3_ROW_SLOT) {
            if (lane == 0) off = (unsigned long long)a.sites_cap * ISB_K3_ROW_SLOT + atomicAdd(a.row_words_total, (unsigned long long)n_words);
            off = __shfl_sync(ISB_FULL, off, 0);
        }
        if ((int64_t)(off + n_words) > a.row_cap) {               // host grows the row storage and re-runs
            if (lane == 0) atomicOr(a.d_err, ISB_DEV_ERR_ROWBUF);
            m.nw = 0;
        }
        if (lane == 0) {
            a.meta[k] = m;
            a.row_off[k] = (int64_t)off;
        }
        bool dup = false;
        if (m.nw > 0) {
            uint32_t *g_any = a.rows + off;
            n_ent = k1f_warp_sum(n_ent);
            int n_bits = 0;
            if (m.nw * 32 <= a.code_ids) {
                // Bit rows by BALLOT.  The site's entries are scattered into a byte map "pair id -> allele code" (ids are
                // distinct unless a pair entered the site twice: plain stores); then lane l owns bit l of every row word:
                // word w of the `any` row is ballot(code[32 w + l] != 0), word w of allele row r is ballot(code == r's
                // base).  A pair seen twice (htslib's overlap quirk) shows up as fewer set bits than entries: then the
                // exact slow path rebuilds the rows with the multiplicity planes.
                for (int i = lane; i < m.nw * 8; i += 32) code4[i] = 0u;
                __syncwarp();
                const int id0 = m.wlo << 5;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (gb[u] >= 0) code[gid[u] - id0] = (uint8_t)(gb[u] + 1);
                for (int g0 = 128; g0 < nc; g0 += 128) {
                    int b4[4], id4[4];
                    group(g0, b4, id4);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (b4[u] >= 0) code[id4[u] - id0] = (uint8_t)(b4[u] + 1);
                }
                __syncwarp();
                int b_of[4] = {0, 0, 0, 0};                        // base (+ 1) of allele row r
                {
                    int r = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((bases >> b) & 1u) b_of[r++] = b + 1;
                }
                for (int w = 0; w < m.nw; ++w) {
                    const int c = code[w * 32 + lane];
                    const unsigned any_w = __ballot_sync(ISB_FULL, c != 0);
                    unsigned mine = lane == 0 ? any_w : 0u;        // lane r + 1 keeps allele row r's word; lanes na + 1 .. 2 na: 0
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r >= na) break;                        // warp-uniform
                        const unsigned row_w = __ballot_sync(ISB_FULL, c == b_of[r]);     // b_of > 0: empty ids never match
                        if (lane == r + 1) mine = row_w;
                    }
                    n_bits += __popc(any_w);
                    if (lane <= 2 * na) g_any[(size_t)lane * m.nw + w] = mine;
                }
                __syncwarp();
                dup = n_bits != n_ent;
                if (dup)
                    k3f_slow_rows(a.rd.seg_start, a.rd.seg_len, a.rd.seg_word, a.rd.seg_pair, a.rd.words, a.rd.n_words, a.n_pairs, a.d_err,
                                  g_any, m.wlo, m.nw, na, bases, glo, nc, abs_pos, lane);
            } else {
                k3f_slow_rows(a.rd.seg_start, a.rd.seg_len, a.rd.seg_word, a.rd.seg_pair, a.rd.words, a.rd.n_words, a.n_pairs, a.d_err,
                              g_any, m.wlo, m.nw, na, bases, glo, nc, abs_pos, lane);
                // a pair with two entries sets a bit of `any` that is already set: detect it by counting
                for (int i = lane; i < m.nw; i += 32) n_bits += __popc(g_any[i]);
                dup = k1f_warp_sum(n_bits) != n_ent;
            }
        }
        if (lane == 0) a.has2[k] = dup ? 1 : 0;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------

static size_t k1f_smem_bytes(int seg_cap, bool m1, bool fuse, int Mg)
{
    size_t b = ((((size_t)seg_cap * (m1 ? 8 : 9)) + 15) & ~(size_t)15) + 16;   // staged table + the zero word
    if (!m1) b += (size_t)Mg * 8 * K1F_THREADS * 4;
    else b += sizeof(int4) * K1F_WARPS * K1F_TILE4;                // the count quads of the tile
    if (fuse) b += 4 * 2 * K1F_THREADS + 4 * 16 + K1F_WARPS * 256;
    return b;
}

// L2 prefetch distance in tiles: 0 = every block prefetches its own tile only (the default); ISB_K1F_PF = n > 0 makes a block
// also ask for the words and the table of the tile n later.  Measured (B200, 2e7 positions at 100x): M = 1: 0 -> 0.428 ms,
// 74 -> 0.424, 148 -> 0.438, 296 -> 0.468; M = 15: 0 -> 3.76 ms, 74 -> 4.49.  The loads are not waiting for DRAM any more;
// data fetched further ahead only displaces data that is still needed.
static int k1f_pf_ahead(const isb_ctx *ctx)
{
    static const int env = getenv("ISB_K1F_PF") ? atoi(getenv("ISB_K1F_PF")) : -1;
    (void)ctx;
    return env >= 0 ? env : 0;
}

// tile bounds + staging capacity of a batch
static int k1f_prepare(isb_ctx *ctx, isb_reads_dev *rd, int32_t start, int32_t L, int *seg_cap, const int32_t *splits = nullptr,
                       int n_splits = 0, int32_t **tile_split_out = nullptr)
{
    cudaStream_t st = ctx->stream;
    if (rd->max_seg_len < 1 || rd->max_seg_len > K1F_MAXLEN)
        return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: max_seg_len must be in [1, 256]");
    if (((uintptr_t)rd->words & 15) != 0) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: words must be 16-byte aligned");
    if (start & 7) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: start must be a multiple of 8 (the stream is position-aligned)");
    const int n_tiles = (L + K1F_TILE - 1) / K1F_TILE;
    int rc;
    if ((rc = isb_ensure(ctx, SL_RD_BOUNDS, sizeof(int64_t) * 5 * (size_t)n_tiles))) return rc;
    int64_t *tile_lo = (int64_t *)ctx->buf[SL_RD_BOUNDS].p, *tile_hi = tile_lo + n_tiles;
    int64_t *tile_wlo = tile_hi + n_tiles, *tile_whi = tile_wlo + n_tiles;
    int32_t *tile_split = tile_split_out ? (int32_t *)(tile_whi + n_tiles) : nullptr;
    k1f_tile_bounds<<<(n_tiles + 7) / 8, 256, 0, st>>>(rd->seg_start, rd->seg_word, rd->n_segs, rd->n_words, start, n_tiles,
                                                       rd->max_seg_len, tile_lo, tile_hi, tile_wlo, tile_whi, splits, n_splits, tile_split);
    ISB_LAUNCH_CHECK();
    if (tile_split_out) *tile_split_out = tile_split;
    rd->n_tiles = n_tiles;
    rd->tile_lo = tile_lo;
    rd->tile_hi = tile_hi;
    rd->tile_wlo = tile_wlo;
    rd->tile_whi = tile_whi;
    // staging capacity: what a tile holds on average + 15 % + 48 (the Poisson spread of ~800 segments is 3.5 %), capped
    int64_t cap = rd->n_segs > 0 ? (int64_t)((double)rd->n_segs / L * (K1F_TILE + rd->max_seg_len) * 1.12) + 40 : 64;
    if (cap < 64) cap = 64;
    if (cap > K1F_SEG_CAP_MAX) cap = K1F_SEG_CAP_MAX;
    *seg_cap = (int)((cap + 3) & ~(int64_t)3);
    return ISB_OK;
}

int isb_k1f_pileup_launch(isb_ctx *ctx, isb_reads_dev *rd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                          int M, int32_t *counts, unsigned long long *nmask)
{
    cudaStream_t st = ctx->stream;
    if (L <= 0) return ISB_OK;
    if (M > 1 && !pair_mm) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: pair_mm is required when M > 1");
    int rc, seg_cap = 0;
    if ((rc = k1f_prepare(ctx, rd, start, L, &seg_cap))) return rc;
    k1f_args a;
    memset(&a, 0, sizeof(a));
    a.rd = *rd; a.pair_mm = pair_mm; a.n_pairs = n_pairs; a.start = start; a.L = L; a.M = M; a.counts = counts;
    a.d_err = ctx->d_err; a.seg_cap = seg_cap; a.pf_ahead = k1f_pf_ahead(ctx);
    const int groups = M == 1 ? 1 : (M + K1F_LEVELS - 1) / K1F_LEVELS;
    const int Mg = M == 1 ? 0 : (M < K1F_LEVELS ? M : K1F_LEVELS);
    const size_t smem = k1f_smem_bytes(seg_cap, M == 1, false, Mg);
    static bool attr_m1[64] = {false}, attr_mm[64] = {false};      // function attributes are per device
    if (nmask) ISB_CUDA(cudaMemsetAsync(nmask, 0, sizeof(unsigned long long) * (size_t)L, st));
    if (M == 1) {
        if (!attr_m1[ctx->device & 63])
            ISB_CUDA(cudaFuncSetAttribute(k1f_pileup<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_m1[ctx->device & 63] = true;
        k1f_pileup<true, false><<<rd->n_tiles, K1F_THREADS, smem, st>>>(a);
        ISB_LAUNCH_CHECK();
    } else {
        if (!attr_mm[ctx->device & 63])
            ISB_CUDA(cudaFuncSetAttribute(k1f_pileup<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_mm[ctx->device & 63] = true;
        k1f_pileup<false, false><<<dim3(rd->n_tiles, groups), K1F_THREADS, smem, st>>>(a);
        ISB_LAUNCH_CHECK();
    }
    if (nmask && rd->n_nev > 0) return isb_k1r_n_events_launch(ctx, rd->n_nev, rd->nev_pos, rd->nev_pair, pair_mm, n_pairs, start, L, M, nmask);
    return ISB_OK;
}

// Fused path: K1f<fused> -> k2q_sites (-> k3f_site_rows -> the linkage back end when `ld` is given).  Counters: [0] SNV
// rows, [1] LD rows, [2] sites, [3] linked site pairs, [4] overflow row words, [5] queued sites; nothing is read back here.
int isb_k1f_profile_launch(isb_ctx *ctx, isb_reads_dev *rd, int64_t n_pairs, int32_t start, int32_t L,
                           unsigned long long *nmask, const isb_k2_fuse *fuse, const isb_k1f_linkage *ld)
{
    cudaStream_t st = ctx->stream;
    int rc, seg_cap = 0;
    ISB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 6 * sizeof(unsigned long long), st));
    if (L <= 0) return ISB_OK;
    int32_t *tile_split = nullptr;
    if ((rc = k1f_prepare(ctx, rd, start, L, &seg_cap, ld ? ld->splits : nullptr, ld ? ld->n_splits : 0, ld ? &tile_split : nullptr))) return rc;
    if ((rc = isb_k2_prepare(ctx, fuse->min_freq))) return rc;
    if (nmask) {                                                  // N events only make level 0 a key of MMcounts
        ISB_CUDA(cudaMemsetAsync(nmask, 0, sizeof(unsigned long long) * (size_t)L, st));
        if ((rc = isb_k1r_n_events_launch(ctx, rd->n_nev, rd->nev_pos, rd->nev_pair, nullptr, n_pairs, start, L, 1, nmask))) return rc;
    }
    const int n_tiles = rd->n_tiles;
    // Site queue: sites with a second base (mostly sequencing errors: their share grows with the coverage) and, with
    // min_cov < 1, every position.  First guess from the mean coverage, grown on demand (isb_k1f_grow);
    // ISB_K1F_QUEUE_INIT forces a small first guess (tests of the re-run path).
    static const int64_t queue_init = getenv("ISB_K1F_QUEUE_INIT") ? atoll(getenv("ISB_K1F_QUEUE_INIT")) : 0;
    {
        const double cov = (double)rd->n_words * 8.0 / (double)L;
        double share = fuse->min_cov < 1 ? 1.0 : 0.10 + cov / 400.0;
        if (share > 1.0) share = 1.0;
        int64_t want = queue_init > 0 ? queue_init : (int64_t)((double)L * share) + 4096;
        if (want > (int64_t)L + 32) want = (int64_t)L + 32;
        if (ctx->queue_cap < want) ctx->queue_cap = want;
    }
    const int64_t qcap = ctx->queue_cap;
    if ((rc = isb_ensure(ctx, SL_Q_TILE, sizeof(int32_t) * 2 * (size_t)n_tiles))) return rc;
    if ((rc = isb_ensure(ctx, SL_Q_POS, sizeof(int32_t) * (size_t)qcap))) return rc;
    if ((rc = isb_ensure(ctx, SL_Q_E, sizeof(int4) * (size_t)qcap))) return rc;
    if ((rc = isb_ensure(ctx, SL_Q_CAND, sizeof(uint32_t) * (size_t)qcap))) return rc;
    k1f_args a;
    memset(&a, 0, sizeof(a));
    a.rd = *rd; a.n_pairs = n_pairs; a.start = start; a.L = L; a.M = 1; a.d_err = ctx->d_err; a.seg_cap = seg_cap; a.pf_ahead = k1f_pf_ahead(ctx);
    a.k2 = *fuse;
    a.thr2 = ctx->d_thr2; a.n_lut = ctx->n_lut;
    a.q_first = (int32_t *)ctx->buf[SL_Q_TILE].p;
    a.q_cnt = a.q_first + n_tiles;
    a.q_pos = (int32_t *)ctx->buf[SL_Q_POS].p;
    a.q_E = (int4 *)ctx->buf[SL_Q_E].p;
    a.q_cand = (uint32_t *)ctx->buf[SL_Q_CAND].p;
    a.q_cap = qcap;
    a.n_queue = ctx->d_counters + 5;

    k2q_args b;
    memset(&b, 0, sizeof(b));
    b.n_tiles = n_tiles; b.q_first = a.q_first; b.q_cnt = a.q_cnt; b.q_pos = a.q_pos; b.q_E = a.q_E; b.q_cand = a.q_cand; b.q_cap = qcap;
    b.start = start; b.L = L; b.nmask = nmask; b.k2 = *fuse; b.thr2 = ctx->d_thr2; b.n_lut = ctx->n_lut; b.lut_default = ctx->lut_default;
    b.n_rows = ctx->d_counters + 0; b.counts = nullptr; b.d_err = ctx->d_err;

    k3f_args c;
    memset(&c, 0, sizeof(c));
    isb_k3_tiles ts;
    memset(&ts, 0, sizeof(ts));
    if (ld) {
        // site slots: 1 % SNV sites with headroom, grown on demand (isb_k1f_grow); ISB_K1F_SITES_INIT forces a small
        // first guess (tests of the re-run path)
        static const int64_t sites_init = getenv("ISB_K1F_SITES_INIT") ? atoll(getenv("ISB_K1F_SITES_INIT")) : 0;
        const int64_t want = sites_init > 0 ? sites_init : (int64_t)L / 72 + 4096;
        if (ctx->sites_cap < want) ctx->sites_cap = want;
        const int64_t cap = ctx->sites_cap;
        if ((rc = isb_ensure(ctx, SL_TILE_SITES, sizeof(int32_t) * 2 * (size_t)n_tiles))) return rc;
        if ((rc = isb_ensure(ctx, SL_SITE_POS, sizeof(int32_t) * (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_SITE_CAND, sizeof(int4) * (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_SITE_META, sizeof(isb_site_meta) * (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_ROW_OFF, sizeof(int64_t) * (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_HAS2, (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_SITE_COUNTS, sizeof(int4) * (size_t)cap))) return rc;
        if ((rc = isb_ensure(ctx, SL_ROWS, sizeof(uint32_t) * ((size_t)cap * (ISB_K3_ROW_SLOT + 8) + 64)))) return rc;
        b.do_ld = 1;
        b.tile_first = (int32_t *)ctx->buf[SL_TILE_SITES].p;
        b.tile_cnt = b.tile_first + n_tiles;
        b.sites_cap = cap;
        b.site_pos = (int32_t *)ctx->buf[SL_SITE_POS].p;
        b.site_counts = (int4 *)ctx->buf[SL_SITE_COUNTS].p;
        b.site_rec = (int4 *)ctx->buf[SL_SITE_CAND].p;
        b.n_splits = ld->n_splits; b.splits = ld->splits; b.tile_split = tile_split; b.tile_lo = rd->tile_lo;
        b.n_sites = ctx->d_counters + 2;
        c.rd = *rd; c.n_pairs = n_pairs; c.start = start; c.L = L;
        c.sites_cap = cap; c.site_rec = b.site_rec;
        c.meta = (isb_site_meta *)ctx->buf[SL_SITE_META].p;
        c.row_off = (int64_t *)ctx->buf[SL_ROW_OFF].p;
        c.has2 = (uint8_t *)ctx->buf[SL_HAS2].p;
        c.rows = (uint32_t *)ctx->buf[SL_ROWS].p;
        c.row_cap = (int64_t)(ctx->buf[SL_ROWS].cap / sizeof(uint32_t));
        c.n_sites = ctx->d_counters + 2;
        c.row_words_total = ctx->d_counters + 4;
        c.d_err = ctx->d_err;
        // pair-id window of a site ~ the pairs whose first mate starts within one fragment length before it: sized from the
        // pair density with headroom, so that the ballot row builder (not the atomic fallback) serves deep coverage too
        int64_t ids = (int64_t)((double)n_pairs / (L > 0 ? L : 1) * 700.0 * 1.5) + 64;
        ids = (ids + 255) / 256 * 256;
        if (ids < K1F_CODE_IDS_MIN) ids = K1F_CODE_IDS_MIN;
        if (ids > K1F_CODE_IDS_MAX) ids = K1F_CODE_IDS_MAX;
        c.code_ids = (int)ids;
        ts.n_tiles = n_tiles; ts.tile_first = b.tile_first; ts.tile_cnt = b.tile_cnt; ts.sites_cap = cap;
        ts.site_pos = b.site_pos; ts.meta = c.meta; ts.row_off = c.row_off; ts.has2 = c.has2; ts.site_counts = b.site_counts;
        ts.rows = c.rows;
    }
    const size_t smem = k1f_smem_bytes(seg_cap, true, true, 0);
    static bool attr[64] = {false};
    if (!attr[ctx->device & 63])
        ISB_CUDA(cudaFuncSetAttribute(k1f_pileup<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr[ctx->device & 63] = true;
    int ts0 = isb_time_begin(ctx, 0);
    k1f_pileup<true, true><<<n_tiles, K1F_THREADS, smem, st>>>(a);
    ISB_LAUNCH_CHECK();
    isb_time_end(ctx, ts0);
    int ts1 = isb_time_begin(ctx, 1);
    {
        // one warp per tile, no grid-stride cap: with ~1.3 waves of capped blocks the last wave left half the SMs idle on
        // small batches (the per-rank share of a strong-scaling run)
        k2q_sites<<<(n_tiles + K2Q_WARPS - 1) / K2Q_WARPS, K2Q_WARPS * 32, 0, st>>>(b);
        ISB_LAUNCH_CHECK();
    }
    isb_time_end(ctx, ts1);
    if (ld) {
        int ts2 = isb_time_begin(ctx, 2);
        k3f_site_rows<<<ctx->sm_count * 8, K3F_WARPS * 32, (size_t)K3F_WARPS * c.code_ids, st>>>(c);
        ISB_LAUNCH_CHECK();
        rc = isb_k3_backend_tiles(ctx, rd, &ts, n_pairs, start, L, nmask, fuse->site_flags, ld->n_splits, ld->splits, ld->min_snp,
                                  ld->rows, ld->cap);
        isb_time_end(ctx, ts2);
        if (rc) return rc;
    }
    return ISB_OK;
}

// After the counters / error word of a fused run are on the host: grow what was too small.  *again = true -> re-run.
int isb_k1f_grow(isb_ctx *ctx, bool *again)
{
    *again = false;
    const unsigned e = *ctx->h_err;
    if (e & ~(ISB_DEV_ERR_SITECAP | ISB_DEV_ERR_ROWBUF | ISB_DEV_ERR_QCAP)) return ISB_OK;   // a real error: the caller reports it
    const int64_t n_sites = (int64_t)ctx->h_counters[2], n_listed = (int64_t)ctx->h_counters[3], n_queue = (int64_t)ctx->h_counters[5];
    if (n_queue > ctx->queue_cap) {
        ctx->queue_cap = n_queue + n_queue / 8 + 1024;
        *again = true;
    }
    if (n_sites > ctx->sites_cap) {
        ctx->sites_cap = n_sites + n_sites / 8 + 1024;
        *again = true;
    }
    if (e & ISB_DEV_ERR_ROWBUF) {
        const size_t need = (size_t)ctx->sites_cap * ISB_K3_ROW_SLOT + (size_t)ctx->h_counters[4];
        int rc = isb_ensure(ctx, SL_ROWS, sizeof(uint32_t) * (need + need / 8 + 1024));
        if (rc) return rc;
        *again = true;
    }
    const int64_t cap_pairs = (int64_t)(ctx->buf[SL_PAIRS].cap / (2 * sizeof(int32_t)));
    if (n_listed > cap_pairs) {
        int rc = isb_ensure(ctx, SL_PAIRS, 2 * sizeof(int32_t) * (size_t)(n_listed + n_listed / 8 + 1024));
        if (rc) return rc;
        *again = true;
    }
    if (*again) {
        cudaError_t ce = cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream);
        if (ce != cudaSuccess) return isb_fail(ctx, ISB_ERR_CUDA, cudaGetErrorString(ce));
    }
    return ISB_OK;
}
