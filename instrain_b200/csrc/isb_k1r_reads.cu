// isb_k1r_reads.cu -- K1r: pileup counts straight from READ-MAJOR aligned segments, sm_100a.
//
// Same result as K1 (isb_k1_pileup.cu): counts[position][mm][A,C,T,G] (+ nmask) of samfile.pileup(...) column iteration
// (inStrain/profile/profile_utilities.py:150-153) + get_base_counts_mm (:268-286) -- but the input is one 4-bit code per
// aligned base, stored once per READ (include/instrain_b200.h, isb_reads_batch), not one 10-byte event per column entry.
// The pileup is a transposition (reads x offsets -> positions), done here on the fly without atomics:
//
//   * a block owns a tile of K1R_TILE = 1024 positions, a thread 8 consecutive positions (one 32-bit word of nibbles);
//   * the segments that can touch the tile (seg_start in (tile_first - max_seg_len, tile_end), a contiguous range of the
//     start-sorted segment table: k1r_tile_bounds) are staged chunk by chunk in shared memory -- start / length / word
//     offset by the threads, the nibble words of the chunk (contiguous in the stream) by ONE TMA 1-D bulk copy;
//   * each thread binary-searches its own candidate range in the chunk and, per candidate, loads the ONE word of the
//     segment that holds its 8 positions (the stream is position-aligned: word k of a segment covers the batch
//     coordinates [8 (start / 8 + k), + 8)).  Codes of non-events and the nibbles outside the segment are 0, so no
//     per-nibble bounds checks are needed;
//   * counting is bit-sliced.  The codes are one-hot (A=1, C=2, T=4, G=8), so the 32 bits of that register are the 32
//     (position, base) indicator bits of the candidate.  M = 1: they are added into eight VERTICAL counter planes
//     (plane j = bit j of 32 independent counters) with a Harley-Seal carry-save tree, 8 candidates per block:
//     7 CSAs + a 5-plane ripple = 24 logic ops per 8 candidates, instead of ~11 per candidate for per-base masks
//     and horizontal adds.  The planes are turned into integers once per <= 248 candidates.
//     M > 1 keeps horizontal 8-bit counters per (mm level, base) in shared memory, laid out [word][thread] so that
//     every lane always hits its own bank, and flushes them to the thread's own cells of `counts` (plain stores).
//
// HBM traffic: 0.5 B per aligned base + ~22 B per segment in, 16*M B per position out (vs 6-10 B per event in for K1).
// The kernel is bound by issue slots / shared-memory bandwidth, not by HBM (profiles/README.md).
#include "isb_common.cuh"
#include "isb_bitslice.cuh"
#include <climits>
#include <cstdlib>

#define K1R_THREADS (K1R_TILE / 8)     // one thread per 8 positions (one word of nibbles)
#define K1R_MAXLEN 256                 // hard cap of max_seg_len
#define K1R_LEVELS 32                  // mm levels per pass of the M > 1 kernel (shared-memory accumulators)
#define K1R_WPS (K1R_MAXLEN / 8 + 1)   // worst-case words per segment incl. its separator
#define K1R_STAGE_IT 9                 // segment-table elements per thread per chunk: seg_cap <= 9 * 128

struct k1r_args {
    isb_reads_dev rd;
    const uint8_t *pair_mm;
    int64_t n_pairs;
    int32_t start, L;
    int M;
    int seg_cap;                       // segments staged per chunk (shared-memory budget / bytes per segment)
    int words_cap;                     // words staged per chunk: seg_cap * (max_seg_len / 8 + 2) + 8
    int32_t *counts;
    unsigned long long *nmask;
    unsigned int *d_err;
};

// tile t covers relative positions [t*K1R_TILE, (t+1)*K1R_TILE): its candidate segment range and their word range
__global__ void __launch_bounds__(256)
k1r_tile_bounds(const int32_t *__restrict__ seg_start, const uint16_t *__restrict__ seg_len,
                const int64_t *__restrict__ seg_word, int64_t n_segs, int32_t start, int n_tiles, int max_seg_len,
                int64_t *__restrict__ tile_lo, int64_t *__restrict__ tile_hi, int64_t *__restrict__ tile_wlo,
                int64_t *__restrict__ tile_whi)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t first = (int64_t)start + (int64_t)t * K1R_TILE;
    const int64_t lo = isb_lower_bound(seg_start, 0, n_segs, first - max_seg_len + 1);
    const int64_t hi = isb_lower_bound(seg_start, lo, n_segs, first + K1R_TILE);
    tile_lo[t] = lo;
    tile_hi[t] = hi;
    tile_wlo[t] = hi > lo ? seg_word[lo] - 1 : 0;
    tile_whi[t] = hi > lo ? seg_word[hi - 1] + (((seg_start[hi - 1] & 7) + seg_len[hi - 1] + 7) >> 3) + 1 : 0;
}

// M > 1: write (or add, once counts hold a partial sum) the thread's shared 8-bit counters to its cells of `counts`
__device__ __forceinline__ void k1r_flush_levels(const k1r_args &a, uint32_t *s_acc, int t, int Mg, int m_base, int32_t P,
                                                 bool add, bool clear)
{
    int4 *c4 = reinterpret_cast<int4 *>(a.counts);
    for (int m = 0; m < Mg; ++m) {
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            w8[j] = s_acc[(size_t)(m * 8 + j) * K1R_THREADS + t];
            if (clear) s_acc[(size_t)(m * 8 + j) * K1R_THREADS + t] = 0u;
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

template <bool kM1>
__global__ void __launch_bounds__(K1R_THREADS) k1r_pileup(k1r_args a)
{
    extern __shared__ __align__(128) unsigned char k1r_smem_raw[];
    uint32_t *s_words = reinterpret_cast<uint32_t *>(k1r_smem_raw);
    // one packed word per segment (a 32-bit load has half the bank conflicts of a 64-bit one when every lane follows its
    // own segment): bits 11.. = staged index of the word that covers tile column 0, + 160; bits 0..10 = tile-relative end + 256
    uint32_t *s_meta = s_words + a.words_cap;
    int32_t *s_start = reinterpret_cast<int32_t *>(s_meta + a.seg_cap);  // relative start (sorted): the candidate search key
    uint8_t *s_mm = reinterpret_cast<uint8_t *>(s_start + a.seg_cap);
    uint64_t *bar = reinterpret_cast<uint64_t *>(k1r_smem_raw + (((size_t)a.words_cap * 4 + (size_t)a.seg_cap * 9 + 7) & ~(size_t)7));
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(bar) + 16);

    const int t = threadIdx.x;
    const int tile = blockIdx.x;
    const int32_t T0 = tile * K1R_TILE;
    const int32_t P = T0 + t * 8;                               // first of the thread's 8 positions (relative)
    const bool active = P < a.L;
    const int maxlen = a.rd.max_seg_len;
    const int64_t lo = a.rd.tile_lo[tile], hi = a.rd.tile_hi[tile];
    const int m_base = kM1 ? 0 : (int)blockIdx.y * K1R_LEVELS;
    const int Mg = kM1 ? 1 : min(K1R_LEVELS, a.M - m_base);

    int c[8][4];                                                  // M = 1: the thread's 32 counters
    uint32_t pl[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};            // M = 1: vertical counter planes (weights 1 .. 128)
    int n8 = 0;                                                   // candidates since the last flush (counters hold <= 255)
    bool spilled = false;                                         // M > 1: counts already hold a partial sum
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) c[k][b] = 0;
    if (!kM1)
        for (int w = 0; w < Mg * 8; ++w) s_acc[w * K1R_THREADS + t] = 0u;
    if (t == 0) {
        isb_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned parity = 0;
    unsigned err = 0;

    for (int64_t c0 = lo; c0 < hi; c0 += a.seg_cap) {
        const int nc = (int)min((int64_t)a.seg_cap, hi - c0);
        __syncthreads();                                          // previous chunk consumed (first pass: mbarrier init visible)
        const int64_t last = c0 + nc - 1;
        int64_t w_first, w_end;
        if (c0 == lo && last == hi - 1) {                         // the usual case: the whole tile in one chunk
            w_first = a.rd.tile_wlo[tile];
            w_end = a.rd.tile_whi[tile];
        } else {
            w_first = __ldg(a.rd.seg_word + c0) - 1;              // the separator in front of the chunk's first segment
            w_end = __ldg(a.rd.seg_word + last) + (((__ldg(a.rd.seg_start + last) & 7) + __ldg(a.rd.seg_len + last) + 7) >> 3) + 1;
        }
        const int64_t wb = w_first & ~(int64_t)3;                 // 16-byte aligned source
        const int64_t wn = ((w_end - wb) + 3) & ~(int64_t)3;
        if (w_first < 0 || wn <= 0 || wn > a.words_cap || wb + wn > a.rd.n_words) {   // layout rules violated (uniform)
            err |= ISB_DEV_ERR_SEG;
            continue;
        }
        if (t == 0) {
            isb_mbar_expect_tx(bar, (unsigned)wn * 4u);
            isb_bulk_g2s(s_words, a.rd.words + wb, (unsigned)wn * 4u, bar);
        }
        // Segment table of the chunk -> shared memory.  All global loads of the (up to K1R_STAGE_IT) elements a thread
        // handles are issued before the first is used, so the block pays ONE memory latency here instead of one per
        // element (ncu: the per-element version spent 36 % of its cycles on this loop's long-scoreboard stalls).
        {
            int32_t r_s[K1R_STAGE_IT], r_prev[K1R_STAGE_IT], r_pid[K1R_STAGE_IT];
            int r_n[K1R_STAGE_IT];
            int64_t r_w[K1R_STAGE_IT];
#pragma unroll
            for (int k = 0; k < K1R_STAGE_IT; ++k) {
                const int i = t + k * K1R_THREADS;
                const int64_t g = c0 + i;
                r_s[k] = 0; r_prev[k] = INT_MIN; r_n[k] = 1; r_w[k] = wb + 1; r_pid[k] = 0;
                if (i < nc) {
                    r_s[k] = __ldg(a.rd.seg_start + g);
                    r_n[k] = __ldg(a.rd.seg_len + g);
                    r_w[k] = __ldg(a.rd.seg_word + g);
                    if (g > 0) r_prev[k] = __ldg(a.rd.seg_start + g - 1);
                    if (!kM1) r_pid[k] = __ldg(a.rd.seg_pair + g);
                }
            }
            int r_mm[K1R_STAGE_IT];
            if (!kM1) {
#pragma unroll
                for (int k = 0; k < K1R_STAGE_IT; ++k) {
                    r_mm[k] = 255;
                    if (t + k * K1R_THREADS < nc && r_pid[k] >= 0 && (int64_t)r_pid[k] < a.n_pairs)
                        r_mm[k] = __ldg(a.pair_mm + r_pid[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K1R_STAGE_IT; ++k) {
                const int i = t + k * K1R_THREADS;
                if (i >= nc) break;
                const int32_t s = r_s[k] - a.start;
                const int n = r_n[k];
                const int64_t wl = r_w[k] - wb;
                if (n < 1 || n > maxlen || s < 0 || (int64_t)s + n > (int64_t)a.L || wl < 1 || wl + (((s & 7) + n + 7) >> 3) + 1 > wn ||
                    r_prev[k] > r_s[k])
                    err |= ISB_DEV_ERR_SEG;
                const int n_c = min(max(n, 0), maxlen);
                const int wl_c = (int)min(max(wl, (int64_t)1), wn - 1);
                const int s_rel = min(max(s - T0, -255), K1R_TILE - 1);   // candidates start in (T0 - 256, T0 + 1024)
                s_meta[i] = ((uint32_t)(wl_c - (s_rel >> 3) + 160) << 11) | (uint32_t)(s_rel + n_c + 256);
                s_start[i] = s;
                if (!kM1) {
                    if (r_mm[k] >= a.M) { err |= ISB_DEV_ERR_MM; r_mm[k] = 255; }
                    s_mm[i] = (uint8_t)r_mm[k];
                }
            }
        }
        __syncthreads();
        isb_mbar_wait(bar, parity);
        parity ^= 1u;
        if (!active) continue;

        // candidates of this thread inside the chunk: seg_start in (P - maxlen, P + 8)
        int cl, ch;
        {
            int l = 0, h = nc;
            const int key = P - maxlen + 1;
            while (l < h) { const int mid = (l + h) >> 1; if (s_start[mid] < key) l = mid + 1; else h = mid; }
            cl = l;
            h = nc;
            const int key2 = P + 8;
            while (l < h) { const int mid = (l + h) >> 1; if (s_start[mid] < key2) l = mid + 1; else h = mid; }
            ch = l;
        }
        // the 8 one-hot nibbles of segment i at the thread's positions (0 where the segment does not reach).  The stream
        // is position-aligned: the thread's column is ONE word of the segment, no shift.
        const int t_off = t - 160, P_end = t * 8 + 256;
        auto fetch = [&](int i) -> uint32_t {
            const uint32_t md = s_meta[i];
            const uint32_t x = s_words[(int)(md >> 11) + t_off];
            return P_end < (int)(md & 0x7ffu) ? x : 0u;            // short segment: those words belong to a later one
        };
        if (kM1) {
            int i = cl;
            for (; i + 16 <= ch; i += 16) {                        // two Harley-Seal blocks per trip: 16 fetches in flight
                uint32_t x[8], y[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { x[u] = fetch(i + u); y[u] = fetch(i + 8 + u); }
                k1r_add8(pl, x);
                k1r_add8(pl, y);
                n8 += 16;
                if (n8 > 239) {                                    // the next trip could overflow 255
                    k1r_planes_to_counts(c, pl);
                    n8 = 0;
                }
            }
            for (; i + 8 <= ch; i += 8) {
                uint32_t x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = fetch(i + u);
                k1r_add8(pl, x);
                n8 += 8;
                if (n8 > 247) {                                    // the next block could overflow 255
                    k1r_planes_to_counts(c, pl);
                    n8 = 0;
                }
            }
            if (i < ch) {
                uint32_t x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = (i + u < ch) ? fetch(min(i + u, ch - 1)) : 0u;
                k1r_add8(pl, x);
                n8 += 8;
                if (n8 > 247) {
                    k1r_planes_to_counts(c, pl);
                    n8 = 0;
                }
            }
        } else {
            for (int i = cl; i < ch;) {
                const int g_end = min(ch, i + 15);
                n8 += g_end - i;
                for (; i < g_end; ++i) {
                    const uint32_t x = fetch(i);
                    const int lv = (int)s_mm[i] - m_base;
                    if ((unsigned)lv < (unsigned)Mg) {             // 8-bit counters per (level, base, even/odd position)
                        uint32_t *acc = s_acc + (size_t)(lv * 8) * K1R_THREADS + t;
                        acc[0 * K1R_THREADS] += x & 0x01010101u;
                        acc[1 * K1R_THREADS] += (x >> 4) & 0x01010101u;
                        acc[2 * K1R_THREADS] += (x >> 1) & 0x01010101u;
                        acc[3 * K1R_THREADS] += (x >> 5) & 0x01010101u;
                        acc[4 * K1R_THREADS] += (x >> 2) & 0x01010101u;
                        acc[5 * K1R_THREADS] += (x >> 6) & 0x01010101u;
                        acc[6 * K1R_THREADS] += (x >> 3) & 0x01010101u;
                        acc[7 * K1R_THREADS] += (x >> 7) & 0x01010101u;
                    }
                }
                if (n8 > 240) {                                    // flush before a byte can overflow
                    k1r_flush_levels(a, s_acc, t, Mg, m_base, P, spilled, true);
                    spilled = true;
                    n8 = 0;
                }
            }
        }
    }
    if (err) atomicOr(a.d_err, err);
    if (!active) return;

    if (kM1) {
        k1r_planes_to_counts(c, pl);
        int4 *c4 = reinterpret_cast<int4 *>(a.counts) + P;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (P + k >= a.L) break;
            c4[k] = make_int4(c[k][0], c[k][1], c[k][2], c[k][3]);
        }
    } else {
        k1r_flush_levels(a, s_acc, t, Mg, m_base, P, spilled, false);
    }
}

// passing non-ACGT read bases ("N events"): the level becomes a key of the position's MMcounts (nmask bit), nothing else
__global__ void __launch_bounds__(256)
k1r_n_events(int64_t n_nev, const int32_t *__restrict__ nev_pos, const int32_t *__restrict__ nev_pair,
             const uint8_t *__restrict__ pair_mm, int64_t n_pairs, int32_t start, int32_t L, int M,
             unsigned long long *__restrict__ nmask, unsigned int *__restrict__ d_err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nev) return;
    const int64_t p = (int64_t)nev_pos[i] - start;
    const int32_t pid = nev_pair[i];
    if (p < 0 || p >= L || pid < 0 || pid >= n_pairs) { atomicOr(d_err, ISB_DEV_ERR_SEG); return; }
    const int mm = M > 1 ? pair_mm[pid] : 0;
    if (mm >= M) { atomicOr(d_err, ISB_DEV_ERR_MM); return; }
    atomicOr(nmask + p, 1ull << mm);
}

int isb_k1r_n_events_launch(isb_ctx *ctx, int64_t n_nev, const int32_t *nev_pos, const int32_t *nev_pair, const uint8_t *pair_mm,
                            int64_t n_pairs, int32_t start, int32_t L, int M, unsigned long long *nmask)
{
    if (n_nev <= 0 || !nmask) return ISB_OK;
    k1r_n_events<<<(unsigned)((n_nev + 255) / 256), 256, 0, ctx->stream>>>(n_nev, nev_pos, nev_pair, pair_mm, n_pairs, start, L, M,
                                                                           nmask, ctx->d_err);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

int isb_k1r_launch(isb_ctx *ctx, isb_reads_dev *rd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                   int M, int32_t *counts, unsigned long long *nmask)
{
    cudaStream_t st = ctx->stream;
    if (L <= 0) return ISB_OK;
    // default: the circular-schedule kernel of isb_k1f_fused.cu (no word staging, conflict-free); ISB_K1R_LEGACY=1 keeps
    // the first K1r (TMA-staged words, free-running per-thread candidate loops) for A/B runs
    static const int legacy = getenv("ISB_K1R_LEGACY") ? atoi(getenv("ISB_K1R_LEGACY")) : 0;
    if (!legacy) return isb_k1f_pileup_launch(ctx, rd, pair_mm, n_pairs, start, L, M, counts, nmask);
    if (rd->max_seg_len < 1 || rd->max_seg_len > K1R_MAXLEN)
        return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: max_seg_len must be in [1, 256]");
    if (M > 1 && !pair_mm) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: pair_mm is required when M > 1");
    if (((uintptr_t)rd->words & 15) != 0) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: words must be 16-byte aligned");
    if (start & 7) return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: start must be a multiple of 8 (the stream is position-aligned)");
    const int n_tiles = (L + K1R_TILE - 1) / K1R_TILE;
    int rc;
    if ((rc = isb_ensure(ctx, SL_RD_BOUNDS, sizeof(int64_t) * 4 * (size_t)n_tiles))) return rc;
    int64_t *tile_lo = (int64_t *)ctx->buf[SL_RD_BOUNDS].p, *tile_hi = tile_lo + n_tiles;
    int64_t *tile_wlo = tile_hi + n_tiles, *tile_whi = tile_wlo + n_tiles;
    k1r_tile_bounds<<<(n_tiles + 255) / 256, 256, 0, st>>>(rd->seg_start, rd->seg_len, rd->seg_word, rd->n_segs, start, n_tiles,
                                                           rd->max_seg_len, tile_lo, tile_hi, tile_wlo, tile_whi);
    ISB_LAUNCH_CHECK();
    rd->n_tiles = n_tiles;
    rd->tile_lo = tile_lo;
    rd->tile_hi = tile_hi;
    rd->tile_wlo = tile_wlo;
    rd->tile_whi = tile_whi;

    k1r_args a;
    a.rd = *rd; a.pair_mm = pair_mm; a.n_pairs = n_pairs; a.start = start; a.L = L; a.M = M; a.counts = counts;
    a.nmask = nmask; a.d_err = ctx->d_err;
    // Shared-memory budget: the staging area should hold the whole candidate set of a tile (then every thread works in
    // every chunk); two blocks per SM.  Bytes per staged segment: its words incl. separator + 4 (meta) + 4 (start) + 1 (mm).
    const int wps = (rd->max_seg_len + 14) / 8 + 2;               // position-aligned data words + up to two separator words
    const size_t per_seg = (size_t)wps * 4 + 9;
    const int groups = M == 1 ? 1 : (M + K1R_LEVELS - 1) / K1R_LEVELS;
    const int Mg = M == 1 ? 0 : (M < K1R_LEVELS ? M : K1R_LEVELS);
    const size_t acc_bytes = (size_t)Mg * 8 * K1R_THREADS * 4;
    const size_t budget = (size_t)110 * 1024;
    const size_t stage_budget = acc_bytes + 24 * 1024 < budget ? budget - acc_bytes : 24 * 1024;
    int seg_cap = (int)((stage_budget - 64) / per_seg);
    const int64_t avg_need = rd->n_segs > 0 ? (int64_t)((double)rd->n_segs / L * (K1R_TILE + rd->max_seg_len) * 1.25) + 32 : 32;
    if (seg_cap > avg_need) seg_cap = (int)avg_need;              // no point staging more than a tile ever holds
    if (seg_cap < 64) seg_cap = 64;
    if (seg_cap > K1R_STAGE_IT * K1R_THREADS) seg_cap = K1R_STAGE_IT * K1R_THREADS;
    a.seg_cap = seg_cap;
    a.words_cap = (seg_cap * wps + 8 + 3) & ~3;
    const size_t smem = (((size_t)a.words_cap * 4 + (size_t)seg_cap * 9 + 7) & ~(size_t)7) + 16 + acc_bytes;
    static bool attr_m1[64] = {false}, attr_mm[64] = {false};      // function attributes are per device
    if (nmask) ISB_CUDA(cudaMemsetAsync(nmask, 0, sizeof(unsigned long long) * (size_t)L, st));
    if (M == 1) {
        if (!attr_m1[ctx->device & 63])
            ISB_CUDA(cudaFuncSetAttribute(k1r_pileup<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_m1[ctx->device & 63] = true;
        k1r_pileup<true><<<n_tiles, K1R_THREADS, smem, st>>>(a);
        ISB_LAUNCH_CHECK();
    } else {
        if (!attr_mm[ctx->device & 63])
            ISB_CUDA(cudaFuncSetAttribute(k1r_pileup<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_mm[ctx->device & 63] = true;
        k1r_pileup<false><<<dim3(n_tiles, groups), K1R_THREADS, smem, st>>>(a);
        ISB_LAUNCH_CHECK();
    }
    if (nmask && rd->n_nev > 0) return isb_k1r_n_events_launch(ctx, rd->n_nev, rd->nev_pos, rd->nev_pair, pair_mm, n_pairs, start, L, M, nmask);
    return ISB_OK;
}
