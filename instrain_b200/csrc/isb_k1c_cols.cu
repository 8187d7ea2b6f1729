// isb_k1c_cols.cu -- K1c: pileup counts from COLUMN WORDS (the pileup-major form of the aligned segments), sm_100a.
//
// Same result as K1 / K1r: counts[position][mm][A,C,T,G] (+ nmask) of samfile.pileup(...) column iteration
// (inStrain/profile/profile_utilities.py:150-153) + get_base_counts_mm (:268-286).  The input is the same one-hot
// nibble word per (read, 8 aligned positions) as the read-major stream, but stored where the pileup needs it
// (include/instrain_b200.h, isb_cols_batch): for every column word (8 consecutive positions) the words of the reads
// that cover it, and the column lists of 8 neighbouring column words (one GROUP = 64 positions) interleaved in
// 32-byte units (8 slots of one column), so that a chunk row of a group is 256 contiguous bytes.  The transposition
// "reads -> columns" that pysam's pileup engine performs per column is done once by the packer
// (isb_cols_from_reads*), so the kernel is a pure stream:
//
//   * a warp takes 4 consecutive groups (256 positions), lane = column word; every lane issues ONE 256-bit load per
//     chunk row (LDG.E.256: a warp instruction reads 1 KB), no shared memory, no atomics, no searches;
//   * counting is bit-sliced as in K1r (isb_bitslice.cuh): 24 logic ops per 8 words;
//   * M = 1 can run the SNV call of K2 (k2_site_m1, isb_k2_site.cuh) in its epilogue: the warp's counts are transposed
//     through shared memory to "lane = position", covT / clonT / site_flags / SNV rows leave the kernel directly with
//     coalesced stores, counts are written only where the linkage stage reads them (flagged sites) unless the caller
//     asks for the full array;
//   * M > 1 gathers pair_mm[id] per word and keeps 8-bit counters per (level, base) in shared memory, [word][thread].
//
// HBM traffic: 0.5 B per aligned base (+ chunk padding, + 4 B id per word at M > 1) in, 16*M B per position out
// (fused M = 1: 9 B per position out).  Measured on B200: the unfused M = 1 kernel runs at 0.81 of the HBM copy peak;
// the fused one at 0.64 (its SNV epilogue is issue-bound) but replaces two kernels and a 32 B/position round trip.
#include "isb_common.cuh"
#include "isb_bitslice.cuh"
#include "isb_k2_site.cuh"
#include "isb_scan.cuh"
#include <cstdlib>

#define K1C_WARPS 4                        // warps per block; a warp owns 256 positions = 32 / ISB_COLS_LANES groups
#define K1C_THREADS (32 * K1C_WARPS)
#define K1C_LEVELS 32                      // mm levels per pass of the M > 1 kernel (shared-memory accumulators)
#ifndef K1C_ROWS
#define K1C_ROWS 2                         // chunk rows (one 256-bit load per lane each) per trip of the M = 1 main loop
#endif
#ifndef K1C_MINB
#define K1C_MINB 1                         // __launch_bounds__ min blocks per SM of the M = 1 kernels
#endif
#define K1C_TILE (256 + 32)                // count quads of a warp's 256 positions + one pad quad per 8 positions
static_assert(ISB_COLS_UNIT == 8 && 32 % ISB_COLS_LANES == 0, "K1c loads one 8-word unit per lane and chunk");

struct k1c_args {
    isb_cols_dev cd;
    const uint8_t *pair_mm;
    int64_t n_pairs;
    int32_t L;
    int32_t start;
    int M;
    int32_t *counts;
    int write_counts;                      // fused kernel: 1 = every position, 0 = only flagged sites (what K3 reads)
    const unsigned long long *nmask;       // fused kernel: read where a position has no A/C/T/G count (may be NULL)
    isb_k2_fuse k2;
    const int32_t *thr2;
    int n_lut, lut_default;
    unsigned long long *n_rows;
    unsigned int *d_err;
};

// one 32-byte unit (8 words of one column) with a single 256-bit load; streamed data: no L1 allocation
__device__ __forceinline__ void k1c_ld256(const void *p, uint32_t (&x)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}

// chunk range of the lane's group (lanes of one group read the same two offsets: broadcast); 0 chunks beyond the batch
__device__ __forceinline__ void k1c_group_range(const k1c_args &a, int64_t g, int64_t &c0, int &nch)
{
    c0 = 0;
    nch = 0;
    if (g >= a.cd.n_groups) return;
    const int64_t lo = __ldg(a.cd.grp_off + g), hi = __ldg(a.cd.grp_off + g + 1);
    if (lo < 0 || hi < lo || hi > a.cd.n_chunks || hi - lo > (1 << 24)) {        // layout rules violated
        atomicOr(a.d_err, ISB_DEV_ERR_SEG);
        return;
    }
    c0 = lo;
    nch = (int)(hi - lo);
}

template <bool kFuse>
__global__ void __launch_bounds__(K1C_THREADS, K1C_MINB) k1c_pileup_m1(k1c_args a)
{
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * K1C_WARPS + (threadIdx.x >> 5);     // the warp's 256 positions
    int64_t c0;
    int nch;                                                                     // per group: lanes of a warp may differ
    k1c_group_range(a, wg * (32 / ISB_COLS_LANES) + lane / ISB_COLS_LANES, c0, nch);
    // the lane's unit of chunk row r: words + ((c0 + r) * LANES + lane % LANES) * 8
    const uint32_t *src = a.cd.words + (c0 * ISB_COLS_LANES + (lane % ISB_COLS_LANES)) * ISB_COLS_UNIT;

    int c[8][4];
    uint32_t pl[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};            // vertical counter planes (weights 1 .. 128)
    int n8 = 0;                                                   // words since the last flush (counters hold <= 255)
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) c[k][b] = 0;

    int ch = 0;
    for (; ch + K1C_ROWS <= nch; ch += K1C_ROWS) {                // K1C_ROWS x 32 bytes per lane in flight
        uint32_t x[K1C_ROWS][8];
#pragma unroll
        for (int r = 0; r < K1C_ROWS; ++r) k1c_ld256(src + (size_t)(ch + r) * ISB_COLS_CHUNK, x[r]);
#pragma unroll
        for (int r = 0; r < K1C_ROWS; ++r) k1r_add8(pl, x[r]);
        n8 += 8 * K1C_ROWS;
        if (n8 > 255 - 8 * K1C_ROWS) {                            // the next trip could overflow 255
            k1r_planes_to_counts(c, pl);
            n8 = 0;
        }
    }
    for (; ch < nch; ++ch) {
        uint32_t x[8];
        k1c_ld256(src + (size_t)ch * ISB_COLS_CHUNK, x);
        k1r_add8(pl, x);
        n8 += 8;
        if (n8 > 247) {
            k1r_planes_to_counts(c, pl);
            n8 = 0;
        }
    }
    k1r_planes_to_counts(c, pl);

    if (!kFuse) {                                                 // counts of the lane's 8 positions (128 contiguous bytes)
        const int64_t P = wg * 256 + lane * 8;
        int4 *c4 = reinterpret_cast<int4 *>(a.counts) + P;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (P + k >= a.L) break;
            c4[k] = make_int4(c[k][0], c[k][1], c[k][2], c[k][3]);
        }
        return;
    }

    // ---- fused SNV call (K2 at M = 1) ------------------------------------------------------------------------------
    // The warp's 256 count quads are transposed through a warp-private shared-memory tile (one pad quad per 8
    // positions: conflict-free both ways), then 8 rounds of "lane = position" run the site arithmetic of K2 with fully
    // coalesced covT / clonT / site_flags stores.  Rows are allocated ONCE per warp (prefix sum of the lanes' row bits,
    // one atomic) and written by re-evaluating the ~1 % of sites that have one.
    __shared__ int4 s_tile[kFuse ? K1C_WARPS * K1C_TILE : 1];
    int4 *tile = s_tile + (threadIdx.x >> 5) * K1C_TILE;
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[lane * 9 + k] = make_int4(c[k][0], c[k][1], c[k][2], c[k][3]);
    __syncwarp();
    const int64_t W0 = wg * 256;
    int4 *counts4 = reinterpret_cast<int4 *>(a.counts);
    unsigned rowmask = 0u;                                        // bit r: position W0 + 32 r + lane emits a raw_snp_table row
    auto site = [&](int64_t p, const int4 &E, int (&C)[4], int &r) -> k2_m1_site {
        C[0] = E.x; C[1] = E.y; C[2] = E.z; C[3] = E.w;
        r = a.k2.ref[p];
        const int T = E.x + E.y + E.z + E.w;
        const int thr_T = (T >= a.k2.min_cov && T < a.n_lut) ? __ldg(a.thr2 + T) : a.lut_default;
        const bool nm0 = T == 0 && a.nmask && (a.nmask[p] & 1ull);
        return k2_site_m1(C, r, nm0, thr_T, a.n_lut, a.lut_default, a.k2.min_cov, a.k2.min_freq);
    };
#pragma unroll 2
    for (int rd = 0; rd < 8; ++rd) {
        const int q = rd * 32 + lane;
        const int64_t p = W0 + q;
        if (p >= a.L) continue;
        const int4 E = tile[q + (q >> 3)];
        int C[4], r;
        const k2_m1_site s = site(p, E, C, r);
        a.k2.covT[p] = s.T;
        a.k2.clonT[p] = s.clon;
        a.k2.site_flags[p] = (uint8_t)s.flags;
        if (a.k2.clonTR)
            a.k2.clonTR[p] = (a.k2.cov_r > 0 && s.T >= a.k2.cov_r) ? k2_rarefied_clon(C, s.T, a.k2.cov_r, a.k2.seed, (int64_t)p + a.start, 0)
                                                                 : CUDART_NAN_F;
        if (a.write_counts || s.flags) counts4[p] = E;            // K3 reads counts at flagged sites only
        rowmask |= (s.is_row ? 1u : 0u) << rd;
    }
    const int my_rows = __popc(rowmask);
    int incl = my_rows;                                           // warp prefix sum of the row counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(ISB_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(ISB_FULL, incl, 31);
    if (total == 0) return;
    unsigned long long base_slot = 0;
    if (lane == 31) base_slot = atomicAdd(a.n_rows, (unsigned long long)total);
    base_slot = __shfl_sync(ISB_FULL, base_slot, 31);
    int64_t slot = (int64_t)base_slot + incl - my_rows;
    while (rowmask) {
        const int rd = __ffs((int)rowmask) - 1;
        rowmask &= rowmask - 1u;
        const int q = rd * 32 + lane;
        const int64_t p = W0 + q;
        const int4 E = tile[q + (q >> 3)];
        int C[4], r;
        const k2_m1_site s = site(p, E, C, r);
        if (slot < a.k2.cap) k2_write_row_m1(a.k2.rows + slot, (int32_t)p + a.start, C, r, s, a.n_lut, a.k2.min_freq);
        ++slot;
    }
}

// M > 1: write (or add, once counts hold a partial sum) the thread's shared 8-bit counters to its cells of `counts`
__device__ __forceinline__ void k1c_flush_levels(const k1c_args &a, uint32_t *s_acc, int t, int Mg, int m_base, int32_t P,
                                                 bool add, bool clear)
{
    int4 *c4 = reinterpret_cast<int4 *>(a.counts);
    for (int m = 0; m < Mg; ++m) {
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            w8[j] = s_acc[(size_t)(m * 8 + j) * K1C_THREADS + t];
            if (clear) s_acc[(size_t)(m * 8 + j) * K1C_THREADS + t] = 0u;
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

__global__ void __launch_bounds__(K1C_THREADS) k1c_pileup_mm(k1c_args a)
{
    extern __shared__ __align__(16) uint32_t k1c_acc[];           // [Mg * 8][K1C_THREADS]: every lane owns its bank
    const int t = threadIdx.x, lane = t & 31;
    const int64_t wg = (int64_t)blockIdx.x * K1C_WARPS + (t >> 5);
    const int m_base = (int)blockIdx.y * K1C_LEVELS;
    const int Mg = min(K1C_LEVELS, a.M - m_base);
    for (int w = 0; w < Mg * 8; ++w) k1c_acc[w * K1C_THREADS + t] = 0u;
    int64_t c0;
    int nch;
    k1c_group_range(a, wg * (32 / ISB_COLS_LANES) + lane / ISB_COLS_LANES, c0, nch);
    const int64_t P64 = wg * 256 + lane * 8;
    const size_t unit0 = (size_t)(c0 * ISB_COLS_LANES + (lane % ISB_COLS_LANES)) * ISB_COLS_UNIT;
    const uint32_t *src = a.cd.words + unit0;
    const int32_t *sid = a.cd.ids + unit0;
    int n8 = 0;
    bool spilled = false;
    unsigned err = 0;
    for (int ch = 0; ch < nch; ++ch) {                            // one chunk row per trip: words + ids, all mm gathers, then counts
        uint32_t x[8], idu[8];
        k1c_ld256(src + (size_t)ch * ISB_COLS_CHUNK, x);
        k1c_ld256(sid + (size_t)ch * ISB_COLS_CHUNK, idu);
        int id[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) id[u] = (int)idu[u];
        int mm[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            mm[u] = -1;
            if (id[u] >= 0) {
                if ((int64_t)id[u] < a.n_pairs) mm[u] = __ldg(a.pair_mm + id[u]);
                else err |= ISB_DEV_ERR_SEG;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (mm[u] < 0) continue;                              // padding
            if (mm[u] >= a.M) { err |= ISB_DEV_ERR_MM; continue; }
            const int lv = mm[u] - m_base;
            if ((unsigned)lv < (unsigned)Mg) {                    // 8-bit counters per (level, base, even / odd position)
                uint32_t *acc = k1c_acc + (size_t)(lv * 8) * K1C_THREADS + t;
                acc[0 * K1C_THREADS] += x[u] & 0x01010101u;
                acc[1 * K1C_THREADS] += (x[u] >> 4) & 0x01010101u;
                acc[2 * K1C_THREADS] += (x[u] >> 1) & 0x01010101u;
                acc[3 * K1C_THREADS] += (x[u] >> 5) & 0x01010101u;
                acc[4 * K1C_THREADS] += (x[u] >> 2) & 0x01010101u;
                acc[5 * K1C_THREADS] += (x[u] >> 6) & 0x01010101u;
                acc[6 * K1C_THREADS] += (x[u] >> 3) & 0x01010101u;
                acc[7 * K1C_THREADS] += (x[u] >> 7) & 0x01010101u;
            }
        }
        n8 += 8;
        if (n8 > 240 && P64 < a.L) {                              // flush before a byte can overflow
            k1c_flush_levels(a, k1c_acc, t, Mg, m_base, (int32_t)P64, spilled, true);
            spilled = true;
            n8 = 0;
        }
    }
    if (err) atomicOr(a.d_err, err);
    if (P64 < a.L) k1c_flush_levels(a, k1c_acc, t, Mg, m_base, (int32_t)P64, spilled, false);
}

int isb_k1c_launch(isb_ctx *ctx, const isb_cols_dev *cd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                   int M, int32_t *counts, unsigned long long *nmask, const isb_k2_fuse *fuse, bool init_nmask)
{
    cudaStream_t st = ctx->stream;
    if (L <= 0) return ISB_OK;
    const int64_t n_groups = ((int64_t)L + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP;
    if (cd->n_groups != n_groups) return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: n_groups must be ceil(L / 64)");
    if (cd->n_chunks < 0 || !cd->grp_off || (cd->n_chunks > 0 && !cd->words))
        return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: null grp_off / words");
    if (((uintptr_t)cd->words & 31) != 0 || ((uintptr_t)cd->ids & 31) != 0 || ((uintptr_t)counts & 15) != 0)
        return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: words and ids must be 32-byte aligned, counts 16-byte aligned");
    if (start & 7) return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: start must be a multiple of 8");
    if (M > 1 && (!pair_mm || !cd->ids)) return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: pair_mm and ids are required when M > 1");
    if (fuse && M != 1) return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: the fused SNV call needs M == 1");
    int rc;
    if (nmask && init_nmask) {
        ISB_CUDA(cudaMemsetAsync(nmask, 0, sizeof(unsigned long long) * (size_t)L, st));
        if ((rc = isb_k1r_n_events_launch(ctx, cd->n_nev, cd->nev_pos, cd->nev_pair, pair_mm, n_pairs, start, L, M, nmask))) return rc;
    }
    k1c_args a;
    memset(&a, 0, sizeof(a));
    a.cd = *cd; a.pair_mm = pair_mm; a.n_pairs = n_pairs; a.L = L; a.start = start; a.M = M; a.counts = counts;
    a.write_counts = 1; a.d_err = ctx->d_err;
    const unsigned blocks = (unsigned)((((int64_t)L + 255) / 256 + K1C_WARPS - 1) / K1C_WARPS);
    if (M == 1 && fuse) {
        if ((rc = isb_k2_prepare(ctx, fuse->min_freq))) return rc;
        a.k2 = *fuse;
        a.thr2 = ctx->d_thr2; a.n_lut = ctx->n_lut; a.lut_default = ctx->lut_default;
        a.n_rows = ctx->d_counters + 0;
        a.nmask = nmask;
        a.write_counts = fuse->full_counts ? 1 : 0;
        k1c_pileup_m1<true><<<blocks, K1C_THREADS, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
    } else if (M == 1) {
        k1c_pileup_m1<false><<<blocks, K1C_THREADS, 0, st>>>(a);
        ISB_LAUNCH_CHECK();
    } else {
        const int groups = (M + K1C_LEVELS - 1) / K1C_LEVELS;
        const int Mg = M < K1C_LEVELS ? M : K1C_LEVELS;
        const size_t smem = (size_t)Mg * 8 * K1C_THREADS * 4;
        static bool attr_mm[64] = {false};                          // function attributes are per device
        if (!attr_mm[ctx->device & 63])
            ISB_CUDA(cudaFuncSetAttribute(k1c_pileup_mm, cudaFuncAttributeMaxDynamicSharedMemorySize, K1C_LEVELS * 8 * K1C_THREADS * 4));
        attr_mm[ctx->device & 63] = true;
        k1c_pileup_mm<<<dim3(blocks, groups), K1C_THREADS, smem, st>>>(a);
        ISB_LAUNCH_CHECK();
    }
    return ISB_OK;
}

// ---- layout conversion: read-major aligned segments -> column words ----------------------------------------------------
// Not on the timed path (the packer produces the layout once per batch; bench.py uses it to lay out the generated
// data set).  One thread per column word walks the start-sorted segment table over the starts that can reach the
// column, so the words of a column keep the TABLE ORDER of their segments (= BAM order: what the linkage stage's
// self-edge rule depends on).

__global__ void __launch_bounds__(256)
k0c_check_segments(isb_reads_dev rd, int32_t start, int32_t L, unsigned int *__restrict__ d_err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rd.n_segs) return;
    const int64_t s = rd.seg_start[i];
    const int n = rd.seg_len[i];
    const int64_t nw = ((s & 7) + n + 7) >> 3;
    const int64_t w = rd.seg_word[i];
    if (n < 1 || n > rd.max_seg_len || s < start || s + n > (int64_t)start + L || w < 0 || w + nw > rd.n_words ||
        (i > 0 && rd.seg_start[i - 1] > s))
        atomicOr(d_err, ISB_DEV_ERR_SEG);
}

template <bool kFill>
__global__ void __launch_bounds__(256)
k0c_columns(isb_reads_dev rd, int32_t start, int32_t L, int64_t n_cols, int32_t *__restrict__ grp_chunks,
            const int64_t *__restrict__ grp_off, uint32_t *__restrict__ words, int32_t *__restrict__ ids)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // n_cols is padded to whole warps
    const bool real = c < n_cols;                                           // a column of an existing group
    const int sl = (int)(c % ISB_COLS_LANES);
    const int64_t W8 = (int64_t)start + c * 8;                              // first coordinate of the column word
    int64_t base = 0;
    int depth = 0;
    if (kFill && real) {
        const int64_t g = c / ISB_COLS_LANES;
        base = (grp_off[g] * ISB_COLS_LANES + sl) * ISB_COLS_UNIT;
        depth = (int)(grp_off[g + 1] - grp_off[g]) * ISB_COLS_UNIT;
    }
    int slot = 0;
    if (real && c * 8 < L) {
        int64_t i = isb_lower_bound(rd.seg_start, 0, rd.n_segs, W8 - rd.max_seg_len + 1);
        for (; i < rd.n_segs; ++i) {
            const int64_t s = __ldg(rd.seg_start + i);
            if (s >= W8 + 8) break;
            if (s + (int64_t)__ldg(rd.seg_len + i) - 1 < W8) continue;    // ends before the column
            if (kFill && slot < depth) {
                const int64_t idx = base + (int64_t)(slot / ISB_COLS_UNIT) * ISB_COLS_CHUNK + (slot % ISB_COLS_UNIT);
                words[idx] = __ldg(rd.words + __ldg(rd.seg_word + i) + ((W8 >> 3) - (s >> 3)));
                ids[idx] = __ldg(rd.seg_pair + i);
            }
            ++slot;
        }
    }
    if (kFill) {
        for (; slot < depth; ++slot) {                                      // padding up to the group's depth
            const int64_t idx = base + (int64_t)(slot / ISB_COLS_UNIT) * ISB_COLS_CHUNK + (slot % ISB_COLS_UNIT);
            words[idx] = 0u;
            ids[idx] = -1;
        }
    } else {
        int mx = slot;
#pragma unroll
        for (int d = ISB_COLS_LANES / 2; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(ISB_FULL, mx, d));   // max over the group's lanes
        if (real && sl == 0) grp_chunks[c / ISB_COLS_LANES] = (mx + ISB_COLS_UNIT - 1) / ISB_COLS_UNIT;
    }
}

struct ChunksFn {
    const int32_t *grp_chunks;
    __device__ int operator()(int64_t i) const { return grp_chunks[i]; }
};
struct GrpOffSink {
    int64_t *grp_off;
    __device__ void operator()(int64_t i, int64_t prefix, int) const { grp_off[i] = prefix; }
};

// grp_off[n_groups + 1] always; words / ids (cap_chunks chunks of ISB_COLS_CHUNK words) when given.  *n_chunks = chunks needed.
int isb_cols_convert(isb_ctx *ctx, const isb_reads_dev *rd, int32_t start, int32_t L, int64_t *grp_off, uint32_t *words,
                     int32_t *ids, int64_t cap_chunks, int64_t *n_chunks)
{
    cudaStream_t st = ctx->stream;
    *n_chunks = 0;
    if (start & 7) return isb_fail(ctx, ISB_ERR_ARG, "column-word conversion: start must be a multiple of 8");
    if (rd->max_seg_len < 1 || rd->max_seg_len > 256)
        return isb_fail(ctx, ISB_ERR_ARG, "column-word conversion: max_seg_len must be in [1, 256]");
    const int64_t n_groups = ((int64_t)L + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP;
    if (n_groups == 0) {
        ISB_CUDA(cudaMemsetAsync(grp_off, 0, sizeof(int64_t), st));
        return ISB_OK;
    }
    int rc;
    if ((rc = isb_ensure(ctx, SL_CD_CNT, sizeof(int32_t) * (size_t)n_groups))) return rc;
    int32_t *grp_chunks = (int32_t *)ctx->buf[SL_CD_CNT].p;
    const int64_t n_cols = n_groups * ISB_COLS_LANES;
    if (rd->n_segs > 0) {
        k0c_check_segments<<<(unsigned)((rd->n_segs + 255) / 256), 256, 0, st>>>(*rd, start, L, ctx->d_err);
        ISB_LAUNCH_CHECK();
    }
    k0c_columns<false><<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(*rd, start, L, n_cols, grp_chunks, nullptr, nullptr, nullptr);
    ISB_LAUNCH_CHECK();
    const int nb = (int)((n_groups + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if ((rc = isb_ensure(ctx, SL_SCAN_TMP, sizeof(int64_t) * (size_t)nb))) return rc;
    int64_t *block_sums = (int64_t *)ctx->buf[SL_SCAN_TMP].p;
    ChunksFn cf{grp_chunks};
    scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(cf, n_groups, block_sums);
    ISB_LAUNCH_CHECK();
    scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nb, ctx->d_counters + 7);
    ISB_LAUNCH_CHECK();
    GrpOffSink sink{grp_off};
    scan_scatter<<<nb, SCAN_THREADS, 0, st>>>(cf, n_groups, block_sums, sink);
    ISB_LAUNCH_CHECK();
    ISB_CUDA(cudaMemcpyAsync(grp_off + n_groups, ctx->d_counters + 7, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(ctx->h_counters + 7, ctx->d_counters + 7, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    *n_chunks = (int64_t)ctx->h_counters[7];
    if (*ctx->h_err & ISB_DEV_ERR_SEG) return ISB_OK;             // the caller reports it (check_dev_err)
    if (!words) return ISB_OK;
    if (!ids) return isb_fail(ctx, ISB_ERR_ARG, "column-word conversion: ids is required with words");
    if (*n_chunks > cap_chunks) return isb_fail(ctx, ISB_ERR_CAPACITY, "column-word conversion: words / ids too small (see n_chunks)");
    k0c_columns<true><<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(*rd, start, L, n_cols, nullptr, grp_off, words, ids);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
