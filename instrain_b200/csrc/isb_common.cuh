// isb_common.cuh -- shared declarations for libinstrain_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/instrain_b200.h"

#define ISB_WARP 32
#define ISB_FULL 0xffffffffu

// device-side error bits (ctx->d_err)
#define ISB_DEV_ERR_ORDER 0x1u        // event outside its position tile: input not position-major
#define ISB_DEV_ERR_MM 0x2u           // pair_mm >= M
#define ISB_DEV_ERR_MULT 0x4u         // a read pair has > 2 qualifying events on one site
#define ISB_DEV_ERR_ROWBUF 0x8u       // linkage bit-row scratch too small (host grows it and re-runs K3)
#define ISB_DEV_ERR_SEG 0x10u         // read-major batch violates its layout rules (order, range, word offsets)
#define ISB_DEV_ERR_SITECAP 0x20u     // fused read-major path: more linkage sites than site slots (host grows them and re-runs)
#define ISB_DEV_ERR_QCAP 0x40u        // fused read-major path: more sites for the general SNV call than queue slots (same remedy)

#define ISB_K3_ROW_SLOT 64            // words of the fixed bit-row slot every linkage site owns (wider rows go to an overflow region)

enum {  // scratch buffer slots of a context (grow-only device allocations)
    SL_REF_POS = 0, SL_BASE, SL_QUAL, SL_READ_ID, SL_PAIR_MM, SL_REF, SL_SPLITS,   // staged inputs
    SL_COUNTS, SL_NMASK, SL_COVT, SL_CLONT, SL_CLONTR, SL_FLAGS, SL_SNV, SL_LD,              // staged outputs
    SL_TILE_OFF,                                                                   // K1 tile event offsets
    SL_PK_OFF, SL_PK_IDBASE, SL_PK_BQD, SL_PK_ESC_EVT, SL_PK_ESC_ID,                  // packed transfer format (K0 inputs)
    SL_K3_TILE_OFF, SL_PAIRS,
    SL_K4_CUM, SL_K4_CLON, SL_K4_STATE, SL_K4_HIST, SL_K4_OFF, SL_K4_OUT,
    SL_RD_START, SL_RD_LEN, SL_RD_PAIR, SL_RD_WORD, SL_RD_WORDS, SL_RD_BOUNDS, SL_RD_NPOS, SL_RD_NPAIR,         // read-major batch (K1r inputs)
    SL_RC_BASE2, SL_RC_PASS, SL_RC_MISW, SL_RC_MISC,                               // compact / delta transfer formats (K0r / K0d inputs)
    SL_RD_CAND, SL_RD_EVOFF, SL_RD_EVB, SL_RD_EVQ, SL_RD_EVID,                          // K3 site events from segments
    SL_CD_OFF, SL_CD_WORDS, SL_CD_IDS, SL_CD_CNT,                                  // column-word batch (K1c inputs) + conversion scratch
    SL_SCAN_TMP, SL_SITE_POS, SL_SITE_META, SL_SITE_WORDS, SL_ROW_OFF, SL_ROWS, SL_MM_MASK, SL_HAS2,
    SL_TILE_SITES, SL_SITE_COUNTS,                                                 // fused read-major path: per-tile site slots, counts per site
    SL_Q_TILE, SL_Q_POS, SL_Q_E, SL_Q_CAND, SL_SITE_CAND,                          // fused read-major path: queue of the sites for the general SNV call
    SL_COUNT
};

struct isb_devbuf {
    void *p;
    size_t cap;
};

struct isb_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int32_t *d_lut;
    int n_lut;
    int lut_default;
    int32_t *d_thr2;                  // K2: max(null-model threshold, min count passing min_freq) per coverage
    double thr2_min_freq;
    cudaStream_t aux_stream;          // chunk pipeline: K2 / K3 of chunk c run here while K1 of chunk c+1 runs on `stream`
    int keep_counters;                // set inside the chunk pipeline: K2 / K3 append to the row counters instead of resetting
    unsigned long long *d_counters;   // [0]=n_snv rows [1]=n_ld rows [2]=n_sites [3]=n_site_pairs [4]=total row words
    unsigned int *d_err;
    unsigned long long *h_counters;   // pinned mirror
    unsigned int *h_err;
    int64_t launches;
    int64_t sites_cap;                // fused read-major path: linkage-site slots allocated so far (grow-only)
    int64_t queue_cap;                // fused read-major path: queue slots (sites for the general SNV call) allocated so far
    isb_devbuf buf[SL_COUNT];
    char err[512];
    uint64_t seed;                    // isb_params.seed of the running call (linkage: normalized columns)
    // optional per-stage device timing (isb_enable_timing): CUDA events recorded on ctx->stream around K1 / K2 / K3
    int timing;
    int n_tev;
    int cap_tev;
    struct isb_tev *tev;
};

struct isb_tev {
    int stage;               // 0 = K1, 1 = K2, 2 = K3
    cudaEvent_t a, b;
};

int isb_time_begin(isb_ctx *ctx, int stage);   // returns slot index or -1 (timing off)
void isb_time_end(isb_ctx *ctx, int slot);

struct isb_site_meta {   // one linkage-eligible site (16 bytes)
    int32_t ev_lo_rel;   // first event of the site, relative to the site's tile (unused; kept for alignment)
    int32_t wlo;         // first 32-bit word of the pair-id window
    int32_t nw;          // words in the window (0: no qualifying event)
    int32_t split;       // split index, -1 if the position is in no split
};

#define ISB_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                     cudaGetErrorString(_e));                                                      \
            return ISB_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define ISB_LAUNCH_CHECK()                                                                         \
    do {                                                                                           \
        ctx->launches++;                                                                           \
        ISB_CUDA(cudaGetLastError());                                                              \
    } while (0)

static inline int isb_fail(isb_ctx *ctx, int code, const char *msg)
{
    snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}

// ---- launchers implemented in the kernel translation units -------------------------------------------------
int isb_k1_launch(isb_ctx *ctx, int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                  const int32_t *read_id, const uint8_t *pair_mm, int32_t start, int32_t L, int M, int min_qual,
                  uint32_t flags, int32_t *counts, unsigned long long *nmask);
int isb_k2_launch(isb_ctx *ctx, int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                  const uint8_t *ref, int32_t start, int min_cov, double min_freq, int32_t *covT, float *clonT,
                  uint8_t *site_flags, isb_snv_row *rows, int64_t cap, float *clonTR = nullptr, int cov_r = 0,
                  uint64_t seed = 0);
int isb_k3_launch(isb_ctx *ctx, int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                  const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start, int32_t L, int M,
                  int min_qual, const int32_t *counts, const unsigned long long *nmask, const uint8_t *site_flags,
                  int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows, int64_t cap);
// read-major batch, device pointers (+ the per-tile segment bounds isb_k1r_launch computes)
#ifndef K1R_TILE
#define K1R_TILE 1024                 // positions per K1r block; also the granularity of the K3 candidate search
#endif
struct isb_k2_fuse;
typedef struct isb_k2_fuse isb_k2_fuse_fwd;
struct isb_reads_dev {
    int64_t n_segs;
    const int32_t *seg_start;
    const uint16_t *seg_len;
    const int32_t *seg_pair;
    const int64_t *seg_word;
    int64_t n_words;
    const uint32_t *words;
    int32_t max_seg_len;
    int64_t n_nev;                    // passing non-ACGT read bases (nmask only)
    const int32_t *nev_pos;
    const int32_t *nev_pair;
    int n_tiles;
    const int64_t *tile_lo;           // [n_tiles] first segment with seg_start > tile_first - max_seg_len
    const int64_t *tile_hi;           // [n_tiles] first segment with seg_start >= tile_first + K1R_TILE
    const int64_t *tile_wlo;          // [n_tiles] word range [wlo, whi) of those segments incl. both separators
    const int64_t *tile_whi;
};
int isb_k1r_launch(isb_ctx *ctx, isb_reads_dev *rd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                   int M, int32_t *counts, unsigned long long *nmask);
// Fused read-major path (M = 1): K1f (pileup + SNV call + bit rows of the linkage sites, isb_k1f_fused.cu) and the
// linkage back end on its per-tile site slots (isb_k3_linkage.cu).  Everything is enqueued on ctx->stream without a
// host round trip; the caller reads the counters / error word afterwards and re-runs after isb_k1f_grow when a scratch
// capacity was exceeded.
struct isb_k1f_linkage {
    int32_t n_splits;
    const int32_t *splits;
    int min_snp;
    isb_ld_row *rows;
    int64_t cap;
};
int isb_k1f_profile_launch(isb_ctx *ctx, isb_reads_dev *rd, int64_t n_pairs, int32_t start, int32_t L,
                           unsigned long long *nmask, const isb_k2_fuse_fwd *fuse, const isb_k1f_linkage *ld);
int isb_k1f_grow(isb_ctx *ctx, bool *again);
int isb_k1f_pileup_launch(isb_ctx *ctx, isb_reads_dev *rd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                          int M, int32_t *counts, unsigned long long *nmask);
// linkage back end over per-tile site slots (tile = K1R_TILE positions): linked site pairs -> LD rows, device-side counts
struct isb_k3_tiles {
    int n_tiles;
    const int32_t *tile_first;        // [n_tiles] first site slot of the tile
    const int32_t *tile_cnt;          // [n_tiles] sites of the tile (slots are contiguous and position-ordered inside a tile)
    int64_t sites_cap;
    const int32_t *site_pos;          // [sites_cap] relative position
    isb_site_meta *meta;
    const int64_t *row_off;
    uint8_t *has2;
    const int4 *site_counts;          // [sites_cap] A,C,T,G counts of the site (M = 1)
    uint32_t *rows;
};
int isb_k3_backend_tiles(isb_ctx *ctx, const isb_reads_dev *rd, const isb_k3_tiles *ts, int64_t n_pairs, int32_t start, int32_t L,
                         const unsigned long long *nmask, const uint8_t *site_flags, int32_t n_splits, const int32_t *splits,
                         int min_snp, isb_ld_row *rows, int64_t cap);
int isb_k3_launch_reads(isb_ctx *ctx, const isb_reads_dev *rd, int64_t n_pairs, const uint8_t *pair_mm, int32_t start,
                        int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                        const uint8_t *site_flags, int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows,
                        int64_t cap);
// column-word batch (include/instrain_b200.h, isb_cols_batch), device pointers
struct isb_cols_dev {
    int64_t n_groups;                 // groups of ISB_COLS_GROUP = 64 positions (ISB_COLS_LANES = 8 column words)
    const int64_t *grp_off;           // [n_groups + 1] chunk offsets
    int64_t n_chunks;
    const uint32_t *words;            // [n_chunks][ISB_COLS_LANES][4]
    const int32_t *ids;               // same indexing: read-pair id of each word, -1 = padding
    int64_t n_nev;
    const int32_t *nev_pos;
    const int32_t *nev_pair;
};
// fused SNV call (M = 1) in the epilogue of K1c: everything isb_k2_launch would take
struct isb_k2_fuse {
    const uint8_t *ref;
    int min_cov;
    double min_freq;
    int32_t *covT;
    float *clonT;
    uint8_t *site_flags;
    isb_snv_row *rows;
    int64_t cap;
    int full_counts;                  // 1: write counts of every position; 0: only of flagged sites (what K3 reads)
    float *clonTR;                    // rarefied clonality (NULL: not computed)
    int cov_r;                        // rarefied_coverage
    uint64_t seed;
};
// init_nmask = false: nmask (when given) already holds the N-event bits of the range (chunked calls set it once per batch)
int isb_k1c_launch(isb_ctx *ctx, const isb_cols_dev *cd, const uint8_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                   int M, int32_t *counts, unsigned long long *nmask, const isb_k2_fuse *fuse, bool init_nmask = true);
int isb_cols_convert(isb_ctx *ctx, const isb_reads_dev *rd, int32_t start, int32_t L, int64_t *grp_off, uint32_t *words,
                     int32_t *ids, int64_t cap_chunks, int64_t *n_chunks);
int isb_k3_launch_cols(isb_ctx *ctx, const isb_cols_dev *cd, int64_t n_pairs, const uint8_t *pair_mm, int32_t start,
                       int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                       const uint8_t *site_flags, int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows,
                       int64_t cap, int32_t col_shift = 0);
int isb_k1r_n_events_launch(isb_ctx *ctx, int64_t n_nev, const int32_t *nev_pos, const int32_t *nev_pair, const uint8_t *pair_mm,
                            int64_t n_pairs, int32_t start, int32_t L, int M, unsigned long long *nmask);
int isb_k2_prepare(isb_ctx *ctx, double min_freq);
int isb_ensure(isb_ctx *ctx, int slot, size_t bytes);
int isb_k2_selftest_division(isb_ctx *ctx, int s_lo, int s_hi, unsigned long long *h_mismatches);
int isb_k4_launch(isb_ctx *ctx, int32_t L, int M, const int32_t *covT, const float *clonT, const unsigned long long *nmask,
                  int n_seg, const int32_t *seg_off, isb_summary_row *out);
int isb_tile_offsets(isb_ctx *ctx, const int32_t *ref_pos, int64_t n, int32_t start, int32_t L, int tp, int n_tiles);
int isb_k0r_launch(isb_ctx *ctx, int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, int64_t n_units, const uint16_t *base2,
                   const uint8_t *pass, int64_t *seg_word, int64_t n_words, uint32_t *words);
int isb_k0d_launch(isb_ctx *ctx, int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, int64_t n_units,
                   const uint8_t *pass, const uint8_t *ref, int32_t start, int32_t L, int64_t n_mis, const uint32_t *mis_word,
                   const uint8_t *mis_code, int64_t *seg_word, int64_t n_words, uint32_t *words);
int isb_k0_launch(isb_ctx *ctx, int64_t n, const int64_t *pos_off, const int32_t *id_base, const uint8_t *bqd,
                  int64_t n_esc, const int64_t *esc_evt, const int32_t *esc_id, int32_t start, int32_t L, int qpass,
                  int32_t *ref_pos, uint8_t *base, uint8_t *qual, int32_t *read_id);

// ---- small device helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t isb_lower_bound(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int64_t key)
{
    // first index in [lo, hi) with a[idx] >= key
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- TMA 1-D bulk copy + mbarrier (sm_90+ PTX; used by K1 and the staged K2) -----------------------------------
__device__ __forceinline__ uint32_t isb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void isb_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(isb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void isb_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(isb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void isb_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(isb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(isb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void isb_mbar_wait(uint64_t *bar, unsigned parity)
{
    const uint32_t addr = isb_smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void isb_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(isb_smem_u32(bar)) : "memory");
}
