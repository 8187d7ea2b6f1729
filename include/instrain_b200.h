/* instrain_b200.h -- C ABI of libinstrain_b200.so: the B200 (sm_100a) implementation of the
 * `inStrain profile` hot path (pileup counts -> per-site SNV call -> pairwise SNV linkage).
 *
 * The reference has no FFI: its seam is the Python call
 *     inStrain.profile.profile_bam(bam, Fdb, sR2M, ISP_loc, **kwargs)      inStrain/profile/__init__.py:7-18
 * and, one level lower, per split
 *     profile_split(samfile, scaffold, start, end, split_number, seq, R2M, null_model, **kwargs)
 *                                                                          inStrain/profile/profile_utilities.py:115-216
 * This library replaces the body of profile_split (every function listed below) for a whole *batch* of
 * scaffolds/splits at once; the Python shim `instrain_b200.profile` binds it with ctypes.
 *
 * Data model (one "batch" = any number of scaffolds concatenated into one int32 coordinate space):
 *   events   columnar, POSITION-MAJOR (sorted by ref_pos; within a position = pileup column order = BAM order):
 *              ref_pos int32, base uint8 (0..3 = A,C,T,G  -- inStrain's order, profile_utilities.py:34-35;
 *              4 = any other in-alignment base), qual uint8 (after htslib's mate-overlap tweak),
 *              read_id int32 (index of the read PAIR = the reference's `query_name` key of R2M)
 *   pair_mm  uint8[n_pairs]   R2M value (summed NM of the pair, filter_reads.py:917-928); 0 in set mode
 *   ref      uint8[L]         reference base codes (0..3, 4 = not A/C/G/T), upper-cased sequence (fasta.py:25-27)
 *   splits   int32[n_splits][2]  (start,end) inclusive, ascending, disjoint, in batch coordinates (fasta.py:56-73);
 *                             linkage never crosses a split (profile_utilities.py:165,184-185)
 * Only reads whose name is in R2M are packed (others are never counted: profile_utilities.py:277-283).
 *
 * Every pointer argument may be a HOST pointer or a DEVICE pointer (detected with cudaPointerGetAttributes);
 * host buffers are staged through context-owned device buffers.  No torch types cross this boundary.
 * All functions return 0 on success or a negative ISB_ERR_* code; isb_last_error() gives the message.
 * One context per GPU; a context is not thread-safe.  There is NO CPU fallback: without a CUDA device
 * isb_create() fails.
 */
#ifndef INSTRAIN_B200_H
#define INSTRAIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISB_ABI_VERSION 2

#define ISB_OK 0
#define ISB_ERR_CUDA (-1)        /* CUDA runtime error (message has the cudaError string) */
#define ISB_ERR_ARG (-2)         /* invalid argument (M out of range, null pointer, misaligned pointer ...) */
#define ISB_ERR_CAPACITY (-3)    /* an output row buffer is too small; n_* hold the required sizes */
#define ISB_ERR_ORDER (-4)       /* events are not position-major */
#define ISB_ERR_UNSUPPORTED (-5) /* input outside the supported envelope (e.g. a pair with >2 reads on one site) */

#define ISB_MAX_MM 64            /* mm levels M = max(pair_mm)+1 must be <= 64 */

/* site_flags byte written by isb_call_snvs: low nibble = the reference's `bases` set (bit b = base b was con/var
 * of a multi-allelic call at some mm), 0x10 = anySNP (snv_utilities.py:84-140): the site takes part in linkage. */
#define ISB_SITE_ANYSNP 0x10

/* class codes of isb_snv_row.cls (calc_snp_class, snv_utilities.py:198-223) */
enum { ISB_CLS_AMBIGUOUS_REFERENCE = 0, ISB_CLS_DIVERGENT_SITE = 1, ISB_CLS_SNS = 2, ISB_CLS_SNV = 3,
       ISB_CLS_CON_SNV = 4, ISB_CLS_POP_SNV = 5 };

/* One row of the reference's raw_snp_table (Stable, snv_utilities.py:119-130).  32 bytes. */
typedef struct {
    int32_t pos;          /* batch coordinate */
    int32_t cnt[4];       /* A,C,T,G counts cumulative over mm' <= mm (mm_counts_to_counts, profile_utilities.py:297-312) */
    int32_t mm;
    uint8_t ref;          /* reference base code */
    uint8_t con;          /* con_base */
    uint8_t var;          /* var_base */
    uint8_t allele_count; /* "morphia" */
    uint8_t cls;          /* ISB_CLS_* */
    uint8_t cryptic;      /* p2c[pos] (snv_utilities.py:137-144) */
    uint8_t pad[2];
} isb_snv_row;

/* One row of the reference's raw_linkage_table (_calc_ld_single, linkage.py:138-240).  64 bytes.
 * r2_normalized / d_prime_normalized (linkage.py:200-228) are the same statistics on min_snp haplotypes re-drawn from the
 * four observed frequencies.  The reference draws them with an UNSEEDED np.random.choice; here the draws come from a
 * counter-based generator keyed by (isb_params.seed, pos_a, pos_b, mm): reproducible, same distribution. */
typedef struct {
    int32_t pos_a, pos_b; /* batch coordinates, pos_a <= pos_b */
    int32_t mm;
    int32_t c_AB, c_Ab, c_aB, c_ab;
    uint8_t allele_A, allele_a, allele_B, allele_b;
    double r2, d_prime;   /* NaN where the reference yields np.nan */
    double r2_normalized, d_prime_normalized;
} isb_ld_row;

typedef struct isb_ctx isb_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
/* null_lut[t] = minimum alt count at coverage t, or -1 where the reference's model dict has no key t;
 * lut_default = model[-1]  (generate_snp_model, snv_utilities.py:14-38; lookup at :173-176). */
isb_ctx *isb_create(int device, const int32_t *null_lut, int n_lut, int lut_default);
void isb_destroy(isb_ctx *ctx);
const char *isb_last_error(const isb_ctx *ctx);   /* ctx may be NULL: error of the last failed isb_create */
int isb_abi_version(void);
/* Run on `stream` (a cudaStream_t) instead of the context's own stream; 0 restores the own stream. */
int isb_set_stream(isb_ctx *ctx, void *stream);
int isb_synchronize(isb_ctx *ctx);
/* Row counts of the last whole-path call, [n_snv, n_ld, n_sites, n_site_pairs], copied to `dst` (device or pinned host
 * memory, 4 x int64) by an asynchronous copy on the context's stream: with ISB_NO_SYNC a multi-GPU caller exchanges the
 * counts on the device and trims the gathered row slabs without ever waiting for the host (bench.py's table gather, the
 * counterpart of the result queue of the reference's task farm, profile_controller.py:243-271). */
int isb_row_counts_async(isb_ctx *ctx, int64_t *dst);

/* ---- stage K1: pileup counts ------------------------------------------------------------------------------- */
/* Replaces pysam's column iteration + get_base_counts_mm (profile_utilities.py:268-286):
 *   counts[p-start][mm][base] += 1  for every event with qual >= min_qual and base < 4
 *   nmask[p-start] |= 1<<mm         for every event with qual >= min_qual and base == 4  (the mm key that the
 *                                   reference's defaultdict creates before P2C[...] raises, :280-281)
 * counts is int32[L][M][4] and is overwritten (not accumulated); nmask is uint64[L] (may be NULL).
 * Events must be position-major unless ISB_K1_ANY_ORDER is set in `flags` (slower global-atomic path). */
#define ISB_K1_ANY_ORDER 0x1
int isb_pileup_counts(isb_ctx *ctx, int64_t n_events, const int32_t *ref_pos, const uint8_t *base,
                      const uint8_t *qual, const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm,
                      int32_t start, int32_t L, int M, int min_qual, uint32_t flags,
                      int32_t *counts, uint64_t *nmask);

/* ---- stage K2: SNV calling ----------------------------------------------------------------------------------- */
/* Replaces update_covT (profile_utilities.py:288-295), update_snp_table / call_snv_site / calc_snp_class /
 * calculate_clonality (snv_utilities.py:40-231) and is_present (readComparer.py:307-316) for L positions.
 *   covT[p][m]   = exact-mm coverage (int32)           clonT[p][m] = clonality of cumulative counts (float32, NaN = unset)
 *   site_flags[p] see ISB_SITE_ANYSNP                   rows: unordered; *n_rows = number produced (also when > cap)
 * clonTR (rarefied clonality) comes out of the whole-path entry points only (isb_result.clonTR). */
int isb_call_snvs(isb_ctx *ctx, int32_t L, int M, const int32_t *counts, const uint64_t *nmask, const uint8_t *ref,
                  int32_t start, int min_cov, double min_freq, int32_t *covT, float *clonT, uint8_t *site_flags,
                  isb_snv_row *rows, int64_t cap, int64_t *n_rows);

/* ---- stage K3: linkage --------------------------------------------------------------------------------------- */
/* Replaces update_linked_reads, calc_mm_SNV_linkage_network, calculate_ld, _iterator_ld_sites,
 * major_minor_allele and _calc_ld_single (linkage.py:14-240, 254-283; the normalized columns with seed 0).
 * Needs the position-major events again plus K1/K2 outputs.  rows unordered; *n_rows as above. */
int isb_linkage(isb_ctx *ctx, int64_t n_events, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start, int32_t L, int M,
                int min_qual, const int32_t *counts, const uint64_t *nmask, const uint8_t *site_flags,
                int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows, int64_t cap, int64_t *n_rows);

/* ---- whole path for one batch (what the profile_bam shim calls) ------------------------------------------------ */
typedef struct {
    int64_t n_events;
    const int32_t *ref_pos;
    const uint8_t *base;
    const uint8_t *qual;
    const int32_t *read_id;
    int64_t n_pairs;
    const uint8_t *pair_mm;
    int32_t start;            /* coordinate of position row 0 */
    int32_t L;
    const uint8_t *ref;
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
} isb_batch;

typedef struct {
    int32_t min_cov;          /* -c/--min_cov 5          (argumentParser.py:107) */
    int32_t min_snp;          /* --min_snp 20            (argumentParser.py:157) */
    int32_t min_qual;         /* min_base_quality=30     (profile_utilities.py:152) */
    uint32_t flags;           /* ISB_SKIP_LINKAGE ... */
    double min_freq;          /* -f/--min_freq 0.05      (argumentParser.py:109) */
    int32_t rarefied_cov;     /* --rarefied_coverage 50  (argumentParser.py:168): clonTR is computed where the cumulative
                               * coverage reaches it; 0 = no rarefied clonality */
    int32_t pad;
    uint64_t seed;            /* key of the counter-based generator behind clonTR and the normalized linkage columns */
} isb_params;
#define ISB_SKIP_LINKAGE 0x2   /* K1+K2 only (BASELINE config 2) */
#define ISB_NO_SYNC 0x4        /* all-device buffers only: enqueue and return; row counts valid after isb_synchronize */
#define ISB_PIPELINE 0x8       /* opt-in: overlap K2/K3 of one position chunk with K1 of the next on a second stream (batches
                                * >= 2^22 positions).  Event columns, measured on B200: +2 % at M = 1, -10 % at M = 15 (K1 is ~70 %
                                * issue-bound, so the co-running latency-bound kernels slow it down almost as much as they hide).
                                * Column words (isb_profile_cols): implemented after the last GPU session of round 1, not yet
                                * measured -- off by default. */

typedef struct {
    /* outputs (host or device); any may be NULL to skip the copy-out (the kernels still run) */
    int32_t *counts;          /* [L][M][4] */
    uint64_t *nmask;          /* [L] */
    int32_t *covT;            /* [L][M] */
    float *clonT;             /* [L][M] */
    uint8_t *site_flags;      /* [L] */
    isb_snv_row *snv;
    int64_t snv_cap;
    isb_ld_row *ld;
    int64_t ld_cap;
    /* filled by the call */
    int64_t n_snv;
    int64_t n_ld;
    int64_t n_sites;          /* linkage-eligible (anySNP) sites */
    int64_t n_site_pairs;     /* site pairs evaluated by the linkage kernel */
    /* rarefied clonality [L][M] (calculate_rarefied_clonality, snv_utilities.py:233-247: the clonality of rarefied_cov bases
     * re-drawn from the site's base frequencies; NaN below rarefied_cov).  Unseeded in the reference; here keyed by
     * (isb_params.seed, position, mm).  NULL or rarefied_cov == 0: not computed. */
    float *clonTR;
} isb_result;

int isb_profile_batch(isb_ctx *ctx, const isb_batch *in, const isb_params *prm, isb_result *out);

/* Same path fed with the PACKED transfer format (~1 B per event + 12 B per position instead of 10 B per event): the
 * host packer's compact form of the same event columns, expanded on the device by kernel K0 into the canonical
 * columns, then K1 -> K2 -> K3 unchanged.  This is what the end-to-end (host buffers in, host tables out) path uses,
 * because PCIe, not the kernels, bounds it.  Events of a position are sorted by pair id (stable w.r.t. BAM order, so
 * results are identical to the BAM-order columns).  See instrain_b200/csrc/isb_k0_expand.cu for the byte layout. */
typedef struct {
    int64_t n_events;
    const int64_t *pos_off;   /* [L+1] CSR offsets of the positions */
    const int32_t *id_base;   /* [L]   pair id of the first event of each position */
    const uint8_t *bqd;       /* [n]   bit7 = qual >= min_qual, bits 4-6 = base code, bits 0-3 = pair-id delta (15 = escape) */
    int64_t n_esc;
    const int64_t *esc_evt;   /* [n_esc] ascending event indices whose delta is > 14 */
    const int32_t *esc_id;    /* [n_esc] their absolute pair ids */
    int64_t n_pairs;
    const uint8_t *pair_mm;
    int32_t start;
    int32_t L;
    const uint8_t *ref;
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
    int32_t min_qual;         /* the threshold the quality bit was computed with; must equal isb_params.min_qual */
} isb_packed_batch;

int isb_profile_batch_packed(isb_ctx *ctx, const isb_packed_batch *in, const isb_params *prm, isb_result *out);

/* ---- READ-MAJOR input: aligned segments ------------------------------------------------------------------------------ */
/* The B200-first layout of the same information, as close to the BAM records as the path allows (what pysam hands
 * the reference one pileup column at a time, profile_utilities.py:150-153, stored once per READ instead of once per
 * column).  A segment is one CIGAR M/=/X block of a kept read: seg_len consecutive reference positions from seg_start,
 * one 4-bit ONE-HOT code per aligned base:  A = 1, C = 2, T = 4, G = 8 when the base survives htslib's base-quality
 * filter AFTER the mate-overlap tweak (i.e. it would be an event with qual >= min_qual), 0 when it does not (such a base
 * is not an event at all) or when it is not A/C/T/G.  Passing non-ACGT bases -- which only make their pair's mm level a
 * key of the position's MMcounts (the nmask bit) -- are listed separately (nev_pos / nev_pair; usually empty).
 * 4 bits per aligned base + ~22 bytes per segment instead of 10 bytes per event: 16x less HBM traffic for the pileup
 * and ~1.8x less PCIe traffic than the packed format.  The pileup kernel (K1f, isb_k1f_fused.cu) transposes on the fly:
 * every thread owns 8 consecutive positions and counts the codes of the segments that cover them with bit-sliced
 * (vertical, carry-save) counters -- no atomics.
 *
 * Layout rules (validated on the device; violations -> ISB_ERR_ORDER):
 *   - segments sorted by seg_start (ascending, ties in any order), all inside [start, start + L), 1 <= seg_len <=
 *     max_seg_len <= 256 (longer blocks are split by the packer);
 *   - words: the stream is POSITION-ALIGNED.  Word k of segment i covers the batch coordinates [8 * (seg_start[i] / 8
 *     + k), + 8): the base at coordinate p sits in nibble (p - seg_start[i]) + seg_start[i] % 8 of the segment's words,
 *     i.e. bits 4 * (n % 8).. of word seg_word[i] + n / 8; the nibbles before the first and after the last base are 0.  A
 *     segment has ceil((seg_start % 8 + seg_len) / 8) data words.  (A pileup thread that owns 8 consecutive
 *     coordinates therefore needs exactly one word of every read that covers them, unshifted.)  `start` must be a
 *     multiple of 8;
 *   - the data words of consecutive segments (in table order) are separated by one or two zero words, i.e.
 *     seg_word[0] >= 1, seg_word[i+1] = seg_word[i] + (data words of i) + 1 or + 2, the stream ends with a zero word and
 *     is padded with zero words to a multiple of 4 (n_words); `words` is 16-byte aligned. */
typedef struct {
    int64_t n_segs;
    const int32_t *seg_start;   /* [n_segs] batch coordinate of the first base */
    const uint16_t *seg_len;    /* [n_segs] */
    const int32_t *seg_pair;    /* [n_segs] read-pair id (index into pair_mm) */
    const int64_t *seg_word;    /* [n_segs] index of the segment's first data word */
    int64_t n_words;
    const uint32_t *words;      /* [n_words] nibble stream */
    int32_t max_seg_len;
    int32_t pad;
    int64_t n_nev;              /* passing non-ACGT read bases */
    const int32_t *nev_pos;     /* [n_nev] batch coordinate */
    const int32_t *nev_pair;    /* [n_nev] read-pair id */
    int64_t n_pairs;
    const uint8_t *pair_mm;     /* [n_pairs]; may be NULL when M == 1 */
    int32_t start;
    int32_t L;
    const uint8_t *ref;         /* [L] (not needed by isb_pileup_reads) */
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
    int32_t pad2;
} isb_reads_batch;

/* the pileup stage alone (K1f, unfused): counts[L][M][4] (+ nmask[L], may be NULL) from a read-major batch */
int isb_pileup_reads(isb_ctx *ctx, const isb_reads_batch *in, int32_t *counts, uint64_t *nmask);
/* The whole path on a read-major batch -- the default path.  M = 1 and no raw counts asked: K1f (pileup + single-allele
 * sites) -> k2q_sites (general SNV call on the site queue) -> k3f_site_rows -> linkage back end; otherwise K1f (counts) ->
 * K2 -> K3 (site events materialised from the segments).  Same results as
 * isb_profile_batch on the event columns of the same reads.  isb_params.min_qual is not used (the codes carry it). */
int isb_profile_reads(isb_ctx *ctx, const isb_reads_batch *in, const isb_params *prm, isb_result *out);

/* ---- READ-MAJOR input, compact TRANSFER format ------------------------------------------------------------------------ */
/* The same aligned segments in 3 bits per aligned base and without word offsets, for batches that cross PCIe (host
 * buffers): ~32 % fewer bytes than isb_reads_batch.  Segments are cut into units of 8 bases (the last unit of a segment
 * zero-padded); units are position-aligned like the words of isb_reads_batch (unit k of segment i covers the batch
 * coordinates [8 * (seg_start[i] / 8 + k), + 8)); unit k of segment i is element unit_off[i] + k of both unit arrays,
 * unit_off = exclusive prefix sum of ceil((seg_start % 8 + seg_len) / 8) in table order (implicit: not transferred).  Base j of a unit: 2-bit code (A,C,T,G = 0..3; inStrain's
 * order, profile_utilities.py:34-35) in bits 2j..2j+1 of base2, event bit j of pass (1 = the base survives htslib's
 * base-quality filter after the mate-overlap tweak and is A/C/T/G).  Passing non-ACGT bases go to nev_pos / nev_pair as
 * in isb_reads_batch.  K0r (isb_k0r_expand.cu) rebuilds seg_word and the canonical nibble stream in device memory, then
 * the kernels of isb_profile_reads run: results are identical.  Same ordering / range rules for the segments. */
typedef struct {
    int64_t n_segs;
    const int32_t *seg_start;   /* [n_segs] ascending */
    const uint16_t *seg_len;    /* [n_segs] 1 .. max_seg_len */
    const int32_t *seg_pair;    /* [n_segs] */
    int64_t n_units;            /* sum of ceil((seg_start % 8 + seg_len) / 8) */
    const uint16_t *base2;      /* [n_units] */
    const uint8_t *pass;        /* [n_units] */
    int32_t max_seg_len;
    int32_t pad;
    int64_t n_nev;
    const int32_t *nev_pos;
    const int32_t *nev_pair;
    int64_t n_pairs;
    const uint8_t *pair_mm;     /* [n_pairs]; may be NULL when M == 1 */
    int32_t start;
    int32_t L;
    const uint8_t *ref;         /* [L] */
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
    int32_t pad2;
} isb_reads_compact;

int isb_profile_reads_compact(isb_ctx *ctx, const isb_reads_compact *in, const isb_params *prm, isb_result *out);

/* ---- READ-MAJOR input, reference-delta TRANSFER format ------------------------------------------------------------------ */
/* The smallest form of the same aligned segments for batches that cross PCIe: reads are almost everywhere identical to
 * the reference, which the call receives anyway.  Per unit of 8 bases (units exactly as in isb_reads_compact) only the
 * event bits `pass` travel (1 byte); every passing A/C/T/G base that DIFFERS from the reference base at its position
 * (or whose reference base is not A/C/T/G) is listed once: mis_word = index of its word in the canonical nibble stream
 * (1 + unit index + segment index: one leading zero word, one separator per segment), mis_code = bits 4-6 the nibble
 * position in the word, bits 0-3 the XOR of the one-hot codes of the reference base (0 if not A/C/T/G) and of the read
 * base.  ~1.1 bits per aligned base + 5 bytes per mismatch instead of 3 bits per base: about half the bytes of
 * isb_reads_compact at 1 % divergence.  K0d (isb_k0r_expand.cu) rebuilds the nibble stream on the device (reference
 * codes masked by the event bits, then the listed nibbles flipped), then the kernels of isb_profile_reads run:
 * results are identical. */
typedef struct {
    int64_t n_segs;
    const int32_t *seg_start;   /* [n_segs] ascending */
    const uint16_t *seg_len;    /* [n_segs] 1 .. max_seg_len */
    const int32_t *seg_pair;    /* [n_segs] */
    int64_t n_units;            /* sum of ceil((seg_start % 8 + seg_len) / 8) */
    const uint8_t *pass;        /* [n_units] */
    int64_t n_mis;
    const uint32_t *mis_word;   /* [n_mis] */
    const uint8_t *mis_code;    /* [n_mis] */
    int32_t max_seg_len;
    int32_t pad;
    int64_t n_nev;
    const int32_t *nev_pos;
    const int32_t *nev_pair;
    int64_t n_pairs;
    const uint8_t *pair_mm;     /* [n_pairs]; may be NULL when M == 1 */
    int32_t start;
    int32_t L;
    const uint8_t *ref;         /* [L] */
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
    int32_t pad2;
} isb_reads_delta;

int isb_profile_reads_delta(isb_ctx *ctx, const isb_reads_delta *in, const isb_params *prm, isb_result *out);
/* The encoder on the host (C++, no GPU; what the host packer runs after isb_pack_scaffold_reads): pass[n_units] and the
 * mismatch entries of a read-major batch against ref[L] (reference codes of the batch coordinates start .. start + L).
 * Returns the number of entries -- only the first cap_mis are stored (call again with larger buffers when it is larger;
 * mis_word / mis_code may be NULL for a counting call) -- or -1 when the segments violate the layout rules of
 * isb_reads_batch or n_units is not the sum of their unit counts. */
int64_t isb_reads_delta_host(int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, const int64_t *seg_word,
                             const uint32_t *words_in, int64_t n_words_in, int32_t start, int32_t L, const uint8_t *ref,
                             uint8_t *pass, int64_t n_units, uint32_t *mis_word, uint8_t *mis_code, int64_t cap_mis);

/* ---- COLUMN-WORD input: the pileup-major form of the aligned segments ------------------------------------------------- */
/* The same one-hot nibble words as isb_reads_batch (one 32-bit word = the codes of 8 consecutive, 8-aligned batch
 * coordinates of ONE read), stored where the pileup needs them instead of where the read is: per COLUMN WORD (8
 * positions) the words of all reads covering it.  This is the transposition pysam's pileup engine performs column by
 * column in the reference (profile_utilities.py:150-153, :275), done once by the packer at word granularity -- the
 * north-star "columnar, position-major" layout with 4 bits per aligned base instead of 10 bytes.
 *   - position p (relative to `start`, a multiple of 8) lies in column word c = p / 8, GROUP g = c / 8 (64 positions),
 *     lane c % 8;
 *   - group g owns chunks [grp_off[g], grp_off[g+1]); a chunk is 8 lanes x 8 words (256 bytes): 8 consecutive slots of
 *     each of the group's 8 columns, a column's 8 slots contiguous (one 32-byte sector).  Slot i of column (g, lane) is
 *     word ((grp_off[g] + i / 8) * 8 + lane) * 8 + i % 8 of `words`, its read-pair id the same element of `ids`.  Every
 *     column of a group is padded to the group's depth 8 * (grp_off[g+1] - grp_off[g]) with word 0 / id -1 (8
 *     neighbouring columns see almost the same reads: ~8 % padding at 100x);
 *   - the words of a column keep the table order of their segments (BAM order);
 *   - passing non-ACGT read bases go to nev_pos / nev_pair as in isb_reads_batch.
 * K1c (isb_k1c_cols.cu) streams it: a warp takes 4 consecutive groups, lane = column word, one 256-bit load per lane and
 * chunk (a warp instruction reads 1 KB: four whole 256-byte chunks), bit-sliced counting, no shared memory, atomics or
 * searches -- an HBM-bound kernel on 0.5 B per aligned base.  The linkage stage reads the entries of an SNV site as one
 * nibble of each word of its column: whole 32-byte sectors at addresses that follow from the position alone. */
#define ISB_COLS_LANES 8                                  /* column words per group */
#define ISB_COLS_UNIT 8                                   /* consecutive slots of one column per chunk (32 bytes) */
#define ISB_COLS_GROUP (8 * ISB_COLS_LANES)               /* positions per group */
#define ISB_COLS_CHUNK (ISB_COLS_UNIT * ISB_COLS_LANES)   /* words per chunk */
typedef struct {
    int64_t n_groups;           /* ceil(L / ISB_COLS_GROUP) */
    const int64_t *grp_off;     /* [n_groups + 1] chunk offsets, grp_off[0] = 0 */
    int64_t n_chunks;           /* = grp_off[n_groups] */
    const uint32_t *words;      /* [n_chunks * ISB_COLS_CHUNK], 32-byte aligned */
    const int32_t *ids;         /* [n_chunks * ISB_COLS_CHUNK], 32-byte aligned; may be NULL for isb_pileup_cols at M == 1 */
    int64_t n_nev;
    const int32_t *nev_pos;
    const int32_t *nev_pair;
    int64_t n_pairs;
    const uint8_t *pair_mm;     /* [n_pairs]; may be NULL when M == 1 */
    int32_t start;
    int32_t L;
    const uint8_t *ref;         /* [L] (not needed by isb_pileup_cols) */
    int32_t n_splits;
    const int32_t *splits;
    int32_t M;
    int32_t pad;
} isb_cols_batch;

/* Layout conversion on the device: read-major batch -> column words.  grp_off[n_groups + 1] is always written and
 * *n_chunks set; words / ids (room for cap_chunks chunks) are filled when given (NULL: sizing call).  ISB_ERR_CAPACITY
 * when cap_chunks < *n_chunks.  Host or device pointers. */
int isb_cols_from_reads(isb_ctx *ctx, const isb_reads_batch *in, int64_t *grp_off, int64_t *n_chunks, uint32_t *words,
                        int32_t *ids, int64_t cap_chunks);
/* The same conversion on the host (C++, no GPU; what the host packer runs after isb_pack_scaffold_reads).  Returns the
 * number of chunks, or -1 when the segments violate the layout rules of isb_reads_batch; words / ids may be NULL
 * (sizing call) and are not written when cap_chunks is too small. */
int64_t isb_cols_from_reads_host(int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, const int32_t *seg_pair,
                                 const int64_t *seg_word, const uint32_t *words_in, int64_t n_words_in, int32_t start,
                                 int32_t L, int64_t *grp_off, uint32_t *words, int32_t *ids, int64_t cap_chunks);
/* stage K1c alone: counts[L][M][4] (+ nmask[L], may be NULL) from a column-word batch */
int isb_pileup_cols(isb_ctx *ctx, const isb_cols_batch *in, int32_t *counts, uint64_t *nmask);
/* K1c -> K2 -> K3 on a column-word batch; same results as isb_profile_reads / isb_profile_batch on the same reads.
 * When M == 1 and the caller asks for neither counts nor nmask (isb_result.counts == isb_result.nmask == NULL), the SNV
 * call runs in the epilogue of the pileup kernel (one pass: the counts of a position never leave the registers,
 * except at linkage sites). */
int isb_profile_cols(isb_ctx *ctx, const isb_cols_batch *in, const isb_params *prm, isb_result *out);

/* ---- stage K4: merge-stage summary reductions (the row after the hot path) ----------------------------- */
/* Numeric core of make_coverage_table (profile_utilities.py:425-506) with mm_counts_to_counts_shrunk (:508-532) and
 * get_basewise_clons (:534-546): per scaffold s (positions [scaffold_off[s], scaffold_off[s+1]) of covT / clonT) and mm
 * level m, over cumulative coverage c_m(p) = sum_{m'<=m} covT[p][m'] and the clonality of the highest level <= m at
 * which clonT is set.  out[s*M + m]; the host derives breadth, mean, std, SEM, nucl_diversity, ANI from these exact
 * sums (instrain_b200/summary.py).  Medians are exact order statistics (4-pass radix select). */
typedef struct {
    int64_t length;        /* positions of the scaffold */
    int64_t nonzero;       /* positions with c_m > 0                         -> breadth */
    int64_t sum_cov;       /* sum of c_m                                      -> coverage (mean) */
    uint64_t sum_cov2;     /* sum of c_m^2                                    -> coverage_std / coverage_SEM */
    int64_t counted;       /* positions with a clonality value                -> breadth_minCov, ANI denominators */
    double sum_clon;       /* sum of those clonalities                        -> nucl_diversity = 1 - mean */
    int32_t cov_med_lo, cov_med_hi;   /* the two middle order statistics of c_m (equal when length is odd) */
    float clon_med_lo, clon_med_hi;   /* same for the clonality values (NaN when counted == 0) */
    int32_t present;       /* level m is a key of the scaffold's covT (the table has a row for it) */
    int32_t pad;
} isb_summary_row;

int isb_scaffold_summary(isb_ctx *ctx, int32_t L, int M, const int32_t *covT, const float *clonT, const uint64_t *nmask,
                         int32_t n_scaffolds, const int32_t *scaffold_off, isb_summary_row *out);

/* number of kernels this library has launched on the context since creation (bench.py's gpu_launches) */
int64_t isb_launch_count(const isb_ctx *ctx);

/* Measurement hooks (bench.py's roofline leg): when enabled, isb_profile_batch brackets K1 / K2 / K3 with CUDA
 * events on the context's stream.  isb_stage_times synchronises, returns the summed device milliseconds and call
 * counts per stage (index 0 = K1 pileup, 1 = K2 SNV, 2 = K3 linkage) since the last call, and resets them. */
/* Self-test of K2's fast correctly-rounded quotient (one reciprocal + 3 FP64 ops instead of an IEEE division): number of
 * pairs (c, s), s_lo <= s <= s_hi, 0 <= c <= s, whose result differs bitwise from the IEEE division; -1 on error. */
int64_t isb_selftest_division(isb_ctx *ctx, int s_lo, int s_hi);
int isb_enable_timing(isb_ctx *ctx, int on);
int isb_stage_times(isb_ctx *ctx, double ms[3], int64_t calls[3]);

/* ---- host packer (C++, no GPU): BAM -> position-major event columns ------------------------------------------------ */
/* Replaces pysam.AlignmentFile(bam) (profile_utilities.py:56) and the htslib pileup engine behind
 * samfile.pileup(..., stepper='nofilter', ignore_overlaps=True, ...) (profile_utilities.py:150-153): BGZF inflate, BAM
 * record decode, htslib 1.10's mate-overlap quality tweak (overlap_push / tweak_overlap_quality, including the
 * cigar_iref2iseq_next in-block counter behaviour the reference's goldens pin), CIGAR expansion of M/=/X bases and a
 * stable counting sort into position-major order.  The BAM must be coordinate-sorted (as inStrain requires). */
void *isb_bam_open(const char *path);                      /* NULL on failure */
void isb_bam_close(void *bam);
int isb_bam_n_refs(void *bam);
const char *isb_bam_ref_name(void *bam, int tid);
int64_t isb_bam_ref_len(void *bam, int tid);
const char *isb_bam_error(void *bam);
/* next record's tid; -1 unmapped tail; -2 clean end of file; -3 read error: a corrupt or truncated file (bad BGZF block,
 * CRC mismatch, record cut short, missing BGZF end-of-file block) -- pysam / htslib raise there, so must the caller
 * (isb_bam_error has the reason).  isb_pack_scaffold* return NULL and isb_filter_open returns NULL in that case. */
int isb_bam_peek_tid(void *bam);
const char *isb_host_last_error(void);                     /* reason of the last failed isb_filter_open on this thread */
/* Reposition at a BGZF virtual offset taken from the BAM's .bai index (first alignment of a scaffold): several readers of
 * one BAM, one per host thread, can then pack different scaffolds concurrently.  0 on success. */
int isb_bam_seek(void *bam, uint64_t voffset);
/* Consume all records of scaffold `tid`; pack the reads whose name is in the list (names_blob + name_off[n_names+1];
 * name_mm[i] = R2M value, 0 in set mode).  Positions are shifted by pos_offset, pair ids start at pair_id_offset and
 * follow BAM order of first appearance.  Returns an events handle (NULL on error). */
void *isb_pack_scaffold(void *bam, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                        const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset);
int64_t isb_events_count(void *events);
int64_t isb_events_pairs(void *events);
int64_t isb_events_reads_seen(void *events);
int64_t isb_events_reads_packed(void *events);
void isb_events_copy(void *events, int32_t *ref_pos, uint8_t *base, uint8_t *qual, int32_t *read_id, uint8_t *pair_mm);
void isb_events_free(void *events);

/* Same reads as a READ-MAJOR batch (isb_reads_batch): one segment per CIGAR M/=/X block (clipped to the scaffold, split at
 * 256 bases), sorted by start, one-hot codes for the bases with quality >= min_qual after the overlap tweak, passing
 * non-ACGT bases in the N-event list.  isb_reads_stream_words() words hold the scaffold's stream ([data words + one zero
 * word] per segment); a batch stream is one leading zero word + the scaffolds' streams + zero padding to a multiple of
 * 4 words, and isb_reads_copy() rebases seg_word to word_base = the index where this scaffold's stream is placed. */
void *isb_pack_scaffold_reads(void *bam, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                              const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset, int min_qual);
/* The same for the reads that overlap the scaffold positions [lo, hi) only (hi <= lo: the whole scaffold): what an index
 * fetch of that region hands htslib's pileup (inStrain/polymorpher.py:287-293).  Reading stops at the first record that
 * starts at or behind hi, so the reader is left inside the scaffold: seek (isb_bam_seek) or close it afterwards.  Used to
 * profile ONE large scaffold by contiguous runs of its splits on several GPUs (instrain_b200.profile.profile_scaffold_run). */
void *isb_pack_scaffold_reads_region(void *bam, int tid, int64_t n_names, const char *names_blob, const int64_t *name_off,
                                     const uint8_t *name_mm, int32_t pos_offset, int32_t pair_id_offset, int min_qual,
                                     int64_t lo, int64_t hi);
int64_t isb_reads_segs(void *reads);
int64_t isb_reads_stream_words(void *reads);
int64_t isb_reads_pairs(void *reads);
int64_t isb_reads_n_events(void *reads);
int64_t isb_reads_nev(void *reads);
int isb_reads_max_len(void *reads);
int64_t isb_reads_reads_seen(void *reads);
int64_t isb_reads_reads_packed(void *reads);
void isb_reads_copy(void *reads, int32_t *seg_start, uint16_t *seg_len, int32_t *seg_pair, int64_t *seg_word,
                    uint32_t *words, int32_t *nev_pos, int32_t *nev_pair, uint8_t *pair_mm, int64_t word_base);
void isb_reads_free(void *reads);

/* ---- host read filter (C++, no GPU): BAM -> sR2M --------------------------------------------------------------------- */
/* Default configuration of the reference's read filter: get_paired_reads (filter_reads.py:885-956),
 * paired_read_filter with pairing_filter='paired_only' (:471-532), filter_scaff2pair2info / evaluate_pair (:201-300,
 * :387-426).  isb_filter_open makes one pass over the BAM; isb_filter_apply applies the thresholds (reference defaults:
 * min_read_ani 0.95, min_mapq -1, max_insert_relative 3, min_insert 50); the kept pairs of a scaffold (names in file
 * order + summed NM) are what isb_pack_scaffold takes as R2M.  tally[6] = pass_pairing_filter, pass_min_read_ani,
 * pass_max_insert, pass_min_insert, pass_min_mapq, filtered_pairs (the mapping_info columns). */
void *isb_filter_open(const char *bam_path);
/* The same pass on n_threads host threads, one reader each, scaffolds taken from a shared counter: first_voffset[tid] = BGZF
 * virtual offset of the scaffold's first alignment (from the .bai index; 0 = no alignments).  Identical tables. */
void *isb_filter_open_mt(const char *bam_path, int n_threads, int n_refs, const uint64_t *first_voffset);
int64_t isb_filter_apply(void *filter, double min_read_ani, int min_mapq, double max_insert_relative, int min_insert);
/* All pairing filters of the reference + priority reads (paired_read_filter, filter_reads.py:471-532): pairing_mode 0 =
 * paired_only (isb_filter_apply), 1 = non_discordant, 2 = all_reads (mates on two scaffolds are merged: summed NM, insert
 * -2); the n_priority names pass the pairing filter regardless.  tally2[3] = unfiltered_priority_reads,
 * filtered_singletons, filtered_priority_reads.  Returns the number of kept names, -1 for an unknown mode. */
int64_t isb_filter_apply2(void *filter, double min_read_ani, int min_mapq, double max_insert_relative, int min_insert,
                          int pairing_mode, int64_t n_priority, const char *names_blob, const int64_t *name_off);
void isb_filter_tally2(void *filter, int tid, int64_t tally2[3]);
int isb_filter_n_refs(void *filter);
double isb_filter_max_insert(void *filter);
void isb_filter_tally(void *filter, int tid, int64_t tally[6]);
/* The other per-scaffold columns of mapping_info: stats[10] = unfiltered_reads, unfiltered_pairs, unfiltered_singletons,
 * mean_mistmaches, mean_insert_distance, mean_mapq_score, mean_pair_length, mean_PID, median_insert (over the pairs that
 * pass the pairing filter; NaN when there are none), and that number of pairs. */
void isb_filter_stats(void *filter, int tid, double stats[10]);
/* The same for any pairing filter: the means / median run over the entries the last isb_filter_apply2 selected (singletons,
 * priority reads and pairs merged over two scaffolds included when the mode keeps them; filter_reads.py:262-276). */
void isb_filter_stats2(void *filter, int tid, double stats[10]);
int64_t isb_filter_n_pairs(void *filter, int tid);
int64_t isb_filter_names_bytes(void *filter, int tid);
void isb_filter_copy(void *filter, int tid, char *names_blob, int64_t *name_off, int32_t *mm);
void isb_filter_free(void *filter);

#ifdef __cplusplus
}
#endif
#endif /* INSTRAIN_B200_H */
