/* oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Plain-C CPU restatement of the inStrain `profile` hot path on columnar event arrays:
 *   orc_pileup_counts  <- get_base_counts_mm            inStrain/profile/profile_utilities.py:268-286
 *                         update_covT                   inStrain/profile/profile_utilities.py:288-295
 *   orc_call_snvs      <- update_snp_table              inStrain/profile/snv_utilities.py:40-145
 *                         call_snv_site                 inStrain/profile/snv_utilities.py:147-196
 *                         calc_snp_class                inStrain/profile/snv_utilities.py:198-223
 *                         calculate_clonality           inStrain/profile/snv_utilities.py:225-231
 *                         is_present                    inStrain/readComparer.py:307-316
 *                         mm_counts_to_counts           inStrain/profile/profile_utilities.py:297-312
 *   orc_linkage        <- update_linked_reads           inStrain/profile/linkage.py:254-283
 *                         calc_mm_SNV_linkage_network   inStrain/profile/linkage.py:14-44
 *                         calculate_ld/_iterator_ld_sites  inStrain/profile/linkage.py:46-131
 *                         major_minor_allele            inStrain/profile/linkage.py:133-136
 *                         _calc_ld_single (non-random part) inStrain/profile/linkage.py:138-198
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the reference's own functions
 * (oracle/ref_harness.py, build container) and against fixtures derived from the reference's
 * golden tables (tests/golden/).  Excluded by design (unseeded RNG in the reference): clonTR,
 * r2_normalized, d_prime_normalized (snv_utilities.py:233-247, linkage.py:200-228).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle.c -o _build/liboracle.so   (oracle/build.py)
 *
 * Conventions: bases 0..3 = A,C,T,G (profile_utilities.py:34-35); 4 = any other in-alignment base.
 * Events must be grouped by position (position-major); within a position the array order is the
 * pileup column order (BAM file order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t pos;
    int32_t cnt[4];      /* cumulative (mm' <= mm) A,C,T,G counts */
    int32_t mm;
    uint8_t ref;         /* 0..3 or 4 (not ACGT) */
    uint8_t con;
    uint8_t var;
    uint8_t allele_count;
    uint8_t cls;         /* 0 AmbiguousReference 1 DivergentSite 2 SNS 3 SNV 4 con_SNV 5 pop_SNV */
    uint8_t cryptic;
    uint8_t pad[2];
} orc_snv_row;           /* 32 bytes */

typedef struct {
    int32_t pos_a, pos_b, mm;
    int32_t c_AB, c_Ab, c_aB, c_ab;
    uint8_t allele_A, allele_a, allele_B, allele_b;
    double r2, d_prime;
} orc_ld_row;            /* 48 bytes */

enum { SITE_ANYSNP = 0x10 };   /* site_flags: low nibble = `bases` set (bit b), 0x10 = anySNP */

/* ---------------------------------------------------------------- K1 */
/* counts[L][M][4] += 1 for events with qual >= min_qual and base < 4; nmask[p] bit mm set for
 * qualifying events with base == 4 (the defaultdict side effect at profile_utilities.py:280-281:
 * table[mm] is created before P2C[...] raises KeyError). */
int orc_pileup_counts(int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                      const int32_t *read_id, const int32_t *pair_mm, int32_t start, int32_t L, int M,
                      int min_qual, int32_t *counts, uint64_t *nmask)
{
    memset(counts, 0, sizeof(int32_t) * (size_t)L * M * 4);
    memset(nmask, 0, sizeof(uint64_t) * (size_t)L);
    for (int64_t e = 0; e < n; ++e) {
        int32_t p = ref_pos[e] - start;
        if (p < 0 || p >= L || qual[e] < min_qual) continue;
        int mm = pair_mm[read_id[e]];
        if (mm < 0 || mm >= M) return -1;
        if (base[e] < 4) counts[((size_t)p * M + mm) * 4 + base[e]] += 1;
        else nmask[p] |= (uint64_t)1 << mm;
    }
    return 0;
}

/* ---------------------------------------------------------------- K2 */
static inline int lut_get(const int32_t *lut, int n_lut, int lut_default, int64_t total)
{
    /* `if total in model: model[total] else model[-1]`  (snv_utilities.py:173-176); lut[t] < 0 = key absent */
    if (total >= 0 && total < n_lut && lut[total] >= 0) return lut[total];
    return lut_default;
}

static inline int argmax4(const int64_t *c)
{
    int b = 0;
    for (int i = 1; i < 4; ++i) if (c[i] > c[b]) b = i;   /* np.argmax: first maximum */
    return b;
}

/* Returns number of SNV rows written (or -needed if cap too small). */
int64_t orc_call_snvs(int32_t L, int M, const int32_t *counts, const uint64_t *nmask, const uint8_t *ref,
                      const int32_t *lut, int n_lut, int lut_default, int min_cov, double min_freq,
                      int32_t start, int32_t *covT, float *clonT, uint8_t *site_flags,
                      orc_snv_row *rows, int64_t cap)
{
    int64_t n_rows = 0;
    for (int32_t p = 0; p < L; ++p) {
        int64_t C[4] = {0, 0, 0, 0};
        int any_snp = 0, cryptic = 0;
        unsigned bases = 0;
        int64_t first_row = n_rows;
        for (int m = 0; m < M; ++m) {
            const int32_t *E = counts + ((size_t)p * M + m) * 4;
            int64_t e_sum = (int64_t)E[0] + E[1] + E[2] + E[3];
            covT[(size_t)p * M + m] = (int32_t)e_sum;
            clonT[(size_t)p * M + m] = NAN;
            int present = e_sum > 0 || ((nmask[p] >> m) & 1);
            if (!present) continue;                      /* mm not a key of MMcounts */
            for (int b = 0; b < 4; ++b) C[b] += E[b];
            int64_t T = C[0] + C[1] + C[2] + C[3];
            if (T >= min_cov) {                          /* snv_utilities.py:95-96, 225-231 */
                double s = (double)T;
                double prob = ((double)C[0] / s) * ((double)C[0] / s) + ((double)C[1] / s) * ((double)C[1] / s)
                            + ((double)C[2] / s) * ((double)C[2] / s) + ((double)C[3] / s) * ((double)C[3] / s);
                clonT[(size_t)p * M + m] = (float)prob;
            }
            if (T < min_cov) continue;                   /* call_snv_site -> (None, 0) */
            int thr = lut_get(lut, n_lut, lut_default, T);
            int i = 0;
            for (int b = 0; b < 4; ++b)
                if (C[b] >= thr && (double)C[b] / (double)T >= min_freq) ++i;
            int con = argmax4(C);
            int is_row = (i > 1) || (i == 1 && con != ref[p]) || (i == 0);
            if (!is_row) {                               /* snp == -1 */
                if (any_snp) cryptic = 1;
                continue;
            }
            int64_t tmp[4] = {C[0], C[1], C[2], C[3]};
            tmp[con] = 0;
            int var = argmax4(tmp);                      /* first index of the max after zeroing con (:110-112) */
            int cls;
            if (ref[p] > 3) cls = 0;
            else if (i == 0) cls = 1;
            else if (i == 1) cls = 2;
            else if (ref[p] == con) cls = 3;
            else if (ref[p] == var) cls = 4;
            else {
                int64_t cr = C[ref[p]];
                cls = (cr >= thr && (double)cr / (double)T >= min_freq) ? 4 : 5;
            }
            if (n_rows < cap) {
                orc_snv_row *r = rows + n_rows;
                memset(r, 0, sizeof(*r));
                r->pos = p + start;
                for (int b = 0; b < 4; ++b) r->cnt[b] = (int32_t)C[b];
                r->mm = m; r->ref = ref[p]; r->con = (uint8_t)con; r->var = (uint8_t)var;
                r->allele_count = (uint8_t)i; r->cls = (uint8_t)cls;
            }
            ++n_rows;
            if (i >= 2) { any_snp = 1; bases |= (1u << con) | (1u << var); }
            else if (i == 1 && any_snp) cryptic = 1;
        }
        if (cryptic)
            for (int64_t r = first_row; r < n_rows && r < cap; ++r) rows[r].cryptic = 1;
        site_flags[p] = (uint8_t)(bases | (any_snp ? SITE_ANYSNP : 0));
    }
    return n_rows <= cap ? n_rows : -n_rows;
}

/* ---------------------------------------------------------------- K3 */
typedef struct { int32_t i, j, mm; uint8_t b1, b2; } combo_t;

static int combo_cmp(const void *x, const void *y)
{
    const combo_t *a = (const combo_t *)x, *b = (const combo_t *)y;
    if (a->i != b->i) return a->i < b->i ? -1 : 1;
    if (a->j != b->j) return a->j < b->j ? -1 : 1;
    if (a->mm != b->mm) return a->mm < b->mm ? -1 : 1;
    return 0;
}

typedef struct { int32_t site; uint8_t base; } entry_t;

static void major_minor(const int64_t *c, int *maj, int *mnr)
{
    /* sorted(d, key=d.get, reverse=True)[:2] -- stable, ties keep A,C,T,G order (linkage.py:133-136) */
    int idx[4] = {0, 1, 2, 3};
    for (int a = 1; a < 4; ++a) {
        int v = idx[a], k = a - 1;
        while (k >= 0 && c[idx[k]] < c[v]) { idx[k + 1] = idx[k]; --k; }
        idx[k + 1] = v;
    }
    *maj = idx[0]; *mnr = idx[1];
}

static void cum_counts(const int32_t *counts, const uint64_t *nmask, int M, int32_t p, int m, int64_t *C, int *present)
{
    C[0] = C[1] = C[2] = C[3] = 0;
    for (int k = 0; k <= m; ++k) {
        const int32_t *E = counts + ((size_t)p * M + k) * 4;
        for (int b = 0; b < 4; ++b) C[b] += E[b];
    }
    const int32_t *E = counts + ((size_t)p * M + m) * 4;
    *present = ((int64_t)E[0] + E[1] + E[2] + E[3]) > 0 || ((nmask[p] >> m) & 1);
}

/* Events position-major.  splits: n_splits x (start, end) inclusive, in the same coordinate space as
 * ref_pos (start = coordinate of counts row 0).  Returns #rows (or -needed). */
int64_t orc_linkage(int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                    const int32_t *read_id, const int32_t *pair_mm, int64_t n_pairs, int32_t start, int32_t L,
                    int M, int min_qual, const int32_t *counts, const uint64_t *nmask, const uint8_t *site_flags,
                    int n_splits, const int32_t *splits, int min_snp, orc_ld_row *rows, int64_t cap)
{
    int64_t n_rows = 0;
    /* per-pair entry lists (linked through arrays), rebuilt per split */
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
    int32_t *tail = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
    for (int64_t k = 0; k < n_pairs; ++k) head[k] = -1;
    size_t ent_cap = 1 << 16, cmb_cap = 1 << 16;
    entry_t *ent = (entry_t *)malloc(sizeof(entry_t) * ent_cap);
    int32_t *ent_next = (int32_t *)malloc(sizeof(int32_t) * ent_cap);
    combo_t *cmb = (combo_t *)malloc(sizeof(combo_t) * cmb_cap);
    int64_t e0 = 0;
    for (int s = 0; s < n_splits; ++s) {
        int32_t s_lo = splits[2 * s], s_hi = splits[2 * s + 1];
        /* events are position-major: advance to the split */
        while (e0 < n && ref_pos[e0] < s_lo) ++e0;
        size_t n_ent = 0; int64_t n_touched = 0;
        int64_t e = e0;
        for (; e < n && ref_pos[e] <= s_hi; ++e) {
            int32_t p = ref_pos[e] - start;
            if (p < 0 || p >= L) continue;
            uint8_t f = site_flags[p];
            if (!(f & SITE_ANYSNP) || qual[e] < min_qual || base[e] > 3 || !((f >> base[e]) & 1)) continue;
            int32_t rid = read_id[e];
            if (n_ent == ent_cap) {
                ent_cap *= 2;
                ent = (entry_t *)realloc(ent, sizeof(entry_t) * ent_cap);
                ent_next = (int32_t *)realloc(ent_next, sizeof(int32_t) * ent_cap);
            }
            ent[n_ent].site = p; ent[n_ent].base = base[e]; ent_next[n_ent] = -1;
            if (head[rid] < 0) { head[rid] = (int32_t)n_ent; touched[n_touched++] = rid; }
            else ent_next[tail[rid]] = (int32_t)n_ent;
            tail[rid] = (int32_t)n_ent;
            ++n_ent;
        }
        e0 = e;   /* positions belong to exactly one split; a read pair straddling a boundary simply has
                   * entries in both splits' lists (read_to_snvs is per split, profile_utilities.py:165) */
        size_t n_cmb = 0;
        for (int64_t t = 0; t < n_touched; ++t) {
            int32_t rid = touched[t];
            int32_t mm = pair_mm[rid];
            for (int32_t a = head[rid]; a >= 0; a = ent_next[a])
                for (int32_t b = ent_next[a]; b >= 0; b = ent_next[b]) {
                    if (n_cmb == cmb_cap) { cmb_cap *= 2; cmb = (combo_t *)realloc(cmb, sizeof(combo_t) * cmb_cap); }
                    cmb[n_cmb].i = ent[a].site; cmb[n_cmb].j = ent[b].site; cmb[n_cmb].mm = mm;
                    cmb[n_cmb].b1 = ent[a].base; cmb[n_cmb].b2 = ent[b].base;
                    ++n_cmb;
                }
            head[rid] = -1;
        }
        qsort(cmb, n_cmb, sizeof(combo_t), combo_cmp);
        size_t g = 0;
        while (g < n_cmb) {
            size_t h = g;
            while (h < n_cmb && cmb[h].i == cmb[g].i && cmb[h].j == cmb[g].j) ++h;
            int32_t p1 = cmb[g].i, p2 = cmb[g].j;
            int64_t K[4][4]; memset(K, 0, sizeof(K));
            size_t q = g;
            while (q < h) {                               /* sorted(mm2combo2counts.items()) */
                int32_t m = cmb[q].mm;
                for (; q < h && cmb[q].mm == m; ++q) K[cmb[q].b1][cmb[q].b2] += 1;
                int64_t C1[4], C2[4]; int pr1, pr2;
                cum_counts(counts, nmask, M, p1, m, C1, &pr1);
                cum_counts(counts, nmask, M, p2, m, C2, &pr2);
                if (!(pr1 && pr2)) continue;              /* mm not in updateMMs */
                int64_t s1 = C1[0] + C1[1] + C1[2] + C1[3], s2 = C2[0] + C2[1] + C2[2] + C2[3];
                if (s1 + s2 < min_snp) continue;
                int A, a, B, b;
                major_minor(C1, &A, &a);
                major_minor(C2, &B, &b);
                if (C1[A] == 0 || C1[a] == 0 || C2[B] == 0 || C2[b] == 0) continue;
                int64_t cAB = K[A][B], cAb = K[A][b], caB = K[a][B], cab = K[a][b];
                int64_t total = cAB + cAb + caB + cab;
                if (!(total > min_snp)) continue;         /* strict (linkage.py:165) */
                double tot = (double)total;
                double fAB = (double)cAB / tot, fAb = (double)cAb / tot, faB = (double)caB / tot, fab = (double)cab / tot;
                double fA = fAB + fAb, fa = fab + faB, fB = fAB + faB, fb = fab + fAb;
                double linkD = fAB - fA * fB;
                double r2 = NAN, dp = NAN;
                if (!(fa == 0 || fA == 0 || fB == 0 || fb == 0)) r2 = linkD * linkD / (fA * fa * fB * fb);
                double linkd = fab - fa * fb;
                if (linkd < 0) {
                    double d1 = -fA * fB, d2 = -fa * fb;
                    dp = linkd / (d1 > d2 ? d1 : d2);
                } else if (linkD > 0) {
                    double d1 = fA * fb, d2 = fa * fB;
                    dp = linkd / (d1 < d2 ? d1 : d2);
                }
                if (n_rows < cap) {
                    orc_ld_row *r = rows + n_rows;
                    memset(r, 0, sizeof(*r));
                    r->pos_a = p1 + start; r->pos_b = p2 + start; r->mm = m;
                    r->c_AB = (int32_t)cAB; r->c_Ab = (int32_t)cAb; r->c_aB = (int32_t)caB; r->c_ab = (int32_t)cab;
                    r->allele_A = (uint8_t)A; r->allele_a = (uint8_t)a; r->allele_B = (uint8_t)B; r->allele_b = (uint8_t)b;
                    r->r2 = r2; r->d_prime = dp;
                }
                ++n_rows;
            }
            g = h;
        }
    }
    free(head); free(tail); free(touched); free(ent); free(ent_next); free(cmb);
    return n_rows <= cap ? n_rows : -n_rows;
}

/* ---------------------------------------------------------------- multi-threaded driver (CPU baseline only) */
/* Runs the three stages over chunks of `chunk_splits` consecutive splits on `n_threads` OpenMP threads and returns the
 * number of SNV / linkage rows.  Used by bench.py's cpu_baseline / --impl reference legs: the reference farms splits to
 * worker processes the same way (profile_controller.py:243-271); rows are counted, not kept. */
#ifdef _OPENMP
#include <omp.h>
#endif

static int64_t lower_bound_i32(const int32_t *a, int64_t n, int64_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

int orc_profile_mt(int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual, const int32_t *read_id,
                   const int32_t *pair_mm, int64_t n_pairs, int M, int min_qual, const uint8_t *ref, int32_t ref_start,
                   const int32_t *lut, int n_lut, int lut_default, int min_cov, double min_freq, int n_splits,
                   const int32_t *splits, int min_snp, int chunk_splits, int n_threads, int64_t *n_snv, int64_t *n_ld)
{
    int64_t tot_snv = 0, tot_ld = 0;
    int failed = 0;
    const int n_chunks = (n_splits + chunk_splits - 1) / chunk_splits;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_snv, tot_ld) reduction(| : failed)
    for (int c = 0; c < n_chunks; ++c) {
        const int s0 = c * chunk_splits, s1 = (s0 + chunk_splits < n_splits) ? s0 + chunk_splits : n_splits;
        const int32_t lo = splits[2 * s0], hi = splits[2 * (s1 - 1) + 1];
        const int32_t L = hi - lo + 1;
        const int64_t e0 = lower_bound_i32(ref_pos, n, lo), e1 = lower_bound_i32(ref_pos, n, (int64_t)hi + 1);
        int32_t *counts = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * M * 4);
        uint64_t *nmask = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)L);
        int32_t *covT = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * M);
        float *clonT = (float *)malloc(sizeof(float) * (size_t)L * M);
        uint8_t *flags = (uint8_t *)malloc((size_t)L);
        int64_t cap_s = (int64_t)L * 2 + 1024, cap_l = (int64_t)L * 8 + 65536;
        orc_snv_row *srows = (orc_snv_row *)malloc(sizeof(orc_snv_row) * (size_t)cap_s);
        orc_ld_row *lrows = (orc_ld_row *)malloc(sizeof(orc_ld_row) * (size_t)cap_l);
        if (!counts || !nmask || !covT || !clonT || !flags || !srows || !lrows) failed |= 1;
        else {
            if (orc_pileup_counts(e1 - e0, ref_pos + e0, base + e0, qual + e0, read_id + e0, pair_mm, lo, L, M, min_qual,
                                  counts, nmask)) failed |= 2;
            int64_t a = orc_call_snvs(L, M, counts, nmask, ref + (lo - ref_start), lut, n_lut, lut_default, min_cov, min_freq,
                                      lo, covT, clonT, flags, srows, cap_s);
            int64_t b = orc_linkage(e1 - e0, ref_pos + e0, base + e0, qual + e0, read_id + e0, pair_mm, n_pairs, lo, L, M,
                                    min_qual, counts, nmask, flags, s1 - s0, splits + 2 * s0, min_snp, lrows, cap_l);
            tot_snv += a < 0 ? -a : a;
            tot_ld += b < 0 ? -b : b;
        }
        free(counts); free(nmask); free(covT); free(clonT); free(flags); free(srows); free(lrows);
    }
    *n_snv = tot_snv;
    *n_ld = tot_ld;
    return failed;
}
