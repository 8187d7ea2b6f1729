// isb_k2_site.cuh -- the per-site arithmetic of the SNV caller shared by K2 (isb_k2_snv.cu) and the fused pileup + SNV
// kernel of the column-word path (isb_k1c_cols.cu): exact quotients, np.argmax, and the single-level (M = 1) site call.
// Reference: call_snv_site / update_snp_table / calc_snp_class / calculate_clonality (inStrain/profile/snv_utilities.py:40-231),
// is_present (inStrain/readComparer.py:307-316).
#pragma once
#include "isb_common.cuh"
#include <math_constants.h>

// Correctly rounded c / s for the four base frequencies of one site from ONE correctly rounded reciprocal
// (Markstein: with y = RN(1/s) and q = RN(c*y), q' = RN(q + (c - s*q)*y) is RN(c/s)).  Three FP64 operations per
// quotient instead of a full IEEE division each.  Used for s <= K2_FAST_DIV_MAX, the range for which
// isb_selftest_division (tests/test_gpu_parity.py) compares it bit for bit with __ddiv_rn for EVERY 0 <= c <= s.
#define K2_FAST_DIV_MAX 65536
__device__ __forceinline__ double k2_quot(double c, double s, double rcp)
{
    const double q = __dmul_rn(c, rcp);
    const double e = __fma_rn(-q, s, c);
    return __fma_rn(e, rcp, q);
}

// ---- counter-based random numbers for the two resampled outputs (clonTR, normalized linkage) -----------------------------
// The reference draws them with an unseeded np.random.choice (snv_utilities.py:241, linkage.py:200), so its own tests drop
// those columns before comparing.  Here every random word is a pure function of (seed, stream tag, site key, word index):
// word k of a site = mix64(mix64(seed + tag) ^ a * K1 ^ b * K2 ^ k * K3) (splitmix64's finaliser).  Reproducible, identical in
// every input layout, restated in numpy by the test oracle for bit-exact parity tests; same distribution as the reference's.
//
// n draws with replacement from 4 categories with counts c[] (total T) = a multinomial, sampled BIT-SLICED: the n trials are
// the bit lanes of one 64-bit word.  Categories are split off one after the other (category i against the rest: a
// binomial with p = c_i / remaining total); a trial compares a lazily generated uniform with p from the most significant
// bit down -- one random word decides that bit for all trials at once, and a trial is decided as soon as its bit differs
// from p's (success iff p has the 1).  Half of the undecided trials are decided per word, so ~log2(n) + 1.3 words serve all
// n trials of a stage (7 for n = 50) instead of one word per two draws; p is taken to 32 bits (the last non-empty category
// takes what remains: exact totals).
#define ISB_RNG_TAG_CLONR 0x636c6f6e54520001ull
#define ISB_RNG_TAG_LD 0x6c646e6f726d0002ull
#define ISB_RNG_K3 0x8cb92ba72f3d8dd7ull
#define ISB_RNG_K4 0xa0761d6478bd642full
__host__ __device__ __forceinline__ uint64_t isb_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// word k of the site keyed (a, b) = isb_mix64(isb_rng_base(seed, tag, a, b) ^ k * K3); the base is computed once per site
__host__ __device__ __forceinline__ uint64_t isb_rng_base(uint64_t seed, uint64_t tag, uint64_t a, uint64_t b)
{
    return isb_mix64(seed + tag) ^ (a * 0x9e3779b97f4a7c15ull) ^ (b * 0xd1b54a32d192ed03ull);
}

// r[i] = number of the n draws that fall into category i (sum = n).  rb = isb_rng_base of the site.
__device__ __forceinline__ void isb_redraw4(const int (&c)[4], int T, int n, uint64_t rb, int (&r)[4])
{
    r[0] = r[1] = r[2] = r[3] = 0;
    for (int q = 0; q * 64 < n; ++q) {                                // 64 trials per chunk (n = 50, 20: one chunk)
        const int nt = min(64, n - q * 64);
        uint64_t left = nt == 64 ? ~0ull : ((1ull << nt) - 1ull);     // trials not yet assigned to a category
        const uint64_t base = rb ^ ((uint64_t)q * ISB_RNG_K4);
        uint64_t kk = 0;                                              // word index * K3
        int rem = T;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (c[i] == 0 || left == 0ull) continue;
            if (c[i] == rem) { r[i] += __popcll(left); left = 0ull; continue; }     // the last non-empty category: all that remain
            const uint32_t P = (uint32_t)(unsigned long long)__dmul_rn(__ddiv_rn((double)c[i], (double)rem), 4294967296.0);
            uint64_t und = left, hit = 0ull;
            for (int b = 31; b >= 0 && und != 0ull; --b) {
                const uint64_t W = isb_mix64(base ^ kk);
                kk += ISB_RNG_K3;
                if ((P >> b) & 1u) { hit |= ~W & und; und &= W; }    // p's bit is 1: a 0 bit of the uniform is below p
                else und &= ~W;                                       // p's bit is 0: a 1 bit of the uniform is above p
            }
            r[i] += __popcll(hit);
            left &= ~hit;
            rem -= c[i];
        }
    }
}

// calculate_rarefied_clonality (snv_utilities.py:233-247): n bases drawn with replacement from the frequencies C / T, then
// the clonality of the drawn counts (double, A,C,T,G order, as calculate_clonality).  pos = batch coordinate.
__device__ __forceinline__ float k2_rarefied_clon(const int (&C)[4], int T, int n, uint64_t seed, int64_t pos, int mm)
{
    const int mx = max(max(C[0], C[1]), max(C[2], C[3]));
    if (mx == T) return 1.0f;                                         // every draw lands on the one base: (n/n)^2 + 0 + 0 + 0
    int r[4];
    isb_redraw4(C, T, n, isb_rng_base(seed, ISB_RNG_TAG_CLONR, (uint64_t)pos, (uint64_t)mm), r);
    const double s = (double)n;
    const double f0 = __ddiv_rn((double)r[0], s), f1 = __ddiv_rn((double)r[1], s);
    const double f2 = __ddiv_rn((double)r[2], s), f3 = __ddiv_rn((double)r[3], s);
    double prob = __dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1));
    prob = __dadd_rn(prob, __dmul_rn(f2, f2));
    prob = __dadd_rn(prob, __dmul_rn(f3, f3));
    return __double2float_rn(prob);
}

__device__ __forceinline__ int k2_argmax4(const int *c)
{
    int b = 0, v = c[0];                                  // np.argmax: first maximum (value tracked: no dynamic indexing)
#pragma unroll
    for (int i = 1; i < 4; ++i) if (c[i] > v) { v = c[i]; b = i; }
    return b;
}

// One site at M = 1 (--skip_mm_profiling): the reference's loop body runs once, `cryptic` can never be set and the site
// has at most one row.  C = A,C,T,G counts, r = reference base code, nm0 = "level 0 is a key of MMcounts although no
// A/C/T/G base was counted" (a passing non-ACGT read base, profile_utilities.py:280-281).
struct k2_m1_site {
    int T;            // coverage (covT)
    float clon;       // clonality (NaN = unset)
    unsigned flags;   // site_flags byte
    bool is_row;      // the site emits a raw_snp_table row
    int i, con, thr;  // allele count ("morphia"), consensus base, threshold used (for the row / class)
};

// thr_T = thr2[T] (the merged integer threshold of coverage T = sum of C) when T < n_lut; unused otherwise.
__device__ __forceinline__ k2_m1_site k2_site_m1(const int (&C)[4], int r, bool nm0, int thr_T,
                                                 int n_lut, int lut_default, int min_cov, double min_freq)
{
    k2_m1_site s;
    s.T = C[0] + C[1] + C[2] + C[3];
    s.clon = CUDART_NAN_F;
    s.flags = 0u;
    s.is_row = false;
    s.i = 0; s.con = 0; s.thr = 0;
    const int T = s.T;
    const bool present = T > 0 || nm0;
    const bool counted = present && T >= min_cov;
    if (!counted) return s;                                           // call_snv_site -> (None, 0)
    const int mx = max(max(C[0], C[1]), max(C[2], C[3]));
    if (mx == T && T > 0) {
        s.clon = 1.0f;                                                // one base only: (T/T)^2 + 0 + 0 + 0 is exactly 1
    } else {                                                          // calculate_clonality, double, A,C,T,G order
        const double sd = (double)T;
        double f0, f1, f2, f3;
        if (T <= K2_FAST_DIV_MAX) {
            const double rcp = __drcp_rn(sd);
            f0 = k2_quot((double)C[0], sd, rcp); f1 = k2_quot((double)C[1], sd, rcp);
            f2 = k2_quot((double)C[2], sd, rcp); f3 = k2_quot((double)C[3], sd, rcp);
        } else {
            f0 = C[0] ? __ddiv_rn((double)C[0], sd) : 0.0; f1 = C[1] ? __ddiv_rn((double)C[1], sd) : 0.0;
            f2 = C[2] ? __ddiv_rn((double)C[2], sd) : 0.0; f3 = C[3] ? __ddiv_rn((double)C[3], sd) : 0.0;
        }
        double prob = __dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1));
        prob = __dadd_rn(prob, __dmul_rn(f2, f2));
        prob = __dadd_rn(prob, __dmul_rn(f3, f3));
        s.clon = __double2float_rn(prob);
    }
    int i = 0;
    if (T < n_lut) {                                                  // integer form of the two-part presence test
        s.thr = thr_T;
#pragma unroll
        for (int b = 0; b < 4; ++b) i += (C[b] >= s.thr);
    } else {
        s.thr = lut_default;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (C[b] >= s.thr && __ddiv_rn((double)C[b], (double)T) >= min_freq) ++i;
    }
    s.i = i;
    s.con = k2_argmax4(C);
    s.is_row = (i > 1) || (i == 1 && s.con != r) || (i == 0);
    if (s.is_row && i >= 2) {
        const int tmp[4] = {s.con == 0 ? 0 : C[0], s.con == 1 ? 0 : C[1], s.con == 2 ? 0 : C[2], s.con == 3 ? 0 : C[3]};
        s.flags = ISB_SITE_ANYSNP | (1u << s.con) | (1u << k2_argmax4(tmp));
    }
    return s;
}

// the raw_snp_table row of a site for which k2_site_m1 said is_row (pos = batch coordinate)
__device__ __forceinline__ void k2_write_row_m1(isb_snv_row *__restrict__ dst_row, int32_t pos, const int (&C)[4], int r,
                                                const k2_m1_site &s, int n_lut, double min_freq)
{
    const int con = s.con, i = s.i, T = s.T, thr = s.thr;
    const int tmp[4] = {con == 0 ? 0 : C[0], con == 1 ? 0 : C[1], con == 2 ? 0 : C[2], con == 3 ? 0 : C[3]};
    const int var = k2_argmax4(tmp);
    int cls;
    if (r > 3) cls = ISB_CLS_AMBIGUOUS_REFERENCE;
    else if (i == 0) cls = ISB_CLS_DIVERGENT_SITE;
    else if (i == 1) cls = ISB_CLS_SNS;
    else if (r == con) cls = ISB_CLS_SNV;
    else if (r == var) cls = ISB_CLS_CON_SNV;
    else {
        const int cr = r == 0 ? C[0] : r == 1 ? C[1] : r == 2 ? C[2] : C[3];                   // is_present(counts[ref], ...)
        const bool pres = T < n_lut ? (cr >= thr) : (cr >= thr && __ddiv_rn((double)cr, (double)T) >= min_freq);
        cls = pres ? ISB_CLS_CON_SNV : ISB_CLS_POP_SNV;
    }
    int4 lo, hi;
    lo.x = pos; lo.y = C[0]; lo.z = C[1]; lo.w = C[2];
    hi.x = C[3]; hi.y = 0;
    hi.z = (r & 0xff) | (con << 8) | (var << 16) | (i << 24);
    hi.w = cls;
    int4 *dst = reinterpret_cast<int4 *>(dst_row);
    dst[0] = lo; dst[1] = hi;
}
