// isb_k2_snv.cu -- K2: per-site null-model SNV caller on sm_100a.
//
// Replaces, per position, update_covT (inStrain/profile/profile_utilities.py:288-295), update_snp_table,
// call_snv_site, calc_snp_class, calculate_clonality (inStrain/profile/snv_utilities.py:40-231),
// mm_counts_to_counts (profile_utilities.py:297-312) and is_present (inStrain/readComparer.py:307-316).
//
// One thread per position runs the reference's short, stateful ascending-mm loop (anySNP / bases / cryptic carry
// state across mm levels).  HBM traffic: 16*M+1 B read, 8*M+1 B written per position, 32 B per SNV row.
// All floating-point tests are IEEE double with explicit round-to-nearest intrinsics (no FMA contraction) so the
// >= min_freq decisions and the float32 clonality are bit-identical to CPython's arithmetic.
#include "isb_common.cuh"
#include <math_constants.h>

#define K2_THREADS 256

// thr2[T] = max(null-model threshold, smallest c with (double)c / (double)T >= min_freq): with it the reference's per-base
// test "c >= model[T] and float(c)/T >= min_freq" (snv_utilities.py:173-180) is ONE integer compare, bit-exact because
// IEEE division is monotone in c.  Built once per (context, min_freq); coverages >= n_lut keep the division.
__global__ void k2_build_thr2(const int32_t *__restrict__ lut, int n_lut, int lut_default, double min_freq,
                              int32_t *__restrict__ thr2)
{
    const int T = blockIdx.x * blockDim.x + threadIdx.x;
    if (T >= n_lut) return;
    const int thr = __ldg(lut + T) >= 0 ? __ldg(lut + T) : lut_default;
    int lo = 0, hi = T + 1;                                  // smallest c in [0, T+1] passing the frequency test
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const bool ok = T > 0 && __ddiv_rn((double)mid, (double)T) >= min_freq;
        if (ok) hi = mid; else lo = mid + 1;
    }
    thr2[T] = max(thr, lo);
}

struct k2_site_state {
    int n_rows;
    int cryptic;
    unsigned bases;
    int any_snp;
};

__device__ __forceinline__ int k2_argmax4(const int *c)
{
    int b = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i) if (c[i] > c[b]) b = i;   // np.argmax: first maximum
    return b;
}

// kWrite = false: compute covT / clonT / flags and count the rows of this site.
// kWrite = true : emit the rows (cryptic already known) to rows[slot...].
template <bool kWrite>
__device__ __forceinline__ k2_site_state
k2_site_loop(int32_t p, int M, const int32_t *__restrict__ counts, unsigned long long nm, int ref,
             const int32_t *__restrict__ thr2, int n_lut, int lut_default, int32_t start, int min_cov, double min_freq,
             int32_t *__restrict__ covT, float *__restrict__ clonT, isb_snv_row *__restrict__ rows, int64_t slot,
             int64_t cap, int cryptic_final)
{
    k2_site_state st = {0, 0, 0u, 0};
    int C[4] = {0, 0, 0, 0};
    for (int m = 0; m < M; ++m) {
        const int4 E = __ldg(reinterpret_cast<const int4 *>(counts) + (size_t)p * M + m);
        const int e_sum = E.x + E.y + E.z + E.w;
        const bool present = e_sum > 0 || ((nm >> m) & 1ull);       // mm is a key of MMcounts
        if (!kWrite) {
            covT[(size_t)p * M + m] = e_sum;                          // update_covT: exact-mm coverage
        }
        float clon = CUDART_NAN_F;
        if (present) {
            C[0] += E.x; C[1] += E.y; C[2] += E.z; C[3] += E.w;       // mm_counts_to_counts(MMcounts, mm)
        }
        const int T = C[0] + C[1] + C[2] + C[3];
        const bool counted = present && T >= min_cov;
        if (counted && !kWrite) {                                     // calculate_clonality, double, A,C,T,G order
            const double s = (double)T;
            // 0/s == +0.0 exactly, so absent bases skip the (expensive) IEEE double division
            const double f0 = C[0] ? __ddiv_rn((double)C[0], s) : 0.0, f1 = C[1] ? __ddiv_rn((double)C[1], s) : 0.0;
            const double f2 = C[2] ? __ddiv_rn((double)C[2], s) : 0.0, f3 = C[3] ? __ddiv_rn((double)C[3], s) : 0.0;
            double prob = __dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1));
            prob = __dadd_rn(prob, __dmul_rn(f2, f2));
            prob = __dadd_rn(prob, __dmul_rn(f3, f3));
            clon = __double2float_rn(prob);
        }
        if (!kWrite) clonT[(size_t)p * M + m] = clon;
        if (!counted) continue;                                       // call_snv_site -> (None, 0)
        int thr, i = 0;
        if (T < n_lut) {                                              // integer form of the two-part presence test
            thr = __ldg(thr2 + T);
#pragma unroll
            for (int b = 0; b < 4; ++b) i += (C[b] >= thr);
        } else {
            thr = lut_default;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (C[b] >= thr && __ddiv_rn((double)C[b], (double)T) >= min_freq) ++i;
        }
        const int con = k2_argmax4(C);
        const bool is_row = (i > 1) || (i == 1 && con != ref) || (i == 0);
        if (!is_row) {                                                // snp == -1
            if (st.any_snp) st.cryptic = 1;
            continue;
        }
        int tmp[4] = {C[0], C[1], C[2], C[3]};
        tmp[con] = 0;
        const int var = k2_argmax4(tmp);
        if (kWrite) {
            int cls;
            if (ref > 3) cls = ISB_CLS_AMBIGUOUS_REFERENCE;
            else if (i == 0) cls = ISB_CLS_DIVERGENT_SITE;
            else if (i == 1) cls = ISB_CLS_SNS;
            else if (ref == con) cls = ISB_CLS_SNV;
            else if (ref == var) cls = ISB_CLS_CON_SNV;
            else {
                const int cr = C[ref];                                // is_present(counts[ref], total, model, min_freq)
                const bool pres = T < n_lut ? (cr >= thr) : (cr >= thr && __ddiv_rn((double)cr, (double)T) >= min_freq);
                cls = pres ? ISB_CLS_CON_SNV : ISB_CLS_POP_SNV;
            }
            const int64_t r = slot + st.n_rows;
            if (r < cap) {
                int4 lo, hi;
                lo.x = p + start; lo.y = C[0]; lo.z = C[1]; lo.w = C[2];
                hi.x = C[3]; hi.y = m;
                hi.z = (ref & 0xff) | (con << 8) | (var << 16) | (i << 24);
                hi.w = cls | (cryptic_final << 8);
                int4 *dst = reinterpret_cast<int4 *>(rows + r);
                dst[0] = lo; dst[1] = hi;
            }
        }
        st.n_rows++;
        if (i >= 2) { st.any_snp = 1; st.bases |= (1u << con) | (1u << var); }
        else if (i == 1 && st.any_snp) st.cryptic = 1;
    }
    return st;
}

__global__ void __launch_bounds__(K2_THREADS)
k2_call_snvs(int32_t L, int M, const int32_t *__restrict__ counts, const unsigned long long *__restrict__ nmask,
             const uint8_t *__restrict__ ref, const int32_t *__restrict__ thr2, int n_lut, int lut_default,
             int32_t start, int min_cov, double min_freq, int32_t *__restrict__ covT, float *__restrict__ clonT,
             uint8_t *__restrict__ site_flags, isb_snv_row *__restrict__ rows, int64_t cap,
             unsigned long long *__restrict__ n_rows)
{
    const int32_t p = blockIdx.x * K2_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = p < L;
    k2_site_state st = {0, 0, 0u, 0};
    unsigned long long nm = 0;
    int r = 4;
    if (active) {
        nm = nmask ? nmask[p] : 0ull;
        r = ref[p];
        st = k2_site_loop<false>(p, M, counts, nm, r, thr2, n_lut, lut_default, start, min_cov, min_freq, covT, clonT,
                                 nullptr, 0, 0, 0);
        site_flags[p] = (uint8_t)(st.bases | (st.any_snp ? ISB_SITE_ANYSNP : 0));
    }
    // warp-aggregated row allocation: one global atomic per warp that has rows
    int incl = st.n_rows;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(ISB_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(ISB_FULL, incl, 31);
    if (total == 0) return;
    unsigned long long base_slot = 0;
    if (lane == 0) base_slot = atomicAdd(n_rows, (unsigned long long)total);
    base_slot = __shfl_sync(ISB_FULL, base_slot, 0);
    if (st.n_rows > 0)
        k2_site_loop<true>(p, M, counts, nm, r, thr2, n_lut, lut_default, start, min_cov, min_freq, covT, clonT, rows,
                           (int64_t)base_slot + incl - st.n_rows, cap, st.cryptic);
}

int isb_k2_launch(isb_ctx *ctx, int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                  const uint8_t *ref, int32_t start, int min_cov, double min_freq, int32_t *covT, float *clonT,
                  uint8_t *site_flags, isb_snv_row *rows, int64_t cap)
{
    cudaStream_t st = ctx->stream;
    if (!ctx->keep_counters) ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 0, 0, sizeof(unsigned long long), st));
    if (L <= 0) return ISB_OK;
    if (!ctx->d_thr2 || ctx->thr2_min_freq != min_freq) {             // (re)build the merged integer threshold table
        if (!ctx->d_thr2) ISB_CUDA(cudaMalloc(&ctx->d_thr2, sizeof(int32_t) * (size_t)ctx->n_lut));
        k2_build_thr2<<<(ctx->n_lut + 255) / 256, 256, 0, st>>>(ctx->d_lut, ctx->n_lut, ctx->lut_default, min_freq, ctx->d_thr2);
        ISB_LAUNCH_CHECK();
        ctx->thr2_min_freq = min_freq;
    }
    k2_call_snvs<<<(L + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, st>>>(
        L, M, counts, nmask, ref, ctx->d_thr2, ctx->n_lut, ctx->lut_default, start, min_cov, min_freq, covT, clonT,
        site_flags, rows, cap, ctx->d_counters + 0);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
