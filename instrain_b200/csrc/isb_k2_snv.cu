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
#include "isb_k2_site.cuh"
#include <math_constants.h>
#include <cstdlib>

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

__global__ void k2_selftest_division(int s_lo, int s_hi, unsigned long long *__restrict__ mismatches)
{
    const int s = s_lo + blockIdx.x;
    if (s > s_hi) return;
    const double ds = (double)s, rcp = __drcp_rn(ds);
    unsigned long long bad = 0;
    for (int c = threadIdx.x; c <= s; c += blockDim.x)
        bad += __double_as_longlong(k2_quot((double)c, ds, rcp)) != __double_as_longlong(__ddiv_rn((double)c, ds));
    if (bad) atomicAdd(mismatches, bad);
}

struct k2_site_state {
    int n_rows;
    int cryptic;
    unsigned bases;
    int any_snp;
    int C[4];                                                         // cumulative A,C,T,G counts up to the current level
};

// Rarefied clonality at M > 1.  A (position, level) needs ~750 instructions of draws only when the level reaches the rarefied
// coverage AND holds more than one base (~10 % of the cells), but a warp whose 32 lanes walk their levels in lock step
// would run that path at almost every level for a few lanes each.  The cells that need draws are therefore queued per warp
// in shared memory and processed DENSELY, one cell per lane, whenever 32 are waiting.
struct k2_rq_entry { int C[4]; int32_t p; int32_t m; };
struct k2_rq { k2_rq_entry *e; int n; };                               // e: [64] per warp; n: warp-uniform

__device__ __forceinline__ void k2_rq_drain(k2_rq &q, bool all, float *clonTR, int M, int cov_r, uint64_t seed,
                                            int32_t start, int cstride = 1)
{
    const int lane = threadIdx.x & 31;
    while (q.n >= 32 || (all && q.n > 0)) {
        const int take = min(q.n, 32);
        __syncwarp();
        if (lane < take) {
            const k2_rq_entry e = q.e[lane];
            const int C[4] = {e.C[0], e.C[1], e.C[2], e.C[3]};
            clonTR[((int64_t)e.p * M + e.m) * cstride] = k2_rarefied_clon(C, C[0] + C[1] + C[2] + C[3], cov_r, seed, (int64_t)e.p + start, e.m);
        }
        const bool has = lane + 32 < q.n;
        k2_rq_entry mv;
        if (has) mv = q.e[lane + 32];
        __syncwarp();
        if (has) q.e[lane] = mv;
        q.n -= take;
    }
    __syncwarp();
}

// The reference's ascending-mm loop over levels [m0, m0 + mc) of ONE site; `st` carries anySNP / bases / cryptic / the
// cumulative counts across calls.  crow[j], cov_row[j], clon_row[j] address level m0 + j (global rows, or the block's
// shared-memory tiles in the staged kernel).
// kWrite = false: compute covT / clonT / flags and count the rows of this site.
// kWrite = true : emit the rows (cryptic already known) to rows[slot...].
template <bool kWrite>
__device__ __forceinline__ void
k2_site_levels(k2_site_state &st, int32_t p, int m0, int mc, const int4 *crow, unsigned long long nm, int ref,
               const int32_t *__restrict__ thr2, int n_lut, int lut_default, int32_t start, int min_cov, double min_freq,
               int32_t *cov_row, float *clon_row, isb_snv_row *__restrict__ rows, int64_t slot, int64_t cap,
               int cryptic_final, float *clonr_row = nullptr, int cov_r = 0, uint64_t seed = 0, k2_rq *rq = nullptr,
               bool valid = true, float *clonTR = nullptr, int M = 0, int cstride = 1)
{
    // rq != nullptr: EVERY lane of the warp runs this loop (valid = false for lanes without a position): the queue uses
    // full-warp ballots
    int *C = st.C;
    for (int j = 0; j < mc; ++j) {
        const int m = m0 + j;
        const int4 E = valid ? crow[j] : make_int4(0, 0, 0, 0);
        const int e_sum = E.x + E.y + E.z + E.w;
        const bool present = e_sum > 0 || ((nm >> m) & 1ull);       // mm is a key of MMcounts
        if (!kWrite) cov_row[j] = e_sum;                              // update_covT: exact-mm coverage
        float clon = CUDART_NAN_F;
        if (present) {
            C[0] += E.x; C[1] += E.y; C[2] += E.z; C[3] += E.w;       // mm_counts_to_counts(MMcounts, mm)
        }
        const int T = C[0] + C[1] + C[2] + C[3];
        const bool counted = present && T >= min_cov;
        if (counted && !kWrite && max(max(C[0], C[1]), max(C[2], C[3])) == T && T > 0) {
            clon = 1.0f;                                              // one base only: (T/T)^2 + 0 + 0 + 0 is exactly 1 (most cells)
        } else if (counted && !kWrite) {                              // calculate_clonality, double, A,C,T,G order
            const double s = (double)T;
            double f0, f1, f2, f3;
            if (T <= K2_FAST_DIV_MAX) {                               // one reciprocal, four 3-op quotients (bit-exact)
                const double rcp = __drcp_rn(s);
                f0 = k2_quot((double)C[0], s, rcp); f1 = k2_quot((double)C[1], s, rcp);
                f2 = k2_quot((double)C[2], s, rcp); f3 = k2_quot((double)C[3], s, rcp);
            } else {                                                  // 0/s == +0.0 exactly: absent bases skip the division
                f0 = C[0] ? __ddiv_rn((double)C[0], s) : 0.0; f1 = C[1] ? __ddiv_rn((double)C[1], s) : 0.0;
                f2 = C[2] ? __ddiv_rn((double)C[2], s) : 0.0; f3 = C[3] ? __ddiv_rn((double)C[3], s) : 0.0;
            }
            double prob = __dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1));
            prob = __dadd_rn(prob, __dmul_rn(f2, f2));
            prob = __dadd_rn(prob, __dmul_rn(f3, f3));
            clon = __double2float_rn(prob);
        }
        if (!kWrite) clon_row[j] = clon;
        if (!kWrite && (clonr_row || rq)) {                           // clonTR[mm][pos]: set where the coverage reaches cov_r
            bool need = false;
            if (valid && clonr_row) {
                if (present && cov_r > 0 && T >= cov_r) {
                    if (max(max(C[0], C[1]), max(C[2], C[3])) == T) clonr_row[m * cstride] = 1.0f;      // one base only: no draws
                    else if (rq) need = true;
                    else clonr_row[m * cstride] = k2_rarefied_clon(st.C, T, cov_r, seed, (int64_t)p + start, m);
                } else {
                    clonr_row[m * cstride] = CUDART_NAN_F;
                }
            }
            if (rq) {
                const unsigned nm_ = __ballot_sync(ISB_FULL, need);
                if (need) {
                    k2_rq_entry e;
                    e.C[0] = C[0]; e.C[1] = C[1]; e.C[2] = C[2]; e.C[3] = C[3]; e.p = p; e.m = m;
                    rq->e[rq->n + __popc(nm_ & ((1u << (threadIdx.x & 31)) - 1u))] = e;
                }
                rq->n += __popc(nm_);
                if (rq->n >= 32) k2_rq_drain(*rq, false, clonTR, M, cov_r, seed, start, cstride);
            }
        }
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
        const int tmp[4] = {con == 0 ? 0 : C[0], con == 1 ? 0 : C[1], con == 2 ? 0 : C[2], con == 3 ? 0 : C[3]};
        const int var = k2_argmax4(tmp);                              // selects, not tmp[con] = 0: keeps C[] in registers
        if (kWrite) {
            int cls;
            if (ref > 3) cls = ISB_CLS_AMBIGUOUS_REFERENCE;
            else if (i == 0) cls = ISB_CLS_DIVERGENT_SITE;
            else if (i == 1) cls = ISB_CLS_SNS;
            else if (ref == con) cls = ISB_CLS_SNV;
            else if (ref == var) cls = ISB_CLS_CON_SNV;
            else {
                const int cr = ref == 0 ? C[0] : ref == 1 ? C[1] : ref == 2 ? C[2] : C[3];   // is_present(counts[ref], ...)
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
}

// Warp-aggregated row allocation (one global atomic per warp that has rows) + the second, row-writing pass of the sites
// that have rows; that pass re-reads the site's counts from global memory (L2-resident, < 1 % of the sites).
__device__ __forceinline__ void
k2_emit_rows(const k2_site_state &st, int32_t p, int M, const int32_t *__restrict__ counts, unsigned long long nm, int ref,
             const int32_t *__restrict__ thr2, int n_lut, int lut_default, int32_t start, int min_cov, double min_freq,
             isb_snv_row *__restrict__ rows, int64_t cap, unsigned long long *__restrict__ n_rows)
{
    const int lane = threadIdx.x & 31;
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
    if (st.n_rows > 0) {
        k2_site_state w = {0, 0, 0u, 0, {0, 0, 0, 0}};
        k2_site_levels<true>(w, p, 0, M, reinterpret_cast<const int4 *>(counts) + (size_t)p * M, nm, ref, thr2, n_lut,
                             lut_default, start, min_cov, min_freq, nullptr, nullptr, rows,
                             (int64_t)base_slot + incl - st.n_rows, cap, st.cryptic);
    }
}

// M = 1 (and the fallback): one thread per position straight from global memory -- with one level a warp's 32 rows are
// one contiguous 512-byte block, already perfectly coalesced.
__global__ void __launch_bounds__(K2_THREADS)
k2_call_snvs(int32_t L, int M, const int32_t *__restrict__ counts, const unsigned long long *__restrict__ nmask,
             const uint8_t *__restrict__ ref, const int32_t *__restrict__ thr2, int n_lut, int lut_default,
             int32_t start, int min_cov, double min_freq, int32_t *__restrict__ covT, float *__restrict__ clonT,
             uint8_t *__restrict__ site_flags, isb_snv_row *__restrict__ rows, int64_t cap,
             unsigned long long *__restrict__ n_rows, float *__restrict__ clonTR, int cov_r, uint64_t seed)
{
    const int32_t p = blockIdx.x * K2_THREADS + threadIdx.x;
    const bool active = p < L;
    k2_site_state st = {0, 0, 0u, 0, {0, 0, 0, 0}};
    unsigned long long nm = 0;
    int r = 4;
    if (active) {
        nm = nmask ? nmask[p] : 0ull;
        r = ref[p];
        k2_site_levels<false>(st, p, 0, M, reinterpret_cast<const int4 *>(counts) + (size_t)p * M, nm, r, thr2, n_lut,
                              lut_default, start, min_cov, min_freq, covT + (size_t)p * M, clonT + (size_t)p * M, nullptr,
                              0, 0, 0, clonTR ? clonTR + (size_t)p * M : nullptr, cov_r, seed);
        site_flags[p] = (uint8_t)(st.bases | (st.any_snp ? ISB_SITE_ANYSNP : 0));
    }
    k2_emit_rows(st, p, M, counts, nm, r, thr2, n_lut, lut_default, start, min_cov, min_freq, rows, cap, n_rows);
}

// M = 1 specialisation (--skip_mm_profiling, the throughput configuration).  With a single level the reference's loop
// body runs once per site: no cumulative state, `cryptic` can never be set (it needs an earlier level with a SNP), and
// a site has at most ONE row -- so rows are allocated with a ballot (one atomic per warp that has rows) and written in
// the same pass.  ncu on the generic kernel at M = 1: 254 warp instructions per 32 sites, 80 % of them integer / control
// overhead of the level loop, the state struct and the second (row-writing) pass that 27 % of the warps entered.
__global__ void __launch_bounds__(K2_THREADS)
k2_call_snvs_m1(int32_t L, const int32_t *__restrict__ counts, const unsigned long long *__restrict__ nmask,
                const uint8_t *__restrict__ ref, const int32_t *__restrict__ thr2, int n_lut, int lut_default,
                int32_t start, int min_cov, double min_freq, int32_t *__restrict__ covT, float *__restrict__ clonT,
                uint8_t *__restrict__ site_flags, isb_snv_row *__restrict__ rows, int64_t cap,
                unsigned long long *__restrict__ n_rows, float *__restrict__ clonTR, int cov_r, uint64_t seed)
{
    const int32_t p = blockIdx.x * K2_THREADS + threadIdx.x;
    const bool active = p < L;
    k2_m1_site s;
    s.is_row = false;
    int C[4] = {0, 0, 0, 0}, r = 4;
    if (active) {
        const int4 E = __ldg(reinterpret_cast<const int4 *>(counts) + p);
        r = ref[p];
        C[0] = E.x; C[1] = E.y; C[2] = E.z; C[3] = E.w;
        const bool nm0 = (E.x + E.y + E.z + E.w) == 0 && nmask && (nmask[p] & 1ull);   // level 0 is a key of MMcounts
        const int T = E.x + E.y + E.z + E.w;
        const int thr_T = (T >= min_cov && T < n_lut) ? __ldg(thr2 + T) : lut_default;
        s = k2_site_m1(C, r, nm0, thr_T, n_lut, lut_default, min_cov, min_freq);
        covT[p] = s.T;
        clonT[p] = s.clon;
        site_flags[p] = (uint8_t)s.flags;
        if (clonTR) clonTR[p] = (cov_r > 0 && T >= cov_r) ? k2_rarefied_clon(C, T, cov_r, seed, (int64_t)p + start, 0) : CUDART_NAN_F;
    }
    const unsigned mask = __ballot_sync(ISB_FULL, s.is_row);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base_slot = 0;
    if (lane == 0) base_slot = atomicAdd(n_rows, (unsigned long long)__popc(mask));
    base_slot = __shfl_sync(ISB_FULL, base_slot, 0);
    if (!s.is_row) return;
    const int64_t slot = (int64_t)base_slot + __popc(mask & ((1u << lane) - 1u));
    if (slot >= cap) return;
    k2_write_row_m1(rows + slot, p + start, C, r, s, n_lut, min_freq);
}

// M > 1: a thread's M count quads are 16*M bytes apart from its neighbour's, so direct loads touch one 32-byte sector per
// 16 useful bytes and the 4-byte covT / clonT stores one sector each.  The staged kernel moves the block's rows through
// shared memory instead.  M <= K2S_MC (every realistic read filter: <= 15 mismatches per pair): the block's input rows
// are ONE contiguous run, fetched with a single TMA 1-D bulk copy (cp.async.bulk + mbarrier complete_tx); the level loop
// runs out of shared memory into two shared output tiles with the global layout, which leave as 128-bit coalesced stores.
// M > K2S_MC: chunks of K2S_MC levels, per-row bulk copies into padded rows, the site state carried in registers.
#define K2S_THREADS 128
#define K2S_MC 16
__global__ void __launch_bounds__(K2S_THREADS)
k2_call_snvs_staged(int32_t L, int M, const int32_t *__restrict__ counts, const unsigned long long *__restrict__ nmask,
                    const uint8_t *__restrict__ ref, const int32_t *__restrict__ thr2, int n_lut, int lut_default,
                    int32_t start, int min_cov, double min_freq, int32_t *__restrict__ covT, float *__restrict__ clonT,
                    uint8_t *__restrict__ site_flags, isb_snv_row *__restrict__ rows, int64_t cap,
                    unsigned long long *__restrict__ n_rows, float *__restrict__ clonTR, int cov_r, uint64_t seed)
{
    extern __shared__ __align__(128) unsigned char k2_smem[];
    const bool whole = M <= K2S_MC;                                   // one chunk: rows keep the global (unpadded) layout
    const int MC = min(M, K2S_MC);
    const int sin = whole ? M : (MC | 1);                             // input row stride in int4 units
    const int sout = whole ? M : (MC | 1);                            // output row stride in words
    int4 *s_in = reinterpret_cast<int4 *>(k2_smem);
    int32_t *s_cov = reinterpret_cast<int32_t *>(s_in + (size_t)K2S_THREADS * sin);
    float *s_clon = reinterpret_cast<float *>(s_cov + (size_t)K2S_THREADS * sout);
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_clon + (size_t)K2S_THREADS * sout);   // 8-byte aligned: K2S_THREADS is even
    k2_rq rq = {reinterpret_cast<k2_rq_entry *>(bar + 2) + (threadIdx.x >> 5) * 64, 0};   // per-warp queue of cells that need draws
    k2_rq *rqp = (clonTR && cov_r > 0) ? &rq : nullptr;
    const int t = threadIdx.x;
    const int32_t p0 = blockIdx.x * K2S_THREADS;
    const int32_t p = p0 + t;
    const bool active = p < L;
    const int npos = min(K2S_THREADS, L - p0);
    if (t == 0) {
        isb_mbar_init(bar, whole ? 1 : K2S_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    k2_site_state st = {0, 0, 0u, 0, {0, 0, 0, 0}};
    unsigned long long nm = 0;
    int r = 4;
    if (whole) {
        if (t == 0) {
            const unsigned bytes = (unsigned)npos * (unsigned)M * 16u;
            isb_mbar_expect_tx(bar, bytes);
            isb_bulk_g2s(s_in, reinterpret_cast<const int4 *>(counts) + (size_t)p0 * M, bytes, bar);
        }
        if (active) {
            nm = nmask ? nmask[p] : 0ull;
            r = ref[p];
        }
        isb_mbar_wait(bar, 0);
        if (active || rqp)                                            // with the draw queue every lane walks the levels
            // clonTR of cell (position, level) is parked in the first word of the cell's (already consumed) input quad and
            // written out coalesced below: a row per thread at stride M straight to global memory was one 32-byte sector
            // per 4-byte store.  Cells are queued for draws only after their quad has been read.
            k2_site_levels<false>(st, p, 0, M, s_in + (size_t)t * M, nm, r, thr2, n_lut, lut_default, start, min_cov,
                                  min_freq, s_cov + (size_t)t * M, s_clon + (size_t)t * M, nullptr, 0, 0, 0,
                                  clonTR && active ? reinterpret_cast<float *>(s_in + (size_t)t * M) : nullptr, cov_r, seed, rqp,
                                  active, reinterpret_cast<float *>(s_in) - (int64_t)p0 * M * 4, M, 4);
        if (rqp) k2_rq_drain(rq, true, reinterpret_cast<float *>(s_in) - (int64_t)p0 * M * 4, M, cov_r, seed, start, 4);
        __syncthreads();
        const int n_out = npos * M;                                   // words; the block's output run starts 16-byte aligned
        const size_t g0 = (size_t)p0 * M;
        const int n4 = n_out >> 2;
        int4 *gc = reinterpret_cast<int4 *>(covT + g0);
        int4 *gl = reinterpret_cast<int4 *>(clonT + g0);
        const int4 *sc = reinterpret_cast<const int4 *>(s_cov), *sl = reinterpret_cast<const int4 *>(s_clon);
        for (int e = t; e < n4; e += K2S_THREADS) { gc[e] = sc[e]; gl[e] = sl[e]; }
        for (int e = (n4 << 2) + t; e < n_out; e += K2S_THREADS) { covT[g0 + e] = s_cov[e]; clonT[g0 + e] = s_clon[e]; }
        if (clonTR) {
            const float *sr = reinterpret_cast<const float *>(s_in);
            for (int e = t; e < n_out; e += K2S_THREADS) clonTR[g0 + e] = sr[(size_t)e * 4];
        }
    } else {
        if (active) {
            nm = nmask ? nmask[p] : 0ull;
            r = ref[p];
        }
        unsigned parity = 0;
        for (int m0 = 0; m0 < M; m0 += MC) {
            const int mc = min(MC, M - m0);
            if (active) {                                             // own row -> own padded shared row
                isb_mbar_expect_tx(bar, (unsigned)mc * 16u);
                isb_bulk_g2s(s_in + (size_t)t * sin, reinterpret_cast<const int4 *>(counts) + (size_t)p * M + m0,
                             (unsigned)mc * 16u, bar);
            } else {
                isb_mbar_arrive(bar);
            }
            isb_mbar_wait(bar, parity);
            parity ^= 1u;
            if (active || rqp)
                k2_site_levels<false>(st, p, m0, mc, s_in + (size_t)t * sin, nm, r, thr2, n_lut, lut_default, start,
                                      min_cov, min_freq, s_cov + (size_t)t * sout, s_clon + (size_t)t * sout, nullptr, 0,
                                      0, 0, clonTR && active ? clonTR + (size_t)p * M : nullptr, cov_r, seed, rqp, active, clonTR, M);
            if (rqp) k2_rq_drain(rq, true, clonTR, M, cov_r, seed, start);
            __syncthreads();
            // two rows per warp pass: 16 lanes per row (mc <= 16), no integer division
            for (int q = (t >> 4); q < npos; q += K2S_THREADS / 16) {
                const int j = t & 15;
                if (j < mc) {
                    const size_t g = (size_t)(p0 + q) * M + m0 + j;
                    covT[g] = s_cov[q * sout + j];
                    clonT[g] = s_clon[q * sout + j];
                }
            }
            if (m0 + MC < M) __syncthreads();                         // tiles are reused by the next chunk
        }
    }
    if (active) site_flags[p] = (uint8_t)(st.bases | (st.any_snp ? ISB_SITE_ANYSNP : 0));
    k2_emit_rows(st, p, M, counts, nm, r, thr2, n_lut, lut_default, start, min_cov, min_freq, rows, cap, n_rows);
}

static size_t k2s_smem_bytes(int M)
{
    const size_t sin = M <= K2S_MC ? (size_t)M : (size_t)(K2S_MC | 1), sout = sin;
    return (size_t)K2S_THREADS * sin * 16 + 2 * (size_t)K2S_THREADS * sout * 4 + 16 + (K2S_THREADS / 32) * 64 * sizeof(k2_rq_entry);
}

// Row counter reset + the merged integer threshold table of (context, min_freq); also called by the fused pileup + SNV
// kernel of the column-word path (isb_k1c_cols.cu), which runs k2_site_m1 in its epilogue.
int isb_k2_prepare(isb_ctx *ctx, double min_freq)
{
    cudaStream_t st = ctx->stream;
    if (!ctx->keep_counters) ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 0, 0, sizeof(unsigned long long), st));
    if (!ctx->d_thr2 || ctx->thr2_min_freq != min_freq) {             // (re)build the merged integer threshold table
        if (!ctx->d_thr2) ISB_CUDA(cudaMalloc(&ctx->d_thr2, sizeof(int32_t) * (size_t)ctx->n_lut));
        k2_build_thr2<<<(ctx->n_lut + 255) / 256, 256, 0, st>>>(ctx->d_lut, ctx->n_lut, ctx->lut_default, min_freq, ctx->d_thr2);
        ISB_LAUNCH_CHECK();
        ctx->thr2_min_freq = min_freq;
    }
    return ISB_OK;
}

int isb_k2_launch(isb_ctx *ctx, int32_t L, int M, const int32_t *counts, const unsigned long long *nmask,
                  const uint8_t *ref, int32_t start, int min_cov, double min_freq, int32_t *covT, float *clonT,
                  uint8_t *site_flags, isb_snv_row *rows, int64_t cap, float *clonTR, int cov_r, uint64_t seed)
{
    cudaStream_t st = ctx->stream;
    int rc = isb_k2_prepare(ctx, min_freq);
    if (rc) return rc;
    if (L <= 0) return ISB_OK;
    static const int variant = getenv("ISB_K2_VARIANT") ? atoi(getenv("ISB_K2_VARIANT")) : -1;   // 0 = generic direct kernel, otherwise M = 1: specialised, M > 1: staged
    if (M > 1 && variant != 0) {
        const size_t smem = k2s_smem_bytes(M);
        static bool attr_set[64] = {false};                          // function attributes are per device
        if (!attr_set[ctx->device & 63]) {
            ISB_CUDA(cudaFuncSetAttribute(k2_call_snvs_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2s_smem_bytes(64)));
            attr_set[ctx->device & 63] = true;
        }
        k2_call_snvs_staged<<<(L + K2S_THREADS - 1) / K2S_THREADS, K2S_THREADS, smem, st>>>(
            L, M, counts, nmask, ref, ctx->d_thr2, ctx->n_lut, ctx->lut_default, start, min_cov, min_freq, covT, clonT,
            site_flags, rows, cap, ctx->d_counters + 0, clonTR, cov_r, seed);
    } else if (M == 1 && variant != 0) {
        k2_call_snvs_m1<<<(L + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, st>>>(
            L, counts, nmask, ref, ctx->d_thr2, ctx->n_lut, ctx->lut_default, start, min_cov, min_freq, covT, clonT,
            site_flags, rows, cap, ctx->d_counters + 0, clonTR, cov_r, seed);
    } else {
        k2_call_snvs<<<(L + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, st>>>(
            L, M, counts, nmask, ref, ctx->d_thr2, ctx->n_lut, ctx->lut_default, start, min_cov, min_freq, covT, clonT,
            site_flags, rows, cap, ctx->d_counters + 0, clonTR, cov_r, seed);
    }
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

// Self-test of the fast quotient: number of (c, s) pairs, s_lo <= s <= s_hi, 0 <= c <= s, whose k2_quot differs from __ddiv_rn.
int isb_k2_selftest_division(isb_ctx *ctx, int s_lo, int s_hi, unsigned long long *h_mismatches)
{
    cudaStream_t st = ctx->stream;
    ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 5, 0, sizeof(unsigned long long), st));
    if (s_hi >= s_lo) {
        k2_selftest_division<<<s_hi - s_lo + 1, 256, 0, st>>>(s_lo, s_hi, ctx->d_counters + 5);
        ISB_LAUNCH_CHECK();
    }
    ISB_CUDA(cudaMemcpyAsync(ctx->h_counters + 5, ctx->d_counters + 5, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    *h_mismatches = ctx->h_counters[5];
    return ISB_OK;
}
