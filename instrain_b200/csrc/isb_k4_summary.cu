// isb_k4_summary.cu -- K4: per-scaffold, per-mm merge-stage summary reductions (the row after the hot path).
//
// Replaces the numeric core of make_coverage_table (inStrain/profile/profile_utilities.py:425-506) with its helpers
// mm_counts_to_counts_shrunk (:508-532) and get_basewise_clons (:534-546): for every scaffold and every mm level,
// over the scaffold's positions,
//   cumulative coverage  c_m(p) = sum_{m' <= m} covT[p][m']      -> breadth, mean, std, SEM, median
//   clonality            the value of the highest level <= m at which clonT[p][.] is set  -> count, mean, median
// The reference does this with one pandas Series.add per mm level and a Python dict update per level (27 % of its
// profile CPU time in its own run report); here it is one pass (k4_cumulate: thread per position, block-aggregated
// atomics) plus exact medians by a 4-pass MSB radix select (k4_select_hist / k4_select_update) on the materialised
// cumulative arrays.  Sums are integers (exact); the clonality sum is a double sum of float32 values.
// The SNV-count columns of the table (calc_snps, snv_utilities.py:249-272) are derived from the SNV rows on the host
// (instrain_b200/summary.py): a few thousand rows, no device work.
#include "isb_common.cuh"
#include <math_constants.h>

#define K4_THREADS 256
#define K4_MC 16                 // mm levels per select launch (shared-memory histogram = MC * 2 * 256 * 4 B = 32 KB)

struct k4_sel_state {            // one per (scaffold, level, query); query 0 = lower middle rank, 1 = upper middle rank
    unsigned long long krem;     // remaining rank inside the bucket chain fixed so far
    unsigned int prefix;         // key bits fixed so far
    int active;
};

__device__ __forceinline__ int k4_segment(const int32_t *__restrict__ off, int n_seg, int32_t p)
{
    int lo = 0, hi = n_seg;      // last s with off[s] <= p
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= p) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

// cumulative arrays + all sums
__global__ void __launch_bounds__(K4_THREADS)
k4_cumulate(int32_t L, int M, const int32_t *__restrict__ covT, const float *__restrict__ clonT,
            const unsigned long long *__restrict__ nmask, int n_seg, const int32_t *__restrict__ seg_off,
            uint32_t *__restrict__ cumcov, float *__restrict__ clonlast, isb_summary_row *__restrict__ out)
{
    __shared__ long long s_i[K4_THREADS / 32][4];
    __shared__ double s_d[K4_THREADS / 32];
    const int32_t p = blockIdx.x * K4_THREADS + threadIdx.x;
    const bool act = p < L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = act ? k4_segment(seg_off, n_seg, p) : -1;
    const int seg0 = k4_segment(seg_off, n_seg, min(blockIdx.x * K4_THREADS, L - 1));
    const bool uniform = __syncthreads_and(!act || seg == seg0);
    const unsigned long long nm = (act && nmask) ? nmask[p] : 0ull;
    unsigned int cum = 0;
    float last = CUDART_NAN_F;
    for (int m = 0; m < M; ++m) {
        int e = 0;
        if (act) {
            e = covT[(size_t)p * M + m];
            cum += (unsigned)e;
            const float c = clonT[(size_t)p * M + m];
            if (!isnan(c)) last = c;
            cumcov[(size_t)p * M + m] = cum;
            clonlast[(size_t)p * M + m] = last;
        }
        long long v_nz = act && cum > 0, v_sum = act ? cum : 0, v_cnt = act && !isnan(last);
        long long v_pres = act && (e > 0 || ((nm >> m) & 1ull));
        unsigned long long v_sq = act ? (unsigned long long)cum * cum : 0ull;
        double v_cl = (act && !isnan(last)) ? (double)last : 0.0;
        if (uniform) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                v_nz += __shfl_xor_sync(ISB_FULL, v_nz, d);
                v_sum += __shfl_xor_sync(ISB_FULL, v_sum, d);
                v_cnt += __shfl_xor_sync(ISB_FULL, v_cnt, d);
                v_pres += __shfl_xor_sync(ISB_FULL, v_pres, d);
                v_sq += __shfl_xor_sync(ISB_FULL, v_sq, d);
                v_cl += __shfl_xor_sync(ISB_FULL, v_cl, d);
            }
            if (lane == 0) {
                s_i[warp][0] = v_nz; s_i[warp][1] = v_sum; s_i[warp][2] = v_cnt; s_i[warp][3] = v_pres;
                s_d[warp] = v_cl;
            }
            // sum of squares goes through a second slot to keep the shared arrays small
            __syncthreads();
            long long t_nz = 0, t_sum = 0, t_cnt = 0, t_pres = 0;
            double t_cl = 0.0;
            if (threadIdx.x == 0) {
                for (int w = 0; w < K4_THREADS / 32; ++w) {
                    t_nz += s_i[w][0]; t_sum += s_i[w][1]; t_cnt += s_i[w][2]; t_pres += s_i[w][3];
                    t_cl += s_d[w];
                }
            }
            __syncthreads();
            if (lane == 0) s_i[warp][0] = (long long)v_sq;
            __syncthreads();
            if (threadIdx.x == 0 && seg0 >= 0) {
                unsigned long long t_sq = 0;
                for (int w = 0; w < K4_THREADS / 32; ++w) t_sq += (unsigned long long)s_i[w][0];
                isb_summary_row *r = out + (size_t)seg0 * M + m;
                if (t_nz) atomicAdd((unsigned long long *)&r->nonzero, (unsigned long long)t_nz);
                if (t_sum) atomicAdd((unsigned long long *)&r->sum_cov, (unsigned long long)t_sum);
                if (t_sq) atomicAdd((unsigned long long *)&r->sum_cov2, t_sq);
                if (t_cnt) atomicAdd((unsigned long long *)&r->counted, (unsigned long long)t_cnt);
                if (t_cnt) atomicAdd(&r->sum_clon, t_cl);
                if (t_pres) atomicOr(&r->present, 1);
            }
            __syncthreads();
        } else if (act && seg >= 0) {                           // block straddles a scaffold boundary: plain atomics
            isb_summary_row *r = out + (size_t)seg * M + m;
            if (v_nz) atomicAdd((unsigned long long *)&r->nonzero, 1ull);
            if (v_sum) atomicAdd((unsigned long long *)&r->sum_cov, (unsigned long long)v_sum);
            if (v_sq) atomicAdd((unsigned long long *)&r->sum_cov2, v_sq);
            if (v_cnt) { atomicAdd((unsigned long long *)&r->counted, 1ull); atomicAdd(&r->sum_clon, v_cl); }
            if (v_pres) atomicOr(&r->present, 1);
        }
    }
}

// ranks of the two middle order statistics: kind 0 = coverage over all positions, kind 1 = clonality over counted ones
__global__ void k4_select_init(int n_seg, int M, const int32_t *__restrict__ seg_off, isb_summary_row *__restrict__ out,
                               int kind, k4_sel_state *__restrict__ st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg * M) return;
    const int s = i / M;
    isb_summary_row *r = out + i;
    const long long len = (long long)seg_off[s + 1] - seg_off[s];
    if (kind == 0) r->length = len;
    const long long n = kind == 0 ? len : r->counted;
    for (int q = 0; q < 2; ++q) {
        k4_sel_state x;
        x.prefix = 0;
        x.active = n > 0;
        x.krem = n > 0 ? (unsigned long long)(q == 0 ? (n - 1) / 2 : n / 2) : 0ull;
        st[(size_t)i * 2 + q] = x;
    }
}

// one radix pass: histogram of byte (key >> shift) & 255 among the keys matching the prefix fixed so far
__global__ void __launch_bounds__(K4_THREADS)
k4_select_hist(int32_t L, int M, int m0, int mc, const uint32_t *__restrict__ cumcov, const float *__restrict__ clonlast,
               int n_seg, const int32_t *__restrict__ seg_off, int kind, int shift, const k4_sel_state *__restrict__ st,
               unsigned int *__restrict__ hist)
{
    extern __shared__ unsigned int s_h[];                     // [mc][2][256]
    const int32_t p = blockIdx.x * K4_THREADS + threadIdx.x;
    const bool act = p < L;
    const int seg = act ? k4_segment(seg_off, n_seg, p) : -1;
    const int seg0 = k4_segment(seg_off, n_seg, min(blockIdx.x * K4_THREADS, L - 1));
    const bool uniform = __syncthreads_and(!act || seg == seg0);
    const int n_h = mc * 2 * 256;
    if (uniform) {
        for (int i = threadIdx.x; i < n_h; i += K4_THREADS) s_h[i] = 0u;
        __syncthreads();
    }
    if (act && seg >= 0) {
        for (int j = 0; j < mc; ++j) {
            const int m = m0 + j;
            unsigned int key;
            if (kind == 0) key = cumcov[(size_t)p * M + m];
            else {
                const float c = clonlast[(size_t)p * M + m];
                if (isnan(c)) continue;
                key = __float_as_uint(c);
            }
            const unsigned int byte = (key >> shift) & 255u;
            for (int q = 0; q < 2; ++q) {
                const k4_sel_state x = st[((size_t)seg * M + m) * 2 + q];
                if (!x.active) continue;
                if (shift < 24 && (key >> (shift + 8)) != (x.prefix >> (shift + 8))) continue;
                if (uniform) atomicAdd(&s_h[(j * 2 + q) * 256 + byte], 1u);
                else atomicAdd(&hist[(((size_t)seg * M + m) * 2 + q) * 256 + byte], 1u);
            }
        }
    }
    if (uniform) {
        __syncthreads();
        if (seg0 >= 0)
            for (int i = threadIdx.x; i < n_h; i += K4_THREADS) {
                const unsigned int v = s_h[i];
                if (v) {
                    const int j = i / 512, q = (i >> 8) & 1, b = i & 255;
                    atomicAdd(&hist[(((size_t)seg0 * M + m0 + j) * 2 + q) * 256 + b], v);
                }
            }
    }
}

__global__ void k4_select_update(int n_items, int shift, const unsigned int *__restrict__ hist, k4_sel_state *__restrict__ st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (seg, m, q)
    if (i >= n_items) return;
    k4_sel_state x = st[i];
    if (!x.active) return;
    const unsigned int *h = hist + (size_t)i * 256;
    unsigned long long acc = 0;
    int b = 0;
    for (; b < 256; ++b) {
        const unsigned long long c = h[b];
        if (x.krem < acc + c) break;
        acc += c;
    }
    if (b == 256) { x.active = 0; st[i] = x; return; }         // cannot happen for consistent inputs
    x.krem -= acc;
    x.prefix |= (unsigned int)b << shift;
    st[i] = x;
}

__global__ void k4_select_store(int n_rows, int kind, const k4_sel_state *__restrict__ st, isb_summary_row *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const k4_sel_state lo = st[(size_t)i * 2], hi = st[(size_t)i * 2 + 1];
    if (kind == 0) {
        out[i].cov_med_lo = lo.active ? (int32_t)lo.prefix : 0;
        out[i].cov_med_hi = hi.active ? (int32_t)hi.prefix : 0;
    } else {
        out[i].clon_med_lo = lo.active ? __uint_as_float(lo.prefix) : CUDART_NAN_F;
        out[i].clon_med_hi = hi.active ? __uint_as_float(hi.prefix) : CUDART_NAN_F;
    }
}

int isb_k4_launch(isb_ctx *ctx, int32_t L, int M, const int32_t *covT, const float *clonT, const unsigned long long *nmask,
                  int n_seg, const int32_t *seg_off, isb_summary_row *out)
{
    cudaStream_t st = ctx->stream;
    int rc;
    const size_t n_rows = (size_t)n_seg * M;
    ISB_CUDA(cudaMemsetAsync(out, 0, sizeof(isb_summary_row) * n_rows, st));
    if (L <= 0 || n_seg <= 0) return ISB_OK;
    if ((rc = isb_ensure(ctx, SL_K4_CUM, sizeof(uint32_t) * (size_t)L * M))) return rc;
    if ((rc = isb_ensure(ctx, SL_K4_CLON, sizeof(float) * (size_t)L * M))) return rc;
    if ((rc = isb_ensure(ctx, SL_K4_STATE, sizeof(k4_sel_state) * n_rows * 2))) return rc;
    if ((rc = isb_ensure(ctx, SL_K4_HIST, sizeof(unsigned int) * n_rows * 2 * 256))) return rc;
    uint32_t *cum = (uint32_t *)ctx->buf[SL_K4_CUM].p;
    float *cl = (float *)ctx->buf[SL_K4_CLON].p;
    k4_sel_state *state = (k4_sel_state *)ctx->buf[SL_K4_STATE].p;
    unsigned int *hist = (unsigned int *)ctx->buf[SL_K4_HIST].p;
    const int grid = (L + K4_THREADS - 1) / K4_THREADS;
    k4_cumulate<<<grid, K4_THREADS, 0, st>>>(L, M, covT, clonT, nmask, n_seg, seg_off, cum, cl, out);
    ISB_LAUNCH_CHECK();
    const int g_rows = (int)((n_rows + 255) / 256), g_items = (int)((n_rows * 2 + 255) / 256);
    for (int kind = 0; kind < 2; ++kind) {
        k4_select_init<<<g_rows, 256, 0, st>>>(n_seg, M, seg_off, out, kind, state);
        ISB_LAUNCH_CHECK();
        for (int shift = 24; shift >= 0; shift -= 8) {
            ISB_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned int) * n_rows * 2 * 256, st));
            for (int m0 = 0; m0 < M; m0 += K4_MC) {
                const int mc = M - m0 < K4_MC ? M - m0 : K4_MC;
                k4_select_hist<<<grid, K4_THREADS, sizeof(unsigned int) * mc * 2 * 256, st>>>(
                    L, M, m0, mc, cum, cl, n_seg, seg_off, kind, shift, state, hist);
                ISB_LAUNCH_CHECK();
            }
            k4_select_update<<<g_items, 256, 0, st>>>((int)(n_rows * 2), shift, hist, state);
            ISB_LAUNCH_CHECK();
        }
        k4_select_store<<<g_rows, 256, 0, st>>>((int)n_rows, kind, state, out);
        ISB_LAUNCH_CHECK();
    }
    return ISB_OK;
}
