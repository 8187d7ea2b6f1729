// isb_k0_expand.cu -- K0: expand the compact host->device transfer format into the canonical event columns in HBM.
//
// The columnar north-star layout costs 10 B per aligned base; over PCIe that, not the kernels, bounds the end-to-end
// rate (measured: 54 GB/s => 5.4e7 positions/s at 100x).  The packed transfer format carries the same information in
// ~1 B per event + 12 B per position:
//   pos_off  int64[L+1]   CSR offsets of the positions (replaces ref_pos, 4 B/event)
//   id_base  int32[L]     pair id of the first event of the position
//   bqd      uint8[n]     bit 7 = quality >= min_qual (the only thing K1/K3 need from the quality),
//                         bits 4-6 = base code (0..4), bits 0-3 = pair-id delta to the previous event of the position
//                         (events of a position are sorted by pair id; 15 = escape, see esc_evt / esc_id)
//   esc_evt int64[n_esc], esc_id int32[n_esc]   sorted event indices with a delta > 14 and their absolute pair ids
// One warp per position: coalesced byte loads, a warp inclusive scan of the deltas, coalesced column stores, so the
// expansion runs at HBM write speed (10 B written per event) and the unchanged K1/K2/K3 follow.
#include "isb_common.cuh"

__global__ void __launch_bounds__(256)
k0_expand_packed(const int64_t *__restrict__ pos_off, const int32_t *__restrict__ id_base,
                 const uint8_t *__restrict__ bqd, int64_t n_esc, const int64_t *__restrict__ esc_evt,
                 const int32_t *__restrict__ esc_id, int32_t start, int32_t L, int qpass,
                 int32_t *__restrict__ ref_pos, uint8_t *__restrict__ base, uint8_t *__restrict__ qual,
                 int32_t *__restrict__ read_id)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp0; p < L; p += n_warps) {
        const int64_t e0 = pos_off[p], e1 = pos_off[p + 1];
        if (e1 <= e0) continue;
        int id_run = id_base[p];                                  // pair id before the current 32-event slice
        for (int64_t eb = e0; eb < e1; eb += 32) {
            const int64_t e = eb + lane;
            const bool in = e < e1;
            const unsigned byte = in ? bqd[e] : 0u;
            int d = in ? (int)(byte & 15u) : 0;
            const bool esc = in && d == 15;
            int id;
            if (__any_sync(ISB_FULL, esc)) {                      // rare: a pair-id jump > 14 inside this slice
                int abs_id = 0;
                if (esc) {
                    int64_t lo = 0, hi = n_esc;
                    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (esc_evt[mid] < e) lo = mid + 1; else hi = mid; }
                    abs_id = esc_id[lo];
                }
                // serial resolve on every lane (identical work, no divergence): running id over the 32 events
                int run = id_run;
                id = 0;
                for (int j = 0; j < 32; ++j) {
                    const int dj = __shfl_sync(ISB_FULL, d, j);
                    const int ej = __shfl_sync(ISB_FULL, (int)esc, j);
                    const int aj = __shfl_sync(ISB_FULL, abs_id, j);
                    run = ej ? aj : run + dj;
                    if (j == lane) id = run;
                }
            } else {
                int incl = d;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const int v = __shfl_up_sync(ISB_FULL, incl, s);
                    if (lane >= s) incl += v;
                }
                id = id_run + incl;
            }
            const int n_in = (int)min((int64_t)32, e1 - eb);
            id_run = __shfl_sync(ISB_FULL, id, n_in - 1);
            if (in) {
                ref_pos[e] = (int32_t)p + start;
                base[e] = (uint8_t)((byte >> 4) & 7u);
                qual[e] = (byte & 0x80u) ? (uint8_t)qpass : (uint8_t)0;
                read_id[e] = id;
            }
        }
    }
}

int isb_k0_launch(isb_ctx *ctx, int64_t n, const int64_t *pos_off, const int32_t *id_base, const uint8_t *bqd,
                  int64_t n_esc, const int64_t *esc_evt, const int32_t *esc_id, int32_t start, int32_t L, int qpass,
                  int32_t *ref_pos, uint8_t *base, uint8_t *qual, int32_t *read_id)
{
    if (L <= 0 || n <= 0) return ISB_OK;
    const int64_t blocks = ((int64_t)L * 32 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)ctx->sm_count * 32 ? blocks : (int64_t)ctx->sm_count * 32);
    k0_expand_packed<<<grid, 256, 0, ctx->stream>>>(pos_off, id_base, bqd, n_esc, esc_evt, esc_id, start, L, qpass, ref_pos,
                                                   base, qual, read_id);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
