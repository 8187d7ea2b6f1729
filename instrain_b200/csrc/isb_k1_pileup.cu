// isb_k1_pileup.cu -- K1: per-position x mm x {A,C,T,G} pileup counts on sm_100a.
//
// Replaces pysam's pileup-column iteration + get_base_counts_mm (inStrain/profile/profile_utilities.py:268-286).
// HBM-bound integer histogram (no tensor cores): algorithmic traffic = 10 B per event read (ref_pos i32,
// base u8, qual u8, read_id i32) + 16*M B per position written.
//
// Main kernel (position-major events): one CTA owns a tile of TP consecutive positions.  Because events are
// sorted by position, the tile's events are ONE contiguous slice [tile_off[t], tile_off[t+1]) of every column,
// so every event byte is read exactly once, fully coalesced with 128-bit loads, and the tile's counters live in
// shared memory.  Equal (position, mm, base) keys inside a warp are merged with __match_any_sync before ONE
// shared-memory atomic per distinct key (a position's ~c events agree on 1-2 bases, so a 32-event slot
// collapses to a handful of atomics).  The finished tile is written back with 128-bit stores: global counters are
// never touched by atomics.
#include "isb_common.cuh"

#define K1_THREADS 256

__global__ void k1_tile_offsets(const int32_t *__restrict__ ref_pos, int64_t n, int32_t start, int32_t L, int TP,
                                int n_tiles, int64_t *__restrict__ tile_off)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    int64_t rel = (int64_t)t * TP;
    if (rel > L) rel = L;
    tile_off[t] = isb_lower_bound(ref_pos, 0, n, (int64_t)start + rel);
}

template <bool kM1>
__global__ void __launch_bounds__(K1_THREADS)
k1_pileup_tiles(const int32_t *__restrict__ ref_pos, const uint8_t *__restrict__ base,
                const uint8_t *__restrict__ qual, const int32_t *__restrict__ read_id,
                const uint8_t *__restrict__ pair_mm, const int64_t *__restrict__ tile_off, int64_t n,
                int32_t start, int32_t L, int M, int TP, int min_qual, int32_t *__restrict__ counts,
                unsigned long long *__restrict__ nmask, unsigned int *__restrict__ d_err)
{
    extern __shared__ __align__(16) int32_t s_cnt[];
    const int tile = blockIdx.x;
    const int p0 = tile * TP;
    const int np = min(TP, L - p0);
    const int n_cnt4 = np * M;                         // int4 elements of this tile
    for (int i = threadIdx.x; i < n_cnt4; i += K1_THREADS) reinterpret_cast<int4 *>(s_cnt)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();

    const int64_t e_lo = tile_off[tile], e_hi = tile_off[tile + 1];
    const int lane = threadIdx.x & 31;
    const int64_t g_lo = e_lo >> 2, g_hi = (e_hi + 3) >> 2;     // groups of 4 events (128-bit loads)
    const int32_t rel0 = start + p0;

    for (int64_t gb = g_lo + (threadIdx.x & ~31); gb < g_hi; gb += K1_THREADS) {
        const int64_t g = gb + lane;
        int32_t pos[4];
        uint32_t b4 = 0, q4 = 0;
        int32_t rid[4] = {0, 0, 0, 0};
        const int64_t e0 = g << 2;
        if (g < g_hi && e0 + 4 <= n) {
            const int4 pv = __ldg(reinterpret_cast<const int4 *>(ref_pos) + g);
            pos[0] = pv.x; pos[1] = pv.y; pos[2] = pv.z; pos[3] = pv.w;
            b4 = __ldg(reinterpret_cast<const uint32_t *>(base) + g);
            q4 = __ldg(reinterpret_cast<const uint32_t *>(qual) + g);
            if (!kM1) {
                const int4 rv = __ldg(reinterpret_cast<const int4 *>(read_id) + g);
                rid[0] = rv.x; rid[1] = rv.y; rid[2] = rv.z; rid[3] = rv.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t e = e0 + j;
                pos[j] = 0;
                if (g < g_hi && e < n) {
                    pos[j] = ref_pos[e];
                    b4 |= (uint32_t)base[e] << (8 * j);
                    q4 |= (uint32_t)qual[e] << (8 * j);
                    if (!kM1) rid[j] = read_id[e];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t e = e0 + j;
            const int b = (b4 >> (8 * j)) & 0xff;
            const int q = (q4 >> (8 * j)) & 0xff;
            bool valid = (e >= e_lo) && (e < e_hi) && (q >= min_qual);
            const int p = pos[j] - rel0;
            if (valid && (p < 0 || p >= np)) { atomicOr(d_err, ISB_DEV_ERR_ORDER); valid = false; }
            int mm = 0;
            if (!kM1 && valid) {
                mm = __ldg(pair_mm + rid[j]);
                if (mm >= M) { atomicOr(d_err, ISB_DEV_ERR_MM); valid = false; }
            }
            if (valid && b >= 4) {                      // in-alignment non-ACGT base: only marks the mm level present
                if (nmask) atomicOr(nmask + p0 + p, 1ull << mm);
                valid = false;
            }
            const int key = valid ? ((p * M + mm) << 2) + b : -1;
            const unsigned peers = __match_any_sync(ISB_FULL, key);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[key], __popc(peers));
        }
    }
    __syncthreads();
    int4 *dst = reinterpret_cast<int4 *>(counts) + (size_t)p0 * M;
    for (int i = threadIdx.x; i < n_cnt4; i += K1_THREADS) dst[i] = reinterpret_cast<const int4 *>(s_cnt)[i];
}

// Any-order fallback: one global atomic per qualifying event (the baseline the tiled kernel is measured against).
__global__ void __launch_bounds__(256)
k1_pileup_atomic(const int32_t *__restrict__ ref_pos, const uint8_t *__restrict__ base, const uint8_t *__restrict__ qual,
                 const int32_t *__restrict__ read_id, const uint8_t *__restrict__ pair_mm, int64_t n, int32_t start,
                 int32_t L, int M, int min_qual, int32_t *__restrict__ counts, unsigned long long *__restrict__ nmask,
                 unsigned int *__restrict__ d_err)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        if (qual[e] < min_qual) continue;
        const int64_t p = (int64_t)ref_pos[e] - start;
        if (p < 0 || p >= L) continue;
        int mm = 0;
        if (M > 1) {
            mm = pair_mm[read_id[e]];
            if (mm >= M) { atomicOr(d_err, ISB_DEV_ERR_MM); continue; }
        }
        const int b = base[e];
        if (b < 4) atomicAdd(counts + ((size_t)p * M + mm) * 4 + b, 1);
        else if (nmask) atomicOr(nmask + p, 1ull << mm);
    }
}

static int k1_tile_positions(int M)
{
    int tp = 1024;                                  // 16 KB of counters at M = 1
    while (tp > 32 && (size_t)tp * M * 16 > 48 * 1024) tp >>= 1;
    return tp;
}

int isb_k1_launch(isb_ctx *ctx, int64_t n, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                  const int32_t *read_id, const uint8_t *pair_mm, int32_t start, int32_t L, int M, int min_qual,
                  uint32_t flags, int32_t *counts, unsigned long long *nmask)
{
    cudaStream_t st = ctx->stream;
    if (nmask) ISB_CUDA(cudaMemsetAsync(nmask, 0, sizeof(unsigned long long) * (size_t)L, st));
    if (L <= 0) return ISB_OK;
    if (flags & ISB_K1_ANY_ORDER) {
        ISB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)L * M * 4, st));
        if (n > 0) {
            int64_t blocks = (n + 255) / 256;
            int grid = (int)(blocks < (int64_t)ctx->sm_count * 16 ? blocks : (int64_t)ctx->sm_count * 16);
            k1_pileup_atomic<<<grid, 256, 0, st>>>(ref_pos, base, qual, read_id, pair_mm, n, start, L, M, min_qual,
                                                    counts, nmask, ctx->d_err);
            ISB_LAUNCH_CHECK();
        }
        return ISB_OK;
    }
    if (((uintptr_t)ref_pos | (uintptr_t)read_id) & 15 || ((uintptr_t)base | (uintptr_t)qual) & 3)
        return isb_fail(ctx, ISB_ERR_ARG, "isb_pileup_counts: event columns must be 16-byte aligned (ref_pos, read_id) / 4-byte aligned (base, qual)");
    const int TP = k1_tile_positions(M);
    const int n_tiles = (L + TP - 1) / TP;
    int rc = isb_ensure(ctx, SL_TILE_OFF, sizeof(int64_t) * ((size_t)n_tiles + 1));
    if (rc) return rc;
    int64_t *tile_off = (int64_t *)ctx->buf[SL_TILE_OFF].p;
    k1_tile_offsets<<<(n_tiles + 1 + 255) / 256, 256, 0, st>>>(ref_pos, n, start, L, TP, n_tiles, tile_off);
    ISB_LAUNCH_CHECK();
    const size_t smem = (size_t)TP * M * 16;
    if (M == 1)
        k1_pileup_tiles<true><<<n_tiles, K1_THREADS, smem, st>>>(ref_pos, base, qual, read_id, pair_mm, tile_off, n,
                                                                 start, L, M, TP, min_qual, counts, nmask, ctx->d_err);
    else
        k1_pileup_tiles<false><<<n_tiles, K1_THREADS, smem, st>>>(ref_pos, base, qual, read_id, pair_mm, tile_off, n,
                                                                  start, L, M, TP, min_qual, counts, nmask, ctx->d_err);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
