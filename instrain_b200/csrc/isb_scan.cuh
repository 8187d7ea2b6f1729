// isb_scan.cuh -- small device-wide ordered exclusive scan / compaction (plumbing for K3; not a hot kernel).
// Three launches: per-block sums, single-CTA scan of the block sums, per-block scan + sink.
#pragma once
#include "isb_common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

template <class F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce(F f, int64_t n, int64_t *__restrict__ block_sums)
{
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) v += f(base + k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(ISB_FULL, v, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}

// single CTA: in-place exclusive scan of block_sums[0..nb), grand total -> *total
static __global__ void __launch_bounds__(1024) scan_blocksums(int64_t *__restrict__ block_sums, int nb,
                                                       unsigned long long *__restrict__ total)
{
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int64_t v = i < nb ? block_sums[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t u = __shfl_up_sync(ISB_FULL, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
            int64_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t u = __shfl_up_sync(ISB_FULL, wi, d);
                if (lane >= d) wi += u;
            }
            s_warp[lane] = wi - w;   // exclusive prefix of warp sums
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t excl = carry + s_warp[warp] + incl - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (unsigned long long)s_carry;
}

// sink(i, exclusive_prefix, value) is called for every i < n, prefixes in index order
template <class F, class Sink>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_scatter(F f, int64_t n, const int64_t *__restrict__ block_off, Sink sink)
{
    __shared__ int64_t s_warp[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        vals[k] = (base + k < n) ? f(base + k) : 0;
        v += vals[k];
    }
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t u = __shfl_up_sync(ISB_FULL, incl, d);
        if (lane >= d) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    int64_t prefix = block_off[blockIdx.x] + woff + incl - v;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) sink(base + k, prefix, vals[k]);
        prefix += vals[k];
    }
}
