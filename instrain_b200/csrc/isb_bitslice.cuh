// isb_bitslice.cuh -- bit-sliced (vertical) counters over one-hot nibble words, shared by the pileup kernels K1r
// (read-major segments, isb_k1r_reads.cu) and K1c (column words, isb_k1c_cols.cu).
//
// A nibble word holds the one-hot codes (A=1, C=2, T=4, G=8) of 8 consecutive positions of one read, so its 32 bits are
// the 32 (position, base) indicator bits of that read.  Eight words at a time are added into eight VERTICAL counter
// planes (plane j = bit j of 32 independent counters) with a Harley-Seal carry-save tree: 7 CSAs + a 5-plane ripple =
// 24 logic ops per 8 words.  The planes are turned into integers once per <= 248 words.
#pragma once
#include <stdint.h>

// carry-save adder on 32 independent bit lanes: h = majority(a, b, c), l = a ^ b ^ c (one LOP3 each)
#define K1R_CSA(h, l, a, b, c)                       \
    {                                                \
        const uint32_t u_ = (a) ^ (b);               \
        const uint32_t h_ = ((a) & (b)) | (u_ & (c)); \
        l = u_ ^ (c);                                \
        h = h_;                                      \
    }

// 8-bit -> 32-bit: byte j of `lo` counts position 2j, byte j of `hi` position 2j+1
__device__ __forceinline__ void k1r_widen(int (&c)[8][4], int b, uint32_t lo, uint32_t hi)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c[2 * j][b] += (int)((lo >> (8 * j)) & 0xffu);
        c[2 * j + 1][b] += (int)((hi >> (8 * j)) & 0xffu);
    }
}

// Vertical (bit-sliced) counters -> per-(position, base) integers.  Plane j holds bit j of 32 independent counters, bit
// lane 4k + b = (position k, base b).  Per base, the eight lanes are pulled out as 0/1 bytes of two words (even / odd
// positions) and summed with weight 2^j, then widened.
__device__ __forceinline__ void k1r_planes_to_counts(int (&c)[8][4], uint32_t (&pl)[8])
{
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t lo = 0u, hi = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            lo += ((pl[j] >> b) & 0x01010101u) << j;
            hi += ((pl[j] >> (b + 4)) & 0x01010101u) << j;
        }
        k1r_widen(c, b, lo, hi);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) pl[j] = 0u;
}

// Harley-Seal block: eight 1-bit inputs per lane into the planes with 7 carry-save adders + one 5-plane ripple
__device__ __forceinline__ void k1r_add8(uint32_t (&pl)[8], const uint32_t (&x)[8])
{
    uint32_t t2a, t2b, t4a, t4b, t8;
    K1R_CSA(t2a, pl[0], pl[0], x[0], x[1]);
    K1R_CSA(t2b, pl[0], pl[0], x[2], x[3]);
    K1R_CSA(t4a, pl[1], pl[1], t2a, t2b);
    K1R_CSA(t2a, pl[0], pl[0], x[4], x[5]);
    K1R_CSA(t2b, pl[0], pl[0], x[6], x[7]);
    K1R_CSA(t4b, pl[1], pl[1], t2a, t2b);
    K1R_CSA(t8, pl[2], pl[2], t4a, t4b);
#pragma unroll
    for (int j = 3; j < 8; ++j) {
        const uint32_t cy = pl[j] & t8;
        pl[j] ^= t8;
        t8 = cy;
    }
}
