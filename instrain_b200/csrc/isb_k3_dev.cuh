// isb_k3_dev.cuh -- device helpers shared by the linkage front ends: the stand-alone K3 kernels (isb_k3_linkage.cu) and
// the fused pileup + SNV call + site-row kernel of the read-major path (isb_k1f_fused.cu).
// Reference: update_linked_reads (inStrain/profile/linkage.py:254-283) -- which (read pair, base) entries a site has.
#pragma once
#include "isb_common.cuh"

// split = last split whose start <= abs_pos, if abs_pos <= its end (else -1).  Splits: fasta.py:56-73.
__device__ __forceinline__ int isb_site_split(const int32_t *__restrict__ splits, int n_splits, int64_t abs_pos)
{
    int s_lo = 0, s_hi = n_splits;
    while (s_lo < s_hi) {
        const int mid = (s_lo + s_hi) >> 1;
        if ((int64_t)__ldg(splits + 2 * mid) <= abs_pos) s_lo = mid + 1; else s_hi = mid;
    }
    int split = s_lo - 1;
    if (split >= 0 && abs_pos > (int64_t)__ldg(splits + 2 * split + 1)) split = -1;
    return split;
}

// One candidate of a site in a read-major batch: is position abs_pos covered by segment g with a passing A/C/T/G base?
__device__ __forceinline__ bool k3r_candidate(const isb_reads_dev &rd, int64_t g, int64_t abs_pos, int &b, int &id)
{
    const int32_t s = __ldg(rd.seg_start + g);
    const int j = (int)(abs_pos - (int64_t)s);
    if (j < 0 || j >= (int)__ldg(rd.seg_len + g)) return false;
    const int jn = j + (s & 7);                                       // position-aligned stream: nibble index in the segment's words
    const uint32_t w = __ldg(rd.words + __ldg(rd.seg_word + g) + (jn >> 3));
    const uint32_t code = (w >> ((jn & 7) << 2)) & 15u;
    if (!code) return false;
    b = __ffs((int)code) - 1;                                       // one-hot A,C,T,G
    id = __ldg(rd.seg_pair + g);
    return true;
}

// Set the bits of one (pair id, base) entry in a site's bit rows  any | ge1[rank(b)] | ge2[rank(b)]  with atomics (the
// exact, slow way: used by the stand-alone front ends and as the duplicate-entry fallback of the fused kernel).
// Returns true when the pair already had an entry on this site.
__device__ __forceinline__ bool k3_row_set(uint32_t *any, const isb_site_meta &m, int na, unsigned bases, int b, int id,
                                           unsigned int *d_err)
{
    const int w = (id >> 5) - m.wlo;
    const uint32_t bit = 1u << (id & 31);
    const int r = __popc(bases & ((1u << b) - 1u));
    const bool dbl = atomicOr(any + w, bit) & bit;                    // second entry of this pair on this site
    uint32_t *ge1 = any + (size_t)(1 + r) * m.nw;
    if (atomicOr(ge1 + w, bit) & bit) {
        uint32_t *ge2 = any + (size_t)(1 + na + r) * m.nw;
        if (atomicOr(ge2 + w, bit) & bit) atomicOr(d_err, ISB_DEV_ERR_MULT);
    }
    return dbl;
}
