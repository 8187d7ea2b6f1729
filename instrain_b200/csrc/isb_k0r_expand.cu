// isb_k0r_expand.cu -- K0r: compact read-major TRANSFER format -> the aligned-segment nibble stream K1r / K3 read, sm_100a.
//
// Only on the host-buffer (end-to-end) path.  The nibble stream costs 4 bits per aligned base + 8 bytes of word offset
// per segment over PCIe; the compact form carries the same information in 3 bits per aligned base (2-bit base code +
// 1 event bit, in position-aligned units of 8 bases: one uint16 + one uint8) and no word offsets: the stream layout is canonical (one
// leading zero word, [data words + one zero word] per segment in table order, zero padding to a multiple of 4 words),
// so seg_word is an exclusive scan of the segments' word counts ceil((start % 8 + seg_len) / 8) done here.  The expanded stream is written to context-owned HBM
// (0.5 B per base at HBM speed, against 0.375 B per base saved on a link that is ~100x slower).
#include "isb_common.cuh"
#include "isb_scan.cuh"

struct UnitsFn {   // data words (= position-aligned 8-base units) of segment i
    const int32_t *seg_start;
    const uint16_t *seg_len;
    __device__ int operator()(int64_t i) const { return ((seg_start[i] & 7) + (int)seg_len[i] + 7) >> 3; }
};
struct SegWordSink {   // seg_word[i] = leading zero word + units before + one separator per earlier segment
    int64_t *seg_word;
    __device__ void operator()(int64_t i, int64_t prefix, int) const { seg_word[i] = 1 + prefix + i; }
};

// one warp per segment, lanes over its units: coalesced 2-byte / 1-byte loads, coalesced word stores
__global__ void __launch_bounds__(256)
k0r_expand_units(int64_t n_segs, const int32_t *__restrict__ seg_start, const uint16_t *__restrict__ seg_len,
                 const int64_t *__restrict__ seg_word,
                 int64_t n_units, const uint16_t *__restrict__ base2, const uint8_t *__restrict__ pass, int64_t n_words,
                 uint32_t *__restrict__ words, unsigned int *__restrict__ d_err)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n_segs; i += n_warps) {
        const int nw = ((seg_start[i] & 7) + (int)seg_len[i] + 7) >> 3;
        const int64_t w0 = seg_word[i];
        const int64_t u0 = w0 - 1 - i;
        if (u0 < 0 || u0 + nw > n_units || w0 + nw + 1 > n_words) {       // table and unit arrays disagree (uniform)
            if (lane == 0) atomicOr(d_err, ISB_DEV_ERR_SEG);
            continue;
        }
        for (int k = lane; k < nw; k += 32) {
            const uint32_t b = base2[u0 + k], p = pass[u0 + k];
            uint32_t w = 0u;
#pragma unroll
            for (int t = 0; t < 8; ++t) w |= (((p >> t) & 1u) << ((b >> (2 * t)) & 3u)) << (4 * t);
            words[w0 + k] = w;
        }
    }
}

// seg_word[] and the nibble stream (n_words = round4(1 + n_units + n_segs)) from the compact columns; all on ctx->stream
int isb_k0r_launch(isb_ctx *ctx, int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, int64_t n_units, const uint16_t *base2,
                   const uint8_t *pass, int64_t *seg_word, int64_t n_words, uint32_t *words)
{
    cudaStream_t st = ctx->stream;
    ISB_CUDA(cudaMemsetAsync(words, 0, sizeof(uint32_t) * (size_t)n_words, st));
    if (n_segs <= 0) return ISB_OK;
    int rc;
    const int nb = (int)((n_segs + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if ((rc = isb_ensure(ctx, SL_SCAN_TMP, sizeof(int64_t) * (size_t)nb))) return rc;
    int64_t *block_sums = (int64_t *)ctx->buf[SL_SCAN_TMP].p;
    UnitsFn uf{seg_start, seg_len};
    scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(uf, n_segs, block_sums);
    ISB_LAUNCH_CHECK();
    scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nb, ctx->d_counters + 6);
    ISB_LAUNCH_CHECK();
    SegWordSink sink{seg_word};
    scan_scatter<<<nb, SCAN_THREADS, 0, st>>>(uf, n_segs, block_sums, sink);
    ISB_LAUNCH_CHECK();
    const int64_t blocks = (n_segs * 32 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)ctx->sm_count * 16 ? blocks : (int64_t)ctx->sm_count * 16);
    k0r_expand_units<<<grid, 256, 0, st>>>(n_segs, seg_start, seg_len, seg_word, n_units, base2, pass, n_words, words, ctx->d_err);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

// ---- reference-delta TRANSFER format (isb_reads_delta) ---------------------------------------------------------------
// Reads are almost everywhere identical to the reference, and the reference crosses PCIe anyway (K2 needs it).  The delta
// format therefore sends, per 8-base unit, only the event bits (`pass`, 1 byte) and, per passing base that DIFFERS from the
// reference, one 5-byte entry (word index of the canonical stream + nibble position + XOR of the two one-hot codes):
// ~1.1 bits + change per aligned base instead of 3.  K0d rebuilds the same nibble stream as K0r: every unit gets the
// one-hot codes of its 8 reference bases masked by the event bits, then the entries flip the differing nibbles.

// One THREAD per segment walks the segment's units (<= 33): the three table loads of every segment are in flight at
// once and the unit loads of a thread are independent of each other (unrolled: issued in batches).  The first version
// (one warp per segment, lanes over units, grid-stride) ran 35 dependent table -> unit -> store chains per warp with
// 20 of 32 lanes busy: 173 us per 1e8 aligned bases, more than the whole K1r -> K2 -> K3 that follows it.
// ref is indexed by batch coordinate - start.
__global__ void __launch_bounds__(256)
k0d_expand_units(int64_t n_segs, const int32_t *__restrict__ seg_start, const uint16_t *__restrict__ seg_len,
                 const int64_t *__restrict__ seg_word, int64_t n_units, const uint8_t *__restrict__ pass,
                 const uint8_t *__restrict__ ref, int32_t start, int32_t L, int64_t n_words, uint32_t *__restrict__ words,
                 unsigned int *__restrict__ d_err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_segs) return;
    const bool ref_aligned = (reinterpret_cast<uintptr_t>(ref) & 7) == 0;
    const int32_t s = seg_start[i];
    const int nw = ((s & 7) + (int)seg_len[i] + 7) >> 3;
    const int64_t w0 = seg_word[i];
    const int64_t u0 = w0 - 1 - i;
    const int64_t r0 = (int64_t)(s & ~7) - start;                          // reference index of the first unit's first base
    if (u0 < 0 || u0 + nw > n_units || w0 + nw + 1 > n_words || r0 < 0 || r0 + 8 * (int64_t)(nw - 1) >= L) {
        atomicOr(d_err, ISB_DEV_ERR_SEG);                                   // table, unit arrays or range disagree
        return;
    }
#pragma unroll 4
    for (int k = 0; k < nw; ++k) {
        const uint32_t ps = __ldg(pass + u0 + k);
        const int64_t r = r0 + 8 * (int64_t)k;
        uint32_t w = 0u;
        if (ref_aligned && r + 8 <= L) {                                    // the unit's 8 reference bases in one load
            const uint2 rr = __ldg(reinterpret_cast<const uint2 *>(ref + r));
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t c = ((t < 4 ? rr.x : rr.y) >> (8 * (t & 3))) & 0xffu;
                if (((ps >> t) & 1u) && c < 4u) w |= (1u << c) << (4 * t);
            }
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t c = (r + t < L) ? (uint32_t)__ldg(ref + r + t) : 4u;
                if (((ps >> t) & 1u) && c < 4u) w |= (1u << c) << (4 * t);
            }
        }
        words[w0 + k] = w;
    }
}

__global__ void __launch_bounds__(256)
k0d_apply_mismatches(int64_t n_mis, const uint32_t *__restrict__ mis_word, const uint8_t *__restrict__ mis_code,
                     int64_t n_words, uint32_t *__restrict__ words, unsigned int *__restrict__ d_err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mis) return;
    const uint32_t w = mis_word[i], e = mis_code[i];
    if ((int64_t)w >= n_words || (e >> 7)) { atomicOr(d_err, ISB_DEV_ERR_SEG); return; }
    atomicXor(words + w, (e & 15u) << (4 * (e >> 4)));
}

int isb_k0d_launch(isb_ctx *ctx, int64_t n_segs, const int32_t *seg_start, const uint16_t *seg_len, int64_t n_units,
                   const uint8_t *pass, const uint8_t *ref, int32_t start, int32_t L, int64_t n_mis, const uint32_t *mis_word,
                   const uint8_t *mis_code, int64_t *seg_word, int64_t n_words, uint32_t *words)
{
    cudaStream_t st = ctx->stream;
    ISB_CUDA(cudaMemsetAsync(words, 0, sizeof(uint32_t) * (size_t)n_words, st));
    if (n_segs <= 0) return ISB_OK;
    int rc;
    const int nb = (int)((n_segs + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if ((rc = isb_ensure(ctx, SL_SCAN_TMP, sizeof(int64_t) * (size_t)nb))) return rc;
    int64_t *block_sums = (int64_t *)ctx->buf[SL_SCAN_TMP].p;
    UnitsFn uf{seg_start, seg_len};
    scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(uf, n_segs, block_sums);
    ISB_LAUNCH_CHECK();
    scan_blocksums<<<1, 1024, 0, st>>>(block_sums, nb, ctx->d_counters + 6);
    ISB_LAUNCH_CHECK();
    SegWordSink sink{seg_word};
    scan_scatter<<<nb, SCAN_THREADS, 0, st>>>(uf, n_segs, block_sums, sink);
    ISB_LAUNCH_CHECK();
    k0d_expand_units<<<(unsigned)((n_segs + 255) / 256), 256, 0, st>>>(n_segs, seg_start, seg_len, seg_word, n_units, pass, ref, start, L,
                                                                       n_words, words, ctx->d_err);
    ISB_LAUNCH_CHECK();
    if (n_mis > 0) {
        k0d_apply_mismatches<<<(unsigned)((n_mis + 255) / 256), 256, 0, st>>>(n_mis, mis_word, mis_code, n_words, words, ctx->d_err);
        ISB_LAUNCH_CHECK();
    }
    return ISB_OK;
}
