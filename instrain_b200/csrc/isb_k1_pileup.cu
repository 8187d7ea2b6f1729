// isb_k1_pileup.cu -- K1: per-position x mm x {A,C,T,G} pileup counts on sm_100a.
//
// Replaces pysam's pileup-column iteration + get_base_counts_mm (inStrain/profile/profile_utilities.py:268-286).
// HBM-bound integer histogram (no tensor cores): algorithmic traffic = 10 B per event read (ref_pos i32,
// base u8, qual u8, read_id i32) + 16*M B per position written.
//
// Main kernel (position-major events): one CTA owns a tile of TP consecutive positions.  Because events are
// sorted by position, the tile's events are ONE contiguous slice [tile_off[t], tile_off[t+1]) of every column,
// so every event byte is read exactly once, fully coalesced with 128-bit loads, and the tile's counters live in
// shared memory.  The finished tile is written back with 128-bit stores: global counters are never touched by
// atomics.  Two implementations of the tile kernel follow: v0 (warp-cooperative, __match_any_sync merge; kept as the
// fallback for unaligned columns and for A/B measurements) and v2 (TMA-staged, lane-serial; the production path).
#include "isb_common.cuh"
#include <stdlib.h>

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

// ---- v2: TMA-staged, lane-serial kernel (the production path) -----------------------------------------------------
// ncu on the warp-cooperative kernels above showed them ISSUE-bound (78 % issue-active, ~250 warp instructions per 128
// events, DRAM at 31 %): every lane spends ~60 instructions per event on unpacking, validation and warp reductions.
// v2 turns the work around (used for M = 1, the --skip_mm_profiling / headline configuration; with per-pair mm levels the
// (position, mm) cell changes almost every event, lane-local accumulation does not pay and v0 stays the default):
//   * one elected thread streams the tile's event slice into shared memory with 1-D TMA bulk copies
//     (cp.async.bulk.shared.global + mbarrier complete_tx), a 3-stage ring, no register staging;
//   * every thread then walks a CONTIGUOUS run of E events of the stage (E = 20 at M = 1: the 80-byte / 20-byte thread
//     strides are bank-conflict-free for the 128-bit position loads and 32-bit base/qual loads).  Position-major order
//     means a run covers 1-2 positions, so the thread counts 4 events at a time with byte-SIMD logic + POPC into four
//     registers and touches shared memory only when the position changes (~2 atomics per 20 events);
//   * the finished tile is written out with 128-bit stores as before.
// Per-thread state of the lane-serial M = 1 walk: current position and its four base counters.
struct k1_run {
    int cur_p;
    unsigned a0, a1, a2, a3;
};

// Flush the counters of position `p` to the tile histogram when `doit`; straight-line predicated code (no
// branch/reconvergence overhead): out-of-tile positions are clamped to a dummy row at index np and reported.
// Flush the counters of position `p` to the tile histogram when `doit`.  ONE branch per flush with four unconditional
// reductions inside: ptxas turns every conditional shared-memory atomic into its own branch + BSSY/BSYNC region
// (29 % of the executed instructions in the previous version), so the per-counter "if (a_k)" tests are gone.
// Out-of-tile positions are clamped to a dummy row at index np and reported.
__device__ __forceinline__ void k1_red(uint32_t smem_addr, unsigned val)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_addr), "r"(val) : "memory");
}

__device__ __forceinline__ void k1_flush(int32_t *s_cnt, int p, int rel0, int np, bool doit, unsigned a0, unsigned a1,
                                         unsigned a2, unsigned a3, unsigned &bad)
{
    if (doit) {
        const unsigned pr = (unsigned)(p - rel0);
        bad |= ((a0 | a1 | a2 | a3) && pr >= (unsigned)np && p != 0x7fffffff) ? 1u : 0u;
        const uint32_t d = isb_smem_u32(s_cnt) + (min(pr, (unsigned)np) << 4);
        k1_red(d + 0, a0);
        k1_red(d + 4, a1);
        k1_red(d + 8, a2);
        k1_red(d + 12, a3);
    }
}

// One 4-event vector of the run.  Position-major order => a leading segment A (events equal to the first position)
// and at most one trailing segment B in the common case; both are counted with byte masks + POPC.
template <bool kInterior>
__device__ __forceinline__ void k1_vec4(k1_run &r, const int4 pv, const uint32_t b4, const uint32_t q4, uint32_t q_add,
                                        bool q_simd, int min_qual, int i, int v_lo, int v_hi, int32_t *s_cnt, int rel0,
                                        int np, int p0, unsigned long long *nmask, unsigned &bad)
{
    uint32_t ok;
    if (q_simd) ok = ((((q4 & 0x7f7f7f7fu) + q_add) | q4) >> 7) & 0x01010101u;
    else ok = ((q4 & 0xff) >= (unsigned)min_qual) | (((q4 >> 8) & 0xff) >= (unsigned)min_qual) << 8 |
              (((q4 >> 16) & 0xff) >= (unsigned)min_qual) << 16 | ((q4 >> 24) >= (unsigned)min_qual) << 24;
    if (!kInterior) {                                                // clip to the tile's slice [v_lo, v_hi)
        const int lo_k = min(max(v_lo - i, 0), 4), hi_k = min(max(v_hi - i, 0), 4);
        const uint32_t m_lo = lo_k >= 4 ? 0u : (0xffffffffu << (8 * lo_k));
        const uint32_t m_hi = hi_k >= 4 ? 0xffffffffu : ~(0xffffffffu << (8 * hi_k));
        ok &= m_lo & m_hi;
    }
    const int32_t ps[4] = {pv.x, pv.y, pv.z, pv.w};
    if (b4 & 0xfcfcfcfcu) {                                          // non-ACGT base(s): rare, exact per-event path
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (((b4 >> (8 * j)) & 0xfc) && ((ok >> (8 * j)) & 1)) {
                const unsigned pr = (unsigned)(ps[j] - rel0);
                if (pr < (unsigned)np) { if (nmask) atomicOr(nmask + p0 + pr, 1ull); }
                else bad |= 1u;
                ok &= ~(1u << (8 * j));
            }
    }
    const uint32_t lo = b4 & 0x01010101u, hi = (b4 >> 1) & 0x01010101u;
    const int same = (pv.y == pv.x) + (pv.z == pv.x) + (pv.w == pv.x);
    const bool multi = (same == 0 && pv.y != pv.w) || (same == 1 && pv.z != pv.w);     // >= 3 positions in 4 events
    if (!multi) {
        const uint32_t mA = 0x01010101u >> (8 * (3 - same));
        const uint32_t okA = ok & mA, okB = ok & ~mA;
        const bool f1 = pv.x != r.cur_p;                               // the run's position ends before this vector
        k1_flush(s_cnt, r.cur_p, rel0, np, f1, r.a0, r.a1, r.a2, r.a3, bad);
        const unsigned keep = f1 ? 0u : 0xffffffffu;
        r.a0 = (r.a0 & keep) + __popc(okA & ~hi & ~lo);
        r.a1 = (r.a1 & keep) + __popc(okA & ~hi & lo);
        r.a2 = (r.a2 & keep) + __popc(okA & hi & ~lo);
        r.a3 = (r.a3 & keep) + __popc(okA & hi & lo);
        const bool f2 = same != 3;                                     // a second position starts inside the vector
        k1_flush(s_cnt, pv.x, rel0, np, f2, r.a0, r.a1, r.a2, r.a3, bad);
        const unsigned b0 = __popc(okB & ~hi & ~lo), b1 = __popc(okB & ~hi & lo);
        const unsigned b2 = __popc(okB & hi & ~lo), b3 = __popc(okB & hi & lo);
        r.a0 = f2 ? b0 : r.a0;
        r.a1 = f2 ? b1 : r.a1;
        r.a2 = f2 ? b2 : r.a2;
        r.a3 = f2 ? b3 : r.a3;
        r.cur_p = pv.w;
    } else {                                                           // coverage < ~3: event by event
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool f = ps[j] != r.cur_p;
            k1_flush(s_cnt, r.cur_p, rel0, np, f, r.a0, r.a1, r.a2, r.a3, bad);
            const unsigned keep = f ? 0u : 0xffffffffu;
            const unsigned one = (ok >> (8 * j)) & 1u;
            const int b = (b4 >> (8 * j)) & 3;
            r.a0 = (r.a0 & keep) + (b == 0 ? one : 0u);
            r.a1 = (r.a1 & keep) + (b == 1 ? one : 0u);
            r.a2 = (r.a2 & keep) + (b == 2 ? one : 0u);
            r.a3 = (r.a3 & keep) + (b == 3 ? one : 0u);
            r.cur_p = ps[j];
        }
    }
}

// kE events per thread per stage (kE*4 and kE bytes thread strides must be bank-conflict-free: kE = 12, 20, 28 ...),
// kT threads, kS stages.
template <bool kM1, int kT, int kE, int kS>
__global__ void __launch_bounds__(kT)
k1_pileup_tiles_tma(const int32_t *__restrict__ ref_pos, const uint8_t *__restrict__ base,
                    const uint8_t *__restrict__ qual, const int32_t *__restrict__ read_id,
                    const uint8_t *__restrict__ pair_mm, const int64_t *__restrict__ tile_off, int64_t n,
                    int32_t start, int32_t L, int M, int TP, int min_qual, int32_t *__restrict__ counts,
                    unsigned long long *__restrict__ nmask, unsigned int *__restrict__ d_err)
{
    constexpr int CH = kE * kT;                                      // events per stage (multiple of 16)
    constexpr int STAGE_BYTES = CH * (kM1 ? 6 : 10);
    static_assert(CH % 16 == 0 && STAGE_BYTES % 128 == 0, "stage geometry");
    extern __shared__ __align__(128) unsigned char s_raw[];
    // layout: [stages][ pos CH*4 | rid CH*4 (M>1) | base CH | qual CH ] | counters (TP+1)*M*16 | mbarriers
    unsigned char *s_stage = s_raw;
    int32_t *s_cnt = reinterpret_cast<int32_t *>(s_raw + kS * STAGE_BYTES);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_raw + kS * STAGE_BYTES + (size_t)(TP + 1) * M * 16);

    const int tile = blockIdx.x;
    const int p0 = tile * TP;
    const int np = min(TP, L - p0);
    const int n_cnt4 = np * M;
    const int tid = threadIdx.x;
    for (int i = tid; i < n_cnt4 + M; i += kT) reinterpret_cast<int4 *>(s_cnt)[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < kS; ++s) isb_mbar_init(s_bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t e_lo = tile_off[tile], e_hi = tile_off[tile + 1];
    const int64_t A = e_lo & ~(int64_t)15;                           // slice start, 16-event aligned for the bulk copies
    const int64_t n_bulk = n & ~(int64_t)15;                         // events that may be fetched with 16-byte granules
    const int n_chunks = (int)((e_hi - A + CH - 1) / CH);
    const int32_t rel0 = start + p0;
    unsigned bad = 0;

    auto issue = [&](int c) {                                        // thread 0 only
        const int st = c % kS;
        const int64_t c_lo = A + (int64_t)c * CH;
        int64_t c_hi = min(c_lo + CH, (e_hi + 15) & ~(int64_t)15);
        c_hi = min(c_hi, n_bulk);
        const unsigned ne = c_hi > c_lo ? (unsigned)(c_hi - c_lo) : 0u;
        unsigned char *sp = s_stage + (size_t)st * STAGE_BYTES;
        isb_mbar_expect_tx(s_bar + st, ne * (kM1 ? 6u : 10u));
        if (ne) {
            isb_bulk_g2s(sp, ref_pos + c_lo, ne * 4u, s_bar + st);
            if (!kM1) isb_bulk_g2s(sp + CH * 4, read_id + c_lo, ne * 4u, s_bar + st);
            isb_bulk_g2s(sp + CH * (kM1 ? 4 : 8), base + c_lo, ne, s_bar + st);
            isb_bulk_g2s(sp + CH * (kM1 ? 5 : 9), qual + c_lo, ne, s_bar + st);
        }
    };
    if (tid == 0)
        for (int c = 0; c < kS - 1 && c < n_chunks; ++c) issue(c);

    // byte-SIMD quality test: high bit of byte j set iff q_j >= min_qual (valid for 1 <= min_qual <= 128)
    const uint32_t q_add = 0x80808080u - 0x01010101u * (uint32_t)min(max(min_qual, 1), 128);
    const bool q_simd = min_qual >= 1 && min_qual <= 128;

    for (int c = 0; c < n_chunks; ++c) {
        const int st = c % kS;
        if (tid == 0 && c + kS - 1 < n_chunks) issue(c + kS - 1);
        unsigned char *sp = s_stage + (size_t)st * STAGE_BYTES;
        const int32_t *s_pos = reinterpret_cast<const int32_t *>(sp);
        const int32_t *s_rid = reinterpret_cast<const int32_t *>(sp + CH * 4);
        const unsigned char *s_base = sp + CH * (kM1 ? 4 : 8);
        const unsigned char *s_qual = sp + CH * (kM1 ? 5 : 9);
        const int64_t c_lo = A + (int64_t)c * CH;
        // the (< 16) events past the last 16-aligned boundary of the whole array cannot be bulk-copied: plain copies
        if (c_lo + CH > n_bulk && n_bulk < n) {
            const int64_t t_lo = max(c_lo, n_bulk), t_hi = min(c_lo + CH, n);
            for (int64_t e = t_lo + tid; e < t_hi; e += kT) {
                const int i = (int)(e - c_lo);
                const_cast<int32_t *>(s_pos)[i] = ref_pos[e];
                if (!kM1) const_cast<int32_t *>(s_rid)[i] = read_id[e];
                const_cast<unsigned char *>(s_base)[i] = base[e];
                const_cast<unsigned char *>(s_qual)[i] = qual[e];
            }
            __syncthreads();
        }
        isb_mbar_wait(s_bar + st, (unsigned)((c / kS) & 1));

        // Run -> thread mapping.  Lanes of one warp take runs that are kT/32 (an ODD number of) runs apart: neighbouring
        // lanes then sit on different positions, so the flush atomics of one warp instruction do not collide on an
        // address (49.8 M bank conflicts with the contiguous mapping), and the odd stride keeps the 128-bit position
        // loads and 32-bit base/qual loads bank-conflict-free.
        constexpr int kW = kT / 32;
        static_assert(kW % 2 == 1, "interleaved mapping needs an odd warp count");
        const int i0 = ((tid & 31) * kW + (tid >> 5)) * kE;
        // valid events of the stage: chunk-relative index in [v_lo, v_hi)
        const int v_lo = (int)max((int64_t)0, e_lo - c_lo), v_hi = (int)min((int64_t)CH, e_hi - c_lo);
        if (kM1) {
            constexpr int NV = kE / 4;
            int4 pv[NV];
            uint32_t b4[NV], q4[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {                           // all loads of the run first: hides the LDS latency
                pv[v] = *reinterpret_cast<const int4 *>(s_pos + i0 + 4 * v);
                b4[v] = *reinterpret_cast<const uint32_t *>(s_base + i0 + 4 * v);
                q4[v] = *reinterpret_cast<const uint32_t *>(s_qual + i0 + 4 * v);
            }
            k1_run r = {0x7fffffff, 0u, 0u, 0u, 0u};
            if (v_lo == 0 && v_hi == CH) {                           // interior chunk (all but the tile's two ends)
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    k1_vec4<true>(r, pv[v], b4[v], q4[v], q_add, q_simd, min_qual, i0 + 4 * v, v_lo, v_hi, s_cnt, rel0, np,
                                  p0, nmask, bad);
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int i = i0 + 4 * v;
                    if (i + 4 <= v_lo || i >= v_hi) continue;
                    k1_vec4<false>(r, pv[v], b4[v], q4[v], q_add, q_simd, min_qual, i, v_lo, v_hi, s_cnt, rel0, np, p0,
                                   nmask, bad);
                }
            }
            k1_flush(s_cnt, r.cur_p, rel0, np, true, r.a0, r.a1, r.a2, r.a3, bad);
        } else {
            // M > 1: the (position, mm) cell changes almost every event, so each event goes straight to the tile
            // histogram with ONE unconditional shared-memory reduction (value 0 when the event does not count; clamped
            // index) -- no per-event branches.  The interleaved run mapping keeps the 32 addresses of a warp
            // instruction on different positions.
            constexpr int NV = kE / 4;
            int4 pv[NV], rv[NV];
            uint32_t b4[NV], q4[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                pv[v] = *reinterpret_cast<const int4 *>(s_pos + i0 + 4 * v);
                rv[v] = *reinterpret_cast<const int4 *>(s_rid + i0 + 4 * v);
                b4[v] = *reinterpret_cast<const uint32_t *>(s_base + i0 + 4 * v);
                q4[v] = *reinterpret_cast<const uint32_t *>(s_qual + i0 + 4 * v);
            }
            const uint32_t s_cnt_u32 = isb_smem_u32(s_cnt);
            const bool interior = v_lo == 0 && v_hi == CH;
            // quality mask of each vector (byte-SIMD), clipped to the tile's slice on boundary chunks; then ALL mm gathers
            // of the run are issued before any of them is used
            uint32_t okm[NV];
            int mmv[NV][4];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int i = i0 + 4 * v;
                uint32_t ok;
                if (q_simd) ok = ((((q4[v] & 0x7f7f7f7fu) + q_add) | q4[v]) >> 7) & 0x01010101u;
                else ok = ((q4[v] & 0xff) >= (unsigned)min_qual) | (((q4[v] >> 8) & 0xff) >= (unsigned)min_qual) << 8 |
                          (((q4[v] >> 16) & 0xff) >= (unsigned)min_qual) << 16 | ((q4[v] >> 24) >= (unsigned)min_qual) << 24;
                if (!interior) {
                    const int lo_k = min(max(v_lo - i, 0), 4), hi_k = min(max(v_hi - i, 0), 4);
                    const uint32_t m_lo = lo_k >= 4 ? 0u : (0xffffffffu << (8 * lo_k));
                    const uint32_t m_hi = hi_k >= 4 ? 0xffffffffu : ~(0xffffffffu << (8 * hi_k));
                    ok &= m_lo & m_hi;
                }
                okm[v] = ok;
                const int32_t rs[4] = {rv[v].x, rv[v].y, rv[v].z, rv[v].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) mmv[v][j] = ((ok >> (8 * j)) & 1u) ? (int)__ldg(pair_mm + rs[j]) : 0;
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int32_t ps[4] = {pv[v].x, pv[v].y, pv[v].z, pv[v].w};
                uint32_t ok = okm[v];
                if (b4[v] & 0xfcfcfcfcu) {                             // non-ACGT base(s): rare, exact per-event path
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (((b4[v] >> (8 * j)) & 0xfc) && ((ok >> (8 * j)) & 1)) {
                            const unsigned pr = (unsigned)(ps[j] - rel0);
                            if (pr < (unsigned)np && mmv[v][j] < M) { if (nmask) atomicOr(nmask + p0 + pr, 1ull << mmv[v][j]); }
                            else bad |= pr < (unsigned)np ? 2u : 1u;
                            ok &= ~(1u << (8 * j));
                        }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned one = (ok >> (8 * j)) & 1u;
                    const unsigned pr = (unsigned)(ps[j] - rel0);
                    const unsigned oob = pr >= (unsigned)np ? 1u : 0u, mmb = mmv[v][j] >= M ? 1u : 0u;
                    bad |= (0u - one) & (oob | (mmb << 1));
                    const unsigned cell = min(pr, (unsigned)np) * (unsigned)M + (unsigned)min(mmv[v][j], M - 1);
                    const unsigned b = (b4[v] >> (8 * j)) & 3u;
                    k1_red(s_cnt_u32 + ((cell << 4) | (b << 2)), one & (oob ^ 1u) & (mmb ^ 1u));
                }
            }
        }
        __syncthreads();                                               // stage consumed: thread 0 may refill it
    }
    if (bad & 1u) atomicOr(d_err, ISB_DEV_ERR_ORDER);
    if (bad & 2u) atomicOr(d_err, ISB_DEV_ERR_MM);
    int4 *dst = reinterpret_cast<int4 *>(counts) + (size_t)p0 * M;
    for (int i = tid; i < n_cnt4; i += kT) dst[i] = reinterpret_cast<const int4 *>(s_cnt)[i];
}

template <bool kM1, int kT, int kE, int kS>
static int k1_launch_tma(isb_ctx *ctx, int n_tiles, size_t cnt_bytes, int M, const int32_t *ref_pos, const uint8_t *base,
                         const uint8_t *qual, const int32_t *read_id, const uint8_t *pair_mm, const int64_t *tile_off,
                         int64_t n, int32_t start, int32_t L, int TP, int min_qual, int32_t *counts,
                         unsigned long long *nmask)
{
    const size_t smem = (size_t)kS * kE * kT * (kM1 ? 6 : 10) + cnt_bytes + (size_t)M * 16 + kS * sizeof(uint64_t);
    auto kern = k1_pileup_tiles_tma<kM1, kT, kE, kS>;
    ISB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, kT, smem, ctx->stream>>>(ref_pos, base, qual, read_id, pair_mm, tile_off, n, start, L, M, TP, min_qual,
                                            counts, nmask, ctx->d_err);
    return ISB_OK;
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

// event offsets of tiles of `tp` positions into ctx->buf[SL_K3_TILE_OFF] (used by K3 to bound its per-site searches)
int isb_tile_offsets(isb_ctx *ctx, const int32_t *ref_pos, int64_t n, int32_t start, int32_t L, int tp, int n_tiles)
{
    int rc = isb_ensure(ctx, SL_K3_TILE_OFF, sizeof(int64_t) * ((size_t)n_tiles + 1));
    if (rc) return rc;
    k1_tile_offsets<<<(n_tiles + 1 + 255) / 256, 256, 0, ctx->stream>>>(ref_pos, n, start, L, tp, n_tiles,
                                                                      (int64_t *)ctx->buf[SL_K3_TILE_OFF].p);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
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
    static int variant = -1, cfg = 0;               // ISB_K1_VARIANT=0: warp-cooperative kernel (A/B + fallback)
    if (variant < 0) {
        const char *v = getenv("ISB_K1_VARIANT");
        variant = v ? atoi(v) : 2;
        const char *c = getenv("ISB_K1_CFG");           // tile-kernel geometry, see k1_launch_tma instantiations below
        cfg = c ? atoi(c) : 0;
    }
    const bool aligned16 = ((((uintptr_t)ref_pos | (uintptr_t)base | (uintptr_t)qual) & 15) == 0) &&
                           (M == 1 || (((uintptr_t)read_id) & 15) == 0);
#define K1_ARGS ref_pos, base, qual, read_id, pair_mm, tile_off, n, start, L, M, TP, min_qual, counts, nmask, ctx->d_err
#define K1_TMA_ARGS ctx, n_tiles, smem, M, ref_pos, base, qual, read_id, pair_mm, tile_off, n, start, L, TP, min_qual, counts, nmask
    if (variant == 0 || !aligned16 || (M > 1 && variant == 1)) {
        if (M == 1) k1_pileup_tiles<true><<<n_tiles, K1_THREADS, smem, st>>>(K1_ARGS);
        else k1_pileup_tiles<false><<<n_tiles, K1_THREADS, smem, st>>>(K1_ARGS);
    } else if (M == 1) {
        int rc2;
        // measured on B200 (2e8 events, c=100): 352x12 0.299 ms, 416x12 0.311, 288x12 0.314, 224x20 0.326, 160x28 0.408
        if (cfg == 1) rc2 = k1_launch_tma<true, 416, 12, 3>(K1_TMA_ARGS);        // 2 CTAs x 13 warps / SM
        else if (cfg == 2) rc2 = k1_launch_tma<true, 224, 20, 3>(K1_TMA_ARGS);   // 2 CTAs x 7 warps / SM
        else if (cfg == 3) rc2 = k1_launch_tma<true, 160, 28, 3>(K1_TMA_ARGS);   // 2 CTAs x 5 warps / SM
        else if (cfg == 4) rc2 = k1_launch_tma<true, 288, 12, 3>(K1_TMA_ARGS);   // 2 CTAs x 9 warps / SM
        else rc2 = k1_launch_tma<true, 352, 12, 3>(K1_TMA_ARGS);                 // default: 2 CTAs x 11 warps / SM
        if (rc2) return rc2;
    } else {
        int rc2 = k1_launch_tma<false, 224, 12, 3>(K1_TMA_ARGS);                 // 2 CTAs x 7 warps / SM at M ~ 15
        if (rc2) return rc2;
    }
#undef K1_ARGS
#undef K1_TMA_ARGS
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
