// isb_synth.cu -- libisb_synth.so: device-side synthetic metagenome -> position-major event columns and/or read-major segments.
//
// BENCH / TEST SUPPORT, not part of the hot path and not part of libinstrain_b200.so.  BASELINE.json's headline
// workload (100 scaffolds x 1 Mb at 100x = 1e10 aligned bases = 100 GB of event columns) cannot be generated on the
// host or shipped over PCIe in bench time, so the data set is synthesised directly in HBM, deterministically from a
// seed, following the model of SURVEY.md section 8d (same model as the CPU generator the parity tests use, different RNG):
//   reference iid uniform; K=4 haplotypes (0.4,0.3,0.2,0.1); Bernoulli(density) SNV sites, alt base uniform among
//   the 3 others, carried by a random non-empty proper subset of haplotypes; 2x150 read pairs, fragment length
//   ~N(350,30) clipped to [200,500] (Irwin-Hall(12) normal), uniform starts; base qualities from the bundled BAM's
//   bins; substitution errors with p = 10^(-q/10); pair mm = mismatches vs reference, pairs with mm >= 15 dropped
//   (1 - mm/300 <= 0.95, inStrain/filter_reads.py:406-408); htslib's mate-overlap quality tweak applied (all-M reads).
// Events are emitted POSITION-MAJOR directly (one thread per position enumerates the fragments covering it), pair
// ids follow fragment-start order (= BAM order of the first mate).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../csrc/isb_scan.cuh"

#define READLEN 150
#define FRAG_MIN 200
#define FRAG_MAX 500
#define MM_DROP 15

struct isbs_params {
    int64_t L;            // positions per scaffold
    int32_t n_scaffolds;
    int32_t coverage;
    double snv_density;
    uint64_t seed;
    int32_t skip_mm;      // 1: pair_mm = 0 for all kept pairs (--skip_mm_profiling)
    int32_t pad;
};

struct isbs_state {
    isbs_params prm;
    int K;                // fragment slots per start position
    uint32_t p_occ;       // slot occupancy threshold (u32)
    int64_t Ltot, n_slots;
    int32_t *slot_id;     // kept-pair id or -1
    uint16_t *slot_F;
    uint8_t *slot_mm;     // mm (255 = unoccupied)
    int64_t *ev_off;      // per-position event offset
    int32_t *cov;
    int64_t *scan_tmp;
    unsigned long long *d_tot;
    int64_t n_events, n_pairs;
    char err[256];
};

static isbs_state g;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rng(uint64_t seed, uint64_t tag, uint64_t i, uint64_t j)
{
    return mix64(mix64(seed + tag * 0x632be59bd9b4e019ull) ^ (i * 0xd1b54a32d192ed03ull) ^ (j * 0x8cb92ba72f3d8dd7ull));
}
enum { TAG_POS = 1, TAG_FRAG = 2, TAG_EV = 3 };

struct pos_info { uint8_t ref, alt, carriers, is_snv; };

__device__ __forceinline__ pos_info pos_draw(uint64_t seed, int64_t p, uint32_t dens24)
{
    const uint64_t r = rng(seed, TAG_POS, (uint64_t)p, 0);
    pos_info o;
    o.ref = r & 3;
    o.is_snv = ((r >> 8) & 0xffffff) < dens24;
    o.alt = (o.ref + 1 + ((r >> 32) & 0xff) % 3) & 3;
    o.carriers = 1 + ((r >> 40) & 0xffff) % 14;
    return o;
}
__device__ __forceinline__ int hap_base(const pos_info &pi, int h) { return (pi.is_snv && ((pi.carriers >> h) & 1)) ? pi.alt : pi.ref; }

// quality bins of the bundled BAM and their error probabilities 10^(-q/10), as 16/24-bit thresholds
__constant__ uint8_t c_qbin[7] = {8, 12, 22, 27, 32, 37, 41};
__constant__ uint32_t c_qcum[7] = {131, 3080, 5308, 8389, 13959, 24707, 65536};     // cumulative p * 65536
__constant__ uint32_t c_perr[7] = {2658984, 1058559, 105856, 33475, 10586, 3347, 1333};   // p_err * 2^24

__device__ __forceinline__ void ev_draw(uint64_t seed, int64_t slot, int mate, int off, int true_base, int &b, int &q)
{
    const uint64_t r = rng(seed, TAG_EV, (uint64_t)slot * 2 + mate, (uint64_t)off);
    const uint32_t u = r & 0xffff;
    int bin = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) bin += (u >= c_qcum[i]);
    q = c_qbin[bin];
    const bool err = ((r >> 16) & 0xffffff) < c_perr[bin];
    b = err ? ((true_base + 1 + ((r >> 40) & 0xff) % 3) & 3) : true_base;
}

__device__ __forceinline__ bool frag_draw(uint64_t seed, int64_t slot, uint32_t p_occ, int &F, int &hap)
{
    const uint64_t r1 = rng(seed, TAG_FRAG, (uint64_t)slot, 0);
    if ((uint32_t)r1 >= p_occ) return false;
    const uint32_t hu = (r1 >> 32) & 0xffff;
    hap = (hu >= 26214) + (hu >= 45875) + (hu >= 58982);          // 0.4, 0.7, 0.9
    uint32_t s = 0;
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const uint64_t r = rng(seed, TAG_FRAG, (uint64_t)slot, k);
        s += (r & 0xffff) + ((r >> 16) & 0xffff) + ((r >> 32) & 0xffff) + ((r >> 48) & 0xffff);
    }
    int f = 350 + (int)((30ll * ((long long)s - 6 * 65536)) >> 16);
    F = min(max(f, FRAG_MIN), FRAG_MAX);
    return true;
}

// one thread per fragment slot: occupancy, fragment length, pair mm, keep flag
__global__ void synth_slots(isbs_params prm, int K, uint32_t p_occ, uint32_t dens24, int64_t n_slots,
                            uint16_t *__restrict__ slot_F, uint8_t *__restrict__ slot_mm)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int64_t xg = s / K;                      // global start position
    const int64_t x = xg % prm.L;                  // within scaffold
    int F, hap;
    slot_F[s] = 0;
    slot_mm[s] = 255;
    if (!frag_draw(prm.seed, s, p_occ, F, hap)) return;
    if (x + F > prm.L) return;                     // fragment would leave the scaffold: no fragment
    int mm = 0;
    for (int mate = 0; mate < 2; ++mate) {
        const int64_t r0 = xg + (mate ? F - READLEN : 0);
        for (int o = 0; o < READLEN; ++o) {
            const pos_info pi = pos_draw(prm.seed, r0 + o, dens24);
            int b, q;
            ev_draw(prm.seed, s, mate, o, hap_base(pi, hap), b, q);
            mm += (b != pi.ref);
        }
    }
    slot_F[s] = (uint16_t)F;
    slot_mm[s] = (uint8_t)min(mm, 254);
}

struct KeepFn {
    const uint8_t *mm;
    __device__ int operator()(int64_t i) const { return mm[i] < MM_DROP ? 1 : 0; }
};
struct SlotIdSink {
    int32_t *slot_id;
    __device__ void operator()(int64_t i, int64_t prefix, int v) const { slot_id[i] = v ? (int32_t)prefix : -1; }
};
struct CovFn {
    const int32_t *cov;
    __device__ int operator()(int64_t i) const { return cov[i]; }
};
struct OffSink {
    int64_t *off;
    __device__ void operator()(int64_t i, int64_t prefix, int) const { off[i] = prefix; }
};

__global__ void synth_pair_mm(const int32_t *__restrict__ slot_id, const uint8_t *__restrict__ slot_mm, int64_t n_slots,
                              int skip_mm, uint8_t *__restrict__ pair_mm)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int32_t id = slot_id[s];
    if (id >= 0) pair_mm[id] = skip_mm ? 0 : slot_mm[s];
}

// one thread per position: enumerate covering fragments in (start, slot, mate) order.
// kFill = false: count events; kFill = true: write them at ev_off[p].
template <bool kFill>
__global__ void synth_events(isbs_params prm, int K, uint32_t dens24, int64_t Ltot,
                             const int32_t *__restrict__ slot_id, const uint16_t *__restrict__ slot_F,
                             const int64_t *__restrict__ ev_off, int32_t *__restrict__ cov,
                             int32_t *__restrict__ ref_pos, uint8_t *__restrict__ base, uint8_t *__restrict__ qual,
                             int32_t *__restrict__ read_id, uint8_t *__restrict__ ref)
{
    const int64_t pg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pg >= Ltot) return;
    const int64_t sc0 = (pg / prm.L) * prm.L;      // first position of this scaffold
    const pos_info pi = kFill ? pos_draw(prm.seed, pg, dens24) : pos_info();
    if (kFill) ref[pg] = pi.ref;
    int64_t w = kFill ? ev_off[pg] : 0;
    int n = 0;
    const int64_t x_lo = max(sc0, pg - (FRAG_MAX - 1));
    for (int64_t xg = x_lo; xg <= pg; ++xg) {
        for (int k = 0; k < K; ++k) {
            const int64_t s = xg * K + k;
            const int32_t id = slot_id[s];
            if (id < 0) continue;
            const int F = slot_F[s];
            const int d = (int)(pg - xg);
            const bool in1 = d < READLEN;
            const bool in2 = d >= F - READLEN && d < F;
            if (!(in1 || in2)) continue;
            if (!kFill) { n += (int)in1 + (int)in2; continue; }
            const uint64_t r1 = rng(prm.seed, TAG_FRAG, (uint64_t)s, 0);
            const uint32_t hu = (r1 >> 32) & 0xffff;
            const int hap = (hu >= 26214) + (hu >= 45875) + (hu >= 58982);
            const int tb = hap_base(pi, hap);
            int b1 = 0, q1 = 0, b2 = 0, q2 = 0;
            if (in1) ev_draw(prm.seed, s, 0, d, tb, b1, q1);
            if (in2) ev_draw(prm.seed, s, 1, d - (F - READLEN), tb, b2, q2);
            if (in1 && in2) {                      // htslib tweak_overlap_quality, a = mate 1
                if (b1 == b2) { q1 = min(200, q1 + q2); q2 = 0; }
                else if (q1 >= q2) { q1 = (int)(0.8 * q1); q2 = 0; }
                else { q2 = (int)(0.8 * q2); q1 = 0; }
            }
            if (in1) { ref_pos[w] = (int32_t)pg; base[w] = (uint8_t)b1; qual[w] = (uint8_t)q1; read_id[w] = id; ++w; }
            if (in2) { ref_pos[w] = (int32_t)pg; base[w] = (uint8_t)b2; qual[w] = (uint8_t)q2; read_id[w] = id; ++w; }
        }
    }
    if (!kFill) cov[pg] = n;
}

// ---- read-major output: the same fragments as aligned segments (one per mate), sorted by start ------------------------
// one thread per position: the segments STARTING there (mate 1 of the fragments starting at pg, then mate 2 of the
// fragments whose second mate starts at pg).  kFill = false counts, kFill = true writes the segment table.
#define SEG_DATA_WORDS ((READLEN + 14) / 8)        // position-aligned words of a READLEN segment at the worst offset
static int g_seg_words = SEG_DATA_WORDS + 1;       // block size: data words (one may stay empty) + 1 (or 2) zero separator words
template <bool kFill>
__global__ void synth_seg_table(isbs_params prm, int K, uint32_t dens24, int64_t Ltot, const int32_t *__restrict__ slot_id,
                                const uint16_t *__restrict__ slot_F, const int64_t *__restrict__ seg_off,
                                int32_t *__restrict__ cnt, int32_t *__restrict__ seg_start, uint16_t *__restrict__ seg_len,
                                int32_t *__restrict__ seg_pair, int64_t *__restrict__ seg_word, int64_t *__restrict__ seg_src,
                                uint8_t *__restrict__ ref, int SEG_WORDS)
{
    const int64_t pg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pg >= Ltot) return;
    const int64_t sc0 = (pg / prm.L) * prm.L;
    int64_t w = kFill ? seg_off[pg] : 0;
    int n = 0;
    for (int k = 0; k < K; ++k) {                                      // mate 1
        const int64_t s = pg * K + k;
        const int32_t id = slot_id[s];
        if (id < 0) continue;
        if (kFill) {
            seg_start[w] = (int32_t)pg; seg_len[w] = READLEN; seg_pair[w] = id; seg_word[w] = 1 + w * SEG_WORDS;
            seg_src[w] = s * 2;
            ++w;
        }
        ++n;
    }
    const int64_t x_lo = max(sc0, pg - (FRAG_MAX - READLEN));
    for (int64_t xg = x_lo; xg <= pg - (FRAG_MIN - READLEN); ++xg) {    // mate 2 starts at xg + F - READLEN
        for (int k = 0; k < K; ++k) {
            const int64_t s = xg * K + k;
            const int32_t id = slot_id[s];
            if (id < 0 || (int64_t)slot_F[s] - READLEN != pg - xg) continue;
            if (kFill) {
                seg_start[w] = (int32_t)pg; seg_len[w] = READLEN; seg_pair[w] = id; seg_word[w] = 1 + w * SEG_WORDS;
                seg_src[w] = s * 2 + 1;
                ++w;
            }
            ++n;
        }
    }
    if (!kFill) cnt[pg] = n;
    else ref[pg] = pos_draw(prm.seed, pg, dens24).ref;
}

// one thread per segment: its READLEN one-hot codes (A=1,C=2,T=4,G=8; 0 = fails min_qual after the overlap tweak)
__global__ void synth_seg_words(isbs_params prm, uint32_t dens24, int64_t n_segs, const int32_t *__restrict__ seg_start,
                                const int64_t *__restrict__ seg_src, const uint16_t *__restrict__ slot_F, int min_qual,
                                uint32_t *__restrict__ words, int SEG_WORDS)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_segs) return;
    const int64_t s = seg_src[i] >> 1;
    const int mate = (int)(seg_src[i] & 1);
    const int F = slot_F[s];
    const int64_t start = seg_start[i];
    const int64_t xg = mate ? start - (F - READLEN) : start;
    const uint64_t r1 = rng(prm.seed, TAG_FRAG, (uint64_t)s, 0);
    const uint32_t hu = (r1 >> 32) & 0xffff;
    const int hap = (hu >= 26214) + (hu >= 45875) + (hu >= 58982);
    uint32_t *dst = words + 1 + i * SEG_WORDS;
    const int sh = (int)(start & 7);               // position-aligned stream: base o sits at nibble o + sh
    for (int wj = 0; wj < SEG_DATA_WORDS; ++wj) {
        uint32_t word = 0;
        for (int nb = 0; nb < 8; ++nb) {
            const int o = wj * 8 + nb - sh;
            if (o < 0) continue;
            if (o >= READLEN) break;
            const int64_t pg = start + o;
            const int d = (int)(pg - xg);
            const pos_info pi = pos_draw(prm.seed, pg, dens24);
            const int tb = hap_base(pi, hap);
            int b1 = 0, q1 = 0, b2 = 0, q2 = 0;
            const bool in1 = d < READLEN, in2 = d >= F - READLEN && d < F;
            if (in1) ev_draw(prm.seed, s, 0, d, tb, b1, q1);
            if (in2) ev_draw(prm.seed, s, 1, d - (F - READLEN), tb, b2, q2);
            if (in1 && in2) {                      // htslib tweak_overlap_quality, a = mate 1
                if (b1 == b2) { q1 = min(200, q1 + q2); q2 = 0; }
                else if (q1 >= q2) { q1 = (int)(0.8 * q1); q2 = 0; }
                else { q2 = (int)(0.8 * q2); q1 = 0; }
            }
            const int b = mate ? b2 : b1, q = mate ? q2 : q1;
            if (q >= min_qual) word |= (1u << b) << (4 * nb);
        }
        dst[wj] = word;                            // the separator words stay zero (buffer is pre-zeroed)
    }
}

#define SYN_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            snprintf(g.err, sizeof(g.err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return -1;                                                                                   \
        }                                                                                                \
    } while (0)

template <class F, class Sink>
static int run_scan(F f, Sink sink, int64_t n, int64_t *tmp, unsigned long long *d_tot, int64_t *total)
{
    const int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    scan_reduce<<<nb, SCAN_THREADS>>>(f, n, tmp);
    scan_blocksums<<<1, 1024>>>(tmp, nb, d_tot);
    scan_scatter<<<nb, SCAN_THREADS>>>(f, n, tmp, sink);
    unsigned long long t = 0;
    SYN_CUDA(cudaMemcpy(&t, d_tot, sizeof(t), cudaMemcpyDeviceToHost));
    *total = (int64_t)t;
    return 0;
}

extern "C" {

const char *isbs_last_error(void) { return g.err; }

static void free_state()
{
    cudaFree(g.slot_id); cudaFree(g.slot_F); cudaFree(g.slot_mm); cudaFree(g.ev_off); cudaFree(g.cov);
    cudaFree(g.scan_tmp); cudaFree(g.d_tot);
    g.slot_id = nullptr; g.slot_F = nullptr; g.slot_mm = nullptr; g.ev_off = nullptr; g.cov = nullptr;
    g.scan_tmp = nullptr; g.d_tot = nullptr;
}

// Phase 1: draw fragments, compute sizes.  Returns 0 and fills n_events / n_pairs.
int isbs_plan(int device, const isbs_params *prm, int64_t *n_events, int64_t *n_pairs)
{
    memset(&g, 0, sizeof(g));
    g.prm = *prm;
    SYN_CUDA(cudaSetDevice(device));
    const double rate = (double)prm->coverage / (2.0 * READLEN);          // fragments per start position
    g.K = (int)(rate * 2.0) + 1;
    g.p_occ = (uint32_t)((rate / g.K) * 4294967296.0);
    g.Ltot = prm->L * prm->n_scaffolds;
    if (g.Ltot >= (1ll << 31)) { snprintf(g.err, sizeof(g.err), "batch coordinate space exceeds int32"); return -1; }
    g.n_slots = g.Ltot * g.K;
    const uint32_t dens24 = (uint32_t)(prm->snv_density * 16777216.0);
    SYN_CUDA(cudaMalloc(&g.slot_id, sizeof(int32_t) * g.n_slots));
    SYN_CUDA(cudaMalloc(&g.slot_F, sizeof(uint16_t) * g.n_slots));
    SYN_CUDA(cudaMalloc(&g.slot_mm, g.n_slots));
    SYN_CUDA(cudaMalloc(&g.ev_off, sizeof(int64_t) * g.Ltot));
    SYN_CUDA(cudaMalloc(&g.cov, sizeof(int32_t) * g.Ltot));
    const int64_t n_max = g.n_slots > g.Ltot ? g.n_slots : g.Ltot;
    SYN_CUDA(cudaMalloc(&g.scan_tmp, sizeof(int64_t) * ((n_max + SCAN_BLOCK - 1) / SCAN_BLOCK + 1)));
    SYN_CUDA(cudaMalloc(&g.d_tot, sizeof(unsigned long long)));
    synth_slots<<<(unsigned)((g.n_slots + 255) / 256), 256>>>(g.prm, g.K, g.p_occ, dens24, g.n_slots, g.slot_F, g.slot_mm);
    SYN_CUDA(cudaGetLastError());
    if (run_scan(KeepFn{g.slot_mm}, SlotIdSink{g.slot_id}, g.n_slots, g.scan_tmp, g.d_tot, &g.n_pairs)) return -1;
    synth_events<false><<<(unsigned)((g.Ltot + 127) / 128), 128>>>(g.prm, g.K, dens24, g.Ltot, g.slot_id, g.slot_F, nullptr,
                                                                   g.cov, nullptr, nullptr, nullptr, nullptr, nullptr);
    SYN_CUDA(cudaGetLastError());
    if (run_scan(CovFn{g.cov}, OffSink{g.ev_off}, g.Ltot, g.scan_tmp, g.d_tot, &g.n_events)) return -1;
    SYN_CUDA(cudaDeviceSynchronize());
    *n_events = g.n_events;
    *n_pairs = g.n_pairs;
    return 0;
}

// Phase 2 (optional, before isbs_fill): the same data set as a read-major batch.  n_segs = 2 * n_pairs segments of READLEN
// bases, n_words = isbs_reads_words(n_segs) words; all buffers caller-owned DEVICE memory.
int64_t isbs_reads_words(int64_t n_segs) { return (1 + n_segs * g_seg_words + 3) / 4 * 4; }
// words per segment block: READLEN/8 data words + 1 separator (default) or + 2 (an odd block size spreads the
// shared-memory banks of K1r's per-lane word fetches)
int isbs_set_seg_words(int w) { if (w < SEG_DATA_WORDS + 1 || w > SEG_DATA_WORDS + 2) return -1; g_seg_words = w; return 0; }

int isbs_fill_reads(int32_t *seg_start, uint16_t *seg_len, int32_t *seg_pair, int64_t *seg_word, uint32_t *words,
                    uint8_t *pair_mm, uint8_t *ref, int min_qual)
{
    const uint32_t dens24 = (uint32_t)(g.prm.snv_density * 16777216.0);
    const int64_t n_segs = 2 * g.n_pairs;
    int64_t *seg_src = nullptr, *seg_off = nullptr;
    int32_t *cnt = nullptr;
    SYN_CUDA(cudaMalloc(&seg_src, sizeof(int64_t) * (size_t)(n_segs + 1)));
    SYN_CUDA(cudaMalloc(&seg_off, sizeof(int64_t) * (size_t)g.Ltot));
    SYN_CUDA(cudaMalloc(&cnt, sizeof(int32_t) * (size_t)g.Ltot));
    synth_seg_table<false><<<(unsigned)((g.Ltot + 127) / 128), 128>>>(g.prm, g.K, dens24, g.Ltot, g.slot_id, g.slot_F, nullptr,
                                                                      cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, g_seg_words);
    SYN_CUDA(cudaGetLastError());
    int64_t total = 0;
    if (run_scan(CovFn{cnt}, OffSink{seg_off}, g.Ltot, g.scan_tmp, g.d_tot, &total)) return -1;
    if (total != n_segs) { snprintf(g.err, sizeof(g.err), "segment count mismatch %lld vs %lld", (long long)total, (long long)n_segs); return -1; }
    synth_seg_table<true><<<(unsigned)((g.Ltot + 127) / 128), 128>>>(g.prm, g.K, dens24, g.Ltot, g.slot_id, g.slot_F, seg_off,
                                                                     nullptr, seg_start, seg_len, seg_pair, seg_word, seg_src, ref, g_seg_words);
    SYN_CUDA(cudaGetLastError());
    SYN_CUDA(cudaMemset(words, 0, sizeof(uint32_t) * (size_t)isbs_reads_words(n_segs)));
    if (n_segs > 0) {
        synth_seg_words<<<(unsigned)((n_segs + 127) / 128), 128>>>(g.prm, dens24, n_segs, seg_start, seg_src, g.slot_F, min_qual, words, g_seg_words);
        SYN_CUDA(cudaGetLastError());
    }
    synth_pair_mm<<<(unsigned)((g.n_slots + 255) / 256), 256>>>(g.slot_id, g.slot_mm, g.n_slots, g.prm.skip_mm, pair_mm);
    SYN_CUDA(cudaGetLastError());
    SYN_CUDA(cudaDeviceSynchronize());
    cudaFree(seg_src); cudaFree(seg_off); cudaFree(cnt);
    return 0;
}

// Phase 3 without event columns: free the plan.
int isbs_free(void)
{
    free_state();
    return 0;
}

// Phase 2: write the columns into caller-owned DEVICE buffers (sizes from isbs_plan), then free the plan.
int isbs_fill(int32_t *ref_pos, uint8_t *base, uint8_t *qual, int32_t *read_id, uint8_t *pair_mm, uint8_t *ref)
{
    const uint32_t dens24 = (uint32_t)(g.prm.snv_density * 16777216.0);
    synth_pair_mm<<<(unsigned)((g.n_slots + 255) / 256), 256>>>(g.slot_id, g.slot_mm, g.n_slots, g.prm.skip_mm, pair_mm);
    SYN_CUDA(cudaGetLastError());
    synth_events<true><<<(unsigned)((g.Ltot + 127) / 128), 128>>>(g.prm, g.K, dens24, g.Ltot, g.slot_id, g.slot_F, g.ev_off,
                                                                  nullptr, ref_pos, base, qual, read_id, ref);
    SYN_CUDA(cudaGetLastError());
    SYN_CUDA(cudaDeviceSynchronize());
    free_state();
    return 0;
}

}  // extern "C"
