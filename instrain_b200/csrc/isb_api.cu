// isb_api.cu -- context, host<->device staging and the extern "C" entry points of libinstrain_b200.so.
// See include/instrain_b200.h for the contract; each entry point cites the reference code it replaces there.
#include "isb_common.cuh"
#include <stdlib.h>

#include <vector>

static char g_create_err[512] = "";

int isb_ensure(isb_ctx *ctx, int slot, size_t bytes)
{
    isb_devbuf &b = ctx->buf[slot];
    if (bytes <= b.cap) return ISB_OK;
    if (b.p) ISB_CUDA(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;     // a little slack so slowly growing batches do not reallocate each time
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        want = bytes;
        ISB_CUDA(cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return ISB_OK;
}

int isb_time_begin(isb_ctx *ctx, int stage)
{
    if (!ctx->timing) return -1;
    if (ctx->n_tev == ctx->cap_tev) {
        const int cap = ctx->cap_tev ? ctx->cap_tev * 2 : 64;
        isb_tev *t = (isb_tev *)realloc(ctx->tev, sizeof(isb_tev) * cap);
        if (!t) return -1;
        ctx->tev = t;
        ctx->cap_tev = cap;
    }
    isb_tev &t = ctx->tev[ctx->n_tev];
    t.stage = stage;
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return -1;
    cudaEventRecord(t.a, ctx->stream);
    return ctx->n_tev++;
}

void isb_time_end(isb_ctx *ctx, int slot)
{
    if (slot >= 0) cudaEventRecord(ctx->tev[slot].b, ctx->stream);
}

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// host pointer -> staged device copy in `slot`; device pointer -> itself
template <class T>
static int stage_in(isb_ctx *ctx, int slot, const T *p, size_t count, const T **out)
{
    if (!p) { *out = nullptr; return ISB_OK; }
    if (is_device_ptr(p)) { *out = p; return ISB_OK; }
    if (count == 0) { *out = nullptr; return ISB_OK; }
    int rc = isb_ensure(ctx, slot, count * sizeof(T) + 16);
    if (rc) return rc;
    ISB_CUDA(cudaMemcpyAsync(ctx->buf[slot].p, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = (const T *)ctx->buf[slot].p;
    return ISB_OK;
}

// output: device pointer -> itself; host pointer or NULL -> scratch in `slot` (copied back by finish_out if host)
template <class T>
static int stage_out(isb_ctx *ctx, int slot, T *p, size_t count, T **out)
{
    if (p && is_device_ptr(p)) { *out = p; return ISB_OK; }
    int rc = isb_ensure(ctx, slot, count * sizeof(T) + 16);
    if (rc) return rc;
    *out = (T *)ctx->buf[slot].p;
    return ISB_OK;
}

template <class T>
static int finish_out(isb_ctx *ctx, T *host, const T *dev, size_t count)
{
    if (!host || (const T *)host == dev || count == 0) return ISB_OK;
    ISB_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return ISB_OK;
}

static int check_dev_err(isb_ctx *ctx)
{
    const unsigned e = *ctx->h_err;
    if (!e) return ISB_OK;
    if (e & ISB_DEV_ERR_ORDER) return isb_fail(ctx, ISB_ERR_ORDER, "events are not position-major (an event lies outside its position tile)");
    if (e & ISB_DEV_ERR_SEG) return isb_fail(ctx, ISB_ERR_ORDER, "read-major batch violates its layout rules (segment order, range or word offsets)");
    if (e & ISB_DEV_ERR_MM) return isb_fail(ctx, ISB_ERR_ARG, "pair_mm value >= M");
    if (e & ISB_DEV_ERR_MULT) return isb_fail(ctx, ISB_ERR_UNSUPPORTED, "a read pair has more than 2 qualifying events on one site");
    return isb_fail(ctx, ISB_ERR_CUDA, "device-side error flag set");
}

static int fetch_status(isb_ctx *ctx)
{
    ISB_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    ISB_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ISB_OK;
}

extern "C" {

int isb_abi_version(void) { return ISB_ABI_VERSION; }

const char *isb_last_error(const isb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

isb_ctx *isb_create(int device, const int32_t *null_lut, int n_lut, int lut_default)
{
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        snprintf(g_create_err, sizeof(g_create_err), "isb_create: no CUDA device (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n_dev || !null_lut || n_lut <= 0) {
        snprintf(g_create_err, sizeof(g_create_err), "isb_create: bad device index %d (of %d) or null model", device, n_dev);
        return nullptr;
    }
    isb_ctx *ctx = new isb_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
#define CREATE_CHECK(call)                                                                                     \
    do {                                                                                                       \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            snprintf(g_create_err, sizeof(g_create_err), "isb_create: %s: %s", #call, cudaGetErrorString(_e)); \
            delete ctx;                                                                                        \
            return nullptr;                                                                                    \
        }                                                                                                      \
    } while (0)
    CREATE_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_CHECK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    {
        int prio_lo = 0, prio_hi = 0;
        CREATE_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CREATE_CHECK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_hi));
    }
    CREATE_CHECK(cudaMalloc(&ctx->d_lut, sizeof(int32_t) * (size_t)n_lut));
    CREATE_CHECK(cudaMemcpy(ctx->d_lut, null_lut, sizeof(int32_t) * (size_t)n_lut, cudaMemcpyHostToDevice));
    ctx->n_lut = n_lut;
    ctx->lut_default = lut_default;
    CREATE_CHECK(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
    CREATE_CHECK(cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
    CREATE_CHECK(cudaMalloc(&ctx->d_err, sizeof(unsigned int)));
    CREATE_CHECK(cudaMemset(ctx->d_err, 0, sizeof(unsigned int)));
    CREATE_CHECK(cudaMallocHost(&ctx->h_counters, 8 * sizeof(unsigned long long)));
    CREATE_CHECK(cudaMallocHost(&ctx->h_err, sizeof(unsigned int)));
    memset(ctx->h_counters, 0, 8 * sizeof(unsigned long long));
    *ctx->h_err = 0;
#undef CREATE_CHECK
    return ctx;
}

void isb_destroy(isb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < SL_COUNT; ++i)
        if (ctx->buf[i].p) cudaFree(ctx->buf[i].p);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_thr2) cudaFree(ctx->d_thr2);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_err) cudaFree(ctx->d_err);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    for (int i = 0; i < ctx->n_tev; ++i) { cudaEventDestroy(ctx->tev[i].a); cudaEventDestroy(ctx->tev[i].b); }
    free(ctx->tev);
    delete ctx;
}

int isb_set_stream(isb_ctx *ctx, void *stream)
{
    if (!ctx) return ISB_ERR_ARG;
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return ISB_OK;
}

int isb_synchronize(isb_ctx *ctx)
{
    if (!ctx) return ISB_ERR_ARG;
    ISB_CUDA(cudaSetDevice(ctx->device));
    int rc = fetch_status(ctx);
    if (rc) return rc;
    return check_dev_err(ctx);
}

int isb_row_counts_async(isb_ctx *ctx, int64_t *dst)
{
    if (!ctx || !dst) return ISB_ERR_ARG;
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemcpyAsync(dst, ctx->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDefault, ctx->stream));
    return ISB_OK;
}

int64_t isb_launch_count(const isb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int64_t isb_selftest_division(isb_ctx *ctx, int s_lo, int s_hi)
{
    if (!ctx || s_lo < 1) return -1;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    unsigned long long bad = 0;
    if (isb_k2_selftest_division(ctx, s_lo, s_hi, &bad) != ISB_OK) return -1;
    return (int64_t)bad;
}

int isb_enable_timing(isb_ctx *ctx, int on)
{
    if (!ctx) return ISB_ERR_ARG;
    ctx->timing = on ? 1 : 0;
    return ISB_OK;
}

int isb_stage_times(isb_ctx *ctx, double ms[3], int64_t calls[3])
{
    if (!ctx || !ms || !calls) return ISB_ERR_ARG;
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) { ms[i] = 0.0; calls[i] = 0; }
    for (int i = 0; i < ctx->n_tev; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->tev[i].a, ctx->tev[i].b) == cudaSuccess) {
            ms[ctx->tev[i].stage] += t;
            calls[ctx->tev[i].stage] += 1;
        }
        cudaEventDestroy(ctx->tev[i].a);
        cudaEventDestroy(ctx->tev[i].b);
    }
    (void)cudaGetLastError();
    ctx->n_tev = 0;
    return ISB_OK;
}

static int check_common(isb_ctx *ctx, int32_t L, int M)
{
    if (!ctx) return ISB_ERR_ARG;
    if (M < 1 || M > ISB_MAX_MM) return isb_fail(ctx, ISB_ERR_ARG, "M (mm levels) must be in [1, 64]");
    if (L < 0) return isb_fail(ctx, ISB_ERR_ARG, "L < 0");
    return ISB_OK;
}

int isb_pileup_counts(isb_ctx *ctx, int64_t n_events, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                      const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start, int32_t L, int M,
                      int min_qual, uint32_t flags, int32_t *counts, uint64_t *nmask)
{
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!counts || (n_events > 0 && (!ref_pos || !base || !qual)) || (M > 1 && n_events > 0 && (!read_id || !pair_mm)))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_pileup_counts: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    const int32_t *d_pos; const uint8_t *d_base, *d_qual, *d_mm; const int32_t *d_rid;
    if ((rc = stage_in(ctx, SL_REF_POS, ref_pos, (size_t)n_events, &d_pos))) return rc;
    if ((rc = stage_in(ctx, SL_BASE, base, (size_t)n_events, &d_base))) return rc;
    if ((rc = stage_in(ctx, SL_QUAL, qual, (size_t)n_events, &d_qual))) return rc;
    if ((rc = stage_in(ctx, SL_READ_ID, M > 1 ? read_id : nullptr, (size_t)n_events, &d_rid))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, M > 1 ? pair_mm : nullptr, (size_t)n_pairs, &d_mm))) return rc;
    int32_t *d_counts; uint64_t *d_nmask = nullptr;
    if ((rc = stage_out(ctx, SL_COUNTS, counts, (size_t)L * M * 4, &d_counts))) return rc;
    if (nmask && (rc = stage_out(ctx, SL_NMASK, nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = isb_k1_launch(ctx, n_events, d_pos, d_base, d_qual, d_rid, d_mm, start, L, M, min_qual, flags, d_counts,
                            (unsigned long long *)d_nmask))) return rc;
    if ((rc = finish_out(ctx, counts, d_counts, (size_t)L * M * 4))) return rc;
    if ((rc = finish_out(ctx, nmask, d_nmask, (size_t)L))) return rc;
    if ((rc = fetch_status(ctx))) return rc;
    return check_dev_err(ctx);
}

int isb_call_snvs(isb_ctx *ctx, int32_t L, int M, const int32_t *counts, const uint64_t *nmask, const uint8_t *ref,
                  int32_t start, int min_cov, double min_freq, int32_t *covT, float *clonT, uint8_t *site_flags,
                  isb_snv_row *rows, int64_t cap, int64_t *n_rows)
{
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!counts || !ref || !n_rows || (cap > 0 && !rows)) return isb_fail(ctx, ISB_ERR_ARG, "isb_call_snvs: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    const int32_t *d_counts; const uint64_t *d_nmask; const uint8_t *d_ref;
    if ((rc = stage_in(ctx, SL_COUNTS, counts, (size_t)L * M * 4, &d_counts))) return rc;
    if ((rc = stage_in(ctx, SL_NMASK, nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = stage_in(ctx, SL_REF, ref, (size_t)L, &d_ref))) return rc;
    int32_t *d_covT; float *d_clonT; uint8_t *d_flags; isb_snv_row *d_rows;
    if ((rc = stage_out(ctx, SL_COVT, covT, (size_t)L * M, &d_covT))) return rc;
    if ((rc = stage_out(ctx, SL_CLONT, clonT, (size_t)L * M, &d_clonT))) return rc;
    if ((rc = stage_out(ctx, SL_FLAGS, site_flags, (size_t)L, &d_flags))) return rc;
    if ((rc = stage_out(ctx, SL_SNV, rows, (size_t)(cap > 0 ? cap : 1), &d_rows))) return rc;
    if ((rc = isb_k2_launch(ctx, L, M, d_counts, (const unsigned long long *)d_nmask, d_ref, start, min_cov, min_freq,
                            d_covT, d_clonT, d_flags, d_rows, cap))) return rc;
    if ((rc = finish_out(ctx, covT, d_covT, (size_t)L * M))) return rc;
    if ((rc = finish_out(ctx, clonT, d_clonT, (size_t)L * M))) return rc;
    if ((rc = finish_out(ctx, site_flags, d_flags, (size_t)L))) return rc;
    if ((rc = fetch_status(ctx))) return rc;
    *n_rows = (int64_t)ctx->h_counters[0];
    const int64_t n_copy = *n_rows < cap ? *n_rows : cap;
    if ((rc = finish_out(ctx, rows, d_rows, (size_t)n_copy))) return rc;
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((rc = check_dev_err(ctx))) return rc;
    if (*n_rows > cap) return isb_fail(ctx, ISB_ERR_CAPACITY, "isb_call_snvs: row buffer too small");
    return ISB_OK;
}

int isb_linkage(isb_ctx *ctx, int64_t n_events, const int32_t *ref_pos, const uint8_t *base, const uint8_t *qual,
                const int32_t *read_id, int64_t n_pairs, const uint8_t *pair_mm, int32_t start, int32_t L, int M,
                int min_qual, const int32_t *counts, const uint64_t *nmask, const uint8_t *site_flags,
                int32_t n_splits, const int32_t *splits, int min_snp, isb_ld_row *rows, int64_t cap, int64_t *n_rows)
{
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!counts || !site_flags || !n_rows || (cap > 0 && !rows) || (n_splits > 0 && !splits) ||
        (n_events > 0 && (!ref_pos || !base || !qual || !read_id)) || (M > 1 && !pair_mm))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_linkage: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    const int32_t *d_pos, *d_rid, *d_counts, *d_splits; const uint8_t *d_base, *d_qual, *d_mm, *d_flags; const uint64_t *d_nmask;
    if ((rc = stage_in(ctx, SL_REF_POS, ref_pos, (size_t)n_events, &d_pos))) return rc;
    if ((rc = stage_in(ctx, SL_BASE, base, (size_t)n_events, &d_base))) return rc;
    if ((rc = stage_in(ctx, SL_QUAL, qual, (size_t)n_events, &d_qual))) return rc;
    if ((rc = stage_in(ctx, SL_READ_ID, read_id, (size_t)n_events, &d_rid))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, pair_mm, (size_t)n_pairs, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_COUNTS, counts, (size_t)L * M * 4, &d_counts))) return rc;
    if ((rc = stage_in(ctx, SL_NMASK, nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = stage_in(ctx, SL_FLAGS, site_flags, (size_t)L, &d_flags))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, splits, (size_t)n_splits * 2, &d_splits))) return rc;
    isb_ld_row *d_rows;
    ctx->seed = 0;
    if ((rc = stage_out(ctx, SL_LD, rows, (size_t)(cap > 0 ? cap : 1), &d_rows))) return rc;
    if ((rc = isb_k3_launch(ctx, n_events, d_pos, d_base, d_qual, d_rid, n_pairs, d_mm, start, L, M, min_qual, d_counts,
                            (const unsigned long long *)d_nmask, d_flags, n_splits, d_splits, min_snp, d_rows, cap))) return rc;
    if ((rc = fetch_status(ctx))) return rc;
    *n_rows = (int64_t)ctx->h_counters[1];
    const int64_t n_copy = *n_rows < cap ? *n_rows : cap;
    if ((rc = finish_out(ctx, rows, d_rows, (size_t)n_copy))) return rc;
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((rc = check_dev_err(ctx))) return rc;
    if (*n_rows > cap) return isb_fail(ctx, ISB_ERR_CAPACITY, "isb_linkage: row buffer too small");
    return ISB_OK;
}

int isb_scaffold_summary(isb_ctx *ctx, int32_t L, int M, const int32_t *covT, const float *clonT, const uint64_t *nmask,
                         int32_t n_scaffolds, const int32_t *scaffold_off, isb_summary_row *out)
{
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!covT || !clonT || !scaffold_off || !out || n_scaffolds < 0)
        return isb_fail(ctx, ISB_ERR_ARG, "isb_scaffold_summary: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    const int32_t *d_cov, *d_off; const float *d_clon; const uint64_t *d_nm;
    // covT / clonT / nmask may still sit in the context's output staging slots (host callers pass host copies)
    if ((rc = stage_in(ctx, SL_COVT, covT, (size_t)L * M, &d_cov))) return rc;
    if ((rc = stage_in(ctx, SL_CLONT, clonT, (size_t)L * M, &d_clon))) return rc;
    if ((rc = stage_in(ctx, SL_NMASK, nmask, (size_t)L, &d_nm))) return rc;
    if ((rc = stage_in(ctx, SL_K4_OFF, scaffold_off, (size_t)n_scaffolds + 1, &d_off))) return rc;
    isb_summary_row *d_out;
    if ((rc = stage_out(ctx, SL_K4_OUT, out, (size_t)n_scaffolds * M, &d_out))) return rc;
    if ((rc = isb_k4_launch(ctx, L, M, d_cov, d_clon, (const unsigned long long *)d_nm, n_scaffolds, d_off, d_out))) return rc;
    if ((rc = finish_out(ctx, out, d_out, (size_t)n_scaffolds * M))) return rc;
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ISB_OK;
}

// K1 -> K2 -> K3 on device-resident inputs; outputs staged / copied as requested by `out`.
static int profile_device(isb_ctx *ctx, isb_reads_dev *rd, isb_cols_dev *cd, int64_t n, const int32_t *d_pos, const uint8_t *d_base, const uint8_t *d_qual,
                          const int32_t *d_rid, int64_t n_pairs, const uint8_t *d_mm, int32_t start, int32_t L, int M,
                          const uint8_t *d_ref, int32_t n_splits, const int32_t *d_splits, const isb_params *prm,
                          isb_result *out)
{
    int rc;
    const bool do_ld = !(prm->flags & ISB_SKIP_LINKAGE);
    int64_t pipe_sites = -1, pipe_pairs = -1;
    int32_t *d_counts = nullptr, *d_covT; uint64_t *d_nmask = nullptr; float *d_clonT; uint8_t *d_flags; isb_snv_row *d_snv; isb_ld_row *d_ld;
    // Read-major segments at M = 1 when the caller does not ask for the raw counts (the reference stores none either,
    // profile_utilities.py:195-216): ONE kernel does pileup + SNV call + the bit rows of the linkage sites (K1f), the
    // linkage back end follows without a host round trip.  No dense counts array exists on this path.
    static const int k1f_env = getenv("ISB_K1F") ? atoi(getenv("ISB_K1F")) : 1;
    const bool fused_reads = rd && M == 1 && !out->counts && k1f_env != 0;
    if (!fused_reads && (rc = stage_out(ctx, SL_COUNTS, out->counts, (size_t)L * M * 4, &d_counts))) return rc;
    if ((!fused_reads || out->nmask || rd->n_nev > 0) && (rc = stage_out(ctx, SL_NMASK, out->nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = stage_out(ctx, SL_COVT, out->covT, (size_t)L * M, &d_covT))) return rc;
    if ((rc = stage_out(ctx, SL_CLONT, out->clonT, (size_t)L * M, &d_clonT))) return rc;
    if ((rc = stage_out(ctx, SL_FLAGS, out->site_flags, (size_t)L, &d_flags))) return rc;
    // rarefied clonality: only when the caller wants it (isb_result.clonTR) and gives a rarefied coverage
    float *d_clonTR = nullptr;
    const int cov_r = out->clonTR && prm->rarefied_cov > 0 ? prm->rarefied_cov : 0;
    if (out->clonTR && (rc = stage_out(ctx, SL_CLONTR, out->clonTR, (size_t)L * M, &d_clonTR))) return rc;
    ctx->seed = prm->seed;
    const int64_t snv_cap = out->snv ? out->snv_cap : 0, ld_cap = out->ld ? out->ld_cap : 0;
    if ((rc = stage_out(ctx, SL_SNV, out->snv, (size_t)(snv_cap > 0 ? snv_cap : 1), &d_snv))) return rc;
    if ((rc = stage_out(ctx, SL_LD, out->ld, (size_t)(ld_cap > 0 ? ld_cap : 1), &d_ld))) return rc;

    // ---- chunk pipeline ------------------------------------------------------------------------------------------
    // (opt-in, ISB_PIPELINE) K1 is HBM-bound, K2 / K3 are latency-bound kernels with little DRAM traffic.  Large batches can be cut at
    // split boundaries into a few chunks: the K1 launches run back to back on the main stream, and K2 + K3 of chunk c
    // run on a second (higher-priority) stream as soon as K1(c) is done, i.e. underneath K1(c+1).  Results are
    // identical (linkage never crosses a split; rows are appended through the same atomic counters).
    std::vector<int32_t> hs;
    std::vector<int> cut;                                    // chunk c = splits [cut[c], cut[c+1])
    const bool want_pipe = !rd && (prm->flags & ISB_PIPELINE) && L >= (1 << 22) && n_splits >= 16 && (!cd || do_ld);
    if (want_pipe) {
        hs.resize((size_t)n_splits * 2);
        ISB_CUDA(cudaMemcpyAsync(hs.data(), d_splits, sizeof(int32_t) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
        ISB_CUDA(cudaStreamSynchronize(ctx->stream));
        bool tiling = hs[0] == start && hs[hs.size() - 1] == start + L - 1;
        for (int i = 0; tiling && i + 1 < n_splits; ++i) tiling = hs[2 * (i + 1)] == hs[2 * i + 1] + 1;
        if (tiling) {                                        // splits tile [start, start+L): cut into <= 8 chunks
            const int n_chunks = 8;
            cut.push_back(0);
            for (int c = 1; c < n_chunks; ++c) {
                const int64_t target = start + (int64_t)L * c / n_chunks;
                int i = cut.back();
                while (i < n_splits && hs[2 * i] < target) ++i;
                if (i > cut.back() && i < n_splits) cut.push_back(i);
            }
            cut.push_back(n_splits);
        }
    }
    static const int fuse_env = getenv("ISB_K1C_FUSE") ? atoi(getenv("ISB_K1C_FUSE")) : 1;
    if (fused_reads) {
        isb_k2_fuse fz = {d_ref, prm->min_cov, prm->min_freq, d_covT, d_clonT, d_flags, d_snv, snv_cap, 0, d_clonTR, cov_r, prm->seed};
        isb_k1f_linkage lk = {n_splits, d_splits, prm->min_snp, d_ld, ld_cap};
        for (int attempt = 0;; ++attempt) {
            if ((rc = isb_k1f_profile_launch(ctx, rd, n_pairs, start, L, (unsigned long long *)d_nmask, &fz, do_ld ? &lk : nullptr))) return rc;
            if (prm->flags & ISB_NO_SYNC) break;                 // the caller checks capacities itself (isb_synchronize)
            if ((rc = fetch_status(ctx))) return rc;
            bool again = false;
            if ((rc = isb_k1f_grow(ctx, &again))) return rc;
            if (!again) break;
            if (attempt == 4) return isb_fail(ctx, ISB_ERR_CUDA, "fused read-major path: scratch sizing did not converge");
        }
    } else if (cut.size() >= 3 && cd) {
        // Column words: the K1c launches (fused with the SNV call at M = 1) of all chunks go back to back on the main
        // stream; chunk c's linkage (+ K2 when not fused) runs on the second stream underneath K1c of the later chunks.
        // K1c is cut at 64-position (group) boundaries rounded UP from the split boundaries, so chunk c's splits are
        // complete once K1c(c) is; K3 addresses the batch's column lists through col_shift.
        const int n_chunks = (int)cut.size() - 1;
        cudaStream_t main_st = ctx->stream, aux = ctx->aux_stream;
        const bool fused = M == 1 && !out->counts && !out->nmask && fuse_env != 0;
        if (fused && cd->n_nev == 0) d_nmask = nullptr;
        std::vector<int64_t> a((size_t)n_chunks + 1);
        a[0] = start;
        for (int c = 1; c < n_chunks; ++c) {
            int64_t x = (int64_t)start + ((((int64_t)hs[2 * cut[c]] - start) + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP) * ISB_COLS_GROUP;
            if (x > (int64_t)start + L) x = (int64_t)start + L;
            a[c] = x < a[c - 1] ? a[c - 1] : x;
        }
        a[n_chunks] = (int64_t)start + L;
        std::vector<cudaEvent_t> ev((size_t)n_chunks + 1);
        for (auto &e : ev) ISB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ISB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(unsigned long long), main_st));
        if (d_nmask) {                                       // N-event bits once for the whole batch
            ISB_CUDA(cudaMemsetAsync(d_nmask, 0, sizeof(uint64_t) * (size_t)L, main_st));
            if ((rc = isb_k1r_n_events_launch(ctx, cd->n_nev, cd->nev_pos, cd->nev_pair, d_mm, n_pairs, start, L, M,
                                              (unsigned long long *)d_nmask))) return rc;
        }
        ctx->keep_counters = 1;
        int ts = isb_time_begin(ctx, 0);
        rc = ISB_OK;
        for (int c = 0; c < n_chunks && rc == ISB_OK; ++c) {
            const int64_t len = a[c + 1] - a[c];
            if (len > 0) {
                const size_t off = (size_t)(a[c] - start);
                isb_cols_dev cdk = *cd;
                cdk.grp_off = cd->grp_off + off / ISB_COLS_GROUP;
                cdk.n_groups = (len + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP;
                isb_k2_fuse fz = {d_ref + off, prm->min_cov, prm->min_freq, d_covT + off * M, d_clonT + off * M, d_flags + off, d_snv,
                                  snv_cap, fuse_env == 2 ? 1 : 0, d_clonTR ? d_clonTR + off * M : nullptr, cov_r, prm->seed};
                rc = isb_k1c_launch(ctx, &cdk, d_mm, n_pairs, (int32_t)a[c], (int32_t)len, M, d_counts + off * M * 4,
                                    d_nmask ? (unsigned long long *)d_nmask + off : nullptr, fused ? &fz : nullptr, false);
            }
            if (rc == ISB_OK) cudaEventRecord(ev[c], main_st);
        }
        isb_time_end(ctx, ts);
        if (rc) { ctx->keep_counters = 0; for (auto &e : ev) cudaEventDestroy(e); return rc; }
        ctx->stream = aux;
        int64_t acc_sites = 0, acc_pairs = 0;
        for (int c = 0; c < n_chunks && rc == ISB_OK; ++c) {
            const int32_t c_lo = hs[2 * cut[c]], c_len = hs[2 * (cut[c + 1] - 1) + 1] - c_lo + 1;
            const size_t off = (size_t)(c_lo - start);
            cudaStreamWaitEvent(aux, ev[c], 0);
            if (!fused) {
                int t2 = isb_time_begin(ctx, 1);
                rc = isb_k2_launch(ctx, c_len, M, d_counts + off * M * 4, d_nmask ? (const unsigned long long *)d_nmask + off : nullptr,
                                   d_ref + off, c_lo, prm->min_cov, prm->min_freq, d_covT + off * M, d_clonT + off * M, d_flags + off,
                                   d_snv, snv_cap, d_clonTR ? d_clonTR + off * M : nullptr, cov_r, prm->seed);
                isb_time_end(ctx, t2);
            }
            if (rc == ISB_OK) {
                int t3 = isb_time_begin(ctx, 2);
                rc = isb_k3_launch_cols(ctx, cd, n_pairs, d_mm, c_lo, c_len, M, d_counts + off * M * 4,
                                        d_nmask ? (const unsigned long long *)d_nmask + off : nullptr, d_flags + off,
                                        cut[c + 1] - cut[c], d_splits + 2 * cut[c], prm->min_snp, d_ld, ld_cap, (int32_t)(c_lo - start));
                isb_time_end(ctx, t3);
                acc_sites += (int64_t)ctx->h_counters[2];
                acc_pairs += (int64_t)ctx->h_counters[3];
            }
        }
        cudaEventRecord(ev[n_chunks], aux);
        ctx->stream = main_st;
        ctx->keep_counters = 0;
        cudaStreamWaitEvent(main_st, ev[n_chunks], 0);
        for (auto &e : ev) cudaEventDestroy(e);
        if (rc) return rc;
        pipe_sites = acc_sites;
        pipe_pairs = acc_pairs;
    } else if (cut.size() >= 3) {
        const int n_chunks = (int)cut.size() - 1;
        cudaStream_t main_st = ctx->stream, aux = ctx->aux_stream;
        std::vector<cudaEvent_t> ev(n_chunks + 1);
        for (auto &e : ev) ISB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ISB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(unsigned long long), main_st));
        int ts = isb_time_begin(ctx, 0);
        for (int c = 0; c < n_chunks; ++c) {
            const int32_t c_lo = hs[2 * cut[c]], c_len = hs[2 * (cut[c + 1] - 1) + 1] - c_lo + 1;
            const size_t off = (size_t)(c_lo - start);
            if ((rc = isb_k1_launch(ctx, n, d_pos, d_base, d_qual, d_rid, d_mm, c_lo, c_len, M, prm->min_qual, 0,
                                    d_counts + off * M * 4, (unsigned long long *)d_nmask + off))) return rc;
            ISB_CUDA(cudaEventRecord(ev[c], main_st));
        }
        isb_time_end(ctx, ts);
        ctx->keep_counters = 1;
        ctx->stream = aux;
        int64_t acc_sites = 0, acc_pairs = 0;
        rc = ISB_OK;
        for (int c = 0; c < n_chunks && rc == ISB_OK; ++c) {
            const int32_t c_lo = hs[2 * cut[c]], c_len = hs[2 * (cut[c + 1] - 1) + 1] - c_lo + 1;
            const size_t off = (size_t)(c_lo - start);
            cudaStreamWaitEvent(aux, ev[c], 0);
            int t2 = isb_time_begin(ctx, 1);
            rc = isb_k2_launch(ctx, c_len, M, d_counts + off * M * 4, (const unsigned long long *)d_nmask + off, d_ref + off,
                               c_lo, prm->min_cov, prm->min_freq, d_covT + off * M, d_clonT + off * M, d_flags + off,
                               d_snv, snv_cap, d_clonTR ? d_clonTR + off * M : nullptr, cov_r, prm->seed);
            isb_time_end(ctx, t2);
            if (rc == ISB_OK && do_ld) {
                int t3 = isb_time_begin(ctx, 2);
                rc = isb_k3_launch(ctx, n, d_pos, d_base, d_qual, d_rid, n_pairs, d_mm, c_lo, c_len, M, prm->min_qual,
                                   d_counts + off * M * 4, (const unsigned long long *)d_nmask + off, d_flags + off,
                                   cut[c + 1] - cut[c], d_splits + 2 * cut[c], prm->min_snp, d_ld, ld_cap);
                isb_time_end(ctx, t3);
                acc_sites += (int64_t)ctx->h_counters[2];
                acc_pairs += (int64_t)ctx->h_counters[3];
            }
        }
        cudaEventRecord(ev[n_chunks], aux);
        ctx->stream = main_st;
        ctx->keep_counters = 0;
        cudaStreamWaitEvent(main_st, ev[n_chunks], 0);
        for (auto &e : ev) cudaEventDestroy(e);
        if (rc) return rc;
        pipe_sites = acc_sites;
        pipe_pairs = acc_pairs;
    } else {
        // Column words at M = 1 when the caller wants neither counts nor nmask: the SNV call runs in K1c's epilogue (one
        // pass; counts are written at flagged sites only, nmask exists only if the batch has N events).
        const bool fused = cd && M == 1 && !out->counts && !out->nmask && fuse_env != 0;
        if (fused && cd->n_nev == 0) d_nmask = nullptr;
        int ts = isb_time_begin(ctx, 0);
        if (cd) {
            isb_k2_fuse fz = {d_ref, prm->min_cov, prm->min_freq, d_covT, d_clonT, d_flags, d_snv, snv_cap, fuse_env == 2 ? 1 : 0,
                              d_clonTR, cov_r, prm->seed};
            rc = isb_k1c_launch(ctx, cd, d_mm, n_pairs, start, L, M, d_counts, (unsigned long long *)d_nmask, fused ? &fz : nullptr);
        } else if (rd) rc = isb_k1r_launch(ctx, rd, d_mm, n_pairs, start, L, M, d_counts, (unsigned long long *)d_nmask);
        else rc = isb_k1_launch(ctx, n, d_pos, d_base, d_qual, d_rid, d_mm, start, L, M, prm->min_qual, 0, d_counts,
                                (unsigned long long *)d_nmask);
        if (rc) return rc;
        isb_time_end(ctx, ts);
        if (!fused) {
            ts = isb_time_begin(ctx, 1);
            if ((rc = isb_k2_launch(ctx, L, M, d_counts, (const unsigned long long *)d_nmask, d_ref, start, prm->min_cov,
                                    prm->min_freq, d_covT, d_clonT, d_flags, d_snv, snv_cap, d_clonTR, cov_r, prm->seed))) return rc;
            isb_time_end(ctx, ts);
        }
        ISB_CUDA(cudaMemsetAsync(ctx->d_counters + 1, 0, 4 * sizeof(unsigned long long), ctx->stream));
        ts = do_ld ? isb_time_begin(ctx, 2) : -1;
        if (do_ld && cd) rc = isb_k3_launch_cols(ctx, cd, n_pairs, d_mm, start, L, M, d_counts, (const unsigned long long *)d_nmask,
                                                 d_flags, n_splits, d_splits, prm->min_snp, d_ld, ld_cap);
        else if (do_ld && rd) rc = isb_k3_launch_reads(ctx, rd, n_pairs, d_mm, start, L, M, d_counts, (const unsigned long long *)d_nmask,
                                                  d_flags, n_splits, d_splits, prm->min_snp, d_ld, ld_cap);
        else if (do_ld) rc = isb_k3_launch(ctx, n, d_pos, d_base, d_qual, d_rid, n_pairs, d_mm, start, L, M, prm->min_qual,
                                           d_counts, (const unsigned long long *)d_nmask, d_flags, n_splits, d_splits,
                                           prm->min_snp, d_ld, ld_cap);
        if (rc) return rc;
        isb_time_end(ctx, ts);
    }
    if (d_counts && (rc = finish_out(ctx, out->counts, d_counts, (size_t)L * M * 4))) return rc;
    if (d_nmask && (rc = finish_out(ctx, out->nmask, d_nmask, (size_t)L))) return rc;
    if ((rc = finish_out(ctx, out->covT, d_covT, (size_t)L * M))) return rc;
    if ((rc = finish_out(ctx, out->clonT, d_clonT, (size_t)L * M))) return rc;
    if (d_clonTR && (rc = finish_out(ctx, out->clonTR, d_clonTR, (size_t)L * M))) return rc;
    if ((rc = finish_out(ctx, out->site_flags, d_flags, (size_t)L))) return rc;
    if (prm->flags & ISB_NO_SYNC) {
        out->n_snv = out->n_ld = out->n_sites = out->n_site_pairs = -1;
        return ISB_OK;
    }
    if ((rc = fetch_status(ctx))) return rc;
    out->n_snv = (int64_t)ctx->h_counters[0];
    out->n_ld = (int64_t)ctx->h_counters[1];
    out->n_sites = pipe_sites >= 0 ? pipe_sites : (int64_t)ctx->h_counters[2];
    out->n_site_pairs = pipe_pairs >= 0 ? pipe_pairs : (int64_t)ctx->h_counters[3];
    if ((rc = finish_out(ctx, out->snv, d_snv, (size_t)(out->n_snv < snv_cap ? out->n_snv : snv_cap)))) return rc;
    if ((rc = finish_out(ctx, out->ld, d_ld, (size_t)(out->n_ld < ld_cap ? out->n_ld : ld_cap)))) return rc;
    ISB_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((rc = check_dev_err(ctx))) return rc;
    if ((out->snv && out->n_snv > snv_cap) || (out->ld && do_ld && out->n_ld > ld_cap))
        return isb_fail(ctx, ISB_ERR_CAPACITY, "isb_profile_batch: row buffer too small (see n_snv / n_ld)");
    return ISB_OK;
}

int isb_profile_batch(isb_ctx *ctx, const isb_batch *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    const int64_t n = in->n_events;
    if (!in->ref || (n > 0 && (!in->ref_pos || !in->base || !in->qual || !in->read_id)) || (M > 1 && !in->pair_mm) ||
        (in->n_splits > 0 && !in->splits))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_batch: null input pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    const int32_t *d_pos, *d_rid, *d_splits; const uint8_t *d_base, *d_qual, *d_mm, *d_ref;
    if ((rc = stage_in(ctx, SL_REF_POS, in->ref_pos, (size_t)n, &d_pos))) return rc;
    if ((rc = stage_in(ctx, SL_BASE, in->base, (size_t)n, &d_base))) return rc;
    if ((rc = stage_in(ctx, SL_QUAL, in->qual, (size_t)n, &d_qual))) return rc;
    if ((rc = stage_in(ctx, SL_READ_ID, in->read_id, (size_t)n, &d_rid))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    return profile_device(ctx, nullptr, nullptr, n, d_pos, d_base, d_qual, d_rid, in->n_pairs, d_mm, in->start, L, M, d_ref, in->n_splits,
                          d_splits, prm, out);
}

int isb_profile_batch_packed(isb_ctx *ctx, const isb_packed_batch *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    const int64_t n = in->n_events;
    if (!in->ref || !in->pos_off || (L > 0 && !in->id_base) || (n > 0 && !in->bqd) || (M > 1 && !in->pair_mm) ||
        (in->n_splits > 0 && !in->splits) || (in->n_esc > 0 && (!in->esc_evt || !in->esc_id)))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_batch_packed: null input pointer");
    if (in->min_qual != prm->min_qual)
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_batch_packed: the quality bit was packed with a different min_qual");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    const int64_t *d_off, *d_esce; const int32_t *d_idb, *d_esci, *d_splits; const uint8_t *d_bqd, *d_mm, *d_ref;
    if ((rc = stage_in(ctx, SL_PK_OFF, in->pos_off, (size_t)L + 1, &d_off))) return rc;
    if ((rc = stage_in(ctx, SL_PK_IDBASE, in->id_base, (size_t)L, &d_idb))) return rc;
    if ((rc = stage_in(ctx, SL_PK_BQD, in->bqd, (size_t)n, &d_bqd))) return rc;
    if ((rc = stage_in(ctx, SL_PK_ESC_EVT, in->esc_evt, (size_t)in->n_esc, &d_esce))) return rc;
    if ((rc = stage_in(ctx, SL_PK_ESC_ID, in->esc_id, (size_t)in->n_esc, &d_esci))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    // canonical columns in context-owned HBM (16 spare events so every column is readable in whole 16-byte granules)
    if ((rc = isb_ensure(ctx, SL_REF_POS, sizeof(int32_t) * ((size_t)n + 16)))) return rc;
    if ((rc = isb_ensure(ctx, SL_READ_ID, sizeof(int32_t) * ((size_t)n + 16)))) return rc;
    if ((rc = isb_ensure(ctx, SL_BASE, (size_t)n + 16))) return rc;
    if ((rc = isb_ensure(ctx, SL_QUAL, (size_t)n + 16))) return rc;
    int32_t *c_pos = (int32_t *)ctx->buf[SL_REF_POS].p, *c_rid = (int32_t *)ctx->buf[SL_READ_ID].p;
    uint8_t *c_base = (uint8_t *)ctx->buf[SL_BASE].p, *c_qual = (uint8_t *)ctx->buf[SL_QUAL].p;
    int qpass = prm->min_qual < 1 ? 1 : (prm->min_qual > 255 ? 255 : prm->min_qual);
    if ((rc = isb_k0_launch(ctx, n, d_off, d_idb, d_bqd, in->n_esc, d_esce, d_esci, in->start, L, qpass, c_pos, c_base,
                            c_qual, c_rid))) return rc;
    return profile_device(ctx, nullptr, nullptr, n, c_pos, c_base, c_qual, c_rid, in->n_pairs, d_mm, in->start, L, M, d_ref, in->n_splits,
                          d_splits, prm, out);
}

// host or device read-major batch -> device pointers (staged through the context's grow-only slots)
static int stage_reads(isb_ctx *ctx, const isb_reads_batch *in, isb_reads_dev *rd, const uint8_t **d_mm)
{
    int rc;
    if (in->n_segs < 0 || in->n_words < 0 || (in->n_words & 3))
        return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: n_segs / n_words invalid (n_words must be a multiple of 4)");
    if (in->n_segs > 0 && (!in->seg_start || !in->seg_len || !in->seg_pair || !in->seg_word || !in->words))
        return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: null segment column");
    memset(rd, 0, sizeof(*rd));
    rd->n_segs = in->n_segs;
    rd->n_words = in->n_words;
    rd->max_seg_len = in->max_seg_len;
    if ((rc = stage_in(ctx, SL_RD_START, in->seg_start, (size_t)in->n_segs, &rd->seg_start))) return rc;
    if ((rc = stage_in(ctx, SL_RD_LEN, in->seg_len, (size_t)in->n_segs, &rd->seg_len))) return rc;
    if ((rc = stage_in(ctx, SL_RD_PAIR, in->seg_pair, (size_t)in->n_segs, &rd->seg_pair))) return rc;
    if ((rc = stage_in(ctx, SL_RD_WORD, in->seg_word, (size_t)in->n_segs, &rd->seg_word))) return rc;
    if ((rc = stage_in(ctx, SL_RD_WORDS, in->words, (size_t)in->n_words, &rd->words))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, d_mm))) return rc;
    if (in->n_nev < 0 || (in->n_nev > 0 && (!in->nev_pos || !in->nev_pair)))
        return isb_fail(ctx, ISB_ERR_ARG, "read-major batch: null N-event column");
    rd->n_nev = in->n_nev;
    if ((rc = stage_in(ctx, SL_RD_NPOS, in->nev_pos, (size_t)in->n_nev, &rd->nev_pos))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPAIR, in->nev_pair, (size_t)in->n_nev, &rd->nev_pair))) return rc;
    return ISB_OK;
}

int isb_pileup_reads(isb_ctx *ctx, const isb_reads_batch *in, int32_t *counts, uint64_t *nmask)
{
    if (!ctx || !in) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!counts || (M > 1 && in->n_segs > 0 && !in->pair_mm)) return isb_fail(ctx, ISB_ERR_ARG, "isb_pileup_reads: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_reads_dev rd;
    const uint8_t *d_mm;
    if ((rc = stage_reads(ctx, in, &rd, &d_mm))) return rc;
    int32_t *d_counts; uint64_t *d_nmask = nullptr;
    if ((rc = stage_out(ctx, SL_COUNTS, counts, (size_t)L * M * 4, &d_counts))) return rc;
    if (nmask && (rc = stage_out(ctx, SL_NMASK, nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = isb_k1r_launch(ctx, &rd, d_mm, in->n_pairs, in->start, L, M, d_counts, (unsigned long long *)d_nmask))) return rc;
    if ((rc = finish_out(ctx, counts, d_counts, (size_t)L * M * 4))) return rc;
    if ((rc = finish_out(ctx, nmask, d_nmask, (size_t)L))) return rc;
    if ((rc = fetch_status(ctx))) return rc;
    return check_dev_err(ctx);
}

int isb_profile_reads(isb_ctx *ctx, const isb_reads_batch *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!in->ref || (M > 1 && !in->pair_mm) || (in->n_splits > 0 && !in->splits))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads: null input pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_reads_dev rd;
    const uint8_t *d_mm, *d_ref; const int32_t *d_splits;
    if ((rc = stage_reads(ctx, in, &rd, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    return profile_device(ctx, &rd, nullptr, 0, nullptr, nullptr, nullptr, nullptr, in->n_pairs, d_mm, in->start, L, M, d_ref,
                          in->n_splits, d_splits, prm, out);
}

// Compact transfer format: copy the 3-bit columns, expand them on the device (K0r) into context-owned HBM, then the same
// K1r -> K2 -> K3 as isb_profile_reads.
int isb_profile_reads_compact(isb_ctx *ctx, const isb_reads_compact *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!in->ref || (M > 1 && !in->pair_mm) || (in->n_splits > 0 && !in->splits))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_compact: null input pointer");
    if (in->n_segs < 0 || in->n_units < 0 || (in->n_segs > 0 && (!in->seg_start || !in->seg_len || !in->seg_pair)) ||
        (in->n_units > 0 && (!in->base2 || !in->pass)) || in->n_nev < 0 || (in->n_nev > 0 && (!in->nev_pos || !in->nev_pair)))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_compact: null or negative-sized segment column");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_reads_dev rd;
    memset(&rd, 0, sizeof(rd));
    rd.n_segs = in->n_segs;
    rd.max_seg_len = in->max_seg_len;
    rd.n_nev = in->n_nev;
    rd.n_words = (1 + in->n_units + in->n_segs + 3) & ~(int64_t)3;
    const uint16_t *d_b2; const uint8_t *d_ps, *d_mm, *d_ref; const int32_t *d_splits;
    if ((rc = stage_in(ctx, SL_RD_START, in->seg_start, (size_t)in->n_segs, &rd.seg_start))) return rc;
    if ((rc = stage_in(ctx, SL_RD_LEN, in->seg_len, (size_t)in->n_segs, &rd.seg_len))) return rc;
    if ((rc = stage_in(ctx, SL_RD_PAIR, in->seg_pair, (size_t)in->n_segs, &rd.seg_pair))) return rc;
    if ((rc = stage_in(ctx, SL_RC_BASE2, in->base2, (size_t)in->n_units, &d_b2))) return rc;
    if ((rc = stage_in(ctx, SL_RC_PASS, in->pass, (size_t)in->n_units, &d_ps))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPOS, in->nev_pos, (size_t)in->n_nev, &rd.nev_pos))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPAIR, in->nev_pair, (size_t)in->n_nev, &rd.nev_pair))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    if ((rc = isb_ensure(ctx, SL_RD_WORD, sizeof(int64_t) * ((size_t)in->n_segs + 2)))) return rc;
    if ((rc = isb_ensure(ctx, SL_RD_WORDS, sizeof(uint32_t) * ((size_t)rd.n_words + 4)))) return rc;
    int64_t *d_seg_word = (int64_t *)ctx->buf[SL_RD_WORD].p;
    uint32_t *d_words = (uint32_t *)ctx->buf[SL_RD_WORDS].p;
    if ((rc = isb_k0r_launch(ctx, in->n_segs, rd.seg_start, rd.seg_len, in->n_units, d_b2, d_ps, d_seg_word, rd.n_words, d_words))) return rc;
    rd.seg_word = d_seg_word;
    rd.words = d_words;
    return profile_device(ctx, &rd, nullptr, 0, nullptr, nullptr, nullptr, nullptr, in->n_pairs, d_mm, in->start, L, M, d_ref,
                          in->n_splits, d_splits, prm, out);
}

// Reference-delta transfer format: copy the event bits and the mismatch list, rebuild the stream on the device (K0d), then
// the same K1r -> K2 -> K3 as isb_profile_reads.
int isb_profile_reads_delta(isb_ctx *ctx, const isb_reads_delta *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!in->ref || (M > 1 && !in->pair_mm) || (in->n_splits > 0 && !in->splits))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_delta: null input pointer");
    if (in->n_segs < 0 || in->n_units < 0 || in->n_mis < 0 || (in->n_segs > 0 && (!in->seg_start || !in->seg_len || !in->seg_pair)) ||
        (in->n_units > 0 && !in->pass) || (in->n_mis > 0 && (!in->mis_word || !in->mis_code)) || in->n_nev < 0 ||
        (in->n_nev > 0 && (!in->nev_pos || !in->nev_pair)))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_delta: null or negative-sized column");
    if (in->start & 7) return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_delta: start must be a multiple of 8");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_reads_dev rd;
    memset(&rd, 0, sizeof(rd));
    rd.n_segs = in->n_segs;
    rd.max_seg_len = in->max_seg_len;
    rd.n_nev = in->n_nev;
    rd.n_words = (1 + in->n_units + in->n_segs + 3) & ~(int64_t)3;
    if (rd.n_words > 0xffffffffll) return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_reads_delta: batch too large for 32-bit word indices");
    const uint8_t *d_ps, *d_mc, *d_mm, *d_ref; const uint32_t *d_mw; const int32_t *d_splits;
    if ((rc = stage_in(ctx, SL_RD_START, in->seg_start, (size_t)in->n_segs, &rd.seg_start))) return rc;
    if ((rc = stage_in(ctx, SL_RD_LEN, in->seg_len, (size_t)in->n_segs, &rd.seg_len))) return rc;
    if ((rc = stage_in(ctx, SL_RD_PAIR, in->seg_pair, (size_t)in->n_segs, &rd.seg_pair))) return rc;
    if ((rc = stage_in(ctx, SL_RC_PASS, in->pass, (size_t)in->n_units, &d_ps))) return rc;
    if ((rc = stage_in(ctx, SL_RC_MISW, in->mis_word, (size_t)in->n_mis, &d_mw))) return rc;
    if ((rc = stage_in(ctx, SL_RC_MISC, in->mis_code, (size_t)in->n_mis, &d_mc))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPOS, in->nev_pos, (size_t)in->n_nev, &rd.nev_pos))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPAIR, in->nev_pair, (size_t)in->n_nev, &rd.nev_pair))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    if ((rc = isb_ensure(ctx, SL_RD_WORD, sizeof(int64_t) * ((size_t)in->n_segs + 2)))) return rc;
    if ((rc = isb_ensure(ctx, SL_RD_WORDS, sizeof(uint32_t) * ((size_t)rd.n_words + 4)))) return rc;
    int64_t *d_seg_word = (int64_t *)ctx->buf[SL_RD_WORD].p;
    uint32_t *d_words = (uint32_t *)ctx->buf[SL_RD_WORDS].p;
    if ((rc = isb_k0d_launch(ctx, in->n_segs, rd.seg_start, rd.seg_len, in->n_units, d_ps, d_ref, in->start, L, in->n_mis, d_mw, d_mc,
                             d_seg_word, rd.n_words, d_words))) return rc;
    rd.seg_word = d_seg_word;
    rd.words = d_words;
    return profile_device(ctx, &rd, nullptr, 0, nullptr, nullptr, nullptr, nullptr, in->n_pairs, d_mm, in->start, L, M, d_ref,
                          in->n_splits, d_splits, prm, out);
}

// host or device column-word batch -> device pointers
static int stage_cols(isb_ctx *ctx, const isb_cols_batch *in, isb_cols_dev *cd, const uint8_t **d_mm)
{
    int rc;
    const int64_t n_groups = ((int64_t)in->L + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP;
    if (in->n_groups != n_groups || in->n_chunks < 0 || !in->grp_off || (in->n_chunks > 0 && !in->words))
        return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: n_groups must be ceil(L / 64); grp_off / words must not be null");
    if (in->n_nev < 0 || (in->n_nev > 0 && (!in->nev_pos || !in->nev_pair)))
        return isb_fail(ctx, ISB_ERR_ARG, "column-word batch: null N-event column");
    memset(cd, 0, sizeof(*cd));
    cd->n_groups = n_groups;
    cd->n_chunks = in->n_chunks;
    cd->n_nev = in->n_nev;
    if ((rc = stage_in(ctx, SL_CD_OFF, in->grp_off, (size_t)n_groups + 1, &cd->grp_off))) return rc;
    if ((rc = stage_in(ctx, SL_CD_WORDS, in->words, (size_t)in->n_chunks * ISB_COLS_CHUNK, &cd->words))) return rc;
    if ((rc = stage_in(ctx, SL_CD_IDS, in->ids, (size_t)in->n_chunks * ISB_COLS_CHUNK, &cd->ids))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPOS, in->nev_pos, (size_t)in->n_nev, &cd->nev_pos))) return rc;
    if ((rc = stage_in(ctx, SL_RD_NPAIR, in->nev_pair, (size_t)in->n_nev, &cd->nev_pair))) return rc;
    if ((rc = stage_in(ctx, SL_PAIR_MM, in->pair_mm, (size_t)in->n_pairs, d_mm))) return rc;
    return ISB_OK;
}

int isb_cols_from_reads(isb_ctx *ctx, const isb_reads_batch *in, int64_t *grp_off, int64_t *n_chunks, uint32_t *words,
                        int32_t *ids, int64_t cap_chunks)
{
    if (!ctx || !in || !grp_off || !n_chunks) return ISB_ERR_ARG;
    int rc = check_common(ctx, in->L, 1);
    if (rc) return rc;
    if (words && (!ids || cap_chunks < 0)) return isb_fail(ctx, ISB_ERR_ARG, "isb_cols_from_reads: ids / cap_chunks invalid");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_reads_dev rd;
    const uint8_t *d_mm;
    isb_reads_batch tmp = *in;
    tmp.pair_mm = nullptr;                                        // not needed here
    tmp.n_pairs = 0;
    if ((rc = stage_reads(ctx, &tmp, &rd, &d_mm))) return rc;
    const size_t n_groups = (size_t)(((int64_t)in->L + ISB_COLS_GROUP - 1) / ISB_COLS_GROUP);
    int64_t *d_off; uint32_t *d_words = nullptr; int32_t *d_ids = nullptr;
    if ((rc = stage_out(ctx, SL_CD_OFF, grp_off, n_groups + 1, &d_off))) return rc;
    if (words) {
        if ((rc = stage_out(ctx, SL_CD_WORDS, words, (size_t)cap_chunks * ISB_COLS_CHUNK, &d_words))) return rc;
        if ((rc = stage_out(ctx, SL_CD_IDS, ids, (size_t)cap_chunks * ISB_COLS_CHUNK, &d_ids))) return rc;
    }
    rc = isb_cols_convert(ctx, &rd, in->start, in->L, d_off, d_words, d_ids, cap_chunks, n_chunks);
    if (rc && rc != ISB_ERR_CAPACITY) return rc;
    const int rc_cap = rc;
    if ((rc = finish_out(ctx, grp_off, d_off, n_groups + 1))) return rc;
    if (words && rc_cap == ISB_OK) {
        if ((rc = finish_out(ctx, words, d_words, (size_t)*n_chunks * ISB_COLS_CHUNK))) return rc;
        if ((rc = finish_out(ctx, ids, d_ids, (size_t)*n_chunks * ISB_COLS_CHUNK))) return rc;
    }
    if ((rc = fetch_status(ctx))) return rc;
    if ((rc = check_dev_err(ctx))) return rc;
    return rc_cap;
}

int isb_pileup_cols(isb_ctx *ctx, const isb_cols_batch *in, int32_t *counts, uint64_t *nmask)
{
    if (!ctx || !in) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!counts || (M > 1 && in->n_chunks > 0 && (!in->pair_mm || !in->ids))) return isb_fail(ctx, ISB_ERR_ARG, "isb_pileup_cols: null pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_cols_dev cd;
    const uint8_t *d_mm;
    if ((rc = stage_cols(ctx, in, &cd, &d_mm))) return rc;
    int32_t *d_counts; uint64_t *d_nmask = nullptr;
    if ((rc = stage_out(ctx, SL_COUNTS, counts, (size_t)L * M * 4, &d_counts))) return rc;
    if (nmask && (rc = stage_out(ctx, SL_NMASK, nmask, (size_t)L, &d_nmask))) return rc;
    if ((rc = isb_k1c_launch(ctx, &cd, d_mm, in->n_pairs, in->start, L, M, d_counts, (unsigned long long *)d_nmask, nullptr))) return rc;
    if ((rc = finish_out(ctx, counts, d_counts, (size_t)L * M * 4))) return rc;
    if ((rc = finish_out(ctx, nmask, d_nmask, (size_t)L))) return rc;
    if ((rc = fetch_status(ctx))) return rc;
    return check_dev_err(ctx);
}

int isb_profile_cols(isb_ctx *ctx, const isb_cols_batch *in, const isb_params *prm, isb_result *out)
{
    if (!ctx || !in || !prm || !out) return ISB_ERR_ARG;
    const int32_t L = in->L;
    const int M = in->M;
    int rc = check_common(ctx, L, M);
    if (rc) return rc;
    if (!in->ref || (M > 1 && !in->pair_mm) || (in->n_splits > 0 && !in->splits) || (in->n_chunks > 0 && !in->ids))
        return isb_fail(ctx, ISB_ERR_ARG, "isb_profile_cols: null input pointer");
    ISB_CUDA(cudaSetDevice(ctx->device));
    ISB_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(unsigned int), ctx->stream));
    isb_cols_dev cd;
    const uint8_t *d_mm, *d_ref; const int32_t *d_splits;
    if ((rc = stage_cols(ctx, in, &cd, &d_mm))) return rc;
    if ((rc = stage_in(ctx, SL_REF, in->ref, (size_t)L, &d_ref))) return rc;
    if ((rc = stage_in(ctx, SL_SPLITS, in->splits, (size_t)in->n_splits * 2, &d_splits))) return rc;
    return profile_device(ctx, nullptr, &cd, 0, nullptr, nullptr, nullptr, nullptr, in->n_pairs, d_mm, in->start, L, M, d_ref,
                          in->n_splits, d_splits, prm, out);
}

}  // extern "C"
