// kpeg_cuda.cu -- C-ABI host side of the B200 decode path (include/kpeg_cuda.h).
//
// Owns a CUDA stream, grow-only device scratch and pinned staging per context, builds the device
// tables from a kpeg_plan and enqueues K0..K3 (kernels.cu).  There is no CPU decode in here: the
// only host arithmetic is table preparation (Huffman LUTs, AAN-prescaled quantisers, the 64 double
// cosines the exact IDCT path needs -- evaluated with the host libm exactly as the reference
// evaluates them, src/MCU.cpp:193).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "host_tables.h"
#include "kernels.cuh"
#include "kpeg_common.h"
#include "kpeg_cuda.h"

using namespace kpeg;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
};

constexpr int MAX_EVENTS = 160;

} // namespace

struct kpeg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool profiling = false;
    cudaEvent_t ev[MAX_EVENTS] = {};
    int ev_stage[MAX_EVENTS] = {}; // stage that ENDS at event i (-1 for the first)
    int nev = 0;
    uint32_t sub_bits = 512;
    int relay_rounds = 8;

    DevBuf scan, words, seg_bit, tile_kept, tile_rst, state, used, seg_hint, start_slot, scan_tiles;
    DevBuf coef, dcdiff, dc, tile_carry, pixels, tables, meta, merged, tie_rec, overflow;
    PinBuf h_tables, h_meta, h_stage;

    kpeg_plan plan_cached;
    bool have_plan = false;

    // last job (for kpeg_cuda_read_coefficients)
    JobGeom last_g = {};
    bool have_last = false;
};

namespace {

int fail(kpeg_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    if (c)
        c->err = buf;
    return code;
}

#define CK(call)                                                                                                       \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(ctx, KPEG_ERR_CUDA, #call, e_);                                                                \
    } while (0)

int ensure(kpeg_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return KPEG_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, KPEG_ERR_NOMEM, "cudaMalloc", e);
    }
    b.cap = want;
    return KPEG_OK;
}

int ensure_pinned(kpeg_ctx *ctx, PinBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return KPEG_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFreeHost(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, KPEG_ERR_NOMEM, "cudaMallocHost", e);
    }
    b.cap = want;
    return KPEG_OK;
}

#define TRY(expr)                                                                                                      \
    do {                                                                                                               \
        int rc_ = (expr);                                                                                              \
        if (rc_ != KPEG_OK)                                                                                            \
            return rc_;                                                                                                \
    } while (0)

int build_tables(kpeg_ctx *ctx, const kpeg_plan *pl, DeviceTables *T)
{
    const char *why = nullptr;
    const int rc = build_device_tables(pl, T, &why);
    return rc == KPEG_OK ? rc : fail(ctx, rc, why);
}

int upload_plan(kpeg_ctx *ctx, const kpeg_plan *pl)
{
    if (ctx->have_plan && memcmp(&ctx->plan_cached, pl, sizeof *pl) == 0)
        return KPEG_OK;
    TRY(ensure_pinned(ctx, ctx->h_tables, sizeof(DeviceTables)));
    TRY(ensure(ctx, ctx->tables, sizeof(DeviceTables)));
    CK(cudaStreamSynchronize(ctx->stream)); // the pinned copy may still be in flight from an earlier plan
    TRY(build_tables(ctx, pl, (DeviceTables *)ctx->h_tables.p));
    CK(cudaMemcpyAsync(ctx->tables.p, ctx->h_tables.p, sizeof(DeviceTables), cudaMemcpyHostToDevice, ctx->stream));
    ctx->plan_cached = *pl;
    ctx->have_plan = true;
    return KPEG_OK;
}

int make_geom(kpeg_ctx *ctx, const kpeg_plan *pl, uint32_t nimages, JobGeom *g)
{
    const char *why = nullptr;
    const int rc = make_job_geom(pl, nimages, ctx->sub_bits, g, &why);
    return rc == KPEG_OK ? rc : fail(ctx, rc, why);
}

// Profiling: an event after every stage; mark(ctx, -1) opens a call.
void mark(kpeg_ctx *ctx, int stage)
{
    if (!ctx->profiling)
        return;
    if (stage < 0)
        ctx->nev = 0;
    if (ctx->nev >= MAX_EVENTS)
        return;
    if (!ctx->ev[ctx->nev])
        cudaEventCreate(&ctx->ev[ctx->nev]);
    cudaEventRecord(ctx->ev[ctx->nev], ctx->stream);
    ctx->ev_stage[ctx->nev] = stage;
    ++ctx->nev;
}

int status_to_rc(kpeg_ctx *ctx, uint32_t st)
{
    if (st == 0)
        return KPEG_OK;
    char buf[160];
    snprintf(buf, sizeof buf, "corrupt entropy-coded data (device status 0x%x:%s%s%s%s%s%s)", st,
             (st & ST_BAD_CODE) ? " bad-code" : "", (st & ST_SLOT_OVERFLOW) ? " slot-overflow" : "",
             (st & ST_SEG_MISMATCH) ? " segment-length" : "", (st & ST_BAD_MARKER) ? " marker-in-scan" : "",
             (st & ST_EXIT_MISMATCH) ? " relay-mismatch" : "", (st & ST_SEG_COUNT) ? " restart-count" : "");
    ctx->err = buf;
    return KPEG_ERR_STREAM;
}

// The whole device pipeline for one job whose stuffed bytes are already at d_scan.
// The caller has opened the profiling window with mark(ctx, -1).
int run_job(kpeg_ctx *ctx, const kpeg_plan *pl, const uint8_t *d_scan, size_t scan_len, uint32_t nimages,
            uint8_t *d_pixels, kpeg_stats *stats, uint32_t *launches_out)
{
    JobGeom g;
    TRY(make_geom(ctx, pl, nimages, &g));
    if (scan_len == 0 || scan_len >= (1ull << 29))
        return fail(ctx, KPEG_ERR_ARG, "entropy-coded segment must be 1 byte .. 512 MiB");
    TRY(upload_plan(ctx, pl));

    const uint32_t S = (uint32_t)scan_len;
    const uint32_t ntiles = (S + UNSTUFF_TILE - 1) / UNSTUFF_TILE;
    const uint32_t nsub_max = (uint32_t)(((uint64_t)S * 8u + g.sub_bits - 1u) / g.sub_bits) + 1u;
    const uint32_t total_mcus = g.nimages * g.mcus_per_image;
    const uint32_t dc_tiles = (total_mcus + DC_TILE - 1) / DC_TILE;
    const size_t words_bytes = ((size_t)S + 3u) / 4u * 4u + 32u;
    const size_t coef_bytes = (size_t)g.total_blocks * 128u;

    TRY(ensure(ctx, ctx->words, words_bytes));
    TRY(ensure(ctx, ctx->seg_bit, ((size_t)g.nseg + 2u) * 4u));
    TRY(ensure(ctx, ctx->tile_kept, (size_t)ntiles * 4u));
    TRY(ensure(ctx, ctx->tile_rst, (size_t)ntiles * 4u));
    TRY(ensure(ctx, ctx->state, (size_t)nsub_max * sizeof(SubState)));
    TRY(ensure(ctx, ctx->used, (size_t)nsub_max * 2u * sizeof(uint32_t)));
    TRY(ensure(ctx, ctx->seg_hint, (size_t)nsub_max * 4u));
    TRY(ensure(ctx, ctx->start_slot, (size_t)nsub_max * 4u));
    TRY(ensure(ctx, ctx->scan_tiles, ((size_t)nsub_max / 1024u + 2u) * sizeof(uint2)));
    TRY(ensure(ctx, ctx->coef, coef_bytes + 256));
    TRY(ensure(ctx, ctx->dcdiff, (size_t)g.total_blocks * 2u + 16));
    TRY(ensure(ctx, ctx->dc, (size_t)g.total_blocks * 2u + 16));
    TRY(ensure(ctx, ctx->tile_carry, (size_t)dc_tiles * 16u));
    // tie records: room for 1/8 of all pixels (typical: ~1 %); beyond that K3 resolves in place
    const uint64_t npix_job = (uint64_t)g.nimages * g.width * g.height;
    const uint32_t tie_cap = (uint32_t)std::min<uint64_t>(npix_job / 8u + 4096u, 1u << 27);
    TRY(ensure(ctx, ctx->tie_rec, (size_t)tie_cap * sizeof(uint4)));
    const size_t overflow_bytes = ((size_t)total_mcus / IDCT_MCUS_PER_CTA + 2u) * sizeof(uint32_t);
    TRY(ensure(ctx, ctx->overflow, overflow_bytes));
    TRY(ensure(ctx, ctx->meta, sizeof(DevMeta)));
    TRY(ensure_pinned(ctx, ctx->h_meta, sizeof(DevMeta)));

    cudaStream_t s = ctx->stream;
    uint32_t launches = 0;
    DevMeta *d_meta = (DevMeta *)ctx->meta.p;

    CK(cudaMemsetAsync(d_meta, 0, sizeof(DevMeta), s));
    CK(cudaMemsetAsync(ctx->words.p, 0, words_bytes, s));
    CK(cudaMemsetAsync(ctx->overflow.p, 0, overflow_bytes, s));
    mark(ctx, KPEG_T_MEMSET);

    UnstuffArgs ua;
    ua.scan = d_scan;
    ua.scan_len = S;
    ua.ntiles = ntiles;
    ua.tile_kept = (uint32_t *)ctx->tile_kept.p;
    ua.tile_rst = (uint32_t *)ctx->tile_rst.p;
    ua.words = (uint8_t *)ctx->words.p;
    ua.seg_bit = (uint32_t *)ctx->seg_bit.p;
    ua.nseg = g.nseg;
    ua.meta = d_meta;
    launch_unstuff(ua, g.sub_bits, s, &launches);
    mark(ctx, KPEG_T_UNSTUFF);

    EntropyArgs ea;
    ea.words = (const uint32_t *)ctx->words.p;
    ea.seg_bit = (const uint32_t *)ctx->seg_bit.p;
    ea.tables = (const DeviceTables *)ctx->tables.p;
    ea.meta = d_meta;
    ea.state = (SubState *)ctx->state.p;
    ea.worklist[0] = (uint32_t *)ctx->used.p;
    ea.worklist[1] = (uint32_t *)ctx->used.p + nsub_max;
    ea.seg_hint = (uint32_t *)ctx->seg_hint.p;
    ea.start_slot = (uint32_t *)ctx->start_slot.p;
    ea.scan_tiles = (uint2 *)ctx->scan_tiles.p;
    ea.coef = (int16_t *)ctx->coef.p;
    ea.dcdiff = (int16_t *)ctx->dcdiff.p;
    ea.nsub_max = nsub_max;
    ea.g = g;

    DcArgs da;
    da.dcdiff = (const int16_t *)ctx->dcdiff.p;
    da.dc = (int16_t *)ctx->dc.p;
    da.tile_carry = (int32_t *)ctx->tile_carry.p;
    da.ntiles = dc_tiles;
    da.g = g;

    IdctArgs ia;
    ia.coef = (const int16_t *)ctx->coef.p;
    ia.dc = (const int16_t *)ctx->dc.p;
    ia.dcdiff = (const int16_t *)ctx->dcdiff.p;
    ia.tables = (const DeviceTables *)ctx->tables.p;
    ia.pixels = d_pixels;
    ia.meta = d_meta;
    ia.tie_rec = (uint4 *)ctx->tie_rec.p;
    ia.tie_cap = tie_cap;
    ia.overflow_mcu = (uint32_t *)ctx->overflow.p;
    ia.g = g;

    launch_entropy_cold(ea, s, &launches);
    mark(ctx, KPEG_T_ENTROPY_COLD);
    int rounds = ctx->relay_rounds < 2 ? 2 : (ctx->relay_rounds > MAX_RELAY_ROUNDS - 1 ? MAX_RELAY_ROUNDS - 1 : ctx->relay_rounds);
    for (int r = 1; r <= rounds; ++r)
        launch_entropy_relay(ea, r, s, &launches);
    mark(ctx, KPEG_T_ENTROPY_RELAY);

    DevMeta *h_meta = (DevMeta *)ctx->h_meta.p;
    uint32_t extra_iterations = 0;
    for (;;) {
        // everything downstream of the relay; optimistic: issued before convergence is known
        // no zero-fill of coef / dcdiff: the final pass writes every slot of every block it owns
        launch_entropy_scan(ea, s, &launches);
        mark(ctx, KPEG_T_ENTROPY_SCAN);
        launch_entropy_write(ea, s, &launches);
        mark(ctx, KPEG_T_ENTROPY_WRITE);
        launch_dc_scan(da, s, &launches);
        mark(ctx, KPEG_T_DC_SCAN);
        launch_idct(ia, s, &launches);
        mark(ctx, KPEG_T_IDCT);
        CK(cudaMemcpyAsync(h_meta, d_meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (h_meta->status & (ST_BAD_MARKER | ST_SEG_COUNT))
            break; // malformed container-level structure: more rounds will not help
        if (h_meta->changed[relay_slot(rounds)] == 0u)
            break; // the last relay round changed nothing: fixed point, results are final
        // Rare: the relay needed more rounds than were pre-issued.  Run two more at a time until a
        // round changes nothing, then redo the downstream stages.
        bool converged = false;
        const uint32_t cap = h_meta->nsub / 2u + 4u;
        while (!converged && extra_iterations < cap) {
            ++extra_iterations;
            for (int k = 0; k < 2; ++k) {
                ++rounds;
                CK(cudaMemsetAsync(&d_meta->changed[relay_slot(rounds)], 0, sizeof(uint32_t), s));
                launch_entropy_relay(ea, rounds, s, &launches);
            }
            mark(ctx, KPEG_T_ENTROPY_RELAY);
            CK(cudaMemcpyAsync(h_meta, d_meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            converged = h_meta->changed[relay_slot(rounds)] == 0u;
        }
        if (!converged)
            return fail(ctx, KPEG_ERR_NOT_CONVERGED, "speculative decode did not reach a fixed point");
        CK(cudaMemsetAsync(&d_meta->status, 0, sizeof(uint32_t), s));
        CK(cudaMemsetAsync(&d_meta->exact_samples, 0, 2 * sizeof(uint32_t), s));
        CK(cudaMemsetAsync(&d_meta->tie_records, 0, 2 * sizeof(uint32_t), s));
    }
    const uint32_t rounds_run = (uint32_t)rounds;

    ctx->last_g = g;
    ctx->have_last = true;
    if (launches_out)
        *launches_out = launches;
    if (stats) {
        stats->width = g.width;
        stats->height = g.height;
        stats->ncomp = g.ncomp;
        stats->scan_bytes = scan_len;
        stats->unstuffed_bytes = h_meta->total_kept;
        stats->segments = h_meta->total_rst + 1u;
        stats->subsequences = h_meta->nsub;
        uint32_t used_rounds = rounds_run;
        if (extra_iterations == 0u) { // rounds that still changed something
            used_rounds = 0;
            for (int r = 1; r <= (int)rounds_run && r < MAX_RELAY_ROUNDS; ++r)
                if (h_meta->changed[r])
                    used_rounds = (uint32_t)r;
        }
        stats->sync_rounds = used_rounds;
        stats->exact_samples = h_meta->tie_records; // pixels with at least one sample on the exact path
        stats->kernel_launches = launches;
    }
    return status_to_rc(ctx, h_meta->status);
}

void fill_times(kpeg_ctx *ctx, kpeg_stats *stats)
{
    if (!stats)
        return;
    for (int i = 0; i < KPEG_T_COUNT; ++i)
        stats->ms[i] = 0.0f;
    stats->ms_total = 0.0f;
    if (!ctx->profiling || ctx->nev < 2)
        return;
    for (int i = 1; i < ctx->nev; ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i - 1], ctx->ev[i]) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const int st = ctx->ev_stage[i];
        if (st >= 0 && st < KPEG_T_COUNT)
            stats->ms[st] += ms;
        stats->ms_total += ms;
    }
}

} // namespace

// ==================================================================================================
// C ABI
// ==================================================================================================

extern "C" int kpeg_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int kpeg_cuda_create(int device, kpeg_ctx **out)
{
    if (!out)
        return KPEG_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return KPEG_ERR_CUDA; // no CPU fallback: the caller gets an error
    }
    if (cudaSetDevice(device) != cudaSuccess)
        return KPEG_ERR_CUDA;
    kpeg_ctx *ctx = new (std::nothrow) kpeg_ctx();
    if (!ctx)
        return KPEG_ERR_NOMEM;
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return KPEG_ERR_CUDA;
    }
    kernels_configure();
    if (const char *sb = getenv("KPEG_SUB_BITS")) {
        const long v = strtol(sb, nullptr, 10);
        if (v >= 64 && v <= 1024 && (v & (v - 1)) == 0)
            ctx->sub_bits = (uint32_t)v;
    }
    if (const char *rr = getenv("KPEG_RELAY_ROUNDS")) {
        const long v = strtol(rr, nullptr, 10);
        if (v >= 2 && v < MAX_RELAY_ROUNDS)
            ctx->relay_rounds = (int)v;
    }
    if (cudaGetLastError() != cudaSuccess) {
        kpeg_cuda_destroy(ctx);
        return KPEG_ERR_CUDA;
    }
    *out = ctx;
    return KPEG_OK;
}

extern "C" void kpeg_cuda_destroy(kpeg_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->scan,   &ctx->words,  &ctx->seg_bit,    &ctx->tile_kept, &ctx->tile_rst, &ctx->state,
                      &ctx->used,   &ctx->seg_hint, &ctx->start_slot, &ctx->scan_tiles, &ctx->coef,    &ctx->dcdiff,   &ctx->dc,
                      &ctx->tile_carry, &ctx->pixels, &ctx->tables,  &ctx->meta,     &ctx->merged, &ctx->tie_rec, &ctx->overflow};
    for (DevBuf *b : bufs)
        if (b->p)
            cudaFree(b->p);
    PinBuf *pins[] = {&ctx->h_tables, &ctx->h_meta, &ctx->h_stage};
    for (PinBuf *b : pins)
        if (b->p)
            cudaFreeHost(b->p);
    for (int i = 0; i < MAX_EVENTS; ++i)
        if (ctx->ev[i])
            cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *kpeg_cuda_last_error(const kpeg_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" int kpeg_cuda_set_profiling(kpeg_ctx *ctx, int on)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    ctx->profiling = on != 0;
    return KPEG_OK;
}

extern "C" int kpeg_cuda_set_tuning(kpeg_ctx *ctx, int sub_bits, int relay_rounds)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    if (sub_bits > 0) {
        if (sub_bits < 64 || sub_bits > 1024 || (sub_bits & (sub_bits - 1)))
            return KPEG_ERR_ARG; // power of two: the kernels address the staged stream with shifts
        ctx->sub_bits = (uint32_t)sub_bits;
    }
    if (relay_rounds > 0) {
        if (relay_rounds < 2 || relay_rounds >= MAX_RELAY_ROUNDS)
            return KPEG_ERR_ARG;
        ctx->relay_rounds = relay_rounds;
    }
    return KPEG_OK;
}

extern "C" void *kpeg_cuda_stream(kpeg_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" void *kpeg_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void kpeg_cuda_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

extern "C" void *kpeg_cuda_device_alloc(kpeg_ctx *ctx, size_t bytes)
{
    if (!ctx)
        return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void kpeg_cuda_device_free(kpeg_ctx *ctx, void *p)
{
    if (ctx && p) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(p);
    }
}

extern "C" int kpeg_cuda_memcpy_h2d(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KPEG_OK;
}

extern "C" int kpeg_cuda_memcpy_d2h(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KPEG_OK;
}

extern "C" int kpeg_cuda_decode_device(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *d_scan, size_t scan_len,
                                       uint8_t *d_pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || !d_scan || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stats)
        memset(stats, 0, sizeof *stats);
    mark(ctx, -1);
    const int rc = run_job(ctx, plan, d_scan, scan_len, 1, d_pixels_out, stats, nullptr);
    fill_times(ctx, stats);
    return rc;
}

extern "C" int kpeg_cuda_decode(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len,
                                uint8_t *pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || !scan || !pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stats)
        memset(stats, 0, sizeof *stats);
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    TRY(ensure(ctx, ctx->scan, scan_len + 64));
    TRY(ensure(ctx, ctx->pixels, npix + 64));
    mark(ctx, -1);
    CK(cudaMemcpyAsync(ctx->scan.p, scan, scan_len, cudaMemcpyHostToDevice, ctx->stream));
    mark(ctx, KPEG_T_H2D);
    const int rc = run_job(ctx, plan, (const uint8_t *)ctx->scan.p, scan_len, 1, (uint8_t *)ctx->pixels.p, stats, nullptr);
    if (rc != KPEG_OK && rc != KPEG_ERR_STREAM)
        return rc;
    CK(cudaMemcpyAsync(pixels_out, ctx->pixels.p, npix, cudaMemcpyDeviceToHost, ctx->stream));
    mark(ctx, KPEG_T_D2H);
    CK(cudaStreamSynchronize(ctx->stream));
    fill_times(ctx, stats);
    return rc;
}

// Batch stream format: the n stuffed scans back to back, each followed by one 2-byte RSTn marker
// (FF D0+(i&7)); to K0 an image boundary is then just another restart boundary.
extern "C" size_t kpeg_batch_packed_size(int n, const size_t *scan_lens)
{
    size_t t = 0;
    for (int i = 0; i < n; ++i)
        t += scan_lens[i] + 2;
    return t;
}

extern "C" int kpeg_batch_pack(int n, const uint8_t *const *scans, const size_t *scan_lens, uint8_t *dst, size_t cap)
{
    size_t o = 0;
    for (int i = 0; i < n; ++i) {
        if (o + scan_lens[i] + 2 > cap)
            return KPEG_ERR_ARG;
        memcpy(dst + o, scans[i], scan_lens[i]);
        o += scan_lens[i];
        dst[o++] = 0xFF;
        dst[o++] = (uint8_t)(0xD0 + (i & 7));
    }
    return KPEG_OK;
}

extern "C" int kpeg_cuda_decode_batch_packed_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                                    size_t packed_len, uint8_t *d_pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || n <= 0 || !d_packed || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stats)
        memset(stats, 0, sizeof *stats);
    mark(ctx, -1);
    const int rc = run_job(ctx, plan, d_packed, packed_len, (uint32_t)n, d_pixels_out, stats, nullptr);
    fill_times(ctx, stats);
    return rc;
}

extern "C" int kpeg_cuda_decode_batch_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_scans,
                                             const uint64_t *scan_offsets, uint8_t *d_pixels_out, kpeg_stats *stats)
{
    // Device-resident scans without separators: re-pack on the device side with n small copies.
    if (!ctx || !plan || n <= 0 || !d_scans || !scan_offsets || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t total = (size_t)(scan_offsets[n] - scan_offsets[0]) + 2u * (size_t)n;
    TRY(ensure(ctx, ctx->scan, total + 64));
    TRY(ensure_pinned(ctx, ctx->h_stage, 16));
    CK(cudaStreamSynchronize(ctx->stream));
    uint8_t *sep = (uint8_t *)ctx->h_stage.p;
    for (int k = 0; k < 8; ++k) {
        sep[2 * k] = 0xFF;
        sep[2 * k + 1] = (uint8_t)(0xD0 + k);
    }
    size_t o = 0;
    for (int i = 0; i < n; ++i) {
        const size_t len = (size_t)(scan_offsets[i + 1] - scan_offsets[i]);
        CK(cudaMemcpyAsync((uint8_t *)ctx->scan.p + o, d_scans + scan_offsets[i], len, cudaMemcpyDeviceToDevice, ctx->stream));
        o += len;
        CK(cudaMemcpyAsync((uint8_t *)ctx->scan.p + o, sep + 2 * (i & 7), 2, cudaMemcpyHostToDevice, ctx->stream));
        o += 2;
    }
    return kpeg_cuda_decode_batch_packed_device(ctx, plan, n, (const uint8_t *)ctx->scan.p, total, d_pixels_out, stats);
}

extern "C" int kpeg_cuda_decode_batch(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *const *scans,
                                      const size_t *scan_lens, uint8_t *const *pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || n <= 0 || !scans || !scan_lens || !pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stats)
        memset(stats, 0, sizeof *stats);
    const size_t total = kpeg_batch_packed_size(n, scan_lens);
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    TRY(ensure_pinned(ctx, ctx->h_stage, total));
    TRY(ensure(ctx, ctx->scan, total + 64));
    TRY(ensure(ctx, ctx->pixels, npix * (size_t)n + 64));
    CK(cudaStreamSynchronize(ctx->stream));
    TRY(kpeg_batch_pack(n, scans, scan_lens, (uint8_t *)ctx->h_stage.p, ctx->h_stage.cap));
    mark(ctx, -1);
    CK(cudaMemcpyAsync(ctx->scan.p, ctx->h_stage.p, total, cudaMemcpyHostToDevice, ctx->stream));
    mark(ctx, KPEG_T_H2D);
    const int rc = run_job(ctx, plan, (const uint8_t *)ctx->scan.p, total, (uint32_t)n, (uint8_t *)ctx->pixels.p, stats, nullptr);
    if (rc != KPEG_OK && rc != KPEG_ERR_STREAM)
        return rc;
    for (int i = 0; i < n; ++i)
        CK(cudaMemcpyAsync(pixels_out[i], (uint8_t *)ctx->pixels.p + npix * (size_t)i, npix, cudaMemcpyDeviceToHost, ctx->stream));
    mark(ctx, KPEG_T_D2H);
    CK(cudaStreamSynchronize(ctx->stream));
    fill_times(ctx, stats);
    return rc;
}

extern "C" int kpeg_cuda_decode_file(kpeg_ctx *ctx, const uint8_t *file, size_t len, uint32_t flags, uint8_t *pixels_out,
                                     size_t cap, kpeg_plan *plan_out, kpeg_stats *stats)
{
    if (!ctx || !file || !pixels_out)
        return KPEG_ERR_ARG;
    kpeg_plan plan;
    size_t off = 0, slen = 0;
    const int prc = kpeg_parse_jfif(file, len, &plan, &off, &slen);
    if (prc != KPEG_OK)
        return fail(ctx, prc, prc == KPEG_ERR_UNSUPPORTED ? "unsupported JPEG coding" : "malformed JFIF container");
    plan.flags = flags;
    if (plan_out)
        *plan_out = plan;
    if ((size_t)plan.width * plan.height * plan.ncomp > cap)
        return fail(ctx, KPEG_ERR_ARG, "pixel buffer too small");
    return kpeg_cuda_decode(ctx, &plan, file + off, slen, pixels_out, stats);
}

extern "C" int kpeg_cuda_read_coefficients(kpeg_ctx *ctx, int16_t *out, size_t cap)
{
    if (!ctx || !out)
        return KPEG_ERR_ARG;
    if (!ctx->have_last)
        return fail(ctx, KPEG_ERR_ARG, "no decode has run on this context");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->last_g.total_blocks * 64u;
    if (cap < n)
        return fail(ctx, KPEG_ERR_ARG, "coefficient buffer too small");
    TRY(ensure(ctx, ctx->merged, n * 2u));
    launch_merge_dc((int16_t *)ctx->merged.p, (const int16_t *)ctx->coef.p, (const int16_t *)ctx->dc.p,
                    (const int16_t *)ctx->dcdiff.p, ctx->last_g.total_blocks, ctx->last_g.flags, ctx->stream);
    CK(cudaMemcpyAsync(out, ctx->merged.p, n * 2u, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return KPEG_OK;
}
