// kpeg_cuda.cu -- C-ABI host side of the B200 decode path (include/kpeg_cuda.h).
//
// A context owns eight LANES (CUDA stream + grow-only device scratch + pinned bookkeeping each) and
// the device tables built from a kpeg_plan.  A job is enqueued on a lane (K0..K3, kernels.cu) and
// finished later (stream sync, device status word, and -- rarely -- extra relay rounds).  Single
// decodes use lane 0; a large host-pointer batch is cut into chunks that rotate through the
// lanes so the H2D copy of one chunk, the kernels of another and the D2H copy of a third overlap
// (the end-to-end path is PCIe-bound: 3 bytes per pixel have to leave the device).
//
// There is no CPU decode in here: the only host arithmetic is table preparation (Huffman LUTs,
// AAN-prescaled quantisers, the 64 double cosines the exact IDCT path needs -- evaluated with the
// host libm exactly as the reference evaluates them, src/MCU.cpp:193).
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
    void *p = nullptr;   // what the kernels get
    size_t cap = 0;
    void *raw = nullptr; // the allocation (== p, or p - GUARD_BYTES in guard mode)
};

// KPEG_GUARD=1 (debugging aid; the pool this was developed on has no compute-sanitizer): every device buffer is
// allocated at exactly the size asked for, between two guard areas of a known byte, which are checked after each job.
constexpr size_t GUARD_BYTES = 256;
constexpr int GUARD_PATTERN = 0xA5;

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
};

constexpr int MAX_EVENTS = 192;
constexpr int NLANES = 8;

struct Copy {
    void *dst;
    const void *src;
    size_t bytes;
    int peer = -1; // >= 0: dst is device memory of that device (peer copy), otherwise host memory
};

// One in-flight job between job_enqueue and job_finish.
struct Job {
    bool active = false;
    JobGeom g = {};
    EntropyArgs ea = {};
    ExpandArgs xa = {};
    IdctArgs ia = {};
    int rounds = 0;
    uint32_t launches = 0;
    uint32_t extra_iterations = 0;
    bool use_records = true;
    bool loop_used = true; // relay rounds 2.. ran as the cooperative device-side loop (DevMeta::relay_rounds is meaningful)
    size_t scan_len = 0;
    std::vector<Copy> d2h; // result copies to (re)issue after the downstream stages
    bool entropy_only = false; // stop after K2: the tiles are the result (one scan of a frame coded one scan per component)
};

struct Lane {
    cudaStream_t stream = nullptr;
    DevBuf scan, words, seg_bit, tile_kept, tile_rst, cls, state, work, seg_hint, start_slot, scan_tiles;
    DevBuf coef, dcdiff, dc, tiles, tie_rec, tie_cnt, pixels, meta, rec, nrec, rec_alt, strip_sub, dcs, dcpre, scan_tiles_dcs;
    PinBuf h_meta;
    // host-buffer batches: pinned staging of the small scans + the separator positions, and their device copy
    PinBuf h_stage, h_ends;
    DevBuf d_ends, d_stage;
    cudaEvent_t ev[MAX_EVENTS] = {};
    int ev_stage[MAX_EVENTS] = {};
    int nev = 0;
    Job job;
};

} // namespace

struct kpeg_ctx {
    int device = 0;
    std::string err;
    bool profiling = false;
    uint32_t sub_bits = 512;
    // no size asked for (set_tuning / KPEG_SUB_BITS): 1024 bits for streams whose restart segments are long (>= 4 Mbit on
    // average: large images without restart markers) and whose blocks are long (>= 96 bits on average: high quality) -- the relay then converges in two rounds instead of five and the
    // cooperative loop, which holds SM slots while it waits out one serial subsequence decode per round, is a third as
    // long: +3 % on the 4K workload), 512 otherwise (segment boundaries are synchronisation points: 512 is 10-17 % faster
    // on the 512x512 batch and on the image with a restart interval per MCU row)
    bool sub_bits_auto = true;
    int relay_rounds = 8;
    bool use_records = true; // final pass = record expansion (KPEG_NO_RECORDS=1: Huffman final pass)
    int split_parts = 4;     // concurrent jobs a device-resident batch is cut into (KPEG_SPLIT)
    int host_chunks = 8;     // pipeline depth of a host-pointer batch (KPEG_HOST_CHUNKS)
    int submit_parts = 1;    // jobs one deferred submission is cut into (KPEG_SUBMIT_SPLIT)
    // bands a large restart-marked image from host memory is cut into, one per lane (KPEG_BANDS; 1 = off)
    int band_parts = 4;
    DevBuf band_counts;     // RSTn markers counted per band (decode_banded)
    PinBuf h_band_counts;
    // result copies of the lanes leave one after another (each waits for the one enqueued before it): copies to the
    // host that run at the same time share the link AND slow each other down (measured: four concurrent 200 MB copies
    // take twice as long as the same four back to back); KPEG_D2H_CHAIN=0 turns the ordering off
    // relay round 2 as a launch of its own, rounds 3.. in the cooperative loop (KPEG_RELAY_ROUND2_WIDE=1).  Off: measured, the
    // loop takes as long without round 2 (every round costs the latency of one serial subsequence decode, ~35 us, whatever
    // the length of its list) and the extra launch costs 2 % of the throughput
    bool relay_round2_wide = false;
    bool d2h_chain = true;
    bool d2h_chain_armed = false;
    cudaEvent_t d2h_done = nullptr;
    Lane lane[NLANES];

    DevBuf tables, merged;
    DevBuf scan_tiles[3]; // frames coded one scan per component: the components' tiles before they are interleaved
    PinBuf h_tables, h_sep;
    cudaEvent_t tables_ready = nullptr;
    kpeg_plan plan_cached;
    bool have_plan = false;

    int last_lane = -1; // lane of the last finished job (kpeg_cuda_read_coefficients)
    JobGeom last_g = {};

    // deferred submissions (kpeg_cuda_submit_* / kpeg_cuda_wait): lanes are handed out round-robin and a
    // lane's previous job is finished only when the lane comes up again, so several batches are in flight
    bool counted = false; // registered with kernels_context_created
    bool guard = false;   // KPEG_GUARD=1
    bool plan_fits_records = true;
    int next_lane = 0;
    int deferred_rc = KPEG_OK;
    std::string deferred_err;
    kpeg_stats deferred_stats = {};
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

#define TRY(expr)                                                                                                      \
    do {                                                                                                               \
        int rc_ = (expr);                                                                                              \
        if (rc_ != KPEG_OK)                                                                                            \
            return rc_;                                                                                                \
    } while (0)

void dev_free(DevBuf &b)
{
    if (b.raw)
        cudaFree(b.raw);
    b.raw = b.p = nullptr;
    b.cap = 0;
}

int ensure(kpeg_ctx *ctx, cudaStream_t stream, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return KPEG_OK;
    if (b.raw) {
        CK(cudaStreamSynchronize(stream));
        dev_free(b);
    }
    const bool guard = ctx->guard;
    const size_t want = guard ? ((bytes + 15u) & ~(size_t)15u) : bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.raw, want + (guard ? 2 * GUARD_BYTES : 0));
    if (e != cudaSuccess) {
        b.raw = nullptr;
        return fail(ctx, KPEG_ERR_NOMEM, "cudaMalloc", e);
    }
    b.p = b.raw;
    if (guard) {
        b.p = (uint8_t *)b.raw + GUARD_BYTES;
        CK(cudaMemsetAsync(b.raw, GUARD_PATTERN, GUARD_BYTES, stream));
        CK(cudaMemsetAsync((uint8_t *)b.p + want, GUARD_PATTERN, GUARD_BYTES, stream));
    }
    b.cap = want;
    return KPEG_OK;
}

int ensure_pinned(kpeg_ctx *ctx, cudaStream_t stream, PinBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return KPEG_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(stream));
        CK(cudaFreeHost(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, KPEG_ERR_NOMEM, "cudaMallocHost", e);
    }
    b.cap = want;
    return KPEG_OK;
}

// ---- plan -> device tables (shared by all lanes) ---------------------------------------------------
// The device tables (Huffman LUTs, quantisers) depend on the components' table selectors and the tables themselves,
// not on the image's dimensions, restart interval or flags: jobs that differ only in those -- the restart-interval
// bands of one image, say -- share the tables and can be in flight together.
bool same_tables(const kpeg_plan *a, const kpeg_plan *b)
{
    kpeg_plan x = *a, y = *b;
    x.width = y.width = x.height = y.height = 0;
    x.restart_interval = y.restart_interval = 0;
    x.flags = y.flags = 0;
    return memcmp(&x, &y, sizeof x) == 0;
}

int upload_plan(kpeg_ctx *ctx, const kpeg_plan *pl)
{
    if (ctx->have_plan && same_tables(&ctx->plan_cached, pl))
        return KPEG_OK;
    // a new plan: nothing may still be reading the old tables
    for (Lane &L : ctx->lane)
        CK(cudaStreamSynchronize(L.stream));
    cudaStream_t s0 = ctx->lane[0].stream;
    TRY(ensure_pinned(ctx, s0, ctx->h_tables, sizeof(DeviceTables)));
    TRY(ensure(ctx, s0, ctx->tables, sizeof(DeviceTables)));
    const char *why = nullptr;
    const int rc = build_device_tables(pl, (DeviceTables *)ctx->h_tables.p, &why);
    if (rc != KPEG_OK)
        return fail(ctx, rc, why);
    CK(cudaMemcpyAsync(ctx->tables.p, ctx->h_tables.p, sizeof(DeviceTables), cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(ctx->tables_ready, s0));
    for (int i = 1; i < NLANES; ++i)
        CK(cudaStreamWaitEvent(ctx->lane[i].stream, ctx->tables_ready, 0));
    ctx->plan_cached = *pl;
    ctx->have_plan = true;
    // symbol records keep 11 magnitude bits (all baseline JPEG has: DC categories 0..11, AC 0..10, T.81
    // F.1.2.1.2 / F.1.2.2.1); a table that declares wider ones is decoded with the Huffman final pass
    ctx->plan_fits_records = true;
    for (int cls = 0; cls < 2; ++cls)
        for (int id = 0; id < 4; ++id) {
            if (!pl->ht_present[cls][id])
                continue;
            int nsym = 0;
            for (int L = 0; L < 16; ++L)
                nsym += pl->ht[cls][id].counts[L];
            for (int i = 0; i < nsym && i < 256; ++i)
                if ((pl->ht[cls][id].symbols[i] & 15) > 11)
                    ctx->plan_fits_records = false;
        }
    return KPEG_OK;
}

// ---- profiling: an event after every stage; mark(.., -1) opens a window on a lane -------------------
void mark(kpeg_ctx *ctx, Lane &L, int stage)
{
    if (!ctx->profiling)
        return;
    if (stage < 0)
        L.nev = 0;
    if (L.nev >= MAX_EVENTS)
        return;
    if (!L.ev[L.nev])
        cudaEventCreate(&L.ev[L.nev]);
    cudaEventRecord(L.ev[L.nev], L.stream);
    L.ev_stage[L.nev] = stage;
    ++L.nev;
}

void add_times(kpeg_ctx *ctx, Lane &L, kpeg_stats *stats)
{
    if (!stats || !ctx->profiling || L.nev < 2)
        return;
    static const bool timeline = getenv("KPEG_TIMELINE") != nullptr; // development: when every stage of every lane ended, relative to lane 0's window
    if (timeline && ctx->lane[0].ev[0]) {
        fprintf(stderr, "[kpeg timeline] lane %d:", (int)(&L - ctx->lane));
        for (int i = 0; i < L.nev; ++i) {
            float at = 0.0f;
            if (cudaEventElapsedTime(&at, ctx->lane[0].ev[0], L.ev[i]) != cudaSuccess)
                cudaGetLastError();
            fprintf(stderr, " %d@%.2f", L.ev_stage[i], at);
        }
        fprintf(stderr, "\n");
    }
    for (int i = 1; i < L.nev; ++i) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, L.ev[i - 1], L.ev[i]) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const int st = L.ev_stage[i];
        if (st >= 0 && st < KPEG_T_COUNT) {
            stats->ms[st] += ms;
            stats->ms_total += ms;
        }
    }
    L.nev = 0;
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

// everything downstream of the relay + the result copies + the bookkeeping read-back
int enqueue_downstream(kpeg_ctx *ctx, Lane &L)
{
    Job &J = L.job;
    cudaStream_t s = L.stream;
    launch_entropy_scan(J.ea, s, &J.launches);
    mark(ctx, L, KPEG_T_ENTROPY_SCAN);
    if (J.use_records) {
        launch_expand(J.xa, s, &J.launches); // K2
        mark(ctx, L, KPEG_T_ENTROPY_WRITE);
    } else {
        // fallback: the Huffman final pass writes a plain coefficient matrix (every slot of every block: no zero-fill)
        // and the DC differences; DC prediction and the conversion into K3's tiles are separate small kernels
        const size_t coef_bytes = (size_t)J.g.total_blocks * 128u;
        TRY(ensure(ctx, s, L.coef, coef_bytes + 256));
        TRY(ensure(ctx, s, L.dcdiff, (size_t)J.g.total_blocks * 2u + 16));
        TRY(ensure(ctx, s, L.dc, (size_t)J.g.total_blocks * 2u + 16));
        J.ea.coef = (int16_t *)L.coef.p;
        J.ea.dcdiff = (int16_t *)L.dcdiff.p;
        launch_entropy_write(J.ea, s, &J.launches);
        mark(ctx, L, KPEG_T_ENTROPY_WRITE);
        launch_dc_integrate(J.g, (const int16_t *)L.dcdiff.p, (int16_t *)L.dc.p, s, &J.launches);
        launch_tiles_from_matrix(J.g, (const int16_t *)L.coef.p, (const int16_t *)L.dc.p, J.xa.tiles, s, &J.launches);
        mark(ctx, L, KPEG_T_DC_SCAN);
    }
    if (!J.entropy_only) {
        CK(launch_idct(J.ia, s, &J.launches));
        mark(ctx, L, KPEG_T_IDCT);
    }
    const bool chain = ctx->d2h_chain && ctx->d2h_done && !J.d2h.empty();
    if (chain && ctx->d2h_chain_armed)
        CK(cudaStreamWaitEvent(s, ctx->d2h_done, 0));
    for (const Copy &c : J.d2h) {
        if (c.peer >= 0)
            CK(cudaMemcpyPeerAsync(c.dst, c.peer, c.src, ctx->device, c.bytes, s));
        else
            CK(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, s));
    }
    if (chain) {
        CK(cudaEventRecord(ctx->d2h_done, s));
        ctx->d2h_chain_armed = true;
    }
    if (!J.d2h.empty())
        mark(ctx, L, KPEG_T_D2H);
    CK(cudaMemcpyAsync(L.h_meta.p, L.meta.p, sizeof(DevMeta), cudaMemcpyDeviceToHost, s));
    return KPEG_OK;
}

int job_finish(kpeg_ctx *ctx, int li, kpeg_stats *stats);

// complete a deferred job; its outcome is reported by the next kpeg_cuda_wait
void finish_deferred(kpeg_ctx *ctx, int li)
{
    const int rc = job_finish(ctx, li, &ctx->deferred_stats);
    if (rc != KPEG_OK && ctx->deferred_rc == KPEG_OK) {
        ctx->deferred_rc = rc;
        ctx->deferred_err = ctx->err;
    }
}

// Enqueue the whole device pipeline for one job whose stuffed bytes are (or will be, in stream
// order) at d_scan.  Optimistic: the stages downstream of the relay are issued before it is known
// whether the pre-issued relay rounds reached the fixed point; job_finish repairs that if not.
int job_enqueue(kpeg_ctx *ctx, int li, const kpeg_plan *pl, const uint8_t *d_scan, size_t scan_len, uint32_t nimages,
                uint8_t *d_pixels, std::vector<Copy> d2h, void *tiles_only = nullptr)
{
    Lane &L = ctx->lane[li];
    Job &J = L.job;
    cudaStream_t s = L.stream;
    if (J.active) // a deferred job still owns this lane's scratch: complete it first
        finish_deferred(ctx, li);
    if (ctx->have_plan && !same_tables(&ctx->plan_cached, pl))
        for (int i = 0; i < NLANES; ++i) // the tables are shared: nothing deferred may outlive them
            if (ctx->lane[i].job.active)
                finish_deferred(ctx, i);
    JobGeom g;
    const char *why = nullptr;
    uint32_t sub_bits = ctx->sub_bits;
    if (ctx->sub_bits_auto) {
        const uint64_t segs = (uint64_t)nimages * (pl->restart_interval ? ((uint64_t)((pl->width + 7u) / 8u) * ((pl->height + 7u) / 8u) + pl->restart_interval - 1u) / pl->restart_interval : 1u);
        // ... and whose blocks are long (>= 96 bits on average, i.e. high quality settings): streams of short blocks
        // re-synchronise within a few hundred bits, their relay needs one or two rounds at 512 bits already (4K q50:
        // 145 Gpixel/s at 512 bits, 139 at 1024; 1080p q90: equal; 4K q95: 80 vs 87)
        const uint64_t blocks = (uint64_t)nimages * ((pl->width + 7u) / 8u) * ((pl->height + 7u) / 8u) * pl->ncomp;
        const uint64_t bits = (uint64_t)scan_len * 8u;
        sub_bits = (bits >= segs * ((uint64_t)4 << 20) && bits >= blocks * 96u) ? 1024u : 512u;
    }
    int rc = make_job_geom(pl, nimages, sub_bits, &g, &why);
    if (rc != KPEG_OK)
        return fail(ctx, rc, why);
    if (scan_len == 0 || scan_len >= (1ull << 29))
        return fail(ctx, KPEG_ERR_ARG, "entropy-coded segment must be 1 byte .. 512 MiB");
    TRY(upload_plan(ctx, pl));

    const uint32_t S = (uint32_t)scan_len;
    const uint32_t ntiles = (S + 15u + UNSTUFF_TILE - 1) / UNSTUFF_TILE; // + 15: chunks are cut on address boundaries
    const uint32_t nsub_max = (uint32_t)(((uint64_t)S * 8u + g.sub_bits - 1u) / g.sub_bits) + 1u;
    const uint32_t total_mcus = g.nimages * g.mcus_per_image;
    const uint32_t nstrips = (total_mcus + IDCT_MCUS_PER_CTA - 1) / IDCT_MCUS_PER_CTA;
    const size_t words_bytes = ((size_t)S + 3u) / 4u * 4u + 64u;

    TRY(ensure(ctx, s, L.words, words_bytes));
    TRY(ensure(ctx, s, L.seg_bit, ((size_t)g.nseg + 2u) * 4u));
    TRY(ensure(ctx, s, L.tile_kept, (size_t)ntiles * 4u));
    TRY(ensure(ctx, s, L.tile_rst, (size_t)ntiles * 4u));
    TRY(ensure(ctx, s, L.cls, (size_t)ntiles * UNSTUFF_THREADS * 4u));
    TRY(ensure(ctx, s, L.state, (size_t)nsub_max * sizeof(SubState)));
    TRY(ensure(ctx, s, L.work, (size_t)nsub_max * 2u * sizeof(uint32_t)));
    TRY(ensure(ctx, s, L.seg_hint, (size_t)nsub_max * 4u));
    TRY(ensure(ctx, s, L.start_slot, (size_t)nsub_max * 4u));
    TRY(ensure(ctx, s, L.scan_tiles, ((size_t)nsub_max / 1024u + 2u) * sizeof(uint2)));
    TRY(ensure(ctx, s, L.strip_sub, ((size_t)nstrips + 1u) * sizeof(uint32_t)));
    TRY(ensure(ctx, s, L.dcs, (size_t)nsub_max * sizeof(long long)));
    TRY(ensure(ctx, s, L.dcpre, (size_t)nsub_max * sizeof(long long)));
    TRY(ensure(ctx, s, L.scan_tiles_dcs, ((size_t)nsub_max / 1024u + 2u) * sizeof(long long)));
    TRY(ensure(ctx, s, L.tiles, (size_t)nstrips * IDCT_MCUS_PER_CTA * g.ncomp * 128u + 256u));
    TRY(ensure(ctx, s, L.tie_rec, (size_t)nstrips * IDCT_TIE_LIST_CAP * sizeof(uint4)));
    TRY(ensure(ctx, s, L.tie_cnt, (size_t)nstrips * sizeof(uint32_t)));
    TRY(ensure(ctx, s, L.meta, sizeof(DevMeta)));
    TRY(ensure_pinned(ctx, s, L.h_meta, sizeof(DevMeta)));
    // records: one per value-carrying symbol.  3/8 of the subsequence's bits covers every table whose value-carrying
    // symbols take at least 3 bits (code >= 2 bits + value >= 1 bit: all of Annex K); a stream that needs more sets
    // ST_REC_OVERFLOW and is redone with the Huffman final pass
    const uint32_t rec_kmax = (g.sub_bits * 3u / 8u + 7u) & ~7u;
    const uint32_t rec_alt_cap = nsub_max / 4u + 1024u; // beyond it a redone subsequence rewrites its interleaved column
    if (ctx->use_records) {
        TRY(ensure(ctx, s, L.rec, (size_t)rec_kmax * ((size_t)nsub_max / 32u + 1u) * 32u * sizeof(uint32_t)));
        TRY(ensure(ctx, s, L.nrec, (size_t)nsub_max * sizeof(uint32_t)));
        TRY(ensure(ctx, s, L.rec_alt, (size_t)rec_alt_cap * rec_kmax * sizeof(uint32_t)));
    }

    J = Job();
    J.use_records = ctx->use_records && ctx->plan_fits_records;
    J.active = true;
    J.g = g;
    J.scan_len = scan_len;
    J.d2h = std::move(d2h);
    J.entropy_only = tiles_only != nullptr;
    void *const tiles = tiles_only ? tiles_only : L.tiles.p;
    DevMeta *d_meta = (DevMeta *)L.meta.p;

    CK(cudaMemsetAsync(d_meta, 0, sizeof(DevMeta), s));
    mark(ctx, L, KPEG_T_MEMSET);

    UnstuffArgs ua;
    ua.scan = d_scan;
    ua.scan_len = S;
    ua.ntiles = ntiles;
    ua.tile_kept = (uint32_t *)L.tile_kept.p;
    ua.tile_rst = (uint32_t *)L.tile_rst.p;
    ua.cls = (uint32_t *)L.cls.p;
    ua.words = (uint8_t *)L.words.p;
    ua.seg_bit = (uint32_t *)L.seg_bit.p;
    ua.nseg = g.nseg;
    ua.meta = d_meta;
    launch_unstuff(ua, g.sub_bits, s, &J.launches);
    mark(ctx, L, KPEG_T_UNSTUFF);

    EntropyArgs &ea = J.ea;
    ea.words = (const uint32_t *)L.words.p;
    ea.seg_bit = (const uint32_t *)L.seg_bit.p;
    ea.tables = (const DeviceTables *)ctx->tables.p;
    ea.meta = d_meta;
    ea.state = (SubState *)L.state.p;
    ea.worklist[0] = (uint32_t *)L.work.p;
    ea.worklist[1] = (uint32_t *)L.work.p + nsub_max;
    ea.seg_hint = (uint32_t *)L.seg_hint.p;
    ea.start_slot = (uint32_t *)L.start_slot.p;
    ea.scan_tiles = (uint2 *)L.scan_tiles.p;
    ea.scan_tiles_dcs = (long long *)L.scan_tiles_dcs.p;
    ea.dcs = (long long *)L.dcs.p;
    ea.dcpre = (long long *)L.dcpre.p;
    ea.rec = J.use_records ? (uint32_t *)L.rec.p : nullptr;
    ea.nrec = (uint32_t *)L.nrec.p;
    ea.rec_kmax = rec_kmax;
    ea.rec_alt = (uint32_t *)L.rec_alt.p;
    ea.rec_alt_cap = rec_alt_cap;
    ea.coef = nullptr; // only the Huffman final pass writes coefficients to global memory (enqueue_downstream)
    ea.dcdiff = nullptr;
    ea.strip_sub = (uint32_t *)L.strip_sub.p;
    ea.strip_slots = k3_strip_slots(g.ncomp);
    ea.nstrips = nstrips;
    ea.nsub_max = nsub_max;
    ea.g = g;

    ExpandArgs &xa = J.xa;
    xa.rec = ea.rec;
    xa.nrec = ea.nrec;
    xa.rec_alt = ea.rec_alt;
    xa.rec_kmax = rec_kmax;
    xa.start_slot = ea.start_slot;
    xa.strip_sub = ea.strip_sub;
    xa.dcpre = ea.dcpre;
    xa.tiles = tiles;
    xa.nstrips = nstrips;
    xa.meta = d_meta;
    xa.g = g;

    IdctArgs &ia = J.ia;
    ia.tiles = tiles;
    ia.nstrips = nstrips;
    ia.tables = (const DeviceTables *)ctx->tables.p;
    ia.pixels = d_pixels;
    ia.tie_rec = (uint4 *)L.tie_rec.p;
    ia.tie_cnt = (uint32_t *)L.tie_cnt.p;
    ia.meta = d_meta;
    ia.g = g;

    launch_entropy_cold(ea, s, &J.launches);
    mark(ctx, L, KPEG_T_ENTROPY_COLD);
    J.rounds = ctx->relay_rounds < 2 ? 2 : (ctx->relay_rounds > MAX_RELAY_ROUNDS - 1 ? MAX_RELAY_ROUNDS - 1 : ctx->relay_rounds);
    launch_entropy_relay(ea, 1, s, &J.launches);
    mark(ctx, L, KPEG_T_ENTROPY_RELAY);
    // (experiment knob, off: round 2 -- 6 % of the subsequences at 4K q95 -- as a wide launch of its own)
    if (ctx->relay_round2_wide)
        launch_entropy_relay(ea, 2, s, &J.launches);
    const int loop_first = ctx->relay_round2_wide ? 3 : 2;
    // the later rounds (0.5 %, 0.04 %, ...) in one cooperative launch (device-side loop, stops at the fixed point); the
    // cap only bounds the loop -- a stream that needs more is finished by job_finish
    J.rounds = MAX_RELAY_ROUNDS - 2;
    if (launch_entropy_relay_loop(ea, loop_first, J.rounds, s, &J.launches) != cudaSuccess) {
        // the driver refused the cooperative launch (MPS / a partitioned device): the same rounds as separate
        // launches; job_finish adds more if these do not reach the fixed point
        J.loop_used = false;
        J.rounds = std::max(loop_first, std::min(ctx->relay_rounds, MAX_RELAY_ROUNDS - 1));
        for (int r = loop_first; r <= J.rounds; ++r)
            launch_entropy_relay(ea, r, s, &J.launches);
    }
    mark(ctx, L, KPEG_T_RELAY_SPARSE);
    return enqueue_downstream(ctx, L);
}

struct NamedBuf {
    const char *name;
    DevBuf *buf;
};

// guard mode: the bytes on either side of every device buffer of the lane must still hold the pattern
int check_guards(kpeg_ctx *ctx, Lane &L)
{
    NamedBuf bufs[] = {{"cls", &L.cls},           {"scan", &L.scan},         {"words", &L.words},       {"seg_bit", &L.seg_bit},
                       {"tile_kept", &L.tile_kept}, {"tile_rst", &L.tile_rst}, {"state", &L.state},       {"work", &L.work},
                       {"seg_hint", &L.seg_hint}, {"start_slot", &L.start_slot}, {"scan_tiles", &L.scan_tiles},
                       {"coef", &L.coef},         {"dcdiff", &L.dcdiff},     {"tiles", &L.tiles},       {"tie_rec", &L.tie_rec},   {"tie_cnt", &L.tie_cnt},     {"strip_sub", &L.strip_sub}, {"dc", &L.dc}, {"dcs", &L.dcs}, {"dcpre", &L.dcpre}, {"scan_tiles_dcs", &L.scan_tiles_dcs}, {"d_ends", &L.d_ends},
                       {"pixels", &L.pixels},     {"meta", &L.meta},
                       {"rec", &L.rec},           {"nrec", &L.nrec},         {"rec_alt", &L.rec_alt},   {"tables", &ctx->tables},
                       {"d_stage", &L.d_stage},   {"scan_tiles[0]", &ctx->scan_tiles[0]}, {"scan_tiles[1]", &ctx->scan_tiles[1]},
                       {"scan_tiles[2]", &ctx->scan_tiles[2]}, {"band_counts", &ctx->band_counts}, {"merged", &ctx->merged}};
    uint8_t host[2 * GUARD_BYTES];
    for (const NamedBuf &nb : bufs) {
        const DevBuf &b = *nb.buf;
        if (!b.raw || b.raw == b.p)
            continue;
        CK(cudaMemcpy(host, b.raw, GUARD_BYTES, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(host + GUARD_BYTES, (const uint8_t *)b.p + b.cap, GUARD_BYTES, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < 2 * GUARD_BYTES; ++i)
            if (host[i] != (uint8_t)GUARD_PATTERN) {
                char msg[160];
                snprintf(msg, sizeof msg, "guard bytes %s device buffer '%s' (%zu bytes) were overwritten at offset %zu",
                         i < GUARD_BYTES ? "before" : "after", nb.name, b.cap, i % GUARD_BYTES);
                return fail(ctx, KPEG_ERR_CUDA, msg);
            }
    }
    return KPEG_OK;
}

// Wait for the job, check the device status word and -- rarely -- run more relay rounds and redo the
// downstream stages.  Adds the job's figures to *stats (which the caller zeroed).
int job_finish(kpeg_ctx *ctx, int li, kpeg_stats *stats)
{
    Lane &L = ctx->lane[li];
    Job &J = L.job;
    if (!J.active)
        return KPEG_OK;
    J.active = false;
    cudaStream_t s = L.stream;
    DevMeta *d_meta = (DevMeta *)L.meta.p;
    DevMeta *h_meta = (DevMeta *)L.h_meta.p;
    for (;;) {
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (h_meta->status & (ST_BAD_MARKER | ST_SEG_COUNT))
            break; // malformed container-level structure: more rounds will not help
        // the loop's grid barrier gave up (its CTAs were not co-resident: another process on the device): rounds up
        // to relay_rounds are complete, the next one may be partly done and is repeated on its partly filled list
        const bool timed_out = (h_meta->status & ST_RELAY_TIMEOUT) != 0u;
        if (J.extra_iterations == 0u && J.loop_used)
            J.rounds = (int)std::max<uint32_t>(h_meta->relay_rounds, 1u); // last round the device loop completed
        const bool converged_now = !timed_out && h_meta->changed[relay_slot(J.rounds)] == 0u;
        if (converged_now && J.use_records && (h_meta->status & ST_REC_OVERFLOW)) {
            // a subsequence held more symbols than the record list: redo the final pass the Huffman way
            J.use_records = false;
            CK(cudaMemsetAsync(&d_meta->status, 0, sizeof(uint32_t), s));
            CK(cudaMemsetAsync(&d_meta->exact_samples, 0, 2 * sizeof(uint32_t), s));
            TRY(enqueue_downstream(ctx, L));
            continue;
        }
        if (converged_now)
            break; // the last relay round changed nothing: fixed point, results are final
        // Rare: the relay needed more rounds than were pre-issued.  Run two more at a time until a
        // round changes nothing, then redo the downstream stages.
        bool converged = false;
        bool repeat = timed_out; // the first round issued here continues a partly filled list: its count is kept
        const uint32_t cap = h_meta->nsub / 2u + 4u;
        while (!converged && J.extra_iterations < cap) {
            ++J.extra_iterations;
            for (int k = 0; k < 2; ++k) {
                ++J.rounds;
                if (!repeat)
                    CK(cudaMemsetAsync(&d_meta->changed[relay_slot(J.rounds)], 0, sizeof(uint32_t), s));
                repeat = false;
                launch_entropy_relay(J.ea, J.rounds, s, &J.launches);
            }
            mark(ctx, L, KPEG_T_RELAY_SPARSE);
            CK(cudaMemcpyAsync(h_meta, d_meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            converged = h_meta->changed[relay_slot(J.rounds)] == 0u;
        }
        if (!converged)
            return fail(ctx, KPEG_ERR_NOT_CONVERGED, "speculative decode did not reach a fixed point");
        CK(cudaMemsetAsync(&d_meta->status, 0, sizeof(uint32_t), s));
        CK(cudaMemsetAsync(&d_meta->exact_samples, 0, 2 * sizeof(uint32_t), s));
        TRY(enqueue_downstream(ctx, L));
    }
    if (ctx->guard)
        TRY(check_guards(ctx, L));
    ctx->last_lane = li;
    ctx->last_g = J.g;
    if (stats) {
        stats->width = J.g.width;
        stats->height = J.g.height;
        stats->ncomp = J.g.ncomp;
        stats->scan_bytes += J.scan_len;
        stats->unstuffed_bytes += h_meta->total_kept;
        stats->segments += h_meta->total_rst + 1u;
        stats->subsequences += h_meta->nsub;
        uint32_t used_rounds = (uint32_t)J.rounds;
        if (J.extra_iterations == 0u) { // rounds that still changed something
            used_rounds = 0;
            for (int r = 1; r <= J.rounds && r < MAX_RELAY_ROUNDS; ++r)
                if (h_meta->changed[r])
                    used_rounds = (uint32_t)r;
        }
        stats->sync_rounds = std::max(stats->sync_rounds, used_rounds);
        stats->exact_samples += h_meta->exact_samples; // IDCT samples re-evaluated in the reference's operation order
        stats->kernel_launches += J.launches;
        add_times(ctx, L, stats);
    }
    return status_to_rc(ctx, h_meta->status & ~(ST_REC_OVERFLOW | ST_RELAY_TIMEOUT));
}

void zero_stats(kpeg_stats *stats)
{
    if (stats)
        memset(stats, 0, sizeof *stats);
}

// One large restart-marked image from host memory: bands of whole MCU rows, each a complete image of the same width
// and tables, go through the lanes one band per lane, so the copy in of one band, the kernels of another and the copy
// out of a third overlap.  The scan is cut at BYTE positions (the first RSTn marker at or after k/parts of its length: a
// memchr over a few hundred bytes); how many restart intervals -- hence MCU rows -- each band holds is counted on the GPU
// from the band's own bytes once they are there (count_restart_markers_kernel: 10 us for 24 MB), so the host never walks
// the scan.  (Finding cut points at given ROWS instead is a marker walk over the whole scan, 12 ms for 96 MB, more
// than the overlap gains: profiles/README.md.)  Needs a restart interval of whole MCU rows.  Synchronous.
// KPEG_ERR_UNSUPPORTED when the image does not qualify (nothing has been enqueued then).
int decode_banded(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len, uint8_t *pixels_out, kpeg_stats *stats)
{
    const uint32_t mx = (plan->width + 7u) / 8u, my = (plan->height + 7u) / 8u, ri = plan->restart_interval;
    if (ri == 0 || ri % mx != 0 || scan_len < (1u << 20))
        return KPEG_ERR_UNSUPPORTED;
    const uint32_t rows_per_interval = ri / mx;
    const uint32_t parts = (uint32_t)std::min(ctx->band_parts, NLANES);
    if (parts < 2 || my < parts * rows_per_interval * 2u)
        return KPEG_ERR_UNSUPPORTED;
    // cut points: the first RSTn marker at or after k / parts of the scan
    size_t begin[NLANES + 1], end[NLANES];
    begin[0] = 0;
    for (uint32_t k = 1; k < parts; ++k) {
        const uint8_t *p = scan + std::max(scan_len * k / parts, begin[k - 1] + 1), *stop = scan + scan_len;
        const uint8_t *cut = nullptr;
        while (p + 1 < stop) {
            const uint8_t *q = (const uint8_t *)memchr(p, 0xFF, (size_t)(stop - 1 - p));
            if (!q)
                break;
            if ((q[1] & 0xF8u) == 0xD0u) {
                cut = q;
                break;
            }
            p = q + 1;
        }
        if (!cut)
            return KPEG_ERR_UNSUPPORTED; // no marker in the rest of the scan: not the stream the plan describes; decoded whole
        end[k - 1] = (size_t)(cut - scan);
        begin[k] = end[k - 1] + 2u;
    }
    end[parts - 1] = scan_len;
    for (int i = 0; i < NLANES; ++i)
        if (ctx->lane[i].job.active)
            finish_deferred(ctx, i);
    // copies in + marker counts, every band on its own lane
    TRY(ensure(ctx, ctx->lane[0].stream, ctx->band_counts, NLANES * sizeof(uint32_t)));
    TRY(ensure_pinned(ctx, ctx->lane[0].stream, ctx->h_band_counts, NLANES * sizeof(uint32_t)));
    uint32_t *d_counts = (uint32_t *)ctx->band_counts.p, *h_counts = (uint32_t *)ctx->h_band_counts.p;
    for (uint32_t k = 0; k < parts; ++k) {
        Lane &L = ctx->lane[k];
        const size_t blen = end[k] - begin[k];
        TRY(ensure(ctx, L.stream, L.scan, blen + 64));
        mark(ctx, L, -1);
        CK(cudaMemcpyAsync(L.scan.p, scan + begin[k], blen, cudaMemcpyHostToDevice, L.stream));
        if (k + 1 < parts) { // the last band holds whatever rows are left
            CK(cudaMemsetAsync(d_counts + k, 0, sizeof(uint32_t), L.stream));
            launch_count_restart_markers((const uint8_t *)L.scan.p, (uint32_t)blen, d_counts + k, L.stream);
            CK(cudaMemcpyAsync(h_counts + k, d_counts + k, sizeof(uint32_t), cudaMemcpyDeviceToHost, L.stream));
        }
        mark(ctx, L, KPEG_T_H2D);
    }
    const size_t row_bytes = (size_t)plan->width * plan->ncomp;
    int rc = KPEG_OK, used[NLANES], nused = 0;
    uint32_t row0 = 0; // first MCU row of the band
    for (uint32_t k = 0; k < parts && rc == KPEG_OK; ++k) {
        Lane &L = ctx->lane[k];
        uint32_t rows = my - row0;
        if (k + 1 < parts) {
            CK(cudaStreamSynchronize(L.stream)); // the band is on the device and its markers are counted
            const uint64_t r = ((uint64_t)h_counts[k] + 1u) * rows_per_interval;
            if (r >= rows) { // more intervals than the frame has rows for: corrupt stream
                rc = fail(ctx, KPEG_ERR_STREAM, "restart markers do not match the restart interval");
                break;
            }
            rows = (uint32_t)r;
        }
        const uint32_t y0 = row0 * 8u, y1 = std::min<uint32_t>((row0 + rows) * 8u, plan->height);
        kpeg_plan bp = *plan;
        bp.height = (uint16_t)(y1 - y0);
        const size_t bpix = row_bytes * (y1 - y0);
        rc = ensure(ctx, L.stream, L.pixels, bpix + 64);
        if (rc != KPEG_OK)
            break;
        rc = job_enqueue(ctx, (int)k, &bp, (const uint8_t *)L.scan.p, end[k] - begin[k], 1, (uint8_t *)L.pixels.p,
                         {Copy{pixels_out + (size_t)y0 * row_bytes, L.pixels.p, bpix}});
        if (rc == KPEG_OK)
            used[nused++] = (int)k;
        row0 += rows;
    }
    const std::string first_err = ctx->err;
    for (int i = 0; i < nused; ++i) { // every enqueued band is completed, whatever happened to the others
        const int frc = job_finish(ctx, used[i], stats);
        if (rc == KPEG_OK && frc != KPEG_OK)
            rc = frc;
        else if (frc != KPEG_OK)
            ctx->err = first_err;
    }
    for (uint32_t k = 0; k < parts; ++k) // lanes whose copies were issued but whose band was never enqueued
        cudaStreamSynchronize(ctx->lane[k].stream);
    if (stats) {
        stats->width = plan->width;
        stats->height = plan->height;
    }
    return rc;
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
    for (Lane &L : ctx->lane) {
        if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) {
            kpeg_cuda_destroy(ctx);
            return KPEG_ERR_CUDA;
        }
    }
    cudaEventCreateWithFlags(&ctx->tables_ready, cudaEventDisableTiming);
    kernels_configure(NLANES);
    kernels_context_created();
    ctx->counted = true;
    if (const char *sb = getenv("KPEG_SUB_BITS")) {
        const long v = strtol(sb, nullptr, 10);
        if (v >= 64 && v <= 1024 && (v & (v - 1)) == 0) {
            ctx->sub_bits = (uint32_t)v;
            ctx->sub_bits_auto = false;
        }
    }
    if (const char *nr = getenv("KPEG_NO_RECORDS"))
        ctx->use_records = !(nr[0] == '1');
    if (const char *e = getenv("KPEG_GUARD"))
        ctx->guard = e[0] == '1';
    if (const char *e = getenv("KPEG_SPLIT"))
        ctx->split_parts = std::max(1, std::min(NLANES, atoi(e)));
    if (const char *e = getenv("KPEG_SUBMIT_SPLIT"))
        ctx->submit_parts = std::max(1, std::min(NLANES, atoi(e)));
    if (const char *e = getenv("KPEG_HOST_CHUNKS"))
        ctx->host_chunks = std::max(1, std::min(64, atoi(e)));
    if (const char *e = getenv("KPEG_BANDS"))
        ctx->band_parts = std::max(1, std::min(NLANES, atoi(e)));
    if (const char *e = getenv("KPEG_RELAY_ROUND2_WIDE"))
        ctx->relay_round2_wide = e[0] != '0';
    if (const char *e = getenv("KPEG_D2H_CHAIN"))
        ctx->d2h_chain = e[0] != '0';
    cudaEventCreateWithFlags(&ctx->d2h_done, cudaEventDisableTiming);
    if (const char *rr = getenv("KPEG_RELAY_ROUNDS")) {
        const long v = strtol(rr, nullptr, 10);
        if (v >= 2 && v < MAX_RELAY_ROUNDS)
            ctx->relay_rounds = (int)v;
    }
    // 8 RSTn separators for packed batches
    cudaMallocHost(&ctx->h_sep.p, 64);
    if (ctx->h_sep.p) {
        ctx->h_sep.cap = 64;
        uint8_t *sep = (uint8_t *)ctx->h_sep.p;
        for (int k = 0; k < 8; ++k) {
            sep[2 * k] = 0xFF;
            sep[2 * k + 1] = (uint8_t)(0xD0 + k);
        }
    }
    if (cudaGetLastError() != cudaSuccess || !ctx->h_sep.p) {
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
    if (ctx->counted)
        kernels_context_destroyed();
    for (Lane &L : ctx->lane) {
        if (L.stream)
            cudaStreamSynchronize(L.stream);
        DevBuf *bufs[] = {&L.cls,  &L.scan, &L.words,    &L.seg_bit,    &L.tile_kept,  &L.tile_rst, &L.state,
                          &L.work, &L.seg_hint, &L.start_slot, &L.scan_tiles, &L.coef,     &L.dcdiff, &L.tiles, &L.tie_rec, &L.tie_cnt,
                          &L.strip_sub, &L.dc, &L.dcs, &L.dcpre, &L.scan_tiles_dcs, &L.pixels, &L.meta,
                          &L.rec,  &L.nrec,       &L.rec_alt};
        for (DevBuf *b : bufs)
            dev_free(*b);
        if (L.h_meta.p)
            cudaFreeHost(L.h_meta.p);
        if (L.h_stage.p)
            cudaFreeHost(L.h_stage.p);
        if (L.h_ends.p)
            cudaFreeHost(L.h_ends.p);
        dev_free(L.d_ends);
        dev_free(L.d_stage);
        for (int i = 0; i < MAX_EVENTS; ++i)
            if (L.ev[i])
                cudaEventDestroy(L.ev[i]);
        if (L.stream)
            cudaStreamDestroy(L.stream);
    }
    dev_free(ctx->tables);
    dev_free(ctx->merged);
    if (ctx->h_tables.p)
        cudaFreeHost(ctx->h_tables.p);
    if (ctx->h_sep.p)
        cudaFreeHost(ctx->h_sep.p);
    if (ctx->tables_ready)
        cudaEventDestroy(ctx->tables_ready);
    if (ctx->d2h_done)
        cudaEventDestroy(ctx->d2h_done);
    for (DevBuf &b : ctx->scan_tiles)
        dev_free(b);
    dev_free(ctx->band_counts);
    if (ctx->h_band_counts.p)
        cudaFreeHost(ctx->h_band_counts.p);
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
        ctx->sub_bits_auto = false;
    } else if (sub_bits < 0) {
        ctx->sub_bits_auto = true; // back to the size chosen per job
    }
    if (relay_rounds > 0) {
        if (relay_rounds < 2 || relay_rounds >= MAX_RELAY_ROUNDS)
            return KPEG_ERR_ARG;
        ctx->relay_rounds = relay_rounds;
    }
    return KPEG_OK;
}

extern "C" void *kpeg_cuda_stream(kpeg_ctx *ctx) { return ctx ? (void *)ctx->lane[0].stream : nullptr; }

extern "C" void *kpeg_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { // pinned for every device's context
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" int kpeg_cuda_host_register(void *p, size_t bytes)
{
    if (!p || !bytes)
        return KPEG_ERR_ARG;
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return KPEG_ERR_CUDA;
    }
    return KPEG_OK;
}

extern "C" void kpeg_cuda_host_unregister(void *p)
{
    if (p && cudaHostUnregister(p) != cudaSuccess)
        cudaGetLastError();
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
        for (Lane &L : ctx->lane)
            cudaStreamSynchronize(L.stream);
        cudaFree(p);
    }
}

extern "C" int kpeg_cuda_memcpy_h2d(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->lane[0].stream));
    CK(cudaStreamSynchronize(ctx->lane[0].stream));
    return KPEG_OK;
}

extern "C" int kpeg_cuda_memcpy_d2h(kpeg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->lane[0].stream));
    CK(cudaStreamSynchronize(ctx->lane[0].stream));
    return KPEG_OK;
}

extern "C" int kpeg_cuda_decode_device(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *d_scan, size_t scan_len,
                                       uint8_t *d_pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || !d_scan || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    mark(ctx, ctx->lane[0], -1);
    TRY(job_enqueue(ctx, 0, plan, d_scan, scan_len, 1, d_pixels_out, {}));
    return job_finish(ctx, 0, stats);
}

extern "C" int kpeg_cuda_decode(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len,
                                uint8_t *pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || !scan || !pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    Lane &L = ctx->lane[0];
    if (L.job.active) // a deferred job still reads this lane's scan / pixel buffers: complete it before they can move
        finish_deferred(ctx, 0);
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    if (ctx->band_parts > 1 && plan->restart_interval != 0 && npix >= ((size_t)64 << 20)) {
        const int brc = decode_banded(ctx, plan, scan, scan_len, pixels_out, stats);
        if (brc != KPEG_ERR_UNSUPPORTED) // UNSUPPORTED: the restart interval does not line up with MCU rows -- whole image below
            return brc;
    }
    TRY(ensure(ctx, L.stream, L.scan, scan_len + 64));
    TRY(ensure(ctx, L.stream, L.pixels, npix + 64));
    mark(ctx, L, -1);
    CK(cudaMemcpyAsync(L.scan.p, scan, scan_len, cudaMemcpyHostToDevice, L.stream));
    mark(ctx, L, KPEG_T_H2D);
    TRY(job_enqueue(ctx, 0, plan, (const uint8_t *)L.scan.p, scan_len, 1, (uint8_t *)L.pixels.p,
                    {Copy{pixels_out, L.pixels.p, npix}}));
    return job_finish(ctx, 0, stats);
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
    zero_stats(stats);
    mark(ctx, ctx->lane[0], -1);
    TRY(job_enqueue(ctx, 0, plan, d_packed, packed_len, (uint32_t)n, d_pixels_out, {}));
    return job_finish(ctx, 0, stats);
}

// Same, with the offsets of the n scans inside the packed stream (packed_offsets[n] = its length): the
// batch is decoded as up to four concurrent jobs, one per lane, so the latency-bound phases of one part
// (late relay rounds, record expansion) overlap the issue-bound phases of the others.
extern "C" int kpeg_cuda_decode_batch_packed_device_split(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                                          const uint64_t *packed_offsets, uint8_t *d_pixels_out,
                                                          kpeg_stats *stats)
{
    if (!ctx || !plan || n <= 0 || !d_packed || !packed_offsets || !d_pixels_out)
        return KPEG_ERR_ARG;
    if (n < 2 || ctx->profiling) // per-kernel event times are only meaningful without a concurrent lane
        return kpeg_cuda_decode_batch_packed_device(ctx, plan, n, d_packed, (size_t)(packed_offsets[n] - packed_offsets[0]),
                                                    d_pixels_out, stats);
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    const int parts = std::max(1, std::min({n, NLANES, ctx->split_parts}));
    for (int li = 0; li < parts; ++li) {
        const int lo = (int)((long long)n * li / parts), hi = (int)((long long)n * (li + 1) / parts);
        mark(ctx, ctx->lane[li], -1);
        const int erc = job_enqueue(ctx, li, plan, d_packed + packed_offsets[lo], (size_t)(packed_offsets[hi] - packed_offsets[lo]),
                                    (uint32_t)(hi - lo), d_pixels_out + npix * (size_t)lo, {});
        if (erc != KPEG_OK) { // the parts already enqueued hold the caller's pointers: complete them before returning
            const std::string why = ctx->err;
            for (int k = 0; k < li; ++k)
                job_finish(ctx, k, nullptr);
            ctx->err = why;
            return erc;
        }
    }
    int rc_all = KPEG_OK;
    for (int li = 0; li < parts; ++li) {
        const int rc = job_finish(ctx, li, stats);
        if (rc != KPEG_OK && (rc_all == KPEG_OK || rc_all == KPEG_ERR_STREAM))
            rc_all = rc; // every lane is completed whatever the outcome; a hard failure outranks a corrupt stream
    }
    return rc_all;
}

// Deferred form of the call above: enqueue and return.  Lanes are taken round-robin; a lane's previous job is
// completed (its status checked, the rare slow path run) only when the lane is needed again or in
// kpeg_cuda_wait, so consecutive batches overlap on the device: the latency-bound phases of one (late relay
// rounds, one-block scans, kernel tails) are filled by the wide kernels of another.  d_packed and
// d_pixels_out must stay valid until kpeg_cuda_wait returns.
extern "C" int kpeg_cuda_submit_batch_packed_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_packed,
                                                    const uint64_t *packed_offsets, uint8_t *d_pixels_out)
{
    if (!ctx || !plan || n <= 0 || !d_packed || !packed_offsets || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    // consecutive submissions already overlap; cutting one further only shortens its kernels (measured)
    const int parts = ctx->profiling ? 1 : std::max(1, std::min({n, NLANES, ctx->submit_parts}));
    for (int k = 0; k < parts; ++k) {
        const int lo = (int)((long long)n * k / parts), hi = (int)((long long)n * (k + 1) / parts);
        const int li = ctx->next_lane;
        ctx->next_lane = (ctx->next_lane + 1) % NLANES;
        if (ctx->lane[li].job.active)
            finish_deferred(ctx, li);
        mark(ctx, ctx->lane[li], -1);
        TRY(job_enqueue(ctx, li, plan, d_packed + packed_offsets[lo], (size_t)(packed_offsets[hi] - packed_offsets[lo]),
                        (uint32_t)(hi - lo), d_pixels_out + npix * (size_t)lo, {}));
    }
    return KPEG_OK;
}

// Complete everything submitted since the last wait.  Returns the first failure among those jobs (its text
// in kpeg_cuda_last_error); *stats receives their accumulated figures.
extern "C" int kpeg_cuda_wait(kpeg_ctx *ctx, kpeg_stats *stats)
{
    if (!ctx)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    for (int k = 0; k < NLANES; ++k) { // oldest submission first
        const int li = (ctx->next_lane + k) % NLANES;
        if (ctx->lane[li].job.active)
            finish_deferred(ctx, li);
    }
    if (stats)
        *stats = ctx->deferred_stats;
    const int rc = ctx->deferred_rc;
    if (rc != KPEG_OK)
        ctx->err = ctx->deferred_err;
    ctx->deferred_rc = KPEG_OK;
    ctx->deferred_err.clear();
    memset(&ctx->deferred_stats, 0, sizeof ctx->deferred_stats);
    return rc;
}

extern "C" int kpeg_cuda_decode_batch_device(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *d_scans,
                                             const uint64_t *scan_offsets, uint8_t *d_pixels_out, kpeg_stats *stats)
{
    // Device-resident scans without separators: re-pack on the device with n small copies.
    if (!ctx || !plan || n <= 0 || !d_scans || !scan_offsets || !d_pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    Lane &L = ctx->lane[0];
    if (L.job.active) // as in kpeg_cuda_decode
        finish_deferred(ctx, 0);
    const size_t total = (size_t)(scan_offsets[n] - scan_offsets[0]) + 2u * (size_t)n;
    TRY(ensure(ctx, L.stream, L.scan, total + 64));
    const uint8_t *sep = (const uint8_t *)ctx->h_sep.p;
    size_t o = 0;
    mark(ctx, L, -1);
    for (int i = 0; i < n; ++i) {
        const size_t len = (size_t)(scan_offsets[i + 1] - scan_offsets[i]);
        CK(cudaMemcpyAsync((uint8_t *)L.scan.p + o, d_scans + scan_offsets[i], len, cudaMemcpyDeviceToDevice, L.stream));
        o += len;
        CK(cudaMemcpyAsync((uint8_t *)L.scan.p + o, sep + 2 * (i & 7), 2, cudaMemcpyHostToDevice, L.stream));
        o += 2;
    }
    mark(ctx, L, KPEG_T_H2D);
    TRY(job_enqueue(ctx, 0, plan, (const uint8_t *)L.scan.p, total, (uint32_t)n, d_pixels_out, {}));
    return job_finish(ctx, 0, stats);
}

// Host-pointer batch, deferred.  Scans go straight from the caller's (ideally pinned) buffers into a lane's
// packed stream (no host-side packing pass); a batch large enough for it to matter is cut into chunks that
// rotate through the lanes, so copies in, kernels and copies out of different chunks -- and of consecutive
// submissions -- overlap.  The end-to-end path is bound by the copy out (3 bytes per pixel over PCIe): the
// point is to keep that copy engine busy from the first chunk to the last.
extern "C" int kpeg_cuda_submit_batch(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *const *scans,
                                      const size_t *scan_lens, uint8_t *const *pixels_out)
{
    if (!ctx || !plan || n <= 0 || !scans || !scan_lens || !pixels_out)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    // chunking: pipeline only when the pixel traffic is worth it (>= 32 MB); chunks of about 8 MB of pixels
    // or more each, at most host_chunks of them
    int per_chunk = n;
    if (n >= 2 && npix * (size_t)n >= (32u << 20)) {
        const int min_imgs = (int)std::max<size_t>(1, (8u << 20) / std::max<size_t>(npix, 1));
        per_chunk = std::max(min_imgs, (n + ctx->host_chunks - 1) / ctx->host_chunks);
    }
    const int nchunks = (n + per_chunk - 1) / per_chunk;
    for (int c = 0; c < nchunks; ++c) {
        const int li = ctx->next_lane;
        ctx->next_lane = (ctx->next_lane + 1) % NLANES;
        Lane &L = ctx->lane[li];
        if (L.job.active) // the lane's previous chunk must be complete before its buffers are reused
            finish_deferred(ctx, li);
        const int i0 = c * per_chunk, i1 = std::min(n, i0 + per_chunk), m = i1 - i0;
        size_t total = 0;
        for (int i = i0; i < i1; ++i)
            total += scan_lens[i] + 2;
        TRY(ensure(ctx, L.stream, L.scan, total + 64));
        TRY(ensure(ctx, L.stream, L.pixels, npix * (size_t)m + 64));
        // Copies in.  A cudaMemcpyAsync costs the host a couple of microseconds whatever its size, so (1) the 2-byte
        // separators are written by a kernel from a table of positions (one small copy per chunk), (2) runs of small
        // scans are gathered in pinned staging memory first and go over as one copy, (3) large scans go straight
        // from the caller's buffer.  With 4096 images of 512x512 this is ~4k copy calls less per batch.
        constexpr size_t SMALL_SCAN = 24u << 10;
        size_t small_bytes = 0;
        for (int i = i0; i < i1; ++i)
            if (scan_lens[i] < SMALL_SCAN)
                small_bytes += scan_lens[i] + 2;
        TRY(ensure_pinned(ctx, L.stream, L.h_ends, (size_t)m * sizeof(uint64_t)));
        TRY(ensure(ctx, L.stream, L.d_ends, (size_t)m * sizeof(uint64_t)));
        if (small_bytes)
            TRY(ensure_pinned(ctx, L.stream, L.h_stage, small_bytes));
        mark(ctx, L, -1);
        uint64_t *ends = (uint64_t *)L.h_ends.p;
        // (0) scans that lie back to back in host memory (a caller that keeps its files in one arena): ONE copy of the
        // whole run into a device staging buffer, and a kernel that moves every scan to its place (the separators need
        // two bytes between scans, so the run cannot be copied to its final place directly)
        bool contiguous = m >= 2;
        for (int i = i0; contiguous && i + 1 < i1; ++i)
            contiguous = scans[i] + scan_lens[i] == scans[i + 1];
        if (contiguous) {
            const size_t run = total - 2u * (size_t)m;
            TRY(ensure(ctx, L.stream, L.d_stage, run + 64));
            size_t o = 0;
            for (int i = i0; i < i1; ++i) {
                o += scan_lens[i];
                ends[i - i0] = o;
                o += 2;
            }
            CK(cudaMemcpyAsync(L.d_stage.p, scans[i0], run, cudaMemcpyHostToDevice, L.stream));
            CK(cudaMemcpyAsync(L.d_ends.p, ends, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, L.stream));
            launch_repack_scans((const uint8_t *)L.d_stage.p, (uint8_t *)L.scan.p, (const uint64_t *)L.d_ends.p, (uint32_t)m, L.stream);
        } else {
            uint8_t *stage = (uint8_t *)L.h_stage.p;
            size_t o = 0, so = 0, run_o = 0, run_so = 0; // device offset, staging offset, start of the open staged run
            auto flush_run = [&]() -> cudaError_t {
                if (so == run_so)
                    return cudaSuccess;
                const cudaError_t e = cudaMemcpyAsync((uint8_t *)L.scan.p + run_o, stage + run_so, so - run_so, cudaMemcpyHostToDevice, L.stream);
                run_so = so;
                return e;
            };
            for (int i = i0; i < i1; ++i) {
                if (scan_lens[i] < SMALL_SCAN) {
                    if (so == run_so)
                        run_o = o;
                    memcpy(stage + so, scans[i], scan_lens[i]);
                    so += scan_lens[i] + 2; // the separator's two bytes travel with the run; the kernel below fills them in
                } else {
                    CK(flush_run());
                    CK(cudaMemcpyAsync((uint8_t *)L.scan.p + o, scans[i], scan_lens[i], cudaMemcpyHostToDevice, L.stream));
                }
                o += scan_lens[i];
                ends[i - i0] = o;
                o += 2;
            }
            CK(flush_run());
            CK(cudaMemcpyAsync(L.d_ends.p, ends, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, L.stream));
            launch_write_separators((uint8_t *)L.scan.p, (const uint64_t *)L.d_ends.p, (uint32_t)m, L.stream);
        }
        mark(ctx, L, KPEG_T_H2D);
        // Copies out: one per run of images whose host buffers follow each other (a caller that decodes into one
        // frame buffer gets ONE copy per chunk)
        std::vector<Copy> d2h;
        for (int i = i0; i < i1;) {
            int j = i + 1;
            while (j < i1 && pixels_out[j] == pixels_out[j - 1] + npix)
                ++j;
            d2h.push_back(Copy{pixels_out[i], (uint8_t *)L.pixels.p + npix * (size_t)(i - i0), npix * (size_t)(j - i)});
            i = j;
        }
        TRY(job_enqueue(ctx, li, plan, (const uint8_t *)L.scan.p, total, (uint32_t)m, (uint8_t *)L.pixels.p, std::move(d2h)));
    }
    return KPEG_OK;
}

// Host-pointer batch, complete on return: a submission followed by a wait of its own.
extern "C" int kpeg_cuda_decode_batch(kpeg_ctx *ctx, const kpeg_plan *plan, int n, const uint8_t *const *scans,
                                      const size_t *scan_lens, uint8_t *const *pixels_out, kpeg_stats *stats)
{
    if (!ctx || !plan || n <= 0 || !scans || !scan_lens || !pixels_out)
        return KPEG_ERR_ARG;
    zero_stats(stats);
    // earlier deferred submissions keep their own outcome for the caller's kpeg_cuda_wait
    for (int i = 0; i < NLANES; ++i)
        if (ctx->lane[i].job.active)
            finish_deferred(ctx, i);
    const int saved_rc = ctx->deferred_rc;
    const std::string saved_err = ctx->deferred_err;
    const kpeg_stats saved_stats = ctx->deferred_stats;
    ctx->deferred_rc = KPEG_OK;
    ctx->deferred_err.clear();
    memset(&ctx->deferred_stats, 0, sizeof ctx->deferred_stats);
    int rc = kpeg_cuda_submit_batch(ctx, plan, n, scans, scan_lens, pixels_out);
    const int wrc = kpeg_cuda_wait(ctx, stats);
    if (rc == KPEG_OK)
        rc = wrc;
    const std::string my_err = ctx->err;
    ctx->deferred_rc = saved_rc;
    ctx->deferred_err = saved_err;
    ctx->deferred_stats = saved_stats;
    ctx->err = my_err;
    return rc;
}

// One band of a tiled decode whose frame lives on another GPU: decode here, then one peer copy of the band's rows
// (NVLink where peer access exists, staged by the driver otherwise).  With dst_device == this context's device the
// kernels write straight into the frame.
extern "C" int kpeg_cuda_decode_to_peer(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *scan, size_t scan_len, int dst_device,
                                        uint8_t *d_dst, kpeg_stats *stats)
{
    if (!ctx || !plan || !scan || !d_dst || dst_device < 0)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    Lane &L = ctx->lane[0];
    if (L.job.active)
        finish_deferred(ctx, 0);
    const size_t npix = (size_t)plan->width * plan->height * plan->ncomp;
    const bool local = dst_device == ctx->device && (reinterpret_cast<uintptr_t>(d_dst) & 7u) == 0;
    TRY(ensure(ctx, L.stream, L.scan, scan_len + 64));
    if (!local) {
        TRY(ensure(ctx, L.stream, L.pixels, npix + 64));
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, dst_device) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(dst_device, 0);
            if (e != cudaSuccess) // already enabled is fine; anything else leaves the staged path
                cudaGetLastError();
        } else {
            cudaGetLastError();
        }
    }
    mark(ctx, L, -1);
    CK(cudaMemcpyAsync(L.scan.p, scan, scan_len, cudaMemcpyHostToDevice, L.stream));
    mark(ctx, L, KPEG_T_H2D);
    std::vector<Copy> out;
    if (!local) {
        Copy c{d_dst, L.pixels.p, npix};
        c.peer = dst_device;
        out.push_back(c);
    }
    TRY(job_enqueue(ctx, 0, plan, (const uint8_t *)L.scan.p, scan_len, 1, local ? d_dst : (uint8_t *)L.pixels.p, std::move(out)));
    return job_finish(ctx, 0, stats);
}

// GPU-side PPM writer: the bytes Image::dumpRawData puts into <name>.ppm (reference src/Image.cpp:108-140: the P6
// header of Image.cpp:124-127, then the R,G,B rows) assembled in device memory, so a consumer that stores or ships the
// file from the GPU never sees a separate header/payload pair.  The payload is written by K3 itself; only the header
// (about a hundred bytes) is copied in.  The file occupies d_out[*ppm_off, *ppm_off + *ppm_len): the offset (< 16)
// keeps the payload 16-byte aligned for K3's vector stores.  One-component images are expanded to R = G = B
// (SURVEY A.8), as kpeg::JPEGDecoder::dumpRawData does on the host.
extern "C" int kpeg_cuda_decode_ppm_device(kpeg_ctx *ctx, const kpeg_plan *plan, const uint8_t *d_scan, size_t scan_len,
                                           uint8_t *d_out, size_t cap, size_t *ppm_off, size_t *ppm_len, kpeg_stats *stats)
{
    if (!ctx || !plan || !d_scan || !d_out || !ppm_off || !ppm_len)
        return KPEG_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_out) & 15u)
        return fail(ctx, KPEG_ERR_ARG, "PPM buffer must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    Lane &L = ctx->lane[0];
    if (L.job.active)
        finish_deferred(ctx, 0);
    char header[160];
    const int hl = kpeg_ppm_header(plan->width, plan->height, header, sizeof header);
    const size_t pad = (16u - (size_t)hl % 16u) % 16u;
    const size_t npx = (size_t)plan->width * plan->height;
    if (pad + (size_t)hl + npx * 3u > cap)
        return fail(ctx, KPEG_ERR_ARG, "PPM buffer too small");
    TRY(ensure_pinned(ctx, L.stream, L.h_stage, 256));
    memcpy(L.h_stage.p, header, (size_t)hl);
    uint8_t *payload = d_out + pad + (size_t)hl;
    uint8_t *k3_out = payload;
    if (plan->ncomp == 1) {
        TRY(ensure(ctx, L.stream, L.pixels, npx + 64));
        k3_out = (uint8_t *)L.pixels.p;
    }
    mark(ctx, L, -1);
    CK(cudaMemcpyAsync(d_out + pad, L.h_stage.p, (size_t)hl, cudaMemcpyHostToDevice, L.stream));
    TRY(job_enqueue(ctx, 0, plan, d_scan, scan_len, 1, k3_out, {}));
    if (plan->ncomp == 1)
        launch_gray_to_rgb((const uint8_t *)L.pixels.p, payload, npx, L.stream);
    *ppm_off = pad;
    *ppm_len = (size_t)hl + npx * 3u;
    return job_finish(ctx, 0, stats);
}

// Planar writer: interleaved R,G,B rows (K3's output) -> three planes [3][H][W] in device memory.
extern "C" int kpeg_cuda_interleaved_to_planar(kpeg_ctx *ctx, const uint8_t *d_rgb, uint8_t *d_planes, size_t npixels)
{
    if (!ctx || !d_rgb || !d_planes)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    Lane &L = ctx->lane[0];
    if (L.job.active)
        finish_deferred(ctx, 0);
    launch_rgb_to_planar(d_rgb, d_planes, npixels, L.stream);
    CK(cudaStreamSynchronize(L.stream));
    CK(cudaGetLastError());
    return KPEG_OK;
}

// A frame coded one scan per component (T.81 A.2.3).  Every scan is a one-component image with tables of its own: K0,
// K1 and K2 run on it unchanged and leave the component's coefficient tiles (DC prediction and the reference's
// DC-difference rule included: both are per component); one kernel interleaves the three tile sets into MCU order and
// K3 runs as for an interleaved file.  Synchronous, on lane 0; not a throughput path.
extern "C" int kpeg_cuda_decode_scans(kpeg_ctx *ctx, const kpeg_plan *frame, const kpeg_scan *scans, int nscans, const uint8_t *file,
                                      size_t file_len, uint8_t *pixels_out, kpeg_stats *stats)
{
    if (!ctx || !frame || !scans || nscans < 1 || !file || !pixels_out)
        return KPEG_ERR_ARG;
    for (int k = 0; k < nscans; ++k)
        if (scans[k].off > file_len || scans[k].len > file_len - scans[k].off)
            return fail(ctx, KPEG_ERR_ARG, "scan outside the file");
    if (nscans == 1 && scans[0].plan.ncomp == frame->ncomp) {
        kpeg_plan pl = scans[0].plan;
        pl.flags = frame->flags;
        return kpeg_cuda_decode(ctx, &pl, file + scans[0].off, scans[0].len, pixels_out, stats);
    }
    if (frame->ncomp != 3 || nscans != 3)
        return fail(ctx, KPEG_ERR_UNSUPPORTED, "a frame is one interleaved scan or one scan per component");
    bool seen[3] = {false, false, false};
    for (int k = 0; k < 3; ++k) {
        if (scans[k].plan.ncomp != 1 || scans[k].comp[0] > 2 || seen[scans[k].comp[0]] || scans[k].plan.width != frame->width ||
            scans[k].plan.height != frame->height)
            return fail(ctx, KPEG_ERR_ARG, "scans do not cover the frame's components once each");
        seen[scans[k].comp[0]] = true;
    }
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    for (int i = 0; i < NLANES; ++i)
        if (ctx->lane[i].job.active)
            finish_deferred(ctx, i);
    Lane &L = ctx->lane[0];
    cudaStream_t s = L.stream;
    JobGeom g3;
    const char *why = nullptr;
    const int grc = make_job_geom(frame, 1, ctx->sub_bits, &g3, &why);
    if (grc != KPEG_OK)
        return fail(ctx, grc, why);
    const uint32_t nstrips = (g3.mcus_per_image + IDCT_MCUS_PER_CTA - 1) / IDCT_MCUS_PER_CTA;
    const uint32_t nmcu_padded = nstrips * IDCT_MCUS_PER_CTA;
    for (int c = 0; c < 3; ++c)
        TRY(ensure(ctx, s, ctx->scan_tiles[c], (size_t)nmcu_padded * 128u + 256u));
    kpeg_stats part;
    uint32_t launches = 0;
    for (int k = 0; k < 3; ++k) {
        kpeg_plan pl = scans[k].plan;
        pl.flags = frame->flags;
        TRY(ensure(ctx, s, L.scan, scans[k].len + 64));
        mark(ctx, L, -1);
        CK(cudaMemcpyAsync(L.scan.p, file + scans[k].off, scans[k].len, cudaMemcpyHostToDevice, s));
        mark(ctx, L, KPEG_T_H2D);
        TRY(job_enqueue(ctx, 0, &pl, (const uint8_t *)L.scan.p, scans[k].len, 1, nullptr, {}, ctx->scan_tiles[scans[k].comp[0]].p));
        memset(&part, 0, sizeof part);
        TRY(job_finish(ctx, 0, stats ? &part : nullptr));
        if (stats) {
            stats->scan_bytes += part.scan_bytes;
            stats->unstuffed_bytes += part.unstuffed_bytes;
            stats->segments += part.segments;
            stats->subsequences += part.subsequences;
            stats->sync_rounds = std::max(stats->sync_rounds, part.sync_rounds);
            stats->kernel_launches += part.kernel_launches;
            for (int t = 0; t < KPEG_T_COUNT; ++t)
                stats->ms[t] += part.ms[t];
            stats->ms_total += part.ms_total;
        }
    }
    // the frame's tables (quantisers of the three components) for K3
    kpeg_plan fp = *frame;
    for (int c = 0; c < 3; ++c) { // K3 needs no Huffman tables, but the device tables are built from a whole plan
        fp.comp_td[c] = scans[0].plan.comp_td[0];
        fp.comp_ta[c] = scans[0].plan.comp_ta[0];
    }
    memcpy(fp.ht, scans[0].plan.ht, sizeof fp.ht);
    memcpy(fp.ht_present, scans[0].plan.ht_present, sizeof fp.ht_present);
    TRY(upload_plan(ctx, &fp));
    const size_t npix = (size_t)frame->width * frame->height * 3u;
    TRY(ensure(ctx, s, L.tiles, (size_t)nmcu_padded * 3u * 128u + 256u));
    TRY(ensure(ctx, s, L.pixels, npix + 64));
    TRY(ensure(ctx, s, L.tie_rec, (size_t)nstrips * IDCT_TIE_LIST_CAP * sizeof(uint4)));
    TRY(ensure(ctx, s, L.tie_cnt, (size_t)nstrips * sizeof(uint32_t)));
    DevMeta *d_meta = (DevMeta *)L.meta.p, *h_meta = (DevMeta *)L.h_meta.p;
    CK(cudaMemsetAsync(d_meta, 0, sizeof(DevMeta), s));
    mark(ctx, L, -1);
    launch_interleave_tiles(ctx->scan_tiles[0].p, ctx->scan_tiles[1].p, ctx->scan_tiles[2].p, L.tiles.p, nmcu_padded, s, &launches);
    IdctArgs ia;
    ia.tiles = L.tiles.p;
    ia.nstrips = nstrips;
    ia.tables = (const DeviceTables *)ctx->tables.p;
    ia.pixels = (uint8_t *)L.pixels.p;
    ia.tie_rec = (uint4 *)L.tie_rec.p;
    ia.tie_cnt = (uint32_t *)L.tie_cnt.p;
    ia.meta = d_meta;
    ia.g = g3;
    CK(launch_idct(ia, s, &launches));
    mark(ctx, L, KPEG_T_IDCT);
    CK(cudaMemcpyAsync(pixels_out, L.pixels.p, npix, cudaMemcpyDeviceToHost, s));
    mark(ctx, L, KPEG_T_D2H);
    CK(cudaMemcpyAsync(h_meta, d_meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (ctx->guard)
        TRY(check_guards(ctx, L));
    ctx->last_lane = 0;
    ctx->last_g = g3; // kpeg_cuda_read_coefficients: the interleaved tiles
    if (stats) {
        stats->width = frame->width;
        stats->height = frame->height;
        stats->ncomp = 3;
        stats->exact_samples += h_meta->exact_samples;
        stats->kernel_launches += launches;
        add_times(ctx, L, stats);
    }
    return KPEG_OK;
}

extern "C" int kpeg_cuda_decode_file(kpeg_ctx *ctx, const uint8_t *file, size_t len, uint32_t flags, uint8_t *pixels_out,
                                     size_t cap, kpeg_plan *plan_out, kpeg_stats *stats)
{
    if (!ctx || !file || !pixels_out)
        return KPEG_ERR_ARG;
    kpeg_plan frame;
    kpeg_scan scans[KPEG_MAX_SCANS];
    int nscans = 0;
    const int prc = kpeg_parse_jfif_scans(file, len, &frame, scans, KPEG_MAX_SCANS, &nscans);
    if (prc != KPEG_OK)
        return fail(ctx, prc, prc == KPEG_ERR_UNSUPPORTED ? "unsupported JPEG coding" : "malformed JFIF container");
    frame.flags = flags;
    if (plan_out)
        *plan_out = frame;
    if ((size_t)frame.width * frame.height * frame.ncomp > cap)
        return fail(ctx, KPEG_ERR_ARG, "pixel buffer too small");
    return kpeg_cuda_decode_scans(ctx, &frame, scans, nscans, file, len, pixels_out, stats);
}

// Any mix of files.  Files with identical plans (dimensions, tables, restart interval) travel together as ONE batch --
// one kernel sequence for the group -- and the groups overlap on the lanes; only when a group fails are its files
// decoded one by one, to find out which of them it was.
extern "C" int kpeg_cuda_decode_files(kpeg_ctx *ctx, int n, const uint8_t *const *files, const size_t *lens, uint32_t flags,
                                      uint8_t *const *pixels_out, const size_t *caps, kpeg_plan *plans_out, int *results,
                                      kpeg_stats *stats)
{
    if (!ctx || n <= 0 || !files || !lens || !pixels_out || !caps)
        return KPEG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    zero_stats(stats);
    struct File {
        kpeg_plan frame;
        kpeg_scan scans[KPEG_MAX_SCANS];
        int nscans = 0, rc = KPEG_OK, group = -1;
    };
    struct Group {
        kpeg_plan plan;
        std::vector<int> members;
        std::vector<const uint8_t *> scans;
        std::vector<size_t> lens;
        std::vector<uint8_t *> outs;
    };
    std::vector<File> fl((size_t)n);
    std::vector<Group> groups;
    std::string first_err;
    int first_rc = KPEG_OK;
    auto note = [&](int i, int rc, const char *what) {
        fl[(size_t)i].rc = rc;
        if (rc != KPEG_OK && first_rc == KPEG_OK) {
            first_rc = rc;
            first_err = std::string("file ") + std::to_string(i) + ": " + what;
        }
    };
    for (int i = 0; i < n; ++i) {
        File &f = fl[(size_t)i];
        if (!files[i] || !pixels_out[i]) {
            note(i, KPEG_ERR_ARG, "null pointer");
            continue;
        }
        const int prc = kpeg_parse_jfif_scans(files[i], lens[i], &f.frame, f.scans, KPEG_MAX_SCANS, &f.nscans);
        if (prc != KPEG_OK) {
            note(i, prc, prc == KPEG_ERR_UNSUPPORTED ? "unsupported JPEG coding" : "malformed JFIF container");
            continue;
        }
        f.frame.flags = flags;
        if (plans_out)
            plans_out[i] = f.frame;
        if ((size_t)f.frame.width * f.frame.height * f.frame.ncomp > caps[i]) {
            note(i, KPEG_ERR_ARG, "pixel buffer too small");
            continue;
        }
        if (f.nscans != 1)
            continue; // one scan per component: decoded on its own below
        kpeg_plan pl = f.scans[0].plan;
        pl.flags = flags;
        size_t g = 0;
        while (g < groups.size() && memcmp(&groups[g].plan, &pl, sizeof pl) != 0)
            ++g;
        if (g == groups.size()) {
            groups.emplace_back();
            groups.back().plan = pl;
        }
        f.group = (int)g;
        groups[g].members.push_back(i);
        groups[g].scans.push_back(files[i] + f.scans[0].off);
        groups[g].lens.push_back(f.scans[0].len);
        groups[g].outs.push_back(pixels_out[i]);
    }
    // earlier deferred submissions keep their own outcome for the caller's kpeg_cuda_wait
    for (int i = 0; i < NLANES; ++i)
        if (ctx->lane[i].job.active)
            finish_deferred(ctx, i);
    const int saved_rc = ctx->deferred_rc;
    const std::string saved_err = ctx->deferred_err;
    const kpeg_stats saved_stats = ctx->deferred_stats;
    auto reset_deferred = [&]() {
        ctx->deferred_rc = KPEG_OK;
        ctx->deferred_err.clear();
        memset(&ctx->deferred_stats, 0, sizeof ctx->deferred_stats);
    };
    auto add_stats = [&](const kpeg_stats &p) {
        if (!stats)
            return;
        stats->scan_bytes += p.scan_bytes;
        stats->unstuffed_bytes += p.unstuffed_bytes;
        stats->segments += p.segments;
        stats->subsequences += p.subsequences;
        stats->sync_rounds = std::max(stats->sync_rounds, p.sync_rounds);
        stats->exact_samples += p.exact_samples;
        stats->kernel_launches += p.kernel_launches;
    };
    // every group is submitted, then ONE wait: the groups overlap on the lanes like consecutive batches do
    reset_deferred();
    int all_rc = KPEG_OK;
    for (Group &g : groups) {
        const int rc = kpeg_cuda_submit_batch(ctx, &g.plan, (int)g.members.size(), g.scans.data(), g.lens.data(), g.outs.data());
        if (all_rc == KPEG_OK)
            all_rc = rc;
    }
    {
        kpeg_stats part;
        const int wrc = kpeg_cuda_wait(ctx, &part);
        if (all_rc == KPEG_OK)
            all_rc = wrc;
        add_stats(part);
    }
    if (all_rc != KPEG_OK) {
        // something failed, and a deferred failure does not say where: group by group, then file by file
        for (Group &g : groups) {
            reset_deferred();
            int rc = kpeg_cuda_submit_batch(ctx, &g.plan, (int)g.members.size(), g.scans.data(), g.lens.data(), g.outs.data());
            kpeg_stats part;
            const int wrc = kpeg_cuda_wait(ctx, &part);
            if (rc == KPEG_OK)
                rc = wrc;
            if (rc == KPEG_OK)
                continue;
            if (g.members.size() == 1) {
                note(g.members[0], rc, ctx->err.c_str());
                continue;
            }
            for (size_t k = 0; k < g.members.size(); ++k) {
                const int one = kpeg_cuda_decode(ctx, &g.plan, g.scans[k], g.lens[k], g.outs[k], &part);
                if (one != KPEG_OK)
                    note(g.members[k], one, ctx->err.c_str());
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        File &f = fl[(size_t)i];
        if (f.rc != KPEG_OK || f.nscans == 1)
            continue;
        kpeg_stats part;
        const int rc = kpeg_cuda_decode_scans(ctx, &f.frame, f.scans, f.nscans, files[i], lens[i], pixels_out[i], &part);
        add_stats(part);
        if (rc != KPEG_OK)
            note(i, rc, ctx->err.c_str());
    }
    ctx->deferred_rc = saved_rc;
    ctx->deferred_err = saved_err;
    ctx->deferred_stats = saved_stats;
    if (results)
        for (int i = 0; i < n; ++i)
            results[i] = fl[(size_t)i].rc;
    if (first_rc != KPEG_OK)
        ctx->err = first_err;
    return first_rc;
}

extern "C" int kpeg_cuda_read_coefficients(kpeg_ctx *ctx, int16_t *out, size_t cap)
{
    if (!ctx || !out)
        return KPEG_ERR_ARG;
    if (ctx->last_lane < 0)
        return fail(ctx, KPEG_ERR_ARG, "no decode has run on this context");
    CK(cudaSetDevice(ctx->device));
    Lane &L = ctx->lane[ctx->last_lane];
    if (L.job.active)
        return fail(ctx, KPEG_ERR_ARG, "the lane of the last decode is busy with a deferred submission");
    const size_t n = (size_t)ctx->last_g.total_blocks * 64u;
    if (cap < n)
        return fail(ctx, KPEG_ERR_ARG, "coefficient buffer too small");
    TRY(ensure(ctx, L.stream, ctx->merged, n * 2u));
    // the coefficient tiles of the lane's last job are still resident: un-bias and un-swizzle them
    launch_matrix_from_tiles(L.tiles.p, (int16_t *)ctx->merged.p, ctx->last_g.total_blocks, L.stream);
    CK(cudaMemcpyAsync(out, ctx->merged.p, n * 2u, cudaMemcpyDeviceToHost, L.stream));
    CK(cudaStreamSynchronize(L.stream));
    CK(cudaGetLastError());
    return KPEG_OK;
}
