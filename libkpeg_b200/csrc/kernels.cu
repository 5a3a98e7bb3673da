// kernels.cu -- hand-written sm_100a kernels of the baseline-JPEG decode hot path.
//
//   K0  unstuff_*      FF00 / RSTn removal + restart-segment table     (Decoder.cpp:532-577, 621-653)
//   K1  entropy_*      Huffman + run-length decode, parallel over fixed-size subsequences with a
//                      self-synchronising relay and a segmented offset scan (Decoder.cpp:655-855)
//   K2  dc_*           DC prediction as a segmented prefix scan           (MCU.cpp:107-108)
//   K3  idct_kernel    dequantise + de-zigzag + 8x8 IDCT + level shift + YCbCr->RGB + interleaved
//                      store, one pass over HBM                           (MCU.cpp:110-279, Image.cpp:51-70)
//
// All file:line citations are relative to /root/reference.  The arithmetic lives in
// entropy_core.h / idct_core.h (host+device inline, also exercised on the CPU by tests/emu).
#include <cuda.h> // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <stddef.h>
#include <stdint.h>

#include <utility>

#include "entropy_core.h"
#include "idct_core.h"
#include "kernels.cuh"
#include "unstuff_core.h"
#include "kpeg_common.h"

namespace kpeg {

// =================================================================================================
// small block-level primitives
// =================================================================================================

template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_sum(uint32_t v, uint32_t *s_warp /*[THREADS/32 + 1]*/, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += o;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < THREADS / 32 ? s_warp[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d)
                wi += o;
        }
        if (lane < THREADS / 32)
            s_warp[lane] = wi - w; // exclusive warp offsets
        if (lane == 31)
            s_warp[THREADS / 32] = wi;
    }
    __syncthreads();
    total = s_warp[THREADS / 32];
    const uint32_t r = s_warp[warp] + incl - v;
    __syncthreads();
    return r;
}

// =================================================================================================
// K0: unstuffing and restart-marker scan
// =================================================================================================
//
// Byte classes inside an entropy-coded segment (T.81 B.1.1.5): FF 00 is a data byte FF; FF D0..D7 is
// a restart marker (segment boundary); FF FF is a fill byte; anything else is flagged.  The
// reference drops only the 00 (Decoder.cpp:631-650) and cannot handle RSTn.

// Classification happens once, here: the keep / RSTn masks of every 16-byte chunk are stored for the
// compaction kernel (4 bytes per 16 of input).  Chunks inside the segment are classified sixteen bytes at a
// time (classify16_swar); the neighbouring bytes come from the adjacent lanes.
__global__ void __launch_bounds__(UNSTUFF_THREADS) unstuff_count_kernel(UnstuffArgs a)
{
    __shared__ uint32_t s_w[UNSTUFF_THREADS / 32 + 1];
    const uint32_t chunk = blockIdx.x * UNSTUFF_THREADS + threadIdx.x;
    // chunks are cut on 16-byte ADDRESS boundaries: chunk c covers stream bytes [16c - mis, 16c - mis + 16), so
    // that a segment that starts anywhere (a part of a packed batch) is still read with aligned 16-byte loads
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(a.scan) & 15u);
    const int32_t start = (int32_t)(chunk * UNSTUFF_BYTES_PER_THREAD) - (int32_t)mis; // segments are < 2^29 bytes
    const int lane = threadIdx.x & 31;
    uint32_t keep = 0, rst = 0, bad = 0;
    const bool whole = start >= 0 && start + 16 <= (int32_t)a.scan_len; // not whole: first / last chunk, or beyond the end
    const bool some = start + 16 > 0 && start < (int32_t)a.scan_len;
    uint32_t b[4] = {0, 0, 0, 0};
    uint32_t in_range = 0xFFFFu;
    if (whole) {
        ld_16bytes(a.scan + start, b);
    } else if (some) {
        // ragged first / last chunk: bytes before the stream read as 00 (no effect on what follows), bytes after
        // it as FF (what classify16 assumes past the end); only the bits of real bytes count
        in_range = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int32_t at = start + i;
            uint32_t v = at < 0 ? 0x00u : 0xFFu;
            if (at >= 0 && at < (int32_t)a.scan_len) {
                v = ld_byte(a.scan + at);
                in_range |= 1u << i;
            }
            b[i >> 2] |= v << (8 * (i & 3));
        }
    }
    // bytes on either side of the chunk: from the neighbouring lanes, from memory at the warp's edges
    uint32_t prev = __shfl_up_sync(0xffffffffu, b[3] >> 24, 1);
    uint32_t next = __shfl_down_sync(0xffffffffu, b[0] & 0xFFu, 1);
    if (__shfl_down_sync(0xffffffffu, some ? 1 : 0, 1) == 0)
        next = 0xFFu; // the next chunk lies past the end
    if (some) {
        if (lane == 0)
            prev = start > 0 ? (uint32_t)ld_byte(a.scan + start - 1) : 0u;
        if (lane == 31)
            next = start + 16 < (int32_t)a.scan_len ? (uint32_t)ld_byte(a.scan + start + 16) : 0xFFu;
        classify16_swar(b, prev, next, keep, rst, bad);
        keep &= in_range;
        rst &= in_range;
    }
    a.cls[chunk] = keep | (rst << 16);
    // only the tile's totals are needed here: one warp-wide integer reduction (REDUX) per warp and one barrier.
    // Both counts in one word: kept <= 4096 per tile, rst <= 2048.
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, __popc(keep) | (__popc(rst) << 16));
    if (lane == 0)
        s_w[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < UNSTUFF_THREADS / 32; ++w)
            tot += s_w[w];
        a.tile_kept[blockIdx.x] = tot & 0xFFFFu;
        a.tile_rst[blockIdx.x] = tot >> 16;
    }
    if (bad)
        atomicOr(&a.meta->status, ST_BAD_MARKER);
}

// One block: exclusive scan of the per-tile counts, stream totals, segment-table prefill.  Every thread
// takes a run of consecutive tiles (serial), the runs are combined with ONE block scan.
__global__ void __launch_bounds__(1024) unstuff_scan_kernel(UnstuffArgs a, uint32_t sub_bits)
{
    __shared__ uint32_t s_w[1024 / 32 + 1];
    const uint32_t per = (a.ntiles + 1023u) / 1024u;
    const uint32_t t0 = threadIdx.x * per, t1 = min(t0 + per, a.ntiles);
    uint32_t sk = 0, sr = 0;
    for (uint32_t t = t0; t < t1; ++t) {
        sk += a.tile_kept[t];
        sr += a.tile_rst[t];
    }
    uint32_t total_kept, total_rst;
    uint32_t ek = block_exclusive_sum<1024>(sk, s_w, total_kept);
    uint32_t er = block_exclusive_sum<1024>(sr, s_w, total_rst);
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t k = a.tile_kept[t], r = a.tile_rst[t];
        a.tile_kept[t] = ek;
        a.tile_rst[t] = er;
        ek += k;
        er += r;
    }
    const uint32_t total_bits = total_kept * 8u;
    for (uint32_t k = threadIdx.x; k < a.nseg + 2u; k += 1024)
        a.seg_bit[k] = k == 0u ? 0u : (k <= a.nseg ? total_bits : 0xFFFFFFFFu);
    // the stream buffer is not zero-filled: define the last (partial) word and the slack the bit window
    // may look at; the compaction kernel, which runs after this one, stores the real tail bytes
    if (threadIdx.x < 12u)
        reinterpret_cast<uint32_t *>(a.words)[(total_kept >> 2) + threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        a.meta->total_kept = total_kept;
        a.meta->total_rst = total_rst;
        a.meta->total_bits = total_bits;
        const uint32_t nsub = (total_bits + sub_bits - 1u) / sub_bits;
        a.meta->nsub = nsub ? nsub : 1u;
        if (total_rst != a.nseg - 1u && total_rst != a.nseg)
            atomicOr(&a.meta->status, ST_SEG_COUNT);
    }
}

// Compaction: the tile's surviving bytes are gathered in shared memory at their final phase within a
// 32-bit word, then written out as whole byte-swapped words (coalesced); only the at most three
// bytes at either end of the tile's output range, which share a word with a neighbouring tile, go out
// as single bytes.  A thread deletes its chunk's dropped bytes in registers (rarely more than one), then
// stores the survivors at their byte offset: whole words where it owns the word, an OR into the zeroed
// buffer where a word is shared with the neighbouring chunk.
__device__ __forceinline__ void delete_byte(uint32_t (&b)[4], uint32_t q) // bytes above q move down by one
{
    const uint32_t t0 = __funnelshift_r(b[0], b[1], 8), t1 = __funnelshift_r(b[1], b[2], 8),
                   t2 = __funnelshift_r(b[2], b[3], 8), t3 = b[3] >> 8;
    const uint32_t wq = q >> 2, low = (1u << (8u * (q & 3u))) - 1u; // bytes of word wq below q stay
    b[0] = wq == 0u ? (b[0] & low) | (t0 & ~low) : b[0];
    b[1] = wq == 1u ? (b[1] & low) | (t1 & ~low) : (wq < 1u ? t1 : b[1]);
    b[2] = wq == 2u ? (b[2] & low) | (t2 & ~low) : (wq < 2u ? t2 : b[2]);
    b[3] = wq == 3u ? (b[3] & low) | (t3 & ~low) : t3;
}

__global__ void __launch_bounds__(UNSTUFF_THREADS) unstuff_write_kernel(UnstuffArgs a)
{
    __shared__ uint32_t s_w[UNSTUFF_THREADS / 32 + 1];
    __shared__ __align__(16) uint8_t s_out[UNSTUFF_TILE + 16];
    const uint32_t chunk = blockIdx.x * UNSTUFF_THREADS + threadIdx.x;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(a.scan) & 15u); // chunks as in the count kernel
    const int32_t start = (int32_t)(chunk * UNSTUFF_BYTES_PER_THREAD) - (int32_t)mis; // segments are < 2^29 bytes
    reinterpret_cast<uint4 *>(s_out)[threadIdx.x] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0)
        reinterpret_cast<uint4 *>(s_out)[UNSTUFF_THREADS] = make_uint4(0, 0, 0, 0);
    const uint32_t cls = a.cls[chunk];
    uint32_t keep = cls & 0xFFFFu;
    const uint32_t rst = cls >> 16;
    uint32_t b[4] = {0, 0, 0, 0};
    if (keep) {
        if (start >= 0 && start + 16 <= (int32_t)a.scan_len) {
            ld_16bytes(a.scan + start, b);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (start + i >= 0 && start + i < (int32_t)a.scan_len)
                    b[i >> 2] |= (uint32_t)ld_byte(a.scan + start + i) << (8 * (i & 3));
        }
    }
    const uint32_t n = __popc(keep);
    uint32_t tot;
    const uint32_t ex = block_exclusive_sum<UNSTUFF_THREADS>(n | (__popc(rst) << 16), s_w, tot); // syncs: s_out is zero
    const uint32_t pos0 = a.tile_kept[blockIdx.x]; // first output byte of the tile
    const uint32_t nkept = tot & 0xFFFFu;
    const uint32_t phase = pos0 & 3u;
    const uint32_t lpos = phase + (ex & 0xFFFFu); // position inside s_out
    if (rst) { // restart markers: the segment that follows starts at the next surviving byte
        uint32_t ridx = a.tile_rst[blockIdx.x] + (ex >> 16);
        uint32_t r = rst;
        while (r) {
            const int i = __ffs(r) - 1;
            r &= r - 1u;
            ++ridx;
            if (ridx < a.nseg)
                a.seg_bit[ridx] = (pos0 - phase + lpos + __popc(keep & ((1u << i) - 1u))) * 8u;
        }
    }
    if (n) {
        uint32_t drop = ~keep & 0xFFFFu;
        while (drop) { // highest dropped byte first: the indices below it stay valid
            const uint32_t q = 31u - (uint32_t)__clz(drop);
            drop &= ~(1u << q);
            delete_byte(b, q);
        }
        // survivors b[0..n) -> s_out[lpos .. lpos + n)
        const uint32_t sh = 8u * (lpos & 3u);
        uint32_t *w = reinterpret_cast<uint32_t *>(s_out) + (lpos >> 2);
        const uint32_t o[5] = {b[0] << sh, __funnelshift_l(b[0], b[1], sh), __funnelshift_l(b[1], b[2], sh),
                               __funnelshift_l(b[2], b[3], sh), sh ? b[3] >> (32u - sh) : 0u};
        const uint32_t first = lpos & 3u, end = first + n; // byte range [first, end) of the five words
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t lo = max(first, 4u * k), hi = min(end, 4u * k + 4u);
            if (hi > lo) {
                if (hi - lo == 4u) {
                    w[k] = o[k];
                } else {
                    const uint32_t m = (hi - 4u * k == 4u ? 0xFFFFFFFFu : (1u << (8u * (hi - 4u * k))) - 1u) &
                                       ~((1u << (8u * (lo - 4u * k))) - 1u);
                    atomicOr(&w[k], o[k] & m);
                }
            }
        }
    }
    __syncthreads();
    // s_out[phase .. phase + nkept) -> global bytes [pos0, pos0 + nkept); word g of the stream is
    // s_out word (g - pos0/4), stored big-endian
    const uint32_t g0 = pos0 >> 2;
    const uint32_t first_full = phase ? 1u : 0u;          // word 0 is shared with the previous tile
    const uint32_t end_byte = phase + nkept;              // in s_out coordinates
    const uint32_t nfull_end = end_byte >> 2;             // words [first_full, nfull_end) are complete
    uint32_t *gw = reinterpret_cast<uint32_t *>(a.words) + g0;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s_out);
    for (uint32_t w = first_full + threadIdx.x; w < nfull_end; w += UNSTUFF_THREADS)
        gw[w] = __byte_perm(sw[w], 0, 0x0123);
    if (threadIdx.x < 8u) {
        // head bytes (phase .. 3 of word 0) and tail bytes (the incomplete last word), one per thread
        uint32_t l;
        if (threadIdx.x < 4u)
            l = threadIdx.x; // candidate head byte
        else
            l = (nfull_end << 2) + (threadIdx.x - 4u); // candidate tail byte
        const bool head = threadIdx.x < 4u && phase && l >= phase && l < end_byte && l < 4u;
        const bool tail = threadIdx.x >= 4u && l < end_byte && l >= phase && (nfull_end >= first_full) &&
                          !(phase && nfull_end == 0u); // word 0 incomplete at both ends: the head threads own it
        if (head || tail)
            a.words[((g0 << 2) + l) ^ 3u] = s_out[l];
    }
}

void launch_unstuff(const UnstuffArgs &a, uint32_t sub_bits, cudaStream_t s, uint32_t *launches)
{
    unstuff_count_kernel<<<a.ntiles, UNSTUFF_THREADS, 0, s>>>(a);
    unstuff_scan_kernel<<<1, 1024, 0, s>>>(a, sub_bits);
    unstuff_write_kernel<<<a.ntiles, UNSTUFF_THREADS, 0, s>>>(a);
    *launches += 3;
}

// =================================================================================================
// K1: entropy decode
// =================================================================================================

// Shared-memory image of what one CTA needs: the lookup tables (staged once per CTA) and the slice of
// the bit stream that belongs to the tile of ENTROPY_THREADS subsequences being decoded.  The slice is
// stored linearly with one padding word after every 32 (position l + (l >> 5)): lanes reading word k
// of their own subsequence (stride 8, 16 or 32 words) then hit 32 different banks.  The slice
// carries four extra words: a symbol may run up to 26 bits past the end of the last subsequence and
// the bit window looks two words ahead.
struct K1Smem {
    uint16_t fast[MAX_LUTS * LUT_SIZE];
    uint16_t longlut[MAX_LUTS * LONG_CAP];
    uint32_t long_n[8];
    uint32_t words[1]; // padded(ENTROPY_THREADS * words_per_subsequence + 4), sized at launch
};

constexpr uint32_t K1_TAIL_WORDS = 4;

__host__ __device__ inline uint32_t k1_pad(uint32_t l) { return l + (l >> 5); }

__host__ __device__ inline size_t k1_smem_bytes(uint32_t sub_bits)
{
    return sizeof(K1Smem) + (size_t)(k1_pad(ENTROPY_THREADS * (sub_bits / 32u) + K1_TAIL_WORDS) + 2u) * sizeof(uint32_t);
}

struct SmemWords {
    uint32_t addr; // shared byte address of words[0]
    uint32_t gw0;  // global word index of the tile's first word
    __device__ __forceinline__ uint32_t operator()(uint32_t gw) const
    {
        const uint32_t l = gw - gw0;
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr + 4u * (l + (l >> 5))));
        return v;
    }
};

struct SmemLuts {
    uint32_t fast_addr; // shared byte address of fast[0]
    uint32_t long_addr; // shared byte address of longlut[0]
    const K1Smem *sm;
    const HuffCanon *canon; // global
    __device__ __forceinline__ uint32_t fast(uint32_t toff, uint32_t idx) const
    {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(fast_addr + 2u * (toff + idx)));
        return v;
    }
    __device__ __forceinline__ uint32_t slow(uint32_t toff, uint32_t win, uint32_t e) const
    {
        const uint32_t ti = toff >> LUT_BITS;
        if (e) { // pointer to a 64-entry sub-table indexed by stream bits 10..15
            uint16_t v;
            const uint32_t idx = ti * (uint32_t)LONG_CAP + ((e >> 5) - 1u) * (uint32_t)SUB_SIZE + ((win >> 16) & (uint32_t)(SUB_SIZE - 1));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(long_addr + 2u * idx));
            return v;
        }
        return huff_slow_lookup(canon[ti], win);
    }
};

// Tables: first level of every table in use, populated part of the second level.  No barrier.
__device__ __forceinline__ void k1_stage_tables(K1Smem &sm, const EntropyArgs &a)
{
    const DeviceTables *t = a.tables;
    const uint32_t nt = a.g.ncomp * 2u;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&t->luts.fast[0][0]);
        uint4 *dst = reinterpret_cast<uint4 *>(sm.fast);
        const uint32_t n = nt * (uint32_t)(LUT_SIZE * sizeof(uint16_t) / 16);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < MAX_LUTS)
        sm.long_n[threadIdx.x] = t->luts.long_n[threadIdx.x];
    for (uint32_t ti = 0; ti < nt; ++ti) {
        const uint32_t n = (t->luts.long_n[ti] * (uint32_t)sizeof(uint16_t) + 15u) / 16u;
        const uint4 *src = reinterpret_cast<const uint4 *>(&t->luts.longlut[ti][0]);
        uint4 *dst = reinterpret_cast<uint4 *>(&sm.longlut[ti * LONG_CAP]);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            dst[i] = __ldg(src + i);
    }
}

// The tile's stream slice (coalesced loads).  Ends with a __syncthreads().
template <int TILE = ENTROPY_THREADS>
__device__ __forceinline__ void k1_stage_stream(K1Smem &sm, const EntropyArgs &a, uint32_t tile, uint32_t total_bits,
                                                uint32_t wlog)
{
    const uint32_t gw0 = (tile * TILE) << wlog;
    const uint32_t nwords = ((uint32_t)TILE << wlog) + K1_TAIL_WORDS;
    const uint32_t total_words = ((total_bits + 31u) >> 5) + 4u; // the stream buffer has zeroed slack beyond this
    for (uint32_t l = threadIdx.x; l < nwords; l += blockDim.x) {
        const uint32_t gw = gw0 + l;
        sm.words[k1_pad(l)] = gw < total_words ? __ldg(a.words + gw) : 0u;
    }
    __syncthreads();
}

// first segment index whose start bit is >= bit  (seg_bit[0..nseg] ascending, seg_bit[nseg] = total_bits)
__device__ __forceinline__ uint32_t first_seg_at_or_after(const uint32_t *seg_bit, uint32_t nseg, uint32_t bit)
{
    uint32_t lo = 0, hi = nseg; // answer in [lo, hi]
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(seg_bit + mid) >= bit)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ SmemLuts k1_luts(const K1Smem &sm, const EntropyArgs &a)
{
    SmemLuts L;
    L.fast_addr = (uint32_t)__cvta_generic_to_shared(sm.fast);
    L.long_addr = (uint32_t)__cvta_generic_to_shared(sm.longlut);
    L.sm = &sm;
    L.canon = a.tables->canon;
    return L;
}

template <int TILE = ENTROPY_THREADS>
__device__ __forceinline__ SmemWords k1_words(const K1Smem &sm, uint32_t tile, uint32_t wlog)
{
    SmemWords W;
    W.addr = (uint32_t)__cvta_generic_to_shared(sm.words);
    W.gw0 = (tile * TILE) << wlog;
    return W;
}

// Persistent over tiles of ENTROPY_THREADS subsequences: tables are staged once per CTA.
__global__ void __launch_bounds__(ENTROPY_THREADS) entropy_cold_kernel(EntropyArgs a, uint32_t wlog)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_raw);
    const uint32_t nsub = a.meta->nsub, total_bits = a.meta->total_bits;
    const uint32_t ntiles = (nsub + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
    if (blockIdx.x >= ntiles)
        return;
    k1_stage_tables(sm, a);
    const SmemLuts L = k1_luts(sm, a);
    StreamView S{a.seg_bit, total_bits};
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        k1_stage_stream(sm, a, tile, total_bits, wlog);
        const uint32_t sub = tile * ENTROPY_THREADS + threadIdx.x;
        if (sub < nsub) {
            const SmemWords W = k1_words(sm, tile, wlog);
            const uint32_t p0 = sub << (wlog + 5u);
            const uint32_t end = min(p0 + a.g.sub_bits, total_bits);
            const uint32_t hint = a.g.nseg > 1u ? first_seg_at_or_after(a.seg_bit, a.g.nseg, p0) : (sub ? 1u : 0u);
            a.seg_hint[sub] = hint;
            const SubState out = decode_span<false>(W, L, S, a.g, end, p0, 0u, 0u, hint, 0u, nullptr, nullptr, nullptr);
            a.state[sub] = out;
        }
        __syncthreads(); // the slice is overwritten by the next tile
    }
}

// Record list layout: groups of 32 consecutive subsequences (one warp), record k of lane l of group g at
// (g * kmax + k) * 32 + l.  Lanes emit / read their k-th record in the same loop iteration, so a warp
// touches one 128-byte line per k, and consecutive k are consecutive lines: a sequential stream per warp.
__device__ __forceinline__ size_t rec_base_index(uint32_t sub, uint32_t kmax)
{
    return ((size_t)(sub >> 5) * kmax) * 32u + (sub & 31u);
}

// A subsequence that is decoded again in a sparse relay round would overwrite its column of that layout
// with lone 4-byte stores, one per 128-byte line (each a DRAM sector read-modify-write).  It gets a
// private, contiguous area of kmax records in rec_alt instead (allocated once, on its first sparse
// visit); nrec[sub] = count | (area index + 1) << 10 tells the final pass where to read.
constexpr uint32_t NREC_MASK = 1023u;

struct GlobalRecorder {
    uint32_t *base; // &rec[rec_base_index(sub)], or the subsequence's private area
    uint32_t kmax;
    uint32_t stride_bytes; // 128 in the warp-interleaved layout, 4 in a private area: one IMAD.WIDE per address
    __device__ __forceinline__ void emit(uint32_t k, uint32_t w) const
    {
        if (k < kmax)
            *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(base) + (size_t)k * stride_bytes) = w;
    }
};

// decode subsequence `sub` from (p, cz) for a relay pass; emits records when the job uses them
template <class Words>
__device__ __forceinline__ SubState relay_decode(const EntropyArgs &a, const Words &W, const SmemLuts &L, const StreamView &S,
                                                 uint32_t sub, uint32_t end, uint32_t p, uint32_t cz, bool sparse)
{
    DecState d;
    dec_init(d, W, S, p, cz >> 8, cz & 0xFFu, a.seg_hint[sub], 0u);
    if (a.rec) {
        GlobalRecorder R;
        R.base = a.rec + rec_base_index(sub, a.rec_kmax);
        R.kmax = a.rec_kmax;
        asm volatile("" : "+r"(R.kmax)); // a register, not a reload of the kernel parameter per symbol
        R.stride_bytes = 128u;
        uint32_t area = 0;
        if (sparse) {
            area = a.nrec[sub] >> 10;
            if (area == 0u) {
                const uint32_t idx = atomicAdd(&a.meta->rec_alt_count, 1u);
                area = idx < a.rec_alt_cap ? idx + 1u : 0u;
            }
            if (area) {
                R.base = a.rec_alt + (size_t)(area - 1u) * a.rec_kmax;
                R.stride_bytes = 4u;
            }
        }
        decode_run<false, true>(d, W, L, S, a.g, end, 0xFFFFFFFFu, NullSink{}, R);
        a.nrec[sub] = min(d.nrec, NREC_MASK) | (area << 10);
        if (d.nrec > a.rec_kmax || (d.st & ST_REC_OVERFLOW))
            atomicOr(&a.meta->status, ST_REC_OVERFLOW);
    } else {
        decode_run<false, false>(d, W, L, S, a.g, end, 0xFFFFFFFFu, NullSink{}, NoRecorder{});
    }
    return dec_exit_state(d);
}

__device__ __forceinline__ int relay_slot_dev(int r) { return r < MAX_RELAY_ROUNDS ? r : 2 + ((r - 2) % (MAX_RELAY_ROUNDS - 2)); }

// Relay: X[i] = decode(i, X[i-1]).  Round 1 visits every subsequence (tiles, like the cold pass) and
// appends i+1 to the work list whenever X[i] changed; later rounds visit only the work list of the
// round before (a few percent of the subsequences, scattered, so they read the stream from global
// memory) until a round appends nothing: the fixed point.
__device__ __forceinline__ void relay_publish(const EntropyArgs &a, uint32_t sub, const SubState &out, uint32_t nsub,
                                              uint32_t *list_out, uint32_t *count_out)
{
    const SubState old = a.state[sub];
    if (old.p != out.p || old.cz != out.cz || old.n != out.n || old.seg != out.seg) {
        uint4 o;
        o.x = out.p;
        o.y = out.n;
        o.z = out.cz;
        o.w = (uint32_t)out.seg;
        __stcg(reinterpret_cast<uint4 *>(&a.state[sub]), o);
        // Only a change of the exit STATE (position, component, zig-zag index) matters downstream; the
        // slot count of a subsequence changes almost always when its entry does, its exit state rarely.
        if ((old.p != out.p || old.cz != out.cz) && sub + 1u < nsub)
            list_out[atomicAdd(count_out, 1u)] = sub + 1u;
    }
}

__global__ void __launch_bounds__(ENTROPY_THREADS, 10) entropy_relay_full_kernel(EntropyArgs a, uint32_t wlog)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_raw);
    const uint32_t nsub = a.meta->nsub, total_bits = a.meta->total_bits;
    const uint32_t ntiles = (nsub + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
    if (blockIdx.x >= ntiles)
        return;
    k1_stage_tables(sm, a);
    const SmemLuts L = k1_luts(sm, a);
    StreamView S{a.seg_bit, total_bits};
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        k1_stage_stream(sm, a, tile, total_bits, wlog);
        const uint32_t sub = tile * ENTROPY_THREADS + threadIdx.x;
        if (sub < nsub) {
            // cold value or already relayed: either is a valid iterate; subsequence 0 starts from the true state
            SubState in;
            in.p = 0;
            in.cz = 0;
            if (sub)
                in = a.state[sub - 1];
            const uint32_t start = sub << (wlog + 5u);
            // without records, a subsequence whose cold decode already started from this state is done
            if (a.rec || (sub != 0u && !(in.p == start && in.cz == 0u))) {
                const SmemWords W = k1_words(sm, tile, wlog);
                const uint32_t end = min((sub + 1u) << (wlog + 5u), total_bits);
                const SubState out = relay_decode(a, W, L, S, sub, end, in.p, in.cz, false);
                if (sub)
                    relay_publish(a, sub, out, nsub, a.worklist[1], &a.meta->changed[1]);
            }
        }
        __syncthreads();
    }
}

// words of ONE subsequence, staged by its own thread (stride odd -> few bank conflicts)
struct PrivateWords {
    uint32_t addr; // shared byte address of this thread's first word
    uint32_t gw0;  // global index of that word
    __device__ __forceinline__ uint32_t operator()(uint32_t gw) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr + 4u * (gw - gw0)));
        return v;
    }
};

__host__ __device__ inline size_t k1_sparse_smem_bytes(uint32_t sub_bits)
{
    return sizeof(K1Smem) + (size_t)ENTROPY_THREADS * (sub_bits / 32u + 5u) * sizeof(uint32_t);
}

__global__ void __launch_bounds__(ENTROPY_THREADS) entropy_relay_sparse_kernel(EntropyArgs a, int round, int slot_prev,
                                                                               int slot_cur, uint32_t wlog)
{
    const uint32_t count = a.meta->changed[slot_prev];
    if (blockIdx.x * ENTROPY_THREADS >= count)
        return; // also the whole grid once the fixed point is reached (count == 0)
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_raw);
    k1_stage_tables(sm, a);
    const uint32_t w = blockIdx.x * ENTROPY_THREADS + threadIdx.x;
    const uint32_t nsub = a.meta->nsub, total_bits = a.meta->total_bits;
    const uint32_t sub = w < count ? a.worklist[(round - 1) & 1][w] : 0xFFFFFFFFu;
    const bool active = sub < nsub;
    uint4 inraw = make_uint4(0, 0, 0, 0);
    const uint32_t stride = (1u << wlog) + 5u;
    uint32_t *mine = sm.words + threadIdx.x * stride;
    if (active) {
        inraw = __ldcg(reinterpret_cast<const uint4 *>(&a.state[sub - 1]));
        const uint32_t j0 = inraw.x >> 5;
        const uint32_t total_words = ((total_bits + 31u) >> 5) + 4u;
        for (uint32_t k = 0; k < stride - 1u; ++k)
            mine[k] = j0 + k < total_words ? __ldg(a.words + j0 + k) : 0u;
    }
    __syncthreads(); // tables staged
    if (!active)
        return;
    const SmemLuts L = k1_luts(sm, a);
    PrivateWords W;
    W.addr = (uint32_t)__cvta_generic_to_shared(mine);
    W.gw0 = inraw.x >> 5;
    StreamView S{a.seg_bit, total_bits};
    const uint32_t end = min((sub + 1u) << (wlog + 5u), total_bits);
    const SubState out = relay_decode(a, W, L, S, sub, end, inraw.x, inraw.z, true);
    relay_publish(a, sub, out, nsub, a.worklist[round & 1], &a.meta->changed[slot_cur]);
}

// The later relay rounds as one launch.  Each round touches only a work list that shrinks quickly
// (5 %, 0.5 %, ... of the subsequences) but costs the full latency of one serial subsequence decode;
// as separate launches every round also paid launch latency, table staging and a tail.  Here a small
// cooperative grid stages the tables once and loops: process the list, grid barrier, next list.
__device__ __forceinline__ uint32_t load_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Returns false when the other CTAs did not arrive within `spin_limit` polls (about a microsecond each once the
// back-off has grown): the grid is not co-resident -- another process or an MPS client holds the SMs the launch-time
// cap assumed free.  The caller then leaves the loop; ST_RELAY_TIMEOUT tells the host to finish the relay with
// per-round launches (entropy_relay_sparse), which need no co-residency.  A decode never hangs on this barrier.
__device__ __forceinline__ bool grid_barrier(uint32_t *counter, uint32_t &generation, uint32_t spin_limit, uint32_t *s_ok)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        ++generation;
        __threadfence();
        atomicAdd(counter, 1u);
        const uint32_t target = generation * gridDim.x;
        // Back off between polls: CTAs without work arrive at once, and a thousand of them re-reading one
        // L2 line back to back slow the memory requests of the CTAs that are still decoding.
        uint32_t ns = 128, polls = 0;
        bool ok = true;
        while (load_acquire_gpu(counter) < target) {
            if (++polls > spin_limit) {
                ok = false;
                break;
            }
            __nanosleep(ns);
            ns = ns < 1024u ? ns * 2u : ns;
        }
        __threadfence();
        *s_ok = ok ? 1u : 0u;
    }
    __syncthreads();
    return *s_ok != 0u;
}

__global__ void __launch_bounds__(ENTROPY_THREADS, 8) entropy_relay_loop_kernel(EntropyArgs a, int first, int last,
                                                                             uint32_t wlog, uint32_t spin_limit)
{
    __shared__ uint32_t s_bar_ok;
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_raw);
    const uint32_t nsub = a.meta->nsub, total_bits = a.meta->total_bits;
    k1_stage_tables(sm, a);
    __syncthreads();
    const SmemLuts L = k1_luts(sm, a);
    StreamView S{a.seg_bit, total_bits};
    const uint32_t stride = (1u << wlog) + 5u;
    uint32_t *mine = sm.words + threadIdx.x * stride;
    PrivateWords W;
    W.addr = (uint32_t)__cvta_generic_to_shared(mine);
    const uint32_t total_words = ((total_bits + 31u) >> 5) + 4u;
    uint32_t generation = 0;
    int round = first;
    // A round whose list fits one CTA is run by CTA 0 alone (the others leave): the tail rounds then cost a
    // __syncthreads() each instead of a grid barrier.
    constexpr uint32_t SOLO_ITEMS = ENTROPY_THREADS;
    bool solo = false;
    for (; round <= last; ++round) {
        const uint32_t count = load_acquire_gpu(&a.meta->changed[relay_slot_dev(round - 1)]);
        if (count == 0u)
            break; // fixed point (uniform across the grid: read after the barrier)
        if (!solo && count <= SOLO_ITEMS) {
            if (blockIdx.x != 0u)
                return;
            solo = true;
        }
        const uint32_t *list_in = a.worklist[(round - 1) & 1];
        uint32_t *list_out = a.worklist[round & 1];
        uint32_t *count_out = &a.meta->changed[relay_slot_dev(round)];
        const uint32_t w0 = solo ? threadIdx.x : blockIdx.x * ENTROPY_THREADS + threadIdx.x;
        const uint32_t wstep = solo ? ENTROPY_THREADS : gridDim.x * ENTROPY_THREADS;
        for (uint32_t w = w0; w < count; w += wstep) {
            const uint32_t sub = __ldcg(list_in + w);
            if (sub >= nsub)
                continue;
            const uint4 inraw = __ldcg(reinterpret_cast<const uint4 *>(&a.state[sub - 1]));
            const uint32_t j0 = inraw.x >> 5;
            for (uint32_t k = 0; k < stride - 1u; ++k)
                mine[k] = j0 + k < total_words ? __ldg(a.words + j0 + k) : 0u;
            W.gw0 = j0;
            const uint32_t end = min((sub + 1u) << (wlog + 5u), total_bits);
            const SubState out = relay_decode(a, W, L, S, sub, end, inraw.x, inraw.z, true);
            relay_publish(a, sub, out, nsub, list_out, count_out);
        }
        if (solo) {
            __threadfence();
            __syncthreads();
        } else if (!grid_barrier(&a.meta->grid_bar, generation, spin_limit, &s_bar_ok)) {
            // round `round` may be incomplete (CTAs that never became resident have not run their share): the host
            // repeats it -- items done twice change nothing and append nothing -- and goes on from there
            if (threadIdx.x == 0)
                atomicOr(&a.meta->status, ST_RELAY_TIMEOUT);
            return;
        }
        // last COMPLETE round, as seen by ANY CTA (after a timeout the CTAs disagree on how far they got; the host
        // repeats the round after the furthest one, the only one that can be partly done)
        if (threadIdx.x == 0)
            atomicMax(&a.meta->relay_rounds, (uint32_t)round);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicMax(&a.meta->relay_rounds, (uint32_t)(round <= last ? round - 1 : last));
}

// Segmented exclusive scan of the slot counts: start_slot[i] = absolute coefficient slot at the
// entry of subsequence i.  An element that crossed a segment boundary carries an absolute value
// (flag f = 1), otherwise a relative count.  Three phases: per-tile aggregate, one-block scan of the
// aggregates, per-tile scan seeded with the tile's carry.
struct SegVal {
    uint32_t f, v;
};

__device__ __forceinline__ SegVal seg_combine(const SegVal &a, const SegVal &b) // a then b
{
    SegVal r;
    r.f = a.f | b.f;
    r.v = b.f ? b.v : a.v + b.v;
    return r;
}

constexpr int SCAN_THREADS = 1024;

// inclusive scan across the block; s_w must hold SCAN_THREADS/32 entries
__device__ __forceinline__ SegVal block_seg_scan(SegVal x, SegVal *s_w, SegVal &aggregate)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        SegVal o;
        o.f = __shfl_up_sync(0xffffffffu, x.f, d);
        o.v = __shfl_up_sync(0xffffffffu, x.v, d);
        if (lane >= d)
            x = seg_combine(o, x);
    }
    if (lane == 31)
        s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        SegVal w = s_w[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            SegVal o;
            o.f = __shfl_up_sync(0xffffffffu, w.f, d);
            o.v = __shfl_up_sync(0xffffffffu, w.v, d);
            if (lane >= d)
                w = seg_combine(o, w);
        }
        s_w[lane] = w;
    }
    __syncthreads();
    if (warp > 0)
        x = seg_combine(s_w[warp - 1], x);
    aggregate = s_w[SCAN_THREADS / 32 - 1];
    return x;
}

__device__ __forceinline__ SegVal load_seg_val(const EntropyArgs &a, uint32_t i, uint32_t nsub)
{
    SegVal e;
    e.f = 0;
    e.v = 0;
    if (i < nsub) {
        const SubState st = a.state[i];
        e.f = st.seg >= 0 ? 1u : 0u;
        e.v = st.n + (e.f ? seg_slot_base(a.g, (uint32_t)st.seg) : 0u);
    }
    return e;
}

__global__ void __launch_bounds__(SCAN_THREADS) entropy_scan_reduce_kernel(EntropyArgs a)
{
    __shared__ SegVal s_w[SCAN_THREADS / 32];
    const uint32_t nsub = a.meta->nsub;
    if (blockIdx.x * SCAN_THREADS >= nsub)
        return;
    SegVal agg;
    block_seg_scan(load_seg_val(a, blockIdx.x * SCAN_THREADS + threadIdx.x, nsub), s_w, agg);
    if (threadIdx.x == 0)
        a.scan_tiles[blockIdx.x] = make_uint2(agg.f, agg.v);
}

// one block: scan_tiles[t] <- value carried INTO tile t; also the final slot of the stream
__global__ void __launch_bounds__(SCAN_THREADS) entropy_scan_tiles_kernel(EntropyArgs a)
{
    __shared__ SegVal s_w[SCAN_THREADS / 32];
    __shared__ SegVal s_carry;
    __shared__ SegVal s_incl[SCAN_THREADS];
    const uint32_t nsub = a.meta->nsub;
    const uint32_t ntiles = (nsub + SCAN_THREADS - 1) / SCAN_THREADS;
    if (threadIdx.x == 0) {
        s_carry.f = 1u; // virtual element -1: absolute slot 0 (segment 0 starts at bit 0)
        s_carry.v = 0u;
    }
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntiles; t0 += SCAN_THREADS) {
        const uint32_t t = t0 + threadIdx.x;
        SegVal e;
        e.f = 0;
        e.v = 0;
        if (t < ntiles) {
            const uint2 r = a.scan_tiles[t];
            e.f = r.x;
            e.v = r.y;
        }
        SegVal agg;
        const SegVal incl = seg_combine(s_carry, block_seg_scan(e, s_w, agg));
        s_incl[threadIdx.x] = incl;
        __syncthreads();
        if (t < ntiles) {
            const SegVal ex = threadIdx.x == 0 ? s_carry : s_incl[threadIdx.x - 1];
            a.scan_tiles[t] = make_uint2(ex.f, ex.v);
            if (t + 1u == ntiles)
                a.meta->final_slot = incl.v;
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1)
            s_carry = incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) entropy_scan_apply_kernel(EntropyArgs a)
{
    __shared__ SegVal s_w[SCAN_THREADS / 32];
    const uint32_t nsub = a.meta->nsub;
    if (blockIdx.x * SCAN_THREADS >= nsub)
        return;
    const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
    SegVal agg;
    SegVal incl = block_seg_scan(load_seg_val(a, i, nsub), s_w, agg);
    const uint2 c = a.scan_tiles[blockIdx.x];
    SegVal carry;
    carry.f = c.x;
    carry.v = c.y;
    incl = seg_combine(carry, incl);
    // incl = absolute slot at the END of subsequence i == entry of subsequence i+1
    if (i + 1u < nsub)
        a.start_slot[i + 1u] = incl.v;
    if (i == 0u)
        a.start_slot[0] = 0u;
}

// ---- final pass -------------------------------------------------------------------------------------
// A tile of WRITE_THREADS consecutive subsequences owns a contiguous range of coefficient slots
// [start_slot[first], start_slot[first of next tile]).  The CTA assembles that range in shared
// memory, WRITE_WIN_BLOCKS blocks at a time (threads whose next symbol lies beyond the window pause
// and resume in the next one), and flushes every window with full 128-byte lines -- so the
// coefficient buffer needs no zero-fill and HBM sees no partial-sector read-modify-write.  Only the
// first and last block of a tile can be shared with a neighbouring tile; those are written
// element-wise, each tile touching exactly its own slots.
constexpr int WRITE_THREADS = 64;
#ifndef KPEG_WIN_BLOCKS
#define KPEG_WIN_BLOCKS 256
#endif
constexpr int WRITE_WIN_BLOCKS = KPEG_WIN_BLOCKS;

struct WriteSmemTail {
    int16_t obuf[WRITE_WIN_BLOCKS * 64]; // slot 0 of a block carries its DC difference until the flush
};

__host__ __device__ inline size_t k1_write_words_bytes(uint32_t sub_bits)
{
    const size_t w = (size_t)(k1_pad(WRITE_THREADS * (sub_bits / 32u) + K1_TAIL_WORDS) + 2u) * sizeof(uint32_t);
    return (w + 15u) & ~(size_t)15u;
}

// byte offset of the window buffers inside the dynamic shared memory block (16-byte aligned)
__host__ __device__ inline size_t k1_write_tail_offset(uint32_t sub_bits)
{
    return (offsetof(K1Smem, words) + k1_write_words_bytes(sub_bits) + 15u) & ~(size_t)15u;
}

__host__ __device__ inline size_t k1_write_smem_bytes(uint32_t sub_bits)
{
    return k1_write_tail_offset(sub_bits) + sizeof(WriteSmemTail);
}

struct SmemSink {
    static constexpr bool bounds_itself = true; // put() drops whatever falls outside the window
    uint32_t obuf_addr; // shared byte address of the window
    uint32_t slot0;     // first slot of the window
    // One predicated store, one address formula: an AC coefficient goes to slot (slot + adv - 1); a DC difference
    // arrives with z == 0 and advance 1, i.e. the same formula puts it into slot 0 of its block -- the slot the
    // coefficient buffer leaves to K2's integrated DC value.  The window flush copies slot 0 of every block to dcdiff[].
    __device__ __forceinline__ void put(bool, uint32_t slot, uint32_t adv, bool valid, int32_t v) const
    {
        const uint32_t off = slot + adv - 1u - slot0;
        if (valid && off < (uint32_t)(WRITE_WIN_BLOCKS * 64))
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(obuf_addr + 2u * off), "h"((uint16_t)v));
    }
};

// Flush window blocks [wb, min(wb + WRITE_WIN_BLOCKS, b_end)): whole 128-byte lines for the blocks the
// tile owns entirely, single elements for the (at most two) blocks shared with a neighbouring tile,
// DC differences of the blocks whose first slot the tile owns.
__device__ __forceinline__ void write_flush_window(const EntropyArgs &a, const WriteSmemTail &tail, uint32_t wb,
                                                   uint32_t b_first, uint32_t b_end, uint32_t s_begin, uint32_t s_end,
                                                   uint32_t t)
{
    const uint32_t we = min(wb + (uint32_t)WRITE_WIN_BLOCKS, b_end);
    if (wb >= we)
        return;
    const uint32_t nblk = we - wb;
    const uint4 *src = reinterpret_cast<const uint4 *>(tail.obuf);
    uint4 *dst = reinterpret_cast<uint4 *>(a.coef) + (size_t)wb * 8u;
    for (uint32_t ci = t; ci < nblk * 8u; ci += WRITE_THREADS) {
        const uint32_t b = wb + (ci >> 3);
        const bool full = (b << 6) >= s_begin && ((b + 1u) << 6) <= s_end;
        if (full)
            dst[ci] = src[ci];
    }
    if (wb == b_first && (s_begin & 63u)) {
        const uint32_t pos = (b_first << 6) + t; // WRITE_THREADS == 64 == slots per block
        if (pos >= s_begin && pos < s_end)
            a.coef[pos] = tail.obuf[t];
    }
    if (we == b_end && (s_end & 63u) && !((b_end - 1u) == b_first && (s_begin & 63u) && wb == b_first)) {
        const uint32_t pos = ((b_end - 1u) << 6) + t;
        if (pos >= s_begin && pos < s_end)
            a.coef[pos] = tail.obuf[((b_end - 1u - wb) << 6) + t];
    }
    for (uint32_t i = t; i < nblk; i += WRITE_THREADS) {
        const uint32_t b = wb + i;
        if ((b << 6) >= s_begin && (b << 6) < s_end)
            a.dcdiff[b] = tail.obuf[i << 6];
    }
}

__global__ void __launch_bounds__(WRITE_THREADS) entropy_write_kernel(EntropyArgs a, uint32_t wlog)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_raw);
    WriteSmemTail &tail = *reinterpret_cast<WriteSmemTail *>(k1_raw + k1_write_tail_offset(a.g.sub_bits));
    const uint32_t nsub = a.meta->nsub, total_bits = a.meta->total_bits;
    const uint32_t ntiles = (nsub + WRITE_THREADS - 1) / WRITE_THREADS;
    if (blockIdx.x >= ntiles)
        return;
    k1_stage_tables(sm, a);
    const SmemLuts L = k1_luts(sm, a);
    StreamView S{a.seg_bit, total_bits};
    const uint32_t total_slots = a.g.total_blocks * 64u;
    const uint32_t t = threadIdx.x;
    uint32_t st = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        k1_stage_stream<WRITE_THREADS>(sm, a, tile, total_bits, wlog);
        const SmemWords W = k1_words<WRITE_THREADS>(sm, tile, wlog);
        const uint32_t sub0 = tile * WRITE_THREADS, sub = sub0 + t;
        // slot range owned by the tile (uniform)
        const uint32_t s_begin = a.start_slot[sub0];
        uint32_t s_end = (sub0 + WRITE_THREADS < nsub) ? a.start_slot[sub0 + WRITE_THREADS] : a.meta->final_slot;
        s_end = min(s_end, total_slots);
        const uint32_t b_first = s_begin >> 6;
        const uint32_t b_end = s_end > s_begin ? (s_end + 63u) >> 6 : b_first; // blocks [b_first, b_end)

        DecState d;
        bool done = true;
        uint32_t end = 0;
        if (sub < nsub) {
            uint32_t p = 0, c = 0, z = 0;
            if (sub) {
                const SubState in = a.state[sub - 1];
                p = in.p;
                c = in.cz >> 8;
                z = in.cz & 0xFFu;
            }
            const uint32_t slot = a.start_slot[sub];
            if ((slot & 63u) != z || ((slot >> 6) % a.g.ncomp) != c)
                st |= ST_EXIT_MISMATCH;
            end = min((sub + 1u) << (wlog + 5u), total_bits);
            dec_init(d, W, S, p, c, z, a.seg_hint[sub], slot);
            done = false;
        }

        for (uint32_t wb = b_first;; wb += WRITE_WIN_BLOCKS) {
            // zero the window
            {
                uint4 *o = reinterpret_cast<uint4 *>(&tail);
                constexpr int N16 = (int)(sizeof(WriteSmemTail) / 16);
#pragma unroll 4
                for (int i = t; i < N16; i += WRITE_THREADS)
                    o[i] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
            if (!done) {
                SmemSink sink;
                sink.obuf_addr = (uint32_t)__cvta_generic_to_shared(tail.obuf);
                sink.slot0 = wb << 6;
                // past the tile's own range (corrupt stream): finish in one go, writes fall outside the window
                const uint32_t limit = wb < b_end ? (wb + WRITE_WIN_BLOCKS) << 6 : 0xFFFFFFFFu;
                decode_run<true, false>(d, W, L, S, a.g, end, limit, sink, NoRecorder{});
                done = d.p >= end;
            }
            const bool all_done = __syncthreads_and(done);
            write_flush_window(a, tail, wb, b_first, b_end, s_begin, s_end, t);
            if (all_done)
                break;
            __syncthreads(); // the window is zeroed again
        }
        if (sub < nsub) {
            st |= d.st;
            const SubState rec = a.state[sub];
            const SubState out = dec_exit_state(d);
            if (out.p != rec.p || out.cz != rec.cz)
                st |= ST_EXIT_MISMATCH;
            if (sub + 1u == nsub && a.meta->final_slot < total_slots)
                st |= ST_SEG_MISMATCH; // the stream ended before the last MCU
        }
        __syncthreads();
    }
    if (st)
        atomicOr(&a.meta->status, st);
}

// Final pass over the symbol records (see entropy_core.h): no Huffman decode, no stream, no tables --
// a thread walks its subsequence's records (coalesced k-major loads, independent of each other, so
// several are in flight), tracks (slot, zig-zag index) and drops values into the shared-memory window.
struct GlobalRecAt {
    const uint32_t *base;
    uint32_t stride_bytes; // 128 in the warp-interleaved list, 4 in a private area: one IMAD.WIDE per address
    __device__ __forceinline__ uint32_t operator()(uint32_t k) const
    {
        return __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(base) + (size_t)k * stride_bytes));
    }
};

__global__ void __launch_bounds__(WRITE_THREADS) entropy_expand_kernel(EntropyArgs a)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    WriteSmemTail &tail = *reinterpret_cast<WriteSmemTail *>(k1_raw);
    const uint32_t nsub = a.meta->nsub;
    const uint32_t ntiles = (nsub + WRITE_THREADS - 1) / WRITE_THREADS;
    const uint32_t total_slots = a.g.total_blocks * 64u;
    const uint32_t t = threadIdx.x;
    uint32_t st = 0;
    // shared addresses of the window: made opaque, or the compiler re-derives them from the CTA's shared window
    // base (S2R + LEA) for every record instead of keeping two registers
    uint32_t win_obuf = (uint32_t)__cvta_generic_to_shared(tail.obuf);
    asm volatile("" : "+r"(win_obuf));
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t sub0 = tile * WRITE_THREADS, sub = sub0 + t;
        const uint32_t s_begin = a.start_slot[sub0];
        uint32_t s_end = (sub0 + WRITE_THREADS < nsub) ? a.start_slot[sub0 + WRITE_THREADS] : a.meta->final_slot;
        s_end = min(s_end, total_slots);
        const uint32_t b_first = s_begin >> 6;
        const uint32_t b_end = s_end > s_begin ? (s_end + 63u) >> 6 : b_first;

        uint32_t k = 0, n = 0, slot = 0, z = 0, expect = 0;
        bool done = true;
        GlobalRecAt R;
        R.base = a.rec + rec_base_index(sub, a.rec_kmax);
        R.stride_bytes = 128u;
        if (sub < nsub) {
            const uint32_t nr = a.nrec[sub];
            n = min(nr & NREC_MASK, a.rec_kmax);
            if (nr >> 10) { // redone in a sparse relay round: private contiguous area
                R.base = a.rec_alt + (size_t)((nr >> 10) - 1u) * a.rec_kmax;
                R.stride_bytes = 4u;
            }
            slot = a.start_slot[sub];
            z = sub ? (a.state[sub - 1].cz & 0xFFu) : 0u;
            if ((slot & 63u) != z)
                st |= ST_EXIT_MISMATCH;
            expect = (sub + 1u < nsub) ? a.start_slot[sub + 1u] : a.meta->final_slot;
            done = false;
        }
        for (uint32_t wb = b_first;; wb += WRITE_WIN_BLOCKS) {
            {
                uint4 *o = reinterpret_cast<uint4 *>(&tail);
                constexpr int N16 = (int)(sizeof(WriteSmemTail) / 16);
#pragma unroll 4
                for (int i = t; i < N16; i += WRITE_THREADS)
                    o[i] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
            if (!done) {
                SmemSink sink;
                sink.obuf_addr = win_obuf;
                sink.slot0 = wb << 6;
                const uint32_t limit = wb < b_end ? (wb + WRITE_WIN_BLOCKS) << 6 : 0xFFFFFFFFu;
                expand_run(k, n, slot, z, st, R, a.g, limit, sink);
                done = k >= n;
            }
            const bool all_done = __syncthreads_and(done);
            write_flush_window(a, tail, wb, b_first, b_end, s_begin, s_end, t);
            if (all_done)
                break;
            __syncthreads();
        }
        if (sub < nsub) {
            if (slot != expect)
                st |= ST_EXIT_MISMATCH;
            if (sub + 1u == nsub && a.meta->final_slot < total_slots)
                st |= ST_SEG_MISMATCH; // the stream ended before the last MCU
        }
        __syncthreads();
    }
    if (st)
        atomicOr(&a.meta->status, st);
}

static uint32_t ilog2(uint32_t v)
{
    uint32_t l = 0;
    while ((1u << l) < v)
        ++l;
    return l;
}

static uint32_t g_k1_grid_cap = 148 * 8;
static uint32_t g_k1_write_grid_cap = 148 * 4;

static uint32_t k1_grid(const EntropyArgs &a)
{
    const uint32_t tiles = (a.nsub_max + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
    return tiles < g_k1_grid_cap ? tiles : g_k1_grid_cap;
}

void launch_entropy_cold(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
    entropy_cold_kernel<<<k1_grid(a), ENTROPY_THREADS, k1_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    ++*launches;
}

void launch_entropy_relay(const EntropyArgs &a, int round, cudaStream_t s, uint32_t *launches)
{
    if (round == 1) {
        const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
        entropy_relay_full_kernel<<<k1_grid(a), ENTROPY_THREADS, k1_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    } else {
        const uint32_t grid = (a.nsub_max + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
        const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
        entropy_relay_sparse_kernel<<<grid, ENTROPY_THREADS, k1_sparse_smem_bytes(a.g.sub_bits), s>>>(
            a, round, relay_slot(round - 1), relay_slot(round), wlog);
    }
    ++*launches;
}

static uint32_t g_sm_count = 148;
static uint32_t g_max_concurrent_loops = 8;
static std::atomic<int> g_live_contexts{0};

void kernels_context_created() { ++g_live_contexts; }
void kernels_context_destroyed() { --g_live_contexts; }
// polls (~1 us each) a CTA waits at the relay loop's grid barrier before it gives up: ~2 s
static uint32_t g_relay_spin_limit = 2000000u;
static int g_relay_loop_per_sm_cap = 0; // experiments: cap of the cooperative relay loop's grid in CTAs per SM (0 = default rule)

cudaError_t launch_entropy_relay_loop(const EntropyArgs &a, int first, int last, cudaStream_t s, uint32_t *launches)
{
    uint32_t wlog = ilog2(a.g.sub_bits / 32u);
    EntropyArgs args = a;
    uint32_t spin_limit = g_relay_spin_limit;
    void *params[] = {&args, &first, &last, &wlog, &spin_limit};
    const size_t smem = k1_sparse_smem_bytes(a.g.sub_bits);
    int per_sm = 2;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_relay_loop_kernel, ENTROPY_THREADS, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    // every lane of a context may be inside its loop at the same time: together they must stay co-resident,
    // or CTAs spinning at one loop's barrier could keep another loop's CTAs from ever being scheduled
    uint32_t cap = g_sm_count * (uint32_t)per_sm / (g_max_concurrent_loops * (uint32_t)std::max(1, g_live_contexts.load()));
    if (g_relay_loop_per_sm_cap > 0) // experiments only
        cap = g_sm_count * (uint32_t)std::min(per_sm, g_relay_loop_per_sm_cap);
    cap = cap < 1u ? 1u : cap;
    // Round 2 redoes a few percent of the subsequences (~6 % at 4K q95) and every later round far fewer, each
    // at the latency of one serial subsequence decode: a small grid is enough, makes the grid barrier cheap,
    // leaves the SMs to the other lanes' kernels, and keeps the sum of the lanes' cooperative grids far
    // below what is co-resident (so concurrent loops can never wait on each other's CTAs).
    uint32_t grid = a.nsub_max / (ENTROPY_THREADS * 12u) + 1u;
    grid = grid < 16u ? 16u : grid;
    grid = grid > cap ? cap : grid;
    const cudaError_t e =
        cudaLaunchCooperativeKernel((const void *)entropy_relay_loop_kernel, dim3(grid), dim3(ENTROPY_THREADS), params, smem, s);
    if (e != cudaSuccess) {
        cudaGetLastError(); // not sticky: the caller issues the rounds as separate launches instead
        return e;
    }
    ++*launches;
    return cudaSuccess;
}

void launch_entropy_scan(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t tiles = (a.nsub_max + SCAN_THREADS - 1) / SCAN_THREADS;
    entropy_scan_reduce_kernel<<<tiles, SCAN_THREADS, 0, s>>>(a);
    entropy_scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>(a);
    entropy_scan_apply_kernel<<<tiles, SCAN_THREADS, 0, s>>>(a);
    *launches += 3;
}

static uint32_t g_k1_expand_grid_cap = 148 * 6;

void launch_entropy_expand(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t tiles = (a.nsub_max + WRITE_THREADS - 1) / WRITE_THREADS;
    const uint32_t grid = tiles < g_k1_expand_grid_cap ? tiles : g_k1_expand_grid_cap;
    entropy_expand_kernel<<<grid, WRITE_THREADS, sizeof(WriteSmemTail), s>>>(a);
    ++*launches;
}

void launch_entropy_write(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
    const uint32_t tiles = (a.nsub_max + WRITE_THREADS - 1) / WRITE_THREADS;
    const uint32_t grid = tiles < g_k1_write_grid_cap ? tiles : g_k1_write_grid_cap;
    entropy_write_kernel<<<grid, WRITE_THREADS, k1_write_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    ++*launches;
}

// =================================================================================================
// K2: DC prediction = segmented prefix sum of the DC differences (MCU.cpp:107-108; predictor reset at
// every restart interval / image start, T.81 F.2.1.3.1 -- the reference never resets)
// =================================================================================================

__device__ __forceinline__ bool mcu_is_reset(const JobGeom &g, uint32_t m)
{
    const uint32_t mi = m % g.mcus_per_image;
    return g.restart_interval ? (mi % g.restart_interval) == 0u : mi == 0u;
}

struct Dc3 {
    int32_t v[3];
    uint32_t f;
};

__device__ __forceinline__ Dc3 dc_combine(const Dc3 &a, const Dc3 &b) // a then b
{
    Dc3 r;
    r.f = a.f | b.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        r.v[c] = b.f ? b.v[c] : a.v[c] + b.v[c];
    return r;
}

__device__ __forceinline__ Dc3 dc_shfl_up(const Dc3 &x, int d)
{
    Dc3 r;
    r.f = __shfl_up_sync(0xffffffffu, x.f, d);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        r.v[c] = __shfl_up_sync(0xffffffffu, x.v[c], d);
    return r;
}

// Inclusive segmented scan across the block of one Dc3 per thread; returns the inclusive value and
// the block aggregate.
__device__ __forceinline__ Dc3 dc_block_scan(Dc3 x, Dc3 *s_w /*[DC_THREADS/32]*/, Dc3 &aggregate)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Dc3 o = dc_shfl_up(x, d);
        if (lane >= d)
            x = dc_combine(o, x);
    }
    if (lane == 31)
        s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        Dc3 w;
        if (lane < DC_THREADS / 32)
            w = s_w[lane];
        else {
            w.f = 0;
            w.v[0] = w.v[1] = w.v[2] = 0;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Dc3 o = dc_shfl_up(w, d);
            if (lane >= d)
                w = dc_combine(o, w);
        }
        if (lane < DC_THREADS / 32)
            s_w[lane] = w;
    }
    __syncthreads();
    if (warp > 0)
        x = dc_combine(s_w[warp - 1], x);
    aggregate = s_w[DC_THREADS / 32 - 1];
    return x;
}

// The thread's DC_MCUS_PER_THREAD consecutive MCUs: their DC differences (as 8-byte words when the run is whole and
// aligned) and the running inclusive values.  The reset predicate (mcu_is_reset) is evaluated once with two
// divisions and then stepped.
__device__ __forceinline__ Dc3 dc_thread_local(const DcArgs &a, uint32_t m0, uint32_t total_mcus, Dc3 incl[DC_MCUS_PER_THREAD])
{
    static_assert(DC_MCUS_PER_THREAD == 4, "the 8-byte access pattern below is written for four MCUs per thread");
    const uint32_t nc = a.g.ncomp;
    int16_t v[DC_MCUS_PER_THREAD * 3];
#pragma unroll
    for (int i = 0; i < DC_MCUS_PER_THREAD * 3; ++i)
        v[i] = 0;
    const int16_t *src = a.dcdiff + (size_t)m0 * nc;
    const bool whole = m0 + DC_MCUS_PER_THREAD <= total_mcus && (reinterpret_cast<uintptr_t>(src) & 7u) == 0;
    if (whole && nc == 3) {
        uint2 w[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            w[i] = __ldg(reinterpret_cast<const uint2 *>(src) + i);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const uint32_t x = (i & 2) ? w[i >> 2].y : w[i >> 2].x;
            v[i] = (int16_t)((i & 1) ? (x >> 16) : (x & 0xFFFFu));
        }
    } else if (whole && nc == 1) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(src));
        v[0] = (int16_t)(w.x & 0xFFFFu), v[3] = (int16_t)(w.x >> 16), v[6] = (int16_t)(w.y & 0xFFFFu), v[9] = (int16_t)(w.y >> 16);
    } else {
#pragma unroll
        for (int k = 0; k < DC_MCUS_PER_THREAD; ++k)
            if (m0 + k < total_mcus)
                for (uint32_t c = 0; c < nc; ++c)
                    v[k * 3 + c] = src[k * nc + c];
    }
    uint32_t mi = m0 % a.g.mcus_per_image;
    uint32_t ri = a.g.restart_interval ? mi % a.g.restart_interval : mi;
    Dc3 run;
    run.f = 0;
    run.v[0] = run.v[1] = run.v[2] = 0;
#pragma unroll
    for (int k = 0; k < DC_MCUS_PER_THREAD; ++k) {
        Dc3 e;
        e.f = 0;
        e.v[0] = e.v[1] = e.v[2] = 0;
        if (m0 + k < total_mcus) {
            e.f = ri == 0u ? 1u : 0u; // == mcu_is_reset(a.g, m0 + k)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                e.v[c] = v[k * 3 + c];
        }
        run = dc_combine(run, e);
        incl[k] = run;
        ++mi, ++ri;
        if (mi == a.g.mcus_per_image)
            mi = 0, ri = 0;
        else if (a.g.restart_interval && ri == a.g.restart_interval)
            ri = 0;
    }
    return run;
}

__global__ void __launch_bounds__(DC_THREADS) dc_reduce_kernel(DcArgs a)
{
    __shared__ Dc3 s_w[DC_THREADS / 32];
    const uint32_t total_mcus = a.g.nimages * a.g.mcus_per_image;
    const uint32_t m0 = (blockIdx.x * DC_THREADS + threadIdx.x) * DC_MCUS_PER_THREAD;
    Dc3 incl[DC_MCUS_PER_THREAD];
    const Dc3 mine = dc_thread_local(a, m0, total_mcus, incl);
    Dc3 agg;
    dc_block_scan(mine, s_w, agg);
    if (threadIdx.x == 0) {
        a.tile_carry[blockIdx.x * 4 + 0] = agg.v[0];
        a.tile_carry[blockIdx.x * 4 + 1] = agg.v[1];
        a.tile_carry[blockIdx.x * 4 + 2] = agg.v[2];
        a.tile_carry[blockIdx.x * 4 + 3] = (int32_t)agg.f;
    }
}

// one block: tile_carry[t] <- exclusive segmented scan (value carried INTO tile t)
__global__ void __launch_bounds__(DC_THREADS) dc_tile_scan_kernel(DcArgs a)
{
    __shared__ Dc3 s_w[DC_THREADS / 32];
    __shared__ Dc3 s_carry;
    __shared__ Dc3 s_incl[DC_THREADS];
    if (threadIdx.x == 0) {
        s_carry.f = 0;
        s_carry.v[0] = s_carry.v[1] = s_carry.v[2] = 0;
    }
    __syncthreads();
    for (uint32_t t0 = 0; t0 < a.ntiles; t0 += DC_THREADS) {
        const uint32_t t = t0 + threadIdx.x;
        Dc3 e;
        e.f = 0;
        e.v[0] = e.v[1] = e.v[2] = 0;
        if (t < a.ntiles) {
            e.v[0] = a.tile_carry[t * 4 + 0];
            e.v[1] = a.tile_carry[t * 4 + 1];
            e.v[2] = a.tile_carry[t * 4 + 2];
            e.f = (uint32_t)a.tile_carry[t * 4 + 3];
        }
        Dc3 agg;
        const Dc3 incl = dc_block_scan(e, s_w, agg);
        const Dc3 carry = s_carry;
        const Dc3 full = dc_combine(carry, incl); // inclusive up to tile t
        // exclusive value = inclusive of t-1: shuffle through shared memory
        s_incl[threadIdx.x] = full;
        __syncthreads();
        if (t < a.ntiles) {
            const Dc3 ex = threadIdx.x == 0 ? carry : s_incl[threadIdx.x - 1];
            a.tile_carry[t * 4 + 0] = ex.v[0];
            a.tile_carry[t * 4 + 1] = ex.v[1];
            a.tile_carry[t * 4 + 2] = ex.v[2];
        }
        __syncthreads();
        if (threadIdx.x == DC_THREADS - 1)
            s_carry = full;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DC_THREADS) dc_apply_kernel(DcArgs a)
{
    __shared__ Dc3 s_w[DC_THREADS / 32];
    const uint32_t total_mcus = a.g.nimages * a.g.mcus_per_image;
    const uint32_t m0 = (blockIdx.x * DC_THREADS + threadIdx.x) * DC_MCUS_PER_THREAD;
    Dc3 incl[DC_MCUS_PER_THREAD];
    const Dc3 mine = dc_thread_local(a, m0, total_mcus, incl);
    Dc3 agg;
    const Dc3 inc = dc_block_scan(mine, s_w, agg);
    // exclusive prefix of this thread = inclusive of the previous thread (through shared memory)
    __shared__ Dc3 s_incl[DC_THREADS];
    s_incl[threadIdx.x] = inc;
    __syncthreads();
    Dc3 pre;
    pre.f = 0;
    pre.v[0] = a.tile_carry[blockIdx.x * 4 + 0];
    pre.v[1] = a.tile_carry[blockIdx.x * 4 + 1];
    pre.v[2] = a.tile_carry[blockIdx.x * 4 + 2];
    if (threadIdx.x > 0)
        pre = dc_combine(pre, s_incl[threadIdx.x - 1]);
    const uint32_t nc = a.g.ncomp;
    int16_t *dst = a.dc + (size_t)m0 * nc;
    Dc3 r[DC_MCUS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < DC_MCUS_PER_THREAD; ++k)
        r[k] = dc_combine(pre, incl[k]);
    const bool whole = m0 + DC_MCUS_PER_THREAD <= total_mcus && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0;
    auto h = [](int32_t x) { return (uint32_t)x & 0xFFFFu; };
    if (whole && nc == 3) { // 12 values = three 8-byte stores
        reinterpret_cast<uint2 *>(dst)[0] = make_uint2(h(r[0].v[0]) | (h(r[0].v[1]) << 16), h(r[0].v[2]) | (h(r[1].v[0]) << 16));
        reinterpret_cast<uint2 *>(dst)[1] = make_uint2(h(r[1].v[1]) | (h(r[1].v[2]) << 16), h(r[2].v[0]) | (h(r[2].v[1]) << 16));
        reinterpret_cast<uint2 *>(dst)[2] = make_uint2(h(r[2].v[2]) | (h(r[3].v[0]) << 16), h(r[3].v[1]) | (h(r[3].v[2]) << 16));
    } else if (whole && nc == 1) {
        reinterpret_cast<uint2 *>(dst)[0] = make_uint2(h(r[0].v[0]) | (h(r[1].v[0]) << 16), h(r[2].v[0]) | (h(r[3].v[0]) << 16));
    } else {
#pragma unroll
        for (int k = 0; k < DC_MCUS_PER_THREAD; ++k)
            if (m0 + k < total_mcus)
                for (uint32_t c = 0; c < nc; ++c)
                    dst[k * nc + c] = (int16_t)r[k].v[c];
    }
}

void launch_dc_scan(const DcArgs &a, cudaStream_t s, uint32_t *launches)
{
    dc_reduce_kernel<<<a.ntiles, DC_THREADS, 0, s>>>(a);
    dc_tile_scan_kernel<<<1, DC_THREADS, 0, s>>>(a);
    dc_apply_kernel<<<a.ntiles, DC_THREADS, 0, s>>>(a);
    *launches += 3;
}

// =================================================================================================
// K3: fused dequantise + de-zigzag + IDCT + level shift + colour conversion + interleaved store
// =================================================================================================
//
// idct_kernel: one CTA reconstructs a strip of IDCT_MCUS_PER_CTA consecutive MCUs (a contiguous chunk
// of the coefficient buffer, and -- when the strip does not wrap -- 8 contiguous runs of pixels).
//   stage 0  the copy engine brings the strip's coefficients (2-D TMA tile, 128-byte swizzle: the per-thread
//            128-byte reads of stage 1 are bank-conflict free) and the quantiser tables into shared memory
//   stage 1  one thread per 8x8 block, all 64 values in registers, two fp32 lanes per instruction (FADD2 / FMUL2 /
//            FFMA2): dequantise (AAN prescale folded into the quantiser), de-zigzag by register renaming, separable
//            fp32 IDCT, rounding, tie-band test (a 64-bit mask per block)
//   stage 2  per pixel row: YCbCr -> RGB on pixel pairs (fp32 with proven margin, double otherwise), pack, store;
//            pixels with a sample inside the tie band are ALSO appended to a global record list
// idct_patch_kernel: one thread per record re-evaluates the flagged samples in the reference's own
// operation order (exact_sample), redoes the colour conversion and rewrites the pixel.  Keeping this
// out of the fused kernel matters: it is a long, strictly serial double-precision chain on ~0.3 % of
// the pixels; inside the strip kernel it kept two of three warps waiting at a barrier.

__device__ __constant__ ZigZagTables c_zz = make_zigzag_tables();

template <int I>
struct ZzNat {
    static constexpr int value = zigzag_to_natural(I);
};

constexpr int IDCT_REC_CAP = 134;
constexpr int IDCT_REC_FIXED = 16;          // slots of the global record list every strip owns
constexpr uint32_t TIE_EMPTY = 0xFFFFFFFFu; // pixel index of an unused slot
constexpr uint32_t IDCT_PREFETCH_AHEAD = 148u * 8u; // strips resident on the device at a time
#ifndef KPEG_IDCT_MIN_CTAS
#define KPEG_IDCT_MIN_CTAS 8
#endif

// Per-block facts the colour stage needs, derived from A = sum |dequantised coefficient| (every sample of the
// block is bounded by A / 4: the 64 basis functions are bounded by 1/4):
//   BLK_NONZERO  some coefficient is non-zero (otherwise all 64 samples are 0)
//   BLK_WIDE     A > COLOUR_SAFE_A: samples may leave the range the fp32 colour path is proven for
constexpr uint32_t BLK_NONZERO = 1u, BLK_WIDE = 2u;
constexpr float COLOUR_SAFE_A = 4.0f * (COLOUR_FAST_RANGE - 8.0f);

// Bit layout of a block's 64-bit tie mask (x = rows 0..3, y = rows 4..7): the bit of sample (row, col) in its word.
// The upper 16 bits hold columns 0..3, the lower 16 columns 4..7; inside, earlier samples sit higher.
__host__ __device__ constexpr int tie_bit(int row, int col) { return ((col & 4) ? 0 : 16) + 15 - (4 * (row & 3) + (col & 3)); }
// bits of one row in its word
__host__ __device__ constexpr uint32_t tie_row_bits(int row) { return (0xFu << (28 - 4 * (row & 3))) | (0xFu << (12 - 4 * (row & 3))); }
// inverse: bit index b of word w (0 = rows 0..3, 1 = rows 4..7) -> sample index row * 8 + col
__device__ __forceinline__ int tie_sample(int w, int b)
{
    const int k = 15 - (b & 15);
    return (4 * w + (k >> 2)) * 8 + ((b & 16) ? 0 : 4) + (k & 3);
}

template <int NC>
struct IdctSmem {
    static constexpr int NM = IDCT_MCUS_PER_CTA;
    static constexpr int NB = NM * NC;
    // The strip's coefficients (16-byte chunks, chunk k of block b at b*8 + (k ^ (b & 7))) are dead once
    // every thread has pulled its block into registers; the rounded samples then reuse the space.
    union {
        uint4 coef[NB * 8];
        float4 samp[NC * 8 * 2 * NM]; // [comp][row][half][mcu] -> 4 samples (rounded, unshifted)
    };
    float2 qpair[NC][32];         // prescaled quantisers in the pair order of the transform (pair_nat)
    float2 qdc[NC][32];           // the same with every AC entry zero: what a block that loses its AC terms (F1) multiplies by
    uint2 tie[NB];                // per block: 64-bit mask of the samples inside the tie band
    uint8_t flag[NB];             // per block: BLK_NONZERO | BLK_WIDE (decides the colour variant of its MCU)
    uint2 rec[IDCT_REC_CAP];      // tie records of this strip (compact form): the first IDCT_REC_FIXED go to the strip's own slots of the global list
    uint32_t nrec, rec_base;
    uint32_t img0, by0, bx0;      // image / block row / block column of the strip's first MCU (divisions done once, by thread 0)
    uint32_t pad_;
    unsigned long long mbar;      // completion barrier of the stage-0 bulk copies
};

// ---- two fp32 lanes per instruction (sm_100a FADD2 / FMUL2 / FFMA2) -------------------------------
// K3 is bound by instruction issue, not by HBM or by the FMA pipe, and most of what it issues are fp32 adds of
// the IDCT butterflies.  Blackwell's packed fp32 instructions do two independent IEEE lanes per issue slot, so the
// transform, the dequantisation, the rounding and the colour arithmetic run on register pairs.  A pair is a
// 64-bit register; packing / unpacking is register naming (mov.b64), not arithmetic.
struct F2 {
    unsigned long long v;
};
__device__ __forceinline__ F2 pack2(float lo, float hi)
{
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ float lo2(F2 a)
{
    float lo, hi;
    unpack2(a, lo, hi);
    return lo;
}
__device__ __forceinline__ float hi2(F2 a)
{
    float lo, hi;
    unpack2(a, lo, hi);
    return hi;
}
__device__ __forceinline__ F2 splat2(float k) { return pack2(k, k); }
__device__ __forceinline__ F2 lane_add(F2 a, F2 b)
{
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 lane_sub(F2 a, F2 b)
{
    F2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 lane_mul(F2 a, F2 b)
{
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 lane_fma(F2 a, F2 b, F2 c)
{
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 lane_mul_k(F2 a, float k) { return lane_mul(a, splat2(k)); }
__device__ __forceinline__ F2 lane_fma_k(F2 a, float k, F2 c) { return lane_fma(a, splat2(k), c); }

// coefficient I (zig-zag index) of a block held as eight 16-byte chunks
template <int I>
__device__ __forceinline__ float chunk_coef(const uint4 (&ch)[8])
{
    constexpr int k = I >> 3, j = (I & 7) >> 1, hi = I & 1;
    const uint32_t w = j == 0 ? ch[k].x : (j == 1 ? ch[k].y : (j == 2 ? ch[k].z : ch[k].w));
    return (float)(short)(hi ? (w >> 16) : (w & 0xFFFFu));
}

template <int NAT>
struct NatZz {
    static constexpr int value = make_zigzag_tables().nat2zz[NAT];
};

// Dequantise (AAN prescale folded into q; q arrives in pair order, see IdctSmem::qpair) and de-zigzag by register
// renaming; on the way, A = sum |c_i| * q_i, the magnitude the tie band is proportional to: |f| * (1 / prescale)
// with the reciprocal an immediate and the absolute value a free operand modifier -- one FFMA per coefficient.
// (fp32 accumulation is off by < 1e-5 relative; the band carries a 10 % margin.)
template <int... Ps>
__device__ __forceinline__ float dequant_dezigzag(const uint4 (&ch)[8], const float2 *qpair, F2 (&P)[32],
                                                  std::integer_sequence<int, Ps...>)
{
    float A[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // four short dependent chains instead of one of 64
    auto one = [&](auto PI) {
        constexpr int p = decltype(PI)::value;
        constexpr int n0 = pair_nat(p, 0), n1 = pair_nat(p, 1);
        const float2 q = qpair[p];
        P[p] = lane_mul(pack2(chunk_coef<NatZz<n0>::value>(ch), chunk_coef<NatZz<n1>::value>(ch)), pack2(q.x, q.y));
        A[p & 1] = fmaf(fabsf(lo2(P[p])), aan_unscale(n0), A[p & 1]);
        A[2 + (p & 1)] = fmaf(fabsf(hi2(P[p])), aan_unscale(n1), A[2 + (p & 1)]);
    };
    (one(std::integral_constant<int, Ps>{}), ...);
    return (A[0] + A[1]) + (A[2] + A[3]);
}

// Packed 2-D transform, first half: the row pass, two rows per instruction.
// P[rp * 8 + c] = rows pair_row(rp, 0), pair_row(rp, 1) at column c, in place.
__device__ __forceinline__ void idct_rows_packed(F2 (&P)[32])
{
#pragma unroll
    for (int rp = 0; rp < 4; ++rp)
        idct8_aan(P[rp * 8 + 0], P[rp * 8 + 1], P[rp * 8 + 2], P[rp * 8 + 3], P[rp * 8 + 4], P[rp * 8 + 5],
                  P[rp * 8 + 6], P[rp * 8 + 7]);
}

// Second half, one column: idct8_aan (idct_core.h) over the eight rows, same operations in the same order.
// With X = (v0, v1), Y = (v4, v7), X2 = (v2, v5), Y2 = (v6, v3) the first two butterfly stages of the even part
// (low lanes) and of the odd part (high lanes) are the same instruction, so they run packed; the rest is scalar on
// the halves.  Nothing is moved between registers.  v[r] = row r of this column.
__device__ __forceinline__ void idct8_column(F2 X, F2 Y, F2 X2, F2 Y2, float (&v)[8])
{
    const F2 S1 = lane_add(X, Y), D1 = lane_sub(X, Y);     // (t10, z11)  (t11, z12)
    const F2 S2 = lane_add(X2, Y2), D2 = lane_sub(X2, Y2); // (t13, z13)  (v2 - v6, z10)
    const F2 E = lane_add(S1, S2), F = lane_sub(S1, S2);   // (e0, o7)    (e3, z11 - z13)
    // even part
    const float t11 = lo2(D1), t13 = lo2(S2), e0 = lo2(E), e3 = lo2(F);
    const float n12 = lane_fma_k(lo2(D2), -1.414213562373095f, t13); // -(t12)
    const float e1 = lane_sub(t11, n12), e2 = lane_add(t11, n12);
    // odd part
    const float z12 = hi2(D1), z10 = hi2(D2), o7 = hi2(E);
    const float z5 = lane_mul_k(lane_add(z10, z12), 1.847759065022573f);
    const float t20 = lane_fma_k(z12, -1.082392200292394f, z5);
    const float t22 = lane_fma_k(z10, -2.613125929752753f, z5);
    const float o6 = lane_sub(t22, o7);
    const float n5 = lane_fma_k(hi2(F), -1.414213562373095f, o6); // -(o5)
    const float o4 = lane_add(t20, n5);
    v[0] = lane_add(e0, o7);
    v[7] = lane_sub(e0, o7);
    v[1] = lane_add(e1, o6);
    v[6] = lane_sub(e1, o6);
    v[2] = lane_sub(e2, n5);
    v[5] = lane_add(e2, n5);
    v[3] = lane_add(e3, o4);
    v[4] = lane_sub(e3, o4);
}

// Four ints -> four bytes with unsigned saturation (cvt.pack.sat: two values per instruction).
__device__ __forceinline__ uint32_t pack4_sat(int a, int b, int c, int d)
{
    // cvt.pack.sat.u8.s32.b32 d, x, y, z:  d = (z << 16) | (sat(x) << 8) | sat(y)
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(d), "r"(c));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;
}

// Rare path (G within 1e-3 of an integer, or samples out of the fp32-safe range): the reference's
// own double expression.  Out of line: eight unrolled call sites.
__device__ __noinline__ uint32_t colour_exact_px(float y, float cb, float cr)
{
    int R, G, B;
    ycc_to_rgb_exact((int)y, (int)cb, (int)cr, R, G, B); // clamped to [0,255]
    return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
}

// ---- exact pass ----------------------------------------------------------------------------------
// Tables of the exact path in shared memory (lane-varying indices would serialise in the constant cache).
struct ExactSmem {
    double cosd[8][8];
    float cc[8][8];
    int32_t qint[MAX_COMP][64];
    unsigned char nat2zz[64];
};

__device__ __forceinline__ void exact_smem_load(ExactSmem &es, const DeviceTables *t)
{
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        (&es.cosd[0][0])[i] = t->cosd[0][i];
        (&es.cc[0][0])[i] = t->cc[0][i];
        es.nat2zz[i] = c_zz.nat2zz[i];
    }
    for (int i = threadIdx.x; i < MAX_COMP * 64; i += blockDim.x)
        (&es.qint[0][0])[i] = t->qint[0][i];
    __syncthreads();
}

struct ExactCtx {
    const int16_t *coef, *dc, *dcdiff;
    uint8_t *pixels;
    uint32_t parity;
};

__device__ __forceinline__ ExactCtx exact_ctx(const IdctArgs &a)
{
    ExactCtx x;
    x.coef = a.coef;
    x.dc = a.dc;
    x.dcdiff = a.dcdiff;
    x.pixels = a.pixels;
    x.parity = a.g.flags & 1u;
    return x;
}

// coefficient I (zig-zag index) of a block held as eight 16-byte chunks, as an integer
template <int I>
__device__ __forceinline__ int chunk_coef_int(const uint4 (&ch)[8])
{
    constexpr int k = I >> 3, j = (I & 7) >> 1, hi = I & 1;
    const uint32_t w = j == 0 ? ch[k].x : (j == 1 ? ch[k].y : (j == 2 ? ch[k].z : ch[k].w));
    return (int)(short)(hi ? (w >> 16) : (w & 0xFFFFu));
}

// The 64 terms of the reference's sum (idct_core.h exact_sample, MCU.cpp:184-198) in ascending natural order
// (== u outer, v inner), fully unrolled: every index is a compile-time constant, the block stays in registers, and
// there is no branch and no load in the chain.  A zero coefficient contributes +-0, which leaves the float
// accumulator unchanged, so evaluating all 64 terms gives the same bits as skipping the zero ones.
template <int... Ns>
__device__ __forceinline__ float exact_terms(const uint4 (&ch)[8], const int32_t *q, const double (&cx)[8],
                                             const double (&cy)[8], float c00, float c01, std::integer_sequence<int, Ns...>)
{
    float sum = 0.0f;
    auto term = [&](auto N) {
        constexpr int nat = decltype(N)::value, zi = NatZz<nat>::value, u = nat >> 3, v = nat & 7;
        const int F = chunk_coef_int<zi>(ch) * q[zi];                                   // MCU.cpp:110-112, :115-120
        const float cc = (u == 0 && v == 0) ? c00 : ((u == 0 || v == 0) ? c01 : 1.0f); // Cu * Cv
        const float t = mul_f32(cc, (float)F);
        const double d = mul_f64(mul_f64((double)t, cx[u]), cy[v]);
        sum = (float)add_f64((double)sum, d); // float accumulator, rounded every term
    };
    (term(std::integral_constant<int, Ns>{}), ...);
    return sum;
}

// The reference's evaluation of sample s of global block gb (component comp).
__device__ __noinline__ int exact_sample_global(const int16_t *coef, const int16_t *dc, const int16_t *dcdiff,
                                                const ExactSmem *es, uint32_t parity, uint32_t gb, uint32_t comp, int s)
{
    const int16_t *blk = coef + (size_t)gb * 64u;
    uint4 ch[8];
    const bool drop_ac = parity && dcdiff[gb] == 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ch[k] = drop_ac ? make_uint4(0, 0, 0, 0) : __ldg(reinterpret_cast<const uint4 *>(blk) + k);
    const int dcv = dc[gb];
    ch[0].x = (ch[0].x & 0xFFFF0000u) | ((uint32_t)dcv & 0xFFFFu);
    const int x = s >> 3, y = s & 7;
    double cx[8], cy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        cx[k] = es->cosd[x][k];
        cy[k] = es->cosd[y][k];
    }
    const float sum = exact_terms(ch, es->qint[comp], cx, cy, es->cc[0][0], es->cc[0][1], std::make_integer_sequence<int, 64>{});
    const float out = (float)mul_f64(0.25, (double)sum);
    return round_half_away(out);
}

// pixel (unshifted integer samples) -> packed bytes, fast path with exact fallback
// eight strips per SM: 228 KB of shared memory, 1 KB of each CTA's share taken by the driver
static_assert(sizeof(IdctSmem<3>) <= (233472 / 8 - 1024), "idct_kernel<3> no longer fits eight CTAs per SM");

template <int NC>
__device__ __forceinline__ uint32_t colour_px(float y, float cb, float cr)
{
    if (NC == 1) {
        const int v = float_bits(y + (RINT_MAGIC + 128.0f)) - RINT_MAGIC_BITS;
        return (uint32_t)clamp_u8(v);
    }
    int R, G, B;
    if (!ycc_to_rgb_fast(y, cb, cr, R, G, B))
        return colour_exact_px(y, cb, cr);
    return (uint32_t)clamp_u8(R) | ((uint32_t)clamp_u8(G) << 8) | ((uint32_t)clamp_u8(B) << 16);
}

// Tie record: x = pixel index (image-major, row-major), y = global MCU index,
// z = sample index in the block | mask of flagged components << 8, w = fast Y | fast Cb << 16 (int16),
// and the fast Cr sample travels in the upper half of z.
__device__ __forceinline__ uint4 make_tie_record(uint32_t pix, uint32_t mcu, int s, uint32_t mask, float y, float cb, float cr)
{
    uint4 r;
    r.x = pix;
    r.y = mcu;
    r.z = (uint32_t)s | (mask << 8) | ((uint32_t)(uint16_t)(int)cr << 16);
    r.w = (uint32_t)(uint16_t)(int)y | ((uint32_t)(uint16_t)(int)cb << 16);
    return r;
}

template <int NC>
__device__ __forceinline__ void resolve_and_store_pixel(const ExactCtx &a, const ExactSmem *es, const uint4 &r)
{
    const uint32_t mcu = r.y;
    uint32_t mask = (r.z >> 8) & 7u;
    const int s = (int)(r.z & 63u);
    float v0 = (float)(short)(r.w & 0xFFFFu), v1 = (float)(short)(r.w >> 16), v2 = (float)(short)(r.z >> 16);
    // lanes flag different components: one call site, lane-varying component, so the warp makes one
    // pass per *number* of flagged components (almost always 1), not one per component
    while (mask) {
        const int c = __ffs(mask) - 1;
        mask &= mask - 1;
        const float e = (float)exact_sample_global(a.coef, a.dc, a.dcdiff, es, a.parity, mcu * NC + c, (uint32_t)c, s);
        v0 = c == 0 ? e : v0;
        v1 = c == 1 ? e : v1;
        v2 = c == 2 ? e : v2;
    }
    const uint32_t px = colour_px<NC>(v0, v1, v2);
    uint8_t *dst = a.pixels + (size_t)r.x * NC;
    dst[0] = (uint8_t)px;
    if (NC == 3) {
        dst[1] = (uint8_t)(px >> 8);
        dst[2] = (uint8_t)(px >> 16);
    }
}

// ---- stage 0 by the copy engine (TMA) ------------------------------------------------------------------
// One elected thread asks for the strip's coefficients as a 2-D tile of the coefficient matrix [blocks][64] through a
// tensor map with the 128-byte swizzle -- 16-byte chunk k of block b lands at chunk (k ^ (b & 7)), which is the
// bank-conflict-free layout stage 1 reads -- and for the two quantiser tables as plain bulk copies.  No thread spends
// an instruction on addresses, loads or stores; rows past the end of the matrix arrive as zeros.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "KPEG_MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra KPEG_MBAR_DONE;\n"
                 "bra KPEG_MBAR_WAIT;\n"
                 "KPEG_MBAR_DONE:\n"
                 "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_tile_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_tile_2d(const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Colour of the eight pixels of one row of an MCU -> 24 unclamped channel values, three variants chosen per MCU:
//   COLOUR_PLAIN    ycc_to_rgb_fast's arithmetic on pixel pairs (FFMA2 / FADD2); its range precondition holds for
//                   the whole MCU (no block is BLK_WIDE) and its flat-chroma special case cannot occur unnoticed
//                   (a pixel with Cb = Cr = 0 fails the G test and takes the exact expression, which is right too)
//   COLOUR_FLAT     both chroma blocks are all-zero (gray-as-YCbCr content): R = G = B = Y + 128
//   COLOUR_GENERAL  ycc_to_rgb_fast itself, pixel by pixel, with its range and flat tests
enum { COLOUR_PLAIN = 0, COLOUR_FLAT = 1, COLOUR_GENERAL = 2 };

// bits 0..23 = R, G, B; bit 24 set when the double expression was needed
__device__ __noinline__ uint32_t colour_px_general(float y, float cb, float cr)
{
    int R, G, B;
    if (!ycc_to_rgb_fast(y, cb, cr, R, G, B))
        return colour_exact_px(y, cb, cr) | (1u << 24);
    return (uint32_t)clamp_u8(R) | ((uint32_t)clamp_u8(G) << 8) | ((uint32_t)clamp_u8(B) << 16);
}

// -> the row's 24 bytes as six words; returns the number of pixels that took the double expression.
// redo (COLOUR_PLAIN only): bit j set = pixel j of the row must be replaced by colour_exact_px.
template <int MODE>
__device__ __forceinline__ uint32_t colour_row8(const float4 (&yy)[2], const float4 (&bb)[2], const float4 (&cc)[2],
                                                uint32_t (&out)[6], uint32_t &redo)
{
    uint32_t exact = 0;
    if constexpr (MODE == COLOUR_GENERAL) {
        const float Y[8] = {yy[0].x, yy[0].y, yy[0].z, yy[0].w, yy[1].x, yy[1].y, yy[1].z, yy[1].w};
        const float Cb[8] = {bb[0].x, bb[0].y, bb[0].z, bb[0].w, bb[1].x, bb[1].y, bb[1].z, bb[1].w};
        const float Cr[8] = {cc[0].x, cc[0].y, cc[0].z, cc[0].w, cc[1].x, cc[1].y, cc[1].z, cc[1].w};
        uint32_t p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p[j] = colour_px_general(Y[j], Cb[j], Cr[j]);
            exact += p[j] >> 24;
            p[j] &= 0xFFFFFFu;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            out[3 * q + 0] = p[4 * q] | (p[4 * q + 1] << 24);
            out[3 * q + 1] = (p[4 * q + 1] >> 8) | (p[4 * q + 2] << 16);
            out[3 * q + 2] = (p[4 * q + 2] >> 16) | (p[4 * q + 3] << 8);
        }
        return exact;
    }
    int px[24];
    if constexpr (MODE == COLOUR_FLAT) {
        const float Y[8] = {yy[0].x, yy[0].y, yy[0].z, yy[0].w, yy[1].x, yy[1].y, yy[1].z, yy[1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
            px[3 * j] = px[3 * j + 1] = px[3 * j + 2] = float_bits(Y[j] + (RINT_MAGIC + 128.0f)) - RINT_MAGIC_BITS;
    } else {
        const F2 magic = splat2(RINT_MAGIC);
        const F2 tiny = splat2(__int_as_float(1)); // 2^-149
        F2 dg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 &y4 = yy[k >> 1], &b4 = bb[k >> 1], &c4 = cc[k >> 1];
            const F2 y = (k & 1) ? pack2(y4.z, y4.w) : pack2(y4.x, y4.y);
            const F2 cb = (k & 1) ? pack2(b4.z, b4.w) : pack2(b4.x, b4.y);
            const F2 cr = (k & 1) ? pack2(c4.z, c4.w) : pack2(c4.x, c4.y);
            // ycc_to_rgb_fast (idct_core.h), two pixels per instruction
            const F2 yr = lane_add(y, splat2(127.501f));
            const F2 yg = lane_add(y, splat2(127.5f));
            const F2 r = lane_fma_k(cr, 1.402f, yr);
            const F2 b = lane_fma_k(cb, 1.772f, yr);
            const F2 g = lane_fma_k(cr, -0.714136f, lane_fma_k(cb, -0.344136f, yg));
            dg[k] = lane_sub(g, lane_sub(lane_add(g, magic), magic));
            // rint() to an integer WITHOUT the magic-number bias: v * 2^-149 is a denormal whose bit pattern is
            // rint(v) itself (round to nearest even, like the magic add) when v >= 0, and has the sign bit set -- a
            // large negative int -- when v < 0, which the saturating pack turns into 0 just as it would -|v|.
            const F2 ri = lane_mul(r, tiny), gi = lane_mul(g, tiny), bi = lane_mul(b, tiny);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * k + h;
                px[3 * j] = float_bits(h ? hi2(ri) : lo2(ri));
                px[3 * j + 1] = float_bits(h ? hi2(gi) : lo2(gi));
                px[3 * j + 2] = float_bits(h ? hi2(bi) : lo2(bi));
            }
        }
        // G within COLOUR_G_BAND of an integer somewhere in the row (0.2 % of the pixels): ONE test per row on the
        // largest |dg|; the caller replaces the listed pixels by the double expression after the row is stored
        float worst = fmaxf(fabsf(lo2(dg[0])), fabsf(hi2(dg[0])));
#pragma unroll
        for (int k = 1; k < 4; ++k)
            worst = fmaxf(worst, fmaxf(fabsf(lo2(dg[k])), fabsf(hi2(dg[k]))));
        if (!(worst < 0.5f - COLOUR_G_BAND)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (!(fabsf((j & 1) ? hi2(dg[j >> 1]) : lo2(dg[j >> 1])) < 0.5f - COLOUR_G_BAND))
                    redo |= 1u << j;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k)
        out[k] = pack4_sat(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
    return exact;
}

template <int NC>
__global__ void __launch_bounds__(NC * IDCT_MCUS_PER_CTA, NC == 3 ? KPEG_IDCT_MIN_CTAS : 16)
    idct_kernel(IdctArgs a, const __grid_constant__ CUtensorMap coef_map)
{
    constexpr int NM = IDCT_MCUS_PER_CTA;
    constexpr int NB = NM * NC;
    extern __shared__ __align__(1024) unsigned char smem_raw[]; // the swizzled tile needs 1024-byte alignment
    IdctSmem<NC> &sm = *reinterpret_cast<IdctSmem<NC> *>(smem_raw);

    const int t = threadIdx.x;
    const int comp = t / NM; // warp-uniform
    const int ml = t % NM;
    const int bl = ml * NC + comp; // block index inside the CTA's strip (MCU-interleaved)
    const uint32_t total_mcus = a.g.nimages * a.g.mcus_per_image;
    const uint32_t mcu0 = blockIdx.x * NM;
    const uint32_t m = mcu0 + ml;
    const uint32_t blk0 = mcu0 * NC;

    // ---- stage 0: quantisers + coefficients -> shared memory, by the copy engine -----------------------
    const uint32_t bar = smem_u32(&sm.mbar);
    if (t == 0) {
        constexpr uint32_t tile_bytes = NB * 128u, table_bytes = NC * 64u * 4u;
        mbar_init(bar, 1);
        mbar_expect_tx(bar, tile_bytes + 2u * table_bytes);
        tma_load_tile_2d(smem_u32(sm.coef), &coef_map, 0, (int)blk0, bar);
        bulk_load(smem_u32(sm.qpair), a.tables->qpair, table_bytes, bar);
        bulk_load(smem_u32(sm.qdc), a.tables->qdc, table_bytes, bar);
        // the strip that will run in this CTA's slot next (CTAs are dispatched in index order): pull it into L2 now
        if (blk0 + IDCT_PREFETCH_AHEAD * NB < a.g.total_blocks)
            tma_prefetch_tile_2d(&coef_map, 0, (int)(blk0 + IDCT_PREFETCH_AHEAD * NB));
        sm.nrec = 0;
        const uint32_t img = mcu0 / a.g.mcus_per_image, mi = mcu0 - img * a.g.mcus_per_image;
        sm.img0 = img;
        sm.by0 = mi / a.g.mcus_x;
        sm.bx0 = mi - sm.by0 * a.g.mcus_x;
    }
    // the block's DC value (from K2) and its DC difference: requested now, needed in stage 1
    int dcv = 0;
    bool drop_ac = false;
    if (m < total_mcus) {
        dcv = a.dc[blk0 + bl];
        drop_ac = (a.g.flags & 1u) && a.dcdiff[blk0 + bl] == 0; // MCU.cpp:97-104
    }
    // One thread waits for the copy engine (its wait acquires the tile), the CTA barrier hands the data on: the other
    // warps sleep at the barrier instead of polling, and the barrier object never needs to be visible to them.
    if (t == 0)
        mbar_wait(bar, 0);
    __syncthreads();

    // ---- stage 1: one thread = one 8x8 block ------------------------------------------------------
    uint4 ch[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ch[k] = sm.coef[bl * 8 + (k ^ (bl & 7))];
    __syncthreads(); // everyone holds its block in registers: `samp` may now overwrite `coef`
    if (m < total_mcus) {
        // the entropy stage leaves slot 0 empty; the integrated DC value comes from K2
        ch[0].x = (ch[0].x & 0xFFFF0000u) | ((uint32_t)dcv & 0xFFFFu);

        F2 P[32];
        const float A = dequant_dezigzag(ch, drop_ac ? sm.qdc[comp] : sm.qpair[comp], P, std::make_integer_sequence<int, 32>{});
        const float thresh = 0.5f - tie_band(A);
        idct_rows_packed(P);

        // (x + 1.5*2^23) - 1.5*2^23 == rint(x) for |x| < 2^22.  A sample is inside the tie band when its distance d
        // from the rounded value exceeds thresh, i.e. when d*d - thresh^2 >= 0: one FFMA2 per sample pair, and the
        // complement of the sign bit is shifted into the block's mask with one funnel shift per sample (no compares,
        // no predicates).  thresh^2 is taken a hair low, so the packed test can only flag more than |d| > thresh does.
        // Bit layout of the masks: tie_bit(row, col) below.
        const float t2 = thresh > 0.0f ? thresh * thresh * (1.0f - 1.0f / 2097152.0f) : 0.0f;
        const F2 neg_t2 = splat2(-t2);
        const F2 magic = splat2(RINT_MAGIC);
        uint32_t keep[2][2]; // [half][word]: sign bits = "outside the band", 16 per half and word
        auto half_block = [&](auto HALF) { // columns 4 half .. 4 half + 3 of all eight rows
            constexpr int half = decltype(HALF)::value;
            float v[4][8]; // [column - 4 half][row]
#pragma unroll
            for (int c = 0; c < 4; ++c)
                idct8_column(P[0 * 8 + 4 * half + c], P[1 * 8 + 4 * half + c], P[2 * 8 + 4 * half + c], P[3 * 8 + 4 * half + c], v[c]);
            uint32_t kl = 0, kh = 0;
#pragma unroll
            for (int row = 0; row < 8; ++row) {
                float r[4];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const F2 x = pack2(v[2 * k][row], v[2 * k + 1][row]);
                    const F2 rr = lane_sub(lane_add(x, magic), magic);
                    const F2 d = lane_sub(x, rr);
                    const F2 e = lane_fma(d, d, neg_t2);
                    unpack2(rr, r[2 * k], r[2 * k + 1]);
                    if (row < 4) {
                        kl = __funnelshift_l((uint32_t)float_bits(lo2(e)), kl, 1);
                        kl = __funnelshift_l((uint32_t)float_bits(hi2(e)), kl, 1);
                    } else {
                        kh = __funnelshift_l((uint32_t)float_bits(lo2(e)), kh, 1);
                        kh = __funnelshift_l((uint32_t)float_bits(hi2(e)), kh, 1);
                    }
                }
                sm.samp[((comp * 8 + row) * 2 + half) * NM + ml] = make_float4(r[0], r[1], r[2], r[3]);
            }
            keep[half][0] = kl;
            keep[half][1] = kh;
        };
        half_block(std::integral_constant<int, 0>{});
        half_block(std::integral_constant<int, 1>{});
        // 16 bits per (half, word), first sample shifted in ends up highest: bit 15 - (4 (row & 3) + (col & 3))
        sm.tie[bl] = make_uint2(~((keep[0][0] << 16) | (keep[1][0] & 0xFFFFu)), ~((keep[0][1] << 16) | (keep[1][1] & 0xFFFFu)));
        sm.flag[bl] = (uint8_t)((A != 0.0f ? BLK_NONZERO : 0u) | (A > COLOUR_SAFE_A ? BLK_WIDE : 0u));
    }
    __syncthreads();

    // ---- stage 2: colour conversion + interleaved store -------------------------------------------
    if (m < total_mcus) {
        uint32_t img = sm.img0, by = sm.by0, bx = sm.bx0 + (uint32_t)ml;
        if (a.g.mcus_x >= (uint32_t)NM) { // the strip wraps at most once
            if (bx >= a.g.mcus_x) {
                bx -= a.g.mcus_x;
                if (++by == a.g.mcus_y) {
                    by = 0;
                    ++img;
                }
            }
        } else { // images narrower than a strip
            img = m / a.g.mcus_per_image;
            const uint32_t mi = m - img * a.g.mcus_per_image;
            by = mi / a.g.mcus_x;
            bx = mi - by * a.g.mcus_x;
        }
        const uint32_t W = a.g.width, H = a.g.height;
        const uint32_t img_pix0 = img * W * H;
        uint8_t *img_base = a.pixels + (size_t)img_pix0 * NC;
        const bool full_w = bx * 8u + 8u <= W;
        const bool vec_ok = full_w && (W % 8u == 0u) && ((reinterpret_cast<uintptr_t>(a.pixels) & 7u) == 0);
        uint32_t colour_exact = 0;
        bool overflow = false;
        uint2 tmask[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c)
            tmask[c] = sm.tie[ml * NC + c];
        int mode = COLOUR_PLAIN;
        if constexpr (NC == 3) {
            const uint32_t f0 = sm.flag[ml * NC], f1 = sm.flag[ml * NC + 1], f2 = sm.flag[ml * NC + 2];
            mode = ((f0 | f1 | f2) & BLK_WIDE) ? COLOUR_GENERAL : (((f1 | f2) & BLK_NONZERO) ? COLOUR_PLAIN : COLOUR_FLAT);
        }
        for (int row = comp; row < 8; row += NC) { // the NC warps of the CTA share the 8 pixel rows
            const uint32_t y = by * 8u + row;
            if (y >= H)
                continue;
            const float4 y0 = sm.samp[((0 * 8 + row) * 2 + 0) * NM + ml];
            const float4 y1 = sm.samp[((0 * 8 + row) * 2 + 1) * NM + ml];
            uint32_t out[2 * NC];
            uint32_t redo = 0;
            if constexpr (NC == 3) {
                const float4 yy[2] = {y0, y1};
                const float4 bb[2] = {sm.samp[((1 * 8 + row) * 2 + 0) * NM + ml], sm.samp[((1 * 8 + row) * 2 + 1) * NM + ml]};
                const float4 cc[2] = {sm.samp[((2 * 8 + row) * 2 + 0) * NM + ml], sm.samp[((2 * 8 + row) * 2 + 1) * NM + ml]};
                uint32_t o6[6];
                if (mode == COLOUR_PLAIN)
                    colour_exact += colour_row8<COLOUR_PLAIN>(yy, bb, cc, o6, redo);
                else if (mode == COLOUR_FLAT)
                    colour_exact += colour_row8<COLOUR_FLAT>(yy, bb, cc, o6, redo);
                else
                    colour_exact += colour_row8<COLOUR_GENERAL>(yy, bb, cc, o6, redo);
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    out[k] = o6[k];
            } else {
                // gray: the reference's colour path with Cb = Cr = 128 gives R = G = B = clamp(Y)
                const float Y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                int v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[j] = float_bits(Y[j] + (RINT_MAGIC + 128.0f)) - RINT_MAGIC_BITS; // integer-valued: exact
                out[0] = pack4_sat(v[0], v[1], v[2], v[3]);
                out[1] = pack4_sat(v[4], v[5], v[6], v[7]);
            }
            uint8_t *dst = img_base + ((size_t)y * W + bx * 8u) * NC;
            if (vec_ok) {
#pragma unroll
                for (int k = 0; k < NC; ++k)
                    reinterpret_cast<uint2 *>(dst)[k] = make_uint2(out[2 * k], out[2 * k + 1]);
            } else {
                const uint32_t nbytes = (full_w ? 8u : W - bx * 8u) * NC;
#pragma unroll
                for (int j = 0; j < 8 * NC; ++j) // compile-time indices: `out` stays in registers
                    if ((uint32_t)j < nbytes)
                        dst[j] = (uint8_t)(out[j >> 2] >> (8 * (j & 3)));
            }
            while (redo) { // rare: this thread's own later byte stores replace what it has just written
                const int j = __ffs(redo) - 1;
                redo &= redo - 1;
                if (bx * 8u + (uint32_t)j >= W)
                    continue;
                const float *sp = reinterpret_cast<const float *>(sm.samp);
                const uint32_t e = colour_exact_px(sp[(((0 * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)],
                                                   sp[(((1 * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)],
                                                   sp[((((NC - 1) * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)]);
                dst[3 * j] = (uint8_t)e;
                dst[3 * j + 1] = (uint8_t)(e >> 8);
                dst[3 * j + 2] = (uint8_t)(e >> 16);
                ++colour_exact;
            }
        }
        // pixels with a sample inside the tie band, in the rows this thread converted: queue them for the exact
        // pass.  One pass over the set bits of the MCU's combined mask after the row loop (most MCUs have none).
        uint32_t plo = 0, phi = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            plo |= tmask[c].x;
            phi |= tmask[c].y;
        }
        if (NC == 3) { // rows comp, comp + 3, comp + 6
            plo &= comp == 0 ? (tie_row_bits(0) | tie_row_bits(3)) : (comp == 1 ? tie_row_bits(1) : tie_row_bits(2));
            phi &= comp == 0 ? tie_row_bits(6) : (comp == 1 ? (tie_row_bits(4) | tie_row_bits(7)) : tie_row_bits(5));
        }
        while (plo | phi) {
            int w, b;
            if (plo) {
                w = 0;
                b = __ffs(plo) - 1;
                plo &= plo - 1;
            } else {
                w = 1;
                b = __ffs(phi) - 1;
                phi &= phi - 1;
            }
            const int s = tie_sample(w, b);
            const int row = s >> 3, j = s & 7;
            if (by * 8u + (uint32_t)row >= H || bx * 8u + (uint32_t)j >= W)
                continue;
            uint32_t cm = 0;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                cm |= (((w ? tmask[c].y : tmask[c].x) >> b) & 1u) << c;
            // fast samples of this pixel, re-read from shared memory
            const float *sp = reinterpret_cast<const float *>(sm.samp);
            float fy, fcb = 0.0f, fcr = 0.0f;
            fy = sp[(((0 * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)];
            if (NC == 3) {
                fcb = sp[(((1 * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)];
                fcr = sp[((((NC - 1) * 8 + row) * 2 + (j >> 2)) * NM + ml) * 4 + (j & 3)];
            }
            // compact strip-local form: x = mcu in strip | sample << 5 | components << 11 | fast Cr << 16,
            // y = fast Y | fast Cb << 16; expanded to the global record when the strip's list is flushed
            const uint2 rec = make_uint2((uint32_t)ml | ((uint32_t)s << 5) | (cm << 11) | ((uint32_t)(uint16_t)(int)fcr << 16),
                                         (uint32_t)(uint16_t)(int)fy | ((uint32_t)(uint16_t)(int)fcb << 16));
            const uint32_t at = atomicAdd(&sm.nrec, 1u); // shared-memory counter: no global atomic unless the strip outgrows its own slots
            if (at < (uint32_t)IDCT_REC_CAP)
                sm.rec[at] = rec;
            else
                overflow = true; // more tied pixels than a strip's list holds: the strip is redone wholesale
        }
        if (colour_exact)
            atomicAdd(&a.meta->colour_exact, colour_exact);
        if (overflow)
            a.overflow_mcu[blockIdx.x] = 1u; // this strip has pixels that did not fit the record list
    }
    // ---- flush the strip's tie records --------------------------------------------------------------
    // Strip i owns slots [i * IDCT_REC_FIXED, (i + 1) * IDCT_REC_FIXED) of the global list and always writes all of
    // them (unused ones as TIE_EMPTY): the common case needs no reservation, so no CTA ends on the round trip of a
    // global atomic.  Records beyond the strip's own slots (rare) go to the shared tail of the list.
    __syncthreads();
    const uint32_t found = sm.nrec, nrec = min(found, (uint32_t)IDCT_REC_CAP);
    auto global_record = [&](uint32_t i) {
        const uint2 c = sm.rec[i];
        const uint32_t rm = mcu0 + (c.x & 31u), rs = (c.x >> 5) & 63u;
        // position of the record's MCU from the strip's origin (as in stage 2: no division when the strip wraps at most once)
        uint32_t img = sm.img0, by = sm.by0, bx = sm.bx0 + (c.x & 31u);
        if (a.g.mcus_x >= (uint32_t)NM) {
            if (bx >= a.g.mcus_x) {
                bx -= a.g.mcus_x;
                if (++by == a.g.mcus_y) {
                    by = 0;
                    ++img;
                }
            }
        } else {
            img = rm / a.g.mcus_per_image;
            const uint32_t mi = rm - img * a.g.mcus_per_image;
            by = mi / a.g.mcus_x;
            bx = mi - by * a.g.mcus_x;
        }
        uint4 r; // the record format of make_tie_record
        r.x = img * a.g.width * a.g.height + (by * 8u + (rs >> 3)) * a.g.width + bx * 8u + (rs & 7u);
        r.y = rm;
        r.z = rs | (((c.x >> 11) & 7u) << 8) | (c.x & 0xFFFF0000u);
        r.w = c.y;
        return r;
    };
    if (t < IDCT_REC_FIXED)
        a.tie_rec[(size_t)blockIdx.x * IDCT_REC_FIXED + t] = (uint32_t)t < nrec ? global_record(t) : make_uint4(TIE_EMPTY, 0, 0, 0);
    if (found <= (uint32_t)IDCT_REC_FIXED) {
        if (t == 0 && found)
            atomicAdd(&a.meta->exact_samples, found); // statistics only: nobody waits for it
        return;
    }
    if (t == 0) {
        sm.rec_base = atomicAdd(&a.meta->tie_records, nrec - (uint32_t)IDCT_REC_FIXED);
        atomicAdd(&a.meta->exact_samples, (uint32_t)IDCT_REC_FIXED);
        if (found > (uint32_t)IDCT_REC_CAP)
            atomicAdd(&a.meta->tie_inline, found - (uint32_t)IDCT_REC_CAP);
    }
    __syncthreads();
    const uint32_t base = gridDim.x * (uint32_t)IDCT_REC_FIXED + sm.rec_base; // the tail starts after every strip's own slots
    for (uint32_t i = IDCT_REC_FIXED + t; i < nrec; i += NB) {
        const uint32_t at = base + (i - IDCT_REC_FIXED);
        if (at < a.tie_cap) {
            a.tie_rec[at] = global_record(i);
        } else {
            a.overflow_mcu[blockIdx.x] = 1u; // global list full
            atomicAdd(&a.meta->tie_inline, 1u);
        }
    }
}

// The record list: strips * IDCT_REC_FIXED slots owned by the strips (sparse: unused slots are TIE_EMPTY) followed by
// a compact tail of meta->tie_records entries.  A warp walks its share of the sparse part 32 slots at a time,
// collects the used ones in a small shared-memory stack and resolves them 32 at a time, so the long serial
// evaluation always runs with full warps.
template <int NC>
__global__ void __launch_bounds__(128) idct_patch_kernel(IdctArgs a)
{
    __shared__ ExactSmem es;
    __shared__ uint4 s_q[4][64];
    exact_smem_load(es, a.tables);
    const ExactCtx x = exact_ctx(a);
    const uint32_t total_mcus_ = a.g.nimages * a.g.mcus_per_image;
    const uint32_t nslots = ((total_mcus_ + IDCT_MCUS_PER_CTA - 1) / IDCT_MCUS_PER_CTA) * (uint32_t)IDCT_REC_FIXED;
    {
        const uint32_t lane = threadIdx.x & 31u, wl = threadIdx.x >> 5;
        const uint32_t nwarps = gridDim.x * 4u, w = blockIdx.x * 4u + wl;
        const uint32_t per = (((nslots + nwarps - 1u) / nwarps) + 31u) & ~31u;
        const uint32_t s0 = w * per, s1 = min(s0 + per, nslots);
        uint32_t qn = 0; // warp-uniform
        for (uint32_t sl = s0; sl < s1; sl += 32u) {
            uint4 r = make_uint4(TIE_EMPTY, 0, 0, 0);
            if (sl + lane < s1)
                r = a.tie_rec[sl + lane];
            const bool used = r.x != TIE_EMPTY;
            const uint32_t bal = __ballot_sync(0xffffffffu, used);
            if (used)
                s_q[wl][qn + __popc(bal & ((1u << lane) - 1u))] = r;
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 32u) {
                r = s_q[wl][qn - 32u + lane];
                __syncwarp();
                qn -= 32u;
                resolve_and_store_pixel<NC>(x, &es, r);
                __syncwarp();
            }
        }
        if (lane < qn)
            resolve_and_store_pixel<NC>(x, &es, s_q[wl][lane]);
    }
    const uint32_t n = a.tie_cap > nslots ? min(a.meta->tie_records, a.tie_cap - nslots) : 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        resolve_and_store_pixel<NC>(x, &es, a.tie_rec[nslots + i]);
    // Overflow (pathologically flat images: more tied pixels than the record list holds).  Strips that
    // reported an overflow are redone wholesale on the exact path: every sample of every pixel.
    if (a.meta->tie_inline == 0u)
        return;
    const uint32_t total_mcus = a.g.nimages * a.g.mcus_per_image;
    const uint32_t nstrips = (total_mcus + IDCT_MCUS_PER_CTA - 1) / IDCT_MCUS_PER_CTA;
    const uint32_t W = a.g.width, H = a.g.height;
    for (uint32_t strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
        if (a.overflow_mcu[strip] == 0u)
            continue;
        for (uint32_t w = threadIdx.x; w < IDCT_MCUS_PER_CTA * 64u; w += blockDim.x) {
            const uint32_t m = strip * IDCT_MCUS_PER_CTA + (w >> 6);
            const int s = (int)(w & 63u);
            if (m >= total_mcus)
                continue;
            const uint32_t img = m / a.g.mcus_per_image, mi = m - img * a.g.mcus_per_image;
            const uint32_t by = mi / a.g.mcus_x, bx = mi - by * a.g.mcus_x;
            const uint32_t px_x = bx * 8u + (uint32_t)(s & 7), px_y = by * 8u + (uint32_t)(s >> 3);
            if (px_x >= W || px_y >= H)
                continue;
            uint4 rec;
            rec.x = img * W * H + px_y * W + px_x;
            rec.y = m;
            rec.z = (uint32_t)s | ((NC == 3 ? 7u : 1u) << 8);
            rec.w = 0;
            resolve_and_store_pixel<NC>(x, &es, rec);
        }
    }
}

static uint32_t g_patch_grid = 148 * 4;

void kernels_configure(int max_concurrent_jobs)
{
    g_max_concurrent_loops = (uint32_t)(max_concurrent_jobs > 0 ? max_concurrent_jobs : 1);
    const int k1max = (int)k1_smem_bytes(1024);
    cudaFuncSetAttribute(entropy_cold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k1max);
    cudaFuncSetAttribute(entropy_relay_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k1max);
    cudaFuncSetAttribute(entropy_relay_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)k1_sparse_smem_bytes(1024));
    {
        int dev = 0, sms = 148, per_sm = 8;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(entropy_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)k1_write_smem_bytes(1024));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_cold_kernel, ENTROPY_THREADS,
                                                          k1_smem_bytes(512)) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        g_k1_grid_cap = (uint32_t)(sms * per_sm);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_write_kernel, WRITE_THREADS,
                                                          k1_write_smem_bytes(512)) != cudaSuccess || per_sm < 1)
            per_sm = 3;
        g_k1_write_grid_cap = (uint32_t)(sms * per_sm);
        cudaFuncSetAttribute(entropy_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WriteSmemTail));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_expand_kernel, WRITE_THREADS,
                                                          sizeof(WriteSmemTail)) != cudaSuccess || per_sm < 1)
            per_sm = 6;
        g_k1_expand_grid_cap = (uint32_t)(sms * per_sm);
        g_patch_grid = (uint32_t)(sms * 4); // warps take contiguous shares of the record list: enough records each to fill whole batches
        if (const char *e = getenv("KPEG_PATCH_PER_SM")) // experiments
            g_patch_grid = (uint32_t)(sms * std::max(1, atoi(e)));
        cudaFuncSetAttribute(entropy_relay_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)k1_sparse_smem_bytes(1024));
        g_sm_count = (uint32_t)sms;
        if (const char *e = getenv("KPEG_RELAY_LOOP_PER_SM"))
            g_relay_loop_per_sm_cap = atoi(e);
        if (const char *e = getenv("KPEG_RELAY_SPIN_LIMIT")) // tests: 0 = give up at the first poll that finds a CTA missing
            g_relay_spin_limit = (uint32_t)strtoul(e, nullptr, 10);
    }
    cudaFuncSetAttribute(idct_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IdctSmem<3>));
    cudaFuncSetAttribute(idct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IdctSmem<1>));
}

// Tensor map of the coefficient matrix [total_blocks][64] of int16, tiles of `box_blocks` whole blocks, 128-byte
// swizzle, zeros past the end.  Encoding is a host-side computation (no driver call reaches the device).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

static bool make_coef_map(CUtensorMap *map, const int16_t *coef, uint32_t total_blocks, uint32_t box_blocks)
{
    const EncodeTiledFn enc = encode_tiled_fn();
    if (!enc)
        return false;
    const cuuint64_t dims[2] = {64, total_blocks};
    const cuuint64_t strides[1] = {128}; // bytes between blocks
    const cuuint32_t box[2] = {64, box_blocks};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t *>(coef), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_idct(const IdctArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t total_mcus = a.g.nimages * a.g.mcus_per_image;
    const uint32_t grid = (total_mcus + IDCT_MCUS_PER_CTA - 1) / IDCT_MCUS_PER_CTA;
    alignas(64) CUtensorMap map;
    if (!make_coef_map(&map, a.coef, a.g.total_blocks, a.g.ncomp * IDCT_MCUS_PER_CTA))
        return cudaErrorNotSupported; // no cuTensorMapEncodeTiled in this driver, or it rejected the geometry
    if (a.g.ncomp == 3)
        idct_kernel<3><<<grid, 3 * IDCT_MCUS_PER_CTA, sizeof(IdctSmem<3>), s>>>(a, map);
    else
        idct_kernel<1><<<grid, IDCT_MCUS_PER_CTA, sizeof(IdctSmem<1>), s>>>(a, map);
    ++*launches;
    return cudaSuccess;
}

void launch_idct_patch(const IdctArgs &a, cudaStream_t s, uint32_t *launches)
{
    if (a.g.ncomp == 3)
        idct_patch_kernel<3><<<g_patch_grid, 128, 0, s>>>(a);
    else
        idct_patch_kernel<1><<<g_patch_grid, 128, 0, s>>>(a);
    ++*launches;
}

// =================================================================================================
// parity hook: coefficients with the DC value merged in and the F1 rule applied
// =================================================================================================
__global__ void merge_dc_kernel(int16_t *out, const int16_t *coef, const int16_t *dc, const int16_t *dcdiff,
                                uint32_t nblocks, uint32_t flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks * 64u)
        return;
    const uint32_t b = i >> 6, z = i & 63u;
    int16_t v;
    if (z == 0u)
        v = dc[b];
    else
        v = ((flags & 1u) && dcdiff[b] == 0) ? (int16_t)0 : coef[i];
    out[i] = v;
}

void launch_merge_dc(int16_t *coef_out, const int16_t *coef, const int16_t *dc, const int16_t *dcdiff, uint32_t nblocks,
                     uint32_t flags, cudaStream_t s)
{
    const uint32_t n = nblocks * 64u;
    merge_dc_kernel<<<(n + 255) / 256, 256, 0, s>>>(coef_out, coef, dc, dcdiff, nblocks, flags);
}

} // namespace kpeg
