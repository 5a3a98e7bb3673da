// kernels.cu -- hand-written sm_100a kernels of the baseline-JPEG decode hot path.
//
//   K0  unstuff_*      FF00 / RSTn removal + restart-segment table     (Decoder.cpp:532-577, 621-653)
//   K1  entropy_*      Huffman + run-length decode, parallel over fixed-size subsequences with a
//                      self-synchronising relay and a segmented offset scan (Decoder.cpp:655-855)
//   K2 + K3 live in k3_fused.cu: record expansion + DC prediction + dequantise + de-zigzag + 8x8 IDCT + level
//                      shift + YCbCr->RGB + interleaved store as ONE kernel (MCU.cpp:93-279, Image.cpp:51-70)
//
// All file:line citations are relative to /root/reference.  The arithmetic lives in
// entropy_core.h / idct_core.h (host+device inline, also exercised on the CPU by tests/emu).
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <stddef.h>
#include <stdint.h>

#include <utility>

#include "entropy_core.h"
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

// Host-buffer batches: the scans of a chunk are copied into the lane's stream buffer one after another, leaving two
// bytes after each; this kernel writes the RSTn separators there (ends[i] = offset just past scan i), so the host
// issues no 2-byte copies.
__global__ void __launch_bounds__(256) write_separators_kernel(uint8_t *scan, const uint64_t *ends, uint32_t n)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) {
        const uint64_t e = ends[i];
        scan[e] = 0xFFu;
        scan[e + 1u] = (uint8_t)(0xD0u + (i & 7u));
    }
}

// Host-buffer batches whose scans lie back to back in host memory travel as ONE copy into a staging buffer; this kernel
// moves scan i from there to its place in the lane's stream buffer (two bytes further on per scan: the separators) and
// writes the RSTn after it.  ends[i] = offset just past scan i in the DESTINATION (as for write_separators_kernel), so
// scan i starts at ends[i-1] + 2 there and at ends[i-1] + 2 - 2 i in the staging buffer.  One CTA per scan and
// round; destination words are written whole (two source words and a funnel shift), the ragged ends byte by byte.
__global__ void __launch_bounds__(256) repack_scans_kernel(const uint8_t *stage, uint8_t *scan, const uint64_t *ends, uint32_t n)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t d0 = i ? ends[i - 1] + 2u : 0u, d1 = ends[i]; // destination byte range of the scan
        const uint8_t *src = stage + (d0 - 2ull * i);
        uint8_t *dst = scan + d0;
        const uint64_t len = d1 - d0;
        const uint64_t to_word = (4u - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
        const uint64_t head = len < to_word ? len : to_word;
        const uint64_t nwords = (len - head) >> 2;
        const uint8_t *sw = src + head; // source of the first whole destination word
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(sw) & 3u);
        const uint32_t *sa = reinterpret_cast<const uint32_t *>(sw - mis);
        uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
        for (uint64_t w = threadIdx.x; w < nwords; w += 256u) {
            const uint32_t lo = __ldg(sa + w);
            dw[w] = mis ? __funnelshift_r(lo, __ldg(sa + w + 1u), 8u * mis) : lo;
        }
        for (uint64_t b = threadIdx.x; b < head; b += 256u)
            dst[b] = src[b];
        for (uint64_t b = head + (nwords << 2) + threadIdx.x; b < len; b += 256u)
            dst[b] = src[b];
        if (threadIdx.x == 0) {
            scan[d1] = 0xFFu;
            scan[d1 + 1u] = (uint8_t)(0xD0u + (i & 7u));
        }
    }
}

void launch_repack_scans(const uint8_t *stage, uint8_t *scan, const uint64_t *ends, uint32_t n, cudaStream_t s)
{
    repack_scans_kernel<<<n < 148u * 8u ? n : 148u * 8u, 256, 0, s>>>(stage, scan, ends, n);
}

// Number of RSTn markers (FF D0 .. FF D7) in scan[0, len): what tells how many restart intervals -- hence MCU rows -- a
// band that was cut at a byte position holds (kpeg_cuda.cu decode_banded).  Inside entropy-coded data an FF is always
// followed by 00 (stuffing), another FF (fill) or a marker, so the two-byte test is exact.  16 bytes per thread, the byte
// after the chunk fetched separately; one atomic per warp.
__global__ void __launch_bounds__(256) count_restart_markers_kernel(const uint8_t *scan, uint32_t len, uint32_t *count)
{
    const uint32_t stride = gridDim.x * 256u * 16u;
    uint32_t n = 0;
    for (uint32_t at = (blockIdx.x * 256u + threadIdx.x) * 16u; at < len; at += stride) {
        const uint32_t m = min(16u, len - at);
        uint32_t prev = __ldg(scan + at);
        for (uint32_t i = 1; i <= m; ++i) {
            const uint32_t cur = at + i < len ? (uint32_t)__ldg(scan + at + i) : 0u;
            n += (prev == 0xFFu && (cur & 0xF8u) == 0xD0u) ? 1u : 0u;
            prev = cur;
        }
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31u) == 0u && n)
        atomicAdd(count, n);
}

void launch_count_restart_markers(const uint8_t *scan, uint32_t len, uint32_t *count, cudaStream_t s)
{
    const uint32_t blocks = (len + 4095u) / 4096u;
    count_restart_markers_kernel<<<blocks < 148u * 8u ? (blocks ? blocks : 1u) : 148u * 8u, 256, 0, s>>>(scan, len, count);
}

void launch_write_separators(uint8_t *scan, const uint64_t *ends, uint32_t n, cudaStream_t s)
{
    write_separators_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(scan, ends, n);
}

// ---- output format helpers (GPU-side PPM / planar writers) ------------------------------------------------------
// gray -> R = G = B triplets (what the reference's colour path yields for Cb = Cr = 128): four pixels
// per thread, one 4-byte load and three 4-byte stores.
__global__ void __launch_bounds__(256) gray_to_rgb_kernel(const uint8_t *gray, uint8_t *rgb, size_t n)
{
    const size_t q = (size_t)blockIdx.x * 256u + threadIdx.x; // quad of pixels
    const size_t i = q * 4u;
    if (i + 4u <= n && ((reinterpret_cast<uintptr_t>(gray) | reinterpret_cast<uintptr_t>(rgb)) & 3u) == 0) {
        const uint32_t g = __ldg(reinterpret_cast<const uint32_t *>(gray) + q);
        const uint32_t a = g & 0xFFu, b = (g >> 8) & 0xFFu, c = (g >> 16) & 0xFFu, d = g >> 24;
        uint32_t *o = reinterpret_cast<uint32_t *>(rgb) + q * 3u;
        o[0] = a * 0x010101u | (b << 24);
        o[1] = b * 0x0101u | (c * 0x0101u << 16);
        o[2] = c | (d * 0x010101u << 8);
    } else {
        for (size_t k = i; k < n && k < i + 4u; ++k)
            rgb[3u * k] = rgb[3u * k + 1u] = rgb[3u * k + 2u] = gray[k];
    }
}

// interleaved R,G,B -> three planes: four pixels per thread (three 4-byte loads, three 4-byte stores)
__global__ void __launch_bounds__(256) rgb_to_planar_kernel(const uint8_t *rgb, uint8_t *planes, size_t n)
{
    const size_t q = (size_t)blockIdx.x * 256u + threadIdx.x;
    const size_t i = q * 4u;
    uint8_t *pr = planes, *pg = planes + n, *pb = planes + 2u * n;
    if (i + 4u <= n && (n & 3u) == 0 && ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(planes)) & 3u) == 0) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(rgb) + q * 3u;
        const uint32_t w0 = __ldg(in), w1 = __ldg(in + 1), w2 = __ldg(in + 2); // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        reinterpret_cast<uint32_t *>(pr)[q] = (__byte_perm(w0, w1, 0x0630) & 0x00FFFFFFu) | ((w2 >> 8) & 0xFFu) << 24;
        reinterpret_cast<uint32_t *>(pg)[q] = ((w0 >> 8) & 0xFFu) | (w1 & 0xFFu) << 8 | (w1 >> 24) << 16 | ((w2 >> 16) & 0xFFu) << 24;
        reinterpret_cast<uint32_t *>(pb)[q] = ((w0 >> 16) & 0xFFu) | ((w1 >> 8) & 0xFFu) << 8 | (w2 & 0xFFu) << 16 | (w2 >> 24) << 24;
    } else {
        for (size_t k = i; k < n && k < i + 4u; ++k) {
            pr[k] = rgb[3u * k];
            pg[k] = rgb[3u * k + 1u];
            pb[k] = rgb[3u * k + 2u];
        }
    }
}

void launch_gray_to_rgb(const uint8_t *gray, uint8_t *rgb, size_t n, cudaStream_t s)
{
    gray_to_rgb_kernel<<<(unsigned)((n + 1023u) / 1024u), 256, 0, s>>>(gray, rgb, n);
}

void launch_rgb_to_planar(const uint8_t *rgb, uint8_t *planes, size_t n, cudaStream_t s)
{
    rgb_to_planar_kernel<<<(unsigned)((n + 1023u) / 1024u), 256, 0, s>>>(rgb, planes, n);
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

#ifndef KPEG_LONGLUT_GLOBAL
#define KPEG_LONGLUT_GLOBAL 0 // 1: second-level Huffman tables read from global memory instead of shared memory
#endif

// Shared-memory image of what one CTA needs: the lookup tables (staged once per CTA) and the slice of
// the bit stream that belongs to the tile of ENTROPY_THREADS subsequences being decoded.  The slice is
// stored linearly with one padding word after every 32 (position l + (l >> 5)): lanes reading word k
// of their own subsequence (stride 8, 16 or 32 words) then hit 32 different banks.  The slice
// carries four extra words: a symbol may run up to 26 bits past the end of the last subsequence and
// the bit window looks two words ahead.
struct K1Smem {
    uint16_t fast[MAX_LUTS * LUT_SIZE];
#if !KPEG_LONGLUT_GLOBAL
    uint16_t longlut[MAX_LUTS * LONG_CAP];
#endif
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
    const uint16_t *glong; // the second-level tables in global memory (KPEG_LONGLUT_GLOBAL)
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
#if KPEG_LONGLUT_GLOBAL
            v = __ldg(glong + idx); // second level from global memory (L1-resident: 6 KB): the shared memory it took is worth two more CTAs per SM
#else
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(long_addr + 2u * idx));
#endif
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
#if !KPEG_LONGLUT_GLOBAL
    for (uint32_t ti = 0; ti < nt; ++ti) {
        const uint32_t n = (t->luts.long_n[ti] * (uint32_t)sizeof(uint16_t) + 15u) / 16u;
        const uint4 *src = reinterpret_cast<const uint4 *>(&t->luts.longlut[ti][0]);
        uint4 *dst = reinterpret_cast<uint4 *>(&sm.longlut[ti * LONG_CAP]);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            dst[i] = __ldg(src + i);
    }
#endif
}

// The tile's stream slice (coalesced loads).  Ends with a __syncthreads().
template <int TILE = ENTROPY_THREADS>
__device__ __forceinline__ void k1_stage_stream(K1Smem &sm, const EntropyArgs &a, uint32_t tile, uint32_t total_bits,
                                                uint32_t wlog)
{
    const uint32_t gw0 = (tile * TILE) << wlog;
    const uint32_t nwords = ((uint32_t)TILE << wlog) + K1_TAIL_WORDS;
    const uint32_t total_words = ((total_bits + 31u) >> 5) + 4u; // the stream buffer has zeroed slack beyond this
    // eight loads in flight per thread: issued one by one (load, wait, store) the staging was a chain of sixteen memory
    // round trips per tile and a quarter of the stall samples of the cold and relay kernels
    constexpr int STAGE_BATCH = 8;
    for (uint32_t l0 = threadIdx.x; l0 < nwords; l0 += STAGE_BATCH * blockDim.x) {
        uint32_t v[STAGE_BATCH];
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const uint32_t l = l0 + (uint32_t)i * blockDim.x, gw = gw0 + l;
            v[i] = (l < nwords && gw < total_words) ? __ldg(a.words + gw) : 0u;
        }
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const uint32_t l = l0 + (uint32_t)i * blockDim.x;
            if (l < nwords)
                sm.words[k1_pad(l)] = v[i];
        }
    }
    __syncthreads();
}

// ---- per-thread stream regions (cold pass, relay round 1) -----------------------------------------------------------
// The relay passes load the two words that cover a symbol in every iteration (entropy_core.h relay_run), so the address
// of word j must cost one multiply-add: every thread gets a REGION of its own -- the words of its subsequence followed
// by the first word of the next one (a symbol that begins in the last word runs into it) -- at stride
// words_per_subsequence + 1, which is odd: lanes reading word k of their regions hit 32 different banks.  In terms of
// the linear slice, word l sits at l + l / words_per_subsequence, and the slot before a region's first word holds a
// second copy of that word (it is the look-ahead word of the region before).
struct RegionWords {
    uint32_t pbase; // shared byte address of the region's first word - 4 * (global index of that word)
    __device__ __forceinline__ uint32_t operator()(uint32_t gw) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(pbase + 4u * gw));
        return v;
    }
    __device__ __forceinline__ void pair(uint32_t j, uint32_t &w0, uint32_t &w1) const
    {
        const uint32_t at = pbase + 4u * j;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(at));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(at));
    }
};

__host__ __device__ inline size_t k1_region_smem_bytes(uint32_t sub_bits)
{
    return sizeof(K1Smem) + (size_t)(ENTROPY_THREADS * (sub_bits / 32u + 1u) + 1u) * sizeof(uint32_t);
}

// The tile's stream slice into the threads' regions (coalesced loads).  Ends with a __syncthreads().
__device__ __forceinline__ void k1_stage_regions(K1Smem &sm, const EntropyArgs &a, uint32_t tile, uint32_t total_bits, uint32_t wlog)
{
    const uint32_t gw0 = (tile * ENTROPY_THREADS) << wlog;
    const uint32_t nwords = ((uint32_t)ENTROPY_THREADS << wlog) + 1u;
    const uint32_t total_words = ((total_bits + 31u) >> 5) + 4u; // the stream buffer has zeroed slack beyond this
    const uint32_t in_sub = (1u << wlog) - 1u;
    constexpr int STAGE_BATCH = 8; // loads in flight per thread
    for (uint32_t l0 = threadIdx.x; l0 < nwords; l0 += STAGE_BATCH * blockDim.x) {
        uint32_t v[STAGE_BATCH];
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const uint32_t l = l0 + (uint32_t)i * blockDim.x, gw = gw0 + l;
            v[i] = (l < nwords && gw < total_words) ? __ldg(a.words + gw) : 0u;
        }
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i) {
            const uint32_t l = l0 + (uint32_t)i * blockDim.x;
            if (l < nwords) {
                const uint32_t at = l + (l >> wlog);
                sm.words[at] = v[i];
                if ((l & in_sub) == 0u && l != 0u)
                    sm.words[at - 1u] = v[i]; // look-ahead word of the region before
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ RegionWords k1_region_words(const K1Smem &sm, uint32_t tile, uint32_t wlog)
{
    // thread t: region at word t * (words per subsequence + 1), first global word (tile * THREADS + t) * words per subsequence
    const uint32_t t = threadIdx.x, wps = 1u << wlog;
    RegionWords W;
    W.pbase = (uint32_t)__cvta_generic_to_shared(sm.words) + 4u * (t * (wps + 1u)) - 4u * (((tile * ENTROPY_THREADS + t) << wlog));
    return W;
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
#if KPEG_LONGLUT_GLOBAL
    L.long_addr = 0;
#else
    L.long_addr = (uint32_t)__cvta_generic_to_shared(sm.longlut);
#endif
    L.glong = &a.tables->luts.longlut[0][0];
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
        k1_stage_regions(sm, a, tile, total_bits, wlog);
        const uint32_t sub = tile * ENTROPY_THREADS + threadIdx.x;
        if (sub < nsub) {
            const RegionWords W = k1_region_words(sm, tile, wlog);
            const uint32_t p0 = sub << (wlog + 5u);
            const uint32_t end = min(p0 + a.g.sub_bits, total_bits);
            const uint32_t hint = a.g.nseg > 1u ? first_seg_at_or_after(a.seg_bit, a.g.nseg, p0) : (sub ? 1u : 0u);
            a.seg_hint[sub] = hint;
            a.state[sub] = relay_span(W, L, S, a.g, end, p0, 0u, 0u, hint);
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
    char *base; // &rec[rec_base_index(sub)], or the subsequence's private area
    uint32_t kmax;
    uint32_t stride_bytes; // 128 in the warp-interleaved layout, 4 in a private area: one IMAD.WIDE per address
    __device__ __forceinline__ void emit(bool on, uint32_t k, uint32_t w) const
    {
#ifdef KPEG_WHATIF_NO_STORE // timing experiment only (results are wrong): how much of the relay is the record stores
        if (on && k < kmax && w == 0x12345u)
#else
        if (on && k < kmax)
#endif
            asm volatile("st.global.u32 [%0], %1;" ::"l"(base + (size_t)k * stride_bytes), "r"(w) : "memory");
    }
};

// decode subsequence `sub` from (p, cz) for a relay pass; emits records when the job uses them
template <class Words>
__device__ __forceinline__ SubState relay_decode(const EntropyArgs &a, const Words &W, const SmemLuts &L, const StreamView &S,
                                                 uint32_t sub, uint32_t end, uint32_t p, uint32_t cz, bool sparse)
{
    DecState d;
    dec_init(d, W, S, p, (cz >> 8) & 3u, cz & 0xFFu, a.seg_hint[sub], 0u);
    if (a.rec) {
        GlobalRecorder R;
        R.base = reinterpret_cast<char *>(a.rec + rec_base_index(sub, a.rec_kmax));
        R.kmax = a.rec_kmax;
        R.stride_bytes = 128u;
        uint32_t area = 0;
        if (sparse) {
            area = a.nrec[sub] >> 10;
            if (area == 0u) {
                const uint32_t idx = atomicAdd(&a.meta->rec_alt_count, 1u);
                area = idx < a.rec_alt_cap ? idx + 1u : 0u;
            }
            if (area) {
                R.base = reinterpret_cast<char *>(a.rec_alt + (size_t)(area - 1u) * a.rec_kmax);
                R.stride_bytes = 4u;
            }
        }
        // registers, not values the compiler recomputes from the kernel parameters for every symbol (it did: seven
        // instructions of address arithmetic per record)
        asm volatile("" : "+l"(R.base), "+r"(R.kmax), "+r"(R.stride_bytes));
#ifdef KPEG_WHATIF_NO_DCS // timing experiment only (results are wrong)
        relay_run<true, false>(d, W, L, S, a.g, end, R);
#else
        relay_run<true, true>(d, W, L, S, a.g, end, R);
#endif
        a.nrec[sub] = min(d.nrec, NREC_MASK) | (area << 10);
        a.dcs[sub] = d.dcs;
        if (d.nrec > a.rec_kmax)
            atomicOr(&a.meta->status, ST_REC_OVERFLOW);
    } else {
        relay_run<false>(d, W, L, S, a.g, end, NoRecorder{});
    }
    return relay_exit_state(d);
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
        if ((old.p != out.p || ((old.cz ^ out.cz) & CZ_STATE_MASK) != 0u) && sub + 1u < nsub)
            list_out[atomicAdd(count_out, 1u)] = sub + 1u;
    }
}

__global__ void __launch_bounds__(ENTROPY_THREADS, 8) entropy_relay_full_kernel(EntropyArgs a, uint32_t wlog)
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
        k1_stage_regions(sm, a, tile, total_bits, wlog);
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
            if (a.rec || (sub != 0u && !(in.p == start && (in.cz & CZ_STATE_MASK) == 0u))) {
                const RegionWords W = k1_region_words(sm, tile, wlog);
                const uint32_t end = min((sub + 1u) << (wlog + 5u), total_bits);
                const SubState out = relay_decode(a, W, L, S, sub, end, in.p, in.cz & CZ_STATE_MASK, false);
                if (sub)
                    relay_publish(a, sub, out, nsub, a.worklist[1], &a.meta->changed[1]);
                else
                    a.state[0] = out; // same state as the cold pass found (the entry is the true one), now with its annotations
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
    __device__ __forceinline__ void pair(uint32_t j, uint32_t &w0, uint32_t &w1) const
    {
        const uint32_t at = addr + 4u * (j - gw0);
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(at));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(at));
    }
};

// the words of one subsequence into its thread's private slice, eight loads in flight
__device__ __forceinline__ void stage_private_words(uint32_t *mine, const uint32_t *words, uint32_t j0, uint32_t n, uint32_t total_words)
{
    for (uint32_t k0 = 0; k0 < n; k0 += 8u) {
        uint32_t v[8];
#pragma unroll
        for (uint32_t i = 0; i < 8u; ++i)
            v[i] = (k0 + i < n && j0 + k0 + i < total_words) ? __ldg(words + j0 + k0 + i) : 0u;
#pragma unroll
        for (uint32_t i = 0; i < 8u; ++i)
            if (k0 + i < n)
                mine[k0 + i] = v[i];
    }
}

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
        stage_private_words(mine, a.words, j0, stride - 1u, total_words);
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
    const SubState out = relay_decode(a, W, L, S, sub, end, inraw.x, inraw.z & CZ_STATE_MASK, true);
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
            stage_private_words(mine, a.words, j0, stride - 1u, total_words);
            W.gw0 = j0;
            const uint32_t end = min((sub + 1u) << (wlog + 5u), total_bits);
            const SubState out = relay_decode(a, W, L, S, sub, end, inraw.x, inraw.z & CZ_STATE_MASK, true);
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
    long long s; // packed DC sums (entropy_core.h, dcs_unpack): reset wherever the slot count is absolute
};

__device__ __forceinline__ SegVal seg_combine(const SegVal &a, const SegVal &b) // a then b
{
    SegVal r;
    r.f = a.f | b.f;
    r.v = b.f ? b.v : a.v + b.v;
    r.s = b.f ? b.s : a.s + b.s;
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
        o.s = __shfl_up_sync(0xffffffffu, x.s, d);
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
            o.s = __shfl_up_sync(0xffffffffu, w.s, d);
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
    e.s = 0;
    if (i < nsub) {
        const SubState st = a.state[i];
        e.f = st.seg >= 0 ? 1u : 0u;
        e.v = st.n + (e.f ? seg_slot_base(a.g, (uint32_t)st.seg) : 0u);
        if (a.rec)
            e.s = a.dcs[i];
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
    if (threadIdx.x == 0) {
        a.scan_tiles[blockIdx.x] = make_uint2(agg.f, agg.v);
        a.scan_tiles_dcs[blockIdx.x] = agg.s;
    }
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
        s_carry.s = 0;
    }
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntiles; t0 += SCAN_THREADS) {
        const uint32_t t = t0 + threadIdx.x;
        SegVal e;
        e.f = 0;
        e.v = 0;
        e.s = 0;
        if (t < ntiles) {
            const uint2 r = a.scan_tiles[t];
            e.f = r.x;
            e.v = r.y;
            e.s = a.scan_tiles_dcs[t];
        }
        SegVal agg;
        const SegVal incl = seg_combine(s_carry, block_seg_scan(e, s_w, agg));
        s_incl[threadIdx.x] = incl;
        __syncthreads();
        if (t < ntiles) {
            const SegVal ex = threadIdx.x == 0 ? s_carry : s_incl[threadIdx.x - 1];
            a.scan_tiles[t] = make_uint2(ex.f, ex.v);
            a.scan_tiles_dcs[t] = ex.s;
            if (t + 1u == ntiles) {
                a.meta->final_slot = incl.v;
                if (incl.v < a.g.total_blocks * 64u)
                    atomicOr(&a.meta->status, ST_SEG_MISMATCH); // the stream ended before the last MCU
            }
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1)
            s_carry = incl;
        __syncthreads();
    }
}

// Third phase.  Besides start_slot[] it is the one place that looks at every subsequence's FINAL relay state, so it
// also (1) turns the annotations of the last decode of every subsequence into device status bits, (2) cross-checks
// the exit state against the slot count (the zig-zag index and component a subsequence ends in must be the ones its
// end slot implies) and the record positions against the restart structure (a subsequence that crossed a boundary
// must end, counting on across it, on the slot the boundary's absolute position gives: T.81 F.2.2.4 -- every
// interval holds exactly Ri MCUs), and (3) notes for every strip of K3 the subsequence its first slot lies in.
__global__ void __launch_bounds__(SCAN_THREADS) entropy_scan_apply_kernel(EntropyArgs a)
{
    __shared__ SegVal s_w[SCAN_THREADS / 32];
    __shared__ SegVal s_incl[SCAN_THREADS];
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
    carry.s = a.scan_tiles_dcs[blockIdx.x];
    incl = seg_combine(carry, incl);
    // incl = absolute slot (and DC predictor values) at the END of subsequence i == entry of subsequence i+1
    if (i + 1u < nsub) {
        a.start_slot[i + 1u] = incl.v;
        a.dcpre[i + 1u] = incl.s;
    }
    if (i == 0u) {
        a.start_slot[0] = 0u;
        a.dcpre[0] = 0;
    }
    s_incl[threadIdx.x] = incl;
    __syncthreads();
    if (i >= nsub)
        return;
    const uint32_t begin = threadIdx.x ? s_incl[threadIdx.x - 1].v : (i ? carry.v : 0u), end = incl.v;
    const SubState st = a.state[i];
    uint32_t status = 0;
    if (a.rec) { // annotations exist only for decodes that emitted records (the Huffman final pass checks for itself)
        status |= (st.cz & CZ_BAD_CODE) ? ST_BAD_CODE : 0u;
        status |= (st.cz & CZ_SLOT_OVERFLOW) ? ST_SLOT_OVERFLOW : 0u;
        status |= (st.cz & CZ_SEG_MISMATCH) ? ST_SEG_MISMATCH : 0u;
        const uint32_t total_slots = a.g.total_blocks * 64u;
        // surplus symbols after the last MCU of the stream are not an error (as in the Huffman final pass); a stream that
        // ENDS before its last MCU is: its last symbol straddles the end like padding would, but too early
        const uint32_t reached = (begin & ~63u) + (st.cz >> CZ_POS_SHIFT);
        if (st.seg >= 0 && reached != end && !((uint32_t)st.seg >= a.g.nseg && end >= total_slots && reached >= total_slots))
            status |= ST_SEG_MISMATCH;
        if ((end & 63u) != (st.cz & 63u) || ((end >> 6) % a.g.ncomp) != ((st.cz >> 8) & 3u))
            status |= ST_EXIT_MISMATCH;
    }
    if (status)
        atomicOr(&a.meta->status, status);
    // strips of K3 whose first slot lies in [begin, end): a subsequence spans at most a handful
    if (a.strip_sub && end > begin) {
        const uint32_t sps = a.strip_slots;
        uint32_t s0 = (begin + sps - 1u) / sps;
        for (uint32_t k = 0; k < 64u && s0 < a.nstrips && s0 * sps < end; ++k, ++s0)
            a.strip_sub[s0] = i;
    }
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
                c = (in.cz >> 8) & 3u;
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
                write_run(d, W, L, S, a.g, end, limit, sink);
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
            if (out.p != rec.p || out.cz != (rec.cz & CZ_STATE_MASK))
                st |= ST_EXIT_MISMATCH;
            if (sub + 1u == nsub && a.meta->final_slot < total_slots)
                st |= ST_SEG_MISMATCH; // the stream ended before the last MCU
        }
        __syncthreads();
    }
    if (st)
        atomicOr(&a.meta->status, st);
}

// ---- fallback DC prediction -----------------------------------------------------------------------------
// Only behind the Huffman final pass (degenerate tables, KPEG_NO_RECORDS=1): MCU.cpp:107-108 over the DC differences
// entropy_write left, one CTA per restart segment (or image), 256 MCUs per step.  The record path needs none of this:
// its predictors come out of the offset scan.
__global__ void __launch_bounds__(256) dc_integrate_kernel(JobGeom g, const int16_t *dcdiff, int16_t *dc)
{
    __shared__ int s_w[8][3];
    __shared__ int s_carry[3];
    const uint32_t seg = blockIdx.x, img = seg / g.segs_per_image, r = seg - img * g.segs_per_image;
    const uint32_t m0 = img * g.mcus_per_image + r * g.restart_interval; // restart_interval == 0: one segment per image
    const uint32_t m1 = min(img * g.mcus_per_image + (g.restart_interval ? (r + 1u) * g.restart_interval : g.mcus_per_image),
                            (img + 1u) * g.mcus_per_image);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3)
        s_carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t mb = m0; mb < m1; mb += 256u) {
        const uint32_t m = mb + threadIdx.x;
        int v[3] = {0, 0, 0};
        if (m < m1)
            for (uint32_t c = 0; c < g.ncomp; ++c)
                v[c] = dcdiff[m * g.ncomp + c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int o = __shfl_up_sync(0xffffffffu, v[c], d);
                if (lane >= d)
                    v[c] += o;
            }
        if (lane == 31)
            for (int c = 0; c < 3; ++c)
                s_w[warp][c] = v[c];
        __syncthreads();
        int before[3] = {s_carry[0], s_carry[1], s_carry[2]};
        for (int w = 0; w < warp; ++w)
            for (int c = 0; c < 3; ++c)
                before[c] += s_w[w][c];
        if (m < m1)
            for (uint32_t c = 0; c < g.ncomp; ++c)
                dc[m * g.ncomp + c] = (int16_t)(v[c] + before[c]);
        __syncthreads();
        if (threadIdx.x == 255)
            for (int c = 0; c < 3; ++c)
                s_carry[c] = v[c] + before[c];
        __syncthreads();
    }
}

void launch_dc_integrate(const JobGeom &g, const int16_t *dcdiff, int16_t *dc, cudaStream_t s, uint32_t *launches)
{
    dc_integrate_kernel<<<g.nseg, 256, 0, s>>>(g, dcdiff, dc);
    ++*launches;
}

static uint32_t ilog2(uint32_t v)
{
    uint32_t l = 0;
    while ((1u << l) < v)
        ++l;
    return l;
}

static uint32_t g_sm_count = 148;
static uint32_t g_k1_grid_cap = 148 * 8;
static uint32_t g_k1_write_grid_cap = 148 * 4;

// CTAs of the persistent cold / relay kernels: what is resident at this subsequence size (the stream regions grow with
// it: 8 CTAs per SM at 512 bits, 6 at 1024), so that the grid is one wave
static uint32_t k1_grid(const EntropyArgs &a)
{
    static uint32_t cap_for[16] = {};
    const uint32_t wlog = ilog2(a.g.sub_bits / 32u) & 15u;
    if (cap_for[wlog] == 0u) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_relay_full_kernel, ENTROPY_THREADS,
                                                          k1_region_smem_bytes(a.g.sub_bits)) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        if (const char *e = getenv("KPEG_K1_PER_SM")) // experiments: fewer CTAs per SM leave room for the other lanes' kernels
            per_sm = std::max(1, std::min(per_sm, atoi(e)));
        cap_for[wlog] = g_sm_count * (uint32_t)per_sm;
    }
    const uint32_t tiles = (a.nsub_max + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
    const uint32_t cap = cap_for[wlog] < g_k1_grid_cap ? cap_for[wlog] : g_k1_grid_cap;
    if (tiles <= cap)
        return tiles ? tiles : 1u;
    // every CTA the same number of tiles: with 3.5 tiles per CTA the last of four rounds would run half empty
    const uint32_t rounds = (tiles + cap - 1u) / cap;
    return (tiles + rounds - 1u) / rounds;
}

void launch_entropy_cold(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
    entropy_cold_kernel<<<k1_grid(a), ENTROPY_THREADS, k1_region_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    ++*launches;
}

void launch_entropy_relay(const EntropyArgs &a, int round, cudaStream_t s, uint32_t *launches)
{
    if (round == 1) {
        const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
        entropy_relay_full_kernel<<<k1_grid(a), ENTROPY_THREADS, k1_region_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    } else {
        const uint32_t grid = (a.nsub_max + ENTROPY_THREADS - 1) / ENTROPY_THREADS;
        const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
        entropy_relay_sparse_kernel<<<grid, ENTROPY_THREADS, k1_sparse_smem_bytes(a.g.sub_bits), s>>>(
            a, round, relay_slot(round - 1), relay_slot(round), wlog);
    }
    ++*launches;
}

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

void launch_entropy_write(const EntropyArgs &a, cudaStream_t s, uint32_t *launches)
{
    const uint32_t wlog = ilog2(a.g.sub_bits / 32u);
    const uint32_t tiles = (a.nsub_max + WRITE_THREADS - 1) / WRITE_THREADS;
    const uint32_t grid = tiles < g_k1_write_grid_cap ? tiles : g_k1_write_grid_cap;
    entropy_write_kernel<<<grid, WRITE_THREADS, k1_write_smem_bytes(a.g.sub_bits), s>>>(a, wlog);
    ++*launches;
}

void kernels_configure(int max_concurrent_jobs)
{
    g_max_concurrent_loops = (uint32_t)(max_concurrent_jobs > 0 ? max_concurrent_jobs : 1);
    const int k1max = (int)k1_region_smem_bytes(1024);
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
                                                          k1_region_smem_bytes(512)) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        g_k1_grid_cap = (uint32_t)(sms * per_sm);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_write_kernel, WRITE_THREADS,
                                                          k1_write_smem_bytes(512)) != cudaSuccess || per_sm < 1)
            per_sm = 3;
        g_k1_write_grid_cap = (uint32_t)(sms * per_sm);
        cudaFuncSetAttribute(entropy_relay_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)k1_sparse_smem_bytes(1024));
        g_sm_count = (uint32_t)sms;
        if (const char *e = getenv("KPEG_RELAY_LOOP_PER_SM"))
            g_relay_loop_per_sm_cap = atoi(e);
        if (const char *e = getenv("KPEG_RELAY_SPIN_LIMIT")) // tests: 0 = give up at the first poll that finds a CTA missing
            g_relay_spin_limit = (uint32_t)strtoul(e, nullptr, 10);
    }
    k3_configure();
}

} // namespace kpeg
