// unstuff_core.h -- byte classification of an entropy-coded segment (host + device inline).
//
// Byte classes (T.81 B.1.1.5): FF 00 is a data byte FF (the 00 is dropped -- what
// JPEGDecoder::byteStuffScanData does with std::string::erase, reference src/Decoder.cpp:621-653);
// FF D0..D7 is a restart marker, i.e. a segment boundary (the reference cannot handle it,
// SURVEY F2); FF FF is a fill byte; any other marker inside the segment is flagged.
#ifndef KPEG_UNSTUFF_CORE_H
#define KPEG_UNSTUFF_CORE_H

#include <string.h>

#include "kpeg_common.h"

namespace kpeg {

KPEG_HD uint8_t ld_byte(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

KPEG_HD void ld_16bytes(const uint8_t *p, uint32_t b[4])
{
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    b[0] = v.x;
    b[1] = v.y;
    b[2] = v.z;
    b[3] = v.w;
#else
    memcpy(b, p, 16);
#endif
}

struct ByteClass {
    uint32_t keep; // bit i: byte i survives
    uint32_t rst;  // bit i: byte i is the second byte of an RSTn marker
    uint32_t bad;  // any unexpected marker
    uint32_t b[4]; // the 16 bytes
};

KPEG_HD ByteClass classify16(const uint8_t *scan, uint32_t len, uint32_t base)
{
    ByteClass c;
    c.keep = c.rst = c.bad = 0;
    if (base + 16u <= len && ((reinterpret_cast<uintptr_t>(scan + base) & 15u) == 0)) {
        ld_16bytes(scan + base, c.b);
    } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t x = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t idx = base + w * 4 + k;
                const uint32_t byte = idx < len ? (uint32_t)ld_byte(scan + idx) : 0xFFu;
                x |= byte << (8 * k);
            }
            c.b[w] = x;
        }
    }
    uint32_t prev = base > 0u && base <= len ? (uint32_t)ld_byte(scan + base - 1) : 0u;
    const uint32_t next16 = base + 16u < len ? (uint32_t)ld_byte(scan + base + 16) : 0xFFu;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t cur = (c.b[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const uint32_t nxt = i < 15 ? (c.b[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu : next16;
        const bool in = base + i < len;
        const bool nxt_in = base + i + 1 < len;
        bool keep;
        if (cur == 0xFFu) {
            const uint32_t n = nxt_in ? nxt : 0xFFu;
            keep = (n == 0x00u);
            if (in && !(n == 0x00u || n == 0xFFu || (n >= 0xD0u && n <= 0xD7u) || n == 0xD9u))
                c.bad = 1;
        } else if (prev == 0xFFu) {
            keep = false;
            if (in && cur >= 0xD0u && cur <= 0xD7u)
                c.rst |= 1u << i;
        } else {
            keep = true;
        }
        if (in && keep)
            c.keep |= 1u << i;
        prev = cur;
    }
    return c;
}

// ---- the same classification, sixteen bytes at a time ------------------------------------------------
// Byte-parallel form for a chunk that lies entirely inside the segment; b[] holds the bytes little-endian
// (byte i = b[i >> 2] >> 8 * (i & 3)), prev / next are the bytes on either side (next = 0xFF past the end).
// Masks carry 0x80 in the byte lanes where a predicate holds.
KPEG_HD uint32_t swar_zero_bytes(uint32_t v) // exact: 0x80 where the byte is 0
{
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu);
}

KPEG_HD uint32_t swar_nibble(uint32_t m) // 0x80-per-byte mask -> 4 bits, byte 0 in bit 0
{
    return (((m >> 7) * 0x00204081u) >> 21) & 0xFu;
}

KPEG_HD void classify16_swar(const uint32_t b[4], uint32_t prev, uint32_t next, uint32_t &keep, uint32_t &rst, uint32_t &bad)
{
    uint32_t ff[4], zz[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        ff[w] = swar_zero_bytes(~b[w]);
        zz[w] = swar_zero_bytes(b[w]);
    }
    keep = 0;
    rst = 0;
    bad = 0;
    uint32_t second = 0; // bytes that follow an FF and are neither 00 nor FF: the second byte of a marker
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t fprev = (ff[w] << 8) | (w ? ff[w - 1] >> 24 : (prev == 0xFFu ? 0x80u : 0u));
        const uint32_t znext = (zz[w] >> 8) | (w < 3 ? zz[w + 1] << 24 : (next == 0x00u ? 0x80000000u : 0u));
        // an FF survives iff a stuffed 00 follows; any other byte survives iff it does not follow an FF
        const uint32_t k = (ff[w] & znext) | (~ff[w] & ~fprev & 0x80808080u);
        keep |= swar_nibble(k) << (4 * w);
        second |= ~ff[w] & ~zz[w] & fprev;
    }
    if (second) { // rare: markers
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t fprev = (ff[w] << 8) | (w ? ff[w - 1] >> 24 : (prev == 0xFFu ? 0x80u : 0u));
            const uint32_t sec = ~ff[w] & ~zz[w] & fprev;
            const uint32_t d0d7 = swar_zero_bytes((b[w] & 0xF8F8F8F8u) ^ 0xD0D0D0D0u);
            const uint32_t d9 = swar_zero_bytes(b[w] ^ 0xD9D9D9D9u);
            rst |= swar_nibble(sec & d0d7) << (4 * w);
            bad |= sec & ~d0d7 & ~d9;
        }
    }
    // (an unexpected marker is flagged where its second byte lies; classify16 flags it at the FF -- the
    // status word is the OR over all chunks either way)
    bad = bad ? 1u : 0u;
}

} // namespace kpeg
#endif
