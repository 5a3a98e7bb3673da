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

} // namespace kpeg
#endif
