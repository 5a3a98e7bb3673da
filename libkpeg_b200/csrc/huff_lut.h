// huff_lut.h -- host-side construction of the device Huffman lookup tables from a DHT table.
//
// Canonical code assignment is the one HuffmanTree::constructHuffmanTree produces level by level
// (reference src/HuffmanTree.cpp:106-157; == ITU-T T.81 Annex C): code = 0; for each length,
// every symbol takes `code++`; then code <<= 1.  Symbol semantics follow decodeScanData
// (src/Decoder.cpp:706-803): DC symbol -> category = sym & 15; AC symbol -> run = sym >> 4,
// category = sym & 15, 0x00 = EOB, 0xF0 = ZRL.
#ifndef KPEG_HUFF_LUT_H
#define KPEG_HUFF_LUT_H

#include <string.h>

#include "kpeg_common.h"

namespace kpeg {

// Fills table `ti` of `set` and its canonical twin.  Returns 0 on success, -1 if the counts
// over-subscribe the code space or exceed 256 symbols.
inline int build_huff_lut(const uint8_t counts[16], const uint8_t *symbols, bool is_ac, LutSet *set, int ti, HuffCanon *canon)
{
    memset(canon, 0, sizeof *canon);
    memset(set->fast[ti], 0, sizeof set->fast[ti]);
    memset(set->longlut[ti], 0, sizeof set->longlut[ti]);
    canon->is_ac = is_ac ? 1u : 0u;
    uint32_t code = 0, idx = 0;
    for (int L = 1; L <= 16; ++L) {
        canon->first_code[L] = (uint16_t)code;
        canon->first_idx[L] = (uint16_t)idx;
        for (uint32_t k = 0; k < counts[L - 1]; ++k) {
            if (idx >= 256 || code >= (1u << L))
                return -1;
            canon->symbols[idx] = symbols[idx];
            ++code;
            ++idx;
        }
        canon->bound[L] = code << (16 - L); // codes of length <= L cover left-aligned windows [0, bound[L])
        code <<= 1;
    }
    // first level: every code of length <= LUT_BITS
    idx = 0;
    for (int L = 1; L <= LUT_BITS; ++L) {
        for (uint32_t k = 0; k < counts[L - 1]; ++k, ++idx) {
            const uint32_t c = canon->first_code[L] + k;
            const uint32_t e = pack_entry((uint32_t)L, symbols[idx], is_ac);
            for (uint32_t w = c << (LUT_BITS - L); w < (c + 1) << (LUT_BITS - L); ++w)
                set->fast[ti][w] = (uint16_t)e;
        }
    }
    // second level: one 64-entry sub-table per 10-bit prefix that starts a longer code
    set->long_n[ti] = 0;
    for (uint32_t prefix = canon->bound[LUT_BITS] >> (16 - LUT_BITS); prefix < (uint32_t)LUT_SIZE; ++prefix) {
        // does any code start with this prefix?  (windows [prefix<<6, (prefix+1)<<6) below bound[16])
        if ((prefix << (16 - LUT_BITS)) >= canon->bound[16])
            break;
        if (set->long_n[ti] + (uint32_t)SUB_SIZE > (uint32_t)LONG_CAP)
            break; // no room: these prefixes keep entry 0 and use the canonical search
        const uint32_t sub = set->long_n[ti] / (uint32_t)SUB_SIZE;
        for (uint32_t low = 0; low < (uint32_t)SUB_SIZE; ++low) {
            const uint32_t w = (prefix << (16 - LUT_BITS)) | low;
            uint32_t e = ENTRY_INVALID;
            for (int L = LUT_BITS + 1; L <= 16; ++L) {
                if (w < canon->bound[L]) {
                    const uint32_t i = canon->first_idx[L] + ((w >> (16 - L)) - canon->first_code[L]);
                    e = pack_entry((uint32_t)L, canon->symbols[i & 255u], is_ac);
                    break;
                }
            }
            set->longlut[ti][sub * (uint32_t)SUB_SIZE + low] = (uint16_t)e;
        }
        set->fast[ti][prefix] = (uint16_t)((sub + 1u) << 5);
        set->long_n[ti] += (uint32_t)SUB_SIZE;
    }
    return 0;
}

} // namespace kpeg
#endif
