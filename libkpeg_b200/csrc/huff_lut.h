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

// Returns 0 on success, -1 if the counts over-subscribe the code space or exceed 256 symbols.
inline int build_huff_lut(const uint8_t counts[16], const uint8_t *symbols, bool is_ac, HuffLut *out)
{
    memset(out, 0, sizeof *out);
    out->is_ac = is_ac ? 1u : 0u;
    uint32_t code = 0, idx = 0;
    for (int L = 1; L <= 16; ++L) {
        out->first_code[L] = (uint16_t)code;
        out->first_idx[L] = (uint16_t)idx;
        for (uint32_t k = 0; k < counts[L - 1]; ++k) {
            if (idx >= 256 || code >= (1u << L))
                return -1;
            uint8_t sym = symbols[idx];
            out->symbols[idx] = sym;
            if (L <= LUT_BITS) {
                uint32_t e = pack_entry((uint32_t)L, sym, is_ac);
                uint32_t lo = code << (LUT_BITS - L), hi = (code + 1) << (LUT_BITS - L);
                for (uint32_t w = lo; w < hi; ++w)
                    out->fast[w] = (uint16_t)e;
            }
            ++code;
            ++idx;
        }
        out->bound[L] = code << (16 - L); // codes of length <= L cover left-aligned windows [0, bound[L])
        code <<= 1;
    }
    out->bound[0] = 0;
    out->bound[17] = 0;
    return 0;
}

} // namespace kpeg
#endif
