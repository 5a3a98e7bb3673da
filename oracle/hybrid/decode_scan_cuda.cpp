// decode_scan_cuda.cpp -- TEST INFRASTRUCTURE: the member function INTEGRATION.md section 2 shows, compiled for real.
//
// oracle/hybrid/build.py copies the reference's sources into a scratch directory, applies the three small edits the
// section lists (a byte copy of the scan next to the '0'/'1' string in scanImageData, src/Decoder.cpp:558-573; the call
// in decodeImageFile, :137-138; a one-line setter in Image.hpp) and compiles them together with THIS file against
// lib/libkpeg_cuda.so: the reference's own parser, logger, Image and PPM writer around this repo's CUDA hot path.
// tests/test_gpu_parity.py::test_reference_with_cuda_hot_path runs the result on lena.jpg and compares the PPM it
// writes with the golden hash of the unmodified reference.
#include <memory>
#include <vector>

#include "Decoder.hpp"   // the reference's, with m_scanBytes / decodeScanDataCUDA() added by build.py
#include "kpeg_cuda.h"

bool kpeg::JPEGDecoder::decodeScanDataCUDA()
{
    kpeg_plan plan{};                                   // POD, see include/kpeg_cuda.h
    plan.width  = (uint16_t)m_image.getWidth();         // set by parseSOF0Segment, src/Decoder.cpp:325-333
    plan.height = (uint16_t)m_image.getHeight();
    plan.ncomp  = 3;                                    // the reference decodes 3 x (1x1) only, :339-356
    plan.flags  = KPEG_FLAG_REF_PARITY;                 // reproduce src/MCU.cpp:97-104 exactly
    for (int c = 0; c < 3; ++c)                         // hard-wired selectors of src/Decoder.cpp:704 / src/MCU.cpp:110
        plan.comp_tq[c] = plan.comp_td[c] = plan.comp_ta[c] = (uint8_t)(c == 0 ? 0 : 1);
    for (int t = 0; t < 2 && t < (int)m_QTables.size(); ++t) {   // m_QTables[t]: 64 UInt16 in zig-zag order, :263-277
        plan.qt_present[t] = 1;
        for (int i = 0; i < 64; ++i)
            plan.qt[t][i] = m_QTables[t][i];
    }
    for (int cls = 0; cls < 2; ++cls)                   // m_huffmanTable[class][id]: array<pair<count, symbols>, 16>
        for (int id = 0; id < 2; ++id) {
            kpeg_huff_spec& h = plan.ht[cls][id];
            int k = 0;
            for (int L = 0; L < 16; ++L) {
                h.counts[L] = (uint8_t)m_huffmanTable[cls][id][L].first;
                for (auto s : m_huffmanTable[cls][id][L].second)
                    h.symbols[k++] = (uint8_t)s;
            }
            plan.ht_present[cls][id] = 1;
        }

    kpeg_ctx* ctx = nullptr;
    if (kpeg_cuda_acquire(0, &ctx) != KPEG_OK)
        return false;                                   // no GPU: no fallback, report the error
    std::vector<uint8_t> rgb((size_t)plan.width * plan.height * 3);
    // m_scanBytes: everything scanImageData read up to, not including, the EOI marker
    const int rc = kpeg_cuda_decode(ctx, &plan, m_scanBytes.data(), m_scanBytes.size(), rgb.data(), nullptr);
    kpeg_cuda_release(0, ctx);
    if (rc != KPEG_OK)
        return false;

    // hand the pixels to Image in the layout dumpRawData reads (src/Image.cpp:129-135; include/Types.hpp:52-76)
    auto px = std::make_shared<std::vector<std::vector<Pixel>>>(plan.height, std::vector<Pixel>(plan.width));
    for (unsigned y = 0; y < plan.height; ++y)
        for (unsigned x = 0; x < plan.width; ++x) {
            const uint8_t* p = &rgb[((size_t)y * plan.width + x) * 3];
            (*px)[y][x].comp[0] = p[0];
            (*px)[y][x].comp[1] = p[1];
            (*px)[y][x].comp[2] = p[2];
        }
    m_image.setPixelPtr(px);
    return true;
}
