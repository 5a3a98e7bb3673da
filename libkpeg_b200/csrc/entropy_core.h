// entropy_core.h -- the Huffman / run-length symbol loop of the B200 decode path, as host+device
// inline functions (the kernels in kernels.cu are thin wrappers that map threads to subsequences).
//
// What it computes is what JPEGDecoder::decodeScanData does bit by bit on a '0'/'1' string
// (reference src/Decoder.cpp:655-855, HuffmanTree::contains src/HuffmanTree.cpp:164-193,
// bitStringtoValue src/Image.cpp:285-302), restated for a massively parallel decoder:
//
//  * the stream (already unstuffed, src/Decoder.cpp:621-653, and stored as big-endian 32-bit words)
//    is cut into fixed-size SUBSEQUENCES; one thread decodes one subsequence;
//  * a decoder state is (bit position p, component c, zig-zag index z).  A thread that does not
//    know its true entry state starts "cold" at (sub*B, 0, 0); JPEG Huffman streams
//    self-synchronise, so after a few hundred bits it is in the true state with high probability.
//    Certainty comes from the relay fixed point computed in kernels.cu: X[i] = decode(i, X[i-1]);
//  * restart-interval (and, in batch mode, image) boundaries are known bit positions at which the
//    state is known to be (boundary, 0, 0) and the absolute output position is known as well; a
//    decoder whose next symbol would straddle the boundary jumps onto it (T.81 F.2.2.4 / E.2.4
//    semantics: the padding 1-bits before a marker can never complete a code).
#ifndef KPEG_ENTROPY_CORE_H
#define KPEG_ENTROPY_CORE_H

#include "kpeg_common.h"

namespace kpeg {

KPEG_HD uint32_t ld_word(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

KPEG_HD uint32_t funnel_left(uint32_t hi, uint32_t lo, uint32_t s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}

// Three-word sliding window over the big-endian word stream; the third word is fetched one
// word ahead of use so the load is off the symbol-to-symbol dependency chain.
struct BitWindow {
    const uint32_t *words;
    uint32_t j, w0, w1, w2;
    KPEG_HD void seek(uint32_t p)
    {
        j = p >> 5;
        w0 = ld_word(words + j);
        w1 = ld_word(words + j + 1);
        w2 = ld_word(words + j + 2);
    }
    // 32 bits starting at bit p; p may have advanced by at most one word since the last call.
    KPEG_HD uint32_t peek(uint32_t p)
    {
        uint32_t jj = p >> 5;
        if (jj != j) {
            w0 = w1;
            w1 = w2;
            j = jj;
            w2 = ld_word(words + j + 2);
        }
        return funnel_left(w0, w1, p & 31u);
    }
};

// Codes longer than LUT_BITS: canonical search.  `win` holds the next 32 stream bits, MSB first.
KPEG_HD uint32_t huff_slow_lookup(const HuffCanon &L, uint32_t win)
{
    const uint32_t w16 = win >> 16;
    for (int len = LUT_BITS + 1; len <= 16; ++len) {
        if (w16 < L.bound[len]) {
            uint32_t idx = L.first_idx[len] + ((w16 >> (16 - len)) - L.first_code[len]);
            return pack_entry((uint32_t)len, L.symbols[idx & 255u], L.is_ac != 0);
        }
    }
    return ENTRY_INVALID;
}

// T.81 F.2.2.1 EXTEND == bitStringtoValue (src/Image.cpp:285-302): leading 1 -> the value itself,
// leading 0 -> value - (2^n - 1).
KPEG_HD int32_t extend_value(uint32_t v, uint32_t n)
{
    return (v >> (n - 1u)) ? (int32_t)v : (int32_t)v - (int32_t)((1u << n) - 1u);
}

struct StreamView {
    const uint32_t *words;   // unstuffed stream, big-endian 32-bit words, >= 3 words of slack after the end
    const uint32_t *seg_bit; // [nseg + 2]: start bit of every restart segment, then total_bits, then 0xFFFFFFFF
    uint32_t total_bits;
};

// Decode from state (p, c, z) until the first symbol boundary at or after `end_bit`.
//   WRITE == false : speculative / relay pass, only the exit state and slot count are produced.
//   WRITE == true  : final pass; `slot` is the absolute coefficient slot at entry, AC coefficients
//                    go to coef[] (zig-zag order, buffer pre-zeroed), DC differences to dcdiff[].
// `k` is a hint: any segment index whose start bit is <= the first boundary after p.
template <bool WRITE>
KPEG_HD SubState decode_span(const StreamView &S, const JobGeom &g, const LutSet &luts, const HuffCanon *canon, uint32_t end_bit,
                             uint32_t p, uint32_t c, uint32_t z, uint32_t k, uint32_t slot, int16_t *coef,
                             int16_t *dcdiff, uint32_t *status_accum)
{
    const uint32_t total_slots = g.total_blocks * 64u;
    const uint32_t nc = g.ncomp;
    while (S.seg_bit[k] <= p)
        ++k;
    uint32_t segend = S.seg_bit[k];
    uint32_t n = 0;
    int32_t seg = -1;
    uint32_t st = 0;
    BitWindow bw;
    bw.words = S.words;
    bw.seek(p);
    while (p < end_bit) {
        const uint32_t win = bw.peek(p);
        const uint32_t ti = c * 2u + (z != 0u ? 1u : 0u);
        uint32_t e = luts.fast[ti][win >> (32 - LUT_BITS)];
        if (e == 0u) { // code longer than LUT_BITS (or no code at all)
            const uint32_t li = (win >> 16) - luts.long_base[ti];
            e = li < luts.long_n[ti] ? (uint32_t)luts.longlut[ti][li] : huff_slow_lookup(canon[ti], win);
        }
        const uint32_t len = e & 31u, size = (e >> 5) & 15u, adv = e >> 9;
        const uint32_t T = len + size;
        if (p + T > segend) {
            // The symbol would straddle a restart / image boundary: we are in its padding.
            if (WRITE && slot != seg_slot_base(g, k) && (k < g.nseg || slot < total_slots))
                st |= ST_SEG_MISMATCH;
            p = segend;
            c = 0;
            z = 0;
            n = 0;
            seg = (int32_t)k;
            if (WRITE)
                slot = seg_slot_base(g, k);
            ++k;
            segend = S.seg_bit[k];
            if (p >= S.total_bits)
                break;
            bw.seek(p);
            continue;
        }
        if (WRITE) {
            if (len > 16u)
                st |= ST_BAD_CODE;
            if (size) {
                const uint32_t raw = (win << len) >> (32u - size);
                const int32_t val = extend_value(raw, size);
                if (slot + adv > total_slots || z + adv > 64u)
                    st |= ST_SLOT_OVERFLOW;
                else if (z == 0u)
                    dcdiff[slot >> 6] = (int16_t)val;
                else
                    coef[slot + adv - 1u] = (int16_t)val;
            }
        }
        p += T;
        uint32_t zn = z + adv;
        zn = zn > 64u ? 64u : zn;
        n += zn - z;
        if (WRITE)
            slot += zn - z;
        z = zn;
        if (z == 64u) {
            z = 0;
            c = (c + 1u == nc) ? 0u : c + 1u;
        }
    }
    if (WRITE && st)
        *status_accum |= st;
    SubState out;
    out.p = p;
    out.n = n;
    out.cz = (c << 8) | z;
    out.seg = seg;
    return out;
}

} // namespace kpeg
#endif
