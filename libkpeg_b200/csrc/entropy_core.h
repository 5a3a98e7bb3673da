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

#ifndef KPEG_BLOCK_END_BRANCH
#define KPEG_BLOCK_END_BRANCH 0 // 1: the end of a block as a branch of the symbol loop (relay_run); 0: selects
#endif

namespace kpeg {

KPEG_HD uint32_t max_u32(uint32_t a, uint32_t b) { return a > b ? a : b; }

KPEG_HD uint32_t funnel_left(uint32_t hi, uint32_t lo, uint32_t s) // the count is taken modulo 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}

// Codes longer than LUT_BITS that the second-level table does not cover: canonical search.
// `win` holds the next 32 stream bits, MSB first.
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

// ---- accessors -------------------------------------------------------------------------------------
// decode_span is a template over where the stream words and the lookup tables live: the kernels
// pass shared-memory accessors (kernels.cu), the CPU single-stepper plain arrays (below).

struct PlainWords { // word `gw` of the big-endian stream from a plain array
    const uint32_t *words;
    KPEG_HD uint32_t operator()(uint32_t gw) const
    {
#if defined(__CUDA_ARCH__)
        return __ldg(words + gw);
#else
        return words[gw];
#endif
    }
    // words j and j + 1: the 64 bits a symbol that begins in word j can touch
    KPEG_HD void pair(uint32_t j, uint32_t &w0, uint32_t &w1) const
    {
        w0 = (*this)(j);
        w1 = (*this)(j + 1u);
    }
};

struct PlainLuts {
    const LutSet *set;
    const HuffCanon *canon;
    // toff = table index * LUT_SIZE
    KPEG_HD uint32_t fast(uint32_t toff, uint32_t idx) const { return (&set->fast[0][0])[toff + idx]; }
    // e = first-level entry with (e & 31) == 0: pointer to a sub-table, or 0
    KPEG_HD uint32_t slow(uint32_t toff, uint32_t win, uint32_t e) const
    {
        const uint32_t ti = toff >> LUT_BITS;
        if (e)
            return set->longlut[ti][((e >> 5) - 1u) * (uint32_t)SUB_SIZE + ((win >> 16) & (uint32_t)(SUB_SIZE - 1))];
        return huff_slow_lookup(canon[ti], win);
    }
};

// T.81 F.2.2.1 EXTEND == bitStringtoValue (src/Image.cpp:285-302): leading 1 -> the value itself,
// leading 0 -> value - (2^n - 1).
KPEG_HD int32_t extend_value(uint32_t v, uint32_t n)
{
    return (v >> (n - 1u)) ? (int32_t)v : (int32_t)v - (int32_t)((1u << n) - 1u);
}

struct StreamView {
    const uint32_t *seg_bit; // [nseg + 2]: start bit of every restart segment, then total_bits, then 0xFFFFFFFF
    uint32_t total_bits;
};

// ---- coefficient sinks (final pass) ------------------------------------------------------------------
// put(is_dc, slot, adv, valid, value): a DC difference belongs to block slot >> 6, an AC coefficient
// to absolute slot (slot + adv - 1); `valid` is false for symbols that carry no value (EOB, ZRL).
struct NullSink {
    static constexpr bool bounds_itself = false;
    KPEG_HD void put(bool, uint32_t, uint32_t, bool, int32_t) const {}
};

// straight to global memory: coef[] must be zero-filled beforehand
struct GlobalSink {
    static constexpr bool bounds_itself = false;
    int16_t *coef;   // [blocks][64]
    int16_t *dcdiff; // [blocks]
    KPEG_HD void put(bool is_dc, uint32_t slot, uint32_t adv, bool valid, int32_t v) const
    {
        if (!valid)
            return;
        if (is_dc)
            dcdiff[slot >> 6] = (int16_t)v;
        else if ((slot & 63u) + adv <= 64u)
            coef[slot + adv - 1u] = (int16_t)v;
    }
};

// ---- coefficient records -------------------------------------------------------------------------
// The relay passes decode every subsequence from (what converges to) its true entry state anyway; they also write
// down what they decoded, so that nothing downstream has to Huffman-decode a third time.  One 32-bit record per
// symbol that CARRIES A VALUE (EOB, ZRL and zero DC differences leave no record):
//     bits 31..16  position: coefficient slot counted from the first slot of the block the subsequence was entered in
//                  (slot = block * 64 + zig-zag index; a DC difference sits in slot 0 of its block)
//     bits 15..0   value + 0x8000 (EXTEND already applied: T.81 F.2.2.1 == bitStringtoValue, src/Image.cpp:285-302)
// A record is self-contained: absolute slot = (start_slot[subsequence] & ~63) + position.  The consumer (K3's
// expansion stage, kernels.cu) needs no running state, so any thread can take any record.  Positions run on
// ACROSS restart / image boundaries: in a valid stream a boundary falls exactly where the slot count says the
// interval ends, so crossing it only rounds the position up to the next block; whether that held is checked once
// per subsequence by the offset scan (RELAY_* annotations below), not per record.
KPEG_HD uint32_t record_pack(uint32_t pos, uint32_t biased_value) { return (pos << 16) + biased_value; }
KPEG_HD uint32_t record_pos(uint32_t r) { return r >> 16; }
KPEG_HD int32_t record_value(uint32_t r) { return (int32_t)(r & 0xFFFFu) - 0x8000; }
constexpr uint32_t COEF_BIAS = 0x8000u; // coefficients travel as value + COEF_BIAS in 16 bits (records, K3's tile)

// value + COEF_BIAS of a symbol with `size` (1..15) magnitude bits that follow the code in the left-aligned window
// (size == 0 is tolerated -- the result is then meaningless and the caller does not store it)
KPEG_HD uint32_t biased_extend(uint32_t win, uint32_t T, uint32_t size)
{
    const uint32_t x = win << (T - size);            // magnitude bits left-aligned
    const uint32_t s = (uint32_t)((int32_t)x >> 31); // all ones: leading 1, the value itself; zero: leading 0, value - (2^size - 1)
    const uint32_t t = x ^ ~s;                       // leading 0: complemented, so that the field below is (2^size - 1) - raw
#if defined(__CUDA_ARCH__)
    const uint32_t u = __funnelshift_rc(t, 0u, 32u - size); // one clamped shift (size == 0: 32 bits, i.e. 0)
#else
    const uint32_t u = (t >> 16) >> (16u - size);
#endif
    // raw, or -((2^size - 1) - raw): one multiply-add by +-1 with the bias as the addend
    return u * (~s | 1u) + COEF_BIAS;
}

// SubState::cz = (component << 8) | zig-zag index in the low 10 bits -- the decoder STATE, the part the relay's
// fixed point is about -- plus annotations of the decode that produced it (they never decide convergence):
constexpr uint32_t CZ_STATE_MASK = 0x3FFu;
constexpr uint32_t CZ_BAD_CODE = 1u << 10;      // a bit pattern that is no Huffman code was consumed
constexpr uint32_t CZ_SLOT_OVERFLOW = 1u << 11; // a coefficient beyond slot 63 of its block
constexpr uint32_t CZ_SEG_MISMATCH = 1u << 12;  // an interval lying wholly inside the subsequence had the wrong number of MCUs
constexpr uint32_t CZ_FLAGS = CZ_BAD_CODE | CZ_SLOT_OVERFLOW | CZ_SEG_MISMATCH;
constexpr uint32_t CZ_POS_SHIFT = 16;           // bits 31..16: position (as in a record) the subsequence ended on

struct NoRecorder {
    KPEG_HD void emit(bool, uint32_t, uint32_t) const {}
};

// Resumable decoder state of one subsequence.
struct DecState {
    uint32_t p, j, sh, w0, w1; // bit position; word index, shift and the two stream words covering it
    uint32_t toff, z;          // table offset (2*component + (z != 0)) * LUT_SIZE, zig-zag index
    uint32_t n;                // slots since entry / since the last boundary crossed
    int32_t seg;               // last boundary crossed, -1 if none
    uint32_t k, segend;        // next boundary: segment index and its start bit
    uint32_t slot;             // absolute slot (final pass only)
    uint32_t st;               // ST_* bits (final pass only)
    uint32_t nrec;             // records emitted (relay passes)
    uint32_t q, qj;            // relay passes: running record position; its value at entry / after the last boundary
    uint32_t flags;            // relay passes: CZ_* annotations
    long long dcs;             // relay passes (DCS): packed sums of the DC differences decoded since entry / the last boundary
};

// ---- DC sums (K2 as a device-wide scan at subsequence granularity) ----------------------------------------
// DC prediction (MCU.cpp:107-108) needs, for every block, the sum of the DC differences of its component since the
// last restart / image start.  The relay passes add up the DC differences each subsequence decodes, the offset scan
// turns them into the predictor values at every subsequence entry, and a K3 strip adds the few DC differences between
// the entry of the subsequence its first slot lies in and that slot.  The three component sums travel as ONE exact
// 64-bit integer s0 + s1 * 2^15 + s2 * 2^30: adding two such numbers adds the components, and a component is
// recovered as long as |s_c| < 2^14 (a sum of differences is a difference of two DC values: |.| <= 4094 in any
// valid baseline stream; garbage decodes give garbage sums, which nothing trusts).
constexpr uint32_t DCS_SHIFT = 15;
KPEG_HD uint32_t dcs_weight(uint32_t c) { return 1u << (DCS_SHIFT * c); }
KPEG_HD void dcs_unpack(long long s, int (&v)[3])
{
    for (int c = 0; c < 3; ++c) {
        const int f = (int)((uint32_t)s & 0x7FFFu);
        v[c] = f >= 0x4000 ? f - 0x8000 : f;
        s = (s - v[c]) >> DCS_SHIFT; // exact: the difference is a multiple of 2^15
    }
}

template <class Words>
KPEG_HD void dec_init(DecState &d, const Words &W, const StreamView &S, uint32_t p, uint32_t c, uint32_t z, uint32_t k,
                      uint32_t slot)
{
    while (S.seg_bit[k] <= p)
        ++k;
    d.k = k;
    d.segend = S.seg_bit[k];
    d.p = p;
    d.j = p >> 5;
    d.sh = p & 31u;
    d.w0 = W(d.j);
    d.w1 = W(d.j + 1u);
    d.toff = (c * 2u + (z != 0u ? 1u : 0u)) * (uint32_t)LUT_SIZE;
    d.z = z;
    d.n = 0;
    d.seg = -1;
    d.slot = slot;
    d.st = 0;
    d.nrec = 0;
    d.q = d.qj = z;
    d.flags = 0;
    d.dcs = 0;
}

KPEG_HD SubState dec_exit_state(const DecState &d)
{
    SubState out;
    out.p = d.p;
    out.n = d.n;
    out.cz = ((d.toff >> (LUT_BITS + 1)) << 8) | d.z;
    out.seg = d.seg;
    return out;
}

// exit state of a relay pass (relay_run): the position runs on across blocks, the zig-zag index is its low six bits
KPEG_HD SubState relay_exit_state(const DecState &d)
{
    SubState out;
    out.p = d.p;
    out.n = d.q - d.qj;
    out.cz = ((d.toff >> (LUT_BITS + 1)) << 8) | (d.q & 63u) | d.flags | (d.q << CZ_POS_SHIFT);
    out.seg = d.seg;
    return out;
}

// ---- relay passes (cold / relay rounds) ----------------------------------------------------------------
// Decode from the state in `d` until the first symbol boundary at or after `end_bit`; produces the exit state, the
// slot count and -- EMIT -- one record per value-carrying symbol.
//
// Loop state: the bit position p alone.  Every iteration loads the two stream words that cover the symbol (W.pair:
// words p / 32 and p / 32 + 1 from the thread's own shared-memory region, one multiply-add for the address) and takes
// the 32-bit window with one funnel shift whose count is p itself (taken modulo 32).  A symbol is at most 27 bits.
// (Keeping the words in registers and rotating a look-ahead word in cost ten instructions per symbol; the two loads
// cost two, and the kernels are bound by instruction issue, not by the latency of the chain.)
//
// Table offset toff = (2*component + (z != 0)) * LUT_SIZE: a DC symbol switches to the component's AC table
// (toff |= LUT_SIZE), the end of a block to the next table in the ring.  q is the running record position
// (block * 64 + zig-zag index, counted from the first slot of the entry block); the zig-zag index is q & 63.
//
// The end of a block (next table of the ring, position rounded up to the next block) is a chain of selects by default;
// KPEG_BLOCK_END_BRANCH=1 builds it as a branch (measured at 4K q95: two dozen symbols per block, so 70 % of a warp's
// iterations have a lane that ends a block and run the branch body for that one lane -- no gain).
//
// Errors of the decode are annotations of the exit state (CZ_*), not device status bits: a speculative decode from a
// wrong entry state meets "errors" that mean nothing; only the annotations of the LAST decode of a subsequence -- the
// one from its true entry state -- survive in state[], and the offset scan turns those into status bits.
//
// DCS: also add up the DC differences per component (dcs_weight / dcs_unpack above).  wd is the weight the NEXT
// symbol's value is added with: 2^(15 component) when that symbol is a DC difference, 0 otherwise -- so the sum costs
// one multiply-add per symbol, no test of what kind of symbol it was.  (biased_extend gives exactly COEF_BIAS for a
// symbol without magnitude bits, i.e. the value 0.)
template <bool EMIT, bool DCS = false, class Words, class Luts, class Rec>
KPEG_HD void relay_run(DecState &d, const Words &W, const Luts &L, const StreamView &S, const JobGeom &g, uint32_t end_bit,
                       const Rec &rec)
{
    const uint32_t ring_last = (g.ncomp * 2u - 1u) * (uint32_t)LUT_SIZE; // the last table of the ring: the last component's AC table
    uint32_t wc = dcs_weight(d.toff >> (LUT_BITS + 1));                    // weight of the current component
    uint32_t wd = (d.toff & (uint32_t)LUT_SIZE) ? 0u : wc;                 // ... if the next symbol is its DC difference
    unsigned long long dcs = (unsigned long long)d.dcs, wsum = wd;         // biased sum; the weights that go with it
    uint32_t p = d.p, toff = d.toff, q = d.q, qj = d.qj;
    uint32_t k = d.k, segend = d.segend, nrec = d.nrec, flags = d.flags;
    int32_t seg = d.seg;
    // one running maximum finds both kinds of bad symbol: a value-carrying symbol must end inside its block
    // (zig-zag index + advance <= 64), a symbol without a value must have a real advance (1, 16 or 64; "no code
    // matches" is the entry with the impossible advance 127).  Only a symbol that reaches the end of its block can be
    // either, so the maximum is taken there.
    uint32_t worst = 0;
    uint32_t blk_end = (q | 63u) + 1u; // position of the next block's slot 0
    auto symbol = [&](uint32_t win, uint32_t e) { // everything a symbol does except moving the bit position
        const uint32_t T = e & 31u, adv = e >> 9, size = (e >> 5) & 15u;
        const uint32_t qn = q + adv;
        if (EMIT || DCS) {
            const uint32_t bv = biased_extend(win, T, size);
            if (EMIT) {
                // no branch: the record word is computed for every symbol (nearly all of them carry a value) and the
                // store alone is predicated
                rec.emit(size != 0u, nrec, record_pack(qn - 1u, bv));
                nrec += size != 0u ? 1u : 0u;
            }
#if KPEG_BLOCK_END_BRANCH
            if (DCS) { // the biased value; COEF_BIAS times the weights used comes off at the end
                dcs += (unsigned long long)bv * wd;
                wd = 0u;
            }
#else
            if (DCS)
                dcs += (unsigned long long)((long long)(int32_t)(bv - COEF_BIAS) * (long long)(int32_t)wd);
#endif
        }
#if KPEG_BLOCK_END_BRANCH
        uint32_t tn = toff | (uint32_t)LUT_SIZE;
        q = qn;
        if (qn >= blk_end) { // end of the block (EOB, or a coefficient in slot 63 -- or beyond: garbage)
            if (EMIT)
                worst = max_u32(worst, size ? qn - (blk_end - 64u) : adv);
            const bool wrap = toff == ring_last;
            tn = wrap ? 0u : toff + (uint32_t)LUT_SIZE; // next table of the ring (a block ends in its AC table)
            q = blk_end;
            blk_end += 64u;
            if (DCS) {
                wc = wrap ? 1u : wc << DCS_SHIFT;
                wd = wc;
                wsum += wc;
            }
        }
        toff = tn;
#else
        // end of the block (EOB, or a coefficient in slot 63 -- or beyond: garbage): selects, not a branch -- with two
        // dozen symbols per block some lane of a warp ends a block in most iterations, and a divergent branch then runs
        // its body for that one lane (measured: 70 % of the iterations at 4K q95)
        const bool endb = qn >= blk_end;
        const bool wrap = toff == ring_last;
        if (EMIT)
            worst = max_u32(worst, size ? qn + 64u - blk_end : adv);
        q = endb ? blk_end : qn;
        blk_end += endb ? 64u : 0u;
        toff = endb ? (wrap ? 0u : toff + (uint32_t)LUT_SIZE) : (toff | (uint32_t)LUT_SIZE); // next table of the ring / the AC table
        if (DCS) {
            wc = endb ? (wrap ? 1u : wc << DCS_SHIFT) : wc;
            wd = endb ? wc : 0u;
        }
#endif
    };
    auto cross_boundary = [&]() { // onto boundary k: state (segend, component 0, DC), position and DC sums restart
        q = (q + 63u) & ~63u;
        if (seg >= 0 && q - qj != seg_slot_base(g, k) - seg_slot_base(g, (uint32_t)seg))
            flags |= CZ_SEG_MISMATCH; // the interval between the last boundary and this one
        qj = q;
        blk_end = q + 64u;
        seg = (int32_t)k;
        p = segend;
        toff = 0;
        if (DCS) { // predictors restart with the interval (T.81 F.2.1.3.1)
            dcs = 0;
            wc = wd = 1u;
            wsum = 1u;
        }
        ++k;
        segend = S.seg_bit[k];
    };
    // Fast lane.  When the next boundary lies at least a symbol beyond end_bit (always, without restart markers,
    // except next to an image end) no symbol of this call can straddle it, and with a word-aligned end_bit
    // "p < end_bit" is "p / 32 < end_bit / 32": the loop carries no boundary test.
    if (p < end_bit && (end_bit & 31u) == 0u && segend >= end_bit + 32u) {
        const uint32_t jend = end_bit >> 5;
        uint32_t j = p >> 5;
        do {
            uint32_t w0, w1;
            W.pair(j, w0, w1);
            const uint32_t win = funnel_left(w0, w1, p); // the count is taken modulo 32
            uint32_t e = L.fast(toff, win >> (32 - LUT_BITS));
            if ((e & 31u) == 0u)
                e = L.slow(toff, win, e);
            symbol(win, e);
            p += e & 31u;
            j = p >> 5;
        } while (j < jend);
    }
    while (p < end_bit) {
        uint32_t w0, w1;
        W.pair(p >> 5, w0, w1);
        const uint32_t win = funnel_left(w0, w1, p);
        uint32_t e = L.fast(toff, win >> (32 - LUT_BITS));
        if ((e & 31u) == 0u) // code longer than LUT_BITS (or no code at all)
            e = L.slow(toff, win, e);
        const uint32_t T = e & 31u;
        if (p + T > segend) {
            // The symbol would straddle a restart / image boundary: we are in its padding (T.81 F.2.2.4 / E.2.4).
            // A valid stream is at the end of an MCU here, so the position is already a multiple of 64.
            cross_boundary();
            if (p >= S.total_bits)
                break;
            continue;
        }
        symbol(win, e);
        p += T;
    }
    // An interval without padding bits can end exactly where the subsequence stops: the boundary is crossed here, not
    // skipped by the next subsequence's dec_init -- every restart of the predictors is then visible to the scans as a
    // boundary some subsequence crossed.
    if (p == segend && p < S.total_bits)
        cross_boundary();
    if (EMIT)
        flags |= worst == ENTRY_ADV_INVALID ? CZ_BAD_CODE : (worst > 64u ? CZ_SLOT_OVERFLOW : 0u);
    d.p = p;
    d.j = p >> 5;
    d.sh = p & 31u;
    d.toff = toff;
    d.z = q & 63u;
    d.q = q;
    d.qj = qj;
    d.n = q - qj;
    d.k = k;
    d.segend = segend;
    d.seg = seg;
    d.nrec = nrec;
    d.flags = flags;
#if KPEG_BLOCK_END_BRANCH
    // the weights actually used: a DC difference that is still to come (wd pending) has not been added
    d.dcs = (long long)(dcs - (unsigned long long)COEF_BIAS * (wsum - wd));
#else
    (void)wsum;
    d.dcs = (long long)dcs;
#endif
}

// ---- Huffman final pass (fallback when records cannot be used) ---------------------------------------
// Decode until the first symbol boundary at or after `end_bit` or until the next symbol would belong to a block at or
// beyond `slot_limit` (a multiple of 64), so that a CTA can assemble its output in shared-memory windows; call again
// with a larger limit to resume.  d.slot is the absolute coefficient slot; values go to sink.put().
template <class Words, class Luts, class Sink>
KPEG_HD void write_run(DecState &d, const Words &W, const Luts &L, const StreamView &S, const JobGeom &g, uint32_t end_bit,
                       uint32_t slot_limit, const Sink &sink)
{
    const uint32_t total_slots = g.total_blocks * 64u;
    const uint32_t ring = g.ncomp * 2u * (uint32_t)LUT_SIZE;
    uint32_t p = d.p, j = d.j, sh = d.sh, w0 = d.w0, w1 = d.w1, toff = d.toff, z = d.z, n = d.n;
    uint32_t k = d.k, segend = d.segend, slot = d.slot, st = d.st;
    int32_t seg = d.seg;
    while (p < end_bit && slot < slot_limit) {
        const uint32_t nxt = W(j + 2u);
        const uint32_t win = funnel_left(w0, w1, sh);
        uint32_t e = L.fast(toff, win >> (32 - LUT_BITS));
        if ((e & 31u) == 0u)
            e = L.slow(toff, win, e);
        const uint32_t T = e & 31u;
        if (p + T > segend) {
            if (slot != seg_slot_base(g, k) && (k < g.nseg || slot < total_slots))
                st |= ST_SEG_MISMATCH;
            p = segend;
            z = 0;
            toff = 0;
            n = 0;
            seg = (int32_t)k;
            slot = seg_slot_base(g, k);
            ++k;
            segend = S.seg_bit[k];
            if (p >= S.total_bits)
                break;
            j = p >> 5;
            sh = p & 31u;
            w0 = W(j);
            w1 = W(j + 1u);
            continue;
        }
        const uint32_t adv = e >> 9;
        const uint32_t size = (e >> 5) & 15u;
        // status bits without branches: an invalid pattern has T == 17 and size == 0
        st |= ((T - size) >> 4) & ((T - size) & 1u);                 // ST_BAD_CODE (== 1): code "length" 17
        st |= (size != 0u && z + adv > 64u) ? ST_SLOT_OVERFLOW : 0u; // a coefficient beyond the block (EOB's 64 is clamped)
        // one store site for DC differences and AC coefficients
        const uint32_t raw = (win << (T - size)) >> ((32u - size) & 31u);
        const int32_t val = extend_value(raw, size | (size == 0u ? 1u : 0u));
        sink.put(z == 0u, slot, adv, size != 0u && slot + adv <= total_slots, val);
        p += T;
        sh += T;
        const bool cross = sh >= 32u;
        sh = cross ? sh - 32u : sh;
        j = cross ? j + 1u : j;
        w0 = cross ? w1 : w0;
        w1 = cross ? nxt : w1;
        uint32_t zn = z + adv;
        zn = zn > 64u ? 64u : zn;
        n += zn - z;
        slot += zn - z;
        if (zn == 64u) {
            z = 0;
            toff += (uint32_t)LUT_SIZE;
            toff = toff == ring ? 0u : toff;
        } else {
            z = zn;
            toff |= (uint32_t)LUT_SIZE;
        }
    }
    d.p = p;
    d.j = j;
    d.sh = sh;
    d.w0 = w0;
    d.w1 = w1;
    d.toff = toff;
    d.z = z;
    d.n = n;
    d.k = k;
    d.segend = segend;
    d.slot = slot;
    d.st = st;
    d.seg = seg;
}

// One-shot conveniences over a whole subsequence (CPU single-stepper, cold pass).
template <class Words, class Luts>
KPEG_HD SubState relay_span(const Words &W, const Luts &L, const StreamView &S, const JobGeom &g, uint32_t end_bit, uint32_t p,
                            uint32_t c, uint32_t z, uint32_t k)
{
    DecState d;
    dec_init(d, W, S, p, c, z, k, 0u);
    relay_run<false>(d, W, L, S, g, end_bit, NoRecorder{});
    return relay_exit_state(d);
}

// coefficients straight to global memory
template <class Words, class Luts>
KPEG_HD SubState write_span(const Words &W, const Luts &L, const StreamView &S, const JobGeom &g, uint32_t end_bit, uint32_t p,
                            uint32_t c, uint32_t z, uint32_t k, uint32_t slot, int16_t *coef, int16_t *dcdiff,
                            uint32_t *status_accum)
{
    DecState d;
    dec_init(d, W, S, p, c, z, k, slot);
    const GlobalSink sink{coef, dcdiff};
    write_run(d, W, L, S, g, end_bit, 0xFFFFFFFFu, sink);
    if (d.st)
        *status_accum |= d.st;
    return dec_exit_state(d);
}

} // namespace kpeg
#endif
