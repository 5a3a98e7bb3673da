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

KPEG_HD uint32_t funnel_left(uint32_t hi, uint32_t lo, uint32_t s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
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

// The same for n in 0..15 where the result for n == 0 is never used (a symbol without a value): no fix-up of n.
KPEG_HD int32_t extend_value_or_any(uint32_t v, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(v, 0u, n - 1u) ? (int32_t)v : (int32_t)v - (int32_t)((1u << n) - 1u); // clamped shift: defined for n == 0
#else
    return extend_value(v, n | (n == 0u ? 1u : 0u));
#endif
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

// ---- symbol records ------------------------------------------------------------------------------
// The relay passes decode every subsequence from (what converges to) its true entry state anyway;
// they also write down what they decoded, one 32-bit record per symbol, so that the final pass is a
// cheap, load-latency-tolerant EXPANSION of records instead of a third serial Huffman decode:
//   a symbol     : the table entry shifted left by 11 (total bits 11-15, magnitude size 16-19, slot
//                  advance 20-26) + the raw magnitude bits as they sit in the stream in bits 0-10 (at most
//                  11 of them in baseline JPEG; EXTEND is applied by the expander).  An invalid bit
//                  pattern is the entry ENTRY_INVALID: advance 127.
//   boundary jump: 0xF0000000 | index of the segment that starts there.
constexpr uint32_t REC_JUMP = 0xF0000000u;
constexpr uint32_t REC_RAW_MASK = 0x7FFu;

// raw = the `size` magnitude bits that follow the code in the left-aligned window (0 when size == 0)
KPEG_HD uint32_t record_raw_bits(uint32_t win, uint32_t T, uint32_t size)
{
    const uint32_t x = win << (T - size);
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(x, 0u, 32u - size); // clamped shift: size == 0 gives 0
#else
    return size ? x >> (32u - size) : 0u;
#endif
}

KPEG_HD uint32_t pack_record(uint32_t e, uint32_t raw) { return (e << 11) + raw; }

struct NoRecorder {
    KPEG_HD void emit(uint32_t, uint32_t) const {}
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
    uint32_t nrec;             // symbol records emitted (relay passes)
};

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

// Decode until the first symbol boundary at or after `end_bit` -- or, in the final pass, until the
// next symbol would belong to a block at or beyond `slot_limit` (a multiple of 64), so that a CTA can
// assemble its output in shared-memory windows; call again with a larger limit to resume.
//   WRITE == false : speculative / relay pass, only the exit state and slot count are produced.
//   WRITE == true  : final pass; d.slot is the absolute coefficient slot, AC coefficients go to
//                    sink.ac(absolute slot, value), DC differences to sink.dc(block, value).
//
// Loop state: bit position p with the words j, j+1 of the stream in registers, sh = p & 31 (a symbol
// is at most 27 bits, so two words always cover it); word j+2 is fetched unconditionally at the top
// of every iteration and rotated in, branch-free, when sh crosses 32 -- lanes of a warp cross word
// boundaries at different symbols, a conditional refill would be executed (mostly masked) by every
// warp on almost every iteration.  Table offset toff = (2*component + (z != 0)) * LUT_SIZE: a DC
// symbol switches to the component's AC table (toff |= LUT_SIZE), the end of a block to the next
// table in the ring.  z is the zig-zag index.
template <bool WRITE, bool EMIT, class Words, class Luts, class Sink, class Rec>
KPEG_HD void decode_run(DecState &d, const Words &W, const Luts &L, const StreamView &S, const JobGeom &g,
                        uint32_t end_bit, uint32_t slot_limit, const Sink &sink, const Rec &rec)
{
    const uint32_t total_slots = g.total_blocks * 64u;
    const uint32_t ring = g.ncomp * 2u * (uint32_t)LUT_SIZE;
    uint32_t p = d.p, j = d.j, sh = d.sh, w0 = d.w0, w1 = d.w1, toff = d.toff, z = d.z, n = d.n;
    uint32_t k = d.k, segend = d.segend, slot = d.slot, st = d.st, nrec = d.nrec;
    int32_t seg = d.seg;
    // Fast lane of the speculative / relay passes.  When the next boundary lies at least a symbol beyond
    // end_bit (always, without restart markers, except next to an image end) no symbol of this call can
    // straddle it, and with a word-aligned end_bit "p < end_bit" is "j < end_bit / 32": the loop carries
    // neither p nor the slot count (64 * blocks completed + z_exit - z_entry, added afterwards).
    if (!WRITE && p < end_bit && (end_bit & 31u) == 0u && segend >= end_bit + 32u) {
        const uint32_t jend = end_bit >> 5, z_entry = z;
        uint32_t nblk = 0;
        while (j < jend) {
            const uint32_t nxt = W(j + 2u);
            const uint32_t win = funnel_left(w0, w1, sh);
            uint32_t e = L.fast(toff, win >> (32 - LUT_BITS));
            if ((e & 31u) == 0u)
                e = L.slow(toff, win, e);
            const uint32_t T = e & 31u, adv = e >> 9;
            if (EMIT) {
                rec.emit(nrec, pack_record(e, record_raw_bits(win, T, (e >> 5) & 15u)));
                ++nrec;
            }
            sh += T;
            {
                const bool cross = sh >= 32u;
                sh = cross ? sh - 32u : sh;
                j = cross ? j + 1u : j;
                w0 = cross ? w1 : w0;
                w1 = cross ? nxt : w1;
            }
            const uint32_t zn = z + adv;
            if (zn >= 64u) {
                z = 0;
                ++nblk;
                toff += (uint32_t)LUT_SIZE;
                toff = toff == ring ? 0u : toff;
            } else {
                z = zn;
                toff |= (uint32_t)LUT_SIZE;
            }
        }
        p = (j << 5) + sh;
        n += nblk * 64u + z - z_entry;
    }
    while (p < end_bit && (!WRITE || slot < slot_limit)) {
        const uint32_t nxt = W(j + 2u); // consumed at the bottom of the iteration, if at all
        const uint32_t win = funnel_left(w0, w1, sh);
        uint32_t e = L.fast(toff, win >> (32 - LUT_BITS));
        if ((e & 31u) == 0u) // code longer than LUT_BITS (or no code at all)
            e = L.slow(toff, win, e);
        const uint32_t T = e & 31u;
        if (p + T > segend) {
            // The symbol would straddle a restart / image boundary: we are in its padding.
            if (WRITE && slot != seg_slot_base(g, k) && (k < g.nseg || slot < total_slots))
                st |= ST_SEG_MISMATCH;
            p = segend;
            z = 0;
            toff = 0;
            n = 0;
            seg = (int32_t)k;
            if (EMIT) {
                rec.emit(nrec, REC_JUMP | (k & 0xFFFFFFu));
                ++nrec;
                if (k > 0xFFFFFFu)
                    st |= ST_REC_OVERFLOW;
            }
            if (WRITE)
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
        if (WRITE) {
            const uint32_t size = (e >> 5) & 15u;
            // status bits without branches: an invalid pattern has T == 17 and size == 0
            st |= ((T - size) >> 4) & ((T - size) & 1u);      // ST_BAD_CODE (== 1): code "length" 17
            st |= (size != 0u && z + adv > 64u) ? ST_SLOT_OVERFLOW : 0u; // a coefficient beyond the block (EOB's 64 is clamped)
            // one store site for DC differences and AC coefficients
            const uint32_t raw = (win << (T - size)) >> ((32u - size) & 31u);
            const int32_t val = extend_value(raw, size | (size == 0u ? 1u : 0u));
            sink.put(z == 0u, slot, adv, size != 0u && slot + adv <= total_slots, val);
        }
        if (EMIT) {
            rec.emit(nrec, pack_record(e, record_raw_bits(win, T, (e >> 5) & 15u)));
            ++nrec;
        }
        p += T;
        sh += T;
        {
            const bool cross = sh >= 32u; // selects, not a branch
            sh = cross ? sh - 32u : sh;
            j = cross ? j + 1u : j;
            w0 = cross ? w1 : w0;
            w1 = cross ? nxt : w1;
        }
        uint32_t zn = z + adv;
        zn = zn > 64u ? 64u : zn;
        n += zn - z;
        if (WRITE)
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
    d.nrec = nrec;
}

// Final pass over records: state is (next record, absolute slot, zig-zag index).  Stops when the
// records are used up or the next symbol belongs to a block at or beyond slot_limit.
template <class RecAt, class Sink>
KPEG_HD void expand_run(uint32_t &k, uint32_t nrec, uint32_t &slot, uint32_t &z, uint32_t &st, const RecAt &rec_at,
                        const JobGeom &g, uint32_t slot_limit, const Sink &sink)
{
    const uint32_t total_slots = g.total_blocks * 64u;
    constexpr int BATCH = 8; // records fetched together: their loads are independent of the running state
    // software pipeline: the batch after the one being expanded is already in flight
    uint32_t nx[BATCH];
    uint32_t adv_max = 0, reach_max = 0;
#pragma unroll
    for (int j = 0; j < BATCH; ++j)
        nx[j] = (k < nrec && slot < slot_limit && k + (uint32_t)j < nrec) ? rec_at(k + (uint32_t)j) : 0u;
    while (k < nrec && slot < slot_limit) {
        uint32_t rr[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
            rr[j] = nx[j];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
            nx[j] = k + (uint32_t)(BATCH + j) < nrec ? rec_at(k + (uint32_t)(BATCH + j)) : 0u;
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            if (k < nrec && slot < slot_limit) { // a record left unused here is fetched again by the next call
                const uint32_t r = rr[j];
                ++k;
                if ((r >> 28) == 0xFu) {
                    const uint32_t seg = r & 0xFFFFFFu;
                    if (slot != seg_slot_base(g, seg) && (seg < g.nseg || slot < total_slots))
                        st |= ST_SEG_MISMATCH;
                    slot = seg_slot_base(g, seg);
                    z = 0;
                } else {
                    const uint32_t size = (r >> 16) & 15u, adv = (r >> 20) & 127u;
                    const bool has_value = size != 0u;
                    uint32_t zn = z + adv;
                    // error flags as two running maxima, folded into st after the loop: an invalid code has the
                    // (otherwise impossible) advance 127; a coefficient may not land beyond slot 63 of its block
                    adv_max = adv > adv_max ? adv : adv_max;
                    reach_max = (has_value && zn > reach_max) ? zn : reach_max;
                    const int32_t val = extend_value_or_any(r & REC_RAW_MASK, size);
                    // a sink that bounds its own writes (the shared-memory window) needs no check against the end of
                    // the coefficient buffer: what lands beyond it stays in the window and is never flushed
                    sink.put(z == 0u, slot, adv, Sink::bounds_itself ? has_value : (has_value && slot + adv <= total_slots), val);
                    zn = zn > 64u ? 64u : zn;
                    slot += zn - z;
                    z = zn == 64u ? 0u : zn;
                }
            }
        }
    }
    st |= adv_max == ENTRY_ADV_INVALID ? ST_BAD_CODE : 0u;
    st |= reach_max > 64u ? ST_SLOT_OVERFLOW : 0u;
}

// One-shot convenience: whole subsequence, coefficients straight to global memory.
template <bool WRITE, class Words, class Luts>
KPEG_HD SubState decode_span(const Words &W, const Luts &L, const StreamView &S, const JobGeom &g, uint32_t end_bit,
                             uint32_t p, uint32_t c, uint32_t z, uint32_t k, uint32_t slot, int16_t *coef,
                             int16_t *dcdiff, uint32_t *status_accum)
{
    DecState d;
    dec_init(d, W, S, p, c, z, k, slot);
    if (WRITE) {
        const GlobalSink sink{coef, dcdiff};
        decode_run<true, false>(d, W, L, S, g, end_bit, 0xFFFFFFFFu, sink, NoRecorder{});
        if (d.st)
            *status_accum |= d.st;
    } else {
        decode_run<false, false>(d, W, L, S, g, end_bit, 0xFFFFFFFFu, NullSink{}, NoRecorder{});
    }
    return dec_exit_state(d);
}

} // namespace kpeg
#endif
