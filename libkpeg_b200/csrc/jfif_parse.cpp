// jfif_parse.cpp -- host-side container parse of the drop-in (no CUDA).
//
// Produces the POD kpeg_plan the kernels consume.  It replaces the marker loop of
// JPEGDecoder::decodeImageFile and the parse* members (reference src/Decoder.cpp:53-75, 90-152,
// 164-530, 579-619) and is a T.81-correct superset of what they accept:
// segments are skipped by their length field, so APPn / COM / DRI are fine where the reference
// stops with "[ FATAL ] Invalid JFIF file"; SOF / SOS table selectors are honoured (the reference
// hard-wires Y -> 0, Cb/Cr -> 1, Decoder.cpp:704, MCU.cpp:110, which is what every file it decodes
// correctly uses); Nf == 1 is accepted.  Unsupported codings map to KPEG_ERR_UNSUPPORTED
// (reference: ResultCode::TERMINATE), malformed data to KPEG_ERR_FORMAT (ResultCode::ERROR).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "kpeg_cuda.h"

namespace {

inline unsigned be16(const uint8_t *p) { return (unsigned)((p[0] << 8) | p[1]); }

// End of the entropy-coded segment that starts at `s`: the first marker that is neither a stuffed
// FF00, an RSTn nor a fill byte.  scanImageData (Decoder.cpp:532-577) stops at the first FF D9 only; both
// agree on well-formed single-scan files, including ones with bytes (or whole images) after the first EOI.
size_t find_scan_end_range(const uint8_t *f, size_t n, size_t s, size_t stop)
{
    // the first position e in [s, stop) with f[e] == FF and f[e + 1] not in {00, D0..D7, FF}; `stop` if there is none.
    // Evaluating positions one by one is the same as the sequential walk that skips two bytes after FF 00 / FF Dx:
    // the byte skipped is never an FF itself.
    size_t e = s;
    while (e < stop && e + 1 < n) {
        const uint8_t *q = (const uint8_t *)memchr(f + e, 0xFF, std::min(stop, n - 1) - e);
        if (!q)
            return stop;
        e = (size_t)(q - f);
        const uint8_t b = f[e + 1];
        if (b == 0x00 || (b >= 0xD0 && b <= 0xD7)) {
            e += 2;
            continue;
        }
        if (b == 0xFF) {
            e += 1;
            continue;
        }
        return e;
    }
    return stop;
}

size_t find_scan_end(const uint8_t *f, size_t n, size_t s)
{
    // Always walk the markers.  Trusting a trailing FF D9 instead would send files that carry data after the first EOI
    // (MPO, concatenated JPEGs, appended previews) to the GPU whole; the reference stops at the first FF D9
    // (Decoder.cpp:546-557), and so does this.  The walk costs ~0.13 ms per MB -- 12 ms for the scan of a 16384x16384
    // image, as long as its decode -- so large scans are walked by several threads, each over its own stretch; the
    // first hit in file order wins.
    constexpr size_t PAR_MIN = (size_t)4 << 20;
    if (n - s < PAR_MIN) {
        const size_t e = find_scan_end_range(f, n, s, n);
        return e;
    }
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 2 ? 2 : (nt > 8 ? 8 : nt);
    std::vector<size_t> hit(nt, n);
    std::vector<std::thread> th;
    const size_t span = (n - s + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t a = s + t * span, b = std::min(n, a + span);
        if (a >= b)
            break;
        th.emplace_back([&, t, a, b] {
            const size_t e = find_scan_end_range(f, n, a, b);
            hit[t] = e < b ? e : n;
        });
    }
    for (auto &x : th)
        x.join();
    for (unsigned t = 0; t < nt; ++t)
        if (hit[t] < n)
            return hit[t];
    return n;
}

} // namespace

// The marker walk.  `pl` accumulates the frame and the tables in force; every SOS adds one kpeg_scan: a plan of its own
// (the scan's components in slots 0.., the tables and restart interval in force at that point -- T.81 B.2.3 allows DHT /
// DQT / DRI between scans) and its entropy-coded segment.  Accepted: ONE interleaved scan of all components in frame
// order (what the reference decodes, Decoder.cpp:461-530), or one scan PER component in any order (T.81 A.2.3: the MCU
// of a non-interleaved scan is one block; the reference reads such SOS headers and then mis-decodes).  Stops as soon as
// every component has its scan.
static int parse_impl(const uint8_t *f, size_t n, kpeg_plan *pl, kpeg_scan *scans, int max_scans, int *nscans)
{
    memset(pl, 0, sizeof *pl);
    *nscans = 0;
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8)
        return KPEG_ERR_FORMAT; // no SOI
    size_t i = 2;
    bool have_frame = false;
    uint8_t comp_id[3] = {0, 0, 0};
    bool comp_done[3] = {false, false, false};
    unsigned covered = 0;
    for (;;) {
        if (i + 2 > n)
            return KPEG_ERR_FORMAT;
        if (f[i] != 0xFF)
            return KPEG_ERR_FORMAT; // Decoder.cpp:126-132
        while (i + 1 < n && f[i + 1] == 0xFF)
            ++i; // fill bytes before a marker (T.81 B.1.1.2)
        if (i + 2 > n)
            return KPEG_ERR_FORMAT;
        const unsigned marker = f[i + 1];
        i += 2;
        if (marker == 0xD8 || marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7))
            continue; // stand-alone markers
        if (marker == 0xD9)
            return KPEG_ERR_FORMAT; // EOI before every component had its scan
        if (i + 2 > n)
            return KPEG_ERR_FORMAT;
        const unsigned L = be16(f + i);
        if (L < 2 || i + L > n)
            return KPEG_ERR_FORMAT;
        const uint8_t *p = f + i + 2;
        const unsigned plen = L - 2;
        switch (marker) {
        case 0xC0: { // SOF0 -- parseSOF0Segment, Decoder.cpp:301-364
            if (plen < 6 || have_frame)
                return KPEG_ERR_FORMAT;
            if (p[0] != 8)
                return KPEG_ERR_UNSUPPORTED;
            const unsigned H = be16(p + 1), W = be16(p + 3), nf = p[5];
            if (nf != 1 && nf != 3)
                return KPEG_ERR_UNSUPPORTED;
            if (plen < 6 + 3 * nf)
                return KPEG_ERR_FORMAT;
            if (W == 0 || H == 0)
                return KPEG_ERR_UNSUPPORTED; // DNL-defined height is not baseline-decodable here
            for (unsigned c = 0; c < nf; ++c) {
                comp_id[c] = p[6 + 3 * c];
                // "Chroma subsampling not yet supported!" (Decoder.cpp:351-356).  A single-component
                // frame is always decoded 1x1 whatever H,V it declares (T.81 A.2.2).
                if (nf == 3 && p[7 + 3 * c] != 0x11)
                    return KPEG_ERR_UNSUPPORTED;
                if (p[8 + 3 * c] > 3)
                    return KPEG_ERR_FORMAT;
                pl->comp_tq[c] = p[8 + 3 * c];
            }
            pl->width = (uint16_t)W;
            pl->height = (uint16_t)H;
            pl->ncomp = (uint8_t)nf;
            have_frame = true;
            break;
        }
        case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7:
        case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
            return KPEG_ERR_UNSUPPORTED; // Decoder.cpp:65-66 (SOF1, SOF2) and the other non-baseline frames
        case 0xDB: { // DQT -- parseQuantizationTable, Decoder.cpp:230-299 (8-bit tables, zig-zag order)
            unsigned k = 0;
            while (k < plen) {
                const unsigned pq = p[k] >> 4, tq = p[k] & 15u;
                if (tq > 3)
                    return KPEG_ERR_FORMAT;
                if (pq != 0)
                    return KPEG_ERR_UNSUPPORTED; // 16-bit tables are not baseline
                if (k + 65 > plen)
                    return KPEG_ERR_FORMAT;
                for (int q = 0; q < 64; ++q)
                    pl->qt[tq][q] = p[k + 1 + q];
                pl->qt_present[tq] = 1;
                k += 65;
            }
            break;
        }
        case 0xC4: { // DHT -- parseHuffmanTable, Decoder.cpp:366-459
            unsigned k = 0;
            while (k < plen) {
                if (k + 17 > plen)
                    return KPEG_ERR_FORMAT;
                const unsigned tc = p[k] >> 4, th = p[k] & 15u;
                if (tc > 1 || th > 3)
                    return KPEG_ERR_FORMAT;
                kpeg_huff_spec *h = &pl->ht[tc][th];
                memset(h, 0, sizeof *h); // a redefinition replaces the table (the reference appends)
                unsigned total = 0;
                for (int q = 0; q < 16; ++q) {
                    h->counts[q] = p[k + 1 + q];
                    total += h->counts[q];
                }
                if (total > 256 || k + 17 + total > plen)
                    return KPEG_ERR_FORMAT;
                memcpy(h->symbols, p + k + 17, total);
                pl->ht_present[tc][th] = 1;
                k += 17 + total;
            }
            break;
        }
        case 0xDD: // DRI -- no counterpart in the reference; T.81 B.2.4.4
            if (plen < 2)
                return KPEG_ERR_FORMAT;
            pl->restart_interval = (uint16_t)be16(p);
            break;
        case 0xDA: { // SOS -- parseSOSSegment, Decoder.cpp:461-530
            if (!have_frame || plen < 1)
                return KPEG_ERR_FORMAT;
            const unsigned ns = p[0];
            if (ns != pl->ncomp && ns != 1)
                return KPEG_ERR_UNSUPPORTED; // two-component scans
            if (plen < 1 + 2 * ns + 3)
                return KPEG_ERR_FORMAT;
            if (*nscans >= max_scans)
                return KPEG_ERR_UNSUPPORTED; // a multi-scan file handed to the single-scan entry point
            kpeg_scan *sc = &scans[*nscans];
            memset(sc, 0, sizeof *sc);
            sc->plan = *pl; // dimensions, tables and restart interval as they stand
            sc->plan.ncomp = (uint8_t)ns;
            for (unsigned s = 0; s < ns; ++s) {
                unsigned c = 0;
                while (c < pl->ncomp && comp_id[c] != p[1 + 2 * s])
                    ++c;
                if (c == pl->ncomp)
                    return KPEG_ERR_FORMAT; // no such component in the frame
                if (ns > 1 && c != s)
                    return KPEG_ERR_UNSUPPORTED; // an interleaved scan must list the components in frame order
                if (comp_done[c])
                    return KPEG_ERR_UNSUPPORTED; // a component coded twice: not sequential baseline
                const unsigned td = p[2 + 2 * s] >> 4, ta = p[2 + 2 * s] & 15u;
                if (td > 3 || ta > 3)
                    return KPEG_ERR_FORMAT;
                if (!pl->qt_present[pl->comp_tq[c]] || !pl->ht_present[0][td] || !pl->ht_present[1][ta])
                    return KPEG_ERR_FORMAT;
                comp_done[c] = true;
                sc->comp[s] = (uint8_t)c;
                sc->plan.comp_tq[s] = pl->comp_tq[c];
                sc->plan.comp_td[s] = (uint8_t)td;
                sc->plan.comp_ta[s] = (uint8_t)ta;
                pl->comp_td[c] = (uint8_t)td;
                pl->comp_ta[c] = (uint8_t)ta;
            }
            for (unsigned s = ns; s < 3; ++s)
                sc->plan.comp_tq[s] = sc->plan.comp_td[s] = sc->plan.comp_ta[s] = 0;
            const size_t s0 = i + L;
            const size_t e = find_scan_end(f, n, s0);
            sc->off = s0;
            sc->len = e - s0;
            ++*nscans;
            covered += ns;
            if (covered == pl->ncomp)
                return KPEG_OK;
            if (e + 1 >= n)
                return KPEG_ERR_FORMAT; // the file ends before the other components' scans
            i = e; // on the marker that ended the segment
            continue;
        }
        default: // APPn, COM, DNL, ...: skipped by length
            break;
        }
        i += L;
    }
}

extern "C" int kpeg_parse_jfif(const uint8_t *f, size_t n, kpeg_plan *pl, size_t *scan_off, size_t *scan_len)
{
    if (!f || !pl || !scan_off || !scan_len)
        return KPEG_ERR_ARG;
    kpeg_scan sc;
    int ns = 0;
    const int rc = parse_impl(f, n, pl, &sc, 1, &ns);
    if (rc != KPEG_OK)
        return rc;
    *pl = sc.plan; // one interleaved scan: its plan is the frame's
    *scan_off = sc.off;
    *scan_len = sc.len;
    return KPEG_OK;
}

extern "C" int kpeg_parse_jfif_scans(const uint8_t *f, size_t n, kpeg_plan *frame, kpeg_scan *scans, int max_scans, int *nscans)
{
    if (!f || !frame || !scans || max_scans < 1 || !nscans)
        return KPEG_ERR_ARG;
    const int rc = parse_impl(f, n, frame, scans, max_scans, nscans);
    if (rc != KPEG_OK || *nscans == 1)
        return rc;
    // one scan per component.  The quantiser of a component is the table in force when ITS scan began (T.81 B.2.3): the
    // frame plan gets them in slots 0..2, component c using slot c, whatever ids the file used.
    kpeg_plan out = *frame;
    memset(out.qt_present, 0, sizeof out.qt_present);
    for (int k = 0; k < *nscans; ++k) {
        const unsigned c = scans[k].comp[0];
        memcpy(out.qt[c], scans[k].plan.qt[scans[k].plan.comp_tq[0]], sizeof out.qt[c]);
        out.qt_present[c] = 1;
        out.comp_tq[c] = (uint8_t)c;
    }
    out.restart_interval = 0; // per scan (kpeg_scan::plan)
    *frame = out;
    return KPEG_OK;
}

// Cuts one restart-marked entropy-coded segment into `parts` bands of whole MCU rows, so that each
// band is itself a complete image of the same width (multi-GPU tiling of very large images,
// BASELINE.json configs[4]; the reference has no tiling at all).  Requires a restart interval that
// is a whole number of MCU rows or divides one (every band then starts right after a marker).
//   out_begin[parts], out_end[parts] : band b is scan[out_begin[b], out_end[b]) -- the RSTn marker
//                                      that separates two bands belongs to neither
//   out_row[parts + 1]               : first MCU row of each band (out_row[parts] = total MCU rows)
// Surplus parts (more parts than MCU rows) come back as empty bands at the end.
extern "C" int kpeg_split_restart_bands(const uint8_t *scan, size_t len, const kpeg_plan *plan, int parts,
                                        uint64_t *out_begin, uint64_t *out_end, uint32_t *out_row)
{
    if (!scan || !plan || parts < 1 || !out_begin || !out_end || !out_row)
        return KPEG_ERR_ARG;
    const uint32_t mx = (plan->width + 7u) / 8u, my = (plan->height + 7u) / 8u;
    const uint32_t ri = plan->restart_interval;
    if (ri == 0 || !((ri % mx) == 0 || (mx % ri) == 0))
        return KPEG_ERR_UNSUPPORTED; // bands must begin on a restart marker
    const uint32_t row_step = (ri % mx) == 0 ? ri / mx : 1u; // MCU rows between candidate cut points
    // balance in CUT UNITS (groups of row_step MCU rows), not in rows: rounding a balanced row down to a cut point
    // can give two bands the same first row when row_step > 1 and parts is close to the number of units
    const uint32_t units = (my + row_step - 1) / row_step;
    const uint32_t used = (uint32_t)parts < units ? (uint32_t)parts : units;
    for (uint32_t k = 0; k <= used; ++k) { // strictly increasing: used <= units
        const uint64_t r = ((uint64_t)units * k / used) * row_step;
        out_row[k] = r < my ? (uint32_t)r : my;
    }
    for (uint32_t k = used + 1; k <= (uint32_t)parts; ++k)
        out_row[k] = my;
    for (int k = 0; k < parts; ++k)
        out_begin[k] = out_end[k] = len;
    out_begin[0] = 0;
    // walk the markers: marker number j (1-based) precedes restart interval j
    uint32_t b = 1;
    uint64_t markers = 0;
    const uint8_t *p = scan, *end = scan + len;
    while (b < used && p + 1 < end) {
        const uint8_t *q = (const uint8_t *)memchr(p, 0xFF, (size_t)(end - 1 - p));
        if (!q)
            break;
        const uint8_t m = q[1];
        if (m >= 0xD0 && m <= 0xD7) {
            ++markers;
            if (((uint64_t)out_row[b] * mx) / ri == markers) {
                out_end[b - 1] = (uint64_t)(q - scan);
                out_begin[b] = (uint64_t)(q + 2 - scan);
                ++b;
            }
            p = q + 2;
        } else {
            p = q + (m == 0xFF ? 1 : 2);
        }
    }
    if (b < used)
        return KPEG_ERR_STREAM; // fewer restart markers than the DRI interval promises
    out_end[used - 1] = len;
    return KPEG_OK;
}

extern "C" int kpeg_ppm_header(int width, int height, char *buf, size_t cap)
{
    // Image::dumpRawData, Image.cpp:124-127
    return snprintf(buf, cap,
                    "P6\n# PPM dump created using libKPEG: https://github.com/TheIllusionistMirage/libKPEG\n%d %d\n255\n",
                    width, height);
}
